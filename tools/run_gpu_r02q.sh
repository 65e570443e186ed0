#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-420; done
echo NO_PDL; KDME_NO_PDL=1 python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-420
echo NO_REFINE; KDME_NO_REFINE=1 python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-420
echo NO_SPLIT; KDME_NO_SPLIT_TILES=1 python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-420
echo TH16; KDME_TILE_H=16 python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-420
