#!/bin/bash
# round-2 first GPU pass: parity tests, quick bench, single-frame tile-height comparison
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150 > gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -c 1500 gpurun_out/bench_a.json
for th in 16 8 4; do KDME_TILE_H=$th python tools/bench_extra.py single > gpurun_out/single_th$th.json 2>&1; cat gpurun_out/single_th$th.json; done
python tools/bench_extra.py single > gpurun_out/single_auto.json 2>&1; cat gpurun_out/single_auto.json
KDME_NO_REFINE=1 python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e > gpurun_out/bench_a_norefine.json 2>/dev/null
tail -c 600 gpurun_out/bench_a_norefine.json
