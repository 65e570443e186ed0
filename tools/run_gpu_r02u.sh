#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q 2>&1 | tail -2
for rl in 0 3 4 5 6 8; do echo "RES_LIMIT=$rl"; KDME_RES_LIMIT=$rl python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-330; done
for th in 16 4; do for rl in 0 2 3 9 10; do echo "TH=$th RES_LIMIT=$rl"; KDME_TILE_H=$th KDME_RES_LIMIT=$rl python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-200; done; done
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_u.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
