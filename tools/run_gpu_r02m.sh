#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_extra.py single 2>&1 | tail -1
python tools/bench_extra.py upsample 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_m.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_m.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
KDME_NO_REFINE=1 python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_m2.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_m2.json').read().strip().splitlines()[-1]); print('bench norefine', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
