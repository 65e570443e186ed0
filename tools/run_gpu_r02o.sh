#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_extra.py single > gpurun_out/single_plain.log 2>&1 && tail -1 gpurun_out/single_plain.log
KDME_NO_REFINE=1 python tools/bench_extra.py single 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_o.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_o.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
ncu --set full --clock-control none --import-source on -k regex:jbf_fast -s 30 -c 1 -o gpurun_out/prof_r02_single_filter python tools/bench_extra.py single > gpurun_out/ncu_s.log 2>&1
tail -1 gpurun_out/ncu_s.log
