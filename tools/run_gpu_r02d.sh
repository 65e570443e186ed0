#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e > gpurun_out/bench_d.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:jbf_fast -s 6 -c 1 -o gpurun_out/prof_r02_jbf_r7 python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
