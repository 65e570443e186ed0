#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:jbf_refine -s 4 -c 1 -o gpurun_out/prof_r02_refine python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_r.log 2>&1
tail -1 gpurun_out/ncu_r.log
