#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q 2>&1 | tail -2
/usr/bin/time -v python bench.py > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err
grep -E "Elapsed|Maximum resident" gpurun_out/bench_full_n1.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_n1.json 2>/dev/null; cat gpurun_out/bench_ref_n1.json | cut -c1-400
python tools/bench_extra.py guided > gpurun_out/extra_guided.json 2>&1; cat gpurun_out/extra_guided.json
python tools/bench_extra.py buffer2d > gpurun_out/extra_buffer2d.json 2>&1
python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:presmooth5 -s 4 -c 1 -o gpurun_out/prof_r02_presmooth5 python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_p.log 2>&1
tail -1 gpurun_out/ncu_p.log
