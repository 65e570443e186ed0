#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py tests/test_gpu_buffers_guided.py -m gpu -x -q 2>&1 | tail -4
python tools/bench_extra.py upsample 2>&1 | tail -1
KDME_UPSAMPLE_DENSE=1 python tools/bench_extra.py upsample 2>&1 | tail -1
python tools/bench_extra.py guided > gpurun_out/guided_plain.log 2>&1 && cat gpurun_out/guided_plain.log | tail -1 &&
ncu --set full --clock-control none --import-source on -k regex:guided_fill -s 2 -c 1 -o gpurun_out/prof_r02_guided_v0 python tools/bench_extra.py guided > gpurun_out/ncu_guided.log 2>&1
tail -2 gpurun_out/ncu_guided.log
