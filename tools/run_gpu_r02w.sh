#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q -k "upsample" 2>&1 | tail -3
python tools/bench_extra.py upsample 2>&1 | tail -1
KDME_NO_REFINE=1 python tools/bench_extra.py upsample 2>&1 | tail -1
