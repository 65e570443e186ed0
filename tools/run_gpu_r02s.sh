#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q -k "upsample_gather" 2>&1 | tail -3
python tools/parity_sweep.py --frames 24 --fuzz 150 > gpurun_out/parity_sweep.json 2> gpurun_out/parity_sweep.err
cat gpurun_out/parity_sweep.json; tail -3 gpurun_out/parity_sweep.err
