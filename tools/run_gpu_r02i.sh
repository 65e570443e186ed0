#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29610 tools/host_copy_ceiling.py > gpurun_out/ceiling_n4.json 2> gpurun_out/ceiling_n4.err
bash tools/run_gpu_r02h.sh 8 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
