#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err
echo "bench wall seconds: $SECONDS"; tail -3 gpurun_out/bench_full_n1.err
python tools/bench_extra.py guided > gpurun_out/extra_guided.json 2>&1; cat gpurun_out/extra_guided.json
python tools/bench_extra.py buffer2d > gpurun_out/extra_buffer2d.json 2>&1
python tools/bench_extra.py single 2>&1 | tail -1
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_full_n1.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'e2e',j['e2e']['value'],'e2e_u16',j['e2e_u16']['value'],'cpu',j['cpu_baseline']['value'],j['cpu_baseline']['cores'])
print('roofline',{k:v for k,v in j['roofline'].items() if k in ('achieved','frac','frac_of_nominal','kernel_ms_per_launch','peak')}, j['roofline']['mufu'])
print('clocks',j['clocks'])
print('upsample',j['extra']['upsample']['ms'],'single',j['extra']['single_frame']['r7_us'],j['extra']['single_frame']['r2_us'])
print('bands',j['extra']['bands_nccl']['ms_per_frame'], j['extra']['bands_nccl']['oracle_seam_check'])
PY
