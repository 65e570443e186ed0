#!/bin/bash
# One-GPU evidence session (gpurun -- bash tools/gpu_session_single.sh): parity tests, smoke, the full bench line, then ONE ncu --set full
# capture of the dominant kernel after the same command ran clean.  Outputs land in gpurun_out/; copy what is cited into profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
python bench.py > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err; tail -2 gpurun_out/bench_full_n1.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_full_n1.json').read().strip().splitlines()[-1])
print('value',j['value'],'ms/step',j['ms_per_step'],'e2e',j['e2e']['value'],'e2e_u16',j['e2e_u16']['value'],'cpu',j['cpu_baseline']['value'],j['cpu_baseline']['cores'])
print('roofline',{k:v for k,v in j['roofline'].items() if k in ('achieved','frac','frac_of_nominal','kernel_ms_per_launch','peak')}, j['roofline']['mufu'])
print('parity',{k:v for k,v in j['parity'].items() if k not in ('stream_frames','config1','tolerance')})
print('upsample',j['extra']['upsample']['ms'],'single',j['extra']['single_frame']['r7_us'],j['extra']['single_frame']['r2_us'])
print('bands',j['extra']['bands_nccl']['ms_per_frame'], j['extra']['bands_nccl']['oracle_seam_check'])
print('sweep',[(r['radius'],round(r['filter_mpixel_s']),round(r['frac_of_fp32_nominal'],3)) for r in j['extra']['sweep']['rows']])
PY
python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:jbf_fast -s 6 -c 1 -o gpurun_out/prof_r02_jbf_r7_v2 python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
