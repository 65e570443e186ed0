#!/bin/bash
mkdir -p gpurun_out
python bench.py --frames 512 --steps 2 --warmup 3 --e2e-frames 256 --cpu-seconds 3 --band-size 4096 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
tail -5 gpurun_out/bench_e.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_e.json').read().strip().splitlines()[-1])
print('value',j['value'],'e2e',j['e2e']['value'],'e2e_u16',j['e2e_u16']['value'])
print('parity',{k:v for k,v in j['parity'].items() if k not in ('stream_frames','config1','tolerance')})
print('config1',j['parity']['config1'])
for k,v in j['extra'].items():
    if k=='sweep': print('sweep',[ (r['radius'],round(r['filter_mpixel_s']),round(r['frac_of_fp32_nominal'],3)) for r in v['rows']])
    else: print(k,v)
print('roofline',{k:v for k,v in j['roofline'].items() if k in ('achieved','frac','frac_of_nominal','mufu','kernel_ms_per_launch')})
PY
python tools/host_copy_ceiling.py > gpurun_out/ceiling_n1.json 2> gpurun_out/ceiling_n1.err; cat gpurun_out/ceiling_n1.json; tail -3 gpurun_out/ceiling_n1.err
