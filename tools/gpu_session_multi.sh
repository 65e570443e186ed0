#!/bin/bash
# N-GPU session (gpurun --gpus N -- bash tools/gpu_session_multi.sh N): the real multi-rank band test, the bare host copy ceiling and the full
# bench line at N ranks.
# N-GPU session: real multi-rank band test, copy ceiling, full bench
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_shard.py -m gpu -q -x -k "two_rank" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/host_copy_ceiling.py > gpurun_out/ceiling_n$N.json 2> gpurun_out/ceiling_n$N.err
cat gpurun_out/ceiling_n$N.json; tail -2 gpurun_out/ceiling_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_full_n$N.json 2> gpurun_out/bench_full_n$N.err
tail -3 gpurun_out/bench_full_n$N.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_full_n$N.json').read().strip().splitlines()[-1])
print('value',j['value'],'e2e',j['e2e']['value'],j['e2e']['host_gbs_total'],'e2e_u16',j['e2e_u16']['value'],j['e2e_u16']['host_gbs_total'])
print('parity',{k:v for k,v in j['parity'].items() if k not in ('stream_frames','config1','tolerance')})
for k,v in j['extra'].items():
    if k=='sweep': print('sweep',[ (r['radius'],round(r['filter_mpixel_s']),round(r['frac_of_fp32_nominal'],3)) for r in v['rows']])
    else: print(k,v)
PY
