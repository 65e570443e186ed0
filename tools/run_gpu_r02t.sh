#!/bin/bash
mkdir -p gpurun_out
for lib in "" "$PWD/kinectdepthmapenhancement_b200/libkdme_regs64.so"; do
  echo "LIB=$lib"
  KDME_LIB_PATH=$lib python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_t.json 2>/dev/null
  python -c "
import json; j=json.loads(open('gpurun_out/bench_t.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
  KDME_LIB_PATH=$lib python tools/bench_extra.py single 2>&1 | tail -1 | cut -c150-330
  KDME_LIB_PATH=$lib python tools/bench_extra.py sweep 2>&1 | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print([(r['radius'],round(r['filter_mpixel_s'])) for r in j['rows']])"
done
