#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py tests/test_shard.py -m gpu -x -q 2>&1 | tail -5
python tools/bench_extra.py single > gpurun_out/single_auto.json 2>&1; cat gpurun_out/single_auto.json
KDME_NO_REFINE=1 python tools/bench_extra.py single 2>&1 | tail -1
KDME_TILE_H=4 python tools/bench_extra.py single 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e > gpurun_out/bench_b.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_single.csv python tools/bench_extra.py single > gpurun_out/ncu_single.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_single.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:40]: print(r[ki][:60], r[gi], r[vi])
PY
