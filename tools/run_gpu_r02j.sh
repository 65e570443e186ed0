#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --frames 1024 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_j.json 2>/dev/null
python -c "
import json; j=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['kernel_ms_per_launch'])"
python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"jbf|presmooth" -s 12 -c 12 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ncu_b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_bench.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:10]: print(r[ki][:70], r[gi], r[vi])
PY
