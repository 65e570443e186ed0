#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jbf.py -m gpu -x -q 2>&1 | tail -2
python tools/bench_extra.py single > gpurun_out/single_plain.log 2>&1 && tail -1 gpurun_out/single_plain.log
KDME_NO_REFINE=1 python tools/bench_extra.py single 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"jbf|presmooth" -s 60 -c 6 --csv --log-file gpurun_out/launches_single.csv python tools/bench_extra.py single > gpurun_out/ncu_single.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_single.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:8]: print(r[ki][:70], r[gi], r[vi])
PY
