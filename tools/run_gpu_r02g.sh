#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for th in 16 8 4; do echo TH=$th; KDME_TILE_H=$th python tools/bench_extra.py single 2>&1 | tail -1; done
echo nosplit; KDME_NO_SPLIT_TILES=1 python tools/bench_extra.py single 2>&1 | tail -1
echo auto; python tools/bench_extra.py single 2>&1 | tail -1
python tools/bench_extra.py guided 2>&1 | tail -1
KDME_GUIDED_GENERIC=1 python tools/bench_extra.py guided 2>&1 | tail -1
python tools/bench_extra.py upsample > gpurun_out/ups_plain.log 2>&1 && tail -1 gpurun_out/ups_plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"jbf|presmooth" -c 12 --csv --log-file gpurun_out/launches_upsample.csv python tools/bench_extra.py upsample > gpurun_out/ncu_ups.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_upsample.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:8]: print(r[ki][:70], r[gi], r[vi])
PY
