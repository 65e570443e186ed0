#!/bin/bash
# One-GPU evidence session (gpurun -- bash tools/gpu_session_single.sh [TAG]): parity tests, smoke, the full bench line, ncu launch lists
# (bench step / one Process call / one Upsampling call) and ONE ncu --set full capture each of the filter, pre-smooth, refinement and
# gather kernels -- every ncu command only after the same command ran clean without ncu.  Outputs land in gpurun_out/; copy what is
# cited into profiles/ (tools/ncu_summary.py condenses the .ncu-rep files; SKIP_TESTS=1 skips pytest and smoke).
TAG=${1:-v3}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
fi
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
python tools/bench_extra.py guided > gpurun_out/extra_guided.json 2>/dev/null; cut -c1-200 gpurun_out/extra_guided.json
B="python bench.py --steps 1 --warmup 3 --frames 128 --no-cpu-baseline --no-e2e --no-extra"
$B > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"presmooth|jbf_" -c 36 --csv --log-file gpurun_out/launches_bench.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:jbf_fast -s 6 -c 1 -f -o gpurun_out/prof_r02_jbf_r7_$TAG $B > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:presmooth5 -s 6 -c 1 -f -o gpurun_out/prof_r02_presmooth5_$TAG $B >> gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jbf_refine -s 6 -c 1 -f -o gpurun_out/prof_r02_refine_$TAG $B >> gpurun_out/ncu_full.log 2>&1
python tools/bench_extra.py upsample > gpurun_out/extra_upsample.json 2>/dev/null &&
ncu --set full --clock-control none --import-source on -k regex:upsample_gather -s 3 -c 1 -f -o gpurun_out/prof_r02_gather_$TAG python tools/bench_extra.py upsample >> gpurun_out/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"presmooth|jbf_" -s 12 -c 12 --csv --log-file gpurun_out/launches_upsample.csv python tools/bench_extra.py upsample > /dev/null 2>&1
python tools/bench_extra.py single > gpurun_out/extra_single.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"presmooth|jbf_" -s 30 -c 24 --csv --log-file gpurun_out/launches_single.csv python tools/bench_extra.py single > /dev/null 2>&1
tail -1 gpurun_out/ncu_full.log
# gpurun brings back at most 64 MiB: condense the reports here and keep only the filter kernel's .ncu-rep
for k in jbf_r7 presmooth5 refine gather; do
  python tools/ncu_summary.py gpurun_out/prof_r02_${k}_$TAG.ncu-rep > gpurun_out/prof_r02_${k}_$TAG.ncu.txt 2>/dev/null
  [ $k != jbf_r7 ] && rm -f gpurun_out/prof_r02_${k}_$TAG.ncu-rep
done
ls -la gpurun_out/*_$TAG.ncu*; du -sh gpurun_out
