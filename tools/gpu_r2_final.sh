#!/bin/bash
# Round-2 closing pass on one GPU: full GPU suite, smoke, the driver-style bench line + reference arm, sanitizers, launch list, ncu captures.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r2f_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_reference.json 2> gpurun_out/r2f_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2f_bench.json'))
r = json.load(open('gpurun_out/r2f_reference.json'))
print('coif4', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'ref', round(r['value']), 'e2e/ref', round(d['e2e']['value'] / r['value'], 1))
for k, v in d.get('workloads', {}).items():
    print(k, round(v['value']), 'frac', round(v['roofline']['frac'], 3), {a: v[a] for a in v if a in ('ms_per_step', 'ms_job')}, 'e2e', (v.get('e2e') or {}).get('value'))
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize.py > gpurun_out/r2f_sanitizer_$tool.log 2>&1; echo "sanitizer $tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize: OK" gpurun_out/r2f_sanitizer_$tool.log | head -4
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2f_ncu_launch.log 2>&1; echo "ncu launches exit $?"
for spec in "sym5 wpt_frame r2f_sym5" "coif4 wpt_frame r2f_coif4" "haar haar_linear r2f_haar" "stft stft_tc511 r2f_stft"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/prof_$3 \
      python bench.py --workload $1 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-workloads > gpurun_out/ncu_$3.log 2>&1
  echo "ncu $1 exit $?"
done
