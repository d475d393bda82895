#!/bin/bash
# Round-2 first pass: GPU tests, the full bench line (all workloads), reference arm, compute-sanitizer, launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2a_bench.json'))
    print('coif4', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), d['e2e'].get('bound'), d['build'])
    for k, v in d.get('workloads', {}).items():
        if 'error' in v:
            print(k, v); continue
        print(k, round(v['value']), 'frac', round(v['roofline']['frac'], 3), {a: v[a] for a in v if a in ('ms_per_step', 'ms_job', 'allreduce_us_by_rank', 'cuda_graph', 'with_dcnn_forward')}, 'e2e', v.get('e2e', {}).get('value'))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -5 gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_reference.json 2> gpurun_out/r2a_reference.err; echo "reference exit $?"; tail -c 700 gpurun_out/r2a_reference.json
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r2a_sanitizer_$tool.log 2>&1; echo "sanitizer $tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize:" gpurun_out/r2a_sanitizer_$tool.log | tail -20
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu_launch.log 2>&1; echo "ncu launches exit $?"
