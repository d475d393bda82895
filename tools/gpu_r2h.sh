#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r2h_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2h_bench.json'))
print('coif4', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), d['e2e'].get('bound'))
for k, v in d.get('workloads', {}).items():
    if 'error' in v:
        print(k, v); continue
    print(k, round(v['value']), 'frac', round(v['roofline']['frac'], 3), {a: v[a] for a in v if a in ('ms_per_step', 'ms_job', 'cuda_graph')}, 'e2e', v.get('e2e', {}).get('value'))
PY
