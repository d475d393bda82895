#!/bin/bash
# quick GPU pass: wpt parity tests + the two packet benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "${1:-wpt or pipeline}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_quick.log
for w in coif4 sym5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "bench $w exit $?"
  python -c "
import json,sys
d=json.load(open('gpurun_out/q_$w.json')); r=d['roofline']
print('$w', round(d['value']), 'frames/s', 'ms', round(d['ms_per_step'],4), 'fma frac', round(r['frac'],3), 'hbm frac', round(r['hbm']['frac'],3), d['clocks'])
" || tail -5 gpurun_out/q_$w.err
done
