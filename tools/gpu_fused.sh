#!/bin/bash
# fused-epilogue pass: new parity tests, then the whole GPU suite, then quick benches (regression check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_epilogue_gpu.py -x -q -m gpu > gpurun_out/pytest_fused.log 2>&1; echo "fused exit $?"; tail -30 gpurun_out/pytest_fused.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
for w in coif4 sym5 stft; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "bench $w exit $?"
  python -c "
import json
d=json.load(open('gpurun_out/q_$w.json')); r=d['roofline']
print('$w', round(d['value']), 'frames/s', 'ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3))
" || tail -5 gpurun_out/q_$w.err
done
