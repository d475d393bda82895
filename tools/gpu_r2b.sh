#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/wpt_phase_timing.py audiodeepfake-detection_b200/libafd_b200_phase.so sym5 coif4 2> gpurun_out/r2b_phase.txt; echo "phase exit $?"; cat gpurun_out/r2b_phase.txt
timeout 600 python bench.py --workload sym5_b128 --steps 50 --warmup 10 --no-workloads --no-e2e --no-cpu-baseline > gpurun_out/r2b_b128.json 2> gpurun_out/r2b_b128.err; echo "b128 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2b_b128.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r2b_sanitizer_$tool.log 2>&1; echo "sanitizer $tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize:" gpurun_out/r2b_sanitizer_$tool.log | tail -8
done
