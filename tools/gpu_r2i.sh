#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r2i_pytest_gpu.log
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize.py > gpurun_out/r2i_sanitizer_$tool.log 2>&1; echo "sanitizer $tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize: OK|Error|hazard" gpurun_out/r2i_sanitizer_$tool.log | head -12
done
