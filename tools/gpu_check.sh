#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (all workloads), ncu launch list + full capture of the top kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tests|bench|ncu|all]
set -u
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
if [[ $what == all || $what == tests ]]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  for w in coif4 sym5 stft haar; do
    extra=""; [[ $w != coif4 ]] && extra="--no-cpu-baseline"
    timeout 600 python bench.py --workload $w $extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
    cat gpurun_out/bench_$w.json
  done
  timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_coif4.csv \
      python bench.py --workload coif4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  for w in coif4 sym5 stft haar; do
    pat="wpt_tree_kernel"; [[ $w == stft ]] && pat="stft_"; [[ $w == haar ]] && pat="haar_fingerprint"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/prof_$w \
        python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
    echo "ncu $w exit $?"
  done
fi
