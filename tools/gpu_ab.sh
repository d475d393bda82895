#!/bin/bash
# same-box A/B of variant builds against the product library: bash tools/gpu_ab.sh "workloads" name1 name2 ...
set -u
mkdir -p gpurun_out
W="$1"; shift
for v in "$@"; do
  echo "== variant $v"
  python tools/ab_bench.py audiodeepfake-detection_b200/libafd_b200.so audiodeepfake-detection_b200/libafd_b200_$v.so $W 2>&1 | tee -a gpurun_out/ab_$v.log
done
