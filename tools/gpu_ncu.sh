#!/bin/bash
# ncu --set full capture of one workload's kernel:  bash tools/gpu_ncu.sh <workload> <kernel regex> [tag]
w=${1:-coif4}; pat=${2:-wpt_tree_kernel}; tag=${3:-$w}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
echo "ncu $w exit $?"; tail -3 gpurun_out/ncu_$tag.log
