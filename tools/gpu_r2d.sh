#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wpt_gpu.py tests/test_fused_epilogue_gpu.py tests/test_published_kats_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -n 5
python tools/ab_bench.py audiodeepfake-detection_b200/libafd_b200_base.so audiodeepfake-detection_b200/libafd_b200.so sym5 coif4 2>&1 | tee gpurun_out/r2d_ab.log
python tools/wpt_phase_timing.py audiodeepfake-detection_b200/libafd_b200_phase.so sym5 coif4 2> gpurun_out/r2d_phase.txt; tail -n 6 gpurun_out/r2d_phase.txt
