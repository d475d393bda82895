#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wpt_gpu.py tests/test_fused_epilogue_gpu.py tests/test_published_kats_gpu.py tests/test_pipeline_gpu.py -q 2>&1 | tail -n 5
for s in 0 400 800 1200 1600 2400; do
  echo "== stagger $s"
  AFD_WPT_STAGGER=$s python tools/ab_bench.py audiodeepfake-detection_b200/libafd_b200_base.so audiodeepfake-detection_b200/libafd_b200.so sym5 coif4 2>&1 | tee -a gpurun_out/r2e_ab.log
done
