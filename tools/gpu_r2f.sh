#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 5
python tools/wpt_diff.py audiodeepfake-detection_b200/libafd_b200_base.so audiodeepfake-detection_b200/libafd_b200.so coif4 4096
python tools/wpt_diff.py audiodeepfake-detection_b200/libafd_b200_base.so audiodeepfake-detection_b200/libafd_b200.so sym5 4096
python tools/wpt_diff.py audiodeepfake-detection_b200/libafd_b200_base.so audiodeepfake-detection_b200/libafd_b200.so db8 2000
for s in 0 600 1000; do
  echo "== stagger $s"
  AFD_WPT_STAGGER=$s python tools/ab_bench.py audiodeepfake-detection_b200/libafd_b200_base.so audiodeepfake-detection_b200/libafd_b200.so sym5 coif4 2>&1 | tee -a gpurun_out/r2f_ab.log
done
