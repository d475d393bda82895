#!/bin/bash
set -u
mkdir -p gpurun_out
bash tools/gpu_ncu.sh haar haar_linear r2_haar
bash tools/gpu_ncu.sh sym5 wpt_frame r2_sym5
bash tools/gpu_ncu.sh coif4 wpt_frame r2_coif4
ls -la gpurun_out/*.ncu-rep
