# bash tools/gpu_stagger_sweep.sh "wavelets" "staggers": kernel-only time of the packet kernel for every (wavelet, AFD_WPT_STAGGER) pair
for w in ${1:-db2 sym4 coif2 sym8 sym10 coif4}; do
  for s in ${2:-0 400 800 1200 1600}; do
    echo -n "$w stagger $s: "; AFD_WPT_STAGGER=$s python tools/ab_bench.py audiodeepfake-detection_b200/libafd_b200.so audiodeepfake-detection_b200/libafd_b200.so $w | sed "s/ B .*//"
  done
done
