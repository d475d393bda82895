#!/bin/bash
# Round-1d measurement pass: bench lines (all workloads + reference arm), ncu launch list and full captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
for w in coif4 sym5 stft haar; do
  extra=""; [[ $w != coif4 ]] && extra="--no-cpu-baseline"
  timeout 600 python bench.py --workload $w $extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
  python -c "
import json
d=json.load(open('gpurun_out/bench_$w.json')); r=d['roofline']
print('$w', round(d['value']), d['unit'], 'ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None)
" || tail -5 gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_stft.csv \
    python bench.py --workload stft --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
for w in stft sym5 haar coif4; do
  pat="wpt_tree_kernel"; [[ $w == stft ]] && pat="stft_"; [[ $w == haar ]] && pat="haar_f"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/prof_$w \
      python bench.py --workload $w --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w exit $?"
done
