#!/bin/bash
# Scaling record on one 8-GPU box: the driver-style bench line at N = 1, 2, 4, 8 (torchrun for N > 1), GPU tests first.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/r2s_pytest_gpu.log
for n in 1 2 4 8; do
  if [[ $n == 1 ]]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_bench_${n}gpu.json 2> gpurun_out/r2s_bench_${n}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) \
        bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2s_bench_${n}gpu.json 2> gpurun_out/r2s_bench_${n}gpu.err
  fi
  python - <<PY
import json
d = json.load(open('gpurun_out/r2s_bench_${n}gpu.json'))
w = d.get('workloads', {})
print('N=${n} coif4', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['copy_ceiling']['value']),
      '| sym5', round(w['sym5']['value']), '| stft', round(w['stft']['value']), '| haar job', round(w['haar']['value']), 'ms', round(w['haar']['ms_job'], 2),
      '| rfft', round(w['rfft']['value']))
PY
done
