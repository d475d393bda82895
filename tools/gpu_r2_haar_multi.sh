#!/bin/bash
set -u
N=${1:-8}
mkdir -p gpurun_out
for n in 1 2 4 $N; do
  [[ $n -gt $N ]] && continue
  if [[ $n == 1 ]]; then
    timeout 600 python bench.py --workload haar --no-workloads --no-e2e --no-cpu-baseline > gpurun_out/r2_haar_job_${n}gpu.json 2> gpurun_out/r2_haar_job_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
        bench.py --gpus $n --workload haar --no-workloads --no-e2e --no-cpu-baseline > gpurun_out/r2_haar_job_${n}gpu.json 2> gpurun_out/r2_haar_job_${n}gpu.err
  fi
  python - <<PY
import json
d = json.load(open('gpurun_out/r2_haar_job_${n}gpu.json'))
print('N=${n}', round(d['value']), 'frames/s  job ms', round(d['ms_per_step'], 3), 'accumulate ms by rank', [round(v, 3) for v in d['ms_accumulate_by_rank']], 'allreduce us', [round(v) for v in d['allreduce_us_by_rank']])
PY
done
