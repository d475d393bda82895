#!/bin/bash
# multi-GPU record: bash tools/gpu_r2_multi.sh N   (full bench line on N GPUs, then the config-5 training step)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/r2_gpus_${N}.csv 2>&1
nvidia-smi topo -m > gpurun_out/r2_topo_${N}.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench N=$N exit $?"
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print('coif4', round(d['value']), 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'ceiling', round(d['e2e']['copy_ceiling']['value']), d['e2e']['copy_ceiling'].get('host_GBs_all_ranks'))
for k, v in d.get('workloads', {}).items():
    if 'error' in v:
        print(k, v); continue
    print(k, round(v['value']), {a: v[a] for a in v if a in ('ms_per_step', 'ms_job', 'ms_accumulate_by_rank', 'allreduce_us_by_rank')}, 'e2e', (v.get('e2e') or {}).get('value'))
PY
tail -n 3 gpurun_out/r2_bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tools/train_step_bench.py > gpurun_out/r2_train_step_${N}gpu.json 2> gpurun_out/r2_train_step_${N}gpu.err; echo "train step N=$N exit $?"; tail -n 2 gpurun_out/r2_train_step_${N}gpu.json
