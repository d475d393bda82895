#!/bin/bash
# r1c pass: parity tests, smoke, PCIe probe, full bench lines for every workload, e2e chunk sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu.csv 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 200 python tools/pcie_probe.py > gpurun_out/pcie_probe.json 2>&1; cat gpurun_out/pcie_probe.json
for c in 128 256 1024; do
  timeout 300 python bench.py --workload coif4 --steps 20 --no-cpu-baseline --e2e-chunk $c > gpurun_out/e2e_chunk_$c.json 2>&1
  python -c "
import json; d=json.load(open('gpurun_out/e2e_chunk_$c.json')); print('chunk $c e2e', round(d['e2e']['value']))"
done
for w in coif4 sym5 stft haar; do
  extra=""; [[ $w != coif4 ]] && extra="--no-cpu-baseline"
  timeout 600 python bench.py --workload $w $extra > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3))"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>&1; tail -c 600 gpurun_out/bench_reference.json
