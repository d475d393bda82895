#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stft_gpu.py -x -q -m gpu > gpurun_out/pytest_stft.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_stft.log
timeout 300 python bench.py --workload stft --no-cpu-baseline > gpurun_out/bench_stft2.json 2> gpurun_out/bench_stft2.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_stft2.json; tail -3 gpurun_out/bench_stft2.err
timeout 200 python tools/haar_probe.py 2>&1 | tail -2
for c in 512 2048; do timeout 300 python bench.py --workload haar --steps 20 --no-cpu-baseline --e2e-chunk $c 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('haar chunk $c e2e', round(d['e2e']['value']), 'value', round(d['value']))"; done
