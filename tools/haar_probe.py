#!/usr/bin/env python
"""Device time of afd_haar_fingerprint_accum per launch for several batch sizes (finds fixed per-launch costs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiodeepfake_detection_b200 as afd

dev = torch.device("cuda:0")
x = torch.randn(4096, 22050, device=dev) * 0.1
res = {}
for B in (64, 296, 512, 1024, 4096):
    acc = afd.FingerprintAccumulator(14, dev)
    xs = x[:B]
    for _ in range(3):
        acc.update(xs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        acc.update(xs)
    e1.record()
    torch.cuda.synchronize()
    res[B] = e0.elapsed_time(e1) / 20
print(json.dumps({"haar_ms_per_launch": res}))
