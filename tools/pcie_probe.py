#!/usr/bin/env python
"""Host<->device copy bandwidth of this box (pinned memory, CUDA events): H2D alone, D2H alone, both at once.
The e2e leg of bench.py is bounded by these numbers; run under gpurun and keep the output beside the bench line."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def timed(fn, iters=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    dev = torch.device("cuda:0")
    res = {}
    for mb in (16, 64, 256):
        n = mb * (1 << 20) // 4
        h_in = torch.empty(n, dtype=torch.float32).pin_memory()
        h_out = torch.empty(n, dtype=torch.float32).pin_memory()
        h_in.fill_(1.0)
        h_out.fill_(0.0)
        d_a = torch.empty(n, dtype=torch.float32, device=dev)
        d_b = torch.ones(n, dtype=torch.float32, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        t_h2d = timed(lambda: d_a.copy_(h_in, non_blocking=True))
        t_d2h = timed(lambda: h_out.copy_(d_b, non_blocking=True))

        def both():
            s1.wait_stream(torch.cuda.current_stream())
            s2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
        t_both = timed(both)
        gb = n * 4 / 1e9
        res[f"{mb}MB"] = {"h2d_GBs": gb / t_h2d, "d2h_GBs": gb / t_d2h, "bidir_each_GBs": gb / t_both}
    print(json.dumps({"pcie_probe": res}))


if __name__ == "__main__":
    main()
