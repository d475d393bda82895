#!/usr/bin/env python
"""Same-box A/B of one library under two environment settings (kernel-only, CUDA events, alternating rounds).
Usage: python tools/ab_env.py VAR valueA valueB [workloads...]"""
import ctypes
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ab_bench import B, N, make_step  # noqa: E402


def main():
    var, va, vb = sys.argv[1:4]
    workloads = sys.argv[4:] or ["coif4", "sym5"]
    lib = ctypes.CDLL(os.path.join(ROOT, "audiodeepfake-detection_b200", "libafd_b200.so"))
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, N, device="cuda", generator=g) * 0.1
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    setenv = ctypes.CDLL(None).setenv          # os.environ alone would not reach getenv() inside the C library
    setenv.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    for w in workloads:
        outs, times = [], [[], []]
        for rnd in range(7):
            for i, val in enumerate((va, vb)):
                setenv(var.encode(), val.encode(), 1)
                fn, out = make_step(lib, w, x, stream)
                for _ in range(5):
                    assert fn() == 0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(40):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                times[i].append(e0.elapsed_time(e1) / 40)
                if rnd == 0:
                    outs.append(out.clone())
        med = [statistics.median(t) for t in times]
        print(f"{w}: {var}={va} {med[0]*1e3:.1f} us ({B/med[0]/1e3:.3f} M/s)  {var}={vb} {med[1]*1e3:.1f} us "
              f"({B/med[1]/1e3:.3f} M/s)  time ratio {med[1]/med[0]:.4f}  identical={torch.equal(outs[0], outs[1])}", flush=True)


if __name__ == "__main__":
    main()
