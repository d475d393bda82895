#!/usr/bin/env python
"""Same-box A/B of two builds of libafd_b200.so (kernel-only, CUDA events, alternating rounds).
Usage: python tools/ab_bench.py <libA.so> <libB.so> [workloads...]   (workloads: coif4 sym5 stft haar)"""
import ctypes
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodeepfake_detection_b200.wavelets import Wavelet  # noqa: E402

N, B = 22050, 4096


def make_step(lib, workload, x, stream):
    xp = ctypes.c_void_p(x.data_ptr())
    if workload not in ("stft", "haar"):
        taps = Wavelet(workload).dec_lo
        F = len(taps)
        c_taps = (ctypes.c_double * F)(*taps)
        T = ctypes.c_int64()
        lib.afd_wpt_out_len(ctypes.c_int64(N), F, 8, ctypes.byref(T))
        out = torch.empty(B, 1, T.value, 256, device="cuda")
        op = ctypes.c_void_p(out.data_ptr())
        return lambda: lib.afd_wpt_forward(xp, ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), c_taps, F, 8, 0,
                                           ctypes.c_float(2.0), 1, ctypes.c_float(1e-12), 0, op, None, stream), out
    if workload == "stft":
        out = torch.empty(B, 1, 101, 256, device="cuda")
        op = ctypes.c_void_p(out.data_ptr())
        return lambda: lib.afd_stft_power(xp, ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), 511, 220,
                                          ctypes.c_float(2.0), 1, ctypes.c_float(1e-12), op, stream), out
    sums = torch.zeros(16384, dtype=torch.float64, device="cuda")
    sp = ctypes.c_void_p(sums.data_ptr())
    return lambda: lib.afd_haar_fingerprint_accum(xp, ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), 14, sp,
                                                  None, stream), sums


def main():
    paths = sys.argv[1:3]
    workloads = sys.argv[3:] or ["coif4", "sym5", "stft"]
    libs = [ctypes.CDLL(os.path.abspath(p)) for p in paths]
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, N, device="cuda", generator=g) * 0.1
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for w in workloads:
        steps = [make_step(lib, w, x, stream) for lib in libs]
        times = [[], []]
        for rnd in range(7):
            for i, (fn, _) in enumerate(steps):
                for _ in range(5):
                    assert fn() == 0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(40):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                times[i].append(e0.elapsed_time(e1) / 40)
        same = torch.equal(steps[0][1], steps[1][1]) if w != "haar" else None
        med = [statistics.median(t) for t in times]
        print(f"{w}: A {med[0]*1e3:.1f} us ({B/med[0]/1e3:.3f} M/s)  B {med[1]*1e3:.1f} us ({B/med[1]/1e3:.3f} M/s)  "
              f"B/A time {med[1]/med[0]:.4f}  identical={same}", flush=True)


if __name__ == "__main__":
    main()
