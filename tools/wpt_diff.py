#!/usr/bin/env python
"""Compare two builds of the library on the packet transform frame by frame (debug aid):
python tools/wpt_diff.py libA.so libB.so [wavelet] [B]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodeepfake_detection_b200.wavelets import Wavelet  # noqa: E402

N = 22050


def run(lib, x, name):
    taps = Wavelet(name).dec_lo
    F = len(taps)
    c_taps = (ctypes.c_double * F)(*taps)
    T = ctypes.c_int64()
    B = x.shape[0]
    lib.afd_wpt_out_len(ctypes.c_int64(N), F, 8, ctypes.byref(T))
    out = torch.zeros(B, 1, T.value, 256, device="cuda")
    rc = lib.afd_wpt_forward(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), c_taps, F, 8,
                             0, ctypes.c_float(2.0), 0, ctypes.c_float(1e-12), 0, ctypes.c_void_p(out.data_ptr()), None,
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.afd_last_error()
    torch.cuda.synchronize()
    return out


def main():
    la, lb = (ctypes.CDLL(os.path.abspath(p)) for p in sys.argv[1:3])
    for l in (la, lb):
        l.afd_last_error.restype = ctypes.c_char_p
    name = sys.argv[3] if len(sys.argv) > 3 else "sym5"
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 600
    x = torch.randn(B, N, device="cuda") * 0.1
    a, b = run(la, x, name), run(lb, x, name)
    scale = a.abs().max()
    err = (a - b).abs().amax(dim=(1, 2, 3)) / scale
    bad = (err > 1e-5).nonzero().flatten().tolist()
    print(f"{name} B={B}: max rel err {float(err.max()):.3e}, bad frames {len(bad)}: {bad[:40]}")
    if bad:
        f = bad[0]
        d = (a[f, 0] - b[f, 0]).abs() / scale          # [T, 256]
        cols = (d.amax(0) > 1e-5).nonzero().flatten().tolist()
        rows = (d.amax(1) > 1e-5).nonzero().flatten().tolist()
        print(f"frame {f}: bad columns {len(cols)}: {cols[:64]}")
        print(f"frame {f}: bad rows {len(rows)}: {rows[:64]}")


if __name__ == "__main__":
    main()
