#!/usr/bin/env python
"""Per-phase cycle table of the packet kernel: loads a -DAFD_WPT_PHASE_TIMING=1 build of the library
(build.build(True, extra_flags=["-DAFD_WPT_PHASE_TIMING=1"], out_path=".../libafd_b200_phase.so")) and runs the
headline shapes; the library prints thread-0 cycles per phase and half frame to stderr after every launch."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiodeepfake_detection_b200.wavelets import Wavelet  # noqa: E402

N, B = 22050, 4096


def main():
    lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
    raw = "--raw" in sys.argv            # log_scale = 0: the epilogue without the log
    names = [a for a in sys.argv[2:] if not a.startswith("--")] or ["sym5", "coif4"]
    x = torch.randn(B, N, device="cuda") * 0.1
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for name in names:
        taps = Wavelet(name).dec_lo
        F = len(taps)
        c_taps = (ctypes.c_double * F)(*taps)
        T = ctypes.c_int64()
        lib.afd_wpt_out_len(ctypes.c_int64(N), F, 8, ctypes.byref(T))
        out = torch.empty(B, 1, T.value, 256, device="cuda")
        for _ in range(3):
            rc = lib.afd_wpt_forward(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N),
                                     c_taps, F, 8, 0, ctypes.c_float(2.0), 0 if raw else 1, ctypes.c_float(1e-12), 0,
                                     ctypes.c_void_p(out.data_ptr()), None, stream)
            assert rc == 0
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
