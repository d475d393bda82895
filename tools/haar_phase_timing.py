#!/usr/bin/env python
"""Per-warp phase table of the linear Haar kernel: loads a -DAFD_HAAR_PHASE_TIMING=1 variant build and runs B = 4096 clips;
the library prints the table to stderr after every launch."""
import ctypes
import os
import sys

import torch

N, B = 22050, 4096
lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
x = torch.randn(B, N, device="cuda") * 0.1
sums = torch.zeros(16384, dtype=torch.float64, device="cuda")
count = torch.zeros(1, dtype=torch.int64, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    rc = lib.afd_haar_fingerprint_accum(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), 14,
                                        ctypes.c_void_p(sums.data_ptr()), ctypes.c_void_p(count.data_ptr()), stream)
    assert rc == 0, rc
torch.cuda.synchronize()
