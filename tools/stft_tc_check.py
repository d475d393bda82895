#!/usr/bin/env python
"""tcgen05 STFT-511 kernel (AFD_STFT_IMPL=tc) against the mma.sync prime-factor kernel and the fp64 DFT: values and time."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.environ.get("AFD_LIB") or os.path.join(ROOT, "audiodeepfake-detection_b200", "libafd_b200.so"))
lib.afd_last_error.restype = ctypes.c_char_p


def run(x, impl, hop=220, log_scale=1, out=None):
    if impl:
        os.environ["AFD_STFT_IMPL"] = impl
    else:
        os.environ.pop("AFD_STFT_IMPL", None)
    B, N = x.shape
    frames = 1 + (N - 1) // hop
    if out is None:
        out = torch.full((B, 1, frames, 256), float("nan"), device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.afd_stft_power(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(x.stride(0)), 511, hop,
                            ctypes.c_float(2.0), log_scale, ctypes.c_float(1e-12), ctypes.c_void_p(out.data_ptr()), stream)
    assert rc == 0, lib.afd_last_error()
    return out


def ref64(x, hop=220):
    w = torch.hann_window(511, periodic=True, dtype=torch.float64, device=x.device)
    s = torch.stft(x.double(), 511, hop, window=w, center=True, pad_mode="reflect", return_complex=True)
    return (s.abs() ** 2).transpose(1, 2).unsqueeze(1)      # [B,1,frames,256]


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, N, hop) in [(3, 22050, 220), (2, 4000, 100), (5, 22050, 242)]:
        x = torch.randn(B, N, device="cuda", generator=g) * 0.1
        a = run(x, "tc", hop, log_scale=0)
        torch.cuda.synchronize()
        r = ref64(x, hop)
        b = run(x, "pfa", hop, log_scale=0)
        torch.cuda.synchronize()
        ea = ((a.double() - r).norm() / r.norm()).item()
        eb = ((b.double() - r).norm() / r.norm()).item()
        print(f"B={B} N={N} hop={hop}: tc rel err {ea:.3e} (max abs {(a.double()-r).abs().max().item():.3e}, nan {int(torch.isnan(a).sum())})  pfa rel err {eb:.3e}", flush=True)
    B, N = 4096, 22050
    x = torch.randn(B, N, device="cuda", generator=g) * 0.1
    outs = {}
    for impl in ("tc", "pfa", "tc", "pfa"):
        out = run(x, impl)
        for _ in range(5):
            run(x, impl, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            run(x, impl, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        outs[impl] = out
        print(f"{impl or 'pfa'}: {ms*1e3:.1f} us  {B/ms/1e3:.3f} M frames/s", flush=True)
    d = (outs["tc"] - outs["pfa"]).abs().max().item()
    print("max |log tc - log pfa| =", d)


if __name__ == "__main__":
    main()
