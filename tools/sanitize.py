#!/usr/bin/env python
"""Small-shape pass over every kernel of the path for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize.py
    compute-sanitizer --tool memcheck  python tools/sanitize.py

Covers wpt_tree_kernel (lattice, direct form, one CTA per SM, extended epilogue, sliced level L-1), stft_tc511_kernel
(tcgen05 / mbarrier pipeline), stft_pfa511_kernel, stft_bluestein_kernel, haar_fast_kernel, haar_fingerprint_kernel,
clip_sum_kernel and the resampler.  Results are compared with the oracle so a run that "passes" also computed the
right thing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import audiodeepfake_detection_b200 as afd  # noqa: E402
from audiodeepfake_detection_b200.wavelets import Wavelet  # noqa: E402
from oracle import ptwt_like, wpt_oracle  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((3, 1, 22050)) * 0.1).astype(np.float32)
    xt = torch.from_numpy(x).to(dev)
    for name, level in (("sym5", 8), ("coif4", 8), ("db20", 8), ("coif10", 8), ("haar", 8), ("sym5", 3), ("db4", 11)):
        got = afd.wavelet_packet_features(xt, Wavelet(name), level).cpu().numpy()
        want = wpt_oracle.packet_features(x.astype(np.float64), Wavelet(name).dec_lo, level, dtype=np.float64)
        print(f"sanitize: packets {name} L{level} rel {rel(got, want):.2e}", flush=True)
        assert rel(got, want) < 1e-5
    # more frames than SMs: the frame kernel's persistent loop, prefetch hand-over between the groups and buffer re-use
    rngb = np.random.default_rng(1)
    xb = (rngb.standard_normal((160, 22050)) * 0.1).astype(np.float32)
    for name in ("sym5", "coif4"):
        got = afd.wavelet_packet_features(torch.from_numpy(xb).to(dev), Wavelet(name), 8).cpu().numpy()
        for f in (0, 77, 148, 159):
            want = wpt_oracle.packet_features(xb[f:f + 1].astype(np.float64), Wavelet(name).dec_lo, 8, dtype=np.float64)
            assert rel(got[f:f + 1], want) < 1e-5, (name, f)
        print(f"sanitize: packets {name} L8 batch 160 ok", flush=True)
    mod = afd.Packets(wavelet_str="sym5", max_lev=8, log_scale=True, loss_less=True, compute_welford=True)
    feats, aux = mod(xt)
    assert torch.isfinite(feats).all() and len(aux) == 256
    print("sanitize: extended epilogue ok", flush=True)

    for impl in ("tc", "pfa"):
        os.environ["AFD_STFT_IMPL"] = impl
        spec = afd.stft_power_features(xt, 511, 220).cpu().numpy()[:, 0]
        want = ptwt_like.stft_power_dft64(x[:, 0]).transpose(0, 2, 1)      # [B, bins, frames] -> [B, frames, bins]
        print(f"sanitize: stft 511/220 impl={impl} rel {rel(spec, want):.2e}", flush=True)
        assert rel(spec, want) < 1e-5
    os.environ.pop("AFD_STFT_IMPL", None)
    spec = afd.stft_power_features(xt, 256, 128).cpu().numpy()
    assert np.isfinite(spec).all()
    print("sanitize: stft 256/128 (bluestein) ok", flush=True)

    for N, level in ((22050, 14), (4097, 10)):
        xx = (rng.standard_normal((5, 1, N)) * 0.1).astype(np.float32)
        got = afd.haar_fingerprint(torch.from_numpy(xx).to(dev), level).cpu().numpy()
        sums, count = wpt_oracle.haar_fingerprint_sums(xx.astype(np.float64), level, dtype=np.float64)
        print(f"sanitize: haar N={N} L{level} rel {rel(got, sums / count):.2e}", flush=True)
        assert rel(got, sums / count) < 1e-5
    xh = (rng.standard_normal((300, 22050)) * 0.1).astype(np.float32)      # two clips per CTA: double-buffered bulk staging
    got = afd.haar_fingerprint(torch.from_numpy(xh).to(dev), 14).cpu().numpy()
    sums, count = wpt_oracle.haar_fingerprint_sums(xh.astype(np.float64), 14, dtype=np.float64)
    print(f"sanitize: haar batch 300 rel {rel(got, sums / count):.2e}", flush=True)
    assert rel(got, sums / count) < 1e-5
    from oracle import resample_oracle
    xr = (rng.standard_normal((3, 6000)) * 0.1).astype(np.float32)
    for orig in (44100, 48000):
        got = afd.framing.resample(torch.from_numpy(xr).to(dev), orig, 22050).cpu().numpy()
        want = resample_oracle.resample(xr, orig, 22050)
        print(f"sanitize: resample {orig} -> 22050 rel {rel(got, want):.2e}", flush=True)
        assert rel(got, want) < 1e-5
    acc = afd.SpectrumFingerprintAccumulator(22050, dev)
    acc.update(xt)
    torch.cuda.synchronize()
    print("sanitize: OK", flush=True)


if __name__ == "__main__":
    main()
