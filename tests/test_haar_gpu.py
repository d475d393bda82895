"""GPU parity of the Haar fingerprint kernel against the oracle."""
import numpy as np
import pytest
import torch

import audiodeepfake_detection_b200 as afd
from oracle import wpt_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,level,B", [(22050, 14, 7), (22050, 8, 3), (22051, 14, 2), (16384, 14, 2), (4097, 10, 5),
                                       (22050, 1, 2), (333, 5, 3), (44100, 14, 2)])
def test_fingerprint_matches_oracle(N, level, B, cuda_device):
    rng = np.random.default_rng(N + level)
    x = (rng.standard_normal((B, 1, N)) * 0.1).astype(np.float32)
    sums, count = oracle.haar_fingerprint_sums(x.astype(np.float64), level, dtype=np.float64)
    acc = afd.FingerprintAccumulator(level, cuda_device).update(torch.from_numpy(x).to(cuda_device))
    assert int(acc.count.item()) == count
    got = acc.mean().cpu().numpy()
    want = sums / count
    assert got.shape == (1 << level,)
    assert np.max(np.abs(got - want)) < 1e-5 * np.max(want)


def test_streaming_updates_equal_one_shot(cuda_device):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(600, 1, 22050, device=cuda_device, generator=g) * 0.1
    one = afd.haar_fingerprint(x, 14)
    acc = afd.FingerprintAccumulator(14, cuda_device)
    for lo in range(0, 600, 250):
        acc.update(x[lo:lo + 250])
    assert torch.allclose(acc.mean(), one, rtol=1e-5, atol=0)   # fp32 partial sums per CTA: grouping changes rounding
    assert int(acc.count.item()) == 600 * 2


def test_reference_entry_point(cuda_device):
    """compute_fingerprint_wpt keeps _compute_fingerprint_wpt's conventions (fingerprints.py:93-99,114-115)."""
    clips = [torch.randn(1, n) * 0.1 for n in (30000, 22050, 22051, 50000)]       # == 1 s is dropped (strict >)
    freqs, mean = afd.compute_fingerprint_wpt(clips, device=cuda_device)
    assert freqs.shape == mean.shape == (16384,) and float(freqs[-1]) == 11025.0
    kept = torch.stack([c[:, :22050] for c in clips if c.shape[-1] > 22050]).numpy()
    sums, count = oracle.haar_fingerprint_sums(kept, 14)
    assert count == 3 * 2
    assert np.max(np.abs(mean.cpu().numpy() - sums / count)) < 1e-5 * np.max(sums / count)


def test_energy_preservation_full_batch(cuda_device):
    """Size-independent check at scale: Haar packets of an even-length-at-every-level signal are an orthonormal
    transform, so a constant clip puts all of its mass in packet 0 and the fingerprint is linear in |scale|."""
    x = torch.full((4096, 16384), 0.25, device=cuda_device)
    fp = afd.haar_fingerprint(x, 14)
    assert abs(float(fp[0]) - 0.25 * 128.0) < 1e-4 and float(fp[1:].abs().max()) < 1e-6
    g = torch.Generator(device="cuda").manual_seed(6)
    y = torch.randn(4096, 22050, device=cuda_device, generator=g)
    assert torch.allclose(afd.haar_fingerprint(3.0 * y, 14), 3.0 * afd.haar_fingerprint(y, 14), rtol=1e-5)


@pytest.mark.parametrize("B,N", [(37, 22050), (3, 2206), (1, 22050)])
def test_rfft_fingerprint_matches_reference_chain(B, N):
    """Mean-spectrum fingerprint (reference fingerprints.py:37-62): column sums on the GPU + one fp64 DFT."""
    from oracle import wpt_oracle

    rng = np.random.default_rng(B)
    clips = (rng.standard_normal((B, 1, N)) * 0.1 + 0.01).astype(np.float32)
    _, want = wpt_oracle.rfft_fingerprint(clips)
    acc = afd.SpectrumFingerprintAccumulator(N, "cuda")
    xt = torch.from_numpy(clips).cuda()
    half = B // 2
    if half:
        acc.update(xt[:half])                  # streaming: two updates equal one
    acc.update(xt[half:])
    got = acc.magnitude().cpu().numpy()
    assert got.shape == want.shape and int(acc.count.item()) == B
    assert np.max(np.abs(got - want)) < 1e-9 * np.max(want) + 1e-12


def test_compute_fingerprint_rfft_surface():
    clips = torch.randn(5, 1, 22050 + 100) * 0.1
    freqs, fp = afd.compute_fingerprint_rfft(clips, seconds=1, amount=4)
    assert freqs.shape == fp.shape == (11026,) and float(freqs[-1]) == 11025.0 and fp.is_cuda
    with pytest.raises(ValueError):
        afd.SpectrumFingerprintAccumulator(22051)


def test_large_batch_matches_oracle(cuda_device):
    """More clips than SMs x 2 (the streaming kernel's double-buffered walk, odd / even clip alignment): against the oracle."""
    rng = np.random.default_rng(77)
    x = (rng.standard_normal((700, 22050)) * 0.1).astype(np.float32)
    got = afd.haar_fingerprint(torch.from_numpy(x).to(cuda_device), 14).cpu().numpy()
    sums, count = oracle.haar_fingerprint_sums(x.astype(np.float64), 14, dtype=np.float64)
    assert count == 700 * 2
    assert np.max(np.abs(got - sums / count)) < 1e-5 * np.max(sums / count)
