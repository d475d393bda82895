"""The resample oracle against committed outputs of torchaudio.functional.resample -- the reference's own call
(src/audiofakedetect/data_loader.py:341-344); generator: tools/make_golden.py."""
import os

import numpy as np
import pytest

from oracle import resample_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_torchaudio.npz")
RATES = (44100, 48000, 24000, 32000)


@pytest.mark.parametrize("orig", RATES)
def test_oracle_matches_torchaudio_outputs(orig):
    g = np.load(GOLDEN)
    x, want = g[f"x_{orig}"], g[f"y_{orig}"]
    got = resample_oracle.resample(x, orig, 22050)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < 2e-6 * np.max(np.abs(want))      # fp32 conv1d accumulation on torchaudio's side
    exact = resample_oracle.resample(x, orig, 22050, dtype=np.float64)
    assert np.max(np.abs(exact - want)) < 3e-5 * np.max(np.abs(want))    # the float64 filter: torchaudio's fp32 taps differ


def test_output_length_and_identity():
    for n, orig in ((44100, 44100), (48000, 48000), (1000, 24000), (22051, 44100)):
        y = resample_oracle.resample(np.zeros((1, n)), orig, 22050)
        assert y.shape[-1] == -(-22050 * n // orig)
    x = np.arange(10.0)[None]
    assert resample_oracle.resample(x, 22050, 22050) is not None and np.array_equal(resample_oracle.resample(x, 5, 5), x)


def test_live_torchaudio_if_importable():
    """Same comparison against a live torchaudio when the image has it (it does in the build container)."""
    torch = pytest.importorskip("torch")
    AF = pytest.importorskip("torchaudio.functional")
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((3, 6000)) * 0.2).astype(np.float32)
    for orig in (44100, 16000 * 3):
        want = AF.resample(torch.from_numpy(x), orig, 22050).numpy()
        got = resample_oracle.resample(x, orig, 22050)
        assert got.shape == want.shape and np.max(np.abs(got - want)) < 2e-6 * np.max(np.abs(want))
