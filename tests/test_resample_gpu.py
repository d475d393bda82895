"""GPU parity of the device resampler / frame cutter against torchaudio's committed outputs and the oracle."""
import os

import numpy as np
import pytest
import torch

import audiodeepfake_detection_b200 as afd
from oracle import resample_oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_torchaudio.npz")


@pytest.mark.parametrize("orig", [44100, 48000, 24000, 32000])
def test_resample_matches_torchaudio_outputs(orig):
    g = np.load(GOLDEN)
    x, want = g[f"x_{orig}"], g[f"y_{orig}"]
    got = afd.framing.resample(torch.from_numpy(x).cuda(), orig, 22050).cpu().numpy()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < 1e-5 * np.max(np.abs(want))      # fp32 tolerance of the path (measured ~3e-7)


@pytest.mark.parametrize("orig,new,n,rows", [(44100, 22050, 44100, 5), (48000, 22050, 48000, 3), (24000, 22050, 24000, 2),
                                             (16000, 22050, 4001, 2), (8, 1, 100, 1), (22050, 22050, 50, 2)])
def test_resample_matches_oracle(orig, new, n, rows):
    rng = np.random.default_rng(orig + n)
    x = (rng.standard_normal((rows, n)) * 0.1).astype(np.float32)
    want = resample_oracle.resample(x, orig, new)
    got = afd.framing.resample(torch.from_numpy(x).cuda(), orig, new).cpu().numpy()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < 1e-5 * max(np.max(np.abs(want)), 1e-30)


def test_cut_frames_resamples_each_window_like_the_reference():
    """data_loader.py:176-182 + :336-344: windows of int(seconds * file_rate) samples, each resampled on its own."""
    rng = np.random.default_rng(0)
    audio = (rng.standard_normal(3 * 44100 + 999) * 0.1).astype(np.float32)
    frames = afd.cut_frames(torch.from_numpy(audio).cuda(), seconds=1, sample_rate=22050, orig_sample_rate=44100)
    assert tuple(frames.shape) == (3, 1, 22050)
    for i in range(3):
        want = resample_oracle.resample(audio[i * 44100:(i + 1) * 44100], 44100, 22050)
        assert np.max(np.abs(frames[i, 0].cpu().numpy() - want)) < 1e-5 * np.max(np.abs(want))
    with pytest.raises(RuntimeError):
        afd.cut_frames(torch.from_numpy(audio).cuda(), sample_rate=22050, orig_sample_rate=16000)
    feats, _ = afd.utterance_features(afd.Packets("sym5", 8, log_scale=True), torch.from_numpy(audio).cuda(),
                                      orig_sample_rate=44100)
    assert tuple(feats.shape) == (3, 1, 256, 95)
