"""CPU restatement of torchaudio.functional.resample (TEST INFRASTRUCTURE ONLY) -- the call the reference makes per
loaded window at src/audiofakedetect/data_loader.py:341-344 with torchaudio's defaults (sinc_interp_hann,
lowpass_filter_width 6, rolloff 0.99).  Follows torchaudio/functional/functional.py::_get_sinc_resample_kernel and
::_apply_sinc_resample_kernel.  torchaudio evaluates the taps in the waveform's dtype (float32) operation by operation;
``dtype=np.float32`` (default) follows that order and reproduces its table to an ulp, ``dtype=np.float64`` is the exact
filter (up to 2e-5 of the signal scale away for ratios like 640:441).  The convolution itself is accumulated in float64.
tests/test_resample_cpu.py pins this file on committed torchaudio outputs (tests/golden/resample_torchaudio.npz).
"""
import math

import numpy as np


def sinc_resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99,
                       dtype=np.float32):
    """-> (taps [new, 2 * width + orig], width, orig, new) with the rates reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    f = dtype
    idx = (np.arange(-width, width + orig).astype(f) / f(orig))[None, :]
    t = (np.arange(0, -new, -1).astype(f) / f(new))[:, None] + idx
    t = t * f(base)
    t = np.clip(t, f(-lowpass_filter_width), f(lowpass_filter_width))
    window = np.cos(t * f(math.pi) / f(lowpass_filter_width) / f(2)) ** 2
    t = t * f(math.pi)
    sinc = np.where(t == 0, f(1), np.sin(t) / np.where(t == 0, f(1), t))
    return (sinc * (window * f(base / orig))).astype(f), width, orig, new


def resample(x, orig_freq: int, new_freq: int, dtype=np.float32):
    """x [..., n] -> [..., ceil(new * n / orig)] (float64 accumulation of ``dtype`` taps)."""
    x = np.asarray(x, dtype=np.float64)
    if int(orig_freq) == int(new_freq):
        return x
    taps, width, orig, new = sinc_resample_taps(orig_freq, new_freq, dtype=dtype)
    taps = taps.astype(np.float64)
    lead, n = x.shape[:-1], x.shape[-1]
    xp = np.pad(x.reshape(-1, n), ((0, 0), (width, width + orig)))
    K = taps.shape[1]
    frames = (xp.shape[1] - K) // orig + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, K, axis=-1)[:, ::orig][:, :frames]      # [rows, frames, K]
    y = np.einsum("rfk,pk->rfp", win, taps).reshape(xp.shape[0], -1)
    target = -(-new * n // orig)
    return y[:, :target].reshape(*lead, target)
