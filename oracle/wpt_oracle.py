"""numpy restatement of the wavelet-packet feature path (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows, step by step:
  * reference wavelet_math.py:182   ptwt.WaveletPacket(data, wavelet, mode="reflect")
      - ptwt ``_get_pad``: pad (2F-3)//2 = F-2 samples on both sides, one more on the right if the node
        length is odd; ptwt ``_fwt_pad``: torch ``F.pad(mode="reflect")`` (whole-sample symmetric, edge
        not repeated); ptwt ``wavedec``: ``conv1d(x_pad, stack([dec_lo[::-1], dec_hi[::-1]]), stride=2)``
        i.e. y[k] = sum_m h[m] * x~[2k+1-m], k = 0 .. floor((n+F-1)/2)-1   (same as pywt's dwt).
      - children paths p+"a" (low-pass), p+"d" (high-pass), recursion down to ``max_lev``.
  * reference wavelet_math.py:185   get_level(max_lev): frequency (Gray-code) ordering of the leaves.
  * reference wavelet_math.py:191-206  stack leaves along a new last axis -> [B, T, P].
  * reference wavelet_math.py:208-218  log(|c|**power + 1e-12), optional sign channel, channel axis.
  * reference wavelet_math.py:263   Packets.forward returns .permute(0, 1, 3, 2).
  * reference fingerprints.py:99-115  Haar level-14 tree, freq order, mean |c| over clips/channel/position.
"""
import numpy as np

from .filters import dec_hi as _dec_hi


def reflect_index(idx, n):
    """Map extended indices onto [0, n) by whole-sample symmetric reflection (torch/pywt 'reflect')."""
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.mod(idx, period)
    return np.where(idx >= n, period - idx, idx)


def out_len(n, F):
    """Length of one analysis step's output: floor((n + F - 1) / 2)  (ptwt _get_pad comment, pywt dwt_coeff_len)."""
    return (n + F - 1) // 2


def level_lengths(n, F, level):
    out = []
    for _ in range(level):
        n = out_len(n, F)
        out.append(n)
    return out


def extension_index(idx, n, mode):
    """Signal-extension index maps as PyWavelets documents them (pywt "Signal extension modes"; pywt.pad is
    numpy.pad for these): 'reflect' = whole-sample symmetric ... x2 x1 | x0 x1 .. xn-1 | xn-2 xn-3 ...;
    'symmetric' = half-sample symmetric ... x1 x0 | x0 x1 .. xn-1 | xn-1 xn-2 ... (pywt's default mode, used by
    the published examples); 'zero' returns -1 for positions outside the signal."""
    if mode == "reflect":
        return reflect_index(idx, n)
    if mode == "symmetric":
        period = 2 * n
        idx = np.mod(idx, period)
        return np.where(idx >= n, period - 1 - idx, idx)
    if mode == "zero":
        return np.where((idx < 0) | (idx >= n), -1, idx)
    raise ValueError(mode)


def dwt_step(x, dec_lo, dtype=np.float32, mode="reflect"):
    """One analysis step on the last axis: returns (lo, hi).  ptwt wavedec(level=1, mode='reflect').

    ``mode`` other than 'reflect' exists only so that the published PyWavelets examples (which use pywt's default
    'symmetric' mode, or 'zero' in ptwt's README) can pin the convolution phase and tap orientation."""
    x = np.asarray(x, dtype=dtype)
    h = np.asarray(dec_lo, dtype=np.float64)
    g = _dec_hi(h)
    F = h.shape[0]
    n = x.shape[-1]
    padl = F - 2
    padr = F - 2 + (n % 2)
    idx = extension_index(np.arange(-padl, n + padr), n, mode)
    if mode == "zero":
        xp = np.where(idx >= 0, x[..., np.maximum(idx, 0)], dtype(0))
    else:
        xp = x[..., idx]
    win = np.lib.stride_tricks.sliding_window_view(xp, F, axis=-1)[..., ::2, :]
    # conv1d is a correlation, ptwt flips the filters: taps applied are h[::-1]
    lo = win @ h[::-1].astype(dtype)
    hi = win @ g[::-1].astype(dtype)
    assert lo.shape[-1] == out_len(n, F)
    return lo.astype(dtype), hi.astype(dtype)


def graycode_paths(level):
    """ptwt WaveletPacket._get_graycode_order / pywt get_level(order='freq')."""
    order = ["a", "d"]
    for _ in range(level - 1):
        order = ["a" + p for p in order] + ["d" + p for p in order[::-1]]
    return order


def natural_paths(level):
    order = [""]
    for _ in range(level):
        order = [p + c for p in order for c in "ad"]
    return order


def wavelet_packet_tree(x, dec_lo, level, dtype=np.float32, mode="reflect"):
    """All nodes down to ``level``: dict path -> array[..., L_level]."""
    tree = {"": np.asarray(x, dtype=dtype)}
    frontier = [""]
    for _ in range(level):
        nxt = []
        for p in frontier:
            lo, hi = dwt_step(tree[p], dec_lo, dtype, mode)
            tree[p + "a"], tree[p + "d"] = lo, hi
            nxt += [p + "a", p + "d"]
        frontier = nxt
    return tree


def packet_coefficients(x, dec_lo, level, order="freq", dtype=np.float32):
    """[B, T, P] stack of level-``level`` leaves (reference wavelet_math.py:185-206)."""
    x = np.asarray(x)
    if x.ndim == 3:  # [B, 1, N] -> [B, N] (old ptwt squeezes the channel axis)
        x = x[:, 0, :]
    tree = wavelet_packet_tree(x, dec_lo, level, dtype)
    paths = graycode_paths(level) if order == "freq" else natural_paths(level)
    return np.stack([tree[p] for p in paths], axis=-1)


def packet_features(x, dec_lo, level=8, log_scale=False, loss_less=False, power=2.0, order="freq",
                    dtype=np.float32):
    """compute_pytorch_packet_representation (wavelet_math.py:167-220): returns [B, C, T, P] (memory order).

    ``Packets.forward`` hands out ``.transpose(0, 1, 3, 2)`` of this, i.e. logical [B, C, P, T].
    """
    wp = packet_coefficients(x, dec_lo, level, order, dtype)
    if log_scale:
        wp_log = np.log(np.abs(wp).astype(dtype) ** dtype(power) + dtype(1e-12)).astype(dtype)
        if loss_less:
            sign = (((wp < 0).astype(dtype) * dtype(-1)) + dtype(0.5)) * dtype(2)
            return np.stack([wp_log, sign], axis=1)
        return wp_log[:, None]
    return wp[:, None]


def haar_fingerprint_sums(clips, level=14, dtype=np.float32):
    """Sum over clips and positions of |c| per frequency-ordered Haar packet (fingerprints.py:99-115).

    Returns (sums[2**level] float64, count) with mean = sums / count, count = clips * channel * positions.
    """
    from .filters import DEC_LO
    x = np.asarray(clips)
    if x.ndim == 3:
        x = x[:, 0, :]
    wp = packet_coefficients(x, DEC_LO["haar"], level, "freq", dtype)  # [B, T, P]
    sums = np.abs(wp).astype(np.float64).sum(axis=(0, 1))
    return sums, wp.shape[0] * wp.shape[1]


def rfft_fingerprint(clips):
    """Mean-spectrum fingerprint exactly as the reference chains it (fingerprints.py:51-62, use = all bins):
    rfft of every clip -> concatenate with an empty zero block -> irfft -> mean over clips -> |rfft|.
    ``clips``: [n, 1, N] (the reference's clip_array); returns (freqs, mean_abs_fft) with N // 2 + 1 entries."""
    clip_array = np.asarray(clips)
    freq_clips = np.fft.rfft(clip_array.astype(np.float64), axis=-1)
    use = freq_clips.shape[-1]
    zeros = np.zeros_like(freq_clips)[:, :, :-use]
    masked_freq = np.concatenate([zeros, freq_clips[:, :, -use:]], -1)
    masked_time_mean = np.mean(np.fft.irfft(masked_freq), 0)[0]
    mean_abs_fft = np.abs(np.fft.rfft(masked_time_mean)[-use:])
    freqs = np.fft.rfftfreq(masked_time_mean.shape[-1], 1.0 / 22050)[-use:]
    return freqs, mean_abs_fft
