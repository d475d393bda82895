"""torch restatement of the reference's CPU packet path, structured like ptwt (TEST INFRASTRUCTURE ONLY).

This is the *CPU baseline* (``bench.py --impl reference`` and the ``cpu_baseline`` leg) and a second,
independently written checker for the numpy oracle.  It follows the call structure the reference executes on
CPU -- one ``F.pad(reflect)`` + one ``F.conv1d(stride=2)`` per tree node, a python loop over the 2^L leaves,
``torch.stack`` and the log epilogue -- so that its run time is representative of
``ptwt.WaveletPacket`` + ``compute_pytorch_packet_representation`` (reference wavelet_math.py:167-220).
ptwt / pywt themselves are third-party, un-pinned (reference requirements.txt:4-5) and not installable in this
image; when they ARE importable ``have_ptwt()`` is true and bench.py / the tests use the real thing instead.

STFT: ``torchaudio.transforms.Spectrogram`` is the reference's own code path (wavelet_math.py:47) and is used
directly; ``stft_power_explicit`` restates it through ``torch.stft`` for boxes without torchaudio.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .filters import DEC_LO, dec_hi


def have_ptwt() -> bool:
    try:
        import ptwt  # noqa: F401
        import pywt  # noqa: F401
        return True
    except Exception:
        return False


class PacketTree:
    """Analysis tree with ptwt.WaveletPacket's surface: ``tree[path]`` and ``get_level(level)``.

    ``lazy=False`` computes every node down to ``max_level`` in the constructor (ptwt <= 0.1.x behaviour:
    the whole tree is built eagerly); ``lazy=True`` computes nodes on first access.
    """

    def __init__(self, data: torch.Tensor, dec_lo, mode: str = "reflect", max_level: int | None = None,
                 lazy: bool = True):
        if mode != "reflect":
            raise ValueError("the reference only uses mode='reflect' (wavelet_math.py:182)")
        if data.dim() == 3:            # old ptwt squeezes [B, 1, N] to [B, N] nodes
            data = data[:, 0, :]
        lo = torch.as_tensor(dec_lo, dtype=torch.float64)
        hi = torch.as_tensor(dec_hi(dec_lo), dtype=torch.float64)
        # conv1d correlates: ptwt flips the decomposition filters
        self.bank = torch.stack([lo.flip(0), hi.flip(0)]).unsqueeze(1).to(data.dtype)   # [2, 1, F]
        self.flen = lo.numel()
        self.nodes = {"": data}
        self.max_level = max_level
        if not lazy:
            if max_level is None:
                raise ValueError("eager construction needs max_level")
            frontier = [""]
            for _ in range(max_level):
                frontier = [c for p in frontier for c in self._split(p)]

    def _split(self, path: str):
        x = self.nodes[path]
        n = x.shape[-1]
        padl = self.flen - 2
        padr = self.flen - 2 + (n % 2)
        xp = F.pad(x.unsqueeze(1), [padl, padr], mode="reflect") if (padl or padr) else x.unsqueeze(1)
        y = F.conv1d(xp, self.bank, stride=2)                      # [B, 2, (n + F - 1) // 2]
        self.nodes[path + "a"], self.nodes[path + "d"] = y[:, 0], y[:, 1]
        return path + "a", path + "d"

    def __getitem__(self, path: str) -> torch.Tensor:
        if path not in self.nodes:
            self[path[:-1]]
            self._split(path[:-1])
        return self.nodes[path]

    @staticmethod
    def get_level(level: int, order: str = "freq"):
        if order == "freq":
            paths = ["a", "d"]
            for _ in range(level - 1):
                paths = ["a" + p for p in paths] + ["d" + p for p in paths[::-1]]
            return paths
        paths = [""]
        for _ in range(level):
            paths = [p + c for p in paths for c in "ad"]
        return paths


class Welford:
    """Running mean / M2 over all but the last axis (restates reference data_loader.py:27-71)."""

    def __init__(self):
        self.count = None

    def update(self, vals: torch.Tensor) -> None:
        if self.count is None:
            self.axes = tuple(range(vals.dim() - 1))
            self.count = torch.zeros(1, dtype=torch.float32)
            self.mean = torch.zeros(vals.shape[-1], dtype=torch.float32)
            self.m2 = torch.zeros(vals.shape[-1], dtype=torch.float32)
        self.count += torch.prod(torch.tensor(vals.shape[:-1]))
        d1 = vals - self.mean
        self.mean += torch.sum(d1 / self.count, self.axes)
        self.m2 += torch.sum(d1 * (vals - self.mean), self.axes)

    def finalize(self):
        return self.mean, torch.sqrt(self.m2 / self.count)


def packet_representation(pt_data: torch.Tensor, dec_lo, max_lev: int = 8, log_scale: bool = False,
                          loss_less: bool = False, power: float = 2.0, block_norm: bool = False,
                          compute_welford: bool = False, block_norm_dict=None, order: str = "freq"):
    """compute_pytorch_packet_representation (reference wavelet_math.py:167-220) on CPU tensors."""
    tree = PacketTree(pt_data, dec_lo, mode="reflect")
    stats = {} if block_norm_dict is None else block_norm_dict
    leaves = []
    for path in tree.get_level(max_lev, order):
        node = tree[path]
        if compute_welford:
            stats.setdefault(path, Welford()).update(node.unsqueeze(-1))
        if block_norm:
            node = node / torch.max(torch.abs(node))
        leaves.append(node)
    wp = torch.stack(leaves, dim=-1)                                  # [B, T, P]
    if log_scale:
        wp_log = torch.log(torch.abs(wp).pow(power) + 1e-12)
        if loss_less:
            sign = ((wp < 0).to(wp.dtype) * (-1) + 0.5) * 2
            return torch.stack([wp_log, sign], 1), stats
        return wp_log.unsqueeze(1), stats
    return wp.unsqueeze(1), stats


def packets_forward(pt_data, wavelet_name: str, max_lev: int = 8, **kw):
    """Packets.forward (reference wavelet_math.py:249-263): logical [B, C, P, T] view."""
    rep, stats = packet_representation(pt_data, DEC_LO[wavelet_name], max_lev, **kw)
    return rep.permute(0, 1, 3, 2), stats


def ptwt_packets_forward(pt_data, wavelet_name: str, max_lev: int = 8, log_scale=True, power=2.0):
    """The real thing, when ptwt + pywt are importable (never in this image)."""
    import ptwt
    import pywt

    tree = ptwt.WaveletPacket(data=pt_data, wavelet=pywt.Wavelet(wavelet_name), mode="reflect")
    wp = torch.stack([tree[n] for n in tree.get_level(max_lev)], dim=-1)
    if log_scale:
        wp = torch.log(torch.abs(wp).pow(power) + 1e-12)
    return wp.unsqueeze(1).permute(0, 1, 3, 2)


def stft_power_explicit(x: torch.Tensor, n_fft: int = 511, hop_length: int = 220, power: float = 2.0,
                        log_scale: bool = False) -> torch.Tensor:
    """STFTLayer.forward (reference wavelet_math.py:56-68) through torch.stft; returns [B, 1, bins, frames]."""
    shape = x.shape
    flat = x.reshape(-1, shape[-1])
    spec = torch.stft(flat, n_fft, hop_length=hop_length, win_length=n_fft,
                      window=torch.hann_window(n_fft, dtype=x.dtype), center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    spec = spec.reshape(shape[:-1] + spec.shape[-2:]).abs().pow(power)
    if x.dim() == 2:
        spec = spec.unsqueeze(1)
    return torch.log(spec + 1e-12) if log_scale else spec


def stft_layer_forward(x: torch.Tensor, n_fft: int = 511, hop_length: int = 220, power: float = 2.0,
                       log_scale: bool = False) -> torch.Tensor:
    """The reference's own STFT code path: torchaudio Spectrogram (+ log)."""
    try:
        from torchaudio.transforms import Spectrogram
    except Exception:
        return stft_power_explicit(x, n_fft, hop_length, power, log_scale)
    spec = Spectrogram(n_fft=n_fft, hop_length=hop_length, power=power)(x)
    if x.dim() == 2:
        spec = spec.unsqueeze(1)
    return torch.log(spec + 1e-12) if log_scale else spec


def stft_power_dft64(x, n_fft: int = 511, hop_length: int = 220):
    """fp64 truth: explicit framing and a direct DFT matrix product.  x: numpy [B, N] -> [B, bins, frames]."""
    import numpy as np

    x = np.asarray(x, dtype=np.float64)
    pad = n_fft // 2
    xp = np.pad(x, [(0, 0), (pad, pad)], mode="reflect")
    frames = 1 + (x.shape[-1] + 2 * pad - n_fft) // hop_length      # torch.stft(center=True)
    idx = np.arange(frames)[:, None] * hop_length + np.arange(n_fft)[None, :]
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)
    seg = xp[:, idx] * win                                           # [B, frames, n_fft]
    k = np.arange(n_fft // 2 + 1)
    ang = -2 * np.pi * ((np.arange(n_fft)[:, None] * k[None, :]) % n_fft) / n_fft
    spec = seg @ np.cos(ang) + 1j * (seg @ np.sin(ang))
    return (np.abs(spec) ** 2).transpose(0, 2, 1)


def haar_fingerprint(clips: torch.Tensor, level: int = 14):
    """_compute_fingerprint_wpt's arithmetic (reference fingerprints.py:99-115) with the torch tree:
    mean |c| over clips, channel and positions for every frequency-ordered Haar packet -> [2^level]."""
    tree = PacketTree(clips, DEC_LO["haar"], mode="reflect")
    leaves = [tree[p] for p in tree.get_level(level, "freq")]
    packets = torch.stack(leaves, -1)
    return packets.abs().double().mean(dim=tuple(range(packets.dim() - 1)))
