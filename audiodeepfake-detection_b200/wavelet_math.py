"""Reference-facing transform modules, backed by hand-written sm_100a kernels.

Mirrors the public surface of the reference's ``src/audiofakedetect/wavelet_math.py`` for the feature
front-end -- same names, argument meaning, return conventions and error behaviour:

  * ``STFTLayer``                              reference wavelet_math.py:25-68
  * ``compute_pytorch_packet_representation``  reference wavelet_math.py:167-220
  * ``Packets``                                reference wavelet_math.py:223-263
  * ``get_transforms``                         reference wavelet_math.py:266-384

Every transform returns ``(features, aux)``; features are logical ``[B, C, P, T]`` views of contiguous
``[B, C, T, P]`` memory (the strides the reference produces), so the DCNN's first ``permute`` is free.
Inputs must be CUDA fp32 tensors; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from math import log
from typing import Optional

import torch

from . import _lib
from .wavelets import get_wavelet


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _as_frames(x: torch.Tensor, what: str) -> torch.Tensor:
    """[B, 1, N] / [B, N] / [N] fp32 CUDA -> contiguous [B, N]."""
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(x)!r}")
    if not x.is_cuda:
        raise RuntimeError(f"{what}: input must live on a CUDA device (B200); this transform has no CPU path")
    if x.dtype != torch.float32:
        raise TypeError(f"{what}: input must be float32, got {x.dtype}")
    if x.dim() == 3:
        if x.shape[1] != 1:
            raise ValueError(f"{what}: expected a single audio channel, got shape {tuple(x.shape)}")
        x = x[:, 0, :]
    elif x.dim() == 1:
        x = x.unsqueeze(0)
    elif x.dim() != 2:
        raise ValueError(f"{what}: expected [B, 1, N] or [B, N], got shape {tuple(x.shape)}")
    if x.stride(-1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def wpt_out_len(n: int, filt_len: int, level: int) -> int:
    out = ctypes.c_int64(0)
    _lib.check("afd_wpt_out_len", _lib.load().afd_wpt_out_len(n, filt_len, level, ctypes.byref(out)))
    return out.value


def wavelet_packet_features(pt_data: torch.Tensor, wavelet, max_lev: int = 8, log_scale: bool = False,
                            loss_less: bool = False, power: float = 2.0, order: str = "freq",
                            log_offset: float = 1e-12) -> torch.Tensor:
    """Fused packet transform; returns contiguous ``[B, C, T, P]`` (C = 2 only with log_scale and loss_less)."""
    x = _as_frames(pt_data, "wavelet_packet_features")
    wav = get_wavelet(wavelet)
    taps = [float(v) for v in wav.dec_lo]
    F = len(taps)
    B, N = x.shape
    T = wpt_out_len(N, F, max_lev)
    C = 2 if (log_scale and loss_less) else 1
    P = 1 << max_lev
    out = torch.empty((B, C, T, P), dtype=torch.float32, device=x.device)
    c_taps = (ctypes.c_double * F)(*taps)
    with torch.cuda.device(x.device):
        rc = _lib.load().afd_wpt_forward(
            ctypes.c_void_p(x.data_ptr()), B, N, x.stride(0) if B > 1 else N, c_taps, F, max_lev,
            _lib.AFD_ORDER_FREQ if order == "freq" else _lib.AFD_ORDER_NATURAL,
            float(power), int(bool(log_scale)), float(log_offset), int(bool(loss_less)),
            ctypes.c_void_p(out.data_ptr()), None, _stream_ptr(x.device))
    _lib.check("afd_wpt_forward", rc)
    return out


def compute_pytorch_packet_representation(
    pt_data: torch.Tensor,
    wavelet,
    max_lev: int = 8,
    log_scale: bool = False,
    loss_less: bool = False,
    power: float = 2.0,
    block_norm: bool = False,
    compute_welford: bool = False,
    block_norm_dict=None,
) -> tuple[torch.Tensor, dict]:
    """Create a packet image ``[B, C, T, P]`` (reference wavelet_math.py:167-220).

    ``block_norm`` (per-node division by the batch-wide max |c|, reference :202-203) and the per-node Welford
    statistics (reference :194-200) are evaluated on the raw coefficients when requested; the default training
    path discards both (``_`` at reference train_classifier.py:966), so the fused kernel is the fast path.
    """
    if block_norm_dict is None:
        block_norm_dict = {}
    if not block_norm and not compute_welford:
        return wavelet_packet_features(pt_data, wavelet, max_lev, log_scale, loss_less, power), block_norm_dict

    # Options that need the raw per-node coefficients: one fused launch for the coefficients, then the
    # statistics / scaling as whole-tensor ops (256 nodes at once instead of the reference's python node loop).
    raw = wavelet_packet_features(pt_data, wavelet, max_lev, False, False, power)[:, 0]  # [B, T, P]
    if compute_welford:
        _update_node_stats(block_norm_dict, raw, max_lev)
    if block_norm:
        raw = raw / raw.abs().amax(dim=(0, 1), keepdim=True)
    if log_scale:
        wp_log = torch.log(raw.abs().pow(power) + 1e-12)
        if loss_less:
            sign = ((raw < 0).to(torch.float32) * (-1) + 0.5) * 2
            return torch.stack([wp_log, sign], 1), block_norm_dict
        return wp_log.unsqueeze(1), block_norm_dict
    return raw.unsqueeze(1), block_norm_dict


def graycode_paths(level: int) -> list[str]:
    """Leaf paths in frequency order (ptwt get_level / reference wavelet_math.py:185)."""
    order = ["a", "d"]
    for _ in range(level - 1):
        order = ["a" + p for p in order] + ["d" + p for p in order[::-1]]
    return order


class NodeStats:
    """Per-node running mean / M2 with the ``WelfordEstimator`` surface (reference data_loader.py:27-71)."""

    def __init__(self):
        self.count = None
        self.mean = None
        self.m2 = None

    def finalize(self):
        return self.mean, torch.sqrt(self.m2 / self.count)


def _update_node_stats(stats: dict, raw: torch.Tensor, level: int) -> None:
    """Chan's parallel update of every node's (count, mean, M2) in three tensor ops."""
    n_b = raw.shape[0] * raw.shape[1]
    mean_b = raw.mean(dim=(0, 1))
    m2_b = ((raw - mean_b) ** 2).sum(dim=(0, 1))
    for p, key in enumerate(graycode_paths(level)):
        st = stats.get(key)
        if st is None:
            st = stats[key] = NodeStats()
            st.count = torch.zeros(1, device=raw.device)
            st.mean = torch.zeros(1, device=raw.device)
            st.m2 = torch.zeros(1, device=raw.device)
        tot = st.count + n_b
        delta = mean_b[p:p + 1] - st.mean
        st.m2 = st.m2 + m2_b[p:p + 1] + delta * delta * st.count * n_b / tot
        st.mean = st.mean + delta * n_b / tot
        st.count = tot


class Packets(torch.nn.Module):
    """Compute wavelet packet representation as module (reference wavelet_math.py:223-263)."""

    def __init__(
        self,
        wavelet_str: str = "sym8",
        max_lev: int = 8,
        log_scale: bool = False,
        loss_less: bool = False,
        power: float = 2.0,
        block_norm: bool = False,
        compute_welford: bool = False,
        block_norm_dict=None,
    ) -> None:
        super().__init__()
        self.wavelet = get_wavelet(wavelet_str)
        self.max_lev = max_lev
        self.log_scale = log_scale
        self.loss_less = loss_less
        self.power = power
        self.block_norm = block_norm
        self.compute_welford = compute_welford
        self.block_norm_dict = block_norm_dict

    def forward(self, pt_data: torch.Tensor) -> tuple[torch.Tensor, dict]:
        packets, block_norm_dict = compute_pytorch_packet_representation(
            pt_data, self.wavelet, self.max_lev, self.log_scale, self.loss_less, self.power,
            block_norm=self.block_norm, compute_welford=self.compute_welford,
            block_norm_dict=self.block_norm_dict,
        )
        return packets.permute(0, 1, 3, 2), block_norm_dict


def stft_out_shape(n: int, n_fft: int, hop: int) -> tuple[int, int]:
    frames, bins = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check("afd_stft_out_shape",
               _lib.load().afd_stft_out_shape(n, n_fft, hop, ctypes.byref(frames), ctypes.byref(bins)))
    return frames.value, bins.value


def stft_power_features(x: torch.Tensor, n_fft: int = 511, hop_length: int = 220, power: float = 2.0,
                        log_scale: bool = False, log_offset: float = 1e-12) -> torch.Tensor:
    """Fused STFT power spectrogram; returns contiguous ``[B, 1, frames, bins]``."""
    xf = _as_frames(x, "stft_power_features")
    B, N = xf.shape
    frames, bins = stft_out_shape(N, n_fft, hop_length)
    out = torch.empty((B, 1, frames, bins), dtype=torch.float32, device=xf.device)
    with torch.cuda.device(xf.device):
        rc = _lib.load().afd_stft_power(
            ctypes.c_void_p(xf.data_ptr()), B, N, xf.stride(0) if B > 1 else N, n_fft, hop_length, float(power),
            int(bool(log_scale)), float(log_offset), ctypes.c_void_p(out.data_ptr()), _stream_ptr(xf.device))
    _lib.check("afd_stft_power", rc)
    return out


class STFTLayer(torch.nn.Module):
    """STFT power-spectrogram module (reference wavelet_math.py:25-68)."""

    def __init__(self, n_fft: int = 511, hop_length: int = 220, log_offset: float = 1e-12,
                 log_scale: bool = False, power: float = 2.0):
        super().__init__()
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.power = power
        self.log_scale = log_scale
        self.log_offset = log_offset     # stored but, like the reference (:66), the literal 1e-12 is applied
        self.block_norm_dict = None

    def forward(self, input: torch.Tensor) -> tuple[torch.Tensor, None]:
        spec = stft_power_features(input, self.n_fft, self.hop_length, self.power, self.log_scale, 1e-12)
        return spec.permute(0, 1, 3, 2), None


class Normalize(torch.nn.Module):
    """``torchvision.transforms.Normalize(mean, std)`` for scalar / per-channel stats (reference :380-382)."""

    def __init__(self, mean, std):
        super().__init__()
        self.mean = mean
        self.std = std

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        mean = torch.as_tensor(self.mean, dtype=x.dtype, device=x.device).reshape(-1, 1, 1)
        std = torch.as_tensor(self.std, dtype=x.dtype, device=x.device).reshape(-1, 1, 1)
        return (x - mean) / std


def get_transforms(args, features: str, device: str, normalization: bool, pbar: bool = False,
                   verbose: bool = True) -> tuple[torch.nn.Sequential, torch.nn.Sequential]:
    """Initialize transformations and normalize (reference wavelet_math.py:266-384).

    ``args`` needs the reference's flag names: transform, num_of_scales, hop_length, log_scale, power, wavelet,
    loss_less ("True"/"False" strings), features, block_norm, mean, std.  Normalisation statistics are taken
    from ``args.mean`` / ``args.std`` (the reference's dataset pass, calc_normalization, is out of scope here;
    see ``normalization_stats`` for the streaming equivalent on tensors).
    """
    if features not in ("none", None):
        raise NotImplementedError("only features='none' is on the accelerated path (lfcc/delta are out of scope)")
    if args.transform == "stft":
        transform = STFTLayer(
            n_fft=args.num_of_scales * 2 - 1,
            hop_length=args.hop_length,
            log_scale=args.features == "none" and args.log_scale,
            power=args.power,
        )
    elif args.transform == "packets":
        transform = Packets(
            wavelet_str=args.wavelet,
            max_lev=int(log(args.num_of_scales, 2)),
            log_scale=args.features == "none" and args.log_scale,
            loss_less=False if args.loss_less == "False" else True,
            power=args.power,
            block_norm_dict=None,
            block_norm=False,
            compute_welford=False,   # reference hard-codes True (:304) and discards the result
        )
    else:
        raise ValueError(f"unknown transform '{args.transform}'")
    transforms = torch.nn.Sequential(transform)
    if getattr(args, "block_norm", False):
        raise NotImplementedError("block_norm needs the reference's dataset pass (calc_normalization)")
    mean = torch.as_tensor(args.mean, dtype=torch.float32, device=device)
    std = torch.as_tensor(args.std, dtype=torch.float32, device=device)
    normalize = torch.nn.Sequential(Normalize(mean, std))
    return transforms, normalize


def normalization_stats(feature_batches) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-channel mean / std over an iterable of feature tensors ``[B, C, P, T]`` -- the quantity
    ``calc_normalization`` (reference wavelet_math.py:387-452) obtains with a WelfordEstimator."""
    count, s, ss = 0, None, None
    for f in feature_batches:
        c = f.shape[1]
        v = f.transpose(0, 1).reshape(c, -1).double()
        s = v.sum(1) if s is None else s + v.sum(1)
        ss = (v * v).sum(1) if ss is None else ss + (v * v).sum(1)
        count += v.shape[1]
    mean = s / count
    std = torch.sqrt(torch.clamp(ss / count - mean * mean, min=0))
    return mean.float(), std.float()
