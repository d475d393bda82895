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
import io
import os
import pickle
from math import log
from typing import Optional

import numpy as np
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
    """Leaf length after ``level`` analysis steps: ``floor((n + F - 1) / 2)`` iterated (== ``afd_wpt_out_len``; kept on
    the host because it sits on the per-step path of small batches)."""
    n, filt_len, level = int(n), int(filt_len), int(level)
    if n < 1 or filt_len < 2 or filt_len % 2 or level < 0:
        raise _lib.AfdError("afd_wpt_out_len", -1, "afd_wpt_out_len: bad argument")
    for _ in range(level):
        n = (n + filt_len - 1) // 2
    return n


class _SameDevice:
    """No-op stand-in for ``torch.cuda.device(dev)`` when ``dev`` already is the current device (the usual case:
    one process per GPU); the real context manager costs two driver calls per step."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_SAME_DEVICE = _SameDevice()


def _on_device(device: torch.device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    return _SAME_DEVICE if index == torch.cuda.current_device() else torch.cuda.device(device)


def _c_taps(wav):
    """``dec_lo`` as a ctypes double array, built once per wavelet object."""
    arr = getattr(wav, "_afd_c_taps", None)
    if arr is None:
        taps = [float(v) for v in wav.dec_lo]
        arr = (ctypes.c_double * len(taps))(*taps)
        try:
            wav._afd_c_taps = arr
        except AttributeError:          # e.g. pywt.Wavelet instances do not take new attributes
            pass
    return arr


def _norm_array(norm, channels: int):
    """(mean, std) scalars / per-channel sequences -> ctypes float[C][2] for the fused Normalize epilogue."""
    if norm is None:
        return None
    if isinstance(norm, HostNorm):                  # converted once by fuse_normalize: no device sync per step
        return norm.array(channels)
    mean, std = norm
    mean = torch.as_tensor(mean, dtype=torch.float32).reshape(-1).tolist()
    std = torch.as_tensor(std, dtype=torch.float32).reshape(-1).tolist()
    if len(mean) == 1:
        mean = mean * channels
    if len(std) == 1:
        std = std * channels
    if len(mean) != channels or len(std) != channels:
        raise ValueError(f"normalisation needs 1 or {channels} mean/std values, got {len(mean)}/{len(std)}")
    flat = []
    for m, sd in zip(mean, std):
        flat += [m, sd]
    return (ctypes.c_float * (2 * channels))(*flat)


class HostNorm:
    """(mean, std) of the fused Normalize epilogue as host floats, read from the device ONCE (``fuse_normalize``);
    ``get_transforms`` / ``calc_normalization`` hand out CUDA tensors, and converting them on every forward would
    be a blocking device-to-host copy per training step."""

    def __init__(self, mean, std):
        self.mean = torch.as_tensor(mean, dtype=torch.float32).reshape(-1).tolist()
        self.std = torch.as_tensor(std, dtype=torch.float32).reshape(-1).tolist()
        self._arrays = {}

    def array(self, channels: int):
        arr = self._arrays.get(channels)
        if arr is None:
            arr = _norm_array((self.mean, self.std), channels)
            self._arrays[channels] = arr
        return arr


def wavelet_packet_features(pt_data: torch.Tensor, wavelet, max_lev: int = 8, log_scale: bool = False,
                            loss_less: bool = False, power: float = 2.0, order: str = "freq",
                            log_offset: float = 1e-12, *, node_scale: Optional[torch.Tensor] = None,
                            norm=None, node_stats: Optional[torch.Tensor] = None,
                            feat_moments: Optional[torch.Tensor] = None, store: bool = True):
    """Fused packet transform; returns contiguous ``[B, C, T, P]`` (C = 2 only with log_scale and loss_less).

    Keyword-only extras map one to one onto ``afd_wpt_forward_ex`` (include/afd_b200.h): ``node_scale`` fp32
    ``[P]`` (block norm), ``norm=(mean, std)`` fused Normalize, ``node_stats`` fp64 ``[3, P]`` and
    ``feat_moments`` fp64 ``[C, 2]`` accumulators; ``store=False`` skips the feature tensor (returns ``None``).
    """
    x = _as_frames(pt_data, "wavelet_packet_features")
    wav = get_wavelet(wavelet)
    c_taps = _c_taps(wav)
    F = len(c_taps)
    B, N = x.shape
    T = wpt_out_len(N, F, max_lev)
    C = 2 if (log_scale and loss_less) else 1
    P = 1 << max_lev
    extended = node_scale is not None or norm is not None or node_stats is not None or feat_moments is not None
    if not store and node_stats is None and feat_moments is None:
        raise ValueError("store=False needs node_stats or feat_moments to accumulate into")
    out = torch.empty((B, C, T, P), dtype=torch.float32, device=x.device) if store else None
    order_id = _lib.AFD_ORDER_FREQ if order == "freq" else _lib.AFD_ORDER_NATURAL
    stride = x.stride(0) if B > 1 else N
    out_ptr = ctypes.c_void_p(out.data_ptr()) if store else None

    def _dev(t, dtype, shape, what):
        if t is None:
            return None
        if t.device != x.device or t.dtype != dtype or tuple(t.shape) != shape or not t.is_contiguous():
            raise ValueError(f"{what} must be a contiguous {dtype} tensor of shape {shape} on {x.device}")
        return ctypes.c_void_p(t.data_ptr())

    with _on_device(x.device):
        if not extended:
            rc = _lib.load().afd_wpt_forward(
                ctypes.c_void_p(x.data_ptr()), B, N, stride, c_taps, F, max_lev, order_id,
                float(power), int(bool(log_scale)), float(log_offset), int(bool(loss_less)),
                out_ptr, None, _stream_ptr(x.device))
            _lib.check("afd_wpt_forward", rc)
        else:
            rc = _lib.load().afd_wpt_forward_ex(
                ctypes.c_void_p(x.data_ptr()), B, N, stride, c_taps, F, max_lev, order_id,
                float(power), int(bool(log_scale)), float(log_offset), int(bool(loss_less)),
                _dev(node_scale, torch.float32, (P,), "node_scale"), _norm_array(norm, C),
                _dev(node_stats, torch.float64, (3, P), "node_stats"),
                _dev(feat_moments, torch.float64, (C, 2), "feat_moments"),
                out_ptr, None, _stream_ptr(x.device))
            _lib.check("afd_wpt_forward_ex", rc)
    return out


def graycode_paths(level: int) -> list[str]:
    """Leaf paths in frequency order (ptwt get_level / reference wavelet_math.py:185)."""
    order = ["a", "d"]
    for _ in range(level - 1):
        order = ["a" + p for p in order] + ["d" + p for p in order[::-1]]
    return order


class NodeStatsTable:
    """Running count / mean / M2 of every leaf node at once (fp64 ``[P]`` on the device).

    One table stands in for the 2^level ``WelfordEstimator`` objects the reference keeps in its
    ``block_norm_dict`` (wavelet_math.py:194-200, data_loader.py:27-71): batches are merged with Chan's parallel
    update from the per-node sums the kernel accumulates, which yields the same mean and ``sqrt(m2 / count)``.
    """

    def __init__(self, packets: int, device):
        self.count = 0
        self.mean = torch.zeros(packets, dtype=torch.float64, device=device)
        self.m2 = torch.zeros(packets, dtype=torch.float64, device=device)

    def merge_sums(self, s: torch.Tensor, q: torch.Tensor, n: int) -> None:
        mean_b = s / n
        m2_b = torch.clamp(q - s * mean_b, min=0)
        tot = self.count + n
        delta = mean_b - self.mean
        self.m2 += m2_b + delta * delta * (self.count * n / tot)
        self.mean += delta * (n / tot)
        self.count = tot


class NodeStats:
    """One node's view of a ``NodeStatsTable`` with the ``WelfordEstimator`` surface (count, mean, m2, finalize)."""

    def __init__(self, table: NodeStatsTable, index: int):
        self.table = table
        self.index = index

    @property
    def count(self) -> torch.Tensor:
        return torch.tensor([float(self.table.count)], device=self.table.mean.device)

    @property
    def mean(self) -> torch.Tensor:
        return self.table.mean[self.index:self.index + 1].float()

    @property
    def m2(self) -> torch.Tensor:
        return self.table.m2[self.index:self.index + 1].float()

    def finalize(self):
        return self.mean, torch.sqrt(self.table.m2[self.index:self.index + 1] / self.table.count).float()


def _stats_table(block_norm_dict: dict, level: int, device) -> NodeStatsTable:
    """The table behind ``block_norm_dict`` (created, and the per-path views filled in, on first use)."""
    for v in block_norm_dict.values():
        if isinstance(v, NodeStats):
            return v.table
        break
    table = NodeStatsTable(1 << level, device)
    for p, key in enumerate(graycode_paths(level)):
        block_norm_dict[key] = NodeStats(table, p)
    return table


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
    *,
    norm=None,
    feat_moments: Optional[torch.Tensor] = None,
    store: bool = True,
) -> tuple[torch.Tensor, dict]:
    """Create a packet image ``[B, C, T, P]`` (reference wavelet_math.py:167-220).

    The keyword-only extras are this package's additions: ``norm=(mean, std)`` fuses the reference's separate
    ``Normalize`` step into the epilogue, ``feat_moments`` (fp64 ``[C, 2]``) accumulates the feature sum / sum of
    squares that ``calc_normalization`` needs, ``store=False`` computes statistics only.

    ``compute_welford`` (per-node running mean / M2, reference :194-200) costs nothing extra: the kernel
    accumulates every node's sum, sum of squares and max |c| while it writes the features.  ``block_norm``
    (``node / max|node|`` over the batch, reference :202-203) needs the maxima before the epilogue, so it runs a
    statistics-only launch first (no feature tensor is written) and a second launch with the per-node scale.
    """
    if block_norm_dict is None:
        block_norm_dict = {}
    extras = dict(norm=norm, feat_moments=feat_moments, store=store)
    if not block_norm and not compute_welford:
        return (wavelet_packet_features(pt_data, wavelet, max_lev, log_scale, loss_less, power, **extras),
                block_norm_dict)

    x = _as_frames(pt_data, "compute_pytorch_packet_representation")
    P = 1 << max_lev
    stats = torch.zeros((3, P), dtype=torch.float64, device=x.device)
    if block_norm:
        wavelet_packet_features(x, wavelet, max_lev, False, False, power, node_stats=stats, store=False)
        scale = (1.0 / stats[2]).to(torch.float32)
        out = wavelet_packet_features(x, wavelet, max_lev, log_scale, loss_less, power, node_scale=scale, **extras)
    else:
        out = wavelet_packet_features(x, wavelet, max_lev, log_scale, loss_less, power, node_stats=stats, **extras)
    if compute_welford:
        n = x.shape[0] * wpt_out_len(x.shape[1], len(get_wavelet(wavelet).dec_lo), max_lev)
        _stats_table(block_norm_dict, max_lev, x.device).merge_sums(stats[0], stats[1], n)
    return out, block_norm_dict


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
        self.fused_norm = None       # (mean, std): apply the reference's Normalize inside the kernel (fuse_normalize)
        self.feat_moments = None     # fp64 [C, 2] accumulator filled by calc_normalization
        self.store = True

    @property
    def channels(self) -> int:
        return 2 if (self.log_scale and self.loss_less) else 1

    def forward(self, pt_data: torch.Tensor) -> tuple[torch.Tensor, dict]:
        packets, block_norm_dict = compute_pytorch_packet_representation(
            pt_data, self.wavelet, self.max_lev, self.log_scale, self.loss_less, self.power,
            block_norm=self.block_norm, compute_welford=self.compute_welford,
            block_norm_dict=self.block_norm_dict,
            norm=self.fused_norm, feat_moments=self.feat_moments, store=self.store,
        )
        if packets is None:
            return None, block_norm_dict
        return packets.permute(0, 1, 3, 2), block_norm_dict


def stft_out_shape(n: int, n_fft: int, hop: int) -> tuple[int, int]:
    frames, bins = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check("afd_stft_out_shape",
               _lib.load().afd_stft_out_shape(n, n_fft, hop, ctypes.byref(frames), ctypes.byref(bins)))
    return frames.value, bins.value


def stft_power_features(x: torch.Tensor, n_fft: int = 511, hop_length: int = 220, power: float = 2.0,
                        log_scale: bool = False, log_offset: float = 1e-12, *, norm=None,
                        feat_moments: Optional[torch.Tensor] = None, store: bool = True):
    """Fused STFT power spectrogram; returns contiguous ``[B, 1, frames, bins]``.

    ``norm`` / ``feat_moments`` (fp64 ``[1, 2]``) / ``store`` as in ``wavelet_packet_features``
    (``afd_stft_power_ex``)."""
    xf = _as_frames(x, "stft_power_features")
    B, N = xf.shape
    frames, bins = stft_out_shape(N, n_fft, hop_length)
    if not store and feat_moments is None:
        raise ValueError("store=False needs feat_moments to accumulate into")
    out = torch.empty((B, 1, frames, bins), dtype=torch.float32, device=xf.device) if store else None
    out_ptr = ctypes.c_void_p(out.data_ptr()) if store else None
    stride = xf.stride(0) if B > 1 else N
    with _on_device(xf.device):
        if norm is None and feat_moments is None:
            rc = _lib.load().afd_stft_power(
                ctypes.c_void_p(xf.data_ptr()), B, N, stride, n_fft, hop_length, float(power),
                int(bool(log_scale)), float(log_offset), out_ptr, _stream_ptr(xf.device))
            _lib.check("afd_stft_power", rc)
        else:
            mom = None
            if feat_moments is not None:
                if (feat_moments.device != xf.device or feat_moments.dtype != torch.float64
                        or feat_moments.numel() != 2 or not feat_moments.is_contiguous()):
                    raise ValueError(f"feat_moments must be a contiguous float64 tensor of 2 elements on {xf.device}")
                mom = ctypes.c_void_p(feat_moments.data_ptr())
            rc = _lib.load().afd_stft_power_ex(
                ctypes.c_void_p(xf.data_ptr()), B, N, stride, n_fft, hop_length, float(power),
                int(bool(log_scale)), float(log_offset), _norm_array(norm, 1), mom, out_ptr, _stream_ptr(xf.device))
            _lib.check("afd_stft_power_ex", rc)
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
        self.fused_norm = None
        self.feat_moments = None
        self.store = True
        self.channels = 1

    def forward(self, input: torch.Tensor) -> tuple[torch.Tensor, None]:
        spec = stft_power_features(input, self.n_fft, self.hop_length, self.power, self.log_scale, 1e-12,
                                   norm=self.fused_norm, feat_moments=self.feat_moments, store=self.store)
        if spec is None:
            return None, None
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


def _opt(args, name: str, default):
    """Optional flag of the reference's DotDict (utils.py:321-395: attribute access is dict lookup, so a missing
    key may raise KeyError instead of AttributeError)."""
    try:
        value = getattr(args, name)
    except (AttributeError, KeyError):
        return default
    return default if value is None else value


class _ArrayUnpickler(pickle.Unpickler):
    """The reference caches ``[mean, std]`` as a pickled list of numpy arrays (wavelet_math.py:449-450); nothing but
    numpy array reconstruction is allowed to run while reading such a file back."""

    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
                ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"),
                ("numpy._core.multiarray", "scalar")}

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"unexpected object {module}.{name} in a mean/std cache file")


def norm_cache_prefix(args) -> Optional[str]:
    """``norm_dir`` of the reference (wavelet_math.py:327-347): the cache file is ``f"{prefix}_mean_std.pkl"``.
    ``None`` when ``args`` does not carry the flags the name is built from (e.g. a bare benchmark config)."""
    try:
        loss_less = "_loss_less" if args.loss_less == "True" else ""
        return (args.log_dir + "/norms/" + args.data_path.replace("/", "_") + "_" + "-".join(args.only_use) + "_"
                + args.transform + "_" + args.wavelet + "_" + str(args.num_of_scales) + "_" + str(args.power)
                + loss_less + "_" + str(args.sample_rate) + "_" + str(args.seconds) + "secs")
    except (AttributeError, KeyError, TypeError):
        return None


def _training_batches(args):
    """The batches the reference's ``calc_normalization`` iterates (wavelet_math.py:406-431).  Dataset and wav I/O are
    the reference's own (out of scope here): an iterable / DataLoader in ``args.norm_batches`` is used as is,
    otherwise the reference's ``get_costum_dataset`` is imported when this module runs inside the reference tree."""
    batches = _opt(args, "norm_batches", None)
    if batches is not None:
        return batches
    get_costum_dataset = None
    for mod in ("audiofakedetect.data_loader", "src.audiofakedetect.data_loader"):
        try:
            get_costum_dataset = __import__(mod, fromlist=["get_costum_dataset"]).get_costum_dataset
            break
        except Exception:      # noqa: BLE001  (the reference package needs ptwt / pywt / librosa at import time)
            continue
    if get_costum_dataset is None:
        raise RuntimeError(
            "get_transforms(normalization=True): no cached '<norm_dir>_mean_std.pkl', no args.norm_batches, and the "
            "reference's audiofakedetect.data_loader (which builds the training set) is not importable here")
    asv = _opt(args, "asvspoof_name", None)
    dataset = get_costum_dataset(
        data_path=args.data_path, ds_type="train", only_use=args.only_use, save_path=args.save_path,
        limit=args.limit_train[0], asvspoof_name=(f"{asv}_T" if asv is not None and "LA" in asv else asv),
        file_type=args.file_type, resample_rate=args.sample_rate, seconds=args.seconds)
    return torch.utils.data.DataLoader(dataset, batch_size=4000, shuffle=False, pin_memory=True,
                                       num_workers=_opt(args, "num_workers", 0))


def get_transforms(args, features: str, device: str, normalization: bool, pbar: bool = False,
                   verbose: bool = True, norm_batches=None) -> tuple[torch.nn.Sequential, torch.nn.Sequential]:
    """Initialize transformations and normalize (reference wavelet_math.py:266-384, same positional signature).

    ``args`` needs the reference's flag names: transform, num_of_scales, hop_length, log_scale, power, wavelet,
    loss_less ("True"/"False" strings), features, block_norm, mean, std.  Normalisation statistics, in the
    reference's order (:349-371): (1) a cached ``<norm_dir>_mean_std.pkl`` (the reference's own file format and
    name) is loaded when it exists; (2) with ``normalization=True`` they are computed by ``calc_normalization`` --
    statistics-only launches over ``norm_batches`` / ``args.norm_batches`` or, inside the reference tree, over the
    reference's training DataLoader -- and cached under the same name; (3) otherwise ``args.mean`` / ``args.std``.
    """
    if features not in ("none", None):
        raise NotImplementedError(
            f"features={features!r}: the reference appends LFCC / ComputeDeltas modules here (wavelet_math.py:316-323); "
            "they are outside the accelerated hot path (SURVEY.md section 2.1 row 5) -- use features='none'")
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
            # the reference hard-codes True (:304): aux is the dict of 2^level per-node estimators.  Here the
            # statistics are a side reduction of the one launch; args.compute_welford=False opts out.
            compute_welford=bool(_opt(args, "compute_welford", True)),
        )
    else:
        raise ValueError(f"unknown transform '{args.transform}'")
    transforms = torch.nn.Sequential(transform)
    block_norm = bool(_opt(args, "block_norm", False))
    welford_dict = None
    prefix = norm_cache_prefix(args)
    if prefix is not None and not block_norm and os.path.exists(f"{prefix}_mean_std.pkl"):      # reference :349-355
        if verbose:
            print("Loading pre calculated mean and std from file.")
        with open(f"{prefix}_mean_std.pkl", "rb") as file:
            mean, std = _ArrayUnpickler(io.BytesIO(file.read())).load()
        mean = torch.from_numpy(np.asarray(mean).astype(np.float32)).to(device)
        std = torch.from_numpy(np.asarray(std).astype(np.float32)).to(device)
    elif normalization:                                                                          # reference :364-367
        if verbose:
            print("computing mean and std values.", flush=True)
        batches = norm_batches if norm_batches is not None else _training_batches(args)
        welford_dict, mean, std = calc_normalization(transforms, batches)
        if prefix is not None and not block_norm:                                                # reference :449-450
            os.makedirs(os.path.dirname(prefix), exist_ok=True)
            with open(f"{prefix}_mean_std.pkl", "wb") as file:
                pickle.dump([mean.cpu().numpy(), std.cpu().numpy()], file)
    else:
        if verbose:
            print("Using default mean and std.")
        mean = torch.as_tensor(args.mean, dtype=torch.float32, device=device)
        std = torch.as_tensor(args.std, dtype=torch.float32, device=device)
    if block_norm:                                        # reference :373-378
        if args.transform != "packets":
            raise ValueError("block_norm is a packet option")
        mean, std = 0.0, 1.0
        transform.block_norm_dict = welford_dict
        transform.compute_welford = False
        transform.block_norm = True
    normalize = torch.nn.Sequential(Normalize(mean, std))
    return transforms, normalize


def merge_moments_across_ranks(moments: torch.Tensor, count: int, group=None) -> tuple[torch.Tensor, int]:
    """Sum the per-rank feature moments ``[C, 2]`` (sum, sum of squares) and element counts over the ranks of
    ``torch.distributed`` (one all-reduce of 2C + 1 fp64): every rank gets the statistics of the whole sharded set."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return moments, count
    packed = torch.cat([moments.reshape(-1), torch.tensor([float(count)], dtype=torch.float64, device=moments.device)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed[:-1].reshape(moments.shape), int(round(float(packed[-1])))


def calc_normalization(transforms: torch.nn.Sequential, audio_batches, distributed: bool = False, group=None) -> tuple:
    """Mean / std of the features over ``audio_batches`` (reference wavelet_math.py:387-452, minus the dataset
    plumbing): returns ``(welford_dict, mean, std)`` like the reference, mean / std per channel.

    One statistics-only launch per batch: the kernels accumulate the feature sum / sum of squares in fp64 on the
    device (``feat_moments``) and never write the feature tensor, so the pass reads 88 KB per frame and writes
    nothing.  The reference runs this pass over the WHOLE training set on every rank; with ``distributed=True`` each
    rank passes its own shard of the batches and the moments meet in one all-reduce."""
    tr = transforms[0]
    dev = None
    moments, count, welford_dict = None, 0, None
    saved = (tr.feat_moments, tr.store, tr.fused_norm)
    try:
        for batch in audio_batches:
            x = batch["audio"] if isinstance(batch, dict) else batch
            if not x.is_cuda:
                x = x.cuda(non_blocking=True)
            if moments is None:
                dev = x.device
                moments = torch.zeros((tr.channels, 2), dtype=torch.float64, device=dev)
            tr.feat_moments, tr.store, tr.fused_norm = moments, False, None
            _, welford_dict = tr(x)
            if isinstance(tr, Packets):
                tr.block_norm_dict = welford_dict             # reference :439
                frames = x.shape[0] if x.dim() > 1 else 1
                count += frames * wpt_out_len(x.shape[-1], len(tr.wavelet.dec_lo), tr.max_lev) * (1 << tr.max_lev)
            else:
                frames = x.shape[0] if x.dim() > 1 else 1
                t, bins = stft_out_shape(x.shape[-1], tr.n_fft, tr.hop_length)
                count += frames * t * bins
    finally:
        tr.feat_moments, tr.store, tr.fused_norm = saved
    if moments is None:
        raise ValueError("calc_normalization: no batches")
    if distributed:
        moments, count = merge_moments_across_ranks(moments, count, group)
    mean = moments[:, 0] / count
    std = torch.sqrt(torch.clamp(moments[:, 1] / count - mean * mean, min=0))
    return welford_dict, mean.float(), std.float()


def fuse_normalize(transforms: torch.nn.Sequential, normalize: torch.nn.Sequential):
    """Fold ``normalize`` (the reference's separate Normalize module, :380-382) into the transform kernel's
    epilogue: returns ``(transforms, identity)`` so callers keep the two-step call pattern
    (train_classifier.py:965-967) while the features touch HBM once."""
    norm = normalize[0]
    transforms[0].fused_norm = HostNorm(norm.mean, norm.std)
    return transforms, torch.nn.Sequential(torch.nn.Identity())


def normalization_stats(feature_batches) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-channel mean / std over an iterable of feature tensors ``[B, C, P, T]`` that already exist
    (``calc_normalization`` is the fused route from audio)."""
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
