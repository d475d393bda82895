"""Haar wavelet-packet "generator fingerprint" (reference scripts/freq_visual/fingerprints.py:85-125).

The reference stacks <= 2500 one-second clips, builds a level-14 Haar packet tree with pywt on the CPU, orders
the 16384 leaves by frequency and takes ``np.mean(np.abs(packets), (0, 1, 2))``.  Here every GPU accumulates
``sum |c|`` per packet with one fused kernel (libafd_b200 ``afd_haar_fingerprint_accum``: in-place tree in
shared memory, register accumulators, one fp64 atomic per packet and CTA), and -- when the clips are sharded
over several ranks -- one NCCL all-reduce of the 16384 sums plus the term count finishes the mean.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional

import torch

from . import _lib
from .wavelet_math import _as_frames, _stream_ptr

SAMPLE_RATE = 22050      # reference fingerprints.py:27


class FingerprintAccumulator:
    """Streaming accumulator: ``update(clips)`` any number of times, then ``mean()``.

    State lives on the device as one fp64 buffer ``[2^level + 1]`` (sums, then the term count) so that a single
    all-reduce covers both.
    """

    def __init__(self, level: int = 14, device: torch.device | str = "cuda"):
        self.level = int(level)
        self.packets = 1 << self.level
        self.device = torch.device(device)
        self.sums = torch.zeros(self.packets, dtype=torch.float64, device=self.device)
        self.count = torch.zeros(1, dtype=torch.int64, device=self.device)

    def update(self, clips: torch.Tensor) -> "FingerprintAccumulator":
        x = _as_frames(clips, "FingerprintAccumulator.update")
        if x.device != self.sums.device:
            raise RuntimeError(f"clips live on {x.device}, accumulator on {self.sums.device}")
        B, N = x.shape
        with torch.cuda.device(x.device):
            rc = _lib.load().afd_haar_fingerprint_accum(
                ctypes.c_void_p(x.data_ptr()), B, N, x.stride(0) if B > 1 else N, self.level,
                ctypes.c_void_p(self.sums.data_ptr()), ctypes.c_void_p(self.count.data_ptr()),
                _stream_ptr(x.device))
        _lib.check("afd_haar_fingerprint_accum", rc)
        return self

    def all_reduce(self, group=None) -> "FingerprintAccumulator":
        """Sum the partial sums and counts of every rank (one 128 KB NCCL all-reduce, stream-ordered after
        the accumulation kernels; no host round trip)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            packed = torch.cat([self.sums, self.count.to(torch.float64)])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            self.sums = packed[:-1].contiguous()
            self.count = packed[-1:].round().to(torch.int64)
        return self

    def mean(self) -> torch.Tensor:
        """``np.mean(np.abs(packets), (0, 1, 2))`` of the reference: fp64 ``[2^level]`` on the device."""
        return self.sums / self.count.to(torch.float64)


def haar_fingerprint(clips: torch.Tensor, level: int = 14, distributed: bool = False, group=None) -> torch.Tensor:
    """Mean |Haar packet coefficient| over clips, channel and positions, frequency ordered -> fp64 ``[2^level]``.

    ``clips``: ``[n, 1, N]`` or ``[n, N]`` fp32 on a CUDA device (the reference's ``clip_array``).  With
    ``distributed=True`` every rank passes its own shard and all ranks receive the global mean.
    """
    acc = FingerprintAccumulator(level, clips.device).update(clips)
    if distributed:
        acc.all_reduce(group)
    return acc.mean()


def compute_fingerprint_wpt(clips: Iterable[torch.Tensor] | torch.Tensor, seconds: int = 1, level: int = 14,
                            amount: Optional[int] = 2500, device: torch.device | str = "cuda",
                            distributed: bool = False):
    """``_compute_fingerprint_wpt`` without the directory scan and the plotting (both out of scope):
    keeps clips longer than ``seconds`` s, cuts them to ``seconds * 22050`` samples, uses the first ``amount``
    (reference :93-99) and returns ``(freqs, mean_packets)`` as the reference does (:114-115, :125)."""
    n = seconds * SAMPLE_RATE
    if isinstance(clips, torch.Tensor):
        if clips.shape[-1] < n:
            raise ValueError(f"clips must hold at least {n} samples")
        batch = clips[..., :n]
        if amount is not None:
            batch = batch[:amount]
    else:
        kept = [c[..., :n] for c in clips if c.shape[-1] > n]
        if amount is not None:
            kept = kept[:amount]
        if not kept:
            raise ValueError("no clip is longer than the requested window")
        batch = torch.stack(kept)
    batch = batch.to(device=device, dtype=torch.float32)
    mean_packets = haar_fingerprint(batch, level, distributed)
    freqs = torch.linspace(0, SAMPLE_RATE // 2, 1 << level, dtype=torch.float64)
    return freqs, mean_packets


class SpectrumFingerprintAccumulator:
    """Streaming accumulator of the mean-spectrum ("rFFT") fingerprint (reference fingerprints.py:37-62).

    ``update(clips)`` adds the clips' per-sample sums (``afd_clip_sum_accum``: one coalesced pass over the clips);
    ``magnitude()`` takes the single real DFT of the mean clip (``afd_rdft_magnitude``).  The reference's
    rfft -> mask(all bins) -> irfft -> mean -> rfft chain equals ``|rfft(mean clip)|`` because rfft/irfft is the
    identity on even-length clips and the mean is linear.  State: fp64 ``[N]`` sums + clip count on the device."""

    def __init__(self, n_samples: int, device: torch.device | str = "cuda"):
        if n_samples % 2:
            raise ValueError("np.fft.irfft returns 2*(bins-1) samples: the reference's chain needs an even clip length")
        self.n = int(n_samples)
        self.device = torch.device(device)
        self.sums = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        self.count = torch.zeros(1, dtype=torch.int64, device=self.device)

    def update(self, clips: torch.Tensor) -> "SpectrumFingerprintAccumulator":
        x = _as_frames(clips, "SpectrumFingerprintAccumulator.update")
        if x.device != self.sums.device or x.shape[1] != self.n:
            raise RuntimeError(f"expected clips of {self.n} samples on {self.sums.device}")
        B, N = x.shape
        with torch.cuda.device(x.device):
            rc = _lib.load().afd_clip_sum_accum(
                ctypes.c_void_p(x.data_ptr()), B, N, x.stride(0) if B > 1 else N,
                ctypes.c_void_p(self.sums.data_ptr()), ctypes.c_void_p(self.count.data_ptr()), _stream_ptr(x.device))
        _lib.check("afd_clip_sum_accum", rc)
        return self

    def all_reduce(self, group=None) -> "SpectrumFingerprintAccumulator":
        """Sum the per-rank sums and counts (one NCCL all-reduce of N + 1 doubles, stream-ordered)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            packed = torch.cat([self.sums, self.count.to(torch.float64)])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            self.sums = packed[:-1].contiguous()
            self.count = packed[-1:].round().to(torch.int64)
        return self

    def magnitude(self) -> torch.Tensor:
        """``np.abs(np.fft.rfft(mean clip))``: fp64 ``[N/2 + 1]`` on the device."""
        count = int(self.count.item())
        if count == 0:
            raise ValueError("no clips accumulated")
        mag = torch.empty(self.n // 2 + 1, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().afd_rdft_magnitude(ctypes.c_void_p(self.sums.data_ptr()), self.n, 1.0 / count,
                                                ctypes.c_void_p(mag.data_ptr()), _stream_ptr(self.device))
        _lib.check("afd_rdft_magnitude", rc)
        return mag


def compute_fingerprint_rfft(clips: Iterable[torch.Tensor] | torch.Tensor, seconds: int = 1,
                             amount: Optional[int] = 2500, device: torch.device | str = "cuda",
                             distributed: bool = False):
    """``_compute_fingerprint_rfft`` without the directory scan, the plots and the wav export (out of scope):
    keeps clips longer than ``seconds`` s, cuts them to ``seconds * 22050`` samples, uses the first ``amount``
    (reference :45-51) and returns ``(freqs, mean_abs_fft)`` (:60-62)."""
    n = seconds * SAMPLE_RATE
    if isinstance(clips, torch.Tensor):
        if clips.shape[-1] < n:
            raise ValueError(f"clips must hold at least {n} samples")
        batch = clips[..., :n]
        if amount is not None:
            batch = batch[:amount]
    else:
        kept = [c[..., :n] for c in clips if c.shape[-1] > n]
        if amount is not None:
            kept = kept[:amount]
        if not kept:
            raise ValueError("no clip is longer than the requested window")
        batch = torch.stack(kept)
    batch = batch.to(device=device, dtype=torch.float32)
    acc = SpectrumFingerprintAccumulator(n, batch.device).update(batch)
    if distributed:
        acc.all_reduce()
    freqs = torch.arange(n // 2 + 1, dtype=torch.float64) * (SAMPLE_RATE / n)     # np.fft.rfftfreq(n, 1 / SAMPLE_RATE)
    return freqs, acc.magnitude()
