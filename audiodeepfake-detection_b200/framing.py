"""Device-side frame cutter: whole utterances -> the fixed-length windows the transforms consume.

The reference cuts every file into ``num_frames // int(seconds * sample_rate)`` consecutive, non-overlapping
windows on the host, one ``torchaudio.load(frame_offset=i * winsize, num_frames=winsize)`` per window
(data_loader.py:178-182, 336-340).  On the device that is a zero-copy view: window i starts at sample
``i * winsize``, so ``[n_windows, winsize]`` rows with row stride ``winsize`` address the utterance in place and the
kernels (which take a row stride) read it without a gather.  The reference's per-window
``torchaudio.functional.resample(audio, sample_rate, resample_rate)`` (data_loader.py:341-344) is ``resample`` below: the same
polyphase windowed-sinc filter as one CUDA kernel (libafd_b200 ``afd_resample``).  File I/O and labels stay out of scope.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def resample(audio: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """``torchaudio.functional.resample(audio, orig_freq, new_freq)`` (torchaudio defaults) on the device.

    ``audio``: fp32 CUDA tensor ``[..., n]``; returns ``[..., ceil(new * n / orig)]``.  Equal rates return the input;
    like the reference (data_loader.py:345-348) the caller decides whether upsampling is allowed -- the filter itself
    handles both directions."""
    if not audio.is_cuda or audio.dtype != torch.float32:
        raise RuntimeError("resample: expected a float32 CUDA tensor (no CPU path)")
    orig_freq, new_freq = int(orig_freq), int(new_freq)
    if orig_freq == new_freq:
        return audio
    lead = audio.shape[:-1]
    x = audio.reshape(-1, audio.shape[-1])
    if x.stride(-1) != 1:
        x = x.contiguous()
    rows, n = x.shape
    n_out = ctypes.c_int64(0)
    lib = _lib.load()
    _lib.check("afd_resample_out_len", lib.afd_resample_out_len(n, orig_freq, new_freq, ctypes.byref(n_out)))
    out = torch.empty((rows, n_out.value), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        for lo in range(0, rows, 65535):
            hi = min(rows, lo + 65535)
            _lib.check("afd_resample", lib.afd_resample(
                ctypes.c_void_p(x[lo:hi].data_ptr()), hi - lo, n, x.stride(0) if rows > 1 else n, orig_freq, new_freq,
                ctypes.c_void_p(out[lo:hi].data_ptr()), n_out.value, stream))
    return out.reshape(*lead, n_out.value)


def cut_frames(audio: torch.Tensor, seconds: float = 1, sample_rate: int = 22050,
               orig_sample_rate: int | None = None) -> torch.Tensor:
    """``[n_samples]`` or ``[1, n_samples]`` -> ``[n_windows, 1, winsize]`` view (no copy); the tail that does not
    fill a window is dropped, as the reference's window table does (data_loader.py:178).

    With ``orig_sample_rate`` the file's windows are cut at the FILE's rate (``int(seconds * orig_sample_rate)`` samples,
    data_loader.py:176-182) and each window is resampled to ``sample_rate`` on its own, exactly like the reference's
    ``__getitem__`` (:336-344); a lower file rate raises as the reference does (:345-348)."""
    if orig_sample_rate is not None and int(orig_sample_rate) != int(sample_rate):
        if orig_sample_rate < sample_rate:
            raise RuntimeError("Sample rate is smaller than desired sample rate. No upsampling possible here.")
        windows = cut_frames(audio, seconds, int(orig_sample_rate))
        return resample(windows, int(orig_sample_rate), int(sample_rate))
    if audio.dim() == 2:
        if audio.shape[0] != 1:
            raise ValueError(f"expected mono audio [1, n], got {tuple(audio.shape)}")
        audio = audio[0]
    if audio.dim() != 1:
        raise ValueError(f"expected [n_samples] or [1, n_samples], got {tuple(audio.shape)}")
    winsize = int(seconds * sample_rate)
    n_windows = audio.shape[0] // winsize
    if n_windows == 0:
        raise ValueError(f"utterance of {audio.shape[0]} samples is shorter than one window of {winsize}")
    if audio.stride(0) != 1:
        audio = audio.contiguous()
    return audio[: n_windows * winsize].view(n_windows, 1, winsize)


def utterance_features(transforms, audio: torch.Tensor, seconds: float = 1, sample_rate: int = 22050,
                       orig_sample_rate: int | None = None):
    """Features of every window of one utterance: ``transforms(cut_frames(audio))`` -> ``([n_windows, C, P, T], aux)``."""
    return transforms(cut_frames(audio, seconds, sample_rate, orig_sample_rate))
