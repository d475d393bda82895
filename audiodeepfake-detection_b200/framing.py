"""Device-side frame cutter: whole utterances -> the fixed-length windows the transforms consume.

The reference cuts every file into ``num_frames // int(seconds * sample_rate)`` consecutive, non-overlapping
windows on the host, one ``torchaudio.load(frame_offset=i * winsize, num_frames=winsize)`` per window
(data_loader.py:178-182, 336-340).  On the device that is a zero-copy view: window i starts at sample
``i * winsize``, so ``[n_windows, winsize]`` rows with row stride ``winsize`` address the utterance in place and the
kernels (which take a row stride) read it without a gather.  File I/O, resampling and labels stay out of scope.
"""
from __future__ import annotations

import torch


def cut_frames(audio: torch.Tensor, seconds: float = 1, sample_rate: int = 22050) -> torch.Tensor:
    """``[n_samples]`` or ``[1, n_samples]`` -> ``[n_windows, 1, winsize]`` view (no copy); the tail that does not
    fill a window is dropped, as the reference's window table does (data_loader.py:178)."""
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


def utterance_features(transforms, audio: torch.Tensor, seconds: float = 1, sample_rate: int = 22050):
    """Features of every window of one utterance: ``transforms(cut_frames(audio))`` -> ``([n_windows, C, P, T], aux)``."""
    return transforms(cut_frames(audio, seconds, sample_rate))
