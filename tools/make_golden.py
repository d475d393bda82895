"""Build tests/golden/ from the read-only reference checkout (run in the build container only).

  * frames.npz        -- every whole 1-s frame (22050 samples, int16) of reference tests/data/**/*.wav
                          (frame cutting as reference data_loader.py:178-182) and the ten 1-s
                          audio-samples/classification_examples/*.wav, with labels (0 real, 1 fake).
  * ckpt_*.pt         -- the three shipped DCNN snapshots (reference models/*.pt), byte-identical copies
                          (binary fixtures, not source).
  * stft_torchaudio.npz -- outputs of the reference's own STFT code path (torchaudio Spectrogram + log) on
                          the first four frames.
"""
import glob
import os
import shutil

import numpy as np
import torch
from scipy.io import wavfile

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

frames, labels, names = [], [], []
for path in sorted(glob.glob(f"{REF}/tests/data/*/*.wav")):
    sr, wav = wavfile.read(path)
    assert sr == 22050 and wav.dtype == np.int16 and wav.ndim == 1
    label = 0 if "/real/" in path else 1
    for i in range(wav.shape[0] // 22050):
        frames.append(wav[i * 22050:(i + 1) * 22050]); labels.append(label)
        names.append(os.path.relpath(path, REF) + f"#{i}")
n_testdata = len(frames)
for path in sorted(glob.glob(f"{REF}/audio-samples/classification_examples/*.wav")):
    sr, wav = wavfile.read(path)
    assert sr == 22050, (path, sr)
    if wav.dtype != np.int16:
        wav = np.clip(np.round(wav.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    if wav.ndim > 1:
        wav = wav[:, 0]
    assert wav.shape[0] >= 22050, (path, wav.shape)
    frames.append(wav[:22050]); labels.append(1); names.append(os.path.relpath(path, REF))
np.savez_compressed(os.path.join(OUT, "frames.npz"), frames=np.stack(frames), labels=np.array(labels, np.int8),
                    names=np.array(names), n_testdata=n_testdata)
print("frames", np.stack(frames).shape, "labels", np.bincount(labels))

for tag in ("sym5", "coif4", "stft"):
    pat = "packets" + tag if tag != "stft" else "stft"
    src = glob.glob(f"{REF}/models/model_{pat}_*.pt")
    assert len(src) == 1, src
    shutil.copyfile(src[0], os.path.join(OUT, f"ckpt_{tag}.pt"))

from torchaudio.transforms import Spectrogram  # noqa: E402  (the reference's STFT code path)
x = torch.from_numpy(np.stack(frames[:4]).astype(np.float32) / 32768.0).unsqueeze(1)
spec = Spectrogram(n_fft=511, hop_length=220, power=2.0)(x)
np.savez_compressed(os.path.join(OUT, "stft_torchaudio.npz"), power=spec.numpy(), log=torch.log(spec + 1e-12).numpy())
print("stft", tuple(spec.shape))

# resample_torchaudio.npz -- outputs of torchaudio.functional.resample (the reference's call at data_loader.py:341-344) for
# the file rates the reference's datasets come in (LJSpeech 22.05 kHz needs none; JSUT 48 kHz, ASVspoof 16 kHz is rejected,
# 44.1 / 24 kHz vocoders): seeded noise + a chirp, short enough to commit.
import torchaudio.functional as AF  # noqa: E402
rng = np.random.default_rng(7)
res = {}
for orig in (44100, 48000, 24000, 32000):
    n = orig // 8 + 37
    tt = np.arange(n) / orig
    sig = (0.3 * rng.standard_normal((2, n)) + 0.5 * np.sin(2 * np.pi * (200 + 3000 * tt) * tt)).astype(np.float32)
    res[f"x_{orig}"] = sig
    res[f"y_{orig}"] = AF.resample(torch.from_numpy(sig), orig, 22050).numpy()
np.savez_compressed(os.path.join(OUT, "resample_torchaudio.npz"), **res)
print("resample", {k: v.shape for k, v in res.items()})
