"""End-to-end checks on the GPU: the reference-facing modules (get_transforms), the shipped checkpoints as
known-answer tests (identical argmax for CUDA features and oracle features), and the host-buffer C-ABI calls."""
import ctypes
import os

import numpy as np
import pytest
import torch

import audiodeepfake_detection_b200 as afd
from audiodeepfake_detection_b200 import _lib
from audiodeepfake_detection_b200.dcnn import load_reference_checkpoint
from oracle import ptwt_like
from oracle import wpt_oracle as oracle
from oracle.filters import DEC_LO

pytestmark = pytest.mark.gpu


class Args(dict):
    __getattr__ = dict.__getitem__


def _args(**kw):
    base = dict(transform="packets", num_of_scales=256, hop_length=220, log_scale=True, power=2.0, wavelet="sym5",
                loss_less="False", features="none", block_norm=False, mean=[-13.6], std=[4.9])
    base.update(kw)
    return Args(base)


@pytest.mark.parametrize("tag,wavelet,tda", [("sym5", "sym5", 1), ("coif4", "coif4", 0), ("stft", None, 0)])
def test_checkpoint_argmax_identical_to_oracle_features(tag, wavelet, tda, golden_dir, golden_frames, cuda_device):
    """north_star: 'the shipped checkpoints in models/ must give identical argmax predictions on the same frames'."""
    frames, labels, _ = golden_frames
    x = torch.from_numpy(frames).unsqueeze(1)
    if wavelet:
        want = torch.from_numpy(oracle.packet_features(frames, DEC_LO[wavelet], 8, log_scale=True)).permute(0, 1, 3, 2)
        args = _args(wavelet=wavelet)
    else:
        want = ptwt_like.stft_layer_forward(x, 511, 220, 2.0, True)
        args = _args(transform="stft")
    mean, std = float(want.mean()), float(want.std())
    args["mean"], args["std"] = [mean], [std]
    transforms, normalize = afd.get_transforms(args, "none", cuda_device, False)
    got, _ = transforms(x.to(cuda_device))
    assert got.shape == want.shape
    model = load_reference_checkpoint(os.path.join(golden_dir, f"ckpt_{tag}.pt"), want.shape[-1], tda).to(cuda_device)
    with torch.no_grad():
        logits_gpu = model(normalize(got)).float().cpu()
        logits_ref = model(((want - mean) / std).to(cuda_device)).float().cpu()
    assert torch.equal(logits_gpu.argmax(-1), logits_ref.argmax(-1))
    assert (logits_gpu.argmax(-1).numpy()[labels == 1] == 1).mean() >= 0.95
    assert float((logits_gpu - logits_ref).abs().max()) < 1e-2 * float(logits_ref.abs().max())


def test_get_transforms_contract(cuda_device):
    x = torch.randn(3, 1, 22050, device=cuda_device)
    tr, norm = afd.get_transforms(_args(wavelet="sym8", num_of_scales=128, loss_less="True"), "none", cuda_device, False)
    feats, aux = tr(x)
    assert tuple(feats.shape) == (3, 2, 128, 187) and isinstance(aux, dict)
    assert norm(feats).shape == feats.shape
    tr, _ = afd.get_transforms(_args(transform="stft"), "none", cuda_device, False)
    assert tuple(tr(x)[0].shape) == (3, 1, 256, 101)
    single = tr(x[0])[0]                      # utils.get_input_dims feeds one un-batched [1, N] item
    assert tuple(single.shape) == (1, 1, 256, 101)
    with pytest.raises(ValueError):
        afd.get_transforms(_args(transform="cwt"), "none", cuda_device, False)


def test_block_norm_and_welford_options(cuda_device):
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((4, 22050)) * 0.1).astype(np.float32)
    xt = torch.from_numpy(x)
    want, stats_ref = ptwt_like.packet_representation(xt, DEC_LO["sym5"], 8, log_scale=True, block_norm=True,
                                                      compute_welford=True)
    got, stats = afd.compute_pytorch_packet_representation(xt.to(cuda_device), afd.Wavelet("sym5"), 8, log_scale=True,
                                                           block_norm=True, compute_welford=True)
    assert got.shape == want.shape
    big = want > -10            # |c| > 7e-3 of the node maximum: log amplification of fp32 noise stays below 1e-3
    assert float((got.cpu() - want)[big].abs().max()) < 1e-3
    assert set(stats) == set(stats_ref)
    for key in ("a" * 8, "d" * 8, "adadadad"):
        m_ref, s_ref = stats_ref[key].finalize()
        m, s = stats[key].finalize()
        assert abs(float(m) - float(m_ref)) < 1e-5 and abs(float(s) - float(s_ref)) < 1e-4 * float(s_ref) + 1e-7


@pytest.mark.parametrize("B,chunk", [(37, 16), (5, 512), (1, 1)])
def test_host_buffer_api_matches_device_api(B, chunk, cuda_device):
    lib = _lib.load()
    rng = np.random.default_rng(B)
    x = np.ascontiguousarray((rng.standard_normal((B, 22050)) * 0.1).astype(np.float32))
    taps = afd.Wavelet("sym5").dec_lo
    c_taps = (ctypes.c_double * 10)(*taps)
    out = np.empty((B, 1, 95, 256), dtype=np.float32)
    T = ctypes.c_int64()
    rc = lib.afd_wpt_forward_host(x.ctypes.data, B, 22050, 22050, c_taps, 10, 8, 0, 2.0, 1, 1e-12, 0, out.ctypes.data,
                                  ctypes.byref(T), 0, chunk)
    assert rc == 0 and T.value == 95
    dev = afd.wavelet_packet_features(torch.from_numpy(x).to(cuda_device), afd.Wavelet("sym5"), 8, log_scale=True)
    assert np.array_equal(out, dev.cpu().numpy())
    spec = np.empty((B, 1, 101, 256), dtype=np.float32)
    assert lib.afd_stft_power_host(x.ctypes.data, B, 22050, 22050, 511, 220, 2.0, 1, 1e-12, spec.ctypes.data, 0, chunk) == 0
    assert np.array_equal(spec, afd.stft_power_features(torch.from_numpy(x).to(cuda_device), log_scale=True).cpu().numpy())
    sums = np.zeros(16384, dtype=np.float64)
    cnt = ctypes.c_int64()
    assert lib.afd_haar_fingerprint_host(x.ctypes.data, B, 22050, 22050, 14, sums.ctypes.data, ctypes.byref(cnt), 0, chunk) == 0
    want, count = oracle.haar_fingerprint_sums(x, 14)
    assert cnt.value == count and np.max(np.abs(sums - want)) < 1e-5 * np.max(want)


def test_strided_rows_and_odd_alignment(cuda_device):
    """Frames inside a larger buffer (row stride > N, rows only 4-byte aligned) must give the same bits."""
    g = torch.Generator(device="cuda").manual_seed(8)
    big = torch.randn(6, 22050 + 7, device=cuda_device, generator=g) * 0.1
    view = big[:, 3:3 + 22050]
    assert not view.is_contiguous()
    w = afd.Wavelet("coif4")
    assert torch.equal(afd.wavelet_packet_features(view, w, 8), afd.wavelet_packet_features(view.contiguous(), w, 8))
    # short filters stage the frame with a 16-byte bulk copy: 8-byte aligned rows (even stride, any even offset) take the same
    # kernel with a 2-float shift and give the same bits; rows that are only 4-byte aligned are served by the two-CTA kernel,
    # whose level 1 runs in the direct form instead of the lattice: same values to fp32 rounding
    big8 = torch.randn(6, 22050 + 8, device=cuda_device, generator=g) * 0.1
    s5 = afd.Wavelet("sym5")
    for off in (2, 4, 6):
        v8 = big8[:, off:off + 22050]
        assert torch.equal(afd.wavelet_packet_features(v8, s5, 8), afd.wavelet_packet_features(v8.contiguous(), s5, 8))
    a, b = afd.wavelet_packet_features(view, s5, 8), afd.wavelet_packet_features(view.contiguous(), s5, 8)
    assert float((a - b).abs().max() / b.abs().max()) < 1e-6
    assert torch.equal(afd.stft_power_features(view), afd.stft_power_features(view.contiguous()))
    assert torch.allclose(afd.haar_fingerprint(view), afd.haar_fingerprint(view.contiguous()), rtol=1e-12)


def test_frame_cutter_is_a_view_and_matches_host_slicing(cuda_device):
    """cut_frames reproduces the reference's window table (data_loader.py:178-182, 336-340) without a copy."""
    wave = torch.randn(1, 22050 * 3 + 777, device=cuda_device) * 0.1
    frames = afd.cut_frames(wave, seconds=1, sample_rate=22050)
    assert frames.shape == (3, 1, 22050) and frames.data_ptr() == wave.data_ptr()
    tr = afd.Packets("sym5", 8, log_scale=True)
    got, _ = afd.utterance_features(tr, wave)
    for i in range(3):
        want, _ = tr(wave[:, i * 22050:(i + 1) * 22050].clone().unsqueeze(0))
        assert torch.equal(got[i:i + 1], want)
    with pytest.raises(ValueError):
        afd.cut_frames(wave[:, :1000])


def test_training_step_consumes_features_on_device(cuda_device):
    """BASELINE config 5 in miniature: fused features (no_grad) -> Normalize -> DCNN fwd/bwd -> Adam, loss falls."""
    from audiodeepfake_detection_b200.train_step import TrainStep

    torch.manual_seed(0)
    tr, norm = afd.get_transforms(_args(), "none", cuda_device, False)
    step = TrainStep(tr, norm, time_len=95, time_dim_add=1, device=cuda_device)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, 1, 22050, generator=g) * 0.1
    x[8:] = x[8:] * torch.linspace(0.2, 1.0, 22050)        # two separable "classes"
    y = torch.cat([torch.zeros(8), torch.ones(8)]).long()
    losses = [float(step(x, y)) for _ in range(12)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
