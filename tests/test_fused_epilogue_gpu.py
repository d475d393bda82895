"""GPU parity of the extended epilogues (afd_wpt_forward_ex / afd_stft_power_ex) against the CPU oracle:
per-node statistics (reference wavelet_math.py:194-200), block-norm scale (:202-203), fused Normalize (:380-382)
and the feature moments of calc_normalization (:436-441)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import ptwt_like
from oracle import wpt_oracle as oracle

import audiodeepfake_detection_b200 as afd
from audiodeepfake_detection_b200 import _lib
from audiodeepfake_detection_b200.wavelets import Wavelet

pytestmark = pytest.mark.gpu

STAT_TOL = 1e-5


def _frames(B, N=22050, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((B, N)) * 0.1).astype(np.float32)


def _oracle_stats(x, name, level, order="freq"):
    raw = oracle.packet_features(x.astype(np.float64), Wavelet(name).dec_lo, level, order=order, dtype=np.float64)[:, 0]
    return raw, np.stack([raw.sum((0, 1)), (raw * raw).sum((0, 1)), np.abs(raw).max((0, 1))])


def _last_level_passes(N, name, level):
    """Number of last-level passes of the kernel plan (> 1: the leaves are produced slice by slice)."""
    taps = Wavelet(name).dec_lo
    c_taps = (ctypes.c_double * len(taps))(*taps)
    passes = ctypes.c_int()
    items = (ctypes.c_int * 24)()
    rs = (ctypes.c_int * 24)()
    smem, ctas, lat = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check("afd_wpt_plan_info", _lib.load().afd_wpt_plan_info(
        N, c_taps, len(taps), level, ctypes.byref(smem), ctypes.byref(ctas), ctypes.byref(lat), ctypes.byref(passes),
        items, rs))
    if ctas.value == 1 and passes.value == level:      # frame kernel: one pass per level, level 1 included
        return 1
    return passes.value - max(level - 2, 0)


@pytest.mark.parametrize("name,level,N,order", [
    ("sym5", 8, 22050, "freq"),        # register-resident statistics (one column pair per thread)
    ("coif4", 8, 22050, "natural"),
    ("haar", 11, 22050, "freq"),       # 512 parents per half tree: column changes per item
    ("sym5", 1, 4000, "freq"),         # level-1 epilogue
    ("db2", 3, 1000, "freq"),          # fewer items than threads
    ("coif4", 8, 24000, "freq"),       # sliced last level (two passes over the leaves)
])
def test_node_stats_match_oracle(name, level, N, order):
    x = _frames(3, N, seed=level)
    raw, want = _oracle_stats(x, name, level, order)
    P = 1 << level
    xt = torch.from_numpy(x).cuda()
    stats = torch.zeros((3, P), dtype=torch.float64, device="cuda")
    got_feats = afd.wavelet_packet_features(xt, Wavelet(name), level, order=order, node_stats=stats)
    plain = afd.wavelet_packet_features(xt, Wavelet(name), level, order=order)
    assert torch.equal(got_feats, plain)                       # statistics do not disturb the features
    got = stats.cpu().numpy()
    scale = np.abs(raw).max() * raw.shape[0] * raw.shape[1]
    assert np.max(np.abs(got[0] - want[0])) < STAT_TOL * scale
    assert np.max(np.abs(got[1] - want[1]) / want[1]) < 1e-4
    assert np.max(np.abs(got[2] - want[2]) / want[2]) < STAT_TOL
    # accumulation: a second, statistics-only launch doubles the sums and keeps the maxima
    none = afd.wavelet_packet_features(xt, Wavelet(name), level, order=order, node_stats=stats, store=False)
    assert none is None
    got2 = stats.cpu().numpy()
    assert np.allclose(got2[:2], 2 * got[:2], rtol=1e-12, atol=1e-12) and np.array_equal(got2[2], got[2])


def test_sliced_last_level_shape_is_covered():
    assert _last_level_passes(24000, "coif4", 8) > 1
    assert _last_level_passes(22050, "coif4", 8) == 1


@pytest.mark.parametrize("loss_less", [False, True])
def test_fused_normalize_scale_and_moments(loss_less):
    x = _frames(4, seed=3)
    xt = torch.from_numpy(x).cuda()
    w = Wavelet("sym5")
    C = 2 if loss_less else 1
    plain = afd.wavelet_packet_features(xt, w, 8, log_scale=True, loss_less=loss_less)
    mean, std = ([-13.6, 0.1], [4.9, 0.9]) if loss_less else ([-13.6], [4.9])
    mom = torch.zeros((C, 2), dtype=torch.float64, device="cuda")
    fused = afd.wavelet_packet_features(xt, w, 8, log_scale=True, loss_less=loss_less, norm=(mean, std), feat_moments=mom)
    m = torch.tensor(mean, device="cuda").view(1, C, 1, 1)
    s = torch.tensor(std, device="cuda").view(1, C, 1, 1)
    want = (plain - m) / s
    assert float((fused - want).abs().max()) < 1e-5
    v = plain.double()
    assert torch.allclose(mom[:, 0], v.sum((0, 2, 3)), rtol=1e-6)
    assert torch.allclose(mom[:, 1], (v * v).sum((0, 2, 3)), rtol=1e-6)
    # block-norm scale: coefficients of column p times node_scale[p]
    scale = torch.rand(256, device="cuda") + 0.5
    raw = afd.wavelet_packet_features(xt, w, 8)
    scaled = afd.wavelet_packet_features(xt, w, 8, node_scale=scale)
    assert float((scaled - raw * scale.view(1, 1, 1, 256)).abs().max()) < 1e-6 * float(raw.abs().max())
    with pytest.raises(_lib.AfdError):
        afd.wavelet_packet_features(xt, w, 8, log_scale=True, norm=([0.0], [0.0]))


def test_block_norm_matches_reference_structure():
    """node / max|node| over the batch, then log (reference wavelet_math.py:202-209), via two fused launches."""
    x = _frames(3, seed=5)
    want, _ = ptwt_like.packet_representation(torch.from_numpy(x), Wavelet("coif4").dec_lo, 8, log_scale=True,
                                              loss_less=True, block_norm=True)
    got, _ = afd.compute_pytorch_packet_representation(torch.from_numpy(x).cuda(), Wavelet("coif4"), 8, log_scale=True,
                                                       loss_less=True, block_norm=True)
    got = got.cpu()
    assert got.shape == want.shape
    assert torch.equal(got[:, 1][want[:, 0] > -10], want[:, 1][want[:, 0] > -10])      # sign channel
    big = want[:, 0] > -10
    assert float((got[:, 0] - want[:, 0])[big].abs().max()) < 1e-3
    assert float(got[:, 0].max()) <= 1e-5                      # every node's maximum maps to log(1 + 1e-12)


def test_welford_dict_accumulates_over_batches():
    x = _frames(6, seed=7)
    xt = torch.from_numpy(x).cuda()
    mod = afd.Packets("sym5", 8, log_scale=True, compute_welford=True)
    _, d = mod(xt[:2])
    mod.block_norm_dict = d
    _, d = mod(xt[2:])
    raw, _ = _oracle_stats(x, "sym5", 8)
    keys = oracle.graycode_paths(8)
    assert list(d) == keys
    for p in (0, 1, 100, 255):
        mean, std = d[keys[p]].finalize()
        col = raw[:, :, p].ravel()
        assert abs(float(mean) - col.mean()) < 1e-6
        assert abs(float(std) - col.std()) < 1e-5 * col.std()
        assert float(d[keys[p]].count) == col.size


@pytest.mark.parametrize("impl", ["tc", "pfa", "bluestein"])
def test_stft_fused_normalize_and_moments(impl):
    x = _frames(5, seed=11)
    xt = torch.from_numpy(x).cuda()
    old = os.environ.get("AFD_STFT_IMPL")
    if impl != "tc":                                  # "tc" = the default dispatch (tcgen05 kernel)
        os.environ["AFD_STFT_IMPL"] = impl
    try:
        plain = afd.stft_power_features(xt, log_scale=True)
        mom = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
        fused = afd.stft_power_features(xt, log_scale=True, norm=(-8.59, 4.63), feat_moments=mom)
        assert float((fused - (plain + 8.59) / 4.63).abs().max()) < 1e-5
        v = plain.double()
        assert torch.allclose(mom[0, 0], v.sum(), rtol=1e-6) and torch.allclose(mom[0, 1], (v * v).sum(), rtol=1e-6)
        assert afd.stft_power_features(xt, log_scale=True, feat_moments=mom, store=False) is None
        assert torch.allclose(mom[0, 0], 2 * v.sum(), rtol=1e-6)
    finally:
        if old is None:
            os.environ.pop("AFD_STFT_IMPL", None)
        else:
            os.environ["AFD_STFT_IMPL"] = old


class _Args(dict):
    __getattr__ = dict.get


@pytest.mark.parametrize("transform", ["packets", "stft"])
def test_calc_normalization_and_fused_normalize(transform):
    """get_transforms(normalization=True) computes mean / std like the reference's calc_normalization pass, and
    fuse_normalize gives the same normalised features as the two-step transforms -> normalize call."""
    x = torch.from_numpy(_frames(12, seed=13)).cuda().unsqueeze(1)
    args = _Args(transform=transform, num_of_scales=256, hop_length=220, log_scale=True, power=2.0, wavelet="sym5",
                 loss_less="False", features="none", block_norm=False, mean=[0.0], std=[1.0])
    batches = [{"audio": x[:5]}, {"audio": x[5:]}]
    tr, norm = afd.get_transforms(args, "none", "cuda", True, norm_batches=batches)
    feats, _ = tr(x)
    mean, std = afd.normalization_stats([feats])
    got_mean, got_std = norm[0].mean, norm[0].std
    assert torch.allclose(got_mean, mean, rtol=1e-5, atol=1e-5) and torch.allclose(got_std, std, rtol=1e-5)
    two_step = norm(feats)
    tr2, ident = afd.fuse_normalize(tr, norm)
    fused, _ = tr2(x)
    assert float((ident(fused) - two_step).abs().max()) < 1e-5
    assert fused.stride() == two_step.stride()
    assert abs(float(fused.mean())) < 1e-4 and abs(float(fused.std()) - 1.0) < 1e-3


def test_get_transforms_writes_and_reloads_the_reference_norm_cache(tmp_path):
    """reference wavelet_math.py:349-367, :449-450: normalization=True computes the statistics over the training batches
    (args.norm_batches stands in for the reference's DataLoader), caches them as <norm_dir>_mean_std.pkl in the reference's
    own pickle format, and the next call -- with or without normalization -- loads that file."""
    import os
    import pickle

    from audiodeepfake_detection_b200 import wavelet_math

    x = torch.from_numpy(_frames(8, seed=21)).cuda().unsqueeze(1)
    args = _Args(transform="packets", num_of_scales=256, hop_length=220, log_scale=True, power=2.0, wavelet="sym5",
                 loss_less="False", features="none", block_norm=False, mean=[0.0], std=[1.0],
                 log_dir=str(tmp_path), data_path="/data/run1", only_use=["ljspeech", "melgan"], sample_rate=22050, seconds=1,
                 norm_batches=[{"audio": x[:4]}, {"audio": x[4:]}])
    tr, norm = afd.get_transforms(args, "none", "cuda", True, verbose=False)
    path = wavelet_math.norm_cache_prefix(args) + "_mean_std.pkl"
    assert os.path.exists(path)
    with open(path, "rb") as fh:
        mean, std = pickle.load(fh)                       # plain numpy arrays, as the reference writes them
    assert abs(float(mean) - float(norm[0].mean)) < 1e-6 and abs(float(std) - float(norm[0].std)) < 1e-6
    feats, aux = tr(x)
    assert len(aux) == 256                                # compute_welford=True as in the reference (:304)
    want_mean, want_std = afd.normalization_stats([feats])
    assert abs(float(mean) - float(want_mean)) < 1e-4 and abs(float(std) - float(want_std)) < 1e-4
    args2 = _Args({k: v for k, v in args.items() if k != "norm_batches"})
    _, norm2 = afd.get_transforms(args2, "none", "cuda", False, verbose=False)          # loads the cache
    assert float(norm2[0].mean) == float(norm[0].mean) and float(norm2[0].std) == float(norm[0].std)
