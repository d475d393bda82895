"""Host-side behaviour of the reference-facing boundary that needs no GPU: the normalisation cache of get_transforms
(reference wavelet_math.py:327-371, :449-450), the quadrature-mirror check on foreign wavelet objects, error messages."""
import os
import pickle

import numpy as np
import pytest
import torch

import audiodeepfake_detection_b200 as afd
from audiodeepfake_detection_b200 import wavelet_math
from audiodeepfake_detection_b200.wavelets import Wavelet, check_orthogonal_pair, get_wavelet


class Args(dict):
    """Like the reference's DotDict (utils.py:321-395): attribute access is a dict lookup."""
    def __getattr__(self, k):
        return self[k]


def _args(tmp_path, **kw):
    base = dict(transform="packets", num_of_scales=256, hop_length=220, log_scale=True, power=2.0, wavelet="sym5",
                loss_less="False", features="none", block_norm=False, mean=[-1.0], std=[2.0],
                log_dir=str(tmp_path), data_path="/data/fake_22050_22050_0.7_x", only_use=["ljspeech", "melgan"],
                sample_rate=22050, seconds=1)
    base.update(kw)
    return Args(base)


def test_norm_cache_file_name_follows_the_reference(tmp_path):
    a = _args(tmp_path)
    want = (str(tmp_path) + "/norms/" + "_data_fake_22050_22050_0.7_x" + "_ljspeech-melgan_packets_sym5_256_2.0_22050_1secs")
    assert wavelet_math.norm_cache_prefix(a) == want
    assert wavelet_math.norm_cache_prefix(_args(tmp_path, loss_less="True")).endswith("_2.0_loss_less_22050_1secs")
    assert wavelet_math.norm_cache_prefix(Args(transform="stft")) is None          # a bare config: no cache


def test_get_transforms_loads_the_reference_pickle_cache(tmp_path):
    a = _args(tmp_path)
    prefix = wavelet_math.norm_cache_prefix(a)
    os.makedirs(os.path.dirname(prefix))
    with open(prefix + "_mean_std.pkl", "wb") as fh:                              # the reference's format (:449-450)
        pickle.dump([np.array([-13.25], dtype=np.float64), np.array([4.5], dtype=np.float64)], fh)
    for normalization in (False, True):                                            # the cache wins either way (:349-355)
        tr, norm = afd.get_transforms(a, "none", "cpu", normalization, verbose=False)
        assert isinstance(tr[0], afd.Packets) and tr[0].compute_welford is True    # reference :304
        assert float(norm[0].mean) == pytest.approx(-13.25) and float(norm[0].std) == pytest.approx(4.5)
    # without a cache and without normalisation: args.mean / args.std (:368-371)
    tr, norm = afd.get_transforms(_args(tmp_path / "other"), "none", "cpu", False, verbose=False)
    assert float(norm[0].mean) == -1.0 and float(norm[0].std) == 2.0


def test_cache_file_cannot_run_code(tmp_path):
    class Boom:
        def __reduce__(self):
            return (os.system, ("echo pwned > /dev/null",))

    a = _args(tmp_path)
    prefix = wavelet_math.norm_cache_prefix(a)
    os.makedirs(os.path.dirname(prefix))
    with open(prefix + "_mean_std.pkl", "wb") as fh:
        pickle.dump([Boom(), Boom()], fh)
    with pytest.raises(pickle.UnpicklingError):
        afd.get_transforms(a, "none", "cpu", False, verbose=False)


def test_normalization_without_any_source_says_what_is_missing(tmp_path):
    with pytest.raises(RuntimeError, match="norm_batches"):
        afd.get_transforms(_args(tmp_path), "none", "cpu", True, verbose=False)


def test_unsupported_feature_heads_name_the_reference_lines(tmp_path):
    with pytest.raises(NotImplementedError, match="316-323"):
        afd.get_transforms(_args(tmp_path), "lfcc", "cpu", False)
    with pytest.raises(ValueError):
        afd.get_transforms(_args(tmp_path, transform="cqt"), "none", "cpu", False)


def test_foreign_wavelet_objects_must_be_quadrature_mirrors():
    class W:
        def __init__(self, lo, hi=None, name="w"):
            self.dec_lo, self.name = lo, name
            if hi is not None:
                self.dec_hi = hi

    sym5 = Wavelet("sym5")
    assert get_wavelet(W(list(sym5.dec_lo), list(sym5.dec_hi))).dec_lo == list(sym5.dec_lo)
    get_wavelet(W(list(sym5.dec_lo)))                                              # no dec_hi: nothing to contradict
    # bior2.2 as pywt tabulates it: dec_hi is NOT the mirror of dec_lo
    lo = [0.0, -0.1767766952966369, 0.3535533905932738, 1.0606601717798214, 0.3535533905932738, -0.1767766952966369]
    hi = [0.0, 0.3535533905932738, -0.7071067811865476, 0.3535533905932738, 0.0, 0.0]
    with pytest.raises(ValueError, match="quadrature mirror"):
        get_wavelet(W(lo, hi, "bior2.2"))
    with pytest.raises(ValueError):
        check_orthogonal_pair(W(list(sym5.dec_lo), list(sym5.dec_hi)[::-1]))
    with pytest.raises(ValueError):
        Wavelet("bior2.2")


def test_wpt_out_len_matches_the_c_abi():
    import ctypes

    from audiodeepfake_detection_b200 import _lib

    for n, f, lev in ((22050, 10, 8), (22050, 24, 8), (22051, 2, 14), (400, 60, 2), (7, 2, 0)):
        out = ctypes.c_int64()
        assert _lib.load().afd_wpt_out_len(n, f, lev, ctypes.byref(out)) == 0
        assert afd.wpt_out_len(n, f, lev) == out.value
    with pytest.raises(_lib.AfdError):
        afd.wpt_out_len(22050, 9, 8)
