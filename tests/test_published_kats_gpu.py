"""The published PyWavelets / ptwt known answers of tests/test_published_kats_cpu.py, run through the CUDA kernels
(C ABI via ctypes).  The numbers are quoted there with their sources; none of them comes from this repository.
Everything the kernels can express is covered: 'reflect' mode only (the reference's mode, wavelet_math.py:182), so the
'symmetric'-mode db2 example is compared on its interior outputs, which do not depend on the extension, and on its
edges against numpy.pad(mode="reflect") (documented by pywt as the same extension).
"""
import numpy as np
import pytest
import torch

import audiodeepfake_detection_b200 as afd
from audiodeepfake_detection_b200.wavelets import Wavelet

from test_published_kats_cpu import (CA_DB2_SYMMETRIC, CD_DB2_SYMMETRIC, SYM3_DEC_HI, SYM3_DEC_LO, WP_DB1, X_DWT,
                                           X_WP)

pytestmark = pytest.mark.gpu


def _packets(x, name, level, order="freq"):
    xt = torch.tensor([x], dtype=torch.float32, device="cuda")
    out = afd.wavelet_packet_features(xt, Wavelet(name), level, order=order)       # [1, 1, T, 2^level]
    torch.cuda.synchronize()
    return out[0, 0].cpu().numpy()


def test_pywt_wavelet_packet_db1_published_nodes_on_the_gpu():
    """[pywt-wp] WaveletPacket([1..8], 'db1'): wp['a'], wp['d'], wp['aa'], wp['ad'], wp['aaa'], wp['aad']."""
    l1 = _packets(X_WP, "db1", 1)
    assert l1.shape == (4, 2)
    assert np.max(np.abs(l1[:, 0] - WP_DB1["a"])) < 2e-6 and np.max(np.abs(l1[:, 1] - WP_DB1["d"])) < 2e-6
    l2 = _packets(X_WP, "db1", 2)                 # frequency order: aa, ad, dd, da
    assert l2.shape == (2, 4)
    assert np.max(np.abs(l2[:, 0] - WP_DB1["aa"])) < 2e-6 and np.max(np.abs(l2[:, 1] - WP_DB1["ad"])) < 2e-6
    l2n = _packets(X_WP, "db1", 2, order="natural")   # natural order: aa, ad, da, dd
    assert np.array_equal(l2n[:, [0, 1, 3, 2]], l2)
    # a linear ramp: d is constant, so dd = 0 and da = -1 (sqrt(2) * -0.7071): pins which of the two is column 2 / 3
    assert np.max(np.abs(l2[:, 2])) < 1e-6 and np.max(np.abs(l2[:, 3] + 1.0)) < 2e-6
    l3 = _packets(X_WP, "db1", 3)                 # frequency order: aaa, aad, add, ada, dda, ddd, dad, daa
    assert l3.shape == (1, 8)
    assert abs(l3[0, 0] - WP_DB1["aaa"][0]) < 4e-6 and abs(l3[0, 1] - WP_DB1["aad"][0]) < 4e-6
    # the rest follows from the printed nodes: ada = sqrt(2) * ad[0], daa = sqrt(2) * da[0]; the other leaves vanish
    assert abs(l3[0, 3] + 2.0 * np.sqrt(2.0)) < 2e-6 and abs(l3[0, 7] + np.sqrt(2.0)) < 2e-6
    assert np.max(np.abs(l3[0, [2, 4, 5, 6]])) < 1e-6


def test_pywt_dwt_db2_published_example_on_the_gpu():
    """[pywt-dwt] pywt.dwt([3, 7, 1, 1, -2, 5, 4, 6], 'db2'): interior coefficients k = 1..3 as printed; the reflect
    edges against numpy.pad(mode='reflect') and the closed-form db2 taps (Daubechies, Table 6.1)."""
    got = _packets(X_DWT, "db2", 1)
    assert got.shape == (5, 2)
    assert np.max(np.abs(got[1:4, 0] - CA_DB2_SYMMETRIC[1:4])) < 2e-6
    assert np.max(np.abs(got[1:4, 1] - CD_DB2_SYMMETRIC[1:4])) < 2e-6
    r3 = np.sqrt(3.0)
    h = np.array([1 - r3, 3 - r3, 3 + r3, 1 + r3]) / (4 * np.sqrt(2.0))
    g = np.array([-h[3], h[2], -h[1], h[0]])
    xp = np.pad(np.asarray(X_DWT, dtype=np.float64), (2, 2), mode="reflect")
    for k in range(5):
        assert abs(got[k, 0] - sum(h[m] * xp[2 * k + 1 - m + 2] for m in range(4))) < 2e-6
        assert abs(got[k, 1] - sum(g[m] * xp[2 * k + 1 - m + 2] for m in range(4))) < 2e-6


def test_sym3_printed_filter_bank_drives_the_kernel():
    """[pywt-wav] Wavelet('sym3').dec_lo / dec_hi as printed in the PyWavelets docs: an impulse response of the GPU
    analysis step reads the taps back (y_lo[k] = h[2k+1-n0] for x = delta(n - n0))."""
    n0 = 20
    x = np.zeros(64, dtype=np.float32)
    x[n0] = 1.0
    got = _packets(x, "sym3", 1)
    for k in range(got.shape[0]):
        m = 2 * k + 1 - n0
        want_lo = SYM3_DEC_LO[m] if 0 <= m < 6 else 0.0
        want_hi = SYM3_DEC_HI[m] if 0 <= m < 6 else 0.0
        assert abs(got[k, 0] - want_lo) < 1e-7 and abs(got[k, 1] - want_hi) < 1e-7


def test_haar_fingerprint_of_the_published_packet_tree():
    """fingerprints.py:101-115 on the [pywt-wp] signal: level-3 Haar packets of [1..8], frequency order, mean |c|."""
    x = torch.tensor([X_WP], dtype=torch.float32, device="cuda")
    fp = afd.haar_fingerprint(x, 3).cpu().numpy()
    want = np.array([WP_DB1["aaa"][0], abs(WP_DB1["aad"][0]), 0, 2.0 * np.sqrt(2.0), 0, 0, 0, np.sqrt(2.0)])
    assert np.max(np.abs(fp - want)) < 4e-6
