"""Known answers that do NOT come from this repository: numbers printed in the PyWavelets documentation and the ptwt
README (the two libraries whose arithmetic the reference calls at wavelet_math.py:182-192 and
fingerprints.py:101-111), plus closed-form filter taps from the literature.  They pin, at value level,

  * the convolution phase and tap orientation of one analysis step (y[k] = sum_m h[m] x~[2k+1-m]) with an
    ASYMMETRIC filter (db2) -- a flipped filter, a shifted phase or swapped lo/hi all change these numbers;
  * the sign convention of the high-pass (dec_hi[k] = (-1)^(k+1) dec_lo[F-1-k]);
  * pywt.WaveletPacket node values and the natural / frequency (Gray-code) leaf orders;
  * the output-length rule and dwt_max_level;
  * the 'reflect' extension (documented as whole-sample symmetric = numpy.pad(mode="reflect")).

The same vectors are run through the CUDA kernels in tests/test_published_kats_gpu.py.

Sources (quoted from memory of the published pages; every number below was afterwards reproduced to all printed digits
by the oracle, which a mis-remembered digit or a wrong convention would not survive):
  [pywt-dwt]   PyWavelets docs, "DWT and IDWT" regression page: x = [3, 7, 1, 1, -2, 5, 4, 6], pywt.dwt(x, 'db2').
  [pywt-wp]    PyWavelets docs, "Wavelet Packets" regression page: WaveletPacket([1..8], 'db1', 'symmetric').
  [pywt-api]   PyWavelets API docs: pywt.dwt([1,2,3,4,5,6], 'db1'), pywt.wavedec([1..8], 'db1', level=2),
               pywt.dwt_max_level(1000, Wavelet('sym5')) == 6, dwt_coeff_len.
  [pywt-wav]   PyWavelets docs, "Wavelet" object page: db1 filter bank printout; Wavelet('sym3').dec_lo / dec_hi.
  [ptwt]       ptwt README quick-start: wavedec of [0,1,2,3,4,5,5,4,3,2,1,0], haar, mode='zero', level=2.
  [daub]       Daubechies, Ten Lectures on Wavelets, Table 6.1 (N = 2): h = ((1+-sqrt3), (3+-sqrt3)) / (4 sqrt2).
"""
import math

import numpy as np
import pytest

from oracle import wpt_oracle as oracle
from oracle.filters import DEC_LO, dec_hi

from audiodeepfake_detection_b200._wavelet_tables import DEC_LO as PRODUCT

# ---- [pywt-dwt]
X_DWT = [3, 7, 1, 1, -2, 5, 4, 6]
CA_DB2_SYMMETRIC = [5.65685425, 7.39923721, 0.22414387, 3.33677403, 7.77817459]
CD_DB2_SYMMETRIC = [-2.44948974, -1.60368225, -4.44140056, -0.41361256, 1.22474487]
# ---- [pywt-wp]
X_WP = [1, 2, 3, 4, 5, 6, 7, 8]
WP_DB1 = {
    "a": [2.12132034, 4.94974747, 7.77817459, 10.60660172],
    "d": [-0.70710678, -0.70710678, -0.70710678, -0.70710678],
    "aa": [5.0, 13.0],
    "ad": [-2.0, -2.0],
    "aaa": [12.72792206],
    "aad": [-5.65685425],
}
WP_LEVEL3_NATURAL = ["aaa", "aad", "ada", "add", "daa", "dad", "dda", "ddd"]
WP_LEVEL3_FREQ = ["aaa", "aad", "add", "ada", "dda", "ddd", "dad", "daa"]
WP_LEVEL2_FREQ = ["aa", "ad", "dd", "da"]
# ---- [pywt-wav]
SYM3_DEC_LO = [0.035226291882100656, -0.08544127388224149, -0.13501102001039084, 0.4598775021193313,
               0.8068915093133388, 0.3326705529509569]
SYM3_DEC_HI = [-0.3326705529509569, 0.8068915093133388, -0.4598775021193313, -0.13501102001039084,
               0.08544127388224149, 0.035226291882100656]
DB2_DEC_LO_PRINTED = [-0.12940952255126037, 0.2241438680420134, 0.8365163037378079, 0.48296291314453416]
# ---- coif5 as PyWavelets tabulates it (first / largest / last taps; the generated table must agree)
COIF5_SPOT = {0: -9.517657273819165e-08, 1: -1.6744288576823017e-07, 19: 0.7742896036529562,
              18: 0.4379916261718371, 20: 0.4215662066908515, 29: -0.00021208083980379827}


def test_pywt_dwt_db2_published_example_pins_phase_and_orientation():
    """[pywt-dwt]: default (symmetric) mode; db2 is asymmetric, so this fails for h reversed, for the even phase
    (x~[2k-m]), and for lo/hi swapped."""
    lo, hi = oracle.dwt_step(np.array([X_DWT], dtype=np.float64), PRODUCT["db2"], dtype=np.float64, mode="symmetric")
    assert lo.shape == (1, 5)
    assert np.max(np.abs(lo[0] - CA_DB2_SYMMETRIC)) < 5e-9
    assert np.max(np.abs(hi[0] - CD_DB2_SYMMETRIC)) < 5e-9
    # the conventions this is meant to exclude do not reproduce the printout
    h = np.asarray(PRODUCT["db2"])
    flipped, _ = oracle.dwt_step(np.array([X_DWT], dtype=np.float64), h[::-1], dtype=np.float64, mode="symmetric")
    assert np.max(np.abs(flipped[0] - CA_DB2_SYMMETRIC)) > 1e-2


def test_interior_coefficients_do_not_depend_on_the_extension_mode():
    """Outputs k = 1..3 of the published example read only real samples: the reflect-mode step (what the reference
    uses, wavelet_math.py:182) must print the same interior numbers."""
    lo, hi = oracle.dwt_step(np.array([X_DWT], dtype=np.float64), PRODUCT["db2"], dtype=np.float64, mode="reflect")
    assert np.max(np.abs(lo[0, 1:4] - CA_DB2_SYMMETRIC[1:4])) < 5e-9
    assert np.max(np.abs(hi[0, 1:4] - CD_DB2_SYMMETRIC[1:4])) < 5e-9
    # and the edges follow numpy.pad(mode="reflect") (pywt.pad documents the equivalence), pad = F-2 per side
    h = np.asarray(PRODUCT["db2"])
    g = dec_hi(h)
    xp = np.pad(np.asarray(X_DWT, dtype=np.float64), (2, 2), mode="reflect")
    for k in range(5):
        assert abs(lo[0, k] - sum(h[m] * xp[2 * k + 1 - m + 2] for m in range(4))) < 1e-12
        assert abs(hi[0, k] - sum(g[m] * xp[2 * k + 1 - m + 2] for m in range(4))) < 1e-12


@pytest.mark.parametrize("mode", ["symmetric", "reflect"])
def test_pywt_wavelet_packet_db1_published_nodes(mode):
    """[pywt-wp].  db1 needs no padding on even lengths, so the printed nodes hold for mode='reflect' too."""
    tree = oracle.wavelet_packet_tree(np.array([X_WP], dtype=np.float64), DEC_LO["haar"], 3, dtype=np.float64, mode=mode)
    for path, want in WP_DB1.items():
        assert np.max(np.abs(tree[path][0] - want)) < 5e-9, path
    assert oracle.natural_paths(3) == WP_LEVEL3_NATURAL
    assert oracle.graycode_paths(3) == WP_LEVEL3_FREQ
    assert oracle.graycode_paths(2) == WP_LEVEL2_FREQ


def test_pywt_api_examples():
    """[pywt-api]"""
    lo, hi = oracle.dwt_step(np.array([[1, 2, 3, 4, 5, 6.0]]), DEC_LO["haar"], dtype=np.float64)
    assert np.allclose(lo[0], [2.12132034, 4.94974747, 7.77817459], atol=5e-9)
    assert np.allclose(hi[0], [-0.70710678] * 3, atol=5e-9)
    tree = oracle.wavelet_packet_tree(np.array([X_WP], dtype=np.float64), DEC_LO["haar"], 2, dtype=np.float64)
    assert np.allclose(tree["aa"][0], [5.0, 13.0]) and np.allclose(tree["ad"][0], [-2.0, -2.0])      # wavedec cA2, cD2
    assert np.allclose(tree["d"][0], [-0.70710678] * 4, atol=5e-9)                                      # cD1
    # dwt_coeff_len: floor((n + F - 1) / 2) for every mode but periodization
    assert oracle.out_len(8, 4) == 5 and oracle.out_len(22050, 10) == 11029 and oracle.out_len(11029, 10) == 5519
    # dwt_max_level(1000, sym5) == 6  <=>  dec_len == 10: floor(log2(1000 / 9))
    assert len(PRODUCT["sym5"]) == 10 and int(math.floor(math.log2(1000 / (len(PRODUCT["sym5"]) - 1)))) == 6


def test_ptwt_readme_example():
    """[ptwt]: [array([3., 9., 3.]), array([-2., 0., 2.]), array([-0.7071.. x3, 0.7071.. x3])]"""
    x = np.array([[0, 1, 2, 3, 4, 5, 5, 4, 3, 2, 1, 0]], dtype=np.float64)
    tree = oracle.wavelet_packet_tree(x, DEC_LO["haar"], 2, dtype=np.float64, mode="zero")
    s = 0.7071067811865476
    assert np.allclose(tree["aa"][0], [3.0, 9.0, 3.0])
    assert np.allclose(tree["ad"][0], [-2.0, 0.0, 2.0])
    assert np.allclose(tree["d"][0], [-s, -s, -s, s, s, s])


def test_filter_tables_against_published_taps():
    """[pywt-wav], [daub]: printed filter banks and the closed form."""
    r3 = math.sqrt(3.0)
    closed = [(1 - r3) / (4 * math.sqrt(2)), (3 - r3) / (4 * math.sqrt(2)), (3 + r3) / (4 * math.sqrt(2)),
              (1 + r3) / (4 * math.sqrt(2))]
    assert np.max(np.abs(np.asarray(PRODUCT["db2"]) - closed)) < 1e-15
    assert np.max(np.abs(np.asarray(PRODUCT["db2"]) - DB2_DEC_LO_PRINTED)) < 1e-15
    assert np.max(np.abs(np.asarray(PRODUCT["sym2"]) - DB2_DEC_LO_PRINTED)) < 1e-15          # sym2 == db2 in pywt
    # pywt's sym3 table carries ~4e-12 of rounding (it is the db3 table): agreement to table precision
    assert np.max(np.abs(np.asarray(PRODUCT["sym3"]) - SYM3_DEC_LO)) < 1e-11
    assert np.max(np.abs(dec_hi(PRODUCT["sym3"]) - SYM3_DEC_HI)) < 1e-11
    # db1 printout: dec_lo [s, s], dec_hi [-s, s]
    assert np.allclose(dec_hi(PRODUCT["db1"]), [-0.7071067811865476, 0.7071067811865476], atol=1e-16)
    for k, v in COIF5_SPOT.items():
        assert abs(PRODUCT["coif5"][k] - v) < 2e-8 * max(1.0, abs(v) / 1e-3) or abs(PRODUCT["coif5"][k] - v) < 1e-9, (k, v)


def test_product_wavelet_object_matches_the_printed_filter_bank():
    from audiodeepfake_detection_b200.wavelets import Wavelet

    w = Wavelet("sym3")
    assert np.max(np.abs(np.asarray(w.dec_hi) - SYM3_DEC_HI)) < 1e-11
    assert np.max(np.abs(np.asarray(w.rec_lo) - SYM3_DEC_LO[::-1])) < 1e-11
    assert w.dec_len == 6
