"""The C-ABI library loads without a GPU and exports every symbol include/afd_b200.h declares; the host-only
entry points (shape inference, validation, error strings) behave as documented.  No compute calls here."""
import ctypes
import os
import re

import pytest

from audiodeepfake_detection_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "afd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afd_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_declare_the_same_symbols():
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.afd_version() >= 200


def test_library_was_built_from_the_sources_on_disk():
    """Build provenance: the sha256 of csrc/ + include/afd_b200.h compiled into the binary (afd_source_hash) equals the
    hash of the sources in the tree -- the prebuilt .so that travels to the GPU box is the committed code."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "_afd_build", os.path.join(os.path.dirname(_lib.LIB_PATH), "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    digest = _lib.load().afd_source_hash().decode()
    assert len(digest) == 64 and digest == build.source_hash() == build.embedded_hash()


@pytest.mark.parametrize("N,F,level,T", [(22050, 10, 8, 95), (22050, 24, 8, 109), (22050, 16, 7, 187),
                                          (22050, 2, 14, 2), (22050, 10, 0, 22050), (1, 2, 3, 1)])
def test_wpt_out_len(N, F, level, T):
    out = ctypes.c_int64(-1)
    assert _lib.load().afd_wpt_out_len(N, F, level, ctypes.byref(out)) == 0
    assert out.value == T


def test_stft_out_shape():
    lib = _lib.load()
    frames, bins = ctypes.c_int64(), ctypes.c_int64()
    assert lib.afd_stft_out_shape(22050, 511, 220, ctypes.byref(frames), ctypes.byref(bins)) == 0
    assert (frames.value, bins.value) == (101, 256)
    assert lib.afd_stft_out_shape(22050, 512, 2, ctypes.byref(frames), ctypes.byref(bins)) == 0
    assert (frames.value, bins.value) == (11026, 257)
    assert lib.afd_stft_out_shape(16000, 255, 100, ctypes.byref(frames), ctypes.byref(bins)) == 0
    assert (frames.value, bins.value) == (160, 128)      # odd n_fft: 1 + (N - 1) // hop, as torch.stft


def test_validation_errors_carry_a_message():
    lib = _lib.load()
    out = ctypes.c_int64()
    assert lib.afd_wpt_out_len(22050, 9, 8, ctypes.byref(out)) == -1          # odd filter length
    assert b"afd_wpt_out_len" in lib.afd_last_error()
    assert lib.afd_wpt_out_len(0, 10, 8, ctypes.byref(out)) == -1
    assert lib.afd_wpt_out_len(22050, 10, 8, None) == -1
    # null pointers are rejected before any CUDA call is made
    assert lib.afd_wpt_forward(None, 1, 22050, 22050, None, 10, 8, 0, 2.0, 1, 1e-12, 0, None, None, None) == -1
    assert lib.afd_stft_power(None, 1, 22050, 22050, 511, 220, 2.0, 1, 1e-12, None, None) == -1
    assert lib.afd_haar_fingerprint_accum(None, 1, 22050, 22050, 14, None, None, None) == -1
    assert lib.afd_wpt_forward_ex(None, 1, 22050, 22050, None, 10, 8, 0, 2.0, 1, 1e-12, 0, None, None, None, None, None,
                                  None, None) == -1
    assert lib.afd_stft_power_ex(None, 1, 22050, 22050, 511, 220, 2.0, 1, 1e-12, None, None, None, None) == -1
    with pytest.raises(_lib.AfdError) as info:
        _lib.check("afd_wpt_out_len", lib.afd_wpt_out_len(22050, 7, 8, ctypes.byref(out)))
    assert info.value.code == -1


def test_product_does_not_import_the_oracle():
    """The product package must never route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "audiodeepfake-detection_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), name


def test_cpu_tensors_are_rejected_not_emulated():
    import torch

    import audiodeepfake_detection_b200 as afd

    with pytest.raises(RuntimeError, match="no CPU path"):
        afd.Packets("sym5", 8)(torch.randn(2, 1, 22050))
    with pytest.raises(RuntimeError, match="no CPU path"):
        afd.STFTLayer()(torch.randn(2, 1, 22050))


def _lattice_resynth(tans, scale):
    """Taps of the filter pair the kernels' rotation chain evaluates: E(z) = R_{J-1} L(z) ... L(z) R_0 with
    R_m = cos(theta_m) [[1, t_m], [-t_m, 1]] and L(z) delaying the second channel by one pair."""
    import numpy as np

    blocks = [np.array([[1.0, tans[0]], [-tans[0], 1.0]])]
    for t in tans[1:]:
        nxt = [np.zeros((2, 2)) for _ in range(len(blocks) + 1)]
        for j, blk in enumerate(blocks):
            nxt[j][0] += blk[0]
            nxt[j + 1][1] += blk[1]
        rot = np.array([[1.0, t], [-t, 1.0]])
        blocks = [rot @ blk for blk in nxt]
    lo = np.concatenate([blk[0] for blk in blocks]) * scale
    hi = np.concatenate([blk[1] for blk in blocks]) * scale
    return lo, hi


@pytest.mark.parametrize("name", ["haar", "db2", "sym5", "db8", "coif4", "sym8", "coif3", "db16"])
def test_lattice_factorisation_reproduces_the_filter_pair(name):
    """The paraunitary lattice the packet kernel evaluates (half the multiplies of the direct form) must be the
    same filter pair: re-synthesise dec_lo / dec_hi from the reported stage tangents."""
    import numpy as np

    from oracle.filters import dec_hi
    from audiodeepfake_detection_b200.wavelets import Wavelet

    taps = np.asarray(Wavelet(name).dec_lo, dtype=np.float64)
    F = len(taps)
    c_taps = (ctypes.c_double * F)(*taps)
    tans = (ctypes.c_double * 32)()
    scale, resid, usable = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    lib = _lib.load()
    assert lib.afd_wpt_lattice_info(c_taps, F, tans, ctypes.byref(scale), ctypes.byref(resid), ctypes.byref(usable)) == 0
    assert usable.value == 1 and resid.value < 2e-9
    lo, hi = _lattice_resynth(list(tans)[:F // 2], scale.value)
    assert np.max(np.abs(lo - taps)) < 2e-9
    assert np.max(np.abs(hi - np.asarray(dec_hi(taps)))) < 2e-9


def test_plan_info_headline_shapes_fit_one_frame_cta():
    """Headline shapes run the frame kernel: one 512-thread CTA per frame, every level one round of its threads
    (level 1: the 512 threads, levels >= 2: the 256 threads of each half-tree group)."""
    from audiodeepfake_detection_b200.wavelets import Wavelet

    lib = _lib.load()
    for name in ("sym5", "coif4"):
        taps = Wavelet(name).dec_lo
        c_taps = (ctypes.c_double * len(taps))(*taps)
        smem, ctas, lat, npass = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        items, rs = (ctypes.c_int * 24)(), (ctypes.c_int * 24)()
        assert lib.afd_wpt_plan_info(22050, c_taps, len(taps), 8, ctypes.byref(smem), ctypes.byref(ctas),
                                     ctypes.byref(lat), ctypes.byref(npass), items, rs) == 0
        assert ctas.value == 1 and lat.value == 1 and smem.value <= 232448
        assert npass.value == 8                                     # levels 1..8
        assert 256 < items[0] <= 512                                # level 1: one round of the whole CTA
        assert all(0 < items[i] <= 256 for i in range(1, npass.value))  # one round of a group's 256 threads per level

def test_plan_info_every_tabulated_wavelet_plans_within_shared_memory():
    """Every filter that ships a table gets a plan for the BASELINE frame length at level 8 that fits one SM's shared memory
    (the frame kernel's tuned last stride falls back to the untuned one when it would not fit)."""
    from audiodeepfake_detection_b200 import _wavelet_tables
    from audiodeepfake_detection_b200.wavelets import Wavelet

    lib = _lib.load()
    names = sorted(getattr(_wavelet_tables, "DEC_LO", getattr(_wavelet_tables, "TABLES", {})).keys())
    assert len(names) >= 40
    for name in names:
        taps = Wavelet(name).dec_lo
        c_taps = (ctypes.c_double * len(taps))(*taps)
        smem, ctas, lat, npass = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        items, rs = (ctypes.c_int * 24)(), (ctypes.c_int * 24)()
        rc = lib.afd_wpt_plan_info(22050, c_taps, len(taps), 8, ctypes.byref(smem), ctypes.byref(ctas),
                                   ctypes.byref(lat), ctypes.byref(npass), items, rs)
        assert rc == 0, (name, rc)
        assert 0 < smem.value <= 232448 and ctas.value in (1, 2) and npass.value >= 7, (name, smem.value, ctas.value, npass.value)   # two-CTA kernel: level 1 is not a pass
