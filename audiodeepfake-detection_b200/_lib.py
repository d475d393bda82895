"""ctypes binding of libafd_b200.so (the C ABI declared in include/afd_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libafd_b200.so")

AFD_ORDER_FREQ = 0
AFD_ORDER_NATURAL = 1

# name -> (restype, argtypes); mirrors include/afd_b200.h one to one
SIGNATURES = {
    "afd_version": (c_int, []),
    "afd_last_error": (c_char_p, []),
    "afd_source_hash": (c_char_p, []),
    "afd_wpt_out_len": (c_int, [c_int64, c_int, c_int, POINTER(c_int64)]),
    "afd_wpt_forward": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(c_double), c_int, c_int, c_int,
                                c_float, c_int, c_float, c_int, c_void_p, POINTER(c_int64), c_void_p]),
    "afd_wpt_forward_ex": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(c_double), c_int, c_int, c_int,
                                   c_float, c_int, c_float, c_int, c_void_p, POINTER(c_float), c_void_p, c_void_p,
                                   c_void_p, POINTER(c_int64), c_void_p]),
    "afd_wpt_forward_host": (c_int, [c_void_p, c_int64, c_int64, c_int64, POINTER(c_double), c_int, c_int, c_int,
                                     c_float, c_int, c_float, c_int, c_void_p, POINTER(c_int64), c_int, c_int64]),
    "afd_stft_power": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_int, c_float,
                               c_void_p, c_void_p]),
    "afd_stft_power_ex": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_int, c_float,
                                  POINTER(c_float), c_void_p, c_void_p, c_void_p]),
    "afd_stft_out_shape": (c_int, [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "afd_stft_power_host": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float, c_int, c_float,
                                    c_void_p, c_int, c_int64]),
    "afd_haar_fingerprint_accum": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                           c_void_p]),
    "afd_haar_fingerprint_host": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                          c_int, c_int64]),
    "afd_clip_sum_accum": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "afd_rdft_magnitude": (c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "afd_resample_out_len": (c_int, [c_int64, c_int, c_int, POINTER(c_int64)]),
    "afd_resample": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "afd_wpt_lattice_info": (c_int, [POINTER(c_double), c_int, POINTER(c_double), POINTER(c_double),
                                     POINTER(c_double), POINTER(c_int)]),
    "afd_wpt_plan_info": (c_int, [c_int64, POINTER(c_double), c_int, c_int, POINTER(c_int), POINTER(c_int),
                                  POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "afd_measure_fp32_fma_tflops": (c_int, [c_int, POINTER(c_double), c_void_p]),
}

_lib = None


class AfdError(RuntimeError):
    """Raised when a libafd_b200 call returns a non-zero status."""

    def __init__(self, fn: str, code: int, message: str):
        super().__init__(f"{fn} failed with status {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the in-tree CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(fn: str, code: int) -> None:
    if code != 0:
        raise AfdError(fn, code, load().afd_last_error().decode("utf-8", "replace"))
