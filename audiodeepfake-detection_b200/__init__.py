"""B200-native feature front-end of gan-police/audiodeepfake-detection (wavelet packets, STFT, Haar fingerprint).

The compute path is hand-written CUDA for sm_100a in ``libafd_b200.so`` (C ABI: include/afd_b200.h); this package
is the thin host-side mirror of the reference's transform-module API.  No CPU fallback exists by design.
"""
from .wavelets import Wavelet, get_wavelet  # noqa: F401
from .wavelet_math import (  # noqa: F401
    Packets,
    STFTLayer,
    Normalize,
    compute_pytorch_packet_representation,
    get_transforms,
    calc_normalization,
    fuse_normalize,
    NodeStats,
    NodeStatsTable,
    normalization_stats,
    stft_power_features,
    wavelet_packet_features,
    wpt_out_len,
    stft_out_shape,
)
from . import fingerprint, framing  # noqa: F401
from .framing import cut_frames, utterance_features  # noqa: F401
from .fingerprint import (  # noqa: F401
    FingerprintAccumulator,
    SpectrumFingerprintAccumulator,
    compute_fingerprint_rfft,
    compute_fingerprint_wpt,
    haar_fingerprint,
)

__version__ = "0.1.0"
