"""Debug aid (r2): does the STFT kernel ever leave an output element unwritten or read outside its input?  Runs it on NaN-poisoned
output buffers, after other kernels have left garbage in shared memory, and on inputs embedded in a NaN-filled buffer.  (The
non-finite values that showed up inside the full GPU suite came from torch.stft on the CPU, see tests/test_stft_gpu.py.)"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import audiodeepfake_detection_b200 as afd
from audiodeepfake_detection_b200.wavelets import Wavelet
def stft(B):
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.randn(B, 22050, device="cuda", generator=g) * 0.1
    out = torch.full((B, 1, 101, 256), float("nan"), device="cuda")   # poison: unwritten elements show up
    import ctypes
    from audiodeepfake_detection_b200 import _lib
    rc = _lib.load().afd_stft_power(ctypes.c_void_p(x.data_ptr()), B, 22050, 22050, 511, 220, 2.0, 0, 1e-12,
                                    ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    return out
for trial in range(3):
    o = stft(5)
    n = torch.isnan(o)
    print("fresh" if trial == 0 else "again", "nan count", int(n.sum()), n.nonzero()[:8].tolist())
x = torch.randn(600, 22050, device="cuda") * 0.1
afd.wavelet_packet_features(x, Wavelet("coif4"), 8)
afd.haar_fingerprint(x, 14)
torch.cuda.synchronize()
for B in (5, 2, 7, 33):
    o = stft(B)
    n = torch.isnan(o)
    print("after wpt/haar B", B, "nan count", int(n.sum()), n.nonzero()[:8].tolist())
# inputs embedded in a NaN-filled buffer: any value read outside [x, x + B*N) shows up unless it is never used
import ctypes
from audiodeepfake_detection_b200 import _lib
for B, off in ((5, 1000), (5, 1002), (2, 1000), (33, 1004)):
    big = torch.full((B * 22050 + 4000,), float("nan"), device="cuda")
    xin = big[off:off + B * 22050].view(B, 22050)
    xin.copy_(torch.randn(B, 22050, device="cuda") * 0.1)
    out = torch.full((B, 1, 101, 256), float("nan"), device="cuda")
    rc = _lib.load().afd_stft_power(ctypes.c_void_p(xin.data_ptr()), B, 22050, 22050, 511, 220, 2.0, 0, 1e-12,
                                    ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    n = torch.isnan(out)
    print("embedded B", B, "offset", off, "rc", rc, "nan count", int(n.sum()), n.nonzero()[:6].tolist())
