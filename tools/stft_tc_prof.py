"""Per-role clock64 breakdown of the tcgen05 STFT kernel (one CTA prints its averages per unit).
Build the instrumented library first (CPU, nvcc only):
  python -c "import importlib.util as u; s=u.spec_from_file_location('b','audiodeepfake-detection_b200/build.py'); m=u.module_from_spec(s); s.loader.exec_module(m); m.build(True, extra_flags=['-DAFD_TC_PROF'], out_path=m.PKG_DIR+'/libafd_b200_prof.so')"
then run this script on the GPU box."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "audiodeepfake-detection_b200", "libafd_b200_prof.so"))
os.environ["AFD_STFT_IMPL"] = "tc"
B, N = 4096, 22050
x = torch.randn(B, N, device="cuda") * 0.1
out = torch.empty(B, 1, 101, 256, device="cuda")
for i in range(2):
    rc = lib.afd_stft_power(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), 511, 220,
                            ctypes.c_float(2.0), 1, ctypes.c_float(1e-12), ctypes.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    print("---", flush=True)
