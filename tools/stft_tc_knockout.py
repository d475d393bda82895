"""Knock-out timing of the tcgen05 STFT kernel: variant libraries built with -DAFD_TC_KO=k (see afd_stft_tc.cu) against the
product library on the same box.  Build them first:  python tools/stft_tc_knockout.py --build   (CPU, nvcc only).
Results are wrong by construction; only the times matter."""
import ctypes, importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if "--build" in sys.argv:
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "audiodeepfake-detection_b200", "build.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    for k in (1, 2, 3, 4, 5, 7, 8):
        print(m.build(True, extra_flags=[f"-DAFD_TC_KO={k}"], out_path=os.path.join(m.PKG_DIR, f"libafd_b200_ko{k}.so")))
    sys.exit(0)
import torch
B, N = 4096, 22050
x = torch.randn(B, N, device="cuda") * 0.1
out = torch.empty(B, 1, 101, 256, device="cuda")
os.environ.pop("AFD_STFT_IMPL", None)
for name in ["libafd_b200.so"] + ["libafd_b200_ko%d.so" % k for k in range(1, 9)]:
    path = os.path.join(ROOT, "audiodeepfake-detection_b200", name)
    if not os.path.exists(path):
        continue
    lib = ctypes.CDLL(path)
    def run():
        return lib.afd_stft_power(ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int64(N), 511, 220,
                                  ctypes.c_float(2.0), 1, ctypes.c_float(1e-12), ctypes.c_void_p(out.data_ptr()), None)
    for _ in range(5):
        assert run() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 30 * 1e3:.1f} us", flush=True)
