"""In-tree build of libafd_b200.so with nvcc for sm_100a (no JIT cache: the .so must travel with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libafd_b200.so")
SOURCES = ["afd_wpt_g0.cu", "afd_wpt_g1.cu", "afd_wpt_g2.cu", "afd_wpt_g3.cu",
           "afd_wpt_x0.cu", "afd_wpt_x1.cu", "afd_wpt_x2.cu", "afd_wpt_x3.cu", "afd_core.cu", "afd_lattice.cu", "afd_wpt.cu",
           "afd_stft.cu", "afd_stft_pfa.cu", "afd_stft_tc.cu", "afd_haar.cu", "afd_rfft.cu", "afd_host.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libafd_b200.so cannot be built")
    return nvcc


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(PKG_DIR, "..", "include", "afd_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str = LIB_PATH,
          ) -> str:
    """Compile every .cu under csrc/ into one shared library (objects built in parallel).

    ``extra_flags`` / ``out_path`` build tuning variants next to the product library for same-box A/B runs
    (tools/ab_bench.py), e.g. ``build(True, extra_flags=["-DAFD_WPT_FFMA2=0"], out_path=".../libafd_b200_prev.so")``."""
    if not force and not _stale() and out_path == LIB_PATH:
        return LIB_PATH
    nvcc = _nvcc()
    variant = out_path != LIB_PATH
    obj_dir = os.path.join(PKG_DIR, "build", "variant") if variant else os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-o", out_path, *objs, "-lcudart"])
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
