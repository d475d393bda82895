"""In-tree build of libafd_b200.so with nvcc for sm_100a (no JIT cache: the .so must travel with the repo)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libafd_b200.so")
SOURCES = ["afd_wpt_g0.cu", "afd_wpt_g1.cu", "afd_wpt_g2.cu", "afd_wpt_g3.cu",
           "afd_wpt_x0.cu", "afd_wpt_x1.cu", "afd_wpt_x2.cu", "afd_wpt_x3.cu", "afd_core.cu", "afd_lattice.cu", "afd_wpt.cu",
           "afd_stft.cu", "afd_stft_pfa.cu", "afd_stft_tc.cu", "afd_haar.cu", "afd_rfft.cu", "afd_resample.cu", "afd_host.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libafd_b200.so cannot be built")
    return nvcc


HASH_MARKER = b"AFD_SOURCE_HASH="


def _source_files() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
        os.path.join(PKG_DIR, "..", "include", "afd_b200.h")]


def source_hash() -> str:
    """sha256 over the names and contents of csrc/* and include/afd_b200.h: what the library is built from."""
    h = hashlib.sha256()
    for path in _source_files():
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    return h.hexdigest()


def embedded_hash(lib_path: str = LIB_PATH):
    """The source hash compiled into a built library (read from the file, no dlopen); ``None`` if absent."""
    try:
        with open(lib_path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    i = blob.find(HASH_MARKER)
    if i < 0:
        return None
    digest = blob[i + len(HASH_MARKER): i + len(HASH_MARKER) + 64]
    return digest.decode("ascii", "replace") if len(digest) == 64 else None


def _deps(path: str, seen=None) -> set[str]:
    """A source file and the local headers it includes (transitively): what its object depends on."""
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if line.startswith("#include \""):
                name = line.split('"')[1]
                _deps(os.path.normpath(os.path.join(os.path.dirname(path), name)), seen)
    return seen


def _stale() -> bool:
    """The binary is current when the source hash compiled into it equals the hash of the sources on disk."""
    return embedded_hash() != source_hash()


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
    digest = source_hash()
    objs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        flags = list(extra_flags)
        only = globals().get("SOURCES_VARIANT_ONLY")
        if variant and only and src not in only:     # variant builds may recompile a few units only
            prod = os.path.join(PKG_DIR, "build", src.replace(".cu", ".o"))
            if os.path.exists(prod):
                objs.append(prod)
                continue
        if src == "afd_core.cu":                     # the one unit that carries the provenance string
            flags.append('-DAFD_SOURCE_HASH="%s"' % digest)
        elif (not force and not variant and os.path.exists(obj)
              and os.path.getmtime(obj) > max(os.path.getmtime(d) for d in _deps(os.path.join(CSRC, src)))):
            objs.append(obj)                         # unchanged unit: keep its object
            continue
        cmd = [nvcc, *NVCC_FLAGS, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out_path, *objs, "-lcudart"])
    return out_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
