"""In-tree build of libarkmpc_b200.so (sm_100a only) with nvcc.  No torch extension machinery: the
product is a plain C-ABI shared library (include/arkmpc_b200.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libarkmpc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(PKG, "lib", "obj")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libarkmpc_b200 cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    out = [os.path.join(ROOT, "include", "arkmpc_b200.h")]
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ark_mpc_b200/lib/libarkmpc_b200.so."""
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    # a translation unit is rebuilt when it or any header changed (not when a sibling .cu did)
    hdr_time = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_time, os.path.getmtime(src)):
            cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", "-o", obj, src]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
        return obj

    # one translation unit per thread: the point kernels take much longer to compile than the scalar ones
    from concurrent.futures import ThreadPoolExecutor

    srcs = sorted(sources(), key=lambda f: (not os.path.basename(f).startswith("curve_"), f))  # slowest units first
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 4))) as pool:
        objs = list(pool.map(compile_one, srcs))
    tmp = LIB + ".tmp"  # link beside the target and rename: a failed link must not remove a working library
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *objs, "-ldl"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
