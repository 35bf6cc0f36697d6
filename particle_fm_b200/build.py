"""Build libpfm_b200.so (hand-written CUDA, sm_100a only) in-tree with nvcc.

The shared library is the product: it has no torch dependency and exports the C ABI declared in
include/pfm_b200.h.  nvcc cross-compiles without a GPU, so this also runs in the CPU container."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpfm_b200.so")
SOURCES = ["pfm_api.cu", "epic_simt.cu", "epic_tc.cu", "epic_train.cu", "tf_simt.cu", "tf_tc.cu", "tf_attn_tc.cu", "tf_train.cu", "xty_tc.cu", "extras.cu", "epic_train_tc.cu", "optim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-warn-spills"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpfm_b200.so cannot be built (there is no CPU fallback)")


OBJ_DIR = os.path.join(PKG, "_obj")          # git-ignored; objects are kept so that only changed translation units recompile


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(PKG, "..", "include", "pfm_b200.h")]


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src, *_headers()] if os.path.exists(d))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "pfm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> lib/libpfm_b200.so; only translation units newer than their object are recompiled."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for s in srcs:       # compile the stale translation units in parallel, then link
        o = os.path.join(OBJ_DIR, os.path.basename(s) + ".o")
        objs.append(o)
        if not force and not _obj_stale(s, o):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            try:
                os.remove(cmd[-1])
            except OSError:
                pass
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs]
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
