"""Build libcuriousrl_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

`python -m curiousrl_b200.build [--force]`.  The library is git-ignored but travels with the
working tree, so a GPU box without a compiler run uses the prebuilt file.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcuriousrl_b200.so")
SOURCES = [os.path.join(CSRC, "ilqr_kernels.cu"), os.path.join(CSRC, "ilqr_generic.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("ilqr_common.cuh", "ilqr_models.cuh", "ilqr_core.cuh", "ilqr_coop.cuh", "ilqr_generic.cuh",
                                               "ilqr_generated.cuh", "ilqr_generic_host.h")] + [
    os.path.join(os.path.dirname(PKG_DIR), "include", "curiousrl_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no PTX for other targets
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-fmad=false",                                  # contraction off: FMAs are explicit in the source
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(PKG_DIR, "_obj")             # one object per translation unit (git-ignored), relinked on change
# headers each translation unit depends on (besides the C ABI header)
UNIT_HEADERS = {
    "ilqr_kernels.cu": ("ilqr_common.cuh", "ilqr_models.cuh", "ilqr_core.cuh", "ilqr_coop.cuh", "ilqr_generic_host.h"),
    "ilqr_generic.cu": ("ilqr_common.cuh", "ilqr_models.cuh", "ilqr_core.cuh", "ilqr_generic.cuh", "ilqr_generated.cuh",
                        "ilqr_generic_host.h"),
}


def find_nvcc():
    cand = shutil.which("nvcc") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if cand and os.path.exists(cand) else None


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS if os.path.exists(p))


def _unit_stale(src, obj):
    if not os.path.exists(obj):
        return True
    built = os.path.getmtime(obj)
    deps = [src, HEADERS[-1]] + [os.path.join(CSRC, h) for h in UNIT_HEADERS[os.path.basename(src)]]
    return any(os.path.getmtime(p) > built for p in deps if os.path.exists(p))


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if missing or older than its sources; return its path.  Every translation unit is
    compiled to its own object (only the changed ones are recompiled) and the objects are linked into the library."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found and %s is missing or stale: the CUDA library cannot be built" % LIB_PATH)
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if not force and not _unit_stale(src, obj):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
