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
SOURCES = [os.path.join(CSRC, "ilqr_kernels.cu"), os.path.join(CSRC, "ilqr_generic.cu"), os.path.join(CSRC, "nn_jacobian.cu")]
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
    "nn_jacobian.cu": (),       # the learned-dynamics Jacobian (tcgen05 tensor-core kernels): only the C ABI header
}


def find_nvcc():
    cand = shutil.which("nvcc") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if cand and os.path.exists(cand) else None


HASH_PATH = LIB_PATH + ".srchash"                  # digest of the sources + flags the library was built from


def source_digest() -> str:
    """SHA-256 over the translation units, their headers and the compiler flags."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in SOURCES + HEADERS:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale():
    """True if the library is missing or was built from other sources.  Decided by CONTENT (the digest the build wrote
    next to the library): `git checkout` and the snapshot that ships the tree to a GPU box reset modification times, which
    made every rank of a torchrun job rebuild a perfectly good library at once.  Without a digest file: by mtime."""
    if not os.path.exists(LIB_PATH):
        return True
    if os.path.exists(HASH_PATH):
        with open(HASH_PATH) as f:
            return f.read().strip() != source_digest()
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS if os.path.exists(p))


def _unit_digest(src) -> str:
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in [src, HEADERS[-1]] + [os.path.join(CSRC, x) for x in UNIT_HEADERS[os.path.basename(src)]]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _unit_stale(src, obj):
    if not os.path.exists(obj) or not os.path.exists(obj + ".srchash"):
        return True
    with open(obj + ".srchash") as f:
        return f.read().strip() != _unit_digest(src)


def _run(cmd, verbose):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if missing or built from other sources; return its path.  Every translation unit is
    compiled to its own object (only the changed ones are recompiled) and the objects are linked into the library.

    Safe under concurrent callers (one process per GPU under torchrun): the whole build holds an exclusive lock, every
    output is written to a temporary name and renamed into place, and staleness is re-checked once the lock is held, so
    the ranks that waited find the finished library and never load a half-written one."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found and %s is missing or stale: the CUDA library cannot be built" % LIB_PATH)
    import fcntl
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():          # another process built it while this one waited for the lock
                return LIB_PATH
            tag = ".tmp%d" % os.getpid()
            objs = []
            for src in SOURCES:
                obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")
                objs.append(obj)
                if not force and not _unit_stale(src, obj):
                    continue
                _run([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj + tag, src], verbose)
                os.replace(obj + tag, obj)
                with open(obj + ".srchash" + tag, "w") as f:
                    f.write(_unit_digest(src))
                os.replace(obj + ".srchash" + tag, obj + ".srchash")
            _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH + tag]
                 + objs, verbose)
            if os.path.exists(HASH_PATH):
                os.remove(HASH_PATH)                    # never a new library under an old digest
            os.replace(LIB_PATH + tag, LIB_PATH)
            with open(HASH_PATH + tag, "w") as f:
                f.write(source_digest())
            os.replace(HASH_PATH + tag, HASH_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
