"""Run-time compiled device models: any `DynamicModelWrapper` gets one at `init` (reference: basic_ilqr.py:50-66 takes
any scenario - `iLQRDynamicModel` / `iLQRObjectiveFunction` lambdify its sympy expressions and numba compiles them on
first use, ilqr_dynamic_model.py:25-37).

`compile_scenario(scenario)`: the generator of csrc/ilqr_generated.cuh prints the scenario's dynamics, Jacobian, cost,
gradient and Hessian (plain and under the log-barrier cost) as structs GenUser / GenUserBarrier; nvcc compiles
csrc/ilqr_user_abi.cu - the general-form kernels over that header behind the same C ABI - for sm_100a into a library of
its own, cached in-tree under curiousrl_b200/_jit/<digest>/ (keyed by the generated text, the kernel sources and the
flags, so an edited scenario is a new library and an unchanged one is reused; the directory travels with the tree).
There is no interpreter fallback: without nvcc and without a cached library `init` raises.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

from . import _native, build as _build
from . import codegen

JIT_DIR = os.path.join(_build.PKG_DIR, "_jit")
USER_MODEL_ID, USER_BARRIER_MODEL_ID = 100, 101
MAX_D = 16
_SOURCES = ("ilqr_user_abi.cu", "ilqr_generic.cu", "ilqr_generic.cuh", "ilqr_generic_host.h", "ilqr_core.cuh", "ilqr_models.cuh",
            "ilqr_common.cuh")
_loaded = {}


def _digest(text: str) -> str:
    h = hashlib.sha256(" ".join(_build.NVCC_FLAGS).encode())
    h.update(text.encode())
    for name in _SOURCES:
        with open(os.path.join(_build.CSRC, name), "rb") as f:
            h.update(f.read())
    with open(_build.HEADERS[-1], "rb") as f:
        h.update(f.read())
    return h.hexdigest()[:20]


def compile_scenario(scenario, verbose: bool = False):
    """-> (ctypes library exporting the C ABI for this scenario, path).  Compiles on first use, then cached."""
    d = len(list(scenario.x_u_var))
    if d > MAX_D:
        raise Exception('Scenario "%s": n + m = %d exceeds the %d components a run-time compiled model holds' % (
            scenario.name, d, MAX_D))
    text = codegen.generate_user_header(scenario)
    key = _digest(text)
    if key in _loaded:
        return _loaded[key]
    folder = os.path.join(JIT_DIR, key)
    lib_path, hdr_path = os.path.join(folder, "model.so"), os.path.join(folder, "model.cuh")
    if not os.path.exists(lib_path):
        nvcc = _build.find_nvcc()
        if nvcc is None:
            raise RuntimeError('curiousrl_b200: scenario "%s" has no precompiled device model and nvcc is not available to '
                               "compile one (there is no CPU fallback)" % scenario.name)
        import fcntl
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, ".lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)                     # one process per GPU may get here at the same time
            try:
                if not os.path.exists(lib_path):
                    with open(hdr_path, "w") as f:
                        f.write(text)
                    tmp = lib_path + ".tmp%d" % os.getpid()
                    cmd = [nvcc] + _build.NVCC_FLAGS + ['-DCRL_USER_HEADER="%s"' % hdr_path, "-shared", "-o", tmp,
                                                        os.path.join(_build.CSRC, "ilqr_user_abi.cu")]
                    proc = subprocess.run(cmd, capture_output=True, text=True)
                    if verbose:
                        print(proc.stderr)
                    if proc.returncode != 0:
                        raise RuntimeError("nvcc failed for scenario %s:\n%s" % (scenario.name, proc.stderr[-4000:]))
                    os.replace(tmp, lib_path)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    lib = ctypes.CDLL(lib_path)
    _native._declare(lib)
    if lib.crl_abi_version() != _native.ABI_VERSION:
        raise RuntimeError("curiousrl_b200: ABI version mismatch in %s" % lib_path)
    _loaded[key] = (lib, lib_path)
    return _loaded[key]
