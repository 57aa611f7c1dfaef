"""Import shim that lets the UNMODIFIED reference run in the build container.  TEST INFRASTRUCTURE.

Used only by `oracle/make_golden.py` (fixture generation) and by the optional
`tests/test_oracle_vs_reference.py`, both of which need `/root/reference`
(absent on the GPU box).  Two things stop the reference from importing here
(SURVEY.md section 8(c)): matplotlib is not installed, and sympy >= 1.13's strict
printers reject N-dim arrays in `lambdify`.  Neither touches the numerics.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CURIOUSRL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CuriousRL"))


def install():
    """Patch sys.modules / sympy.lambdify and put the reference on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np
    import sympy as sp
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        names = ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.transforms")
        for name in names:
            sys.modules.setdefault(name, types.ModuleType(name))
        for name in names[1:]:
            setattr(sys.modules["matplotlib"], name.split(".")[1], sys.modules[name])
    if not getattr(sp.lambdify, "_accepts_ndim_arrays", False):
        plain = sp.lambdify

        def lambdify(args, expr, modules=None, **kw):
            if isinstance(expr, (sp.NDimArray, np.ndarray)):
                expr = expr.tolist()
            return plain(args, expr, modules, **kw)

        lambdify._accepts_ndim_arrays = True
        sp.lambdify = lambdify
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
