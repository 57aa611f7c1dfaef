"""Import shim that lets the UNMODIFIED reference run in the build container.  TEST INFRASTRUCTURE.

Used by the fixture generators (`oracle/make_golden*.py`), by `tests/test_oracle_vs_reference.py` (skipped where no
reference tree is found) and by `oracle/ref_runner.py` (the CPU arm of bench.py).  The reference is looked for at
$CURIOUSRL_REFERENCE, `/root/reference` (build container only) and `baseline/_ref/` (the install made by
`oracle/vendor_reference.sh`, which travels to the GPU box).  Two things stop the reference from importing here
(SURVEY.md section 8(c)): matplotlib is not installed, and sympy >= 1.13's strict
printers reject N-dim arrays in `lambdify`.  Neither touches the numerics.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")      # oracle/vendor_reference.sh (travels to the GPU box)


def _find_root():
    """$CURIOUSRL_REFERENCE, else the read-only tree of the build container, else the vendored install."""
    for cand in (os.environ.get("CURIOUSRL_REFERENCE"), "/root/reference", VENDORED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "CuriousRL")):
            return cand
    return os.environ.get("CURIOUSRL_REFERENCE", "/root/reference")


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CuriousRL"))


def vendored() -> bool:
    return os.path.isdir(os.path.join(VENDORED_ROOT, "CuriousRL"))


def install(root=None):
    """Patch sys.modules / sympy.lambdify and put the reference on sys.path (idempotent)."""
    global REFERENCE_ROOT
    if root is not None:
        REFERENCE_ROOT = root
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
