import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(tag):
    return np.load(os.path.join(GOLDEN_DIR, tag + ".npz"))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = load_golden(tag)
        return cache[tag]

    return get


_SCENARIOS = {}


def make_scenario(model, T):
    """curiousrl_b200 scenario objects, cached (sympy construction costs ~0.3 s)."""
    key = (model, T)
    if key not in _SCENARIOS:
        import curiousrl_b200.scenario.dynamic_model as models
        _SCENARIOS[key] = getattr(models, model)(T=T)
    return _SCENARIOS[key]


_PROBLEMS = {}


def make_problem(model, T, use_numba=None):
    """Oracle Problem objects, cached (symbolic differentiation costs ~1 s).

    RoboticArm defaults to the numba substrate: CPython evaluates the `x**2` of the
    lambdified dynamics with libm pow(), numba (what the reference runs) with x*x, and
    the two differ by 1 ulp in a few Jacobian entries - enough to break bit equality.
    """
    if use_numba is None:
        use_numba = model == "RoboticArmTracking"
    key = (model, T, use_numba)
    if key not in _PROBLEMS:
        from oracle import ilqr_numpy
        _PROBLEMS[key] = ilqr_numpy.Problem(make_scenario(model, T), use_numba=use_numba)
    return _PROBLEMS[key]
