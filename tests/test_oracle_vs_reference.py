"""The CPU oracle against the UNMODIFIED reference, run live on problems that are NOT in the committed fixtures.

Needs a reference tree (the build container's /root/reference, or baseline/_ref/ installed by
oracle/vendor_reference.sh); skipped where there is none.  The golden-vector tests (tests/test_oracle_golden.py) pin the
oracle to recorded outputs of the reference; this one closes the loop on fresh inputs - other seeds, other horizons -
so a fixture that was generated wrongly, or an oracle that only memorises the fixtures' cases, would show here.
Bit equality: the oracle runs the reference's own substrate (sympy lambdify, numba, NumPy / LAPACK).
"""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no reference tree on this machine")

CASES = [("TwoLinkPlanarManipulator", 60, 7), ("ThreeLinkPlanarManipulator", 45, 8), ("RoboticArmTracking", 80, 9)]


@pytest.fixture(scope="module")
def reference():
    ref_shim.install()
    from CuriousRL.utils.Logger import logger as ref_logger
    import CuriousRL.scenario.dynamic_model as ref_models
    from CuriousRL.algorithm.ilqr_solver import BasiciLQR
    ref_logger.set_level(45)
    return ref_models, BasiciLQR


@pytest.mark.parametrize("model,T,seed", CASES)
def test_fresh_problems_bit_equal(reference, model, T, seed):
    ref_models, RefBasiciLQR = reference
    import curiousrl_b200.scenario.dynamic_model as models
    from curiousrl_b200.utils.synthetic import synthetic_batch
    from oracle import ilqr_numpy as orc
    algo = RefBasiciLQR(max_line_search=10, max_iter=60)
    algo.init(getattr(ref_models, model)(T=T))
    prob = orc.Problem(getattr(models, model)(T=T), use_numba=True)
    x0s, targets = synthetic_batch(model, 2, seed=seed)
    for b in range(2):
        ap = algo.get_obj_add_param()
        ap[:, 0], ap[:, 1] = targets[b]
        algo.set_obj_add_param(ap)
        algo.set_init_state(x0s[b].reshape(-1, 1))
        algo.set_obj_fun_value(np.inf)
        n_iter = {"n": 0}
        plain = algo.__class__.forward_pass

        def counting(*a, _p=plain, **k):
            n_iter["n"] += 1
            return _p(algo, *a, **k)

        algo.forward_pass = counting
        algo.solve()
        del algo.forward_pass
        mine = orc.solve(prob, x0s[b], targets[b], max_line_search=10, max_iter=60)
        assert mine["iters"] == n_iter["n"]
        assert mine["J"] == float(algo.get_obj_fun_value())
        assert np.array_equal(mine["X"], algo._trajectory[:, :, 0])
