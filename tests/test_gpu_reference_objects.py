"""`init(scenario)` takes scenario objects that are NOT curiousrl_b200 classes ("keep the
reference class"): what is read is the reference's property surface (CuriousRL/scenario/dynamic_model/dynamic_model.py:55-101,
scen_wrapper.py:2-16), the device model is chosen by `scenario.name` and checked against the object's own sympy f and cost.

* a stand-in object: a plain class with the reference's property names and nothing else;
* the reference's OWN scenario objects, where a reference tree is on the machine (baseline/_ref/ travels to the GPU box:
  oracle/vendor_reference.sh) - solved through the reference's call sequence and compared with this package's classes.
"""
import numpy as np
import pytest
import torch

from conftest import make_scenario
from oracle import ref_shim

pytestmark = pytest.mark.gpu


class _Foreign:
    """The reference's scenario surface, re-typed: no curiousrl_b200 base class, no extra attributes."""

    def __init__(self, name, src):
        self._name = name
        self._f, self._xu, self._constr = src.dynamic_function, src.x_u_var, np.array(src.constr, dtype=np.float64)
        self._x0, self._u0 = np.array(src.init_state), np.array(src.init_action)
        self._l, self._pv, self._p = src.obj_fun, src.add_param_var, None if src.add_param is None else np.array(src.add_param)
        self._n, self._m, self._T = int(src.n), int(src.m), int(src.T)

    def with_model(self): return True
    def is_action_discrete(self): return False
    def is_output_image(self): return False
    def play(self): pass
    dynamic_function = property(lambda s: s._f)
    x_u_var = property(lambda s: s._xu)
    constr = property(lambda s: s._constr)
    init_state = property(lambda s: s._x0)
    init_action = property(lambda s: s._u0)
    obj_fun = property(lambda s: s._l)
    add_param_var = property(lambda s: s._pv)
    add_param = property(lambda s: s._p)
    n = property(lambda s: s._n)
    m = property(lambda s: s._m)
    T = property(lambda s: s._T)
    name = property(lambda s: s._name)


def _algos():
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, LogBarrieriLQR
    logger.set_level(45)
    return BasiciLQR, LogBarrieriLQR


@pytest.mark.parametrize("model,T", [("TwoLinkPlanarManipulator", 40), ("ThreeLinkPlanarManipulator", 40),
                                     ("RoboticArmTracking", 40), ("VehicleTracking", 40)])
def test_stand_in_scenario_object(model, T):
    BasiciLQR, LogBarrieriLQR = _algos()
    own = make_scenario(model, T)
    foreign = _Foreign(model, own)
    a, b = BasiciLQR(max_line_search=10, max_iter=30).init(foreign), BasiciLQR(max_line_search=10, max_iter=30).init(own)
    a.solve()
    b.solve()
    assert np.array_equal(a._trajectory, b._trajectory) and a.get_obj_fun_value() == b.get_obj_fun_value()
    assert a.get_obj_add_param() is foreign.add_param                       # quirk Q8: the scenario's own array
    LogBarrieriLQR(max_line_search=10, max_iter=5).init(foreign)            # verified against the object's own expressions


def test_edited_scenario_of_a_known_name_is_refused():
    """Dispatch is by name; a scenario that keeps the name but changes the dynamics or the cost must not run the compiled
    default expressions silently (LogBarrieriLQR on the generated models skipped this check)."""
    BasiciLQR, LogBarrieriLQR = _algos()
    own = make_scenario("VehicleTracking", 40)
    edited = _Foreign("VehicleTracking", own)
    edited._l = own.obj_fun * 2
    with pytest.raises(Exception, match="obj_fun"):
        BasiciLQR(max_line_search=10).init(edited)
    with pytest.raises(Exception, match="obj_fun"):
        LogBarrieriLQR(max_line_search=10).init(edited)
    edited = _Foreign("VehicleTracking", own)
    edited._f = own.dynamic_function * 1.01
    with pytest.raises(Exception, match="dynamic_function"):
        LogBarrieriLQR(max_line_search=10).init(edited)


@pytest.mark.skipif(not ref_shim.available(), reason="no reference tree on this machine (bash oracle/vendor_reference.sh)")
@pytest.mark.parametrize("model,T", [("TwoLinkPlanarManipulator", 50), ("ThreeLinkPlanarManipulator", 50),
                                     ("RoboticArmTracking", 50), ("CartPoleSwingUp1", 60)])
def test_the_references_own_scenario_objects(model, T):
    """Example/ThreeLinkPlanarManipulatorInteractive.py:8-30 with only the solver import swapped: the reference's scenario
    class, this package's BasiciLQR - same result as with this package's scenario class, and the re-solve sequence works."""
    ref_shim.install()
    import CuriousRL.scenario.dynamic_model as ref_models
    BasiciLQR, LogBarrieriLQR = _algos()
    ref_scenario = getattr(ref_models, model)(T=T)
    own = make_scenario(model, T)
    a = BasiciLQR(max_line_search=10, max_iter=40).init(ref_scenario)
    b = BasiciLQR(max_line_search=10, max_iter=40).init(own)
    a.solve()
    b.solve()
    assert np.array_equal(a._trajectory, b._trajectory) and a.get_obj_fun_value() == b.get_obj_fun_value()
    if ref_scenario.add_param is not None and model != "CartPoleSwingUp1":
        for algo in (a, b):                                               # retarget + reset + re-solve
            ap = algo.get_obj_add_param()
            ap[:, 0], ap[:, 1] = 1.5, -0.5
            algo.set_obj_add_param(ap)
            algo.set_obj_fun_value(np.inf)
            algo.solve()
        assert np.array_equal(a._trajectory, b._trajectory)
        assert ref_scenario.add_param[0, 0] == 1.5                         # in-place edit reached the reference object (Q8)
    LogBarrieriLQR(max_line_search=10, max_iter=3).init(ref_scenario)
