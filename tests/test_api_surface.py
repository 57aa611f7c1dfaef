"""The reference's Python surface is kept: scenario property bag, solver constructor arguments and
accessors, logger result channel."""
import inspect

import numpy as np
import pytest

from conftest import make_scenario


@pytest.mark.parametrize("model,n,m,T", [("TwoLinkPlanarManipulator", 6, 2, 100), ("ThreeLinkPlanarManipulator", 8, 3, 100),
                                         ("RoboticArmTracking", 6, 2, 100)])
def test_scenario_property_bag(model, n, m, T):
    sc = make_scenario(model, T)
    assert sc.name == model and (sc.n, sc.m, sc.T) == (n, m, T)
    assert sc.with_model() and not sc.is_action_discrete() and not sc.is_output_image()
    assert sc.init_state.shape == (n, 1) and sc.init_action.shape == (T, m, 1) and not sc.init_action.any()
    assert sc.constr.shape == (n + m, 2) and np.all(np.isinf(sc.constr[:n]))
    assert sc.add_param.shape == (T, 2) and np.all(sc.add_param == 2)
    assert len(sc.x_u_var) == n + m and len(sc.add_param_var) == 2
    assert sc.dynamic_function.shape == (n,)
    assert sc.init_state[-1, 0] == sum(sc.link_lengths)


def test_scenario_constructor_arguments():
    import curiousrl_b200.scenario.dynamic_model as models
    sc = models.ThreeLinkPlanarManipulator(T=17, x=-1.5, y=0.25)
    assert sc.T == 17 and np.all(sc.add_param[:, 0] == -1.5) and np.all(sc.add_param[:, 1] == 0.25)
    free = models.RoboticArmTracking(is_with_constraints=False, T=5)
    assert np.all(np.isinf(free.constr)) and np.all(models.RoboticArmTracking(T=5).constr[6:] == [-500, 500])
    assert list(inspect.signature(models.RoboticArmTracking.__init__).parameters)[1:] == ["is_with_constraints", "T", "x", "y"]
    assert list(inspect.signature(models.TwoLinkPlanarManipulator.__init__).parameters)[1:] == ["T", "x", "y"]


def test_solver_constructor_and_accessors_without_gpu():
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm import AlgoWrapper
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR, iLQRWrapper
    logger.set_level(45)
    params = list(inspect.signature(BasiciLQR.__init__).parameters)[1:8]
    assert params == ["max_iter", "is_check_stop", "stopping_criterion", "max_line_search", "gamma", "line_search_method",
                      "stopping_method"]
    sig = inspect.signature(BasiciLQR.__init__).parameters
    assert (sig["max_iter"].default, sig["max_line_search"].default, sig["gamma"].default) == (1000, 50, 0.5)
    assert sig["stopping_criterion"].default == 1e-6 and sig["line_search_method"].default == "vanilla"
    algo = BasiciLQR(max_line_search=10)
    assert isinstance(algo, iLQRWrapper) and isinstance(algo, AlgoWrapper)
    assert algo.kwargs["max_line_search"] == 10 and algo.kwargs["max_iter"] == 1000
    assert algo.get_obj_fun_value() == np.inf
    algo.set_obj_fun_value(3.0)
    assert algo.get_obj_fun_value() == 3.0
    for name in ("init", "solve", "forward_pass", "backward_pass", "set_init_state", "set_obj_add_param",
                 "get_obj_add_param", "solve_batch"):
        assert callable(getattr(algo, name))
    with pytest.raises(ValueError):
        BasiciLQR(line_search_method="armijo")
    with pytest.raises(ValueError):
        BasiciLQR(stopping_method="absolute")


def test_unsupported_scenario_raises_instead_of_falling_back():
    from curiousrl_b200.algorithm.ilqr_solver import DeviceModel
    # a name without a precompiled model and no scenario object to compile one from (curiousrl_b200/jit.py): refused
    with pytest.raises(Exception, match="no precompiled device model"):
        DeviceModel("TwoWheeledRobot", 100, -1.0, 1.0)


def test_logger_result_channel(tmp_path, monkeypatch):
    from curiousrl_b200.utils.Logger import Logger, logger
    assert Logger() is logger                       # singleton
    logger.set_level(45)
    logger.set_is_save_json(False)
    logger.save_to_json(trajectory=[[1.0], [2.0]])
    assert logger.read_from_json()["trajectory"] == [[1.0], [2.0]]
    monkeypatch.chdir(tmp_path)
    logger.set_folder_name("run_a").set_is_use_logger(True).set_is_save_json(True)
    for i in range(3):
        logger.save_to_json(trajectory=[[float(i)]])
    assert logger.read_from_json()["trajectory"] == [[2.0]]
    assert logger.read_from_json("run_a", 1)["trajectory"] == [[1.0]]
    with pytest.raises(Exception):
        logger.read_from_json("run_a", 9)
    with pytest.raises(Exception):
        logger.read_from_json("missing_folder")
    logger.set_is_save_json(False)


def test_synthetic_generator_known_values():
    from curiousrl_b200.utils.synthetic import synthetic_batch
    x0, tgt = synthetic_batch("TwoLinkPlanarManipulator", 8, seed=0)        # SURVEY.md section 8(c) check values
    assert abs(x0[0, 0] - 0.860555661424686) < 1e-15 and abs(x0[0, 1] - 0.726357844699773) < 1e-15
    assert np.allclose(tgt[0], [0.993161608794206, 0.479060083934160], atol=1e-15)
    x3, t3 = synthetic_batch("ThreeLinkPlanarManipulator", 8, seed=0)
    assert abs(x3[0, 1] - 0.230770222962508) < 1e-15 and np.allclose(t3[0], [1.379289511932583, -0.444909015229402], atol=1e-15)
    assert np.allclose(x3[:, 6], np.sin(x3[:, 0]) + 2 * np.sin(x3[:, 0] + x3[:, 2]) + 2 * np.sin(x3[:, 0] + x3[:, 2] + x3[:, 4]))
    xr, _ = synthetic_batch("RoboticArmTracking", 4, seed=1)
    assert np.allclose(xr[:, 5], np.cos(xr[:, 0]) + 2 * np.cos(xr[:, 2]))
    assert np.all(np.hypot(t3[:, 0], t3[:, 1]) <= 5.0)


def test_scheduling_options_reach_the_c_abi_struct():
    """coop_threshold / occupancy are scheduling knobs only: defaulted, forwarded into crl_solver_opts, and the
    ctypes struct has the ABI-v2 layout of include/curiousrl_b200.h (no GPU needed)."""
    import ctypes
    from curiousrl_b200 import _native
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    names = [f[0] for f in _native.SolverOpts._fields_]
    assert names == ["max_iter", "is_check_stop", "stopping_criterion", "max_line_search", "gamma", "line_search_method",
                     "stopping_method", "grid_blocks", "reserved", "coop_threshold", "reserved2"]
    assert ctypes.sizeof(_native.SolverOpts) == 56
    sig = inspect.signature(BasiciLQR.__init__).parameters
    assert sig["coop_threshold"].default == 0 and sig["occupancy"].default == 0
    o = BasiciLQR(max_line_search=10, coop_threshold=-1, occupancy=20)._opts(max_iter=7, is_check_stop=False)
    assert (o.max_iter, o.is_check_stop, o.max_line_search, o.reserved, o.coop_threshold, o.reserved2) == (7, 0, 10, 20, -1, 0)


def test_log_barrier_constructor_matches_reference():
    """LogBarrieriLQR keeps the reference's constructor (log_barrier_ilqr.py:14-22) - same names, order and defaults."""
    import inspect
    from curiousrl_b200.algorithm.ilqr_solver import LogBarrieriLQR
    sig = inspect.signature(LogBarrieriLQR.__init__).parameters
    names = list(sig)[1:9]
    assert names == ["max_iter", "is_check_stop", "stopping_criterion", "max_line_search", "gamma", "t", "line_search_method",
                     "stopping_method"]
    assert sig["max_iter"].default == 1000 and sig["max_line_search"].default == 50 and sig["gamma"].default == 0.5
    assert sig["t"].default == [0.5, 1., 2., 5., 10., 20., 50., 100.]
    assert sig["line_search_method"].default == "feasibility" and sig["stopping_method"].default == "relative"
    for name in ("init", "solve", "forward_pass", "backward_pass", "set_init_state", "set_obj_add_param", "get_obj_add_param",
                 "set_obj_fun_value", "get_obj_fun_value", "solve_batch"):
        assert callable(getattr(LogBarrieriLQR, name))


def test_cart_pole_and_vehicle_scenarios_have_the_reference_surface():
    """The scenario classes behind the generated device models: sizes, box, initial state and cost-parameter schedules
    of the reference classes (cart_pole_swingup1.py:16-55, cart_pole_swingup2.py:16-59, vehicle_tracking.py:15-46,
    car_parking.py:23-73), and the unconstrained variants."""
    import numpy as np
    import curiousrl_b200.scenario.dynamic_model as models
    spec = {"CartPoleSwingUp1": (4, 1, 150, 5), "CartPoleSwingUp2": (5, 1, 150, 6), "VehicleTracking": (4, 2, 150, None),
            "CarParking": (4, 2, 500, 2)}
    for name, (n, m, T, p) in spec.items():
        sc = getattr(models, name)()
        assert (sc.name, sc.n, sc.m, sc.T) == (name, n, m, T)
        assert sc.init_state.shape == (n, 1) and sc.init_action.shape == (T, m, 1) and not sc.init_action.any()
        assert sc.constr.shape == (n + m, 2) and len(sc.x_u_var) == n + m and len(sc.dynamic_function) == n
        assert sc.with_model() and not sc.is_action_discrete() and not sc.is_output_image()
        if p is None:
            assert sc.add_param is None and sc.add_param_var is None
        else:
            assert sc.add_param.shape == (T, p) and len(sc.add_param_var) == p
            assert np.all(sc.add_param[:-1] == sc.add_param[0]) and not np.array_equal(sc.add_param[-1], sc.add_param[0])
        free = getattr(models, name)(is_with_constraints=False, T=7)
        assert free.T == 7 and np.all(np.isinf(free.constr))
    assert np.array_equal(models.CartPoleSwingUp1().constr[2], [-1.0, 1.0])          # the cart position is bounded
    assert np.array_equal(models.CarParking().constr[0], [-1.0, np.inf])             # the wall
    assert np.array_equal(models.VehicleTracking().constr[4:], [[-0.6, 0.6], [-3.0, 3.0]])
