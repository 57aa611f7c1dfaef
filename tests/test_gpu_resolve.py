"""Batched re-solve (receding horizon) and the batched result channel on the GPU.

Reference behaviour: Example/ThreeLinkPlanarManipulatorInteractive.py:18-31 - new target, `set_obj_fun_value(inf)`,
`set_init_state(<current state>)`, `solve()`.  `resolve_batch` does that for every problem of a batch at once; the
check is the CPU oracle re-solving the same problems one by one from the same states (same iteration counts, cost
and trajectory <= 1e-9), both cold (the reference's restart from `init_action`) and warm-started.
"""
import numpy as np
import pytest
import torch

from conftest import make_problem, make_scenario

pytestmark = pytest.mark.gpu

MODEL, T = "ThreeLinkPlanarManipulator", 30


def make_algo(**kw):
    from curiousrl_b200 import logger
    from curiousrl_b200.algorithm.ilqr_solver import BasiciLQR
    logger.set_level(45)
    return BasiciLQR(max_line_search=10, **kw).init(make_scenario(MODEL, T))


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))


def test_resolve_batch_matches_oracle_resolves(golden):
    from oracle import ilqr_numpy as orc          # checker only
    g = golden("ThreeLink_T30")
    prob = make_problem(MODEL, T)
    algo = make_algo()
    x0, tgt = g["batch_x0"][:6], g["batch_target"][:6]
    first = algo.solve_batch(x0, tgt)
    X1 = first.trajectory.cpu().numpy()
    steps = np.array([0, 3, 5, 1, 7, 2])
    new_tgt = tgt[::-1].copy()
    n = algo.dynamic_model.n
    for warm in (False, True):
        res = algo.resolve_batch(first, steps, new_tgt, warm_start=warm).to_host()
        checked = 0
        for b in range(6):
            U0 = None
            if warm:
                U0 = np.zeros((T, X1.shape[2] - n))
                U0[:T - steps[b]] = X1[b, steps[b]:, n:]
            ref = orc.solve(prob, X1[b, steps[b], :n], new_tgt[b], U0=None if U0 is None else U0[:, :, None],
                            max_line_search=10)
            if int(res.iterations[b]) != ref["iters"]:
                continue                          # knife-edge problem (see test_gpu_solve_parity); counted below
            checked += 1
            assert abs(res.obj_fun_value[b] - ref["J"]) <= 1e-9 * abs(ref["J"])
            assert rel(res.trajectory[b], ref["X"][:, :, 0] if ref["X"].ndim == 3 else ref["X"]) <= 1e-8
        assert checked >= 5, "warm=%s: only %d of 6 re-solves took the oracle's iteration count" % (warm, checked)


def test_resolve_at_step_zero_repeats_the_solve(golden):
    """Restarting every problem from step 0 with its own target and the scenario's init_action is the same solve."""
    g = golden("ThreeLink_T30")
    algo = make_algo()
    first = algo.solve_batch(g["batch_x0"], g["batch_target"])
    again = algo.resolve_batch(first, 0, g["batch_target"])
    for f in ("trajectory", "obj_fun_value", "iterations", "converged"):
        assert torch.equal(getattr(first, f), getattr(again, f))
    with pytest.raises(ValueError):
        algo.resolve_batch(first, T, g["batch_target"])
    with pytest.raises(ValueError):
        algo.resolve_batch(algo.solve_batch(g["batch_x0"], g["batch_target"], return_trajectory=False), 0)


def test_solve_batch_publishes_to_the_batch_channel(golden, tmp_path, monkeypatch):
    """solve_batch leaves its result on the logger's batched channel like solve() leaves the trajectory on the JSON
    channel: by reference in memory, as a binary record under logs/<folder>/stream/ with set_is_save_json(True)."""
    from curiousrl_b200 import logger
    from curiousrl_b200.utils.result_stream import ResultStream
    monkeypatch.chdir(tmp_path)
    g = golden("ThreeLink_T30")
    algo = make_algo()
    saved = (logger.IS_SAVE_JSON, logger.FOLDER_NAME)
    try:
        res = algo.solve_batch(g["batch_x0"], g["batch_target"])
        assert logger.read_batch()["trajectory"] is res.trajectory
        logger.set_folder_name("gpu_stream").set_is_save_json(True)
        res = algo.solve_batch(g["batch_x0"], g["batch_target"])
        res2 = algo.resolve_batch(res, 4, g["batch_target"], warm_start=True)
        st = ResultStream("gpu_stream")
        assert len(st) == 2
        rec0, rec1 = logger.read_batch(no_iter=0), logger.read_batch()
        assert np.array_equal(rec0["trajectory"], res.trajectory.cpu().numpy())
        assert np.array_equal(rec1["iterations"], res2.iterations.cpu().numpy())
        assert rec0["meta"]["solver"] == "BasiciLQR" and rec0["meta"]["T"] == T
        one = np.asarray(st.trajectory_as_json(0, index=2)["trajectory"])
        assert one.shape == (T, res.trajectory.shape[2], 1)
        assert np.array_equal(one[:, :, 0], res.trajectory[2].cpu().numpy())
    finally:
        logger.IS_SAVE_JSON, logger.FOLDER_NAME = saved
