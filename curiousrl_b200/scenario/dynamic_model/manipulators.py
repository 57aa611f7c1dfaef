"""The three planar-arm tracking scenarios of the batched-iLQR path.

Symbolic definitions (sympy) with the same constructor arguments, state
ordering, constants and cost weights as the reference classes

* TwoLinkPlanarManipulator   - two_link_planar_manipulator.py:14-53   (n=6, m=2)
* ThreeLinkPlanarManipulator - three_link_planar_manipulator.py:15-68 (n=8, m=3)
* RoboticArmTracking         - robotic_arm_tracking.py:15-76          (n=6, m=2)

State layout: (theta_1, dtheta_1, ..., theta_J, dtheta_J, ee_x, ee_y); the two
end-effector entries are *outputs* of the current joint angles (they do not
feed back).  Cost: sum_i w_i (xu_i - r_i)^2 with the target (p0, p1) on the
end-effector entries; `add_param` is the per-time-step target [T, 2].

The CUDA device models in csrc/models.cuh are hand-written from the same
formulas; tests check the two against each other through the oracle.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from .dynamic_model import DynamicModelWrapper
from ...utils.Logger import logger

_H = 0.01  # Euler step, all three models


def _tracking_cost(xu, weights, n_joint):
    """sum_i w_i (xu_i - r_i)^2, r = target on the two end-effector slots."""
    target = sp.symbols("p:2")
    ref = np.zeros(len(xu), dtype=object)
    ref[:] = 0.0
    ref[2 * n_joint] = target[0]
    ref[2 * n_joint + 1] = target[1]
    err = np.asarray(xu, dtype=object) - ref
    return err @ np.diag(np.asarray(weights, dtype=np.float64)) @ err, target


def _box(n, m, u_limit):
    rows = [[-np.inf, np.inf]] * n + [[-u_limit, u_limit]] * m
    return np.asarray(rows, dtype=np.float64)


class _PlanarArm(DynamicModelWrapper):
    """Shared construction + a matplotlib stick-figure replay for the arm scenarios."""

    #: True -> link angles accumulate (theta_1, theta_1+theta_2, ...); False -> absolute angles
    _cumulative_angles = True

    def _finish(self, xu, f, weights, n_joint, u_limit, T, x, y):
        n, m = 2 * n_joint + 2, n_joint
        x0 = np.zeros((n, 1), dtype=np.float64)
        x0[n - 1, 0] = float(sum(self.link_lengths))          # arm stretched along +y
        cost, target = _tracking_cost(xu, weights, n_joint)
        goal = np.hstack([x * np.ones((T, 1)), y * np.ones((T, 1))])
        super().__init__(dynamic_function=f, x_u_var=xu, constr=_box(n, m, u_limit),
                         init_state=x0, init_action=np.zeros((T, m, 1)), obj_fun=cost,
                         add_param_var=target, add_param=goal)

    @property
    def link_lengths(self):
        return tuple(getattr(self, "l%d" % (i + 1)) for i in range(self._n_joint))

    def joint_positions(self, state):
        """Cartesian joint positions [J+1, 2] for one state vector (base at the origin)."""
        pts = np.zeros((self._n_joint + 1, 2))
        ang = 0.0
        for j, length in enumerate(self.link_lengths):
            th = float(state[2 * j])
            ang = ang + th if self._cumulative_angles else th
            pts[j + 1] = pts[j] + length * np.array([np.sin(ang), np.cos(ang)])
        return pts

    def play(self, logger_folder=None, no_iter=-1):
        """Replay the last (or a logged) trajectory as a stick figure.

        Reads the same channel as the reference `play()` implementations
        (`logger.read_from_json(...)["trajectory"]`, e.g. two_link_planar_manipulator.py:66);
        needs matplotlib, which is optional.
        """
        import matplotlib.pyplot as plt
        reach = float(sum(self.link_lengths)) + 1.0
        fig, ax = self.create_plot(xlim=(-reach, reach), ylim=(-reach, reach))
        traj = np.asarray(logger.read_from_json(logger_folder, no_iter)["trajectory"])
        (line,) = ax.plot([], [], "o-", lw=3)
        self._is_interrupted = False
        for t in range(self.T):
            self.play_trajectory_current = traj[t, :, 0]
            pts = self.joint_positions(self.play_trajectory_current)
            line.set_data(pts[:, 0], pts[:, 1])
            fig.canvas.draw()
            plt.pause(0.001)
            if self._is_interrupted:
                return
        self._is_interrupted = True


def _kinematic_chain(n_joint, lengths):
    """Double-integrator joints + forward-kinematics outputs (cumulative angles)."""
    d = 3 * n_joint + 2
    xu = sp.symbols("x_u:%d" % d)
    n = 2 * n_joint + 2
    rows = []
    for j in range(n_joint):
        rows.append(xu[2 * j] + _H * xu[2 * j + 1])
        rows.append(xu[2 * j + 1] + _H * xu[n + j])
    ang, ex, ey = 0, 0, 0
    for j in range(n_joint):
        ang = ang + xu[2 * j]
        ex = ex + lengths[j] * sp.sin(ang)
        ey = ey + lengths[j] * sp.cos(ang)
    rows += [ex, ey]
    return xu, sp.Array(rows)


class TwoLinkPlanarManipulator(_PlanarArm):
    """Two torque-free joints driven by accelerations; target tracking at (x, y)."""
    _n_joint = 2

    def __init__(self, T=100, x=2, y=2):
        self.l1, self.l2 = 1, 2
        xu, f = _kinematic_chain(2, (self.l1, self.l2))
        self._finish(xu, f, [0., 10., 0., 10., 1e4, 1e4, 1., 1.], 2, 100, T, x, y)


class ThreeLinkPlanarManipulator(_PlanarArm):
    """Three-joint version (the widest model of the path: n=8, m=3)."""
    _n_joint = 3

    def __init__(self, T=100, x=2, y=2):
        self.l1, self.l2, self.l3 = 1, 2, 2
        xu, f = _kinematic_chain(3, (self.l1, self.l2, self.l3))
        self._finish(xu, f, [0., 10., 0., 10., 0., 10., 1e4, 1e4, 1., 1., 1.], 3, 100, T, x, y)


class RoboticArmTracking(_PlanarArm):
    """Rigid two-link arm with gravity: ddtheta = H(theta)^-1 (Q(theta, dtheta) + tau)."""
    _n_joint = 2
    _cumulative_angles = False

    def __init__(self, is_with_constraints=True, T=100, x=2, y=2):
        m1, m2, g = 1, 2, 9.8
        self.l1, self.l2 = 1, 2
        l1, l2 = self.l1, self.l2
        xu = sp.symbols("x_u:8")
        th1, w1, th2, w2, _, _, tau1, tau2 = xu
        dl = th1 - th2
        coupling = (1 / 2) * m2 * l1 * l2 * sp.cos(dl)
        inertia = sp.Matrix([[((1 / 3) * m1 + m2) * l1 ** 2, coupling],
                             [coupling, (1 / 3) * m2 * l2 ** 2]])
        inertia_inv = inertia.inv()
        bias = np.asarray([
            [-0.5 * m2 * l1 * l2 * (w2 ** 2) * sp.sin(dl) + 0.5 * m1 * g * l1 * sp.sin(th1) + m2 * g * l1 * sp.sin(th1)],
            [0.5 * m2 * l1 * l2 * (w1 ** 2) * sp.sin(dl) + 0.5 * m2 * g * l2 * sp.sin(th2)]])
        torque = np.asarray([[tau1], [tau2]])
        acc = inertia_inv @ bias + inertia_inv @ torque
        f = sp.Array([th1 + _H * w1,
                      w1 + _H * acc[0, 0],
                      th2 + _H * w2,
                      w2 + _H * acc[1, 0],
                      l1 * sp.sin(th1) + l2 * sp.sin(th2),
                      l1 * sp.cos(th1) + l2 * sp.cos(th2)])
        self._finish(xu, f, [0., 10., 0., 10., 1e4, 1e4, 0.01, 0.01], 2,
                     500 if is_with_constraints else np.inf, T, x, y)
