"""The cart-pole and vehicle scenarios.

Symbolic definitions (sympy) with the reference classes' constructor arguments, state ordering,
constants, bounds and cost parameters:

* CartPoleSwingUp1 - cart_pole_swingup1.py:16-55  n=4 (angle, angular velocity, position, velocity), m=1 (force);
                     position in [-1, 1], force in [-10, 10]; cost sum_i c_i xu_i^2 with time-varying weights
                     c [T, 5] through `add_param` (terminal weights in the last row)
* CartPoleSwingUp2 - cart_pole_swingup2.py:16-59  n=5 (position, velocity, sin, cos, angular velocity), m=1;
                     the angle is carried as (sin, cos) and re-derived with atan2; weights c [T, 6]
* VehicleTracking  - vehicle_tracking.py:15-46    n=4 (x, y, heading, speed), m=2 (steering, acceleration);
                     rear-axle kinematic model with sqrt / asin; constant quadratic cost, no add_param
* CarParking       - car_parking.py:23-73         the same vehicle, x >= -1; pseudo-Huber costs
                     sqrt(x^2 + p^2) - p switched between a running and a terminal set by `add_param` [T, 2]

These exercise what the three arm scenarios do not: time-varying cost parameters, state-dependent cost
Hessians, finite STATE bounds (the clamp in the rollouts is live), per-action bounds and dense Jacobians.
The device models are generated from exactly these expressions (curiousrl_b200/codegen.py ->
csrc/ilqr_generated.cuh).
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from .dynamic_model import DynamicModelWrapper

_INF = np.inf


def _bounds(rows, constrained):
    rows = np.asarray(rows, dtype=np.float64)
    if not constrained:
        rows = np.tile(np.asarray([[-_INF, _INF]]), (rows.shape[0], 1))
    return rows


def _terminal_schedule(T, running, terminal):
    """add_param [T, p]: `running` for tau < T-1, `terminal` in the last row."""
    table = np.zeros((T, len(running)), dtype=np.float64)
    table[:T - 1] = np.asarray(running, dtype=np.float64)
    table[T - 1] = np.asarray(terminal, dtype=np.float64)
    return table


class CartPoleSwingUp1(DynamicModelWrapper):
    """Pole hanging down at rest (angle pi); swing it up with a bounded force on the cart."""

    def __init__(self, is_with_constraints=True, T=150):
        cart_mass, pole_mass, half_len, h, grav = 1, 0.1, 0.5, 0.02, 9.8
        xu = sp.symbols("x_u:5")
        ang, ang_vel, vel, force = xu[0], xu[1], xu[3], xu[4]
        gamma = (force + pole_mass * half_len * (ang_vel ** 2) * sp.sin(ang)) / (cart_mass + pole_mass)
        ang_acc = (grav * sp.sin(ang) - sp.cos(ang) * gamma) / (
            half_len * ((4 / 3) - ((pole_mass * (sp.cos(ang) ** 2)) / (cart_mass + pole_mass))))
        acc = gamma - (pole_mass * half_len * ang_acc * sp.cos(ang)) / (cart_mass + pole_mass)
        f = sp.Array([ang + h * ang_vel, ang_vel + h * ang_acc, xu[2] + h * vel, vel + h * acc])
        weights = sp.symbols("c:5")
        cost = xu @ np.diag(np.asarray(weights)) @ xu
        super().__init__(dynamic_function=f, x_u_var=xu,
                         constr=_bounds([[-_INF, _INF], [-_INF, _INF], [-1, 1], [-_INF, _INF], [-10, 10]], is_with_constraints),
                         init_state=np.asarray([np.pi, 0, 0, 0], dtype=np.float64).reshape(-1, 1),
                         init_action=np.zeros((T, 1, 1)), obj_fun=cost, add_param_var=weights,
                         add_param=_terminal_schedule(T, (1, 0.1, 1, 1, 0.1), (10000, 1000, 0, 0, 0)))


class CartPoleSwingUp2(DynamicModelWrapper):
    """The same plant with the angle carried as (sin, cos): no wrap-around in the cost."""

    def __init__(self, is_with_constraints=True, T=150):
        h, cart_mass, pole_mass, half_len, grav = 0.02, 1, 0.1, 0.5, 9.8
        xu = sp.symbols("x_u:6")
        s, c, ang_vel, force = xu[2], xu[3], xu[4], xu[5]
        ang_next = sp.atan2(s, c) + h * ang_vel
        gamma = (force + pole_mass * half_len * (ang_vel ** 2) * s) / (cart_mass + pole_mass)
        ang_acc = (grav * s - c * gamma) / (half_len * ((4 / 3) - ((pole_mass * (c ** 2)) / (cart_mass + pole_mass))))
        acc = gamma - (pole_mass * half_len * ang_acc * c) / (cart_mass + pole_mass)
        f = sp.Array([xu[0] + h * xu[1], xu[1] + h * acc, sp.sin(ang_next), sp.cos(ang_next), ang_vel + h * ang_acc])
        weights = sp.symbols("c:6")
        ref = np.asarray([0, 0, 0, 1, 0, 0])
        cost = (xu - ref) @ np.diag(np.asarray(weights)) @ (xu - ref)
        super().__init__(dynamic_function=f, x_u_var=xu,
                         constr=_bounds([[-1, 1]] + [[-_INF, _INF]] * 4 + [[-10, 10]], is_with_constraints),
                         init_state=np.asarray([0, 0, 0.001, -1, 0], dtype=np.float64).reshape(-1, 1),
                         init_action=np.zeros((T, 1, 1)), obj_fun=cost, add_param_var=weights,
                         add_param=_terminal_schedule(T, (0.1, 0.1, 1, 1, 0.1, 0.1), (0, 0, 10000, 10000, 1000, 0)))


def _rear_axle_vehicle(xu):
    """x' = x + b cos(heading), y' = y + b sin(heading), heading' = heading + asin(h/d v sin(steer)),
    v' = v + h a, with b = d + h v cos(steer) - sqrt(d^2 - h^2 v^2 sin^2(steer)) (wheel base d = 3, h = 0.1)."""
    h, d = 0.1, 3
    h_over_d = h / d
    speed, steer = xu[3], xu[4]
    b = d + h * speed * sp.cos(steer) - sp.sqrt(d ** 2 - (h ** 2) * (speed ** 2) * (sp.sin(steer) ** 2))
    return sp.Array([xu[0] + b * sp.cos(xu[2]), xu[1] + b * sp.sin(xu[2]),
                     xu[2] + sp.asin(h_over_d * speed * sp.sin(steer)), speed + h * xu[5]])


class VehicleTracking(DynamicModelWrapper):
    """Track the line y = -10 at speed 8 from rest at the origin, heading +y."""

    def __init__(self, is_with_constraints=True, T=150):
        xu = sp.symbols("x_u:6")
        weights = np.diag([0., 1., 1., 1., 10., 10.])
        ref = np.asarray([0., -10., 0., 8., 0., 0.])
        cost = (xu - ref) @ weights @ (xu - ref)
        super().__init__(dynamic_function=_rear_axle_vehicle(xu), x_u_var=xu,
                         constr=_bounds([[-_INF, _INF]] * 4 + [[-0.6, 0.6], [-3, 3]], is_with_constraints),
                         init_state=np.asarray([0, 0, np.pi / 2, 0], dtype=np.float64).reshape(-1, 1),
                         init_action=np.zeros((T, 2, 1)), obj_fun=cost)


def _pseudo_huber(x, p):
    return sp.sqrt((x ** 2) + (p ** 2)) - p


class CarParking(DynamicModelWrapper):
    """Park at the origin, heading 0, at rest; the wall x >= -1 is a state bound."""

    def __init__(self, is_with_constraints=True, T=500):
        xu = sp.symbols("x_u:6")
        switch = sp.symbols("s:2")
        running = _pseudo_huber(xu[0], 0.01) + _pseudo_huber(xu[1], 0.01) + _pseudo_huber(xu[2], 0.01)
        terminal = (1000 * _pseudo_huber(xu[0], 0.01) + 1000 * _pseudo_huber(xu[1], 0.01)
                    + 1000 * _pseudo_huber(xu[2], 0.01) + 100 * _pseudo_huber(xu[3], 0.01))
        effort = xu[4] ** 2 + xu[5] ** 2
        cost = switch[0] * running + switch[1] * terminal + effort
        super().__init__(dynamic_function=_rear_axle_vehicle(xu), x_u_var=xu,
                         constr=_bounds([[-1, _INF]] + [[-_INF, _INF]] * 3 + [[-0.6, 0.6], [-3, 3]], is_with_constraints),
                         init_state=np.asarray([1, -1, np.pi / 2, 0], dtype=np.float64).reshape(-1, 1),
                         init_action=np.zeros((T, 2, 1)), obj_fun=cost, add_param_var=switch,
                         add_param=_terminal_schedule(T, (1., 0.), (0., 1.)))
