"""Synthetic batch generator of the benchmark workload.

Random initial joint angles/velocities with the end-effector entries set by the
model's own output map, and a constant target drawn uniformly from the reach
disc.  Draw order is fixed (PCG64, vectorised over the batch) so that every
consumer - GPU bench, oracle, golden fixtures - sees the same problems.
"""
from __future__ import annotations

import numpy as np

#: model name -> (number of joints, link lengths, cumulative angles?)
ARM_GEOMETRY = {
    "TwoLinkPlanarManipulator": (2, (1.0, 2.0), True),
    "ThreeLinkPlanarManipulator": (3, (1.0, 2.0, 2.0), True),
    "RoboticArmTracking": (2, (1.0, 2.0), False),
}


def synthetic_batch(model_name: str, batch: int, seed: int = 0):
    """Return (x0 [B, n], target [B, 2]) in float64."""
    n_joint, lengths, cumulative = ARM_GEOMETRY[model_name]
    rng = np.random.default_rng(seed)
    th = rng.uniform(-np.pi, np.pi, (batch, n_joint))
    thd = rng.uniform(-1.0, 1.0, (batch, n_joint))
    reach = float(sum(lengths))
    r = reach * np.sqrt(rng.uniform(0.0, 1.0, batch))
    phi = rng.uniform(-np.pi, np.pi, batch)
    n = 2 * n_joint + 2
    x0 = np.zeros((batch, n), dtype=np.float64)
    x0[:, 0:2 * n_joint:2] = th
    x0[:, 1:2 * n_joint:2] = thd
    ang = np.cumsum(th, axis=1) if cumulative else th
    x0[:, n - 2] = (np.asarray(lengths) * np.sin(ang)).sum(axis=1)
    x0[:, n - 1] = (np.asarray(lengths) * np.cos(ang)).sum(axis=1)
    target = np.stack([r * np.cos(phi), r * np.sin(phi)], axis=1)
    return x0, target


def synthetic_init_states(scenario, batch: int, seed: int = 0, spread: float = 0.05):
    """Seeded batch for the scenarios without a target (cart-pole / vehicle): the scenario's own initial state
    plus uniform noise of half-width `spread` on every entry, x0 [B, n] float64 (PCG64, one vectorised draw)."""
    rng = np.random.default_rng(seed)
    base = np.asarray(scenario.init_state, dtype=np.float64).reshape(1, -1)
    return base + spread * rng.uniform(-1.0, 1.0, (batch, base.shape[1]))
