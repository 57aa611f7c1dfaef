"""Multi-GPU path on CPU: world_size-2 gloo run of the shard/gather logic (the only cross-rank step).

Each rank "solves" its contiguous shard with the host build of the device math (stand-in for the
GPU kernel, which needs a device) and the terminal gather must reproduce the unsharded result in
the original batch order - also for batch sizes that do not divide evenly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_partitions_exactly():
    from curiousrl_b200.distributed import shard_range
    for B in (0, 1, 7, 8, 65536, 1048577):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from curiousrl_b200.distributed import gather_results, shard_range
    from curiousrl_b200.utils.synthetic import synthetic_batch
    from hostsim import HostSim
    T = 30
    x0, tgt = synthetic_batch("ThreeLinkPlanarManipulator", batch, seed=3)
    lo, hi = shard_range(batch, rank, world)
    hs = HostSim("ThreeLinkPlanarManipulator")
    res = [hs.solve(x0[b], tgt[b], T) for b in range(lo, hi)]
    J = torch.tensor([r["J"] for r in res], dtype=torch.float64)
    it = torch.tensor([r["iters"] for r in res], dtype=torch.int32)
    cv = torch.tensor([int(r["converged"]) for r in res], dtype=torch.uint8)
    X = torch.tensor(np.stack([r["X"] for r in res])) if res else torch.zeros((0, T, 11), dtype=torch.float64)
    g = gather_results(J, it, cv, batch, trajectory=X)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), J=g.obj_fun_value.numpy(), it=g.iterations.numpy(),
             cv=g.converged.numpy(), X=g.trajectory.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 7])
def test_two_rank_gather_matches_unsharded(tmp_path, batch):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostsim import HostSim, build
    build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    from curiousrl_b200.utils.synthetic import synthetic_batch
    x0, tgt = synthetic_batch("ThreeLinkPlanarManipulator", batch, seed=3)
    hs = HostSim("ThreeLinkPlanarManipulator")
    ref = [hs.solve(x0[b], tgt[b], 30) for b in range(batch)]
    for rank in range(2):
        g = np.load(os.path.join(tmp_path, "rank%d.npz" % rank))
        assert np.array_equal(g["J"], np.array([r["J"] for r in ref]))
        assert np.array_equal(g["it"], np.array([r["iters"] for r in ref], dtype=np.int32))
        assert np.array_equal(g["cv"], np.array([int(r["converged"]) for r in ref], dtype=np.uint8))
        assert np.array_equal(g["X"], np.stack([r["X"] for r in ref]))


# ----------------------------------------------------------------------------- solve_sharded itself, all solver kinds
class _HostAlgo:
    """Stand-in for a solver object on a box without a GPU: `solve_batch` with the product's signature, executed by the
    host build of the device math (generated VehicleTracking model, plain or under the log barrier)."""

    def __init__(self, barrier):
        from types import SimpleNamespace
        from hostsim.generic import GenericHostSim
        import curiousrl_b200.scenario.dynamic_model as models
        self.T = 20
        self.sc = models.VehicleTracking(T=self.T)
        self.hs = GenericHostSim("VehicleTracking", barrier=barrier)
        self.barrier = barrier
        self.ns = SimpleNamespace
        if barrier:
            self._real_dev = object()
        self.t = [0.5, 2.0, 10.0]

    def solve_batch(self, init_state, add_param=None, return_trajectory=True):
        assert add_param is None
        T, rows = self.T, []
        for x0 in np.asarray(init_state):
            if self.barrier:
                st = np.stack([np.stack([t * np.ones(T), self.t[max(s - 1, 0)] * np.ones(T)], axis=1) for s, t in enumerate(self.t)])
                rows.append(self.hs.solve(x0, None, T, max_iter=30, line_search_method=1, stages=st))
            else:
                rows.append(self.hs.solve(x0, None, T, max_iter=30))
        n = len(rows)
        out = self.ns(obj_fun_value=torch.tensor([r["J"] for r in rows], dtype=torch.float64),
                      iterations=torch.tensor([r["iters"] for r in rows], dtype=torch.int32),
                      converged=torch.tensor([int(r["converged"]) for r in rows], dtype=torch.uint8),
                      trajectory=torch.tensor(np.stack([r["X"] for r in rows])) if n and return_trajectory else
                      (torch.zeros((0, T, 6), dtype=torch.float64) if return_trajectory else None))
        if self.barrier:
            out.inner_iterations = (torch.tensor(np.stack([r["inner_iters"] for r in rows]), dtype=torch.int32) if n
                                    else torch.zeros((0, len(self.t)), dtype=torch.int32))
            out.real_obj_fun_value = out.obj_fun_value * 0.5
        return out


def _sharded_worker(rank, world, port, batch, barrier, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from curiousrl_b200.distributed import solve_sharded
    from curiousrl_b200.utils.synthetic import synthetic_init_states
    algo = _HostAlgo(barrier)
    x0 = synthetic_init_states(algo.sc, batch, seed=4)
    g = solve_sharded(algo, x0, None, gather_trajectory=True)
    extra = dict(inner=g.inner_iterations.numpy(), Jreal=g.real_obj_fun_value.numpy()) if barrier else {}
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), J=g.obj_fun_value.numpy(), it=g.iterations.numpy(),
             cv=g.converged.numpy(), X=g.trajectory.numpy(), **extra)
    dist.destroy_process_group()


@pytest.mark.parametrize("barrier", [False, True])
def test_solve_sharded_two_ranks(tmp_path, barrier):
    """solve_sharded with add_param=None (scenario-owned parameters) for a plain solver and for a log-barrier solver
    (per-loop iteration counts and real costs gathered too), ragged batch: every rank ends with the unsharded result."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hostsim.generic import lib
    lib()
    batch = 5
    mp.spawn(_sharded_worker, args=(2, _free_port(), batch, barrier, str(tmp_path)), nprocs=2, join=True)
    from curiousrl_b200.utils.synthetic import synthetic_init_states
    algo = _HostAlgo(barrier)
    ref = algo.solve_batch(synthetic_init_states(algo.sc, batch, seed=4))
    for rank in range(2):
        g = np.load(os.path.join(tmp_path, "rank%d.npz" % rank))
        assert np.array_equal(g["J"], ref.obj_fun_value.numpy()) and np.array_equal(g["it"], ref.iterations.numpy())
        assert np.array_equal(g["cv"], ref.converged.numpy()) and np.array_equal(g["X"], ref.trajectory.numpy())
        if barrier:
            assert np.array_equal(g["inner"], ref.inner_iterations.numpy())
            assert np.array_equal(g["Jreal"], ref.real_obj_fun_value.numpy())
