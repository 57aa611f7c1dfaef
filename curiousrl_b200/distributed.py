"""Data-parallel sharding of the trajectory batch: one process per GPU, no data-path collective.

Trajectories are independent (the reference solves them one at a time through one stateful
object, Example/*.py:18-31), so rank g simply owns the contiguous slice
[g*B/G, (g+1)*B/G) and runs the persistent solve kernel on it.  The only communication is
the terminal gather of per-trajectory cost (8 B), iteration count (4 B) and convergence flag
(1 B), packed into 16-byte records and sent by ONE all_gather; trajectories stay sharded unless asked for.
Backend: NCCL over NVLink on GPUs; gloo works for the same code on CPU tensors (tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


RECORD_BYTES = 16      # one gathered record per trajectory: float64 J | int32 iterations | uint8 converged | 3 pad


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (batch % world) ranks get one extra trajectory."""
    base, extra = divmod(int(batch), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class GatheredResult:
    obj_fun_value: torch.Tensor    # [B] float64, rank-major = original batch order
    iterations: torch.Tensor       # [B] int32
    converged: torch.Tensor        # [B] uint8
    trajectory: Optional[torch.Tensor] = None   # [B, T, d] only if requested
    inner_iterations: Optional[torch.Tensor] = None     # [B, len(t)] int32  (LogBarrieriLQR)
    real_obj_fun_value: Optional[torch.Tensor] = None   # [B] float64         (LogBarrieriLQR: barrier-free cost)


def _all_gather_ragged(local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """all_gather of 1-D/N-D tensors whose leading sizes differ by at most one (padded to the max)."""
    world = len(counts)
    width = max(counts)
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(c == width for c in counts):
        return out
    return torch.cat([out[r * width:r * width + counts[r]] for r in range(world)], dim=0)


def gather_results(obj_fun_value: torch.Tensor, iterations: torch.Tensor, converged: torch.Tensor, batch: int,
                   trajectory: Optional[torch.Tensor] = None, group=None) -> GatheredResult:
    """Terminal gather: every rank ends up with the whole batch's J / iterations / flags."""
    world = dist.get_world_size(group)
    counts = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    # ONE collective for the three per-trajectory scalars: records of 16 bytes (J: 8, iterations: 4, flag: 1, pad: 3)
    n = obj_fun_value.shape[0]
    rec = torch.zeros((n, RECORD_BYTES), dtype=torch.uint8, device=obj_fun_value.device)
    rec[:, 0:8] = obj_fun_value.to(torch.float64).contiguous().view(torch.uint8).reshape(n, 8)
    rec[:, 8:12] = iterations.to(torch.int32).contiguous().view(torch.uint8).reshape(n, 4)
    rec[:, 12] = converged.to(torch.uint8)
    all_rec = _all_gather_ragged(rec, counts, group)
    J = all_rec[:, 0:8].contiguous().view(torch.float64).reshape(-1)
    it = all_rec[:, 8:12].contiguous().view(torch.int32).reshape(-1)
    cv = all_rec[:, 12].contiguous()
    X = _all_gather_ragged(trajectory, counts, group) if trajectory is not None else None
    return GatheredResult(J, it, cv, X)


def solve_sharded(algo, init_state, add_param=None, gather_trajectory: bool = False, group=None, **kw) -> GatheredResult:
    """Solve this rank's slice of a replicated [B, ...] problem set and gather the per-trajectory results.

    `init_state` [B, n] and `add_param` [B, p] / [B, T, p] are the FULL batch (host or device); only the local slice
    is uploaded.  `add_param=None`: every problem uses the scenario's own parameters (the scenarios without a target).
    `algo` is a BasiciLQR or a LogBarrieriLQR; the latter's per-loop iteration counts and barrier-free costs are
    gathered too (4 * len(t) + 8 more bytes per trajectory).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = int(init_state.shape[0])
    lo, hi = shard_range(B, rank, world)
    wants_traj = gather_trajectory or hasattr(algo, "_real_dev")      # the log-barrier solver needs them for its real cost
    res = algo.solve_batch(init_state[lo:hi], None if add_param is None else add_param[lo:hi],
                           return_trajectory=wants_traj, **kw)
    out = gather_results(res.obj_fun_value, res.iterations, res.converged, B,
                         trajectory=res.trajectory if gather_trajectory else None, group=group)
    counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    if getattr(res, "inner_iterations", None) is not None:
        out.inner_iterations = _all_gather_ragged(res.inner_iterations, counts, group)
    if getattr(res, "real_obj_fun_value", None) is not None:
        out.real_obj_fun_value = _all_gather_ragged(res.real_obj_fun_value, counts, group)
    return out
