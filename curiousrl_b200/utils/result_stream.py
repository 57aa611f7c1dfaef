"""Batched, binary result / replay channel.

The reference's channel is `Logger.save_to_json(trajectory=X.tolist())` once per iteration and
`Logger.read_from_json(folder_name, no_iter)` on the replay side
(/root/reference/CuriousRL/utils/Logger.py:66-99; consumers: every scenario's `play()`, e.g.
two_link_planar_manipulator.py:66): one JSON line of nested `[T][d][1]` lists per record, ~2 ms
per 100-step trajectory - more than the whole numeric iteration, and unusable for 10^6 trajectories.

A `ResultStream` keeps the semantics (append-only records, `read(no_iter)` with -1 = last, a
missing record raises the reference's message) and changes the encoding: a record is a set of
named arrays with a leading batch axis (`trajectory [B, T, d]`, `obj_fun_value [B]`,
`iterations [B]`, `converged [B]`, ...), each stored as its own `.npy` file so that a reader
memory-maps it and touches only the trajectories it looks at.

Layout under `logs/<folder>/stream/`:
    index.jsonl            one line per record: {"record": k, "fields": {name: {"shape", "dtype"}}, "meta": {...}}
    <k>/<name>.npy         the arrays of record k
A record's arrays are written first and its index line last, so a reader never sees a partial
record.  `trajectory_as_json(no_iter, index)` returns trajectory `index` of a record in the
reference's own encoding (`{"trajectory": [T][d][1] lists}`) for consumers written against
`read_from_json`.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np

_INDEX = "index.jsonl"


def _to_numpy(value) -> np.ndarray:
    """torch tensors (any device) and array-likes -> a host ndarray; CUDA tensors go through pinned memory."""
    if hasattr(value, "detach") and hasattr(value, "cpu"):          # torch.Tensor without importing torch here
        t = value.detach()
        if t.is_cuda:
            import torch
            host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            host.copy_(t, non_blocking=False)
            return host.numpy()
        return t.cpu().numpy()
    return np.asarray(value)


class ResultStream:
    def __init__(self, folder_name: str, root: str = "logs"):
        self.path = os.path.join(root, folder_name, "stream")
        self.folder_name = folder_name

    # ------------------------------------------------------------------ index
    def _index_path(self):
        return os.path.join(self.path, _INDEX)

    def _records(self):
        p = self._index_path()
        if not os.path.isfile(p):
            return []
        with open(p) as fh:
            return [json.loads(line) for line in fh if line.strip()]

    def __len__(self):
        return len(self._records())

    # ------------------------------------------------------------------ write
    def append(self, meta: Optional[dict] = None, **arrays) -> int:
        """Append one record; every value needs a leading batch axis of the same length.  Returns its number."""
        if not arrays:
            raise ValueError("a record needs at least one array")
        host = {name: _to_numpy(v) for name, v in arrays.items() if v is not None}
        sizes = {a.shape[0] if a.ndim else None for a in host.values()}
        if None in sizes or len(sizes) != 1:
            raise ValueError("every array of a record needs the same leading batch axis, got %s"
                             % {k: v.shape for k, v in host.items()})
        k = len(self)
        rec_dir = os.path.join(self.path, str(k))
        os.makedirs(rec_dir, exist_ok=True)
        fields = {}
        for name, a in host.items():
            a = np.ascontiguousarray(a)
            np.save(os.path.join(rec_dir, name + ".npy"), a, allow_pickle=False)
            fields[name] = {"shape": list(a.shape), "dtype": str(a.dtype)}
        line = json.dumps({"record": k, "fields": fields, "meta": meta or {}})
        with open(self._index_path(), "a") as fh:                    # the index line publishes the record
            fh.write(line + "\n")
            fh.flush()
            os.fsync(fh.fileno())
        return k

    # ------------------------------------------------------------------ read
    def _resolve(self, no_iter: int) -> dict:
        recs = self._records()
        if not recs:
            raise Exception('No result stream is saved in the log folder "%s". '
                            "Call logger.set_is_save_json(True) before solve_batch()." % self.folder_name)
        if no_iter == -1:
            return recs[-1]
        if no_iter < 0 or no_iter >= len(recs):
            raise Exception("The number of iteration exceeds the maximum iteration number!")   # Logger.py:95
        return recs[no_iter]

    def read(self, no_iter: int = -1, mmap: bool = True) -> Dict[str, np.ndarray]:
        """Record `no_iter` (-1 = the last) as {name: array}; arrays are memory-mapped read-only unless mmap=False."""
        rec = self._resolve(no_iter)
        rec_dir = os.path.join(self.path, str(rec["record"]))
        out = {name: np.load(os.path.join(rec_dir, name + ".npy"), mmap_mode="r" if mmap else None, allow_pickle=False)
               for name in rec["fields"]}
        out["meta"] = rec.get("meta", {})
        return out

    def trajectory_as_json(self, no_iter: int = -1, index: int = 0) -> dict:
        """Trajectory `index` of a record in the reference's encoding: {"trajectory": [T][d][1] nested lists}."""
        rec = self.read(no_iter)
        if "trajectory" not in rec:
            raise Exception("record %d holds no trajectories" % no_iter)
        traj = np.asarray(rec["trajectory"][index], dtype=np.float64)
        return {"trajectory": traj[:, :, None].tolist()}
