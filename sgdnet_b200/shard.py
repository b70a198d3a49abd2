"""Fold / alpha sharding of cv_sgdnet over ranks.

The fits of one cv_sgdnet call - #alpha full-data fits and #alpha x #folds fold fits (R/cv_sgdnet.R:160-200) - are
arithmetically independent once every fit has its own sampling stream (api.cv_sgdnet, per-fit seeds). One process
per GPU: X and y are replicated, the fits are dealt to the ranks longest-first, each rank runs its share as one
batch of concurrent CTAs (`sgdnet_fit_batch_*`), and ONE collective ends the run: an all_gather of the per-fit
deviance rows (n_lambda doubles per fit). Nothing is exchanged inside the solver loop.

`torch.distributed` is the only plumbing: backend "nccl" on GPUs (bench / production), "gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def assign_longest_first(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy LPT: items sorted by decreasing cost, each to the currently lightest rank. Deterministic (ties by
    index), so every rank computes the same assignment without talking to the others."""
    order = sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for k in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(k)
        load[r] += float(costs[k])
    return [sorted(v) for v in out]


class Shard:
    """What api.cv_sgdnet needs from the process group: my rank, which fits are mine, and the final gather."""

    def __init__(self, rank: int = 0, world: int = 1, device=None):
        self.rank, self.world, self.device = rank, world, device

    @classmethod
    def from_torch(cls):
        """Rank / world of the initialised default process group (single process when there is none)."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                import torch
                dev = None
                if dist.get_backend() == "nccl":
                    dev = torch.device("cuda", torch.cuda.current_device())
                return cls(dist.get_rank(), dist.get_world_size(), dev)
        except ImportError:
            pass
        return cls()

    def mine(self, costs: Sequence[float]) -> List[int]:
        return assign_longest_first(costs, self.world)[self.rank]

    def owner_of(self, costs: Sequence[float]) -> np.ndarray:
        owner = np.zeros(len(costs), dtype=np.int64)
        for r, items in enumerate(assign_longest_first(costs, self.world)):
            owner[items] = r
        return owner

    def all_gather_rows(self, rows: np.ndarray, mine: Sequence[int]) -> np.ndarray:
        """rows: [n_items, width] float64 with this rank's items filled in (others arbitrary). Returns the array with
        every rank's items filled in, on every rank. One all_gather of a dense [n_items, width] block per rank."""
        if self.world == 1:
            return rows
        import torch
        import torch.distributed as dist
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        mask = np.zeros(rows.shape[0], dtype=np.float64)
        mask[list(mine)] = 1.0
        payload = np.concatenate([np.where(np.isnan(rows), 0.0, rows) * mask[:, None],
                                  np.isnan(rows).astype(np.float64) * mask[:, None], mask[:, None]], axis=1)
        t = torch.from_numpy(payload)
        if self.device is not None:
            t = t.to(self.device)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t)
        out = np.full_like(rows, np.nan)
        width = rows.shape[1]
        for part in parts:
            a = part.cpu().numpy()
            own = a[:, -1] > 0.5
            vals = a[:, :width].copy()
            vals[a[:, width:2 * width] > 0.5] = np.nan
            out[own] = vals[own]
        return out

    def broadcast_object(self, obj, src: int):
        if self.world == 1:
            return obj
        import torch.distributed as dist
        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src, device=self.device)
        return box[0]
