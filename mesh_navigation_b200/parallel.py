"""One-goal-per-GPU sharding of batched multi-goal potential fields.

The path shards perfectly across queries (independent goals, the map is replicated read-only on
every GPU) and not at all inside one query, so the only collective is the final gather of the
potential arrays.  One process per GPU; `torch.distributed` (NCCL over NVLink/NVSwitch on the GPU
box, gloo in the CPU tests) is plumbing only -- there is no compute step that a collective follows
tile by tile, hence nothing to fuse.

goal k -> rank k mod N.  Each rank computes its goals in chunks; the all-gather of chunk i runs
asynchronously while chunk i+1 is being computed.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def shard_indices(n_goals: int, rank: int, world: int) -> np.ndarray:
    """indices of the goals owned by `rank` (goal k -> rank k mod world)"""
    return np.arange(rank, n_goals, world, dtype=np.int64)


def goals_per_rank(n_goals: int, world: int) -> List[int]:
    return [len(range(r, n_goals, world)) for r in range(world)]


def unshard_order(n_goals: int, world: int) -> np.ndarray:
    """row permutation that maps the rank-major gathered buffer back to goal order"""
    per = goals_per_rank(n_goals, world)
    pad = max(per)
    order = np.empty(n_goals, dtype=np.int64)
    for r in range(world):
        idx = shard_indices(n_goals, r, world)
        order[idx] = r * pad + np.arange(len(idx))
    return order


def sharded_potentials(compute_chunk: Callable[[np.ndarray, "object"], None], n_goals: int, V: int, *, rank: int,
                       world: int, device, chunk: int = 64, dist=None, torch=None):
    """Run `compute_chunk(goal_indices, out_tensor)` for this rank's goals and all-gather every rank's
    fields.  Returns a [n_goals, V] float32 tensor in goal order on `device` (every rank gets all fields).

    compute_chunk fills out_tensor[:len(goal_indices)] (float32, [chunk, V], on `device`) with the
    potential fields of the given global goal indices.
    """
    if torch is None:
        import torch as _t
        torch = _t
    mine = shard_indices(n_goals, rank, world)
    pad = max(goals_per_rank(n_goals, world))
    local = torch.empty((pad, V), dtype=torch.float32, device=device)
    if pad > len(mine):
        local[len(mine):].fill_(float("inf"))                 # padding rows of the ragged last shard
    if world == 1:
        for s in range(0, len(mine), chunk):
            idx = mine[s:s + chunk]
            compute_chunk(idx, local[s:s + len(idx)])
        return local[:n_goals]
    gathered = torch.empty((world * pad, V), dtype=torch.float32, device=device)
    views = [gathered[r * pad:(r + 1) * pad] for r in range(world)]
    pending = []
    # chunk boundaries are identical on every rank (pad is global) so the collectives line up
    for s in range(0, pad, chunk):
        e = min(pad, s + chunk)
        idx = mine[s:min(e, len(mine))]
        if len(idx):
            compute_chunk(idx, local[s:s + len(idx)])
        outs = [v[s:e] for v in views]
        pending.append(dist.all_gather(outs, local[s:e], async_op=True))   # overlaps with the next chunk
    for p in pending:
        p.wait()
    # rank-major buffer -> goal order without a copy: row (i, r) of the [pad, world, V] view is goal i*world + r
    return gathered.view(world, pad, V).transpose(0, 1)


def goal_order_rows(result, n_goals: int, V: int):
    """materialise the result of sharded_potentials as a contiguous [n_goals, V] tensor (tests / consumers)"""
    if result.dim() == 2:
        return result[:n_goals]
    return result.reshape(-1, V)[:n_goals]
