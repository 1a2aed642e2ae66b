"""CPU tests of the N>1 path: world_size-2 gloo run of the goal sharding + all-gather plumbing
(mesh_navigation_b200/parallel.py).  The compute step is a stand-in (there is no CPU product path);
what is tested is goal -> rank assignment, ragged shards, chunked async gathers and goal-order output."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mesh_navigation_b200 import parallel as P


def test_shard_bookkeeping():
    for n, w in [(1024, 8), (10, 4), (7, 2), (3, 8), (1, 1)]:
        seen = np.concatenate([P.shard_indices(n, r, w) for r in range(w)])
        assert sorted(seen.tolist()) == list(range(n))
        assert sum(P.goals_per_rank(n, w)) == n
        order = P.unshard_order(n, w)
        assert len(set(order.tolist())) == n


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_goals, V, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def compute(idx, out):           # stand-in for mnb_cvp_batch on this rank's goals
        calls.append(idx.copy())
        for j, g in enumerate(idx):
            out[j] = torch.arange(V, dtype=torch.float32) * 0.5 + float(g)

    res = P.sharded_potentials(compute, n_goals, V, rank=rank, world=world, device=torch.device("cpu"), chunk=chunk,
                               dist=dist, torch=torch)
    full = P.goal_order_rows(res, n_goals, V)
    expect = torch.arange(V, dtype=torch.float32)[None, :] * 0.5 + torch.arange(n_goals, dtype=torch.float32)[:, None]
    ok = bool(torch.equal(full, expect))
    mine = np.concatenate(calls) if calls else np.zeros(0, np.int64)
    q.put((rank, ok, mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_goals,chunk", [(11, 2), (8, 64), (3, 1), (11, 0), (3, 0)])
def test_sharded_gather_world2(n_goals, chunk):
    world, V = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_goals, V, chunk, q)) for r in range(world)]
    [p.start() for p in procs]
    out = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    for rank, ok, mine in out:
        assert ok, f"rank {rank}: gathered fields are not in goal order"
        assert mine == list(range(rank, n_goals, world))
