"""World-size-2 CPU tests (gloo) of the multi-rank host logic: contiguous
env-index shards with seeds base + GLOBAL index, and the exact int64 all-reduce of the episode counters.
The per-shard numbers come from the oracle here (no GPU); tests/test_gpu_scale.py checks on the GPU
that a sharded batch equals the unsharded one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_delivery_b200.sharding import reduce_episode_stats, shard_range
from oracle import oracle_env as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _counters(envs, actions):
    """episodes, moves, on_time, late, pickups, steps, eplen, starved -- what mdx_episode_stats reports."""
    c = np.zeros(8, np.int64)
    for s in range(actions.shape[0]):
        for n, e in enumerate(envs):
            before = list(e.rcell), list(e.carry)
            t0 = e.t
            r, d = e.step(actions[s, n])
            moved = sum(a != b for a, b in zip(before[0], e.rcell))
            deliveries = sum(1 for a, b in zip(before[1], e.carry) if a and not b)
            ontime = sum(1 for a, b in zip(before[1], e.carry) if a and not b and t0 <= e.p_deadline[a - 1])
            c += [int(d), moved, ontime, deliveries - ontime, sum(1 for a, b in zip(before[1], e.carry) if b and not a), 1,
                  e.t if d else 0, 0]
            if d:
                e.reset()
    return c


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    grid = orc.parse_map_text("\n".join(" ".join("1" if (r in (0, 6) or c in (0, 6)) else "0" for c in range(7)) for r in range(7)))
    envs = [orc.OracleEnv(grid, 20, 3, 6, seed=100 + g) for g in range(lo, hi)]
    for e in envs:
        e.reset()
    acts = np.random.RandomState(9).randint(0, 15, size=(50, n_total, 3))[:, lo:hi]
    local = torch.from_numpy(_counters(envs, acts))
    total = reduce_episode_stats(local.clone())
    if rank == 0:
        out.put((local.tolist(), total.tolist()))
    else:
        out.put((local.tolist(), None))
    dist.destroy_process_group()


def test_two_rank_sharding_and_stat_reduction():
    n_total, world = 10, 2
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_range(65536, r, 8) for r in range(8)][3] == (24576, 32768)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    locals_ = [np.array(r[0]) for r in res]
    total = next(np.array(r[1]) for r in res if r[1] is not None)
    assert np.array_equal(total, locals_[0] + locals_[1])
    # the same 10 envs unsharded give the same totals: results do not depend on the number of ranks
    grid = orc.parse_map_text("\n".join(" ".join("1" if (r in (0, 6) or c in (0, 6)) else "0" for c in range(7)) for r in range(7)))
    envs = [orc.OracleEnv(grid, 20, 3, 6, seed=100 + g) for g in range(n_total)]
    for e in envs:
        e.reset()
    acts = np.random.RandomState(9).randint(0, 15, size=(50, n_total, 3))
    assert np.array_equal(_counters(envs, acts), total)
    assert total[5] == 50 * n_total and total[0] >= 2 * n_total
