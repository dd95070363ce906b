"""CPU, world_size 2, gloo: the data-parallel host logic -- bucketed all-reduce of the flat gradient
arena driven by the engine's ready/done hooks, parameter broadcast, tile partitioning."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200.dist import GradAllReducer, partition


class _Store:
    def __init__(self, total):
        self.total = total
        self.p_arena = torch.zeros(total)
        self.m_arena = torch.zeros(total)
        self.g_arena = torch.zeros(total)
        self.dirty = False


class _FakeEngine:
    def __init__(self, total):
        self.store = _Store(total)
        self.grad_ready_hook = None
        self.grad_done_hook = None


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 10_000
    eng = _FakeEngine(total)
    red = GradAllReducer(eng, bucket_mb=4 * 3000 / (1 << 20))   # 3000-element buckets
    eng.store.p_arena.fill_(float(rank + 1))
    eng.store.m_arena.fill_(float(rank + 5))
    red.broadcast_parameters(0)
    ok = bool((eng.store.p_arena == 1).all() and (eng.store.m_arena == 5).all() and eng.store.dirty)
    # two "backward" passes; layers complete front to back
    for step in range(2):
        g = torch.arange(total, dtype=torch.float32) * (rank + 1) + step
        flushed_before_done = 0
        for (s, e) in ((0, 64), (64, 2500), (2500, 3500), (3500, 9000), (9000, total)):
            eng.store.g_arena[s:e] = g[s:e]
            eng.grad_ready_hook(s, e)
            flushed_before_done = red.start
        eng.grad_done_hook()
        expect = torch.arange(total, dtype=torch.float32) * sum(r + 1 for r in range(world)) + step * world
        ok = ok and torch.equal(eng.store.g_arena, expect) and flushed_before_done >= 3500 and red.start == 0
    lo, hi = partition(121, rank, world)
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == (0, 61) and res[1][2] == (61, 121)
