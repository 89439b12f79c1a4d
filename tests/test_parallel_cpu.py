"""CPU, world_size 2, gloo: the sharding / gather plumbing of the multi-GPU path (no collective on the hot path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_batch, q):
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    par = importlib.import_module('sea-attention_b200.parallel')
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full_in = torch.randn(total_batch, 3, 5, generator=g)
        (mine,) = par.shard_batch([full_in], world, rank)
        b, e = par.shard_bounds(total_batch, world, rank)
        assert mine.shape[0] == e - b
        ctx_local = mine * 2.0 + 1.0                      # stand-in for the per-shard forward (no cross-rank dependency)
        full = par.all_gather_context(ctx_local, total_batch)
        ok = torch.equal(full, full_in * 2.0 + 1.0)
        slow = par.max_over_ranks(float(rank + 1), torch.device('cpu'))
        q.put((rank, bool(ok), slow))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('total_batch', [2, 5, 8])
def test_batch_sharding_and_gather_gloo(total_batch):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(slow == 2.0 for _, _, slow in res)


def test_shard_bounds_cover_everything(sea):
    par = __import__('importlib').import_module('sea-attention_b200.parallel')
    for total in (0, 1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [par.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_bounds(4, 2, 2)


class _StubModule:
    """Stands in for PerlinAttention.forward_query_block on the CPU: context row t depends on query row t only."""
    def forward_query_block(self, q, k, v, t0, t1):
        import types
        N, H, T, d = q.shape
        return types.SimpleNamespace(context_layer=(q[:, :, t0:t1] * 2.0 + 1.0).permute(0, 2, 1, 3).reshape(N, t1 - t0, H * d))


def _qb_worker(rank, world, port, T, out_q):
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    par = importlib.import_module('sea-attention_b200.parallel')
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        q = torch.randn(2, 3, T, 4, generator=g)
        mine = par.forward_query_sharded(_StubModule(), q, q, q, world, rank, gather=False)
        t0, t1 = par.query_block_bounds(T, world, rank)
        full = par.forward_query_sharded(_StubModule(), q, q, q, world, rank, gather=True)
        want = (q * 2.0 + 1.0).permute(0, 2, 1, 3).reshape(2, T, 12)
        out_q.put((rank, mine.shape[1] == t1 - t0, bool(torch.equal(full, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('T', [100, 256, 700])
def test_query_block_sharding_and_gather_gloo(T):
    """world_size 2, gloo: query-block bounds tile [0, T), each rank computes only its rows, the optional all-gather rebuilds the
    whole context on every rank (blocks of unequal size, including an empty one when T <= the 128-row alignment)."""
    ctx = mp.get_context('spawn')
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_qb_worker, args=(r, 2, port, T, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out_q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for _, a, b in res), res
