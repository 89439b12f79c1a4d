"""Multi-GPU plumbing for the SEA attention hot path.

The path shards with NO exchange step: every stage is independent across batch items, and -- for long contexts --
across query blocks given replicated K/V.  One process per GPU; `torch.distributed` (NCCL over NVLink/NVSwitch on the
GPU box, gloo in the CPU tests) is used (a) for barriers / max-over-ranks timing, (b) when a caller asks for
the full context tensor on every rank (`all_gather_context`), and (c) for the one exchange a sharded long-context prefill has: the
exclusive scan of the Performer state sums across ranks (`performer_exchanged`, an all-gather of ~3 MB per rank), which replaces
every rank's recomputation of the prefix before its rows.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `total` units (batch items or query rows) owned by `rank`;
    the first `total % world_size` ranks get one extra unit."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank {rank} / world size {world_size}')
    base, rem = divmod(total, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(tensors, world_size: int, rank: int):
    """Batch sharding: slices dim 0 of every tensor (q, k, v, masks ...) to this rank's items."""
    n = tensors[0].shape[0]
    b, e = shard_bounds(n, world_size, rank)
    return [t[b:e] for t in tensors]


def all_gather_context(ctx_local: torch.Tensor, total_batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Assembles the full context_layer [N, T, H*d] on every rank from batch shards of possibly unequal size.
    The only collective the path ever needs, and only on request (NCCL all-gather over NVLink on the GPU box)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total_batch, world, r) for r in range(world)]
    max_n = max(e - b for b, e in sizes)
    pad = torch.zeros((max_n,) + tuple(ctx_local.shape[1:]), dtype=ctx_local.dtype, device=ctx_local.device)
    pad[: ctx_local.shape[0]] = ctx_local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def query_block_bounds(T: int, world_size: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Query rows [t0, t1) of `rank` for a long-context prefill sharded by query block: contiguous blocks of equal
    size rounded to `align` rows (the row-block of the attention kernels); causal nnz per row is ~ H * k, so equal blocks balance
    the sparse stages."""
    per = -(-T // world_size)
    per = -(-per // align) * align
    t0 = min(rank * per, T)
    return t0, min(t0 + per, T)


def forward_query_sharded(module, q, k, v, world_size: int, rank: int, gather: bool = False, group: Optional[dist.ProcessGroup] = None):
    """One rank's share of a query-block sharded prefill: q, k, v [N,H,T,d] replicated on every rank; returns this rank's context
    rows [N, t1-t0, H*d], or -- gather=True -- the whole [N,T,H*d] on every rank (one all-gather over NVLink / NVSwitch, the only
    collective; ranks whose block is empty contribute nothing)."""
    T = q.shape[2]
    t0, t1 = query_block_bounds(T, world_size, rank)
    if t1 > t0:
        ctx = module.forward_query_block(q, k, v, t0, t1).context_layer          # (one block per rank: the prefix is computed inside)
    else:
        ctx = torch.zeros((q.shape[0], 0, q.shape[1] * q.shape[3]), dtype=q.dtype, device=q.device)
    if not gather:
        return ctx
    per = query_block_bounds(T, world_size, 0)[1]
    pad = torch.zeros((ctx.shape[0], per, ctx.shape[2]), dtype=ctx.dtype, device=ctx.device)
    pad[:, : ctx.shape[1]] = ctx
    bufs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(bufs, pad, group=group)
    parts = []
    for r in range(world_size):
        b, e = query_block_bounds(T, world_size, r)
        parts.append(bufs[r][:, : e - b])
    return torch.cat(parts, dim=1)


def contiguous_query_range(T: int, world_size: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Rank `rank`'s CONTIGUOUS share [t0, t1) of T query rows, boundaries on multiples of `align`.  For long contexts the attention
    is the O(T k) gather kernel -- a constant number of entries per row -- so equal contiguous ranges are balanced."""
    blocks = -(-T // align)
    b0, b1 = shard_bounds(blocks, world_size, rank)
    return min(b0 * align, T), min(b1 * align, T)


def performer_range_state(module, q, k, v, t0: int, t1: int, is_last: bool):
    """Step 1 of performer_exchanged for the rows [t0, t1) of one rank: the chunk sums of its range [h0, t1) (h0 = t0 - halo) and the
    [2, state] tensor the ranks exchange -- row 0 the sums of the range, row 1 the sums of its last `halo` rows (zero for the last rank),
    which the next rank's range overlaps.  -> (workspace or None, state, h0)."""
    from . import ops
    halo = 4 * len(module._cnn_convs()[0])
    h0 = max(t0 - halo, 0)
    w = module._weights_fp32()
    pos, proj = w['pos'], w['proj']
    if t1 <= t0:
        n = int(ops._lib.load().sea_performer_mma_state_floats(q.shape[0], q.shape[1], q.shape[3], proj.shape[0]))
        return None, torch.zeros((2, n), dtype=torch.float32, device=q.device), h0
    ws, total = ops.performer_causal_range_sums(k[:, :, h0:t1], v[:, :, h0:t1], pos[h0:t1], proj)
    if not is_last and t1 - halo >= h0:
        _, tail = ops.performer_causal_range_sums(k[:, :, t1 - halo:t1], v[:, :, t1 - halo:t1], pos[t1 - halo:t1], proj)
    else:
        tail = torch.zeros_like(total)
    return ws, torch.stack([total, tail]), h0


def performer_range_finish(module, q, k, v, t0: int, t1: int, ws, states_before: List[torch.Tensor]):
    """Step 2: prefix from state(h0) = sum over the ranks before of (total - tail), then the outputs of [h0, t1).
    -> (ctx [N,H,t1-h0,2d], cumavg [N,H,t1-h0,d], h0), the `performer=` argument of PerlinAttention.forward_query_block."""
    from . import ops
    halo = 4 * len(module._cnn_convs()[0])
    h0 = max(t0 - halo, 0)
    w = module._weights_fp32()
    init = None
    for st in states_before:
        part = st[0] - st[1]
        init = part if init is None else init + part
    ctx, cumavg = ops.performer_causal_range_out(q[:, :, h0:t1], k[:, :, h0:t1], v[:, :, h0:t1], w['pos'][h0:t1], w['proj'], ws, init, h0)
    return ctx, cumavg, h0


def performer_exchanged(module, q, k, v, world_size: int, rank: int, group: Optional[dist.ProcessGroup] = None, t_range: Optional[Tuple[int, int]] = None):
    """The linear-attention stage of a query-block sharded prefill WITHOUT redundant prefix work ("one tiny exclusive-scan
    exchange of the Performer state"): every rank runs the chunk sums over its own rows [h0, t1) only (h0 = t0 - halo: the predictor CNN of
    its first rows looks `halo` rows back), the ranks all-gather two small state tensors -- the sums of their range and of its last `halo`
    rows, which the next rank's range overlaps -- and each rank starts its prefix from the sums of everything before h0:
        state(h0_r) = sum_{r' < r} (total_r' - tail_r')            (S, z and the running sum of v; fp32)
    This is the one collective of the path (NCCL all-gather of ~3 MB per rank over NVLink / NVSwitch); K and V stay replicated.
    -> ((ctx, cumavg, h0) or None for an empty range, (t0, t1))."""
    T = q.shape[2]
    t0, t1 = contiguous_query_range(T, world_size, rank) if t_range is None else t_range
    ws, state, h0 = performer_range_state(module, q, k, v, t0, t1, is_last=rank == world_size - 1)
    if world_size > 1:
        gathered = [torch.empty_like(state) for _ in range(world_size)]
        dist.all_gather(gathered, state, group=group)
    else:
        gathered = [state]
    if t1 <= t0:
        return None, (t0, t1)
    return performer_range_finish(module, q, k, v, t0, t1, ws, gathered[:rank]), (t0, t1)


def max_over_ranks(value: float, device, group: Optional[dist.ProcessGroup] = None) -> float:
    """Device-side timing reduction used by bench.py: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
