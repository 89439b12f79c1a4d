"""Multi-GPU plumbing for the SEA attention hot path.

The path shards with NO exchange step: every stage is independent across batch items, and -- for long contexts --
across query blocks given replicated K/V.  One process per GPU; `torch.distributed` (NCCL over NVLink/NVSwitch on the
GPU box, gloo in the CPU tests) is used only (a) for barriers / max-over-ranks timing and (b) when a caller asks for
the full context tensor on every rank (`all_gather_context`).  Nothing here launches a collective on the hot path.
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


def max_over_ranks(value: float, device, group: Optional[dist.ProcessGroup] = None) -> float:
    """Device-side timing reduction used by bench.py: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
