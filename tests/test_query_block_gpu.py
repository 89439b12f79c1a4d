"""GPU: query-block sharding of a causal prefill (SURVEY 8e / BASELINE configs[4]): the blocks of a partition of [0, T), each computed
with nothing from the others (replicated K / V, locally recomputed Performer prefix, 8-row CNN halo), concatenate to the
unsharded forward."""
import importlib

import pytest
import torch
import transformers

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _module(sea, H, d, T, P, k, nbf=8):
    torch.manual_seed(5)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    m = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV)
    m.benchmarking = True
    return m


@pytest.mark.parametrize('H,d,T,P,k,dtype,world', [(32, 64, 2048, 256, 64, torch.bfloat16, 4), (4, 64, 1000, 64, 16, torch.bfloat16, 3),
                                                   (4, 32, 700, 64, 16, torch.float32, 2), (8, 128, 1024, 128, 32, torch.float32, 4),
                                                   # BASELINE configs[4] / [3] head shapes on the tensor-core path for those head dims
                                                   (32, 128, 2048, 256, 128, torch.bfloat16, 4), (32, 80, 1536, 256, 64, torch.bfloat16, 3)])
def test_query_blocks_concatenate_to_the_full_prefill(sea, H, d, T, P, k, dtype, world):
    par = importlib.import_module(sea.__name__ + '.parallel')
    m = _module(sea, H, d, T, P, k)
    g = torch.Generator().manual_seed(T)
    q = (torch.randn(1, H, T, d, generator=g) * d ** -0.5).to(dtype).to(DEV)
    kk = torch.randn(1, H, T, d, generator=g).to(dtype).to(DEV)
    v = torch.randn(1, H, T, d, generator=g).to(dtype).to(DEV)
    from oracle import sea_oracle as so
    am = so.causal_additive_mask(T, dtype, 1).to(DEV)
    with torch.no_grad():
        full = m(q, kk, v, q, kk, v, q, kk, am, None, None)
        parts, probs = [], []
        covered = 0
        for r in range(world):
            t0, t1 = par.query_block_bounds(T, world, r)
            assert t0 == covered
            covered = t1
            if t1 > t0:
                o = m.forward_query_block(q, kk, v, t0, t1)
                assert o.context_layer.shape == (1, t1 - t0, H * d)
                parts.append(o.context_layer)
                probs.append(o.estimated_attention_probs)
        assert covered == T
    torch.cuda.synchronize()
    ctx = torch.cat(parts, dim=1)
    pr = torch.cat(probs, dim=2)
    # the predictor of a block sees exactly the rows the full run sees (8 halo rows cover both dilated convs): same probabilities,
    # hence the same top-k and the same context, up to the summation order inside kernels whose tiling depends on the row count
    torch.testing.assert_close(pr, full.estimated_attention_probs, rtol=1e-4, atol=1e-7)
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-3, atol=1e-5)
    bad = (ctx.float() - full.context_layer.float()).abs() > tol['atol'] + tol['rtol'] * full.context_layer.float().abs()
    assert float(bad.float().mean()) < 1e-3, float(bad.float().mean())


def test_query_block_argument_checks(sea):
    m = _module(sea, 4, 64, 256, 64, 16)
    x = torch.randn(1, 4, 256, 64, device=DEV).bfloat16()
    with pytest.raises(sea.SeaError):
        m.forward_query_block(x, x, x, 128, 128)
    with pytest.raises(sea.SeaError):
        m.forward_query_block(x, x, x, 0, 300)


@pytest.mark.parametrize('H,d,T,P,k', [(32, 80, 16384, 256, 64), (32, 128, 8192, 256, 128)])
def test_long_context_fast_path_equals_csr_path(sea, H, d, T, P, k):
    """BASELINE configs[3] (OPT-2.7B head shape at its full T = 16384) and a configs[4] point at sizes the CPU oracle cannot reach:
    the production path (tensor-core Performer / MLP, attention straight from the top-k bits, clamped pixels since T/P = 64 >= k)
    against the CSR-materialising path of the same module, plus the size-independent properties of the CSR: strictly causal
    columns, per-row nnz <= H*k + slack, crow monotone."""
    m = _module(sea, H, d, T, P, k)
    g = torch.Generator().manual_seed(T + d)
    q = (torch.randn(1, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(1, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(1, H, T, d, generator=g).bfloat16().to(DEV)
    with torch.no_grad():
        fast = m(q, kk, v, q, kk, v, q, kk, None, None, None)
        m.output_attentions = True
        slow = m(q, kk, v, q, kk, v, q, kk, None, None, None)
        m.output_attentions = False
    torch.cuda.synchronize()
    assert torch.equal(fast.estimated_attention_probs, slow.estimated_attention_probs)
    torch.testing.assert_close(fast.context_layer.float(), slow.context_layer.float(), rtol=2e-2, atol=2e-2)
    crow, col = slow.partial_attention_mask.crow_indices()[0], slow.partial_attention_mask.col_indices()[0]
    nnz_row = crow[1:] - crow[:-1]
    assert bool((nnz_row >= 0).all()) and int(crow[-1]) == col.numel()
    assert int(nnz_row.max()) <= H * (k + -(-T // P))                                   # the reference's own max_col_z bound (causal_resize_m_to_t.py:946)
    rows = torch.repeat_interleave(torch.arange(T, device=DEV), nnz_row)
    assert bool(((col % T) <= rows).all()) and bool((col // T < H).all())                 # strictly causal, valid head
    assert torch.isfinite(fast.context_layer.float()).all()


def test_query_blocks_with_a_shared_performer_prefix(sea):
    """A rank that walks several blocks computes the linear-attention stage once (performer_prefix) and hands it to every block:
    identical outputs to blocks that recompute their own prefix."""
    H, d, T, P, k = 32, 128, 1536, 256, 64
    m = _module(sea, H, d, T, P, k)
    g = torch.Generator().manual_seed(3)
    q = (torch.randn(1, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(1, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(1, H, T, d, generator=g).bfloat16().to(DEV)
    with torch.no_grad():
        perf = m.performer_prefix(q, kk, v, 1400)
        for t0, t1 in ((0, 300), (300, 1111), (1111, 1400)):
            a = m.forward_query_block(q, kk, v, t0, t1, performer=perf)
            b = m.forward_query_block(q, kk, v, t0, t1)
            assert torch.equal(a.estimated_attention_probs, b.estimated_attention_probs)
            assert torch.equal(a.context_layer, b.context_layer)
        with pytest.raises(sea.SeaError):
            m.forward_query_block(q, kk, v, 1300, 1500, performer=perf)            # the prefix does not cover the block


@pytest.mark.parametrize('H,d,T,P,k,world', [(32, 128, 2048, 256, 128, 4), (32, 64, 1900, 256, 64, 3), (8, 80, 1000, 128, 32, 2)])
def test_exchanged_performer_state_reproduces_the_full_prefill(sea, H, d, T, P, k, world):
    """parallel.performer_exchanged without the network: every "rank" computes the sums of its own rows only, the [2, state] tensors
    are passed around in place of the all-gather, and each rank's blocks, started from sum(total - tail) of the ranks before it,
    concatenate to the unsharded forward."""
    par = importlib.import_module(sea.__name__ + '.parallel')
    m = _module(sea, H, d, T, P, k)
    g = torch.Generator().manual_seed(T)
    q = (torch.randn(1, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(1, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(1, H, T, d, generator=g).bfloat16().to(DEV)
    with torch.no_grad():
        full = m(q, kk, v, q, kk, v, q, kk, None, None, None)
        ranges = [par.contiguous_query_range(T, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == T and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        step1 = [par.performer_range_state(m, q, kk, v, t0, t1, is_last=(r == world - 1)) for r, (t0, t1) in enumerate(ranges)]
        parts, probs = [], []
        for r, (t0, t1) in enumerate(ranges):
            if t1 <= t0:
                continue
            perf = par.performer_range_finish(m, q, kk, v, t0, t1, step1[r][0], [s_[1] for s_ in step1[:r]])
            assert perf[2] == max(t0 - 8, 0) and perf[0].shape[2] == t1 - perf[2]
            o = m.forward_query_block(q, kk, v, t0, t1, performer=perf)
            parts.append(o.context_layer)
            probs.append(o.estimated_attention_probs)
    torch.cuda.synchronize()
    pr = torch.cat(probs, dim=2)
    # the prefix state reaches a row through a different summation order (per-rank totals instead of per-chunk prefixes): bf16-level noise
    # on the Performer output, near-ties of the top-k may move
    err = (pr - full.estimated_attention_probs).abs()
    assert float(err.max()) < 2e-2 and float(err.mean()) < 2e-5, (float(err.max()), float(err.mean()))
    ctx = torch.cat(parts, dim=1)
    bad = (ctx.float() - full.context_layer.float()).abs() > 3e-2 + 3e-2 * full.context_layer.float().abs()
    assert float(bad.float().mean()) < 1e-2, float(bad.float().mean())          # (moved near-ties of the top-k: a few rows pick other pixels)
