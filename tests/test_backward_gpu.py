"""GPU parity tests of the sparse-attention backward: sea_sparse_attention_bits_bwd through the C ABI against
autograd (fp64) of the reference's dense masked-attention expression restricted to the mask (oracle.sparse_attention_grads).
Tolerances: rtol 1e-3 for fp32 inputs, 2e-2 for bf16 (north-star)."""
import pytest
import torch

from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _case(sea, N, H, T_DST, T_SRC, P, k, d, causal, seed):
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand(N, H, T_DST, P, generator=g) < min(1.0, 2.0 * k / P)).float()
    mask[0, 1 % H, 3] = 0                                       # a (row, head) with no alive pixel
    mask[0, 0, 5] = 1                                           # a head with every pixel alive
    crow, col, Z = so.resize_from_m_to_t_csr(mask, k, T_SRC, causal)
    alive = so.flat_csr_to_dense(crow, col, torch.ones(col.shape), T_SRC, H) > 0
    q = torch.randn(N, H, T_DST, d, generator=g) * d ** -0.5
    kk = torch.randn(N, H, T_SRC, d, generator=g)
    v = torch.randn(N, H, T_SRC, d, generator=g)
    scales = torch.randn(N, H, T_DST, 2, generator=g)
    dout = torch.randn(N, T_DST, H * d, generator=g)
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    return bits, alive, q, kk, v, scales, dout


@pytest.mark.parametrize('N,H,T_DST,T_SRC,P,k,d,causal,with_avg', [
    (2, 3, 100, 100, 32, 8, 64, True, True), (1, 2, 96, 96, 32, 8, 32, True, True), (1, 2, 64, 64, 32, 8, 128, True, True),
    (1, 2, 512, 512, 32, 4, 64, True, True),                    # T/P > k: clamped, sub-sampled pixels
    (2, 4, 64, 64, 32, 8, 64, False, False), (1, 3, 40, 128, 64, 8, 64, True, False),
])
def test_sparse_attention_backward_fp32(sea, N, H, T_DST, T_SRC, P, k, d, causal, with_avg):
    bits, alive, q, kk, v, scales, dout = _case(sea, N, H, T_DST, T_SRC, P, k, d, causal, seed=P + d + T_DST)
    out_r, dq_r, dk_r, dv_r, ds_r = so.sparse_attention_grads(alive, q, kk, v, scales, dout, use_scaler=True, with_avg=with_avg)
    avg = None
    if with_avg:
        avg = (torch.cumsum(v, 2) / torch.arange(1, T_SRC + 1).view(1, 1, -1, 1)).to(DEV)
    dq, dk, dv, ds = sea.ops.sparse_attention_from_bits_backward(bits, q.to(DEV), kk.to(DEV), v.to(DEV), scales.to(DEV), avg, dout.to(DEV),
                                                                 P, k, True, causal)
    for mine, ref in ((dq, dq_r), (dk, dk_r), (dv, dv_r)):
        torch.testing.assert_close(mine.cpu(), ref, rtol=1e-3, atol=1e-4)
    if not with_avg:
        ds_r = ds_r.clone(); ds_r[..., 1] = 0                   # without the running-mean branch s1 is unused by the forward
    torch.testing.assert_close(ds.cpu(), ds_r, rtol=1e-3, atol=1e-4)


def test_sparse_attention_backward_no_scaler(sea):
    N, H, T, P, k, d = 1, 2, 80, 32, 8, 64
    bits, alive, q, kk, v, scales, dout = _case(sea, N, H, T, T, P, k, d, True, seed=9)
    _, dq_r, dk_r, dv_r, ds_r = so.sparse_attention_grads(alive, q, kk, v, scales, dout, use_scaler=False, with_avg=True)
    avg = (torch.cumsum(v, 2) / torch.arange(1, T + 1).view(1, 1, -1, 1)).to(DEV)
    dq, dk, dv, ds = sea.ops.sparse_attention_from_bits_backward(bits, q.to(DEV), kk.to(DEV), v.to(DEV), scales.to(DEV), avg, dout.to(DEV), P, k, False, True)
    for mine, ref in ((dq, dq_r), (dk, dk_r), (dv, dv_r), (ds, ds_r)):
        torch.testing.assert_close(mine.cpu(), ref, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('N,H,T,P,k,d', [(1, 4, 200, 32, 8, 64), (1, 32, 256, 256, 64, 64)])
def test_sparse_attention_autograd_bf16(sea, N, H, T, P, k, d):
    """autograd.Function wrapper: bf16 forward (block / gather kernel) + backward vs the fp64 oracle."""
    bits, alive, q, kk, v, scales, dout = _case(sea, N, H, T, T, P, k, d, True, seed=3)
    qb, kb, vb = (t.bfloat16() for t in (q, kk, v))
    out_r, dq_r, dk_r, dv_r, ds_r = so.sparse_attention_grads(alive, qb.float(), kb.float(), vb.float(), scales, dout.bfloat16().float())
    qd, kd, vd = (t.to(DEV).requires_grad_(True) for t in (qb, kb, vb))
    sd = scales.to(DEV).requires_grad_(True)
    avg = (torch.cumsum(vb.float(), 2) / torch.arange(1, T + 1).view(1, 1, -1, 1)).bfloat16().to(DEV)
    out = sea.ops.sparse_attention_from_bits_autograd(bits, qd, kd, vd, sd, avg, P, k)
    torch.testing.assert_close(out.float().cpu(), out_r, rtol=2e-2, atol=2e-2)
    out.backward(dout.bfloat16().to(DEV))
    for mine, ref in ((qd.grad, dq_r), (kd.grad, dk_r), (vd.grad, dv_r), (sd.grad, ds_r)):
        assert mine is not None
        torch.testing.assert_close(mine.float().cpu(), ref, rtol=2e-2, atol=2e-2)


def test_module_forward_backward_through_sparse_path(sea):
    """PerlinAttention.forward with inputs that require grad: context_layer backpropagates into q, k, v through the
    sparse attention (mask and scales are treated as constants of the step)."""
    import transformers
    N, H, d, T, P, k, nbf = 1, 4, 64, 192, 32, 8, 8
    torch.manual_seed(11)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV)
    mk = lambda s: (torch.randn(N, H, T, d, device=DEV) * s).bfloat16().requires_grad_(True)
    q, kk, v = mk(d ** -0.5), mk(1.0), mk(1.0)
    am = so.causal_additive_mask(T, torch.bfloat16, N).to(DEV)
    out = mod(q, kk, v, q, kk, v, q, kk, am, None, None)
    dout = torch.randn_like(out.context_layer)
    out.context_layer.backward(dout)
    for t in (q, kk, v):
        assert t.grad is not None and torch.isfinite(t.grad.float()).all() and float(t.grad.float().abs().sum()) > 0
    # the same numbers as the operator called by hand on the step's own mask / scales
    with torch.no_grad():
        mod.output_attentions = False
        w = mod._weights_fp32()
        ctx, cumavg = sea.ops.performer_causal(q, kk, v, w['pos'], w['proj'])
        cnn_in, scales, _ = sea.ops.predictor_mlp(ctx, v, w, mod.attention_predictor_dec_row_splits, P // mod.attention_predictor_dec_row_down_scale)
    probs = out.estimated_attention_probs
    kpr, _ = mod._shape_consts(H, P, T, T, q.device)
    bits = sea.ops.topk_mask_bits(probs, kpr, 'causal_batch')
    dq, dk, dv, _ = sea.ops.sparse_attention_from_bits_backward(bits, q.detach(), kk.detach(), v.detach(), scales, cumavg, dout, P, k, True, True)
    torch.testing.assert_close(q.grad.float(), dq.bfloat16().float(), rtol=1e-2, atol=1e-3)
    torch.testing.assert_close(v.grad.float(), dv.bfloat16().float(), rtol=1e-2, atol=1e-3)
