"""GPU: non-causal (BERT) variant -- stages and the whole layer against the oracle and the reference fixture."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF
import transformers

from conftest import golden_layer
from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('N,H,T,d,F', [(1, 2, 64, 64, 266), (2, 3, 50, 32, 27), (1, 1, 100, 64, 33)])
def test_performer_noncausal_matches_oracle(sea, N, H, T, d, F):
    g = torch.Generator().manual_seed(T + F)
    q = torch.randn(N, H, T, d, generator=g) * d ** -0.5
    k = torch.randn(N, H, T, d, generator=g)
    v = torch.randn(N, H, T, d, generator=g)
    proj = torch.randn(F, d, generator=g)
    ctx = sea.ops.performer_noncausal(q.to(DEV), k.to(DEV), v.to(DEV), proj.to(DEV))
    v2 = torch.cat([so.v_identity_grid(N, H, T, d), v], -1)
    ref = so.performer_noncausal(q, k, v2, proj)
    torch.testing.assert_close(ctx.cpu(), ref, rtol=1e-3, atol=1e-5)


def test_v_identity_closed_form_equals_grid_sample():
    """The closed form used by the kernel equals the reference's F.grid_sample construction (attention.py:462-495)."""
    N, H, T, d = 1, 2, 37, 16
    eye = torch.eye(d).view(1, 1, d, d).expand(N, H, d, d)
    cs = torch.ones(N, 1, 1, T).cumsum(-1)
    ty = ((cs - 1.0) / ((torch.full((N,), float(T)) - 1.0).view(N, 1, 1, 1) + 1e-8) * 2 - 1).view(N, T, 1, 1).expand(N, T, d, 1)
    tx = (torch.arange(d) / (d - 1) * 2 - 1).view(1, 1, d, 1).expand(N, T, d, 1)
    ref = TF.grid_sample(eye, torch.cat([tx, ty], -1), mode='bilinear', align_corners=True, padding_mode='zeros')
    torch.testing.assert_close(so.v_identity_grid(N, H, T, d), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('stride_t,up,relu', [(2, 1, True), (1, 1, True), (1, 2, False)])
def test_conv3x3_cl_matches_torch(sea, stride_t, up, relu):
    N, T, W, C, O = 2, 21, 12, 16, 8
    g = torch.Generator().manual_seed(stride_t * 10 + up)
    x = torch.randn(N, T, W, C, generator=g)
    wt = torch.randn(O, C, 3, 3, generator=g) * 0.1
    b = torch.randn(O, generator=g)
    y = sea.ops.conv3x3_cl(x.to(DEV), wt.to(DEV), b.to(DEV), stride_t=stride_t, up=up, relu=relu)
    xin = x.permute(0, 3, 1, 2)
    if up > 1:
        xin = TF.interpolate(xin, scale_factor=(up, 1), mode='nearest')
    ref = TF.conv2d(xin, wt, b, stride=(stride_t, 1), padding=1)
    if relu:
        ref = torch.relu(ref)
    torch.testing.assert_close(y.cpu().permute(0, 3, 1, 2), ref, rtol=1e-3, atol=1e-4)


def test_bert_tail_matches_torch(sea):
    N, H, Tin, Win, T, P = 2, 3, 18, 16, 17, 32
    y = torch.randn(N, Tin, Win, H, generator=torch.Generator().manual_seed(2))
    probs, scores = sea.ops.bert_tail(y.to(DEV), T, P, want_scores=True)
    ref = TF.interpolate(y.permute(0, 3, 1, 2), (T, P), mode='bilinear')
    torch.testing.assert_close(scores.cpu(), ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(probs.cpu(), torch.softmax(ref, -1), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize('N,H,T,P,k,ties', [(2, 3, 40, 16, 4, False), (1, 4, 64, 32, 8, True), (1, 12, 30, 128, 64, True),
                                           (2, 12, 130, 128, 64, True), (1, 12, 512, 128, 64, False), (1, 2, 9, 32, 64, False)])
def test_topk_batch_bit_exact(sea, N, H, T, P, k, ties):
    g = torch.Generator().manual_seed(P + T)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g), -1)
    if ties:
        probs = probs[..., : P // 4].repeat_interleave(4, dim=-1).contiguous()
    tl_ = torch.full((N,), T, dtype=torch.long)
    ref = so.topk_mask_noncausal(probs, k, 1.0, tl_, 'batch')
    kpi = torch.clamp_min(torch.round(tl_ * H * (k * 1.0 * P / tl_)), 1)
    bits = sea.ops.topk_mask_bits_batch(probs.to(DEV), kpi.to(DEV))                      # multi-CTA selection
    assert torch.equal(sea.ops.bits_to_mask(bits, H, P).cpu(), ref)
    bits1 = sea.ops.topk_mask_bits_batch(probs.to(DEV), kpi.to(DEV), single_cta=True)     # one CTA per item
    assert torch.equal(bits1.cpu(), bits.cpu())


@pytest.mark.parametrize('N,H,T,P,k,ties', [(2, 3, 40, 16, 4, False), (1, 12, 130, 128, 64, True), (2, 4, 64, 32, 8, True)])
def test_topk_head_mode_bit_exact(sea, N, H, T, P, k, ties):
    """k_flatten_dim='head' (attention.py:838-842): one top-k group per (item, head) over its T*P keys."""
    g = torch.Generator().manual_seed(P + T + 1)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g), -1)
    if ties:
        probs = probs[..., : P // 4].repeat_interleave(4, dim=-1).contiguous()
    tl_ = torch.full((N,), T, dtype=torch.long)
    ref = so.topk_mask_noncausal(probs, k, 1.0, tl_, 'head')
    kpi = torch.clamp_min(torch.round(tl_ * (k * 1.0 * P / tl_)), 1).view(N, 1).expand(N, H).reshape(-1)
    bits = sea.ops.topk_mask_bits_batch(probs.to(DEV), kpi.to(DEV), group_heads=1)
    assert torch.equal(sea.ops.bits_to_mask(bits, H, P).cpu(), ref)


def test_bert_avg_matches_oracle(sea):
    N, H, T, P, d = 2, 3, 50, 16, 32
    g = torch.Generator().manual_seed(4)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g), -1)
    v = torch.randn(N, H, T, d, generator=g)
    avg = sea.ops.bert_avg(probs.to(DEV), v.to(DEV))
    wts = so.resize_from_m_to_t_dense(probs.mean(-2, keepdim=True), 0.0, torch.zeros(N, 1, 1, T), target_width=T, is_causal=False)
    ref = (v * wts.transpose(-1, -2)).sum(-2, keepdim=True)
    torch.testing.assert_close(avg.cpu(), ref, rtol=1e-3, atol=1e-5)


def _bert_module(sea, m, sd, k_flatten_dim='batch'):
    cfg = transformers.BertConfig(hidden_size=m['H'] * m['d'], num_attention_heads=m['H'], max_position_embeddings=m['T'])
    pc = sea.PerlinAttentionConfig(performer_nb_factor=m['nbf'], k=m['k'], attention_predictor_length=m['P'], causal=False, k_flatten_dim=k_flatten_dim)
    mod = sea.PerlinAttention(cfg, pc).eval()
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert not unexpected
    mod = mod.to(DEV)
    mod.benchmarking = True
    return mod


def test_bert_layer_matches_reference_fixture_fp32(sea):
    g, m, sd = golden_layer('layer_bert_h4_t64')
    H, T, P, d = m['H'], m['T'], m['P'], m['d']
    mod = _bert_module(sea, m, sd)
    mod.output_attentions = True
    q, k, v = (torch.from_numpy(g[x]).to(DEV) for x in 'qkv')
    mask = torch.zeros(1, 1, 1, T, device=DEV)
    out = mod(q, k, v, q, k, v, q, k, mask, None, None)
    torch.testing.assert_close(out.estimated_attention_probs_m.cpu(), torch.from_numpy(g['sparse.estimated_attention_probs']), rtol=2e-3, atol=1e-6)
    pm = out.partial_attention_mask
    assert np.array_equal(pm.crow_indices().cpu().numpy(), g['sparse.crow'])
    assert np.array_equal(pm.col_indices().cpu().numpy(), g['sparse.col'])
    torch.testing.assert_close(out.partial_attention_probs.values().cpu(), torch.from_numpy(g['sparse.probs_values']), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(out.context_layer.cpu(), torch.from_numpy(g['sparse.context_layer']), rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize('mode', ['batch', 'query', 'causal_batch'])
def test_bert_layer_matches_oracle(sea, mode):
    N, H, d, T, P, k, nbf = 2, 4, 64, 96, 32, 8, 4
    torch.manual_seed(7)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=False, k_flatten_dim=mode)
    mod = sea.PerlinAttention(cfg, pc).eval()
    sd = {k_: v_.detach().clone().float() for k_, v_ in mod.state_dict().items()}
    mod = mod.to(DEV)
    mod.output_attentions = True
    g = torch.Generator().manual_seed(3)
    q, kk, v = (torch.randn(N, H, T, d, generator=g) for _ in range(3))
    qd, kd, vd = q.to(DEV), kk.to(DEV), v.to(DEV)
    out = mod(qd, kd, vd, qd, kd, vd, qd, kd, torch.zeros(N, 1, 1, T, device=DEV), None, None)
    b = so.perlin_forward_noncausal(sd, q, kk, v, k_top=k, P=P, k_flatten_dim=mode, sparse=True, keep_dense=True)
    torch.testing.assert_close(out.estimated_attention_probs.cpu(), b['estimated_attention_probs'], rtol=2e-3, atol=1e-6)
    pm = out.partial_attention_mask
    mine = so.flat_csr_to_dense(pm.crow_indices().cpu().long(), pm.col_indices().cpu().long(), torch.ones(pm.col_indices().shape), T, H).numpy()
    assert float((mine == b['partial_attention_mask'].numpy()).mean()) >= 0.999
    rows_same = torch.from_numpy((mine == b['partial_attention_mask'].numpy()).all(axis=(0, 1, 3)))
    torch.testing.assert_close(out.context_layer.cpu()[:, rows_same], b['context_layer'][:, rows_same], rtol=1e-3, atol=2e-5)
    # bf16: attention straight from the bit mask, BERT mean broadcast over the rows
    mod.output_attentions = False
    qb, kb, vb = qd.bfloat16(), kd.bfloat16(), vd.bfloat16()
    out2 = mod(qb, kb, vb, qb, kb, vb, qb, kb, torch.zeros(N, 1, 1, T, device=DEV, dtype=torch.bfloat16), None, None)
    assert out2.context_layer.dtype == torch.bfloat16 and out2.partial_attention_mask is None
    assert float((out2.context_layer.float().cpu() - b['context_layer']).abs().mean()) < 3e-2


def _padded_mask(lengths, T, dtype=torch.float32):
    fmin = torch.finfo(torch.float16).min / 2 if dtype != torch.float32 else torch.finfo(torch.float32).min / 2
    mask = torch.zeros(len(lengths), 1, 1, T, dtype=dtype)
    for n, L in enumerate(lengths):
        mask[n, :, :, L:] = fmin
    return mask


def test_bert_padded_batch_matches_reference_fixture_fp32(sea):
    """Right-padded non-causal batch (attention.py:401-449, 482, 512-514, 777-778, 837, 1209-1219) against the unmodified
    reference's dense path (tests/golden/layer_bert_padded_h4_t64.npz, lengths 64 / 45 / 23 of T = 64) on the valid query rows, and
    against the oracle's sparse branch bit for bit on the CSR."""
    g, m, sd = golden_layer('layer_bert_padded_h4_t64')
    N, H, T, P, d, k = m['N'], m['H'], m['T'], m['P'], m['d'], m['k']
    lengths = [int(x) for x in g['lengths']]
    mod = _bert_module(sea, m, sd)
    mod.output_attentions = True
    q, kk, v = (torch.from_numpy(g[x]) for x in 'qkv')
    v_in = v.clone().to(DEV)
    out = mod(q.to(DEV), kk.to(DEV), v_in, q.to(DEV), kk.to(DEV), v_in, q.to(DEV), kk.to(DEV), _padded_mask(lengths, T).to(DEV), None, None)
    assert torch.equal(v_in.cpu(), v)                                            # the caller's v is not modified (a masked copy is used)
    valid = torch.arange(T).view(1, T) < torch.tensor(lengths).view(N, 1)
    rows4, rows3 = valid.view(N, 1, T, 1), valid.view(N, T, 1)
    ref_probs = torch.from_numpy(g['dense.estimated_attention_probs'])
    torch.testing.assert_close(out.estimated_attention_probs.cpu() * rows4, ref_probs * rows4, rtol=2e-3, atol=1e-6)
    b = so.perlin_forward_noncausal(sd, q, kk, v, k_top=k, P=P, sparse=True, keep_dense=True, lengths=torch.tensor(lengths))
    pm = out.partial_attention_mask
    crow, col = pm.crow_indices().cpu(), pm.col_indices().cpu()
    mine = so.flat_csr_to_dense(crow.long(), col.long(), torch.ones(col.shape), T, H).numpy().astype(bool)
    theirs = b['partial_attention_mask'].numpy().astype(bool)
    vt = rows4.expand(N, H, T, T).numpy()
    agree = float(((mine == theirs) | ~vt).mean())
    assert agree >= 0.999, agree
    for n in range(N):                                                           # no column of a padded token
        nnz = int(crow[n, -1])
        assert nnz > 0 and int((col[n, :nnz] % T).max()) < lengths[n]
    same = torch.from_numpy(((mine == theirs) | ~vt).all(axis=(1, 3))) & valid
    assert float(same.float().sum() / valid.float().sum()) > 0.9
    ref_ctx = torch.from_numpy(g['dense.context_layer'])
    # rows whose interpolated mask equals the reference dense path's: same context
    dense_alive = np.unpackbits(g['dense.partial_attention_mask_alive'])[:N * H * T * T].reshape(N, H, T, T).astype(bool)
    same_ref = torch.from_numpy(((mine == dense_alive) | ~vt).all(axis=(1, 3))) & valid
    assert float(same_ref.float().sum() / valid.float().sum()) > 0.9
    torch.testing.assert_close(out.context_layer.cpu()[same_ref], ref_ctx[same_ref], rtol=1e-3, atol=3e-5)
    torch.testing.assert_close(out.context_layer.cpu()[same], b['context_layer'][same], rtol=1e-3, atol=3e-5)
    # bf16 runs the same path; a mask that is not right padding is refused
    mod.output_attentions = False
    qb, kb, vb = q.bfloat16().to(DEV), kk.bfloat16().to(DEV), v.bfloat16().to(DEV)
    out2 = mod(qb, kb, vb, qb, kb, vb, qb, kb, _padded_mask(lengths, T, torch.bfloat16).to(DEV), None, None)
    d2 = (out2.context_layer.float().cpu() - ref_ctx).abs() * rows3
    assert float((d2 > 5e-2 + 5e-2 * ref_ctx.abs()).float().mean()) < 0.03
    bad = _padded_mask(lengths, T)
    bad[1, :, :, 3] = bad[1, :, :, -1]
    with pytest.raises(sea.SeaError):
        mod(qb, kb, vb, qb, kb, vb, qb, kb, bad.to(DEV).bfloat16(), None, None)
