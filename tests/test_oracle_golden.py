"""CPU: pins the oracle (oracle/sea_oracle.py) against
  * the fixtures produced by running the unmodified reference (oracle/make_golden.py), and
  * the reference's own known-answer vectors (notebook tables)."""
import numpy as np
import pytest
import torch

import os

from conftest import ROOT, golden_layer, load_golden
from oracle import sea_oracle as so

LAYERS = ['layer_causal_h4_t128', 'layer_causal_h3_t100']


def _bits(g, key, shape):
    n = int(np.prod(shape))
    return np.unpackbits(g[key])[:n].reshape(shape)


def test_kat_notebook_nnz_per_row():
    # src/poc/neko/visualize_ops_causal_resize.ipynb:29-35 (CSR) and :51-57 (dense)
    g = load_golden('kat_causal_resize')
    N, H, T, P, K = g['meta'].tolist()
    # top-k stage: this input has rows of exactly equal probabilities, where the reference's torch.topk is
    # implementation-defined; the oracle (lower index wins) must keep the same MULTISET of keys per row.
    mine = so.topk_mask_causal_batch(torch.from_numpy(g['probs']), float(K), floor_variant=True).numpy().astype(bool)
    ref = g['compressed_mask'].astype(bool)
    pr = g['probs']
    for t in range(T):
        assert np.array_equal(np.sort(pr[0, 0, t][mine[0, 0, t]]), np.sort(pr[0, 0, t][ref[0, 0, t]]))
    # CSR stage on the reference's own compressed mask: bit-exact, and equal to the notebook table
    mask = torch.from_numpy(g['compressed_mask'])
    crow, col, Z = so.resize_from_m_to_t_csr(mask, K, T, True)
    assert np.diff(crow.numpy()[0]).tolist() == g['nnz_per_row_notebook'].tolist()
    assert abs(np.diff(crow.numpy()[0])[-32:].mean() - 13.6562) < 1e-3
    assert np.array_equal(crow.numpy(), g['crow']) and np.array_equal(col.numpy(), g['col'])
    dense = so.flat_csr_to_dense(crow, col, torch.ones(col.shape), T, H)
    assert np.array_equal(dense.numpy().astype(np.uint8), g['dense'])


def test_kat_notebook_causal_conv():
    # src/poc/neko/test_causal_conv.ipynb:65, table :42-47
    g = load_golden('kat_causal_conv')
    y = so.causal_conv2d(*(torch.from_numpy(g[k]) for k in ('x', 'weight', 'weight_mask', 'bias')), 3, 1, 1, stride=2)
    assert np.array_equal(y.numpy(), g['out'])
    assert y.long()[0, 0].tolist() == g['table'].tolist()


@pytest.mark.parametrize('name', LAYERS)
def test_dense_path_stages_match_reference(name):
    g, m, sd = golden_layer(name)
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    b = so.perlin_forward_causal(sd, q, k, v, k_top=m['k'], P=m['P'], sparse=False, keep_dense=True)
    for key in ['performer_context_layer', 't_attention_predictor', 'estimated_attention_score', 'estimated_attention_probs',
                'estimated_scales', 'average_context_layer']:
        torch.testing.assert_close(b[key], torch.from_numpy(g['dense.' + key]), rtol=1e-3, atol=2e-5, msg=key)
    assert np.array_equal(so.per_item_top_k_causal(m['H'], m['k'], 1.0, m['P'], m['T']), g['dense.per_item_top_k'].reshape(-1))
    # top-k on the reference's own probabilities: alive sets equal up to exact ties (the reference's sort is
    # unstable, attention.py:880-885); the oracle's contract is "lower flat index wins".
    H, T, P = m['H'], m['T'], m['P']
    probs = torch.from_numpy(g['dense.estimated_attention_probs'])
    mine = so.topk_mask_causal_batch(probs, m['k'], 1.0).numpy().astype(bool)
    ref = _bits(g, 'dense.mask_before_interp_alive', (1, H, T, P)).astype(bool)
    pr = probs.transpose(1, 2).reshape(T, H * P).numpy()
    a1 = mine.transpose(0, 2, 1, 3).reshape(T, H * P)
    a2 = ref.transpose(0, 2, 1, 3).reshape(T, H * P)
    for t in range(T):
        assert np.array_equal(np.sort(pr[t][a1[t]]), np.sort(pr[t][a2[t]])), f'row {t}: alive key multisets differ'


@pytest.mark.parametrize('name', LAYERS)
def test_csr_and_sparse_attention_match_reference_triton(name):
    """Given the reference's compressed mask, the oracle reproduces the (interpreted) Triton kernels:
    crow/col bit-exact, probabilities and context within fp32 tolerance."""
    g, m, sd = golden_layer(name)
    H, T, P, d = m['H'], m['T'], m['P'], m['d']
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    mask_m = torch.from_numpy(_bits(g, 'sparse.mask_before_interp', (1, H, T, P)).astype(np.float32))
    crow, col, Z = so.resize_from_m_to_t_csr(mask_m, m['k'], T, True)
    assert np.array_equal(crow.numpy(), g['sparse.crow'])
    assert np.array_equal(col.numpy(), g['sparse.col'].astype(np.int64))
    s = so.flat_csr_masked_bmm(q, k, crow, col)
    p = so.flat_csr_softmax(s, crow, col, H, T)
    scales = torch.from_numpy(g['dense.estimated_scales'])
    p = so.flat_csr_elmul_rowscale(p, crow, col, torch.sigmoid(scales[..., 0]), T)
    torch.testing.assert_close(p, torch.from_numpy(g['sparse.probs_values']), rtol=1e-3, atol=1e-6)
    ctx = so.flat_csr_sdbmm(p, crow, col, v, H)
    avg = torch.from_numpy(g['dense.average_context_layer'])
    a = torch.sigmoid(scales[..., 1:2])
    out = (ctx * a + (1 - a) * avg).permute(0, 2, 1, 3).reshape(1, T, H * d)
    torch.testing.assert_close(out, torch.from_numpy(g['sparse.context_layer']), rtol=1e-3, atol=2e-5)


def test_dense_resize_matches_reference():
    g, m, sd = golden_layer('layer_causal_h4_t128')
    H, T, P = m['H'], m['T'], m['P']
    alive = torch.from_numpy(_bits(g, 'dense.mask_before_interp_alive', (1, H, T, P)).astype(np.float32))
    fmin = so.fp_min_for(torch.float32)
    cm = so.causal_additive_mask(T)
    pm = so.resize_from_m_to_t_dense((1 - alive) * fmin, fmin, cm, T, True, m['k'], 1.0).masked_fill(cm < -1, fmin)
    ref = _bits(g, 'dense.partial_attention_mask_alive', (1, H, T, T))
    assert np.array_equal((pm > -1).numpy().astype(np.uint8), ref)


def test_bert_layer_oracle_matches_reference():
    """Non-causal (BERT, k_flatten_dim='batch') oracle forward vs the reference's dense and Triton paths."""
    g, m, sd = golden_layer('layer_bert_h4_t64')
    H, T, P, d = m['H'], m['T'], m['P'], m['d']
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    for sparse in (False, True):
        b = so.perlin_forward_noncausal(sd, q, k, v, k_top=m['k'], P=P, sparse=sparse, keep_dense=True)
        for key in ['performer_context_layer', 't_attention_predictor', 'estimated_attention_score', 'estimated_attention_probs',
                    'estimated_scales', 'average_context_layer', 'partial_context_layer_1']:
            torch.testing.assert_close(b[key], torch.from_numpy(g['dense.' + key]), rtol=1e-3, atol=2e-5, msg=key)
        alive = _bits(g, 'dense.mask_before_interp_alive', (1, H, T, P))
        assert np.array_equal(alive, b['partial_attention_mask_before_interp'].numpy().astype(np.uint8))
        if sparse:
            assert np.array_equal(g['sparse.crow'], b['crow_indices'].numpy())
            assert np.array_equal(g['sparse.col'].astype(np.int64), b['col_indices'].numpy())
            torch.testing.assert_close(b['context_layer'], torch.from_numpy(g['sparse.context_layer']), rtol=1e-3, atol=2e-5)
        else:
            assert np.array_equal(_bits(g, 'dense.partial_attention_mask_alive', (1, H, T, T)), b['partial_attention_mask'].numpy().astype(np.uint8))
            torch.testing.assert_close(b['context_layer'], torch.from_numpy(g['dense.partial_context_layer']), rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize('name', LAYERS)
def test_backward_oracle_forward_matches_reference_dense_path(name):
    """Pins the expression oracle.sparse_attention_grads differentiates to the reference: on the reference's own
    mask, q, k, v and estimated_scales its forward value equals the reference dense path's partial_context_layer."""
    g, meta, sd = golden_layer(name)
    q, k, v = (torch.from_numpy(g[x]) for x in ('q', 'k', 'v'))
    N, H, T, d = q.shape
    alive = torch.from_numpy(_bits(g, 'dense.partial_attention_mask_alive', (N, H, T, T)).astype(bool))
    scales = torch.from_numpy(g['dense.estimated_scales'])
    dout = torch.ones(N, T, H * d)
    out, dq, dk, dv, ds = so.sparse_attention_grads(alive, q, k, v, scales, dout, use_scaler=True, with_avg=True)
    torch.testing.assert_close(out, torch.from_numpy(g['dense.partial_context_layer']), rtol=1e-4, atol=1e-5)
    # finite-difference spot check of the gradient itself (fp64 autograd vs central differences on one coordinate)
    eps = 1e-3
    qp, qm = q.clone(), q.clone()
    qp[0, 1, 7, 3] += eps; qm[0, 1, 7, 3] -= eps
    fp = so.sparse_attention_grads(alive, qp, k, v, scales, dout)[0].double().sum()
    fm = so.sparse_attention_grads(alive, qm, k, v, scales, dout)[0].double().sum()
    assert abs(float((fp - fm) / (2 * eps)) - float(dq[0, 1, 7, 3])) < 5e-3 * max(1.0, abs(float(dq[0, 1, 7, 3])))


@pytest.mark.parametrize('causal,d,F,T', [(True, 64, 33, 96), (True, 32, 27, 50), (False, 64, 266, 40), (False, 16, 11, 24)])
def test_performer_restatements_pinned_to_in_tree_jax_original(causal, d, F, T):
    """The only Performer specification under /root/reference is the JAX original it vendors
    (_lra_benchmarks/models/performer/performer_attention.py); oracle/jax_performer_numpy.py is its line-by-line fp64 numpy
    port.  Both restatements the checker relies on -- the `performer_pytorch.FastAttention` stand-in that the imported
    reference ran with when the fixtures were generated, and sea_oracle.performer_* -- must equal it (the two published
    implementations differ only in how the normaliser is stabilised, a ~1e-6-relative term)."""
    import sys
    from oracle import jax_performer_numpy as jp
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'third_party_restated'))
    import performer_pytorch as pp
    g = torch.Generator().manual_seed(d + T)
    q = torch.randn(1, 2, T, d, generator=g, dtype=torch.float64) * (d ** -0.5 if causal else 1.0)
    k = torch.randn(1, 2, T, d, generator=g, dtype=torch.float64)
    v = torch.randn(1, 2, T, 2 * d, generator=g, dtype=torch.float64)
    fa = pp.FastAttention(d, F, causal=causal, generalized_attention=causal).double()
    proj = fa.projection_matrix
    want = torch.from_numpy(jp.fast_attention(q.numpy(), k.numpy(), v.numpy(), proj.numpy(), causal))
    torch.testing.assert_close(fa(q, k, v), want, rtol=1e-5, atol=1e-8)
    mine = so.performer_causal(q, k, v, proj) if causal else so.performer_noncausal(q, k, v, proj)
    torch.testing.assert_close(mine.double(), want, rtol=2e-4, atol=1e-6)       # sea_oracle computes the features in fp32


def test_stateful_decode_ops_match_reference_fixture():
    """a17: the oracle's restatement of attention_state.py's three stateful ops against the outputs of the unmodified reference
    classes (tests/golden/state_ops.npz, oracle/make_golden.py::golden_state_ops), driven token by token the same way."""
    import os
    import numpy as np
    import torch
    from oracle import sea_oracle as so
    fx = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'state_ops.npz'))
    chunks = [int(c) for c in fx['chunks']]
    qf, kf, v = (torch.from_numpy(fx[n]) for n in ('qf', 'kf', 'v'))
    perf, ca = so.StatefulCausalPerformerOracle(), so.StatefulCumAvgOracle()
    outs_p, outs_a, t = [], [], 0
    for c in chunks:
        outs_p.append(perf(qf[:, :, t:t + c], kf[:, :, :t + c], v[:, :, :t + c]))
        outs_a.append(ca(v[:, :, :t + c], c))
        t += c
    assert torch.equal(torch.cat(outs_p, dim=-2), torch.from_numpy(fx['performer_out']))
    assert torch.equal(torch.cat(outs_a, dim=-2), torch.from_numpy(fx['cumavg_out']))
    # windowed CNN: two dilated causal convs + ReLU from the fixture's weights, through the oracle's own conv
    w0, m0, b0 = (torch.from_numpy(fx['cnn.0.' + n]) for n in ('weight', 'weight_mask', 'bias'))
    w2, m2, b2 = (torch.from_numpy(fx['cnn.2.' + n]) for n in ('weight', 'weight_mask', 'bias'))

    def cnn(x):
        y = torch.relu(so.causal_conv2d(x, w0, m0, b0, 3, 2, 2))
        return torch.relu(so.causal_conv2d(y, w2, m2, b2, 3, 2, 2))
    x = torch.from_numpy(fx['cnn_x'])
    st = so.StatefulCausalCNNOracle()
    outs, t = [], 0
    for c in chunks:
        outs.append(st(cnn, x[:, :, t:t + c], c))
        t += c
    got = torch.cat(outs, dim=-2)
    torch.testing.assert_close(got, torch.from_numpy(fx['cnn_out']), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got, torch.from_numpy(fx['cnn_full']), rtol=1e-5, atol=1e-6)


def test_padded_causal_layer_matches_reference_fixture():
    """Padded query rows in the causal model (attention.py:401-449, 512-514, 928-931): the oracle's `dst_valid` branch against the
    unmodified reference run on a 2-item batch whose second item has 45 real rows of 64 (make_golden.py::golden_layer_padded)."""
    g, m, sd = golden_layer('layer_causal_padded_h3_t64')
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    N, H, T, P = m['N'], m['H'], m['T'], m['P']
    valid = (torch.arange(T).view(1, T) < torch.from_numpy(g['lengths']).view(N, 1)).float()
    b = so.perlin_forward_causal(sd, q, k, v, k_top=m['k'], P=P, sparse=False, dst_valid=valid)
    for key in ['performer_context_layer', 'estimated_attention_probs', 'estimated_scales', 'average_context_layer']:
        torch.testing.assert_close(b[key], torch.from_numpy(g['dense.' + key]), rtol=1e-3, atol=2e-5, msg=key)
    ref_alive = _bits(g, 'dense.mask_before_interp_alive', (N, H, T, P)).astype(bool)
    mine = b['partial_attention_mask_before_interp'].numpy().astype(bool)
    # padded rows: nothing alive; real rows: the same alive keys up to exact ties
    assert not mine[1, :, 45:].any() and not ref_alive[1, :, 45:].any()
    probs = b['estimated_attention_probs']
    for n in range(N):
        pr = probs[n].transpose(0, 1).reshape(T, H * P).numpy()
        a1 = mine[n].transpose(1, 0, 2).reshape(T, H * P)
        a2 = ref_alive[n].transpose(1, 0, 2).reshape(T, H * P)
        for t in range(T):
            assert np.array_equal(np.sort(pr[t][a1[t]]), np.sort(pr[t][a2[t]])), f'item {n} row {t}: alive key multisets differ'
    same = torch.from_numpy((mine == ref_alive).all(axis=(1, 3)))                               # rows with identical masks [N,T]
    ctx = b['context_layer']
    ref_ctx = torch.from_numpy(g['dense.context_layer'])
    assert same.float().mean() > 0.5          # (exact ties of the x4-upsampled predictor are frequent at P = 16; the tie rule is the only freedom)
    torch.testing.assert_close(ctx[same], ref_ctx[same], rtol=1e-3, atol=2e-5)


def test_query_skips_layer_matches_reference_fixture():
    """QUERY_SKIPS=2 (attention.py:598, 617-619, 640-644) against the unmodified reference's run."""
    g, m, sd = golden_layer('layer_causal_skips2_h3_t64')
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    b = so.perlin_forward_causal(sd, q, k, v, k_top=m['k'], P=m['P'], sparse=False, query_skips=int(g['skips']))
    for key in ['estimated_attention_probs', 'estimated_scales']:
        torch.testing.assert_close(b[key], torch.from_numpy(g['dense.' + key]), rtol=1e-3, atol=2e-5, msg=key)
    ref_alive = _bits(g, 'dense.mask_before_interp_alive', (1, m['H'], m['T'], m['P'])).astype(bool)
    mine = b['partial_attention_mask_before_interp'].numpy().astype(bool)
    same = torch.from_numpy((mine == ref_alive).all(axis=(1, 3)))
    assert same.float().mean() > 0.5
    torch.testing.assert_close(b['context_layer'][same], torch.from_numpy(g['dense.context_layer'])[same], rtol=1e-3, atol=2e-5)


def test_deeper_cnn_layer_matches_reference_fixture():
    """PERLIN_HOTFIX_OPT_DEEPER=1 (attention.py:246-263: a third dilated causal conv) against the unmodified reference's run."""
    g, m, sd = golden_layer('layer_causal_deeper_h3_t64')
    assert 'attention_predictor_cnn.1.module.net.7.module.weight' in sd
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    b = so.perlin_forward_causal(sd, q, k, v, k_top=m['k'], P=m['P'], sparse=False)
    for key in ['estimated_attention_score', 'estimated_attention_probs', 'estimated_scales']:
        torch.testing.assert_close(b[key], torch.from_numpy(g['dense.' + key]), rtol=1e-3, atol=2e-5, msg=key)
    ref_alive = _bits(g, 'dense.mask_before_interp_alive', (1, m['H'], m['T'], m['P'])).astype(bool)
    mine = b['partial_attention_mask_before_interp'].numpy().astype(bool)
    same = torch.from_numpy((mine == ref_alive).all(axis=(1, 3)))
    assert same.float().mean() > 0.5
    torch.testing.assert_close(b['context_layer'][same], torch.from_numpy(g['dense.context_layer'])[same], rtol=1e-3, atol=2e-5)


def test_bert_padded_batch_oracle_matches_reference():
    """Right-padded non-causal batch (lengths 64 / 45 / 23 of T = 64) against the unmodified reference's dense path: every float buffer,
    the top-k mask, the interpolated mask and the context -- on the VALID query rows (the reference leaves the rows of padded
    queries to the caller's masking).  The oracle's sparse branch (CSR interpolation with the token length as width) is then held
    to its own dense branch on those rows."""
    g, m, sd = golden_layer('layer_bert_padded_h4_t64')
    N, H, T, P, d = m['N'], m['H'], m['T'], m['P'], m['d']
    lengths = torch.from_numpy(g['lengths']).long()
    q, k, v = (torch.from_numpy(g[x]) for x in 'qkv')
    valid = torch.arange(T).view(1, T) < lengths.view(N, 1)                                   # [N,T]
    rows4 = valid.view(N, 1, T, 1)
    bd = so.perlin_forward_noncausal(sd, q, k, v, k_top=m['k'], P=P, sparse=False, keep_dense=True, lengths=lengths)
    for key in ['performer_context_layer', 'estimated_attention_score', 'estimated_scales']:
        ref = torch.from_numpy(g['dense.' + key])
        torch.testing.assert_close(bd[key] * rows4, ref * rows4, rtol=1e-3, atol=2e-5, msg=key)
    torch.testing.assert_close(bd['average_context_layer'], torch.from_numpy(g['dense.average_context_layer']), rtol=1e-3, atol=2e-5)
    alive = _bits(g, 'dense.mask_before_interp_alive', (N, H, T, P)).astype(bool)
    mine = bd['partial_attention_mask_before_interp'].numpy().astype(bool)
    v4 = rows4.expand(N, H, T, P).numpy()
    assert np.array_equal(alive & v4, mine & v4)
    dense_alive = _bits(g, 'dense.partial_attention_mask_alive', (N, H, T, T)).astype(bool)
    vt = rows4.expand(N, H, T, T).numpy()
    assert np.array_equal(dense_alive & vt, bd['partial_attention_mask'].numpy().astype(bool) & vt)
    ref_ctx = torch.from_numpy(g['dense.context_layer'])
    rows3 = valid.view(N, T, 1)
    torch.testing.assert_close(bd['context_layer'] * rows3, ref_ctx * rows3, rtol=1e-3, atol=2e-5)
    # sparse branch: same rows, same context (the two interpolation rules agree up to exact .5 pixel edges)
    bs = so.perlin_forward_noncausal(sd, q, k, v, k_top=m['k'], P=P, sparse=True, keep_dense=True, lengths=lengths)
    same = torch.from_numpy((bs['partial_attention_mask'].numpy().astype(bool) == bd['partial_attention_mask'].numpy().astype(bool)).all(axis=(1, 3))) & valid
    assert float(same.float().sum() / valid.float().sum()) > 0.9
    torch.testing.assert_close(bs['context_layer'][same], ref_ctx[same], rtol=1e-3, atol=2e-5)
    # no column of a padded token is ever selected
    for n in range(N):
        nnz = int(bs['crow_indices'][n, -1])
        assert int((bs['col_indices'][n, :nnz] % T).max()) < int(lengths[n])
