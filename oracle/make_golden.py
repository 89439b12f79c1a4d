"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference; see oracle/ref_harness.py) in the build container.

    python oracle/make_golden.py            # writes tests/golden/*.npz (minutes; Triton interpreter)

The reference's dense torch path runs natively on CPU; its Triton kernels run through Triton's numpy
interpreter (TRITON_INTERPRET=1) with the shims documented in ref_harness.py.  Nothing of the
reference is copied: inputs that only a nested reference function can produce (the Perlin-noise
scores of causal_resize_m_to_t.py:1023-1044) are obtained by compiling that function's AST from the
reference file at run time.
"""
import ast
import os
import sys
import time

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from oracle import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(_HERE), 'tests', 'golden')

# Keys of the reference state_dict the hot path never touches (attention.py:185-189, 315-318): dropped
# from the fixtures to keep them small.
_UNUSED = ('attention_predictor_enc_per_layer.', 'norm_performer.', 'norm_partial.', 'norm_random.', 'norm.',
           'performer_proj_updater.', 'attention_predictor_enc_head_embd')


def _np(t):
    return t.detach().cpu().numpy()


def _extract_nested_function(path, outer, inner):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == outer:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == inner:
                    mod = ast.Module(body=[sub], type_ignores=[])
                    ns = {'torch': torch, 'math': __import__('math')}
                    exec(compile(mod, path, 'exec'), ns)
                    return ns[inner]
    raise KeyError(inner)


def golden_kat_causal_resize():
    """Default config of causal_resize_m_to_t.py:1136-1165 (causal N1 H1 T64 T_M8 K16, seed 42); the
    expected per-row nnz is printed in src/poc/neko/visualize_ops_causal_resize.ipynb:29-35 / :51-57."""
    rh.load_reference()
    import torch.nn.functional as F
    from src.utils import seed
    from src.models.perlin_attention.ops import resize_from_m_to_t_csr, resize_from_m_to_t, flat_csr_to_dense
    from src.models.perlin_attention.ops.kernels.causal_topk_masking import causal_topk_masking
    path = os.path.join(rh.REFERENCE_ROOT, 'src/models/perlin_attention/ops/kernels/causal_resize_m_to_t.py')
    rand_perlin_2d = _extract_nested_function(path, 'test_config', 'rand_perlin_2d')
    N, H, T, T_M, K = 1, 1, 64, 8, 16
    seed()
    FP_MIN = torch.finfo(torch.float16).min * 0.5
    scores = F.interpolate(rand_perlin_2d((128, 128), (16, 16)).view(1, 1, 128, 128), (T, T_M)).expand(N, H, T, T_M).contiguous()
    probs = torch.softmax(scores, dim=-1)
    cm = ((torch.arange(T).view(1, T) > torch.arange(T).view(T, 1)) * FP_MIN).view(1, 1, T, T)
    cmask = causal_topk_masking(probs, k=K * 1.0, attention_mask=cm[:, :, -1:, :], dst_attention_mask=cm[:, :, :, :1],
                                causal_attention_mask=cm, is_causal=True)
    csr = resize_from_m_to_t_csr(cmask, 0, K, target_width=T, is_causal=True, oversampled=1.0)
    dense_from_csr = flat_csr_to_dense(csr, T, H)
    dense = resize_from_m_to_t(cmask, 0, attention_mask=cm, target_width=T, is_causal=True, k=K, oversampled=1.0)
    nnz_notebook = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 14, 15, 15, 16, 15, 15, 14, 15, 15, 12, 14, 13,
                    14, 15, 15, 16, 12, 12, 14, 15, 13, 15, 15, 15, 15, 16, 11, 11, 11, 12, 12, 12, 12, 13, 14, 12, 13, 13,
                    14, 14, 14, 14, 15, 16, 15, 16, 15, 16]
    got = dense_from_csr[0, 0].sum(-1).long().tolist()
    assert got == nnz_notebook, 'reference run does not reproduce its own notebook table'
    assert dense[0, 0].sum(-1).long().tolist() == nnz_notebook
    np.savez_compressed(os.path.join(GOLDEN, 'kat_causal_resize.npz'),
                        probs=_np(probs), compressed_mask=_np(cmask), crow=_np(csr.crow_indices()),
                        col=_np(csr.col_indices()), dense=_np(dense).astype(np.uint8),
                        nnz_per_row_notebook=np.array(nnz_notebook, dtype=np.int64), meta=np.array([N, H, T, T_M, K]))
    print('kat_causal_resize ok')


def golden_kat_causal_conv():
    """src/poc/neko/test_causal_conv.ipynb:65 with the printed table :42-47."""
    rh.load_reference()
    from src.models.perlin_attention.modules import CausalConv2d
    x = torch.eye(8).view(1, 1, 8, 8)
    c = CausalConv2d(1, 3, 3, padding=1, stride=2, causal=True)
    c.bias.data.fill_(0)
    c.weight.data.fill_(1)
    out = c(x)
    table = [[1, 0, 0, 0], [2, 2, 0, 0], [0, 2, 2, 0], [0, 0, 2, 2]]
    assert out.long()[0, 0].tolist() == table
    np.savez_compressed(os.path.join(GOLDEN, 'kat_causal_conv.npz'), x=_np(x), weight=_np(c.weight),
                        weight_mask=_np(c.weight_mask), bias=_np(c.bias), out=_np(out), table=np.array(table))
    print('kat_causal_conv ok')


def _run_layer(m, q, k, v, mask, benchmarking):
    from src.utils import get_bench
    m.benchmarking = benchmarking
    get_bench().activate_temp_buffers = True
    get_bench().reset_temp_buffers()
    with torch.no_grad():
        out = m(q, k, v, q, k, v, q, k, mask, None, None)
    bufs = {k_: v_[-1] for k_, v_ in get_bench().buffers.items()}
    get_bench().activate_temp_buffers = False
    return out, bufs


def golden_layer(name, H, d, T, k, P, nbf, causal, k_flatten_dim=None, seed_inputs=1234):
    """One PerlinAttention layer, dense path (benchmarking=False) and sparse path (benchmarking=True,
    Triton interpreted), fp32, random-init, no padding (test_perlin_opt_causality.py:110-173 style)."""
    t0 = time.time()
    m = rh.build_reference_attention(H, d, T, k, P, nbf, causal, k_flatten_dim=k_flatten_dim)
    g = torch.Generator().manual_seed(seed_inputs)
    q = torch.randn(1, H, T, d, generator=g) * (d ** -0.5 if causal else 1.0)   # OPT pre-scales q (perlin_opt.py:562)
    kk = torch.randn(1, H, T, d, generator=g)
    v = torch.randn(1, H, T, d, generator=g)
    if causal:
        mask = rh.causal_additive_mask(T, torch.float32)
    else:
        mask = torch.zeros(1, 1, 1, T)
    out_d, buf_d = _run_layer(m, q, kk, v, mask, False)
    out_s, buf_s = _run_layer(m, q, kk, v, mask, True)
    csr = out_s.partial_attention_mask
    sd = {k_: _np(v_) for k_, v_ in m.state_dict().items() if not k_.startswith(_UNUSED)}
    keep = ['performer_context_layer', 't_attention_predictor', 'estimated_attention_score_dec_row',
            'estimated_attention_score', 'estimated_attention_probs', 'per_item_top_k', 'estimated_scales',
            'average_context_layer', 'partial_context_layer_1', 'partial_context_layer']
    fx = {'sd.' + k_: v_ for k_, v_ in sd.items()}
    fx.update(q=_np(q), k=_np(kk), v=_np(v))
    for b in keep:
        if b in buf_d and buf_d[b] is not None:
            fx['dense.' + b] = _np(buf_d[b]).astype(np.float32)
    fx['dense.mask_before_interp_alive'] = np.packbits((_np(buf_d['partial_attention_mask_before_interp']) > -1))
    fx['dense.partial_attention_mask_alive'] = np.packbits((_np(buf_d['partial_attention_mask']) > -1))
    fx['sparse.mask_before_interp'] = np.packbits(_np(buf_s['partial_attention_mask_before_interp']) > 0.5)
    fx['sparse.crow'] = _np(csr.crow_indices()).astype(np.int64)
    fx['sparse.col'] = _np(csr.col_indices()).astype(np.int32)
    fx['sparse.probs_values'] = _np(out_s.partial_attention_probs.values()).astype(np.float32)
    fx['sparse.context_layer'] = _np(out_s.context_layer).astype(np.float32)
    fx['sparse.estimated_attention_probs'] = _np(out_s.estimated_attention_probs).astype(np.float32)
    fx['meta'] = np.array([1, H, d, T, k, P, nbf, int(causal)])
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)
    err = (out_d.context_layer - out_s.context_layer).abs().max().item()
    print(f'{name}: dense-vs-sparse context max err {err:.3e}  nnz {int(csr.crow_indices()[0, -1])}  ({time.time() - t0:.0f}s)')


def golden_layer_padded(name, H, d, T, k, P, nbf, lengths, seed_inputs=4321):
    """Causal layer with PADDED query rows (attention.py:401-449: a row t is padded iff causal_mask[n,0,t,0] <= -1; :512-514 zeroes
    v / v_for_atten there, :928-931 kills its top-k): batch of len(lengths) items, item n has lengths[n] real rows, the rest of its
    mask rows are fully masked.  Dense path only (benchmarking=False)."""
    N = len(lengths)
    m = rh.build_reference_attention(H, d, T, k, P, nbf, True)
    g = torch.Generator().manual_seed(seed_inputs)
    q = torch.randn(N, H, T, d, generator=g) * d ** -0.5
    kk = torch.randn(N, H, T, d, generator=g)
    v = torch.randn(N, H, T, d, generator=g)
    mask = rh.causal_additive_mask(T, torch.float32).expand(N, 1, T, T).clone()
    fmin = float(mask.min())
    for n, L in enumerate(lengths):
        mask[n, :, L:, :] = fmin
    out_d, buf_d = _run_layer(m, q.clone(), kk.clone(), v.clone(), mask, False)       # (the reference masks v in place)
    sd = {k_: _np(v_) for k_, v_ in m.state_dict().items() if not k_.startswith(_UNUSED)}
    fx = {'sd.' + k_: v_ for k_, v_ in sd.items()}
    fx.update(q=_np(q), k=_np(kk), v=_np(v), lengths=np.array(lengths))
    for b in ('performer_context_layer', 'estimated_attention_probs', 'estimated_scales', 'average_context_layer', 'partial_context_layer'):
        fx['dense.' + b] = _np(buf_d[b]).astype(np.float32)
    fx['dense.mask_before_interp_alive'] = np.packbits((_np(buf_d['partial_attention_mask_before_interp']) > -1))
    fx['dense.context_layer'] = _np(out_d.context_layer).astype(np.float32)
    fx['meta'] = np.array([N, H, d, T, k, P, nbf, 1])
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)
    print(name, 'written; context abs mean', float(out_d.context_layer.abs().mean()))


def golden_layer_query_skips(name, H, d, T, k, P, nbf, skips, seed_inputs=777):
    """QUERY_SKIPS > 1 (attention.py:598, 617-619, 640-644): the predictor MLP + CNN run on every `skips`-th query row and every
    result is repeated `skips` times.  Dense path, causal, no padding."""
    m = rh.build_reference_attention(H, d, T, k, P, nbf, True)
    g = torch.Generator().manual_seed(seed_inputs)
    q = torch.randn(1, H, T, d, generator=g) * d ** -0.5
    kk = torch.randn(1, H, T, d, generator=g)
    v = torch.randn(1, H, T, d, generator=g)
    mask = rh.causal_additive_mask(T, torch.float32)
    os.environ['QUERY_SKIPS'] = str(skips)
    try:
        out_d, buf_d = _run_layer(m, q, kk, v, mask, False)
    finally:
        del os.environ['QUERY_SKIPS']
    sd = {k_: _np(v_) for k_, v_ in m.state_dict().items() if not k_.startswith(_UNUSED)}
    fx = {'sd.' + k_: v_ for k_, v_ in sd.items()}
    fx.update(q=_np(q), k=_np(kk), v=_np(v), skips=np.array(skips))
    for b in ('estimated_attention_probs', 'estimated_scales'):
        fx['dense.' + b] = _np(buf_d[b]).astype(np.float32)
    fx['dense.mask_before_interp_alive'] = np.packbits((_np(buf_d['partial_attention_mask_before_interp']) > -1))
    fx['dense.context_layer'] = _np(out_d.context_layer).astype(np.float32)
    fx['meta'] = np.array([1, H, d, T, k, P, nbf, 1])
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)
    print(name, 'written')


def golden_layer_deeper(name, H, d, T, k, P, nbf, seed_inputs=99):
    """PERLIN_HOTFIX_OPT_DEEPER=1 (attention.py:246-263): the predictor CNN with a third dilated causal conv.  The switch is read
    by the reference's constructor.  Dense path, causal, no padding."""
    os.environ['PERLIN_HOTFIX_OPT_DEEPER'] = '1'
    try:
        m = rh.build_reference_attention(H, d, T, k, P, nbf, True)
    finally:
        del os.environ['PERLIN_HOTFIX_OPT_DEEPER']
    g = torch.Generator().manual_seed(seed_inputs)
    q = torch.randn(1, H, T, d, generator=g) * d ** -0.5
    kk = torch.randn(1, H, T, d, generator=g)
    v = torch.randn(1, H, T, d, generator=g)
    mask = rh.causal_additive_mask(T, torch.float32)
    out_d, buf_d = _run_layer(m, q, kk, v, mask, False)
    sd = {k_: _np(v_) for k_, v_ in m.state_dict().items() if not k_.startswith(_UNUSED)}
    assert 'attention_predictor_cnn.1.module.net.7.module.weight' in sd
    fx = {'sd.' + k_: v_ for k_, v_ in sd.items()}
    fx.update(q=_np(q), k=_np(kk), v=_np(v))
    for b in ('estimated_attention_score', 'estimated_attention_probs', 'estimated_scales'):
        fx['dense.' + b] = _np(buf_d[b]).astype(np.float32)
    fx['dense.mask_before_interp_alive'] = np.packbits((_np(buf_d['partial_attention_mask_before_interp']) > -1))
    fx['dense.context_layer'] = _np(out_d.context_layer).astype(np.float32)
    fx['meta'] = np.array([1, H, d, T, k, P, nbf, 1])
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)
    print(name, 'written')


def golden_layer_bert_padded(name, H, d, T, k, P, nbf, lengths, seed_inputs=2468):
    """Non-causal (BERT) layer with a RIGHT-PADDED batch: attention_mask [N,1,1,T] is fully negative on the padded tokens of each item
    (attention.py:401-449, 482, 512-514, 777-778, 837, 1209-1219; dense interpolation width = token length, resize_m_to_t.py:36-47).
    Dense path only (benchmarking=False): the reference's sparse non-causal path ignores padding in the interpolation
    (causal_resize_m_to_t.py:955-957, "TODO confirm correctness")."""
    N = len(lengths)
    m = rh.build_reference_attention(H, d, T, k, P, nbf, False, k_flatten_dim='batch')
    g = torch.Generator().manual_seed(seed_inputs)
    q = torch.randn(N, H, T, d, generator=g)
    kk = torch.randn(N, H, T, d, generator=g)
    v = torch.randn(N, H, T, d, generator=g)
    fmin = float(torch.finfo(torch.float32).min / 2)
    mask = torch.zeros(N, 1, 1, T)
    for n, L in enumerate(lengths):
        mask[n, :, :, L:] = fmin
    out_d, buf_d = _run_layer(m, q.clone(), kk.clone(), v.clone(), mask, False)       # (the reference masks v in place)
    sd = {k_: _np(v_) for k_, v_ in m.state_dict().items() if not k_.startswith(_UNUSED)}
    fx = {'sd.' + k_: v_ for k_, v_ in sd.items()}
    fx.update(q=_np(q), k=_np(kk), v=_np(v), lengths=np.array(lengths))
    for b in ('performer_context_layer', 'estimated_attention_score', 'estimated_attention_probs', 'estimated_scales', 'average_context_layer',
              'partial_context_layer'):
        if b in buf_d and buf_d[b] is not None:
            fx['dense.' + b] = _np(buf_d[b]).astype(np.float32)
    fx['dense.mask_before_interp_alive'] = np.packbits((_np(buf_d['partial_attention_mask_before_interp']) > -1))
    fx['dense.partial_attention_mask_alive'] = np.packbits((_np(buf_d['partial_attention_mask']) > -1))
    fx['dense.context_layer'] = _np(out_d.context_layer).astype(np.float32)
    fx['meta'] = np.array([N, H, d, T, k, P, nbf, 0])
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)
    print(name, 'written; buffers', sorted(k_ for k_ in fx if k_.startswith('dense.')))


def golden_layer_training(name, H, d, T, k, P, nbf, N=2, seed_inputs=1357):
    """TRAINING branch of the causal layer (attention.py:680-765, 1066-1133, 1328-1332): module in train() mode, benchmarking=False, teacher
    tensors given -> the distillation loss: KL + MSE of the resized estimated scores, KL + MSE of the dense scores, MSE of the context.
    The only randomness of the branch is the 10 % index jitter of resize_from_m_to_t(training=True) (resize_m_to_t.py:40-45, python
    `random.random()`): it is switched off for the fixture by pinning random.random to 1.0, so that the run is deterministic.
    FORWARD ONLY: in fp32 without autocast the reference's own backward cannot run -- `partial_attention_probs.masked_fill_` (:1120) writes
    into the softmax output autograd saved (softmax_bf16 only makes a copy when it casts, i.e. under half-precision autocast on a GPU);
    gradients are therefore checked against autograd of the oracle's restatement of this forward (tests/test_training_*.py)."""
    import random
    m = rh.build_reference_attention(H, d, T, k, P, nbf, True)
    m.train()
    m.benchmarking = False
    g = torch.Generator().manual_seed(seed_inputs)
    q = torch.randn(N, H, T, d, generator=g) * d ** -0.5
    kk = torch.randn(N, H, T, d, generator=g)
    v = torch.randn(N, H, T, d, generator=g)
    scores_truth = torch.randn(N, H, T, T, generator=g) * 2.0
    context_truth = torch.randn(N, T, H * d, generator=g) * 0.3
    mask = rh.causal_additive_mask(T, torch.float32).expand(N, 1, T, T).clone()
    from src.utils import get_bench
    real_random = random.random
    random.random = lambda: 1.0
    try:
        with torch.no_grad():
            out_nt = m(q, kk, v, q, kk, v, q, kk, mask, None, None)                       # no teacher: loss == 0
            get_bench().activate_temp_buffers = True
            get_bench().reset_temp_buffers()
            out = m(q, kk, v, q, kk, v, q, kk, mask, scores_truth.clone(), context_truth.clone())
            bufs = {k_: v_[-1] for k_, v_ in get_bench().buffers.items()}
            get_bench().activate_temp_buffers = False
    finally:
        random.random = real_random
    assert float(out_nt.loss) == 0.0
    sd = {k_: _np(v_) for k_, v_ in m.state_dict().items() if not k_.startswith(_UNUSED)}
    fx = {'sd.' + k_: v_ for k_, v_ in sd.items()}
    fx.update(q=_np(q), k=_np(kk), v=_np(v), scores_truth=_np(scores_truth), context_truth=_np(context_truth))
    fx['loss'] = np.array(float(out.loss))
    fx['context_layer'] = _np(out.context_layer).astype(np.float32)
    fx['estimated_attention_probs_m'] = _np(out.estimated_attention_probs_m).astype(np.float32)
    fx['estimated_attention_probs'] = _np(out.estimated_attention_probs).astype(np.float32)          # resized to [N,H,T,T] (:1338)
    fx['partial_attention_probs'] = _np(out.partial_attention_probs).astype(np.float32)
    fx['dense_attention_probs'] = _np(out.dense_attention_probs).astype(np.float32)
    fx['partial_attention_mask_alive'] = np.packbits(_np(out.partial_attention_mask) > -1)
    # the reference's own top-k selection (its CPU sort breaks the exact ties of the x4-upsampled predictor arbitrarily): lets a checker
    # evaluate the loss on the SAME mask
    fx['mask_before_interp_alive'] = np.packbits(_np(bufs['partial_attention_mask_before_interp']) > -1)
    fx['meta'] = np.array([N, H, d, T, k, P, nbf, 1])
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **fx)
    print(name, 'written; loss', float(out.loss))


def golden_state_ops():
    """The reference's three stateful decode ops (attention_state.py:43-98 StatefulCausalPerformer, :142-187 StatefulCausalCNN,
    :205-224 StatefulCumAvg) driven exactly as PerlinAttention drives them during a token-by-token decode, on seeded inputs."""
    rh.load_reference()
    from src.models.perlin_attention import attention_state as ast_
    from src.models.perlin_attention.modules import CausalConv2d
    g = torch.Generator().manual_seed(99)
    N, H, T, F_, E = 2, 3, 37, 9, 10
    chunks = [5, 1, 1, 3, 1, 16, 1, 9]                                     # prompt of 5 tokens, then single tokens / small chunks
    assert sum(chunks) == T
    parent = ast_.PerlinAttentionState(None)
    parent.num_heads, parent.head_dim, parent.embd_dim = H, E, H * E
    # --- performer recurrence (operates on whatever q / k it is given; fp64 running sums, eps 1e-12)
    qf = torch.rand(N, H, T, F_, generator=g) + 1e-3
    kf = torch.rand(N, H, T, F_, generator=g) + 1e-3
    v = torch.randn(N, H, T, E, generator=g)
    perf = ast_.StatefulCausalPerformer(parent, None)
    outs, t = [], 0
    for c in chunks:
        outs.append(perf(qf[:, :, t:t + c], kf[:, :, :t + c], v[:, :, :t + c]))
        t += c
    perf_out = torch.cat(outs, dim=-2)
    # --- running mean
    ca = ast_.StatefulCumAvg(parent)
    outs, t = [], 0
    for c in chunks:
        outs.append(ca(v[:, :, :t + c], c))
        t += c
    cumavg_out = torch.cat(outs, dim=-2)
    # --- windowed causal CNN: two dilated causal 3x3 convs + ReLU, the receptive field of the predictor CNN (8 rows < window 24)
    C, Wd = 4, 6
    torch.manual_seed(7)
    cnn = torch.nn.Sequential(CausalConv2d(C, C, 3, padding=2, dilation=2, causal=True), torch.nn.ReLU(),
                              CausalConv2d(C, C, 3, padding=2, dilation=2, causal=True), torch.nn.ReLU())
    x = torch.randn(N, C, T, Wd, generator=g)
    st = ast_.StatefulCausalCNN(parent)
    outs, t = [], 0
    with torch.no_grad():
        for c in chunks:
            outs.append(st(cnn, x[:, :, t:t + c], c))
            t += c
        cnn_out = torch.cat(outs, dim=-2)
        cnn_full = cnn(x)
    sd = {k_: _np(v_) for k_, v_ in cnn.state_dict().items()}
    np.savez_compressed(os.path.join(GOLDEN, 'state_ops.npz'), chunks=np.array(chunks), qf=_np(qf), kf=_np(kf), v=_np(v), performer_out=_np(perf_out),
                        cumavg_out=_np(cumavg_out), cnn_x=_np(x), cnn_out=_np(cnn_out), cnn_full=_np(cnn_full),
                        **{'cnn.' + k_: v_ for k_, v_ in sd.items()})
    print('state_ops: windowed CNN vs full CNN max err', float((cnn_out - cnn_full).abs().max()))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if '--state-only' in sys.argv:
        return golden_state_ops()
    if '--skips-only' in sys.argv:
        return golden_layer_query_skips('layer_causal_skips2_h3_t64', H=3, d=32, T=64, k=6, P=16, nbf=4, skips=2)
    if '--training-only' in sys.argv:
        return golden_layer_training('layer_causal_training_h3_t48', H=3, d=32, T=48, k=6, P=16, nbf=4)
    if '--bert-padded-only' in sys.argv:
        return golden_layer_bert_padded('layer_bert_padded_h4_t64', H=4, d=64, T=64, k=8, P=32, nbf=1, lengths=[64, 45, 23])
    if '--deeper-only' in sys.argv:
        return golden_layer_deeper('layer_causal_deeper_h3_t64', H=3, d=32, T=64, k=6, P=16, nbf=4)
    if '--padded-only' in sys.argv:
        return golden_layer_padded('layer_causal_padded_h3_t64', H=3, d=32, T=64, k=6, P=16, nbf=4, lengths=[64, 45])
    golden_kat_causal_resize()
    golden_kat_causal_conv()
    golden_layer('layer_causal_h4_t128', H=4, d=64, T=128, k=8, P=32, nbf=8, causal=True)
    golden_layer('layer_causal_h3_t100', H=3, d=32, T=100, k=6, P=16, nbf=4, causal=True)
    golden_state_ops()
    golden_layer_padded('layer_causal_padded_h3_t64', H=3, d=32, T=64, k=6, P=16, nbf=4, lengths=[64, 45])
    golden_layer_query_skips('layer_causal_skips2_h3_t64', H=3, d=32, T=64, k=6, P=16, nbf=4, skips=2)
    golden_layer_deeper('layer_causal_deeper_h3_t64', H=3, d=32, T=64, k=6, P=16, nbf=4)
    golden_layer_training('layer_causal_training_h3_t48', H=3, d=32, T=48, k=6, P=16, nbf=4)
    if '--with-bert' in sys.argv:
        golden_layer_bert_padded('layer_bert_padded_h4_t64', H=4, d=64, T=64, k=8, P=32, nbf=1, lengths=[64, 45, 23])
        golden_layer('layer_bert_h4_t64', H=4, d=64, T=64, k=8, P=32, nbf=1, causal=False, k_flatten_dim='batch')


if __name__ == '__main__':
    main()
