"""Whole-layer GPU parity at the BASELINE.json configs themselves (not toy shapes), through the drop-in module and the
C ABI, against the CPU oracle (oracle/sea_oracle.py, pinned to the unmodified reference by tests/test_oracle_golden.py)
on the same seeded inputs:

  NS  north-star      H32 d64  T4096 P256 k64  nbf8   bf16 production path (tcgen05 / register top-k / block attention) + fp32
  C2  OPT-125m        H12 d64  T2048 P256 k64  nbf8   bf16 (zero-padded channels on the tcgen05 kernels) + fp32
  C4' OPT-2.7B head   H32 d80  T2048 P256 k64  nbf8   (configs[3] shape at a length the CPU oracle finishes in seconds)
  C5' layer sweep     H32 d128 T4096 P256 k128 nbf8   (configs[4], first point of the sweep)

Gates (north star): top-k alive set and CSR crow/col BIT-EXACT given identical estimated probabilities; fp outputs rtol
1e-3 (fp32) / 2e-2 (bf16); end-to-end mask agreement >= 99.9 % -- counted on the CAUSAL HALF only (the acausal half is
trivially equal).  For bf16 the end-to-end agreement is a property of bf16 rounding in the predictor (near-ties of the
top-k move); it is reported, gated at the level the arithmetic holds, and the test names the first stage whose error
exceeds 2e-2 if any does.
"""
import functools

import numpy as np
import pytest
import torch
import transformers

from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

CONFIGS = {
    'NS': dict(H=32, d=64, T=4096, P=256, k=64, nbf=8),
    'C2': dict(H=12, d=64, T=2048, P=256, k=64, nbf=8),
    'C4h': dict(H=32, d=80, T=2048, P=256, k=64, nbf=8),
    'C5h': dict(H=32, d=128, T=4096, P=256, k=128, nbf=8),
}


@functools.lru_cache(maxsize=None)
def _case(sea_name, name):
    """Module weights, bf16-representable inputs and the oracle's buffers for one config (computed once per session)."""
    import importlib
    sea = importlib.import_module(sea_name)
    c = CONFIGS[name]
    H, d, T, P, k, nbf = (c[x] for x in ('H', 'd', 'T', 'P', 'k', 'nbf'))
    torch.manual_seed(42)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval()
    sd = {k_: v_.detach().clone().float() for k_, v_ in mod.state_dict().items()}
    g = torch.Generator().manual_seed(42)
    # inputs exactly representable in bf16, so the fp32 and the bf16 runs (and the oracle) see identical numbers
    q = (torch.randn(1, H, T, d, generator=g) * d ** -0.5).bfloat16().float()
    kk = torch.randn(1, H, T, d, generator=g).bfloat16().float()
    v = torch.randn(1, H, T, d, generator=g).bfloat16().float()
    with torch.no_grad():
        ref = so.perlin_forward_causal(sd, q, kk, v, k_top=k, P=P, sparse=True, keep_dense=False)
    keep = ('performer_context_layer', 't_attention_predictor', 'estimated_attention_score', 'estimated_attention_probs',
            'partial_attention_mask_before_interp', 'estimated_scales', 'crow_indices', 'col_indices', 'context_layer',
            'average_context_layer')
    ref = {k_: ref[k_] for k_ in keep}
    return mod, sd, (q, kk, v), ref


def _dense_bool(crow, col, H, T):
    """CSR -> bool [H,T,T] (N = 1), numpy, 1 byte per element."""
    crow = crow.reshape(-1).cpu().numpy().astype(np.int64)
    nnz = int(crow[-1])
    col = col.reshape(-1)[:nnz].cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(T), np.diff(crow))
    out = np.zeros((H, T, T), dtype=bool)
    out[col // T, rows, col % T] = True
    return out


def _causal_half_agreement(a, b):
    H, T, _ = a.shape
    assert not np.triu(a[0], 1).any() and not np.triu(b[0], 1).any()          # both strictly causal
    diff = int((a != b).sum())
    return 1.0 - diff / (H * T * (T + 1) / 2), (a == b).all(axis=(0, 2))


def _run(sea, name, dtype):
    mod, sd, (q, kk, v), ref = _case(sea.__name__, name)
    c = CONFIGS[name]
    H, d, T, P, k = c['H'], c['d'], c['T'], c['P'], c['k']
    m = mod.to(DEV)
    m.benchmarking = True
    m.output_attentions = True
    qd, kd, vd = (t.to(dtype).to(DEV) for t in (q, kk, v))
    am = so.causal_additive_mask(T, dtype, 1).to(DEV)
    with torch.no_grad():
        out = m(qd, kd, vd, qd, kd, vd, qd, kd, am, None, None)
        m.output_attentions = False
        out_fast = m(qd, kd, vd, qd, kd, vd, qd, kd, am, None, None)       # the path bench.py times
    torch.cuda.synchronize()
    return out, out_fast, ref, (H, d, T, P, k)


def _check_masks_bit_exact_given_probs(out, H, T, P, k):
    """north star: 'top-k indices and CSR column/row-pointer construction must be bit-exact given identical estimated
    probabilities' -- the oracle's top-k + interpolation are fed the kernel's own probabilities."""
    probs = out.estimated_attention_probs.float().cpu()
    mask_m = so.topk_mask_causal_batch(probs, k)
    crow_r, col_r, Z_r = so.resize_from_m_to_t_csr(mask_m, k, T, True)
    pm = out.partial_attention_mask
    assert pm.crow_indices().dtype == torch.int64 and pm.col_indices().dtype == torch.int64
    assert torch.equal(pm.crow_indices().cpu(), crow_r), 'crow differs from the oracle given identical probabilities'
    assert pm.col_indices().shape[-1] == Z_r
    assert torch.equal(pm.col_indices().cpu(), col_r), 'col differs from the oracle given identical probabilities'
    # exact-size CSR like the reference's (causal_resize_m_to_t.py:757-762): nnz == crow[-1] == number of stored columns.
    # (torch's own `check_sparse_tensor_invariants` cannot be used on this format: the reference emits the columns of a pixel
    # in DESCENDING order, :561-572, which torch's sortedness invariant rejects for the reference's tensors just the same.)
    assert int(pm.crow_indices()[0, -1]) == pm.col_indices().shape[-1] == pm.values().shape[-1]
    assert out.partial_attention_probs.values().shape == pm.values().shape
    return crow_r, col_r


@pytest.mark.parametrize('name', ['C2', 'NS', 'C4h', 'C5h'])
def test_config_fp32_matches_oracle(sea, name):
    out, out_fast, ref, (H, d, T, P, k) = _run(sea, name, torch.float32)
    torch.testing.assert_close(out.estimated_attention_probs.cpu(), ref['estimated_attention_probs'], rtol=2e-3, atol=1e-6)
    _check_masks_bit_exact_given_probs(out, H, T, P, k)
    pm = out.partial_attention_mask
    mine = _dense_bool(pm.crow_indices(), pm.col_indices(), H, T)
    theirs = _dense_bool(ref['crow_indices'], ref['col_indices'], H, T)
    agree, rows_same = _causal_half_agreement(mine, theirs)
    print(f'[{name} fp32] causal-half mask agreement {agree:.6f}; identical rows {int(rows_same.sum())}/{T}')
    assert agree >= 0.999, agree
    rows = torch.from_numpy(rows_same)
    assert rows.float().mean() > 0.9
    torch.testing.assert_close(out.context_layer.cpu()[:, rows], ref['context_layer'][:, rows], rtol=1e-3, atol=3e-5)
    torch.testing.assert_close(out_fast.context_layer.cpu()[:, rows], ref['context_layer'][:, rows], rtol=1e-3, atol=3e-5)


def _rel_err(a, b):
    """|a-b| against the 2e-2 contract: max over elements of |a-b| / (|b| + floor), floor = 2 % of the tensor's rms."""
    a, b = a.float(), b.float()
    floor = 0.02 * float(b.pow(2).mean().sqrt()) + 1e-12
    return float(((a - b).abs() / (b.abs() + floor)).max())


@pytest.mark.parametrize('name', ['C2', 'NS', 'C4h', 'C5h'])
def test_config_bf16_matches_oracle(sea, name):
    out, out_fast, ref, (H, d, T, P, k) = _run(sea, name, torch.bfloat16)
    assert out.context_layer.dtype == torch.bfloat16
    # bit-exact masks given the kernel's own probabilities holds in every dtype
    _check_masks_bit_exact_given_probs(out, H, T, P, k)
    probs = out.estimated_attention_probs.float().cpu()
    # per-stage tensor at the bf16 tolerance: estimated probabilities (the product of Performer -> MLP -> CNN -> softmax)
    p_err = float((probs - ref['estimated_attention_probs']).abs().max())
    p_rel = _rel_err(probs, ref['estimated_attention_probs'])
    pm = out.partial_attention_mask
    mine = _dense_bool(pm.crow_indices(), pm.col_indices(), H, T)
    theirs = _dense_bool(ref['crow_indices'], ref['col_indices'], H, T)
    agree, rows_same = _causal_half_agreement(mine, theirs)
    iou = float((mine & theirs).sum() / max(1, (mine | theirs).sum()))
    print(f'[{name} bf16] causal-half mask agreement {agree:.6f} (IoU {iou:.4f}); identical rows {int(rows_same.sum())}/{T}; '
          f'estimated_attention_probs max abs err {p_err:.3e}, rel {p_rel:.3e}')
    # End-to-end agreement of a bf16 pipeline with the fp32 oracle is limited by bf16 rounding of the predictor
    # activations moving near-ties of the top-k (a 2^-9 relative step on keys that the x4-upsampled predictor makes
    # nearly equal).  North-star gate 99.9 %; the measured value is printed above.
    assert agree >= 0.999, f'{name}: causal-half agreement {agree:.6f} < 0.999 (probs rel err {p_rel:.3e})'
    rows = torch.from_numpy(rows_same)
    if rows.any():
        torch.testing.assert_close(out.context_layer.float().cpu()[:, rows], ref['context_layer'][:, rows], rtol=2e-2, atol=2e-2)
    # the fast path (attention straight from the bit mask: block / gather kernel) must equal the CSR-driven path
    torch.testing.assert_close(out_fast.context_layer.float().cpu(), out.context_layer.float().cpu(), rtol=2e-2, atol=2e-2)
    # every row against the oracle, mask differences included: a moved near-tie swaps one pixel for another of (nearly) the same
    # estimated probability, which changes the few context elements that pixel dominates; measured 0.1-0.3 % of the elements
    d = (out_fast.context_layer.float().cpu() - ref['context_layer']).abs()
    off = float((d > 5e-2 + 5e-2 * ref['context_layer'].abs()).float().mean())
    print(f'[{name} bf16] context elements outside 5e-2 (all rows, mask differences included): {100 * off:.3f} %; mean abs err {float(d.mean()):.2e}')
    assert off < 0.01 and float(d.mean()) < 5e-3


@pytest.mark.parametrize('name', ['C2', 'NS', 'C4h', 'C5h'])
def test_config_bf16_stages_in_isolation(sea, name):
    """Each dense stage of the bf16 production path fed with the ORACLE's input of that stage (rounded to bf16), compared with
    the oracle's output of that stage at rtol 2e-2: names the stage whose rounding is out of contract, if any."""
    mod, sd, (q, kk, v), ref = _case(sea.__name__, name)
    c = CONFIGS[name]
    H, d, T, P, k = c['H'], c['d'], c['T'], c['P'], c['k']
    m = mod.to(DEV)
    w = m._weights_fp32()
    bf = lambda t: t.bfloat16().to(DEV)
    # a2+a3 Performer
    ctx, avg = sea.ops.performer_causal(bf(q), bf(kk), bf(v), w['pos'], w['proj'])
    torch.testing.assert_close(ctx.float().cpu(), ref['performer_context_layer'], rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(avg.float().cpu(), ref['average_context_layer'], rtol=2e-2, atol=2e-2)
    # a4 MLP (+ the first CNN LayerNorm) from the oracle's Performer output
    S, W = 2, P // 4
    cnn_in, scales, _ = sea.ops.predictor_mlp(bf(ref['performer_context_layer']), bf(v), w, S, W)
    t_ref = ref['t_attention_predictor']
    dec = so.predictor_dec_row(t_ref, sd, S)
    x0 = so.layer_norm(dec, sd['attention_predictor_cnn.0.module.weight'], sd['attention_predictor_cnn.0.module.bias'])
    torch.testing.assert_close(scales.cpu(), ref['estimated_scales'], rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(cnn_in.float().cpu().permute(0, 3, 1, 2), x0, rtol=2e-2, atol=4e-2)
    # a5 first conv from the oracle's CNN input
    p_ = 'attention_predictor_cnn.1.module.net.'
    y_ref = torch.relu(so.causal_conv2d(x0, sd[p_ + '0.module.weight'], sd[p_ + '0.module.weight_mask'], sd[p_ + '0.module.bias'], 3, 2, 2))
    y = sea.ops.causal_conv3x3_dil2_relu(bf(x0.permute(0, 2, 3, 1).contiguous()), w['conv1_w'], w['conv1_b'])
    # a K = 9 x 2H dot product of bf16 operands: error relative to the tensor's scale
    scale = float(y_ref.abs().mean())
    assert float((y.float().cpu().permute(0, 3, 1, 2) - y_ref).abs().max()) <= 2e-2 * max(1.0, float(y_ref.abs().max())), 'conv1 out of the bf16 contract'
    assert float((y.float().cpu().permute(0, 3, 1, 2) - y_ref).abs().mean()) <= 2e-2 * scale
