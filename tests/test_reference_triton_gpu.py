"""GPU: the UNMODIFIED reference (oracle/_ref, the verbatim copy `make -C oracle` carries to the GPU box) with its Triton
kernels COMPILED for the B200 (Triton 3.6, one shim: `tl.math.round = libdevice.round`) against this repo's CUDA
path, on the same inputs, on the device.  This is the secondary oracle of SURVEY 8c: it settles what the numpy-interpreted
fixtures cannot -- e.g. that the compiled kernel's `div.full.f32` sub-sampling of clamped pixels (T/P > k,
causal_resize_m_to_t.py:561-572) gives the same column ids as the IEEE division used here.

Skipped when no copy of the reference travelled (oracle/_ref absent) or Triton cannot compile on the box.
"""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ref_ops():
    if not rh.reference_available():
        pytest.skip('no copy of the reference on this box (oracle/_ref absent)')
    try:
        rh.load_reference(interpret_triton=False)
        import src.models.perlin_attention.ops as ops
        # one tiny launch proves that Triton compiles and runs here
        x = torch.ones(1, 1, 4, 2, device=DEV)
        ops.resize_from_m_to_t_csr(x, 0, 2, target_width=4, is_causal=True, oversampled=1.0)
    except Exception as e:       # pragma: no cover - depends on the box
        pytest.skip(f'reference Triton path does not run here: {e!r}'[:300])
    return ops


def _topk_mask(sea, N, H, T, P, k, seed):
    g = torch.Generator().manual_seed(seed)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g) * 2, -1).to(DEV)
    kpr = torch.from_numpy(np.tile(so.per_item_top_k_causal(H, k, 1.0, P, T), N)).to(DEV)
    bits = sea.ops.topk_mask_bits(probs, kpr, 'causal_batch')
    return sea.ops.bits_to_mask(bits, H, P), bits


CSR_CASES = [
    # N, H, T, P, k, causal
    (1, 32, 4096, 256, 64, True),        # north-star shape
    (1, 12, 2048, 256, 64, True),        # C2
    (2, 3, 1000, 96, 16, True),          # P not a power of two, two batch items
    (1, 2, 4096, 32, 16, True),          # T/P = 128 > k: clamped, sub-sampled pixels (div.full.f32 in the compiled reference)
    (1, 4, 8192, 64, 32, True),          # T/P = 128 > k = 32
    (1, 32, 8192, 128, 16, True),        # C5-like clamp regime: T/P = 64 > k = 16
    (1, 4, 512, 128, 64, False),         # non-causal (BERT)
]


@pytest.mark.parametrize('N,H,T,P,k,causal', CSR_CASES)
def test_csr_interpolation_equals_compiled_reference(sea, ref_ops, N, H, T, P, k, causal):
    mask, bits = _topk_mask(sea, N, H, T, P, k, seed=T + P + k)
    ref = ref_ops.resize_from_m_to_t_csr(mask, 0, k, target_width=T, is_causal=causal, oversampled=1.0)
    mine = sea.ops.resize_from_m_to_t_csr(mask, 0, k, target_width=T, is_causal=causal, oversampled=1.0)
    torch.cuda.synchronize()
    assert mine.shape == ref.shape
    if P & (P - 1):
        # P not a power of two: `scales = target_width / original_width` (causal_resize_m_to_t.py:642) is IEEE division on the CPU but
        # L * (1 / P) in torch's CUDA true-division kernel, so the reference's own CPU and GPU runs differ in a few pixel edges.  This
        # repo follows the CPU arithmetic (the fixtures of tests/golden).  Pin both statements exactly: the compiled reference equals
        # the oracle evaluated with the CUDA scale, this repo equals the oracle evaluated with the IEEE scale.
        crow_g, col_g, _ = so.resize_from_m_to_t_csr(mask.cpu(), k, T, causal, scale_mode='cuda_reciprocal')
        crow_c, col_c, _ = so.resize_from_m_to_t_csr(mask.cpu(), k, T, causal, scale_mode='ieee')
        assert torch.equal(ref.crow_indices().cpu(), crow_g), 'compiled reference != oracle with the CUDA reciprocal scale'
        assert torch.equal(mine.crow_indices().cpu(), crow_c), 'this repo != oracle with the IEEE scale'
        for n in range(N):
            assert torch.equal(ref.col_indices()[n, :int(crow_g[n, -1])].cpu(), col_g[n, :int(crow_g[n, -1])])
            assert torch.equal(mine.col_indices()[n, :int(crow_c[n, -1])].cpu(), col_c[n, :int(crow_c[n, -1])])
        rows = int(((crow_g[:, 1:] - crow_g[:, :-1]) != (crow_c[:, 1:] - crow_c[:, :-1])).sum())
        print(f'[csr {N},{H},{T},{P},{k},{causal}] P not a power of two: reference-on-GPU == oracle(L * (1/P)), this repo == oracle(L / P); '
              f'{rows} of {N * T} rows differ in length between the two scales')
        return
    assert torch.equal(mine.crow_indices(), ref.crow_indices().to(mine.crow_indices().dtype)), 'crow differs from the compiled reference'
    # compare the live part of every item (the reference leaves the tail past an item's nnz at whatever its buffer held)
    for n in range(N):
        nnz = int(ref.crow_indices()[n, -1])
        assert torch.equal(mine.col_indices()[n, :nnz], ref.col_indices()[n, :nnz].to(mine.col_indices().dtype)), \
            f'col differs from the compiled reference (item {n})'
    clamped = (T + P - 1) // P > k
    print(f'[csr {N},{H},{T},{P},{k},{causal}] nnz {int(ref.crow_indices()[0, -1])} clamped pixels: {clamped}: bit-exact')


@pytest.mark.parametrize('H,T,P,k,d,dtype', [(32, 4096, 256, 64, 64, torch.float32), (12, 2048, 256, 64, 64, torch.bfloat16),
                                            (4, 1024, 64, 32, 80, torch.float32), (8, 2048, 128, 128, 128, torch.float32)])
def test_flat_csr_ops_equal_compiled_reference(sea, ref_ops, H, T, P, k, d, dtype):
    """flat_csr_masked_bmm -> softmax -> elmul -> sdbmm: the reference's four Triton ops vs this repo's four ops, chained."""
    N = 1
    mask, bits = _topk_mask(sea, N, H, T, P, k, seed=H + d)
    csr_ref = ref_ops.resize_from_m_to_t_csr(mask, 0, k, target_width=T, is_causal=True, oversampled=1.0)
    csr = sea.ops.resize_from_m_to_t_csr(mask, 0, k, target_width=T, is_causal=True, oversampled=1.0)
    g = torch.Generator().manual_seed(d)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).to(dtype).to(DEV)
    kk = torch.randn(N, H, T, d, generator=g).to(dtype).to(DEV)
    v = torch.randn(N, H, T, d, generator=g).to(dtype).to(DEV)
    scaler = torch.sigmoid(torch.randn(N, H, T, 1, generator=g)).to(DEV).expand(N, H, T, T)
    s_r = ref_ops.flat_csr_masked_bmm(q, kk, csr_ref)
    s_m = sea.ops.flat_csr_masked_bmm(q, kk, csr)
    tol = dict(rtol=1e-3, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    if dtype == torch.float32:
        torch.testing.assert_close(s_m.values().float(), s_r.values().float(), **tol)
    else:
        # bf16: both sides round a 64-term dot product to bf16 with their own summation order; judge both against the fp64 truth
        nnz = int(csr.crow_indices()[0, -1])
        rows = torch.repeat_interleave(torch.arange(T, device=DEV), csr.crow_indices()[0, 1:] - csr.crow_indices()[0, :-1])
        c = csr.col_indices()[0, :nnz].long()
        truth = (q[0, c // T, rows].double() * kk[0, c // T, c % T].double()).sum(-1)
        e_m = (s_m.values()[0, :nnz].double() - truth).abs()
        e_r = (s_r.values()[0, :nnz].double() - truth).abs()
        bound = 2e-2 * truth.abs() + 2e-2
        assert bool((e_m <= bound).all()), 'masked_bmm (bf16) outside rtol 2e-2 of the fp64 truth'
        print(f'[flat_csr bf16] masked_bmm max abs err vs fp64 truth: this repo {float(e_m.max()):.3e}, compiled reference {float(e_r.max()):.3e}')
    p_r = ref_ops.flat_csr_softmax(s_r, H, T)
    p_m = sea.ops.flat_csr_softmax(s_r, H, T)
    torch.testing.assert_close(p_m.values().float(), p_r.values().float(), rtol=1e-3, atol=1e-6)
    e_r = ref_ops.flat_csr_elmul(p_r, scaler)
    e_m = sea.ops.flat_csr_elmul(p_r, scaler)
    torch.testing.assert_close(e_m.values().float(), e_r.values().float(), rtol=1e-3, atol=1e-7)
    c_r = ref_ops.flat_csr_sdbmm(e_r, v, P)
    c_m = sea.ops.flat_csr_sdbmm(e_r, v, P)
    torch.testing.assert_close(c_m.float(), c_r.float(), **tol)
    # and the fused kernel the module runs (attention straight from the bit mask) against the reference's chain
    if dtype == torch.bfloat16 and d in (32, 64, 128):
        scales = torch.zeros(N, H, T, 2, device=DEV)
        scales[..., 0] = torch.logit(scaler[..., 0].clamp(1e-6, 1 - 1e-6))
        scales[..., 1] = 30.0                                       # sigmoid -> 1: no running-mean mix
        avg = torch.zeros(N, H, T, d, dtype=dtype, device=DEV)
        fused = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True)
        torch.testing.assert_close(fused.float(), c_r.float().permute(0, 2, 1, 3).reshape(N, T, H * d), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('H,d,T,P,k,nbf', [(12, 64, 2048, 256, 64, 8), (32, 64, 4096, 256, 64, 8)])
def test_whole_layer_equals_reference_sparse_path_on_gpu(sea, ref_ops, H, d, T, P, k, nbf):
    """The reference module itself (benchmarking=True: Triton sparse path) on the B200 in fp32 vs the drop-in module in fp32
    and bf16, same weights (state_dict copied), same inputs: estimated probabilities, CSR mask, context."""
    import time
    # the reference in true fp32: cuDNN convolutions default to TF32 on this GPU, which alone moves its probabilities by ~4e-3 relative
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_mod = rh.build_reference_attention(H, d, T, k, P, nbf, True).to(DEV)
    ref_mod.benchmarking = True
    g = torch.Generator().manual_seed(7)
    q = (torch.randn(1, H, T, d, generator=g) * d ** -0.5).bfloat16().float().to(DEV)
    kk = torch.randn(1, H, T, d, generator=g).bfloat16().float().to(DEV)
    v = torch.randn(1, H, T, d, generator=g).bfloat16().float().to(DEV)
    am = so.causal_additive_mask(T, torch.float32, 1).to(DEV)
    with torch.no_grad():
        ro = ref_mod(q, kk, v, q, kk, v, q, kk, am, None, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ref_mod(q, kk, v, q, kk, v, q, kk, am, None, None)
        torch.cuda.synchronize()
        ref_ms = (time.perf_counter() - t0) / 3 * 1e3
    import transformers
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval()
    missing, unexpected = mod.load_state_dict(ref_mod.state_dict(), strict=False)
    assert not unexpected, unexpected
    mod = mod.to(DEV)
    mod.benchmarking = True
    mod.output_attentions = True
    with torch.no_grad():
        mo = mod(q, kk, v, q, kk, v, q, kk, am, None, None)
    torch.cuda.synchronize()
    torch.testing.assert_close(mo.estimated_attention_probs, ro.estimated_attention_probs.float(), rtol=2e-3, atol=1e-6)
    r_crow, r_col = ro.partial_attention_mask.crow_indices(), ro.partial_attention_mask.col_indices()
    m_crow, m_col = mo.partial_attention_mask.crow_indices(), mo.partial_attention_mask.col_indices()

    def dense(crow, col):
        out = torch.zeros((H, T, T), dtype=torch.bool, device=DEV)
        nnz = int(crow[0, -1])
        rows = torch.repeat_interleave(torch.arange(T, device=DEV), crow[0, 1:] - crow[0, :-1])
        c = col[0, :nnz].long()
        out[c // T, rows, c % T] = True
        return out
    a, b = dense(m_crow, m_col), dense(r_crow, r_col)
    agree = 1.0 - float((a != b).sum()) / (H * T * (T + 1) / 2)
    rows_same = (a == b).all(dim=2).all(dim=0)
    err = float((mo.context_layer[:, rows_same] - ro.context_layer[:, rows_same].float()).abs().max())
    print(f'[ref-gpu H{H} T{T}] reference (Triton, fp32, eager) {ref_ms:.2f} ms/layer-forward on this GPU; fp32 causal-half mask agreement '
          f'{agree:.6f}, identical rows {int(rows_same.sum())}/{T}, context max abs err on them {err:.2e}')
    # ties between equal keys are implementation-defined in the reference (torch.sort, unstable) -- everything else must agree
    assert agree >= 0.999
    torch.testing.assert_close(mo.context_layer[:, rows_same], ro.context_layer[:, rows_same].float(), rtol=1e-3, atol=3e-5)
    # bf16 production path vs the reference's fp32 result
    mod.output_attentions = False
    with torch.no_grad():
        amb = so.causal_additive_mask(T, torch.bfloat16, 1).to(DEV)
        qb, kb, vb = q.bfloat16(), kk.bfloat16(), v.bfloat16()
        bo = mod(qb, kb, vb, qb, kb, vb, qb, kb, amb, None, None)
    # every row, mask differences included: bf16 rounding of the predictor moves top-k near-ties, which swaps one pixel for another of
    # (nearly) the same estimated probability and changes the few context elements that pixel dominates
    dd = (bo.context_layer.float() - ro.context_layer.float()).abs()
    off = float((dd > 5e-2 + 5e-2 * ro.context_layer.float().abs()).float().mean())
    print(f'[ref-gpu H{H} T{T}] bf16 production path vs the reference (fp32): {100 * off:.3f} % of the context elements outside 5e-2, '
          f'mean abs err {float(dd.mean()):.2e}')
    assert off < 0.01 and float(dd.mean()) < 5e-3
