"""GPU parity tests of the operator boundary (the 7 exports of the reference's ops/__init__.py plus the
top-k), called through the C ABI (ops.py -> ctypes -> libsea_b200.so) and compared with the CPU oracle
on the same seeded inputs.  Integer / index results: bit-exact.  Floating point: rtol 1e-3 (fp32)."""
import numpy as np
import pytest
import torch

from conftest import golden_layer, load_golden
from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rand_mask(N, H, T, P, k, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g), -1)
    if ties:
        probs = probs[..., : max(P // 4, 1)].repeat_interleave(4, dim=-1)[..., :P].contiguous()
    return probs, so.topk_mask_causal_batch(probs, k)


@pytest.mark.parametrize('N,H,T,P,k,ties', [
    (2, 3, 50, 16, 4, False), (1, 4, 128, 32, 8, True), (1, 12, 64, 256, 64, True), (1, 32, 96, 256, 64, True),
    (1, 1, 64, 8, 16, True), (2, 5, 33, 24, 3, False), (1, 64, 8, 256, 64, True),
])
def test_topk_bit_exact(sea, N, H, T, P, k, ties):
    probs, ref = _rand_mask(N, H, T, P, k, seed=N * 1000 + H * 10 + T, ties=ties)
    kpr = torch.from_numpy(np.tile(so.per_item_top_k_causal(H, k, 1.0, P, T), N))
    bits = sea.ops.topk_mask_bits(probs.to(DEV), kpr.to(DEV), 'causal_batch')
    got = sea.ops.bits_to_mask(bits, H, P).cpu()
    assert torch.equal(got, ref)
    # round trip float mask -> bits -> float mask
    assert torch.equal(sea.ops.bits_to_mask(sea.ops.mask_to_bits(ref.to(DEV)), H, P).cpu(), ref)


def test_topk_edge_cases(sea):
    N, H, T, P = 1, 2, 4, 32
    probs = torch.zeros(N, H, T, P)          # all keys equal (+0 and -0 mixed): lowest indices win
    probs[0, 1, 1, :] = -0.0
    for kval in (1.0, 5.0, 64.0, 1000.0):
        kpr = torch.full((N * T,), kval)
        got = sea.ops.bits_to_mask(sea.ops.topk_mask_bits(probs.to(DEV), kpr.to(DEV)), H, P).cpu()
        alive = so.topk_alive_rows(probs.transpose(1, 2).reshape(N * T, H * P).numpy() + 0.0, kpr.numpy())
        ref = torch.from_numpy(alive.reshape(N, T, H, P)).permute(0, 2, 1, 3).float()
        assert torch.equal(got, ref), kval
    # padded query rows are dead (attention.py:928-931)
    rv = torch.tensor([[1, 0, 1, 0]], dtype=torch.uint8)
    got = sea.ops.bits_to_mask(sea.ops.topk_mask_bits(probs.to(DEV), torch.full((4,), 3.0).to(DEV), row_valid=rv.to(DEV)), H, P).cpu()
    assert got[0, :, 1].sum() == 0 and got[0, :, 3].sum() == 0 and got[0, :, 0].sum() == 3


def test_topk_query_mode(sea):
    N, H, T, P = 2, 3, 17, 24
    g = torch.Generator().manual_seed(5)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g), -1)
    kq = torch.tensor([5.0, 9.0])
    got = sea.ops.bits_to_mask(sea.ops.topk_mask_bits(probs.to(DEV), kq.to(DEV), 'query'), H, P).cpu()
    alive = so.topk_alive_rows(probs.reshape(N * H * T, P).numpy(), kq.view(N, 1).expand(N, H * T).reshape(-1).numpy())
    assert torch.equal(got, torch.from_numpy(alive.reshape(N, H, T, P)).float())


CSR_CASES = [
    # N, H, T_DST, T_SRC, P, k, causal
    (1, 1, 64, 64, 8, 16, True), (2, 3, 100, 100, 16, 8, True), (1, 4, 257, 257, 32, 16, True),
    (1, 2, 1024, 1024, 8, 16, True),      # T/P > k: clamp + sub-sampling branch
    (1, 3, 100, 100, 16, 6, True),        # clamp active on late rows only
    (1, 4, 40, 128, 32, 8, True),         # T_DST < T_SRC (last rows of a longer context)
    (2, 4, 64, 64, 32, 8, False),         # non causal
    (1, 12, 96, 96, 24, 5, True),         # P not a power of two
]


@pytest.mark.parametrize('N,H,T_DST,T_SRC,P,k,causal', CSR_CASES)
@pytest.mark.parametrize('idx_dtype', [torch.int64, torch.int32])
def test_csr_interpolation_bit_exact(sea, N, H, T_DST, T_SRC, P, k, causal, idx_dtype):
    g = torch.Generator().manual_seed(T_DST + P)
    mask = (torch.rand(N, H, T_DST, P, generator=g) < min(1.0, 1.5 * k / P)).float()
    mask[0, :, 0, :] = 0           # an empty row
    crow_r, col_r, Z_r = so.resize_from_m_to_t_csr(mask, k, T_SRC, causal)
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    crow, col, Z = sea.ops.csr_from_bits(bits, H, P, k, T_SRC, is_causal=causal, index_dtype=idx_dtype)
    assert Z == Z_r
    assert torch.equal(crow.cpu().long(), crow_r)
    assert torch.equal(col.cpu().long(), col_r)
    # over-allocated (sync-free) variant: same prefix, zero tail
    crow2, col2, Z2 = sea.ops.csr_from_bits(bits, H, P, k, T_SRC, is_causal=causal, index_dtype=idx_dtype, z_alloc=Z_r + 77)
    assert torch.equal(col2.cpu().long()[:, :Z_r], col_r) and int(col2[:, Z_r:].abs().sum()) == 0


def test_resize_csr_op_matches_reference_fixture(sea):
    """resize_from_m_to_t_csr (the op, torch CSR in/out) against the reference's own (Triton-interpreted) output."""
    g = load_golden('kat_causal_resize')
    N, H, T, P, K = g['meta'].tolist()
    csr = sea.resize_from_m_to_t_csr(torch.from_numpy(g['compressed_mask']).to(DEV), 0, K, target_width=T)
    assert csr.is_sparse_csr and tuple(csr.shape) == (N, T, H * T)
    assert np.array_equal(csr.crow_indices().cpu().numpy(), g['crow'])
    assert np.array_equal(csr.col_indices().cpu().numpy(), g['col'])
    assert np.diff(csr.crow_indices().cpu().numpy()[0]).tolist() == g['nnz_per_row_notebook'].tolist()
    dense = sea.flat_csr_to_dense(csr, T, H)
    assert np.array_equal(dense.cpu().numpy().astype(np.uint8), g['dense'])
    for name in ('layer_causal_h4_t128', 'layer_causal_h3_t100'):
        g, m, _ = golden_layer(name)
        H, T, P = m['H'], m['T'], m['P']
        mm = np.unpackbits(g['sparse.mask_before_interp'])[:H * T * P].reshape(1, H, T, P).astype(np.float32)
        csr = sea.resize_from_m_to_t_csr(torch.from_numpy(mm).to(DEV), 0, m['k'], target_width=T)
        assert np.array_equal(csr.crow_indices().cpu().numpy(), g['sparse.crow'])
        assert np.array_equal(csr.col_indices().cpu().numpy(), g['sparse.col'].astype(np.int64))


def _csr_case(N, H, T, P, k, d, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    probs, mask = _rand_mask(N, H, T, P, k, seed)
    crow, col, Z = so.resize_from_m_to_t_csr(mask, k, T, True)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).to(dtype)
    kk = torch.randn(N, H, T, d, generator=g).to(dtype)
    v = torch.randn(N, H, T, d, generator=g).to(dtype)
    return crow, col, Z, q, kk, v


def _mk_csr(crow, col, vals, shape):
    return torch.sparse_csr_tensor(crow.to(DEV), col.to(DEV), vals.to(DEV), size=shape)


@pytest.mark.parametrize('N,H,T,P,k,d', [(2, 3, 100, 16, 8, 32), (1, 4, 128, 32, 8, 64), (1, 2, 96, 32, 8, 80), (1, 2, 64, 16, 8, 128)])
def test_flat_csr_ops_match_oracle(sea, N, H, T, P, k, d):
    crow, col, Z, q, kk, v = _csr_case(N, H, T, P, k, d, seed=d + T)
    shape = (N, T, H * T)
    mask = _mk_csr(crow, col, torch.ones(N, Z), shape)
    # a9
    s = sea.flat_csr_masked_bmm(q.to(DEV), kk.to(DEV), mask)
    s_ref = so.flat_csr_masked_bmm(q, kk, crow, col)
    torch.testing.assert_close(s.values().cpu(), s_ref, rtol=1e-3, atol=1e-5)
    # a10
    p = sea.flat_csr_softmax(s, H, T)
    p_ref = so.flat_csr_softmax(s_ref, crow, col, H, T)
    torch.testing.assert_close(p.values().cpu(), p_ref, rtol=1e-3, atol=1e-6)
    # a11 with the stride-0 row scaler of attention.py:1170
    rs = torch.sigmoid(torch.randn(N, H, T, generator=torch.Generator().manual_seed(3)))
    p2 = sea.flat_csr_elmul(p, rs.to(DEV).view(N, H, T, 1).expand(N, H, T, T))
    p2_ref = so.flat_csr_elmul_rowscale(p_ref, crow, col, rs, T)
    torch.testing.assert_close(p2.values().cpu(), p2_ref, rtol=1e-3, atol=1e-6)
    # a12
    o = sea.flat_csr_sdbmm(p2, v.to(DEV), P)
    o_ref = so.flat_csr_sdbmm(p2_ref, crow, col, v, H)
    assert o.dtype == torch.float32 and tuple(o.shape) == (N, H, T, d)
    torch.testing.assert_close(o.cpu(), o_ref, rtol=1e-3, atol=1e-5)
    # to_dense
    dn = sea.flat_csr_to_dense(p2, T, H)
    torch.testing.assert_close(dn.cpu(), so.flat_csr_to_dense(crow, col, p2.values().cpu(), T, H), rtol=0, atol=0)


def test_flat_csr_ops_empty_and_ragged(sea):
    """Empty rows, an empty batch item next to a full one, rows whose heads are absent."""
    N, H, T, d = 2, 3, 8, 32
    crow = torch.zeros(N, T + 1, dtype=torch.int64)
    cols0 = []
    for t in range(T):
        if t % 3 == 0:
            continue                      # empty row
        cols0 += [0 * T + j for j in range(t + 1)] + [2 * T + t]      # head 1 absent everywhere
        crow[0, t + 1:] = len(cols0)
    Z = len(cols0)
    col = torch.zeros(N, Z, dtype=torch.int64)
    col[0] = torch.tensor(cols0)
    g = torch.Generator().manual_seed(0)
    q, kk, v = (torch.randn(N, H, T, d, generator=g) for _ in range(3))
    mask = _mk_csr(crow, col, torch.ones(N, Z), (N, T, H * T))
    s = sea.flat_csr_masked_bmm(q.to(DEV), kk.to(DEV), mask)
    p = sea.flat_csr_softmax(s, H, T)
    o = sea.flat_csr_sdbmm(p, v.to(DEV), 4)
    s_ref = so.flat_csr_masked_bmm(q, kk, crow, col)
    p_ref = so.flat_csr_softmax(s_ref, crow, col, H, T)
    torch.testing.assert_close(p.values().cpu()[0], p_ref[0], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(o.cpu(), so.flat_csr_sdbmm(p_ref, crow, col, v, H), rtol=1e-3, atol=1e-5)
    assert float(o[1].abs().sum()) == 0.0 and float(o[0, 1].abs().sum()) == 0.0


@pytest.mark.parametrize('causal', [True, False])
def test_dense_resize_matches_oracle(sea, causal):
    N, H, T, P, k = 2, 3, 70, 16, 8
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, H, T, P, generator=g)
    if causal:
        am = so.causal_additive_mask(T, torch.float32, N)
    else:
        am = torch.zeros(N, 1, 1, T)
        am[1, 0, 0, 50:] = so.fp_min_for(torch.float32)         # padded tail on item 1
    ref = so.resize_from_m_to_t_dense(x, -7.0, am, T, causal, k, None)
    got = sea.resize_from_m_to_t(x.to(DEV), -7.0, am.to(DEV), target_width=T, is_causal=causal, k=k)
    assert torch.equal(got.cpu(), ref)


def test_fused_sparse_attention_matches_unfused_ops(sea):
    N, H, T, P, k, d = 1, 4, 128, 32, 8, 64
    crow, col, Z, q, kk, v = _csr_case(N, H, T, P, k, d, seed=77)
    g = torch.Generator().manual_seed(9)
    scales = torch.randn(N, H, T, 2, generator=g)
    avg = v.cumsum(-2) / torch.arange(1, T + 1).view(1, 1, T, 1)
    s_ref = so.flat_csr_masked_bmm(q, kk, crow, col)
    p_ref = so.flat_csr_elmul_rowscale(so.flat_csr_softmax(s_ref, crow, col, H, T), crow, col, torch.sigmoid(scales[..., 0]), T)
    ctx = so.flat_csr_sdbmm(p_ref, crow, col, v, H)
    a = torch.sigmoid(scales[..., 1:2])
    ref = (ctx * a + (1 - a) * avg).permute(0, 2, 1, 3).reshape(N, T, H * d)
    for idt in (torch.int64, torch.int32):
        out, pv = sea.ops.sparse_attention(crow.to(DEV).to(idt), col.to(DEV).to(idt), q.to(DEV), kk.to(DEV), v.to(DEV),
                                           scales.to(DEV), avg.to(DEV), use_scaler=True, want_probs=True)
        torch.testing.assert_close(out.cpu(), ref, rtol=1e-3, atol=2e-5)
        torch.testing.assert_close(pv.cpu(), p_ref, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize('d,idt', [(64, torch.int32), (128, torch.int32), (32, torch.int64), (64, torch.int64)])
def test_fused_sparse_attention_v2_bf16_with_head_ptr(sea, d, idt):
    """16-bit warp-per-(row, head) kernel driven by the head_ptr index that sea_csr_fill emits."""
    N, H, T, P, k = 2, 4, 160, 32, 8
    g = torch.Generator().manual_seed(d)
    probs, mask = _rand_mask(N, H, T, P, k, seed=d + 1)
    mask[1, 2] = 0                  # a head with no entries at all in item 1
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    crow, col, Z, hp = sea.ops.csr_from_bits(bits, H, P, k, T, True, idt, want_head_ptr=True)
    crow_r, col_r, Z_r = so.resize_from_m_to_t_csr(mask, k, T, True)
    assert torch.equal(col.cpu().long(), col_r)
    # head_ptr is the lower bound of h*T inside each row
    hp_c, crow_c = hp.cpu().long(), crow.cpu().long()
    assert torch.equal(hp_c[:, :, 0], crow_c[:, :-1]) and torch.equal(hp_c[:, :, H], crow_c[:, 1:])
    for n in range(N):
        for t in (0, 7, T - 1):
            row = col_r[n, crow_r[n, t]:crow_r[n, t + 1]]
            for h in range(H + 1):
                assert int(hp_c[n, t, h]) == int(crow_r[n, t]) + int((row < h * T).sum())
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16()
    kk = torch.randn(N, H, T, d, generator=g).bfloat16()
    v = torch.randn(N, H, T, d, generator=g).bfloat16()
    scales = torch.randn(N, H, T, 2, generator=g)
    avg = (v.float().cumsum(-2) / torch.arange(1, T + 1).view(1, 1, T, 1)).bfloat16()
    s_ref = so.flat_csr_masked_bmm(q.float(), kk.float(), crow_r, col_r)
    p_ref = so.flat_csr_elmul_rowscale(so.flat_csr_softmax(s_ref, crow_r, col_r, H, T), crow_r, col_r, torch.sigmoid(scales[..., 0]), T)
    ctx = so.flat_csr_sdbmm(p_ref, crow_r, col_r, v.float(), H)
    a = torch.sigmoid(scales[..., 1:2])
    ref = (ctx * a + (1 - a) * avg.float()).permute(0, 2, 1, 3).reshape(N, T, H * d)
    out, pv = sea.ops.sparse_attention(crow, col, q.to(DEV), kk.to(DEV), v.to(DEV), scales.to(DEV), avg.to(DEV), True, True, head_ptr=hp)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=2e-2, atol=1e-2)
    torch.testing.assert_close(pv.cpu(), p_ref, rtol=1e-3, atol=1e-6)
    # the binary-search kernel (no head_ptr) must agree
    out2, pv2 = sea.ops.sparse_attention(crow, col, q.to(DEV), kk.to(DEV), v.to(DEV), scales.to(DEV), avg.to(DEV), True, True)
    torch.testing.assert_close(out.float().cpu(), out2.float().cpu(), rtol=2e-2, atol=1e-2)


@pytest.mark.parametrize('N,H,T_DST,T_SRC,P,k,d,causal', [
    (2, 4, 160, 160, 32, 8, 64, True), (1, 2, 1024, 1024, 32, 4, 64, True),     # second: T/P > k -> clamped, sub-sampled pixels
    (1, 3, 40, 128, 64, 8, 128, True), (2, 4, 64, 64, 32, 8, 32, False), (1, 32, 70, 70, 256, 64, 64, True),
    # d = 64 and no clamped pixel -> the tile-skipping block kernel (block_attn.cu): ragged row blocks, T_DST < T_SRC, non-causal
    (1, 2, 1000, 1000, 64, 32, 64, True), (1, 3, 40, 128, 64, 8, 64, True), (2, 4, 200, 200, 32, 8, 64, False),
    (1, 2, 2048, 2048, 256, 64, 64, True),
    (1, 2, 300, 300, 96, 16, 64, True),           # P not a power of two: the float pixel-edge path of the mask expansion
])
def test_attention_from_bits_equals_csr_path(sea, N, H, T_DST, T_SRC, P, k, d, causal):
    """The bit-mask driven kernel enumerates exactly the entries sea_csr_fill would emit."""
    g = torch.Generator().manual_seed(P + d)
    mask = (torch.rand(N, H, T_DST, P, generator=g) < min(1.0, 2.0 * k / P)).float()
    mask[0, 1, 3] = 0
    mask[0, 0, 5] = 1               # a head with every pixel alive
    if T_DST >= 256:
        mask[0, 0, 128:256] = 0     # a whole 128-row query block of one head without any alive pixel (no active tile)
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    crow, col, Z, hp = sea.ops.csr_from_bits(bits, H, P, k, T_SRC, causal, torch.int32, want_head_ptr=True)
    q = (torch.randn(N, H, T_DST, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(N, H, T_SRC, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(N, H, T_SRC, d, generator=g).bfloat16().to(DEV)
    scales = torch.randn(N, H, T_DST, 2, generator=g).to(DEV)
    avg = torch.randn(N, H, T_DST, d, generator=g).bfloat16().to(DEV)
    ref, _ = sea.ops.sparse_attention(crow, col, q, kk, v, scales, avg, True, False, head_ptr=hp)
    out = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, causal)
    torch.testing.assert_close(out.float().cpu(), ref.float().cpu(), rtol=1e-2, atol=1e-2)
    assert float((out.float() - ref.float()).abs().mean()) < 1e-3
    # both kernels behind the op: the per-(row, head) gather kernel and (where the shape allows) the block kernel
    out_g = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, causal, kernel='gather')
    torch.testing.assert_close(out_g.float().cpu(), ref.float().cpu(), rtol=1e-2, atol=1e-2)
    if sea._lib.load().sea_block_attention_workspace_bytes(N, H, T_DST, T_SRC, d, P, k, sea._lib.SEA_DTYPE_BF16) > 0:
        out_b = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, causal, kernel='block')
        torch.testing.assert_close(out_b.float().cpu(), ref.float().cpu(), rtol=1e-2, atol=1e-2)
        out_h = sea.ops.sparse_attention_from_bits(bits, q.half(), kk.half(), v.half(), scales, avg.half(), P, k, True, causal, kernel='block')
        torch.testing.assert_close(out_h.float().cpu(), ref.float().cpu(), rtol=2e-2, atol=2e-2)
    else:
        assert d != 64 or (T_SRC + P - 1) // P + 1 > k


@pytest.mark.parametrize('H,T,P,k,kernel', [(2, 2048, 256, 64, 'block'), (3, 1000, 64, 32, 'block'), (2, 2048, 256, 64, 'gather')])
def test_attention_from_bits_matches_oracle_directly(sea, H, T, P, k, kernel):
    """The tcgen05 block attention (and the gather kernel) against the CPU oracle's flat-CSR chain itself
    (flat_csr_masked_bmm -> softmax -> elmul -> sdbmm -> mix; attention.py:1151-1173, 1237-1250), not via the repo's CSR kernel."""
    N, d = 1, 64
    g = torch.Generator().manual_seed(T + P)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g) * 2, -1)
    mask = so.topk_mask_causal_batch(probs, k)
    crow, col, Z = so.resize_from_m_to_t_csr(mask, k, T, True)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16()
    kk = torch.randn(N, H, T, d, generator=g).bfloat16()
    v = torch.randn(N, H, T, d, generator=g).bfloat16()
    scales = torch.randn(N, H, T, 2, generator=g)
    avg = (v.float().cumsum(-2) / torch.arange(1, T + 1).view(1, 1, T, 1)).bfloat16()
    s = so.flat_csr_masked_bmm(q.float(), kk.float(), crow, col)
    p = so.flat_csr_elmul_rowscale(so.flat_csr_softmax(s, crow, col, H, T), crow, col, torch.sigmoid(scales[..., 0]), T)
    ctx = so.flat_csr_sdbmm(p, crow, col, v.float(), H)
    a = torch.sigmoid(scales[..., 1:2])
    ref = (ctx * a + (1 - a) * avg.float()).permute(0, 2, 1, 3).reshape(N, T, H * d)
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    out = sea.ops.sparse_attention_from_bits(bits, q.to(DEV), kk.to(DEV), v.to(DEV), scales.to(DEV), avg.to(DEV), P, k, True, True, kernel=kernel)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=2e-2, atol=1e-2)
    assert float((out.float().cpu() - ref).abs().mean()) < 2e-3


@pytest.mark.parametrize('tok,c0', [(0, 15.0), (0, -25.0), (0, 0.0), (0, 35.0), (1, 35.0), (1, -25.0), (2, 35.0), (2, -25.0)])
def test_block_attention_score_range(sea, tok, c0):
    """The block kernel subtracts min(q_t . k_1, q_t . k_2) instead of the row maximum.  One of the first tokens as an attention sink (~30
    and ~70 nats above the other scores: as the only probe token that would underflow every other alive probability; the exact range of the
    scheme is 110 nats above the reference within an 8-tile chunk), as an anti-sink (~50
    nats below them: every alive probability overflows the fixed offset and the kernel has to move the reference) and as a zero row,
    against the gather kernel's max-subtracted fp32 softmax (attention.py:1159-1165)."""
    N, H, T, P, k, d = 1, 2, 512, 64, 16, 64
    g = torch.Generator().manual_seed(17)
    mask = (torch.rand(N, H, T, P, generator=g) < 0.4).float()
    mask[..., 0] = (torch.rand(N, H, T, generator=g) < 0.5).float()        # the first pixel (tokens 0 ..) alive in about half of the rows
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    u = torch.nn.functional.normalize(torch.randn(d, generator=g), dim=0)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5 + 2.0 * u).bfloat16().to(DEV)
    kk = torch.randn(N, H, T, d, generator=g)
    kk[:, :, tok] = c0 * u
    kk = kk.bfloat16().to(DEV)
    v = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    scales = torch.randn(N, H, T, 2, generator=g).to(DEV)
    avg = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    ref = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True, kernel='gather')
    out = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True, kernel='block')
    assert torch.isfinite(out.float()).all()
    torch.testing.assert_close(out.float().cpu(), ref.float().cpu(), rtol=2e-2, atol=2e-2)


def test_attention_from_bits_strided_inputs(sea):
    """q / k / v as transposed views of [N,T,H,d] buffers (what a caller gets from .view(...).transpose(1, 2)): the TMA tensor
    maps and the gather kernels take the strides as they are."""
    N, H, T, P, k, d = 2, 4, 300, 32, 16, 64
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(N, H, T, P, generator=g) < 0.4).float()
    bits = sea.ops.mask_to_bits(mask.to(DEV))
    mk = lambda s: (torch.randn(N, T, H, d, generator=g) * s).bfloat16().to(DEV)
    qt, kt, vt = mk(d ** -0.5), mk(1.0), mk(1.0)
    q, kk, v = qt.transpose(1, 2), kt.transpose(1, 2), vt.transpose(1, 2)
    assert not q.is_contiguous()
    scales = torch.randn(N, H, T, 2, generator=g).to(DEV)
    avg = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    for kern in ('block', 'gather'):
        ref = sea.ops.sparse_attention_from_bits(bits, q.contiguous(), kk.contiguous(), v.contiguous(), scales, avg, P, k, True, True, kernel=kern)
        out = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True, kernel=kern)
        assert torch.equal(out, ref), kern
