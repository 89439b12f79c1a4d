"""Out-of-bounds write detector for the kernels added in round 2, without a memory checker (compute-sanitizer is not available on the GPU
pool): every output of a C-ABI call is carved out of a larger buffer filled with a sentinel byte, the call runs on a ragged shape
(T not a multiple of the 128-row tile, H not dividing 128, partly filled slabs), and the guard bytes on both sides of every output must
come back untouched.  The results themselves are compared with the CPU oracle in tests/test_umma_gpu.py on the same kind of shapes."""
import math

import pytest
import torch
import transformers

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GUARD = 4096          # bytes on each side
SENT = 0x7B


class Guarded:
    """A tensor of `shape` / `dtype` in the middle of a sentinel-filled byte buffer."""

    def __init__(self, shape, dtype):
        n = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((GUARD + n + GUARD,), SENT, dtype=torch.uint8, device=DEV)
        self.t = self.raw[GUARD:GUARD + n].view(dtype).view(shape)
        self.n = n

    def check(self, name):
        torch.cuda.synchronize()
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.n:]
        assert bool((lo == SENT).all()), f'{name}: bytes BEFORE the output were written'
        assert bool((hi == SENT).all()), f'{name}: bytes AFTER the output were written'
        assert not bool((self.raw[GUARD:GUARD + self.n] == SENT).all()), f'{name}: the output was never written'


def _weights(sea, H, d, T, P):
    torch.manual_seed(H + d)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=8, k=8, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval().to(DEV)
    return mod, mod._weights_fp32()


@pytest.mark.parametrize('N,H,d,T,P,c_out', [(1, 12, 80, 37, 256, 24), (2, 32, 128, 131, 256, 64), (1, 20, 96, 19, 128, 40), (1, 12, 64, 45, 256, 64),
                                             (1, 3, 32, 70, 128, 8)])
def test_mlp_mma_writes_only_its_outputs(sea, N, H, d, T, P, c_out):
    lib, S, W = sea._lib, 2, P // 4
    mod, w = _weights(sea, H, d, T, P)
    ctx = torch.randn(N, H, T, 2 * d, device=DEV).bfloat16()
    v = torch.randn(N, H, T, d, device=DEV).bfloat16()
    cnn_in, scales = Guarded((N, T, W, c_out), torch.bfloat16), Guarded((N, H, T, 2), torch.float32)
    ws = Guarded((int(lib.load().sea_predictor_mlp_mma_workspace_bytes(d, S, W)),), torch.uint8)
    lib.call('sea_predictor_mlp_mma_fwd', ctx.data_ptr(), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
             w['enc_w'].data_ptr(), w['enc_b'].data_ptr(), w['enc_ln_w'].data_ptr(), w['enc_ln_b'].data_ptr(),
             w['dec_w'].data_ptr(), w['dec_b'].data_ptr(), w['cnn_ln_w'].data_ptr(), w['cnn_ln_b'].data_ptr(),
             w['scl_w'].data_ptr(), w['scl_b'].data_ptr(), cnn_in.t.data_ptr(), scales.t.data_ptr(), ws.t.data_ptr(),
             N, H, T, d, S, W, c_out, torch.cuda.current_stream().cuda_stream)
    cnn_in.check('cnn_in'); scales.check('scales'); ws.check('packed weights')
    assert torch.isfinite(cnn_in.t.float()).all() and torch.isfinite(scales.t).all()
    # and the same numbers as the allocating wrapper
    a_in, a_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W, force_mma=True, c_out=c_out)
    assert torch.equal(a_in, cnn_in.t) and torch.equal(a_sc, scales.t)


@pytest.mark.parametrize('N,H,d,T,nbf', [(1, 2, 80, 131, 8), (2, 3, 128, 257, 8), (1, 2, 96, 40, 8), (1, 3, 32, 129, 4), (1, 2, 64, 515, 8)])
def test_performer_slabs_write_only_their_outputs(sea, N, H, d, T, nbf):
    lib = sea._lib
    F = int(d * math.log(d) / nbf)
    assert lib.load().sea_performer_mma_supported(lib.SEA_DTYPE_BF16, d, F)
    g = torch.Generator().manual_seed(T)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    k = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    pos = torch.randn(T + 3, d, generator=g).to(DEV)
    proj = torch.randn(F, d, generator=g).to(DEV)
    ctx, avg = Guarded((N, H, T, 2 * d), torch.bfloat16), Guarded((N, H, T, d), torch.bfloat16)
    ws = Guarded((int(lib.load().sea_performer_mma_workspace_floats(N, H, T, d, F)),), torch.float32)
    lib.call('sea_performer_causal_mma_fwd', q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
             v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), pos.data_ptr(), proj.data_ptr(), ctx.t.data_ptr(), avg.t.data_ptr(), ws.t.data_ptr(),
             N, H, T, d, F, torch.cuda.current_stream().cuda_stream)
    ctx.check('ctx'); avg.check('cumavg'); ws.check('chunk-sum workspace')
    assert torch.isfinite(ctx.t.float()).all() and torch.isfinite(avg.t.float()).all()
    c2, a2 = sea.ops.performer_causal(q, k, v, pos, proj)
    assert torch.equal(c2, ctx.t) and torch.equal(a2, avg.t)


@pytest.mark.parametrize('H,d,T,P,k', [(3, 80, 150, 128, 16), (2, 96, 70, 128, 8), (2, 128, 140, 256, 32)])
def test_gather_attention_writes_only_its_output(sea, H, d, T, P, k):
    import numpy as np
    from oracle import sea_oracle as so
    lib, N = sea._lib, 1
    g = torch.Generator().manual_seed(d)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    probs = torch.softmax(torch.randn(N, H, T, P, generator=g) * 2, -1).to(DEV)
    kpr = torch.from_numpy(np.tile(so.per_item_top_k_causal(H, k, 1.0, P, T), N)).to(DEV)
    bits = sea.ops.topk_mask_bits(probs, kpr, 'causal_batch')
    scales = torch.randn(N, H, T, 2, generator=g).to(DEV)
    avg = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    out = Guarded((N, T, H * d), torch.bfloat16)
    lib.call('sea_sparse_attention_bits_fwd', bits.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), kk.data_ptr(), kk.stride(0), kk.stride(1),
             kk.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), scales.data_ptr(), avg.data_ptr(), T * d, d, 1, lib.SEA_DTYPE_BF16, out.t.data_ptr(),
             N, H, T, T, d, P, k, 1, torch.cuda.current_stream().cuda_stream)
    out.check('context')
    ref = sea.ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True, kernel='gather')
    assert torch.equal(ref, out.t)
