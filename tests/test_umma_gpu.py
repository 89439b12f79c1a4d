"""GPU: the tcgen05/TMA implicit-GEMM kernels (bf16) against the fp32 SIMT kernels / the oracle."""
import pytest
import torch

from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _conv_ref(x_cl, weight, bias):
    """x_cl [N,T,W,C] float -> oracle conv (NCHW) -> channels-last"""
    x = x_cl.permute(0, 3, 1, 2).float()
    wm = torch.zeros_like(weight)
    wm[:, :, :3, :] = 1
    y = torch.relu(so.causal_conv2d(x, weight, wm, bias, 3, 2, 2))
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('N,T,W', [(1, 8, 64), (2, 37, 64), (1, 130, 32), (1, 64, 16), (1, 5, 128), (3, 700, 64), (1, 601, 64)])
def test_conv3x3_umma_matches_oracle(sea, N, T, W):
    C = O = 64
    g = torch.Generator().manual_seed(T * 7 + W)
    x = torch.randn(N, T, W, C, generator=g).bfloat16()
    weight = torch.zeros(O, C, 5, 3)
    weight[:, :, :3, :] = torch.randn(O, C, 3, 3, generator=g) * 0.05
    bias = torch.randn(O, generator=g) * 0.1
    assert sea.ops.conv_umma_supported(torch.bfloat16, W, C, O)
    y = sea.ops.causal_conv3x3_dil2_relu(x.to(DEV), weight.to(DEV), bias.to(DEV))
    torch.cuda.synchronize()
    # reference on the bf16-rounded operands (the kernel multiplies bf16 x bf16 exactly, accumulates in fp32)
    ref = _conv_ref(x.float(), weight.bfloat16().float(), bias)
    torch.testing.assert_close(y.float().cpu(), ref, rtol=2e-2, atol=2e-2)
    # and against the fp32 SIMT kernel fed the same bf16 input
    y2 = sea.ops.causal_conv3x3_dil2_relu(x.to(DEV), weight.to(DEV), bias.to(DEV), force_simt=True)
    torch.testing.assert_close(y.float().cpu(), y2.float().cpu(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('N,T,W', [(1, 8, 64), (2, 37, 64), (1, 130, 32)])
def test_conv1x1_umma_matches_oracle(sea, N, T, W):
    C, O = 64, 32
    g = torch.Generator().manual_seed(T + W)
    x = torch.randn(N, T, W, C, generator=g).bfloat16()
    weight = torch.randn(O, C, generator=g) * 0.1
    bias = torch.randn(O, generator=g)
    y = sea.ops.conv1x1_umma(x.to(DEV), weight.to(DEV), bias.to(DEV))
    ref = x.float() @ weight.bfloat16().float().t() + bias
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('N,T', [(1, 8), (2, 37), (3, 700), (1, 601), (1, 2)])
def test_conv3x3_conv1x1_fused_matches_separate(sea, N, T):
    """the second 3x3 conv with the 1x1 conv fused behind it == the two tcgen05 kernels run one after the other (same bf16
    rounding of the activation between them), and == the oracle within the bf16 tolerance"""
    W = C = 64
    g = torch.Generator().manual_seed(T * 3 + N)
    x = torch.randn(N, T, W, C, generator=g).bfloat16().to(DEV)
    w2 = torch.zeros(64, C, 5, 3)
    w2[:, :, :3, :] = torch.randn(64, C, 3, 3, generator=g) * 0.05
    b2 = torch.randn(64, generator=g) * 0.1
    w3 = torch.randn(32, 64, generator=g) * 0.1
    b3 = torch.randn(32, generator=g)
    assert sea.ops.conv3x3_conv1x1_supported(torch.bfloat16, W, C, 64, 32)
    y3 = sea.ops.causal_conv3x3_relu_conv1x1(x, w2.to(DEV), b2.to(DEV), w3.to(DEV), b3.to(DEV))
    y = sea.ops.causal_conv3x3_dil2_relu(x, w2.to(DEV), b2.to(DEV))
    y3_sep = sea.ops.conv1x1_umma(y, w3.to(DEV), b3.to(DEV))
    torch.cuda.synchronize()
    torch.testing.assert_close(y3.cpu(), y3_sep.cpu(), rtol=1e-5, atol=1e-5)
    ref = _conv_ref(x.float().cpu(), w2.bfloat16().float(), b2).bfloat16().float() @ w3.bfloat16().float().t() + b3
    torch.testing.assert_close(y3.cpu(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('ties', [False, True])
@pytest.mark.parametrize('N,H,T,W,P,k', [(1, 32, 40, 64, 256, 64), (2, 8, 33, 16, 64, 8), (1, 4, 20, 8, 32, 4), (1, 32, 12, 32, 128, 16),
                                        (1, 32, 300, 64, 256, 8), (1, 16, 90, 64, 256, 16), (1, 12, 50, 64, 256, 64),
                                        (1, 16, 37, 128, 512, 32), (1, 8, 21, 256, 1024, 16)])      # 16 / 32 pixels per lane: 4 / 8 conv outputs under a lane
def test_tail_topk_fused_matches_oracle(sea, N, H, T, W, P, k, ties):
    import numpy as np
    g = torch.Generator().manual_seed(P + T)
    y3 = torch.randn(N, T, W, H, generator=g)
    bias = torch.randn(H, generator=g)
    ln_w = 1 + 0.1 * torch.randn(P, generator=g)
    ln_b = 0.1 * torch.randn(P, generator=g)
    if ties:        # freshly initialised LayerNorm: the x(P/W) upsample makes runs of exactly equal probabilities
        ln_w, ln_b = torch.ones(P), torch.zeros(P)
    kpr = torch.from_numpy(np.tile(so.per_item_top_k_causal(H, k, 1.0, P, T), N))
    probs, bits = sea.ops.predictor_tail_topk(y3.to(DEV), bias.to(DEV), ln_w.to(DEV), ln_b.to(DEV), kpr.to(DEV), P)
    # oracle: [N,H,T,W] -> nearest x(P/W) -> bias pad columns -> area resize -> LN -> softmax
    y = y3.permute(0, 3, 1, 2)
    u = y.repeat_interleave(P // W, dim=-1)
    pad = bias.view(1, H, 1, 1).expand(N, H, T, 1)
    u = so.area_resize_width(torch.cat([pad, u, pad], dim=-1), P)
    ref = torch.softmax(so.layer_norm(u, ln_w, ln_b), dim=-1)
    torch.testing.assert_close(probs.cpu(), ref, rtol=1e-4, atol=1e-7)
    # top-k bit-exact given the kernel's own probabilities
    mask_ref = so.topk_mask_causal_batch(probs.cpu(), k)
    assert torch.equal(sea.ops.bits_to_mask(bits, H, P).cpu(), mask_ref)
    # and equal to the standalone top-k kernel
    bits2 = sea.ops.topk_mask_bits(probs, kpr.to(DEV), 'causal_batch')
    assert torch.equal(bits.cpu(), bits2.cpu())


@pytest.mark.parametrize('N,H,T,P', [(1, 32, 64, 256), (2, 32, 37, 256), (1, 16, 50, 128), (1, 64, 9, 64), (1, 4, 70, 256),
                                     (1, 12, 45, 256), (2, 12, 33, 64), (1, 20, 19, 128)])        # H need not divide 128
def test_mlp_umma_matches_simt(sea, N, H, T, P):
    import transformers
    d, S, W = 64, 2, P // 4
    torch.manual_seed(P + H)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=8, k=8, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval().to(DEV)
    for n_, p_ in mod.named_parameters():
        if p_.ndim == 1:
            p_.data.add_(0.1 * torch.randn_like(p_))
    w = mod._weights_fp32()
    ctx = torch.randn(N, H, T, 2 * d, device=DEV).bfloat16()
    v = torch.randn(N, H, T, d, device=DEV).bfloat16()
    assert sea._lib.load().sea_predictor_mlp_umma_supported(1, H, d, S, W)
    a_in, a_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W)
    b_in, b_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W, force_simt=True)
    torch.cuda.synchronize()
    # bf16 operands (weights and the GELU output are rounded to bf16 before each GEMM): the north star's bf16 tolerance, 2e-2,
    # against the repo's own SIMT kernel ...
    torch.testing.assert_close(a_sc.cpu(), b_sc.cpu(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(a_in.float().cpu(), b_in.float().cpu(), rtol=2e-2, atol=2e-2)
    # ... and directly against the CPU oracle (attention.py:190-196,242-245,289-291 + the CNN's first LayerNorm, :268)
    sd = {k_: v_.detach().float().cpu() for k_, v_ in mod.state_dict().items()}
    t_ref = so.predictor_enc(torch.cat([ctx.float().cpu(), v.float().cpu()], -1), sd)
    x0 = so.layer_norm(so.predictor_dec_row(t_ref, sd, S), sd['attention_predictor_cnn.0.module.weight'], sd['attention_predictor_cnn.0.module.bias'])
    torch.testing.assert_close(a_sc.cpu(), so.predictor_dec_scaler(t_ref, sd), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(a_in.float().cpu().permute(0, 3, 1, 2), x0, rtol=2e-2, atol=2e-2)
    if 2 * H < 64:      # zero-padded channels for the 64-channel tcgen05 convolutions (OPT-125m: H = 12)
        p_in, p_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W, c_out=64)
        assert p_in.shape == (N, T, W, 64)
        assert torch.equal(p_in[..., :2 * H], a_in) and torch.equal(p_sc, a_sc)
        assert float(p_in[..., 2 * H:].float().abs().max()) == 0.0


@pytest.mark.parametrize('N,H,T,nbf,d', [(1, 2, 128, 8, 64), (2, 3, 200, 8, 64), (1, 4, 515, 8, 64), (1, 2, 96, 5, 64), (1, 1, 300, 16, 64), (1, 1, 40, 32, 64),
                                         # other head dims run the same kernels once per 128-column slab of [pos | v]
                                         (1, 2, 300, 8, 128), (2, 2, 131, 8, 80), (1, 3, 260, 8, 32), (1, 1, 140, 4, 32), (1, 2, 200, 8, 96),
                                         (1, 1, 257, 10, 128), (1, 2, 129, 6, 80)])
def test_performer_mma_matches_oracle(sea, N, H, T, nbf, d):
    import math
    F = int(d * math.log(d) / nbf)
    g = torch.Generator().manual_seed(T + nbf)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16()
    k = torch.randn(N, H, T, d, generator=g).bfloat16()
    v = torch.randn(N, H, T, d, generator=g).bfloat16()
    pos = torch.randn(T + 3, d, generator=g)
    proj = torch.randn(F, d, generator=g)
    assert sea._lib.load().sea_performer_mma_supported(1, d, F)
    ctx, avg = sea.ops.performer_causal(q.to(DEV), k.to(DEV), v.to(DEV), pos.to(DEV), proj.to(DEV))
    torch.cuda.synchronize()
    v2 = torch.cat([pos[:T].view(1, 1, T, d).expand(N, H, T, d), v.float()], -1)
    ref = so.performer_causal(q.float(), k.float(), v2, proj)
    torch.testing.assert_close(ctx.float().cpu(), ref, rtol=2e-2, atol=2e-2)
    avg_ref = v.float().cumsum(-2) / torch.arange(1, T + 1).view(1, 1, T, 1)
    torch.testing.assert_close(avg.float().cpu(), avg_ref, rtol=2e-2, atol=1e-2)
    # cross-check against the fp32 SIMT kernel on the same bf16 inputs
    ctx2, avg2 = sea.ops.performer_causal(q.to(DEV), k.to(DEV), v.to(DEV), pos.to(DEV), proj.to(DEV), force_simt=True)
    torch.testing.assert_close(ctx.float().cpu(), ctx2.float().cpu(), rtol=2e-2, atol=2e-2)


def test_tail_topk_fused_row_counts(sea):
    """The fused tail also emits pass 1 of the CSR interpolation (per-row entry counts)."""
    import numpy as np
    N, H, T, W, P, k = 2, 8, 70, 16, 64, 8
    g = torch.Generator().manual_seed(1)
    y3 = torch.randn(N, T, W, H, generator=g)
    bias = torch.randn(H, generator=g)
    ln_w, ln_b = torch.ones(P), torch.zeros(P)
    kpr = torch.from_numpy(np.tile(so.per_item_top_k_causal(H, k, 1.0, P, T), N))
    probs, bits, crow = sea.ops.predictor_tail_topk(y3.to(DEV), bias.to(DEV), ln_w.to(DEV), ln_b.to(DEV), kpr.to(DEV), P, count_k=k)
    crow2, col2, Z2 = sea.ops.csr_from_bits(bits, H, P, k, T, True, torch.int32)          # unfused count
    crow3, col3, Z3, hp = sea.ops.csr_from_bits(bits, H, P, k, T, True, torch.int32, z_alloc=Z2, want_head_ptr=True, crow_counts=crow)
    assert torch.equal(crow3.cpu(), crow2.cpu()) and torch.equal(col3.cpu(), col2.cpu())
    mask = sea.ops.bits_to_mask(bits, H, P).cpu()
    crow_r, col_r, Z_r = so.resize_from_m_to_t_csr(mask, k, T, True)
    assert torch.equal(crow3.cpu().long(), crow_r) and torch.equal(col3.cpu().long(), col_r)


@pytest.mark.parametrize('N,H,d,T,P', [(1, 32, 128, 40, 256), (2, 32, 80, 37, 256), (1, 12, 80, 45, 256), (1, 16, 128, 50, 128),
                                       (1, 20, 96, 19, 128), (1, 32, 64, 33, 256), (1, 4, 32, 70, 256)])
def test_mlp_mma_any_head_dim_matches_oracle(sea, N, H, d, T, P):
    """csrc/mlp_mma.cu (warp-level tensor-core MLP for any head dim: OPT-2.7B d = 80, the long-context sweep d = 128) against the CPU
    oracle (attention.py:190-196,242-245,289-291 + the CNN's first LayerNorm, :268) and the repo's fp32 SIMT kernel."""
    import transformers
    S, W = 2, P // 4
    torch.manual_seed(P + H + d)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=8, k=8, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval().to(DEV)
    for n_, p_ in mod.named_parameters():
        if p_.ndim == 1:
            p_.data.add_(0.1 * torch.randn_like(p_))
    w = mod._weights_fp32()
    ctx = torch.randn(N, H, T, 2 * d, device=DEV).bfloat16()
    v = torch.randn(N, H, T, d, device=DEV).bfloat16()
    assert sea._lib.load().sea_predictor_mlp_mma_supported(1, H, d, S, W)
    a_in, a_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W, force_mma=True)
    b_in, b_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W, force_simt=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(a_sc.cpu(), b_sc.cpu(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(a_in.float().cpu(), b_in.float().cpu(), rtol=2e-2, atol=3e-2)
    sd = {k_: v_.detach().float().cpu() for k_, v_ in mod.state_dict().items()}
    t_ref = so.predictor_enc(torch.cat([ctx.float().cpu(), v.float().cpu()], -1), sd)
    x0 = so.layer_norm(so.predictor_dec_row(t_ref, sd, S), sd['attention_predictor_cnn.0.module.weight'], sd['attention_predictor_cnn.0.module.bias'])
    torch.testing.assert_close(a_sc.cpu(), so.predictor_dec_scaler(t_ref, sd), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(a_in.float().cpu().permute(0, 3, 1, 2), x0, rtol=2e-2, atol=3e-2)
    if 2 * H < 64:      # zero-padded channels for the 64-channel tcgen05 convolutions
        p_in, p_sc, _ = sea.ops.predictor_mlp(ctx, v, w, S, W, c_out=64, force_mma=True)
        assert p_in.shape == (N, T, W, 64)
        assert torch.equal(p_in[..., :2 * H], a_in) and torch.equal(p_sc, a_sc)
        assert float(p_in[..., 2 * H:].float().abs().max()) == 0.0
