"""Caller-side fusions: SeaOPTAttention = q/k/v projections as one GEMM + strided q/k/v views + no [N,1,T,T] mask
+ PerlinAttention + out_proj, against (a) the CPU oracle fed with separately projected q, k, v (what perlin_opt.py:559-633 does)
and (b) the unfused composition through the repo's own PerlinAttention."""
import pytest
import torch
import transformers

from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _block(sea, E, H, T, P, k, nbf, seed=0):
    torch.manual_seed(seed)
    cfg = transformers.BertConfig(hidden_size=E, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)
    return sea.SeaOPTAttention(E, H, cfg, pc).eval()


def _separate_qkv(blk, x):
    """perlin_opt.py:562-600: three Linear calls, the query scaling, `_shape` transposes."""
    N, T, E = x.shape
    H, d = blk.num_heads, blk.head_dim
    shp = lambda t: t.view(N, T, H, d).transpose(1, 2).contiguous()
    q = shp(torch.nn.functional.linear(x, blk.q_proj.weight, blk.q_proj.bias) * blk.scaling)
    k = shp(torch.nn.functional.linear(x, blk.k_proj.weight, blk.k_proj.bias))
    v = shp(torch.nn.functional.linear(x, blk.v_proj.weight, blk.v_proj.bias))
    return q, k, v


@pytest.mark.parametrize('E,H,T,P,k,nbf', [(256, 4, 128, 32, 8, 8), (192, 3, 100, 16, 6, 8)])
def test_block_matches_oracle_fp32(sea, E, H, T, P, k, nbf):
    blk = _block(sea, E, H, T, P, k, nbf)
    sd = {k_[len('perlin_self_attention.attention.'):]: v_.detach().clone().float() for k_, v_ in blk.state_dict().items()
          if k_.startswith('perlin_self_attention.attention.')}
    x = torch.randn(2, T, E, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        q, kk, v = _separate_qkv(blk, x)
        ref = so.perlin_forward_causal(sd, q, kk, v, k_top=k, P=P, sparse=True)
        ref_out = torch.nn.functional.linear(ref['context_layer'], blk.out_proj.weight, blk.out_proj.bias)
        blk = blk.to(DEV)
        out, probs, present = blk(x.to(DEV))                               # attention_mask=None: causal, nothing padded
        mask = so.causal_additive_mask(T, torch.float32, 2).to(DEV)
        out_m, _, _ = blk(x.to(DEV), attention_mask=mask)                  # an explicit mask gives the same result
    assert probs is None and len(present) == 2 and present[0].shape == (2, H, T, E // H)
    assert torch.equal(out, out_m)
    # rows whose top-k mask equals the oracle's (near-ties can move with the GEMM's summation order) agree to fp32 tolerance
    close = ((out.cpu() - ref_out).abs() <= 1e-4 + 1e-3 * ref_out.abs()).all(dim=-1)
    assert close.float().mean() > 0.9, float(close.float().mean())


def test_block_equals_unfused_composition_bf16(sea):
    E, H, T, P, k, nbf = 512, 8, 256, 64, 16, 8
    blk = _block(sea, E, H, T, P, k, nbf, seed=1).to(DEV).bfloat16()
    x = torch.randn(1, T, E, generator=torch.Generator().manual_seed(5)).bfloat16().to(DEV)
    with torch.no_grad():
        out, _, _ = blk(x)
        q, kk, v = _separate_qkv(blk, x)
        ctx = blk.attention(q, kk, v, q, kk, v, q, kk, None, None, None).context_layer
        ref = torch.nn.functional.linear(ctx, blk.out_proj.weight, blk.out_proj.bias)
    # bf16: the fused [3E,E] GEMM pre-scales the q weights instead of scaling the product -- one bf16 rounding apart
    d = (out.float() - ref.float()).abs()
    assert float((d > 3e-2 + 3e-2 * ref.float().abs()).float().mean()) < 0.02 and float(d.mean()) < 5e-3


def test_block_decode_with_cache_matches_prefill(sea):
    """use_cache: (k, v, state) in past_key_value (perlin_opt.py:575-581, 627-628); rows of a token-by-token decode equal the
    rows of the full prefill (the property test_perlin_opt_cache.py checks)."""
    E, H, T, P, k, nbf = 256, 4, 48, 32, 8, 8
    blk = _block(sea, E, H, T, P, k, nbf, seed=2).to(DEV)
    x = torch.randn(1, T, E, generator=torch.Generator().manual_seed(7)).to(DEV)
    T0 = 40
    with torch.no_grad():
        full, _, _ = blk(x)
        out0, _, past = blk(x[:, :T0], use_cache=True)
        assert len(past) == 3 and past[2] is not None and past[2].t == T0
        rows = [out0]
        for t in range(T0, T):
            o, _, past = blk(x[:, t:t + 1], past_key_value=past, use_cache=True)
            rows.append(o)
        assert past[0].shape[2] == T and past[2].t == T
    dec = torch.cat(rows, dim=1)
    close = ((dec - full).abs() <= 1e-4 + 1e-3 * full.abs()).all(dim=-1)
    assert close.float().mean() > 0.9, float(close.float().mean())


def test_block_rejects_what_it_does_not_do(sea):
    blk = _block(sea, 128, 2, 16, 8, 4, 8).to(DEV)
    x = torch.zeros(1, 16, 128, device=DEV)
    with pytest.raises(sea.SeaError):
        blk(x, key_value_states=x)
    with pytest.raises(sea.SeaError):
        blk(x.cpu())
    sd = blk.state_dict()
    assert 'q_proj.weight' in sd and 'out_proj.bias' in sd and 'perlin_self_attention.attention.performer.projection_matrix' in sd
