"""Caller-side fusions: the decoder self-attention block around `PerlinAttention`, i.e. what
`OPTAttention.forward` + `PerlinSelfAttention.forward` do around the layer (reference: perlin_opt/perlin_opt.py:559-633,
434-477; perlin_attention/self_attention.py:75-264), with the copies and O(T^2) tensors of that glue removed:

  * q / k / v projections as ONE GEMM over the concatenated weight [3E, E] (cuBLAS, a plain library GEMM); the OPT query
    scaling d^-1/2 (perlin_opt.py:562) is folded into the packed q rows;
  * no `_shape(...).contiguous()` transposes (perlin_opt.py:573-600): q, k, v are strided VIEWS [N,H,T,d] of the GEMM output
    [N,T,3,H,d] -- every kernel behind PerlinAttention takes (sn, sh, st) strides, the TMA descriptors are built from them;
  * no `[N,1,T,T]` additive causal mask: `attention_mask=None` means "causal, nothing padded" (the reference materialises
    4 T^2 bytes per item and reads them back, attention.py:401-449); a mask, when given, is only inspected for padded rows;
  * the context comes out of the attention kernel already permuted to [N,T,H*d] (a14) and feeds `out_proj` directly;
  * decode: (k, v, state) travel in `past_key_value` like the reference (perlin_opt.py:575-581, 627-628).

Parameter names equal the reference's (`q_proj`, `k_proj`, `v_proj`, `out_proj`, `perlin_self_attention.attention.*`), so a
reference OPT checkpoint's attention block loads with `load_state_dict`.  LoRA adapters (self_attention.py:104-120) stay the
caller's: with `lora_in_approx_enabled` call PerlinAttention directly with a separate `v_for_atten`.
"""
from typing import Optional, Tuple

import torch
from torch import nn

from ._lib import SeaError
from .attention import PerlinAttention
from .config import PerlinAttentionConfig, get_default_config


class _SelfAttentionHolder(nn.Module):
    """Keeps PerlinAttention under the reference's name `perlin_self_attention.attention` (self_attention.py:52)."""

    def __init__(self, attention: PerlinAttention):
        super().__init__()
        self.attention = attention


class SeaOPTAttention(nn.Module):
    """Drop-in for the `attention_method == 'perlin'` path of the reference's OPTAttention (decoder self-attention)."""

    def __init__(self, embed_dim: int, num_heads: int, config=None, perlin_config: PerlinAttentionConfig = None, bias: bool = True):
        super().__init__()
        if embed_dim % num_heads:
            raise SeaError(f'embed_dim {embed_dim} is not divisible by num_heads {num_heads}')
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.scaling = self.head_dim ** -0.5
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        pc = perlin_config if perlin_config is not None else get_default_config()
        if not pc.causal:
            raise SeaError('SeaOPTAttention is the decoder (causal) block')
        if config is None:
            import types
            config = types.SimpleNamespace(hidden_size=embed_dim, num_attention_heads=num_heads, max_position_embeddings=2048)
        self.perlin_self_attention = _SelfAttentionHolder(PerlinAttention(config, pc))
        self.pconfig = pc
        self.benchmarking = False
        self.last_loss = None
        self._qkv = None              # (stamp, W [3E,E], b [3E] or None): packed projection, q rows pre-scaled

    @property
    def attention(self) -> PerlinAttention:
        return self.perlin_self_attention.attention

    # ------------------------------------------------------------------------------------------------
    def invalidate_packed(self):
        self._qkv = None
        self.attention.invalidate_packed()

    def _packed_qkv(self, dtype):
        """[q * d^-1/2 ; k ; v] weights as one [3E, E] matrix in the compute dtype.  Re-made on every call unless the attention
        module's packings are frozen (`attention.freeze_packed_weights()`), for the same reason as ops.PackedWeights."""
        src = [self.q_proj.weight, self.k_proj.weight, self.v_proj.weight]
        if self.q_proj.bias is not None:
            src += [self.q_proj.bias, self.k_proj.bias, self.v_proj.bias]
        stamp = tuple((int(t.data_ptr()), int(t._version)) for t in src) + (dtype,)
        hit = self._qkv
        if hit is not None and hit[0] == stamp and self.attention._packed.frozen:
            return hit[1], hit[2]
        with torch.no_grad():
            w = torch.cat([self.q_proj.weight * self.scaling, self.k_proj.weight, self.v_proj.weight], dim=0).to(dtype).contiguous()
            b = None
            if self.q_proj.bias is not None:
                b = torch.cat([self.q_proj.bias * self.scaling, self.k_proj.bias, self.v_proj.bias], dim=0).to(dtype).contiguous()
        self._qkv = (stamp, w, b)
        return w, b

    def _apply(self, fn, *args, **kwargs):
        self._qkv = None
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._qkv = None
        return super()._load_from_state_dict(*args, **kwargs)

    # ------------------------------------------------------------------------------------------------
    def forward(self, hidden_states: torch.Tensor, key_value_states: Optional[torch.Tensor] = None,
                past_key_value: Optional[Tuple[torch.Tensor]] = None, attention_mask: Optional[torch.Tensor] = None,
                layer_head_mask: Optional[torch.Tensor] = None, output_attentions: bool = False, use_cache: bool = False):
        """hidden_states [N, T_new, E] -> (attn_output [N, T_new, E], partial_attention_probs CSR or None, past_key_value).
        Same contract as perlin_opt.py:559-633 for decoder self-attention; cross-attention and head masks raise."""
        if key_value_states is not None or layer_head_mask is not None:
            raise SeaError('SeaOPTAttention: cross-attention / layer_head_mask are not part of the SEA decoder path (perlin_opt.py:601-602)')
        if not hidden_states.is_cuda:
            raise SeaError('SeaOPTAttention runs on CUDA tensors only; there is no CPU path')
        if self.training:
            raise SeaError('SeaOPTAttention is an inference path (the training branch is not implemented)')
        N, T_new, E = hidden_states.shape
        H, d = self.num_heads, self.head_dim
        att = self.attention
        w, b = self._packed_qkv(hidden_states.dtype)
        qkv = torch.nn.functional.linear(hidden_states, w, b).view(N, T_new, 3, H, d)        # one GEMM; q already scaled
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))                         # strided views [N,H,T_new,d], no copies
        cache = use_cache or self.pconfig.use_cache or past_key_value is not None
        state = None
        if past_key_value is not None:
            k = torch.cat([past_key_value[0], k], dim=2)
            v = torch.cat([past_key_value[1], v], dim=2)
            state = past_key_value[2] if len(past_key_value) > 2 else None
        elif cache:
            k, v = k.contiguous(), v.contiguous()          # the cache outlives this call: do not pin the whole qkv buffer
        att.output_attentions = bool(output_attentions)
        if cache:
            was = self.pconfig.use_cache
            self.pconfig.use_cache = True
            try:
                mask = attention_mask if state is None else None
                out = att(q, k, v, q, k, v, q, k, mask, None, None, state)
            finally:
                self.pconfig.use_cache = was
        else:
            out = att(q, k, v, q, k, v, q, k, attention_mask, None, None, None)
        self.last_loss = out.loss
        attn_output = torch.nn.functional.linear(out.context_layer, self.out_proj.weight.to(hidden_states.dtype),
                                                 None if self.out_proj.bias is None else self.out_proj.bias.to(hidden_states.dtype))
        present = (k, v, out.state) if cache else (k, v)
        return attn_output, (out.partial_attention_probs if output_attentions else None), present
