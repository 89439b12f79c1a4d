"""Drop-in `PerlinAttention` (reference: src/models/perlin_attention/attention.py:133-1359).

Same constructor `(config, perlin_config)`, same parameter / buffer names (so reference checkpoints
load with `load_state_dict`), same 12-argument `forward` and the same `PerlinAttentionOutput` tuple --
but the forward is a fixed sequence of hand-written sm_100a kernels reached through the C ABI of
libsea_b200.so (see include/sea_b200.h, ops.py).  The nn.Modules below only HOLD parameters; their
torch forward is never used, and there is no CPU / eager fallback: unsupported modes raise.

What runs on the device per call (causal prefill, the reference's `benchmarking` branch, attention.py
:518-573, 595-673, 774-947, 1036-1042, 1151-1173, 1208-1282):
  performer (3 launches) -> predictor MLP -> conv x2 -> predictor tail (+softmax) -> grouped top-k ->
  CSR count+scan -> CSR fill -> fused sparse attention (+scaler, +running-mean mix, +permute)
with no host synchronisation inside the chain (the reference has >= 8 `.item()`/`nonzero()` syncs).  The one host
read a default call does is the padding check (`check_padding`, N*T booleans); set `check_padding = False` for sync-free
calls / CUDA-graph capture.  `output_attentions` adds one nnz read-back (the reference's own `.item()`).
"""
import math
import os
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import ops
from .attention_state import PerlinAttentionState
from ._lib import SeaError
from .config import PerlinAttentionConfig, get_default_config


class PerlinAttentionOutput(NamedTuple):
    """Field-for-field the reference's output tuple (attention.py:84-106)."""
    loss: torch.Tensor
    context_layer: torch.Tensor
    partial_attention_probs: torch.Tensor
    partial_attention_mask: torch.Tensor
    estimated_attention_probs_m: torch.Tensor
    estimated_attention_probs: torch.Tensor
    dense_attention_probs: torch.Tensor
    key_for_score: torch.Tensor
    state: object

    def to(self, device):
        mv = lambda t: t.to(device) if isinstance(t, torch.Tensor) else t
        return PerlinAttentionOutput(*[mv(f) for f in self])


# ----------------------------------------------------------------------------- parameter containers
class _Holder(nn.Module):
    """Keeps a child under a fixed attribute name so state_dict keys equal the reference's
    (`ModuleBenchmark.module`, attention.py:108-121; `KeepRes.net`, modules.py:42-55)."""

    def __init__(self, attr: str, child: nn.Module):
        super().__init__()
        setattr(self, attr, child)

    def forward(self, *a, **kw):
        raise SeaError('parameter container: the CUDA path never calls torch forwards')


class _CausalConvParams(nn.Module):
    """Parameters of the reference's CausalConv2d (modules.py:96-142): weight [O,C,2k-1,k] whose first k
    rows are the live taps, a 0/1 `weight_mask` buffer, bias [O]; PyTorch Conv2d default init."""

    def __init__(self, in_ch: int, out_ch: int, ksize: int):
        super().__init__()
        seed_conv = nn.Conv2d(in_ch, out_ch, ksize)
        w = torch.zeros((out_ch, in_ch, 2 * ksize - 1, ksize))
        w[:, :, :ksize, :] = seed_conv.weight.data
        m = torch.zeros_like(w)
        m[:, :, :ksize, :] = 1.0
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(seed_conv.bias.data.clone())
        self.register_buffer('weight_mask', m)


def _orthogonal_chunk(cols):
    qm, _ = torch.linalg.qr(torch.randn((cols, cols)), mode='reduced')
    return qm.t()


def gaussian_orthogonal_random_matrix(nb_rows: int, nb_columns: int) -> torch.Tensor:
    """Performer projection (FAVOR+, Choromanski et al.): stacked orthogonal blocks rescaled by the
    norms of fresh Gaussian rows (performer-pytorch `ortho_scaling=0`)."""
    blocks = [_orthogonal_chunk(nb_columns) for _ in range(nb_rows // nb_columns)]
    rem = nb_rows - (nb_rows // nb_columns) * nb_columns
    if rem > 0:
        blocks.append(_orthogonal_chunk(nb_columns)[:rem])
    mat = torch.cat(blocks)
    mult = torch.randn((nb_rows, nb_columns)).norm(dim=1)
    return torch.diag(mult) @ mat


class _PerformerParams(nn.Module):
    """Holds `projection_matrix` under the name the reference's `FastAttention` registers it
    (attention.py:159-164) and offers `redraw_projection_matrix` for `ProjectionUpdater`."""

    def __init__(self, dim_heads: int, nb_features: int, causal: bool):
        super().__init__()
        self.dim_heads, self.nb_features, self.causal = dim_heads, nb_features, causal
        self.generalized_attention = causal
        self.register_buffer('projection_matrix', gaussian_orthogonal_random_matrix(nb_features, dim_heads))

    @torch.no_grad()
    def redraw_projection_matrix(self, device=None):
        self.projection_matrix.copy_(gaussian_orthogonal_random_matrix(self.nb_features, self.dim_heads))


class ProjectionUpdater(nn.Module):
    """Mirror of src/models/common/performer.py:5-36 (training-time feature redraw)."""

    def __init__(self, instance, feature_redraw_interval):
        super().__init__()
        self.instance = instance
        self.feature_redraw_interval = feature_redraw_interval
        self.register_buffer('calls_since_last_redraw', torch.tensor(0))

    def fix_projections_(self):
        self.feature_redraw_interval = None

    def redraw_projections(self, device=None):
        if not self.training:
            return
        if self.feature_redraw_interval is not None and self.calls_since_last_redraw >= self.feature_redraw_interval:
            for m in self.instance.modules():
                if isinstance(m, _PerformerParams):
                    m.redraw_projection_matrix(device)
            self.calls_since_last_redraw.zero_()
            return
        self.calls_since_last_redraw += 1


def _k_per_row_causal(H: int, k: float, k_oversample: float, P: int, T_SRC: int, T_DST: int, device) -> torch.Tensor:
    """per_item_top_k of attention.py:849,856,866, evaluated with the same torch fp32 ops (round half
    even, clamp >= 1).  Depends on shapes only, so it is cached by the module."""
    tl_ = torch.arange(T_SRC - T_DST + 1, T_SRC + 1, dtype=torch.long, device=device)
    kt = H * ((k * k_oversample * P) / tl_)
    return torch.clamp_min(torch.round(kt), 1).float().contiguous()


def _csr_alloc_upper_bound(H: int, k: int, P: int, T_SRC: int, T_DST: int, k_per_row: torch.Tensor) -> int:
    """Upper bound of the nnz of one batch item: a row keeps at most K_t pixels, each at most
    min(ceil(L/P), k) wide, and never more than H*L entries (L = causal length)."""
    L = torch.arange(T_SRC - T_DST + 1, T_SRC + 1, dtype=torch.float64)
    width = torch.clamp(torch.ceil(L / P) + 1, max=float(k))   # +1: slack for fp32 rounding of the pixel bounds
    per_row = torch.minimum(k_per_row.double().cpu().clamp(max=float(H * P)) * width, H * L)
    return int(per_row.sum().item()) + 32


def _csr_outputs(crow, col, pvals, size):
    """partial_attention_mask / partial_attention_probs exactly as the reference returns them (causal_resize_m_to_t.py:757-762):
    batched CSR with int64 indices and Z = max_n nnz_n columns (rows of items with fewer entries are zero-padded at the tail).
    The kernels work on int32 indices in a buffer allocated at a shape-derived upper bound; `output_attentions` is off the hot
    path, so the nnz is read back once here (the reference's own `.item()`, :667) and the buffers are trimmed."""
    Z = int(crow[:, -1].max().item())
    crow64, col64 = crow.to(torch.int64), col[:, :Z].to(torch.int64).contiguous()
    ones = torch.ones((col64.shape[0], Z), dtype=torch.float32, device=col.device)
    vals = pvals[:, :Z].contiguous()
    if col64.shape[0] > 1:       # the tail past an item's own nnz must be zero (it is uninitialised in the over-allocated buffer)
        dead = torch.arange(Z, device=col.device).view(1, Z) >= crow64[:, -1:].to(col.device)
        col64.masked_fill_(dead, 0)
    return torch.sparse_csr_tensor(crow64, col64, ones, size=size), torch.sparse_csr_tensor(crow64, col64, vals, size=size)


class PerlinAttention(nn.Module):
    def __init__(self, config, perlin_config: PerlinAttentionConfig = None):
        super().__init__()
        self.config = config
        self.pconfig = perlin_config if perlin_config is not None else get_default_config()
        pc = self.pconfig
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        H, d, P = self.num_attention_heads, self.attention_head_size, pc.attention_predictor_length

        # set from outside by the reference harnesses (benchmark_bert.py:172-173)
        self.benchmarking = False
        # analogue of HF `output_attentions`: materialise partial_attention_probs / _mask CSR tensors
        self.output_attentions = False
        # one tiny host read of N*T booleans to reject padded batches (set False under CUDA graphs)
        self.check_padding = True
        # estimated_attention_probs [N,H,T,P] fp32 is an OUTPUT of the reference (attention.py:1349-1359), but its OPT caller drops it
        # (perlin_opt.py:477) unless output_attentions; False = the fused tail keeps the probabilities in registers (top-k only) and
        # the two estimated_* fields of the output tuple are None (134 MB less HBM traffic per layer at the north-star shape)
        self.keep_estimated_probs = True
        # decode (use_cache): one native call per token (sea_decode_step) instead of the per-op python sequence; False keeps the latter
        # (same kernels, same results; it is what the tests compare the native step with)
        self.decode_native = True
        self._decode_ws = {}

        self.performer_nb_features = int(d * math.log(d) / pc.performer_nb_factor)
        self.performer = _PerformerParams(d, self.performer_nb_features, causal=pc.causal)
        self.performer_proj_updater = ProjectionUpdater(self.performer, 1000)
        self.register_buffer('attention_predictor_enc_head_embd', torch.eye(H))
        self.attention_predictor_enc_per_layer = nn.Sequential(
            nn.Linear(3 * d * H, 2 * d * H), nn.LayerNorm(2 * d * H), nn.GELU())
        self.attention_predictor_enc = nn.Sequential(nn.Linear(3 * d, 2 * d), nn.LayerNorm(2 * d), nn.GELU())
        if not pc.causal:
            self.attention_predictor_dec_row_down_scale = 2
            self.attention_predictor_dec_row_splits = 4
            S = 4
            self.attention_predictor_dec_row_out_ch = (P // 2) * S
            self.attention_predictor_dec_row = nn.Sequential(nn.Linear(2 * d, self.attention_predictor_dec_row_out_ch), nn.Identity())
            self.attention_predictor_cnn = nn.Sequential(_Holder('net', nn.Sequential(
                nn.Conv2d(S * H, 4 * H, 3, padding=1, stride=(2, 1)), nn.ReLU(),
                nn.Conv2d(4 * H, 4 * H, 3, padding=1), nn.ReLU(),
                nn.Identity(),
                nn.Conv2d(4 * H, H, 3, padding=1))))
        else:
            inner_ch = int(os.environ.get('PERLIN_HOTFIX_OPT_INNER_CH', '2'))
            # PERLIN_HOTFIX_OPT_DEEPER=1 (attention.py:246-263): a third dilated causal conv + ReLU in the predictor CNN
            self._deeper = int(os.environ.get('PERLIN_HOTFIX_OPT_DEEPER', '0')) == 1
            self.attention_predictor_dec_row_down_scale = 4
            self.attention_predictor_dec_row_splits = inner_ch
            self.attention_predictor_dec_row_out_ch = (P // 4) * inner_ch
            self.attention_predictor_dec_row = nn.Sequential(nn.Linear(2 * d, self.attention_predictor_dec_row_out_ch), nn.Identity())
            C = inner_ch * H
            convs = []
            for _ in range(3 if self._deeper else 2):
                convs += [_Holder('module', _CausalConvParams(C, C, 3)), nn.ReLU()]
            self.attention_predictor_cnn = nn.Sequential(
                _Holder('module', nn.LayerNorm(P // 4)),
                _Holder('module', _Holder('net', nn.Sequential(
                    *convs,
                    _Holder('module', nn.Identity()),
                    _Holder('module', _CausalConvParams(C, H, 1))))),
                _Holder('module', nn.LayerNorm(P)))
        self.attention_predictor_dec_scaler = nn.Sequential(nn.Linear(2 * d, 2))
        self.norm_performer = nn.LayerNorm(config.hidden_size)
        self.norm_partial = nn.LayerNorm(config.hidden_size)
        self.norm_random = nn.LayerNorm(config.hidden_size)
        self.norm = nn.LayerNorm(config.hidden_size)
        self.register_buffer('_v_eye', None, persistent=False)
        self.v_eye_learned = nn.Parameter(torch.rand((1, 1, d, d)))
        max_pos = config.max_position_embeddings if hasattr(config, 'max_position_embeddings') else 2048
        self.v_eye_learned_causal = nn.Parameter(torch.randn((1, 1, max_pos, d)))
        self._shape_cache = {}
        self._packed = ops.PackedWeights()
        self._padded_cache = None
        self._w_cache = None

    # ------------------------------------------------------------------------------------------------
    def _cnn_convs(self):
        """(the dilated 3x3 CausalConv2d parameter holders in order, the 1x1 one) of the causal predictor CNN: net indices
        0, 2 (, 4 with PERLIN_HOTFIX_OPT_DEEPER) and 5 (7)."""
        net = self.attention_predictor_cnn[1].module.net
        n3 = 3 if getattr(self, '_deeper', False) else 2
        return [net[2 * i].module for i in range(n3)], net[2 * n3 + 1].module

    def _weights_fp32(self):
        """fp32 contiguous views / copies of the predictor parameters, keyed like the C entries name them.  Frozen module
        (freeze_packed_weights): made once and reused -- a bf16 / fp16 module would otherwise convert ~20 tensors per call."""
        if self._packed.frozen and self._w_cache is not None:
            return dict(self._w_cache)
        w = self._weights_fp32_uncached()
        self._w_cache = dict(w) if self._packed.frozen else None
        return w

    def _weights_fp32_uncached(self):
        f = lambda t: t.detach().float().contiguous()
        enc, dec, scl, cnn = self.attention_predictor_enc, self.attention_predictor_dec_row, self.attention_predictor_dec_scaler, self.attention_predictor_cnn
        w = {
            'enc_w': f(enc[0].weight), 'enc_b': f(enc[0].bias), 'enc_ln_w': f(enc[1].weight), 'enc_ln_b': f(enc[1].bias),
            'dec_w': f(dec[0].weight), 'dec_b': f(dec[0].bias), 'scl_w': f(scl[0].weight), 'scl_b': f(scl[0].bias),
            'proj': f(self.performer.projection_matrix),
        }
        if self.pconfig.causal:
            c3x3, c1x1 = self._cnn_convs()
            w.update({
                'cnn_ln_w': f(cnn[0].module.weight), 'cnn_ln_b': f(cnn[0].module.bias),
                'conv3_w': f(c1x1.weight).reshape(c1x1.weight.shape[0], -1), 'conv3_b': f(c1x1.bias),
                'out_ln_w': f(cnn[2].module.weight), 'out_ln_b': f(cnn[2].module.bias),
                'pos': f(self.v_eye_learned_causal).reshape(-1, self.attention_head_size),
            })
            for i, c in enumerate(c3x3):        # 'conv1', 'conv2' (, 'conv2b': the DEEPER variant's third dilated conv)
                name = ('conv1', 'conv2', 'conv2b')[i]
                w[name + '_w'], w[name + '_b'] = f(c.weight), f(c.bias)
        else:
            net = cnn[0].net
            w.update({'conv1_w': f(net[0].weight), 'conv1_b': f(net[0].bias), 'conv2_w': f(net[2].weight), 'conv2_b': f(net[2].bias),
                      'conv3_w': f(net[5].weight), 'conv3_b': f(net[5].bias)})
        return w

    # ------------------------------------------------------------------------------------------------ weight packings
    def freeze_packed_weights(self, on: bool = True):
        """Declare the predictor weights constant (inference loops, CUDA-graph capture, benchmarks): the bf16 packings and the
        zero-padded conv weights made by the next forward are reused by later ones.  Not frozen (default): they are re-made on
        every call, so ANY way of writing a parameter -- including `.data` writes, which bump no version counter -- is seen."""
        self._packed.freeze(on)
        self._padded_cache = None
        self._w_cache = None
        return self

    def invalidate_packed(self):
        """Drop every cached weight packing (call after modifying parameters of a frozen module)."""
        self._packed.invalidate()
        self._padded_cache = None
        self._w_cache = None

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        if getattr(self, '_packed', None) is not None:
            self.invalidate_packed()
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode: bool = True):
        if mode and getattr(self, '_packed', None) is not None:
            self._packed.freeze(False)
            self._padded_cache = None
            self._w_cache = None
        return super().train(mode)

    def _padded_conv_weights(self, w, C, H):
        """Zero-padded copies of the CNN weights for the 64-channel tcgen05 kernels (conv 3x3: [C,C,5,3] -> [64,64,5,3]; 1x1:
        [H,C] -> [32,64]); cached until a source parameter changes."""
        c3x3, c1x1 = self._cnn_convs()
        src = tuple(t for c in c3x3 + [c1x1] for t in (c.weight, c.bias))
        stamp = tuple((int(t.data_ptr()), int(t._version)) for t in src)
        hit = self._padded_cache
        if hit is not None and hit[0] == stamp and self._packed.frozen:
            return hit[1]
        dev = w['conv1_w'].device
        out = {}
        for name in ('conv1', 'conv2', 'conv2b')[:len(c3x3)]:
            wt = torch.zeros((64, 64, 5, 3), dtype=torch.float32, device=dev)
            wt[:C, :C] = w[name + '_w']
            b = torch.zeros((64,), dtype=torch.float32, device=dev)
            b[:C] = w[name + '_b']
            out[name + '_w'], out[name + '_b'] = wt, b
        w3 = torch.zeros((32, 64), dtype=torch.float32, device=dev)
        w3[:H, :C] = w['conv3_w']
        b3 = torch.zeros((32,), dtype=torch.float32, device=dev)
        b3[:H] = w['conv3_b']
        out['conv3_w'], out['conv3_b'] = w3, b3
        self._padded_cache = (stamp, out)
        return out

    def _shape_consts(self, H, P, T_SRC, T_DST, device):
        key = (H, P, T_SRC, T_DST, self.pconfig.k, self.pconfig.k_oversample, str(device))
        hit = self._shape_cache.get(key)
        if hit is None:
            kpr = _k_per_row_causal(H, self.pconfig.k, self.pconfig.k_oversample, P, T_SRC, T_DST, device)
            hit = (kpr, _csr_alloc_upper_bound(H, self.pconfig.k, P, T_SRC, T_DST, kpr))
            self._shape_cache[key] = hit
        return hit

    def forward(self, q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask,
                attention_scores_truth, context_layer_truth, last_state=None):
        pc = self.pconfig
        dynamic_k = int(os.environ.get('DYNAMIC_K', '0'))     # attention.py:348-351
        if dynamic_k > 0:
            pc.k = dynamic_k
        query_skips = int(os.environ.get('QUERY_SKIPS', '1'))      # attention.py:598
        if query_skips != 1 and (not pc.causal or pc.use_cache or last_state is not None):
            raise SeaError('QUERY_SKIPS > 1 (attention.py:598) is implemented for the causal prefill only')
        if not q.is_cuda:
            raise SeaError('PerlinAttention (sea-attention_b200) runs on CUDA tensors only; there is no CPU path')
        if (pc.use_cache or last_state is not None) and not pc.causal:
            raise SeaError('use_cache / PerlinAttentionState is only defined for the causal model (attention_state.py)')
        if (pc.use_cache or last_state is not None) and getattr(self, '_deeper', False):
            raise SeaError('use_cache with the PERLIN_HOTFIX_OPT_DEEPER predictor is not implemented (the decode state keeps two CNN windows)')
        if pc.attention_predictor_method != 'mlp' or pc.attention_predictor_backend != 'performer' or pc.attention_predictor_enc_per_layer:
            raise SeaError('only the mlp predictor with the performer backend is implemented')
        if pc.context_output_method != 'mix' or pc.random_lookup or pc.out_add_performer_context:
            raise SeaError("only context_output_method='mix' without random lookup is implemented")
        # (k_oversample: the sparse path only scales per_item_top_k with it, attention.py:837-853; the CSR interpolation ignores its
        # `oversampled` argument, causal_resize_m_to_t.py:638 -- both are followed here)
        if self.training or attention_scores_truth is not None or context_layer_truth is not None:
            # training branch (the reference's benchmarking=False path with its distillation losses, attention.py:680-765, 1066-1133,
            # 1328-1332): dense O(T^2) by definition, composed of differentiable torch operations on the GPU + the CUDA top-k (training.py)
            if not pc.causal:
                raise SeaError('the training branch is implemented for the causal model only')
            if pc.use_cache or last_state is not None or query_skips != 1:
                raise SeaError('the training branch takes no decode state and no QUERY_SKIPS')
            from . import training
            return training.forward_train(self, q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask,
                                          attention_scores_truth, context_layer_truth, PerlinAttentionOutput)
        if not pc.causal:
            return self._forward_noncausal(q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask)
        if pc.k_flatten_dim != 'causal_batch' or not pc.k_flatten:
            raise SeaError("causal PerlinAttention needs k_flatten_dim='causal_batch' (perlin_opt.py:227-231)")

        if pc.use_cache or last_state is not None:
            return self._forward_causal_stateful(q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask, last_state)
        return self._forward_causal_prefill(q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask,
                                            query_skips=query_skips)

    def forward_query_block(self, q, k, v, t0: int, t1: int, performer=None):
        """Query-block sharded causal prefill (BASELINE configs[4]): q, k, v [N,H,T,d] hold the whole sequence; returns
        the PerlinAttentionOutput of query rows [t0, t1) only (context_layer [N, t1-t0, H*d]).  Concatenating the blocks of a
        partition of [0, T) reproduces forward() on the whole sequence; blocks need nothing from each other.
        performer = performer_prefix(q, k, v, t_end >= t1): a rank that walks SEVERAL blocks computes the linear-attention stage once
        over its longest prefix and hands it to every block, instead of recomputing the prefix sums per block.  A triple
        (ctx, cumavg, t_start) covers rows [t_start, ...) only (parallel.performer_exchanged: each rank computes its own range and the
        ranks exchange their state sums)."""
        if not self.pconfig.causal:
            raise SeaError('query-block sharding is defined for the causal model')
        if self.training:
            raise SeaError('forward_query_block is an inference path')
        return self._forward_causal_prefill(q, k, v, q, k, v, q, k, None, block=(t0, t1), performer=performer)

    def performer_prefix(self, q, k, v, t_end: int = None):
        """(ctx [N,H,t_end,2d], running mean of v [N,H,t_end,d]) of the first t_end tokens: stage a2 + a3 (+ a13) alone, for
        forward_query_block(..., performer=...)."""
        t_end = q.shape[2] if t_end is None else int(t_end)
        w = self._weights_fp32()
        return ops.performer_causal(q[:, :, :t_end], k[:, :, :t_end], v[:, :, :t_end], w['pos'], w['proj'])

    # ------------------------------------------------------------------------------------------------
    def _forward_causal_prefill(self, q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask, capture=None,
                                block=None, query_skips: int = 1, performer=None):
        """Causal prefill (T_DST == T_SRC).  `capture` (dict) receives the CNN intermediates the decode state is built from.

        block = (t0, t1): query-block sharding of a long prefill: q / k / v hold the whole sequence (K and V are
        replicated on every rank), and only the query rows [t0, t1) are produced -- context [N, t1-t0, H*d], probabilities
        [N,H,t1-t0,P].  No exchange with the other blocks is needed: the Performer prefix sums are recomputed locally over [0, t1)
        (linear, the cheapest stage), the predictor MLP and the dilated causal convolutions run on rows [t0-8, t1) -- each conv
        looks 4 rows back, so 8 halo rows (12 with the DEEPER predictor) make every kept row exact --, top-k is per row (K_t uses the absolute t), and the
        sparse attention takes T_DST = t1-t0 query rows against T_SRC = t1 source tokens."""
        pc = self.pconfig
        N, H, T, d = q.shape
        assert attention_mask is None or attention_mask.shape == (N, 1, T, T), f'causal additive mask must be [N,1,T,T], got {tuple(attention_mask.shape)}'
        assert k.shape == (N, H, T, d) and v.shape == (N, H, T, d)
        if v_for_atten.shape != v.shape:
            raise SeaError('v_for_atten must have the shape of v')
        # LoRA in the approximation (self_attention.py:104-120): the Performer sums take v_for_atten, everything else (performer_value,
        # running mean, sparse attention) takes v
        lora_v = None if v_for_atten.data_ptr() == v.data_ptr() else v_for_atten
        P = pc.attention_predictor_length
        t0, t1 = (0, T) if block is None else (int(block[0]), int(block[1]))
        if not (0 <= t0 < t1 <= T):
            raise SeaError(f'bad query block [{t0}, {t1}) of {T} rows')
        c3x3, c1x1 = self._cnn_convs()
        h0 = max(t0 - 4 * len(c3x3), 0)           # first halo row: every dilated conv looks 4 rows back
        if block is not None:
            if capture is not None or self.output_attentions:
                raise SeaError('a query block returns context and probabilities only (no decode state, no CSR tensors)')
            q, k, v = q[:, :, :t1], k[:, :, :t1], v[:, :, :t1]
            q_for_atten, k_for_atten = q_for_atten[:, :, :t1], k_for_atten[:, :, :t1]
            lora_v = None if lora_v is None else lora_v[:, :, :t1]
            q_for_score, k_for_score = q_for_score[:, :, t0:t1], k_for_score[:, :, :t1]
        row_valid = None
        if self.check_padding and attention_mask is not None:
            # dst_attention_mask = causal_attention_mask[:,:,:,:1] (attention.py:432); the reference reads the
            # whole [N,1,T,T] mask and syncs (:434) -- only the first column matters.
            valid_b = attention_mask[:, 0, :, 0] > -1
            if not bool(valid_b.all()):
                # padded query rows (:512-514, :928-931): v and v_for_atten are zeroed there, their top-k is empty.  (The reference
                # zeroes the caller's v in place; a masked copy is used here.)
                if block is not None:
                    raise SeaError('padded rows and query-block sharding are not combined')
                row_valid = valid_b
                v = v * valid_b.view(N, 1, T, 1).to(v.dtype)
        w = self._weights_fp32()
        S = self.attention_predictor_dec_row_splits
        W = P // self.attention_predictor_dec_row_down_scale
        k_per_row, z_alloc = self._shape_consts(H, P, t1, t1 - t0, q.device)

        # a2+a3 (+ running mean for a13)
        if performer is not None:
            ts = int(performer[2]) if len(performer) > 2 else 0
            if block is None or lora_v is not None or row_valid is not None or ts > h0 or ts + performer[0].shape[2] < t1:
                raise SeaError('a precomputed Performer prefix belongs to a query block and must cover its rows (halo included)')
            ctx, cumavg = performer[0], performer[1]          # rows [ts, ...): sliced to the block below
        elif lora_v is not None:
            if row_valid is not None:
                raise SeaError('padded rows together with a separate v_for_atten are not implemented')
            ctx, _ = ops.performer_causal(q_for_atten, k_for_atten, lora_v, w['pos'], w['proj'])
            _, cumavg = ops.performer_causal(q_for_atten, k_for_atten, v, w['pos'], w['proj'])          # (running mean of v; off the hot path)
        elif row_valid is None:
            ctx, cumavg = ops.performer_causal(q_for_atten, k_for_atten, v, w['pos'], w['proj'])
        else:
            # v_for_atten = cat(v_eye_learned_causal, v) is zeroed on padded rows as a whole: the position half differs per batch item
            parts = [ops.performer_causal(q_for_atten[n:n + 1], k_for_atten[n:n + 1], v[n:n + 1],
                                          w['pos'].reshape(-1, d)[:T] * row_valid[n].view(T, 1).to(w['pos'].dtype), w['proj']) for n in range(N)]
            ctx, cumavg = torch.cat([p_[0] for p_ in parts], dim=0), torch.cat([p_[1] for p_ in parts], dim=0)
        v_mlp = v
        if query_skips > 1:
            # QUERY_SKIPS (attention.py:617-619, 640-644): the predictor MLP + CNN see every query_skips-th row only (consecutively, as
            # a shorter sequence) and every result -- scores and scales -- is repeated query_skips times
            if block is not None or capture is not None or T % query_skips:
                raise SeaError('QUERY_SKIPS > 1 needs T % QUERY_SKIPS == 0 and is not combined with query blocks / the decode state')
            ctx, v_mlp = ctx[:, :, ::query_skips].contiguous(), v[:, :, ::query_skips].contiguous()
        if block is not None:
            ts = (int(performer[2]) if len(performer) > 2 else 0) if performer is not None else 0
            ctx, v_mlp = ctx[:, :, h0 - ts:t1 - ts].contiguous(), v[:, :, h0:t1]
            cumavg = cumavg[:, :, t0 - ts:t1 - ts]
        # a4
        # (weight packings of the tensor-core kernels are cached per module and re-made only when a parameter changes)
        pk = self._packed
        enc, dec, scl = self.attention_predictor_enc, self.attention_predictor_dec_row, self.attention_predictor_dec_scaler
        w['_src_mlp'] = (enc[0].weight, dec[0].weight, scl[0].weight)
        # models whose 2H != 64 (e.g. OPT-125m, H = 12): run the 64-channel tcgen05 MLP / conv kernels on zero-padded channels
        pad_c = (q.dtype == torch.bfloat16 and d == 64 and S * H < 64 and H <= 32 and P % 32 == 0 and W in (16, 32, 64)
                 and ops.conv_umma_supported(q.dtype, W, 64, 64) and ctx.is_contiguous())
        # tensor-core tail: 1x1 conv before the upsample (tcgen05), then tail + softmax + top-k in one kernel
        tc_tail = pad_c or (q.dtype == torch.bfloat16 and ops.conv_umma_supported(q.dtype, W, S * H, H) and P % 32 == 0)
        cw = self._padded_conv_weights(w, S * H, H) if pad_c else w
        # ... and where the shape allows (W = 64, 64 -> 64 -> 32 channels) that 1x1 conv runs inside the second 3x3 conv's kernel
        fuse_c3 = tc_tail and ops.conv3x3_conv1x1_supported(q.dtype, W, 64, cw['conv2_w'].shape[0], cw['conv3_w'].shape[0])
        y = y3 = None
        cnn_in, scales, _ = ops.predictor_mlp(ctx, v_mlp, w, S, W, packed=pk, **({'c_out': 64} if pad_c else {}))
        # the dilated convs 'conv1', 'conv2' (, 'conv2b' with PERLIN_HOTFIX_OPT_DEEPER); the last one may carry the 1x1 conv
        names = ('conv1', 'conv2', 'conv2b')[:len(c3x3)]
        y1 = cnn_in
        for name, cm in zip(names[:-1], c3x3[:-1]):
            y1 = ops.causal_conv3x3_dil2_relu(y1, cw[name + '_w'], cw[name + '_b'], packed=pk, slot=name, src=cm.weight)
            if capture is not None and name == 'conv1':
                capture['cnn_in'], capture['conv1'] = cnn_in, y1
        if fuse_c3:
            y3 = ops.causal_conv3x3_relu_conv1x1(y1, cw[names[-1] + '_w'], cw[names[-1] + '_b'], cw['conv3_w'], cw['conv3_b'], packed=pk, slot=names[-1],
                                                 src=c3x3[-1].weight, slot3='conv3', src3=c1x1.weight)
        else:
            y = ops.causal_conv3x3_dil2_relu(y1, cw[names[-1] + '_w'], cw[names[-1] + '_b'], packed=pk, slot=names[-1], src=c3x3[-1].weight)
        if query_skips > 1:
            if y3 is not None:
                y3 = y3.repeat_interleave(query_skips, dim=1)
            else:
                y = y.repeat_interleave(query_skips, dim=1)
            scales = scales.repeat_interleave(query_skips, dim=2)
        if t0 > h0:                               # drop the halo rows (their conv outputs saw zero padding instead of real rows)
            if y3 is not None:
                y3 = y3[:, t0 - h0:].contiguous()
            else:
                y = y[:, t0 - h0:].contiguous()
            scales = scales[:, :, t0 - h0:].contiguous()
        # a5 .. a7
        kpr = k_per_row.repeat(N) if N > 1 else k_per_row
        if tc_tail:
            if y3 is None:
                y3 = ops.conv1x1_umma(y, cw['conv3_w'], cw['conv3_b'], packed=pk, slot='conv3', src=c1x1.weight)
            if pad_c:
                y3 = y3[..., :H].contiguous()
            want_probs = self.keep_estimated_probs or self.output_attentions or row_valid is not None
            res = ops.predictor_tail_topk(y3, w['conv3_b'], w['out_ln_w'], w['out_ln_b'], kpr, P, want_probs=want_probs,
                                          count_k=pc.k if self.output_attentions else 0)
            probs, bits, crow_counts = res if len(res) == 3 else (res[0], res[1], None)
            if row_valid is not None:             # (off the hot path: the grouped top-k again, with the padded rows' keys zeroed)
                bits, crow_counts = ops.topk_mask_bits(probs, kpr, 'causal_batch', row_valid=row_valid), None
        else:
            probs, _ = ops.predictor_tail(y, w['conv3_w'], w['conv3_b'], w['out_ln_w'], w['out_ln_b'], P)
            bits = ops.topk_mask_bits(probs, kpr, 'causal_batch', row_valid=row_valid)
            crow_counts = None
        if not self.output_attentions and ops.attention_bits_supported(q.dtype, d, P):
            # a8 + a9-a14 in one kernel: the CSR column list is a pure function of the bit mask, so it is only
            # materialised when the caller asks for the CSR tensors (output_attentions)
            if torch.is_grad_enabled() and any(t.requires_grad for t in (q_for_score, k_for_score, v)):
                # training through the sparse path: q, k, v receive the gradient of the masked attention, the
                # scaler and the running-mean mix; the predictor (Performer / MLP / CNN -> mask, scales) is not differentiated
                # here -- in the reference it learns from its own distillation losses (attention.py:707-765), not built yet
                context = ops.sparse_attention_from_bits_autograd(bits, q_for_score, k_for_score, v, scales, cumavg, P, pc.k,
                                                                  use_scaler=pc.partial_attention_scaler, is_causal=True)
            else:
                context = ops.sparse_attention_from_bits(bits, q_for_score, k_for_score, v, scales, cumavg, P, pc.k,
                                                         use_scaler=pc.partial_attention_scaler, is_causal=True)
            pvals = crow = col = None
            Z = 0
        else:
            # a8 (int32 indices internally; no host sync: col is allocated at a shape-derived upper bound)
            crow, col, Z, head_ptr = ops.csr_from_bits(bits, H, P, pc.k, t1, is_causal=True, index_dtype=torch.int32, z_alloc=z_alloc,
                                                       want_head_ptr=True, crow_counts=crow_counts)
            # a9-a14
            context, pvals = ops.sparse_attention(crow, col, q_for_score, k_for_score, v, scales, cumavg,
                                                  use_scaler=pc.partial_attention_scaler, want_probs=self.output_attentions,
                                                  head_ptr=head_ptr)
        partial_probs = partial_mask = None
        if self.output_attentions:
            partial_mask, partial_probs = _csr_outputs(crow, col, pvals, (N, t1 - t0, H * t1))
        return PerlinAttentionOutput(
            loss=0, context_layer=context, partial_attention_probs=partial_probs, partial_attention_mask=partial_mask,
            estimated_attention_probs_m=probs, estimated_attention_probs=probs, dense_attention_probs=None,
            key_for_score=k_for_score, state=None)

    # ------------------------------------------------------------------------------------------------
    def _decode_native_planned(self, plan, st, qn, kc, v, N, H, d, F, P, S, C_win, T_new):
        """The native decode step with every per-module constant taken from the plan (see _forward_causal_stateful)."""
        pc = self.pconfig
        contexts, prob_rows = [], []
        mid, kpr0, ws_ptr, ws_n = plan['mid'], plan['kpr'], plan['ws'], plan['ws_n']
        kargs = (kc.data_ptr(), kc.stride(0), kc.stride(1), kc.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2))
        tail = (N, H, d, F, P, S, C_win)
        k_, scaler = int(pc.k), int(bool(pc.partial_attention_scaler))
        dev = qn.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            for i in range(T_new):
                t = st.t
                new = PerlinAttentionState(t=t + 1, performer=torch.empty_like(st.performer), cnn_in_win=torch.empty_like(st.cnn_in_win),
                                           conv1_win=torch.empty_like(st.conv1_win))
                context = torch.empty((N, 1, H * d), dtype=qn.dtype, device=dev)
                probs = torch.empty((N, H, 1, P), dtype=torch.float32, device=dev)
                qi = qn[:, :, i:i + 1]
                ops._lib.call('sea_decode_step', qi.data_ptr(), qi.stride(0), qi.stride(1), *kargs, *mid, kpr0 + 4 * t,
                              st.performer.data_ptr(), new.performer.data_ptr(), st.cnn_in_win.data_ptr(), new.cnn_in_win.data_ptr(),
                              st.conv1_win.data_ptr(), new.conv1_win.data_ptr(), context.data_ptr(), probs.data_ptr(), ws_ptr, ws_n,
                              *tail, t, k_, scaler, stream)
                st = new
                contexts.append(context)
                prob_rows.append(probs)
        context = contexts[0] if T_new == 1 else torch.cat(contexts, dim=1)
        probs = prob_rows[0] if T_new == 1 else torch.cat(prob_rows, dim=2)
        return PerlinAttentionOutput(
            loss=0, context_layer=context, partial_attention_probs=None, partial_attention_mask=None,
            estimated_attention_probs_m=probs, estimated_attention_probs=probs, dense_attention_probs=None,
            key_for_score=kc, state=st)

    def _forward_causal_stateful(self, q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask, last_state):
        """use_cache / decode path (reference attention_state.py + attention.py:559-572, 627-639, 1222-1236).
        q holds the T_new NEW query tokens, k / v all T_SRC tokens seen so far (the caller's KV cache).
          * last_state is None: prompt prefill through the normal kernels + construction of the state;
          * otherwise: the new tokens are advanced one by one -- incremental Performer / running mean, the predictor MLP on one
            token, the two dilated causal convs on a 5-row window, top-k of one row, sparse attention of one query row.
        Row t of a decode equals row t of a prefill (the reference checks the same property in test_perlin_opt_cache.py)."""
        pc = self.pconfig
        N, H, T_new, d = q.shape
        T_SRC = k.shape[2]
        P = pc.attention_predictor_length
        S = self.attention_predictor_dec_row_splits
        W = P // self.attention_predictor_dec_row_down_scale
        w = self._weights_fp32()
        F = w['proj'].shape[0]
        if last_state is None:
            if T_new != T_SRC:
                raise SeaError('use_cache without a state: the first call must be the prompt prefill (T_DST == T_SRC)')
            cap = {}
            out = self._forward_causal_prefill(q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask, capture=cap)
            st = PerlinAttentionState(t=T_SRC, performer=ops.performer_state_new(N, H, d, F, q.device))
            ops.performer_state_build(k_for_atten, v, w['pos'], w['proj'], st.performer)

            def last_rows(x):            # last 4 rows, zero-padded at the front for prompts shorter than 4 tokens
                win = torch.zeros((N, 4) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
                n_ = min(4, x.shape[1])
                win[:, 4 - n_:] = x[:, x.shape[1] - n_:]
                return win
            st.cnn_in_win, st.conv1_win = last_rows(cap['cnn_in']), last_rows(cap['conv1'])
            return out._replace(state=st)

        if v_for_atten.data_ptr() != v.data_ptr():
            raise SeaError('v_for_atten must alias v (LoRA-in-approximation is not implemented)')
        if last_state.t + T_new != T_SRC:
            raise SeaError(f'state has consumed {last_state.t} tokens, got {T_new} new queries but {T_SRC} keys')
        if T_SRC > self.v_eye_learned_causal.shape[2]:
            raise SeaError(f'decode position {T_SRC} exceeds max_position_embeddings = {self.v_eye_learned_causal.shape[2]} '
                           f'(v_eye_learned_causal has no row for it)')
        # functional update like the reference: the caller's state object is left untouched.  Only the Performer sums are advanced in
        # place by the kernel, so only they are copied; the CNN windows are rebuilt (torch.cat) every step anyway.
        st = PerlinAttentionState(t=last_state.t, performer=last_state.performer, cnn_in_win=last_state.cnn_in_win,
                                  conv1_win=last_state.conv1_win)
        pk = self._packed
        C_win = st.cnn_in_win.shape[-1]
        pad_c = C_win != S * H                       # the state was built on the zero-padded 64-channel path
        # Frozen module (inference loop): everything about the native step that does not change from token to token -- ~30 parameter
        # pointers, the packing slots, the workspace -- is resolved once and kept beside the cached fp32 weights (dropped with them).
        plan_key = (N, H, d, P, S, C_win, int(pc.k), float(pc.k_oversample), q.dtype, q.device, bool(pc.partial_attention_scaler),
                    int(self.v_eye_learned_causal.shape[2]))
        plan = self._w_cache.get('_decode_plan') if (self.decode_native and pk.frozen and self._w_cache is not None) else None
        if plan is not None and plan['key'] != plan_key:
            plan = None
        if plan is not None and (q_for_atten.data_ptr() == q_for_score.data_ptr() and k_for_atten.data_ptr() == k_for_score.data_ptr()
                                 and q.dtype == k.dtype == v.dtype and k.stride(-1) == 1 and v.stride(-1) == 1 and q.stride(-1) == 1
                                 and k_for_score.stride() == k.stride() and st.cnn_in_win.is_contiguous() and st.conv1_win.is_contiguous()):
            return self._decode_native_planned(plan, st, q_for_score, k_for_score, v, N, H, d, F, P, S, C_win, T_new)
        net = self.attention_predictor_cnn[1].module.net
        enc, dec, scl = self.attention_predictor_enc, self.attention_predictor_dec_row, self.attention_predictor_dec_scaler
        w['_src_mlp'] = (enc[0].weight, dec[0].weight, scl[0].weight)
        wp = self._padded_conv_weights(w, S * H, H) if pad_c else w
        # per_item_top_k of every position up to max_position_embeddings, computed once (it depends on shapes only)
        max_pos = self.v_eye_learned_causal.shape[2]
        key = ('decode_kpr', H, P, pc.k, pc.k_oversample, max_pos, str(q.device))
        kpr_all = self._shape_cache.get(key)
        if kpr_all is None:
            kpr_all = _k_per_row_causal(H, pc.k, pc.k_oversample, P, max_pos, max_pos, q.device)
            self._shape_cache[key] = kpr_all
        contexts, prob_rows = [], []
        native = (self.decode_native and q_for_atten.data_ptr() == q_for_score.data_ptr() and k_for_atten.data_ptr() == k_for_score.data_ptr()
                  and q.dtype == k.dtype == v.dtype and k.stride(-1) == 1 and v.stride(-1) == 1 and q.stride(-1) == 1
                  and k_for_score.stride() == k.stride() and st.cnn_in_win.is_contiguous() and st.conv1_win.is_contiguous())
        if native:
            lib = ops._lib.load()
            dcode = ops._dtype_code(q)
            if q.dtype == torch.bfloat16 and lib.sea_predictor_mlp_umma_supported(dcode, H, d, S, W):
                mlp_ws, fresh_m = pk.get('mlp', w['_src_mlp'], lib.sea_predictor_mlp_umma_workspace_bytes(), q.device)
            elif q.dtype == torch.bfloat16 and lib.sea_predictor_mlp_mma_supported(dcode, H, d, S, W):
                mlp_ws, fresh_m = pk.get('mlp_mma', w['_src_mlp'], lib.sea_predictor_mlp_mma_workspace_bytes(d, S, W), q.device)
            else:
                mlp_ws, fresh_m = None, False
            if q.dtype == torch.bfloat16 and C_win == 64 and ops.conv_umma_supported(q.dtype, W, C_win, C_win):
                nb = lib.sea_conv_umma_workspace_bytes(C_win, C_win)
                c1_ws, fresh_1 = pk.get('conv1', (net[0].module.weight,), nb, q.device)
                c2_ws, fresh_2 = pk.get('conv2', (net[2].module.weight,), nb, q.device)
            else:
                c1_ws = c2_ws = None
                fresh_1 = fresh_2 = False
            repack = int(fresh_m or fresh_1 or fresh_2)
            wkey = (N, H, d, P, S, C_win, int(pc.k), q.dtype, str(q.device))
            ws = self._decode_ws.get(wkey)
            if ws is None:
                nbytes = int(lib.sea_decode_step_workspace_bytes(N, H, d, P, S, C_win, int(pc.k), dcode))
                ws = torch.empty((nbytes + 256,), dtype=torch.uint8, device=q.device)
                ws = ws[(-ws.data_ptr()) % 256:][:nbytes]
                self._decode_ws = {wkey: ws}
            qn = q_for_score
            if pk.frozen and not repack and self._w_cache is not None:
                _pp = lambda t_: None if t_ is None else t_.data_ptr()
                self._w_cache['_decode_plan'] = {
                    'key': plan_key, 'keep': (w, wp, mlp_ws, c1_ws, c2_ws, ws, kpr_all), 'kpr': kpr_all.data_ptr(), 'ws': ws.data_ptr(), 'ws_n': ws.numel(),
                    'mid': (dcode, w['pos'].data_ptr(), w['proj'].data_ptr(),
                            w['enc_w'].data_ptr(), w['enc_b'].data_ptr(), w['enc_ln_w'].data_ptr(), w['enc_ln_b'].data_ptr(),
                            w['dec_w'].data_ptr(), w['dec_b'].data_ptr(), w['cnn_ln_w'].data_ptr(), w['cnn_ln_b'].data_ptr(),
                            w['scl_w'].data_ptr(), w['scl_b'].data_ptr(),
                            wp['conv1_w'].data_ptr(), wp['conv1_b'].data_ptr(), wp['conv2_w'].data_ptr(), wp['conv2_b'].data_ptr(),
                            w['conv3_w'].data_ptr(), w['conv3_b'].data_ptr(), w['out_ln_w'].data_ptr(), w['out_ln_b'].data_ptr(),
                            _pp(mlp_ws), _pp(c1_ws), _pp(c2_ws), 0)}
        else:
            st.performer = st.performer.clone()          # the per-op kernel advances the sums in place
        for i in range(T_new):
            t = st.t
            if native:
                # (the new token's k / v rows are rows t of the caches; every intermediate lives in the workspace)
                new = PerlinAttentionState(t=t + 1, performer=torch.empty_like(st.performer), cnn_in_win=torch.empty_like(st.cnn_in_win),
                                           conv1_win=torch.empty_like(st.conv1_win))
                context = torch.empty((N, 1, H * d), dtype=q.dtype, device=q.device)
                probs = torch.empty((N, H, 1, P), dtype=torch.float32, device=q.device)
                qi = qn[:, :, i:i + 1]
                with torch.cuda.device(q.device):
                    ops._lib.call('sea_decode_step', qi.data_ptr(), qi.stride(0), qi.stride(1),
                                  k_for_score.data_ptr(), k_for_score.stride(0), k_for_score.stride(1), k_for_score.stride(2),
                                  v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), dcode,
                                  w['pos'].data_ptr(), w['proj'].data_ptr(),
                                  w['enc_w'].data_ptr(), w['enc_b'].data_ptr(), w['enc_ln_w'].data_ptr(), w['enc_ln_b'].data_ptr(),
                                  w['dec_w'].data_ptr(), w['dec_b'].data_ptr(), w['cnn_ln_w'].data_ptr(), w['cnn_ln_b'].data_ptr(),
                                  w['scl_w'].data_ptr(), w['scl_b'].data_ptr(),
                                  wp['conv1_w'].data_ptr(), wp['conv1_b'].data_ptr(), wp['conv2_w'].data_ptr(), wp['conv2_b'].data_ptr(),
                                  w['conv3_w'].data_ptr(), w['conv3_b'].data_ptr(), w['out_ln_w'].data_ptr(), w['out_ln_b'].data_ptr(),
                                  None if mlp_ws is None else mlp_ws.data_ptr(), None if c1_ws is None else c1_ws.data_ptr(),
                                  None if c2_ws is None else c2_ws.data_ptr(), repack,
                                  kpr_all.data_ptr() + 4 * t,
                                  st.performer.data_ptr(), new.performer.data_ptr(), st.cnn_in_win.data_ptr(), new.cnn_in_win.data_ptr(),
                                  st.conv1_win.data_ptr(), new.conv1_win.data_ptr(),
                                  context.data_ptr(), probs.data_ptr(), ws.data_ptr(), ws.numel(),
                                  N, H, d, F, P, S, C_win, t, int(pc.k), int(bool(pc.partial_attention_scaler)),
                                  torch.cuda.current_stream(q.device).cuda_stream)
                repack = 0
                st = new
                contexts.append(context)
                prob_rows.append(probs)
                continue
            qa, ka, v1 = q_for_atten[:, :, i:i + 1], k_for_atten[:, :, t:t + 1], v[:, :, t:t + 1]
            ctx, cumavg = ops.performer_causal_state(qa, ka, v1, w['pos'], w['proj'], st.performer, t)
            cnn_row, scales, _ = ops.predictor_mlp(ctx, v1, w, S, W, packed=pk, c_out=C_win if pad_c else None)      # [N,1,W,C]
            x_win = torch.cat([st.cnn_in_win, cnn_row], dim=1)                                                       # rows t-4 .. t
            y1 = ops.causal_conv3x3_dil2_relu(x_win, wp['conv1_w'], wp['conv1_b'], packed=pk, slot='conv1', src=net[0].module.weight)[:, 4:5]
            y1_win = torch.cat([st.conv1_win, y1], dim=1)
            y2 = ops.causal_conv3x3_dil2_relu(y1_win, wp['conv2_w'], wp['conv2_b'], packed=pk, slot='conv2', src=net[2].module.weight)[:, 4:5]
            st.cnn_in_win, st.conv1_win = x_win[:, 1:], y1_win[:, 1:]            # views of fresh tensors: the next cat copies them
            y2 = (y2[..., :S * H] if pad_c else y2).contiguous()          # [:, 4:5] is a strided view when N > 1
            probs, _ = ops.predictor_tail(y2, w['conv3_w'], w['conv3_b'], w['out_ln_w'], w['out_ln_b'], P)             # [N,H,1,P]
            kpr = kpr_all[t:t + 1]
            bits = ops.topk_mask_bits(probs, kpr.repeat(N) if N > 1 else kpr, 'causal_batch')
            qs, ks, vs = q_for_score[:, :, i:i + 1], k_for_score[:, :, :t + 1], v[:, :, :t + 1]
            if ops.attention_bits_supported(q.dtype, d, P):
                context = ops.sparse_attention_from_bits(bits, qs, ks, vs, scales, cumavg, P, pc.k, use_scaler=pc.partial_attention_scaler,
                                                         is_causal=True, kernel='gather')
            else:
                crow, col, Z, head_ptr = ops.csr_from_bits(bits, H, P, pc.k, t + 1, is_causal=True, index_dtype=torch.int32, want_head_ptr=True)
                context, _ = ops.sparse_attention(crow, col, qs, ks, vs, scales, cumavg, use_scaler=pc.partial_attention_scaler, want_probs=False,
                                                  head_ptr=head_ptr)
            contexts.append(context)
            prob_rows.append(probs)
            st.t = t + 1
        context = contexts[0] if T_new == 1 else torch.cat(contexts, dim=1)
        probs = prob_rows[0] if T_new == 1 else torch.cat(prob_rows, dim=2)
        return PerlinAttentionOutput(
            loss=0, context_layer=context, partial_attention_probs=None, partial_attention_mask=None,
            estimated_attention_probs_m=probs, estimated_attention_probs=probs, dense_attention_probs=None,
            key_for_score=k_for_score, state=st)

    # ------------------------------------------------------------------------------------------------
    def _forward_noncausal(self, q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask):
        """BERT variant (attention.py with causal=False): eye-grid v_for_atten, FAVOR+ Performer, strided CNN + bilinear
        resize, 'batch' / 'query' / 'causal_batch' top-k, non-causal CSR, sparse attention, probability-weighted mean mix."""
        pc = self.pconfig
        N, H, T, d = q.shape
        assert attention_mask.shape == (N, 1, 1, T), f'non-causal additive mask must be [N,1,1,T], got {tuple(attention_mask.shape)}'
        if v_for_atten.data_ptr() != v.data_ptr():
            raise SeaError('v_for_atten must alias v (LoRA-in-approximation is not implemented)')
        lengths = None
        if self.check_padding:
            valid = attention_mask[:, 0, 0, :] > -1                                      # [N,T]
            if not bool(valid.all()):
                # right-padded batch (attention.py:401-449): item n has lengths[n] real tokens.  v and v_for_atten are zeroed on the padded
                # tokens (:512-514; a masked copy here, the reference writes into the caller's v), the identity grid follows the rank among
                # the valid tokens (:482), the probabilities of padded query rows are zeroed before the top-k (:777-778), the interpolation
                # width is the token length (the reference's dense path, resize_m_to_t.py:36-47), the average context skips padded tokens.
                lengths = valid.sum(-1).to(torch.int32)
                if not bool((valid == (torch.arange(T, device=q.device).view(1, T) < lengths.view(N, 1))).all()):
                    raise SeaError('non-causal padding must be right padding (valid tokens first); other masks are not implemented')
                if bool((lengths < 1).any()):
                    raise SeaError('a batch item without any valid token')
                v = v * valid.view(N, 1, T, 1).to(v.dtype)
        P = pc.attention_predictor_length
        w = self._weights_fp32()
        S, W = self.attention_predictor_dec_row_splits, P // self.attention_predictor_dec_row_down_scale
        ctx = ops.performer_noncausal(q_for_atten, k_for_atten, v, w['proj'], lengths=lengths)
        cnn_in, scales, _ = ops.predictor_mlp(ctx, v, w, S, W)                      # [N,T,W,4H], no leading LayerNorm
        y = ops.conv3x3_cl(cnn_in, w['conv1_w'], w['conv1_b'], stride_t=2, relu=True)
        y = ops.conv3x3_cl(y, w['conv2_w'], w['conv2_b'], relu=True)
        y = ops.conv3x3_cl(y, w['conv3_w'], w['conv3_b'], up=2, relu=False)       # nearest (2,1) upsample folded into the conv
        probs, _ = ops.bert_tail(y, T, P)
        kf = float(pc.k) * float(pc.k_oversample) * P
        tl_ = torch.full((N,), T, dtype=torch.long, device=q.device) if lengths is None else lengths.long()
        row_valid = None
        if lengths is not None:
            row_valid = valid
            probs = probs * valid.view(N, 1, T, 1).to(probs.dtype)                       # :777-778
        mode = pc.k_flatten_dim if pc.k_flatten else 'query'
        if mode == 'batch':
            kpi = torch.clamp_min(torch.round(tl_ * H * (kf / tl_)), 1)                                         # attention.py:837,856,866
            bits = ops.topk_mask_bits_batch(probs, kpi)
        elif mode == 'head':
            kpi = torch.clamp_min(torch.round(tl_ * (kf / tl_)), 1).view(N, 1).expand(N, H).reshape(-1)          # :838-842
            bits = ops.topk_mask_bits_batch(probs, kpi, group_heads=1)
        elif mode == 'query':
            kpi = torch.clamp_min(torch.round(kf / tl_), 1)                                                      # :853
            bits = ops.topk_mask_bits(probs, kpi, 'query', row_valid=row_valid)
        elif mode == 'causal_batch':
            kpi = torch.clamp_min(torch.round(H * (kf / tl_)), 1).view(N, 1).expand(N, T).reshape(-1)             # :846
            bits = ops.topk_mask_bits(probs, kpi, 'causal_batch', row_valid=row_valid)
        else:
            raise SeaError(f"k_flatten_dim='{mode}' is not implemented")
        avg = ops.bert_avg(probs, v, lengths=lengths)
        partial_probs = partial_mask = None
        if lengths is None and not self.output_attentions and ops.attention_bits_supported(q.dtype, d, P):
            context = ops.sparse_attention_from_bits(bits, q_for_score, k_for_score, v, scales, avg, P, pc.k,
                                                     use_scaler=pc.partial_attention_scaler, is_causal=False)
        else:
            # exact-size CSR (one host read of the nnz, like the reference's own .item(), causal_resize_m_to_t.py:667)
            crow, col, Z, head_ptr = ops.csr_from_bits(bits, H, P, pc.k, T, is_causal=False, index_dtype=torch.int32, want_head_ptr=True,
                                                       lengths=lengths)
            context, pvals = ops.sparse_attention(crow, col, q_for_score, k_for_score, v, scales, avg,
                                                  use_scaler=pc.partial_attention_scaler, want_probs=self.output_attentions, head_ptr=head_ptr)
            if self.output_attentions:
                partial_mask, partial_probs = _csr_outputs(crow, col, pvals, (N, T, H * T))
        return PerlinAttentionOutput(
            loss=0, context_layer=context, partial_attention_probs=partial_probs, partial_attention_mask=partial_mask,
            estimated_attention_probs_m=probs, estimated_attention_probs=probs, dense_attention_probs=None,
            key_for_score=k_for_score, state=None)
