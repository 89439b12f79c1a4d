"""Host-side mirror of the reference operator package `src/models/perlin_attention/ops/__init__.py:1-7`
(same names, argument meaning and error behaviour), backed by the sm_100a kernels of libsea_b200.so.
Host synchronisations: none in the stage ops; the standalone CSR ops read the nnz back once, like the reference's `.item()`.

Every function takes CUDA tensors, allocates its outputs with torch (ownership convention of the
reference: ops return fresh tensors, inputs are never mutated) and launches on torch's current stream.
Nothing here computes on the CPU and there is no fallback: a missing library or a CPU tensor raises.
"""
import math
import os
from typing import Optional

import torch

from . import _lib
from ._lib import SeaError

_DTYPES = {torch.float32: _lib.SEA_DTYPE_F32, torch.bfloat16: _lib.SEA_DTYPE_BF16, torch.float16: _lib.SEA_DTYPE_F16}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise SeaError(f'unsupported dtype {t.dtype}')


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SeaError('sea-attention_b200 ops run on CUDA tensors only (no CPU fallback)')


def _p(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _dense(t: torch.Tensor, dtype=None) -> torch.Tensor:
    """Stage ops hand raw pointers with implicit dense strides to the library: make the tensor contiguous (and of the
    expected dtype) instead of silently reading a strided view as if it were dense."""
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t if t.is_contiguous() else t.contiguous()


def _on_tensor_device(fn):
    """Runs an op with the device of its first CUDA tensor argument current, so that the stream handed to the library and the
    context the kernels launch in belong to the tensors' device (modules placed on a non-current GPU, e.g. HF `device_map`)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapped


def _inner_contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.stride(-1) == 1 else t.contiguous()


def _idx64(crow: torch.Tensor, col: torch.Tensor) -> int:
    if crow.dtype != col.dtype or crow.dtype not in (torch.int32, torch.int64):
        raise SeaError('crow/col indices must both be int32 or both int64')
    return 1 if crow.dtype == torch.int64 else 0


# --------------------------------------------------------------------------------------------- masks
def mask_to_bits(mask: torch.Tensor) -> torch.Tensor:
    """0/1 float mask [N,H,T,P] -> u32 bit rows [N,T,ceil(H*P/32)] (bit h*P+m)."""
    _cuda(mask)
    N, H, T, P = mask.shape
    m = _inner_contig(mask.float())
    bits = torch.empty((N, T, (H * P + 31) // 32), dtype=torch.int32, device=mask.device)
    _lib.call('sea_mask_float_to_bits', m.data_ptr(), m.stride(0), m.stride(1), m.stride(2), bits.data_ptr(), N, H, T, P, _stream())
    return bits


def bits_to_mask(bits: torch.Tensor, H: int, P: int) -> torch.Tensor:
    _cuda(bits)
    N, T, _ = bits.shape
    out = torch.empty((N, H, T, P), dtype=torch.float32, device=bits.device)
    _lib.call('sea_mask_bits_to_float', bits.data_ptr(), out.data_ptr(), N, H, T, P, _stream())
    return out


def topk_mask_bits(probs: torch.Tensor, k_per_group: torch.Tensor, group_mode: str = 'causal_batch',
                   row_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Grouped top-k (reference attention.py:774-947).  probs [N,H,T,P] fp32; k_per_group is the
    reference's `per_item_top_k` ([N*T] for 'causal_batch', [N] for 'query')."""
    _cuda(probs, k_per_group, row_valid)
    N, H, T, P = probs.shape
    pr = _inner_contig(probs.float())
    kg = k_per_group.reshape(-1).float().contiguous()
    mode = {'causal_batch': 0, 'query': 1}[group_mode]
    if kg.numel() != (N * T if mode == 0 else N):
        raise SeaError('k_per_group has the wrong number of groups')
    rv = None if row_valid is None else row_valid.reshape(N, T).to(torch.uint8).contiguous()
    bits = torch.empty((N, T, (H * P + 31) // 32), dtype=torch.int32, device=probs.device)
    _lib.call('sea_topk_mask_bits', pr.data_ptr(), pr.stride(0), pr.stride(1), pr.stride(2), kg.data_ptr(), _p(rv),
              bits.data_ptr(), N, H, T, P, mode, _stream())
    return bits


# --------------------------------------------------------------------------------------------- a8
def _lengths_arg(lengths, N, device):
    if lengths is None:
        return None
    ln = lengths.reshape(-1).to(device=device, dtype=torch.int32).contiguous()
    if ln.numel() != N:
        raise SeaError(f'lengths must hold one token count per batch item ({N}), got {ln.numel()}')
    return ln


def csr_from_bits(bits: torch.Tensor, H: int, P: int, k: int, T_SRC: int, is_causal: bool = True,
                  index_dtype=torch.int64, z_alloc: Optional[int] = None, want_head_ptr: bool = False, crow_counts=None, lengths=None):
    """bit mask -> (crow, col, Z).  z_alloc=None reads the exact nnz back (one host sync, like the
    reference's `.item()` at causal_resize_m_to_t.py:667); an int skips the sync and over-allocates.
    lengths (non-causal only): valid tokens per item of a right-padded batch = the interpolation width of that item's rows."""
    N, T_DST, _ = bits.shape
    idx64 = 1 if index_dtype == torch.int64 else 0
    ln = _lengths_arg(lengths, N, bits.device)
    if ln is not None and is_causal:
        raise SeaError('csr_from_bits: per-item lengths belong to the non-causal interpolation')
    if crow_counts is not None:      # per-row counts already produced by the fused predictor tail: only scan them
        crow = crow_counts
        if crow.dtype != index_dtype:
            raise SeaError('crow_counts dtype must equal index_dtype')
        _lib.call('sea_crow_scan', crow.data_ptr(), idx64, N, T_DST, _stream())
    else:
        crow = torch.empty((N, T_DST + 1), dtype=index_dtype, device=bits.device)
        _lib.call('sea_csr_count_len', bits.data_ptr(), crow.data_ptr(), idx64, N, H, T_DST, P, T_SRC, int(k), int(is_causal), _p(ln), _stream())
    Z = int(crow[:, -1].max().item()) if z_alloc is None else int(z_alloc)
    col = torch.empty((N, Z), dtype=index_dtype, device=bits.device)
    hp = torch.empty((N, T_DST, H + 1), dtype=torch.int32, device=bits.device) if (want_head_ptr and P % 32 == 0) else None
    _lib.call('sea_csr_fill_len', bits.data_ptr(), crow.data_ptr(), col.data_ptr(), idx64, Z, _p(hp), N, H, T_DST, P, T_SRC, int(k),
              int(is_causal), _p(ln), _stream())
    if want_head_ptr:
        return crow, col, Z, hp
    return crow, col, Z


def resize_from_m_to_t_csr(x, masked_fill_value, k, target_width=None, training=False, need_assert=False, is_causal=True,
                           max_col_z=None, benchmarking=False, oversampled=None):
    """Mirror of ops/kernels/causal_resize_m_to_t.py:910-921.  x: 0/1 mask [N,H,T_DST,T_M] ->
    batched torch.sparse_csr_tensor [N, T_DST, H*T_SRC] (int64 indices, values = ones of x.dtype)."""
    assert not training
    assert masked_fill_value == 0
    _cuda(x)
    N, H, T_DST, T_M = x.shape
    T_SRC = target_width if target_width is not None else T_DST
    bits = mask_to_bits(x)
    crow, col, Z = csr_from_bits(bits, H, T_M, k, T_SRC, is_causal=is_causal, index_dtype=torch.int64)
    values = torch.ones((N, Z), dtype=x.dtype, device=x.device)
    return torch.sparse_csr_tensor(crow_indices=crow, col_indices=col, values=values, size=(N, T_DST, H * T_SRC))


def flat_csr_to_dense(csr: torch.Tensor, T_SRC: int, H: int) -> torch.Tensor:
    """Mirror of ops/kernels/flat_csr_to_dense.py:3 -> [N,H,T_DST,T_SRC]."""
    assert csr.is_sparse_csr
    N, T_DST, H_T = csr.shape
    crow, col, values = csr.crow_indices(), csr.col_indices(), csr.values()
    _cuda(crow)
    out = torch.empty((N, H, T_DST, T_SRC), dtype=torch.float32, device=crow.device)
    vals = values.float().contiguous()
    _lib.call('sea_flat_csr_to_dense', crow.data_ptr(), col.data_ptr(), _idx64(crow, col), vals.data_ptr(), col.shape[-1],
              out.data_ptr(), N, H, T_DST, T_SRC, _stream())
    return out.to(values.dtype)


# --------------------------------------------------------------------------------------------- a9-a12
def _rebuild(csr, values):
    return torch.sparse_csr_tensor(crow_indices=csr.crow_indices(), col_indices=csr.col_indices(), values=values, size=csr.shape)


def flat_csr_masked_bmm(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor, max_z_per_row: int = None):
    """Mirror of ops/kernels/flat_csr_masked_bmm.py:137."""
    assert mask.is_sparse_csr
    assert a.ndim == b.ndim
    assert a.ndim == 4
    N, H, T_DST, HID = a.shape
    assert b.shape[:2] == (N, H)
    _, _, T_SRC, HID = b.shape
    assert mask.shape == (N, T_DST, H * T_SRC)
    _cuda(a, b)
    if a.dtype != b.dtype:
        raise SeaError('a and b must share a dtype')
    a, b = _inner_contig(a), _inner_contig(b)
    crow, col = mask.crow_indices(), mask.col_indices()
    Z = col.shape[-1]
    out = torch.zeros((N, Z), dtype=torch.float32, device=a.device)
    _lib.call('sea_flat_csr_masked_bmm', crow.data_ptr(), col.data_ptr(), _idx64(crow, col), Z,
              a.data_ptr(), a.stride(0), a.stride(1), a.stride(2), b.data_ptr(), b.stride(0), b.stride(1), b.stride(2),
              _dtype_code(a), out.data_ptr(), N, H, T_DST, T_SRC, HID, _stream())
    return _rebuild(mask, out.to(mask.values().dtype))


def flat_csr_softmax(scores: torch.Tensor, H: int, T_SRC: int, max_z_per_row: int = None):
    """Mirror of ops/kernels/flat_csr_softmax.py:127."""
    assert scores.is_sparse_csr
    crow, col = scores.crow_indices(), scores.col_indices()
    _cuda(crow)
    vin = scores.values().float().contiguous()
    N, R1 = crow.shape
    out = torch.zeros_like(vin)
    _lib.call('sea_flat_csr_softmax', crow.data_ptr(), col.data_ptr(), _idx64(crow, col), col.shape[-1], vin.data_ptr(),
              out.data_ptr(), N, H, R1 - 1, T_SRC, _stream())
    return _rebuild(scores, out.to(scores.values().dtype))


def flat_csr_elmul(probs: torch.Tensor, dense: torch.Tensor, max_z_per_row: int = None):
    """Mirror of ops/kernels/flat_csr_elmul.py:110 (dense may be a stride-0 expanded view)."""
    assert probs.is_sparse_csr
    N, T_DST, H_T = probs.shape
    _N, H, _T_DST, T = dense.shape
    assert T_DST == _T_DST
    assert N == _N
    assert H_T == H * T
    crow, col = probs.crow_indices(), probs.col_indices()
    _cuda(crow, dense)
    vin = probs.values().float().contiguous()
    d = dense if dense.dtype == torch.float32 else dense.float()
    out = torch.zeros_like(vin)
    _lib.call('sea_flat_csr_elmul', crow.data_ptr(), col.data_ptr(), _idx64(crow, col), col.shape[-1], vin.data_ptr(),
              out.data_ptr(), d.data_ptr(), d.stride(0), d.stride(1), d.stride(2), d.stride(3), N, H, T_DST, T, _stream())
    return _rebuild(probs, out.to(probs.values().dtype))


def flat_csr_sdbmm(scores: torch.Tensor, value_layer: torch.Tensor, T_M: int, max_z_per_row: int = None, benchmarking: bool = False):
    """Mirror of ops/kernels/flat_csr_sdbmm.py:323 -> dense fp32 [N,H,T_DST,HID]."""
    assert scores.is_sparse_csr
    crow, col = scores.crow_indices(), scores.col_indices()
    _cuda(crow, value_layer)
    N, R1 = crow.shape
    _N, H, T_SRC, HID = value_layer.shape
    assert N == _N
    _N, T_DST, HT_SRC = scores.shape
    assert HT_SRC == H * T_SRC
    v = _inner_contig(value_layer)
    vals = scores.values().float().contiguous()
    out = torch.empty((N, H, T_DST, HID), dtype=torch.float32, device=v.device)
    _lib.call('sea_flat_csr_sdbmm', crow.data_ptr(), col.data_ptr(), _idx64(crow, col), col.shape[-1], vals.data_ptr(),
              v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), _dtype_code(v), out.data_ptr(), N, H, T_DST, T_SRC, HID, _stream())
    return out


# --------------------------------------------------------------------------------------------- a16
def resize_from_m_to_t(x: torch.Tensor, masked_fill_value: float, attention_mask: torch.Tensor, target_width: int = None,
                       training=False, is_causal=True, k=None, oversampled=None):
    """Mirror of ops/kernels/resize_m_to_t.py:6 (inference form: training=False, oversampled None/1.0)."""
    assert masked_fill_value is not None
    if training:
        raise SeaError('resize_from_m_to_t: the training-time jitter (resize_m_to_t.py:40-45) is not part of the hot path')
    if oversampled is not None and float(oversampled) != 1.0:
        raise SeaError('resize_from_m_to_t: k_oversample != 1.0 is not supported')
    _cuda(x, attention_mask)
    N, H, T1, T_M = x.shape
    _N, _H, _TQ, _TK = attention_mask.shape
    assert _H == 1
    T2 = target_width if target_width is not None else T1
    if is_causal:
        assert attention_mask.shape == (N, 1, T1, T2)
    else:
        assert attention_mask.shape == (N, 1, 1, T2), f"{attention_mask.shape} == {T2}"
    am = _inner_contig(attention_mask.float())
    xs = x.float().contiguous()
    out = torch.empty((N, H, T1, T2), dtype=torch.float32, device=x.device)
    _lib.call('sea_resize_m_to_t_dense', xs.data_ptr(), float(masked_fill_value), am.data_ptr(), am.stride(0),
              am.stride(2) if is_causal else 0, out.data_ptr(), N, H, T1, T_M, T2, _stream())
    return out.to(x.dtype)


# --------------------------------------------------------------------------------------------- dense stages
def performer_causal(q, k, v, pos_emb, proj, want_cumavg=True, force_simt=False):
    """a2+a3 (+ running mean of v).  q,k,v [N,H,T,D]; pos_emb fp32 [>=T, D]; proj fp32 [F, D]
    -> ctx [N,H,T,2D], cumavg [N,H,T,D] (dtype of q)."""
    _cuda(q, k, v, pos_emb, proj)
    N, H, T, D = q.shape
    F = proj.shape[0]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    pos = pos_emb.reshape(-1, D).float().contiguous()
    if pos.shape[0] < T:
        raise SeaError(f'v_eye_learned_causal holds {pos.shape[0]} positions < T={T}')
    pj = proj.float().contiguous()
    ctx = torch.empty((N, H, T, 2 * D), dtype=q.dtype, device=q.device)
    avg = torch.empty((N, H, T, D), dtype=q.dtype, device=q.device) if want_cumavg else None
    lib = _lib.load()
    if not force_simt and lib.sea_performer_mma_supported(_DTYPES.get(q.dtype, -1), D, F) and all(
            t.stride(i) % 8 == 0 for t in (q, k, v) for i in range(3)):
        ws = torch.empty((lib.sea_performer_mma_workspace_floats(N, H, T, D, F),), dtype=torch.float32, device=q.device)
        _lib.call('sea_performer_causal_mma_fwd', q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
                  k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
                  pos.data_ptr(), pj.data_ptr(), ctx.data_ptr(), _p(avg), ws.data_ptr(), N, H, T, D, F, _stream())
        return ctx, avg
    ws = torch.empty((lib.sea_performer_workspace_floats(N, H, T, D, F),), dtype=torch.float32, device=q.device)
    _lib.call('sea_performer_causal_fwd', q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
              k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
              pos.data_ptr(), pj.data_ptr(), _dtype_code(q), ctx.data_ptr(), _p(avg), ws.data_ptr(), N, H, T, D, F, _stream())
    return ctx, avg


def performer_range_supported(q, proj) -> bool:
    """Can the ranged (sharded-sequence) Performer entries take these tensors?  (bf16, a head dim / feature count the MMA kernels cover)"""
    return bool(q.is_cuda and _lib.load().sea_performer_mma_supported(_DTYPES.get(q.dtype, -1), q.shape[-1], proj.shape[0]))


def performer_causal_range_sums(k, v, pos_rows, proj):
    """Phase 1 of the Performer over a RANGE of a sequence sharded over ranks: k, v [N,H,Tr,D] and pos_rows fp32 [Tr,D] are
    the rows of the range.  -> (workspace with the per-chunk sums, total fp32 [state_floats]): `total` = (S, z, sum of v) of the range,
    the quantity the ranks exchange."""
    _cuda(k, v, pos_rows, proj)
    N, H, Tr, D = k.shape
    F = proj.shape[0]
    k, v = _inner_contig(k), _inner_contig(v)
    pos = _dense(pos_rows.reshape(-1, D), torch.float32)
    pj = _dense(proj, torch.float32)
    lib = _lib.load()
    if pos.shape[0] < Tr or not lib.sea_performer_mma_supported(_DTYPES.get(k.dtype, -1), D, F):
        raise SeaError('performer_causal_range_sums: unsupported dtype / head dim / feature count, or too few position rows')
    ws = torch.empty((lib.sea_performer_mma_workspace_floats(N, H, Tr, D, F),), dtype=torch.float32, device=k.device)
    total = torch.zeros((lib.sea_performer_mma_state_floats(N, H, D, F),), dtype=torch.float32, device=k.device)     # (the kernels skip the zero padding rows)
    _lib.call('sea_performer_causal_mma_range', None, 0, 0, 0, k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
              pos.data_ptr(), pj.data_ptr(), None, None, ws.data_ptr(), None, total.data_ptr(), N, H, Tr, D, F, 0, 1, _stream())
    return ws, total


def performer_causal_range_out(q, k, v, pos_rows, proj, ws, init, t_off: int, want_cumavg=True):
    """Phase 2: outputs of the range whose per-chunk sums `ws` holds (performer_causal_range_sums on the same rows), starting from `init`
    = the summed totals of everything before the range (None = the range starts the sequence); t_off = absolute position of its first
    row.  -> ctx [N,H,Tr,2D], cumavg [N,H,Tr,D]."""
    _cuda(q, k, v, pos_rows, proj, ws, init)
    N, H, Tr, D = q.shape
    F = proj.shape[0]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    pos = _dense(pos_rows.reshape(-1, D), torch.float32)
    pj = _dense(proj, torch.float32)
    ini = None if init is None else _dense(init, torch.float32)
    ctx = torch.empty((N, H, Tr, 2 * D), dtype=q.dtype, device=q.device)
    avg = torch.empty((N, H, Tr, D), dtype=q.dtype, device=q.device) if want_cumavg else None
    _lib.call('sea_performer_causal_mma_range', q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
              v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), pos.data_ptr(), pj.data_ptr(), ctx.data_ptr(), _p(avg), ws.data_ptr(), _p(ini), None,
              N, H, Tr, D, F, int(t_off), 2, _stream())
    return ctx, avg


class PackedWeights:
    """Per-module holder of the bf16 weight packings the tensor-core kernels keep in their workspace.

    Default (not frozen): the packing kernels run on EVERY call (three tiny launches, ~10 us) -- nothing is trusted across calls,
    because PyTorch offers no reliable change stamp: `Tensor._version` is not bumped by writes through `.data`
    (`p.data.copy_()`, HF `_init_weights`, DeepSpeed / apex master -> bf16 copies, LoRA merges).
    `freeze()` (PerlinAttention.freeze_packed_weights) declares the weights constant: packings made by the next call are then
    reused (weight pointer NULL in the C call) until `invalidate()` -- for inference loops, CUDA-graph capture and the
    benchmark.  Even when frozen a packing is dropped when a source tensor's (data_ptr, _version) changes."""

    def __init__(self):
        self._slots = {}
        self.frozen = False

    def freeze(self, on: bool = True):
        self.frozen = bool(on)
        if not on:
            self._slots.clear()

    def invalidate(self):
        self._slots.clear()

    def get(self, slot, sources, nbytes, device):
        """-> (workspace tensor, fresh): fresh = the caller must let the kernel re-pack."""
        stamp = tuple((int(t.data_ptr()), int(t._version)) for t in sources) + (str(device), int(nbytes))
        hit = self._slots.get(slot)
        if hit is not None and hit[0][-2:] == stamp[-2:]:
            if self.frozen and hit[0] == stamp:
                return hit[1], False
            self._slots[slot] = (stamp, hit[1])
            return hit[1], True               # same buffer (stable address under CUDA graphs), packed again
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        self._slots[slot] = (stamp, ws)
        return ws, True


def predictor_mlp(ctx, v, w, S: int, W: int, want_t_pred=False, force_simt=False, packed: 'PackedWeights' = None, c_out: int = None,
                  force_mma=False):
    """a4.  `w` = dict of fp32 contiguous weights (enc_w, enc_b, enc_ln_w, enc_ln_b, dec_w, dec_b, cnn_ln_w, cnn_ln_b,
    scl_w, scl_b) -> cnn_in [N,T,W,H*S] channels-last, scales fp32 [N,H,T,2], t_pred or None."""
    _cuda(ctx, v)
    N, H, T, D2 = ctx.shape
    D = D2 // 2
    v = _inner_contig(v)
    c_out = H * S if c_out is None else int(c_out)      # > H*S: zero-padded channels (tensor-core path only)
    cnn_in = torch.empty((N, T, W, c_out), dtype=ctx.dtype, device=ctx.device)
    scales = torch.empty((N, H, T, 2), dtype=torch.float32, device=ctx.device)
    lib = _lib.load()
    if (not want_t_pred and not force_simt and not force_mma and ctx.is_contiguous() and w.get('cnn_ln_w') is not None
            and lib.sea_predictor_mlp_umma_supported(_DTYPES.get(ctx.dtype, -1), H, D, S, W)
            and v.stride(0) % 8 == 0 and v.stride(1) % 8 == 0 and v.stride(2) % 8 == 0):
        nbytes = lib.sea_predictor_mlp_umma_workspace_bytes()
        if packed is not None:
            ws, fresh = packed.get('mlp', w.get('_src_mlp', (w['enc_w'], w['dec_w'], w['scl_w'])), nbytes, ctx.device)
        else:
            ws, fresh = torch.empty((nbytes,), dtype=torch.uint8, device=ctx.device), True
        wp = (lambda t: t.data_ptr()) if fresh else (lambda t: None)
        _lib.call('sea_predictor_mlp_umma_fwd_ex', ctx.data_ptr(), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
                  wp(w['enc_w']), w['enc_b'].data_ptr(), w['enc_ln_w'].data_ptr(), w['enc_ln_b'].data_ptr(),
                  wp(w['dec_w']), w['dec_b'].data_ptr(), w['cnn_ln_w'].data_ptr(), w['cnn_ln_b'].data_ptr(),
                  wp(w['scl_w']), w['scl_b'].data_ptr(), cnn_in.data_ptr(), scales.data_ptr(), ws.data_ptr(),
                  N, H, T, D, S, W, c_out, _stream(), kernels=2 if fresh else 1)
        return cnn_in, scales, None
    if (not want_t_pred and not force_simt and ctx.is_contiguous() and w.get('cnn_ln_w') is not None
            and lib.sea_predictor_mlp_mma_supported(_DTYPES.get(ctx.dtype, -1), H, D, S, W)
            and v.stride(0) % 8 == 0 and v.stride(1) % 8 == 0 and v.stride(2) % 8 == 0 and W * c_out * 2 <= 32 * 1024 and c_out % 8 == 0):
        # any head dim (D = 80, 128, ...): warp-level tensor-core GEMMs with streamed weights (csrc/mlp_mma.cu)
        nbytes = lib.sea_predictor_mlp_mma_workspace_bytes(D, S, W)
        if packed is not None:
            ws, fresh = packed.get('mlp_mma', w.get('_src_mlp', (w['enc_w'], w['dec_w'], w['scl_w'])), nbytes, ctx.device)
        else:
            ws, fresh = torch.empty((nbytes,), dtype=torch.uint8, device=ctx.device), True
        wp = (lambda t: t.data_ptr()) if fresh else (lambda t: None)
        _lib.call('sea_predictor_mlp_mma_fwd', ctx.data_ptr(), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
                  wp(w['enc_w']), w['enc_b'].data_ptr(), w['enc_ln_w'].data_ptr(), w['enc_ln_b'].data_ptr(),
                  wp(w['dec_w']), w['dec_b'].data_ptr(), w['cnn_ln_w'].data_ptr(), w['cnn_ln_b'].data_ptr(),
                  wp(w['scl_w']), w['scl_b'].data_ptr(), cnn_in.data_ptr(), scales.data_ptr(), ws.data_ptr(),
                  N, H, T, D, S, W, c_out, _stream(), kernels=2 if fresh else 1)
        return cnn_in, scales, None
    if c_out != H * S:
        raise SeaError('predictor_mlp: padded output channels need a tensor-core kernel (bf16, contiguous ctx)')
    t_pred = torch.empty((N, H, T, D2), dtype=ctx.dtype, device=ctx.device) if want_t_pred else None
    _lib.call('sea_predictor_mlp_fwd', ctx.data_ptr(), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), _dtype_code(ctx),
              w['enc_w'].data_ptr(), w['enc_b'].data_ptr(), w['enc_ln_w'].data_ptr(), w['enc_ln_b'].data_ptr(),
              w['dec_w'].data_ptr(), w['dec_b'].data_ptr(), _p(w.get('cnn_ln_w')), _p(w.get('cnn_ln_b')),
              w['scl_w'].data_ptr(), w['scl_b'].data_ptr(), cnn_in.data_ptr(), scales.data_ptr(), _p(t_pred),
              N, H, T, D, S, W, _stream())
    return cnn_in, scales, t_pred


def conv_umma_supported(dtype, W, C, O) -> bool:
    return bool(_lib.load().sea_conv_umma_supported(_DTYPES.get(dtype, -1), W, C, O))


def _conv_ws(C, O, device):
    return torch.empty((_lib.load().sea_conv_umma_workspace_bytes(C, O),), dtype=torch.uint8, device=device)


def _packed_conv_ws(packed, slot, weight, src, C, O, device):
    nbytes = _lib.load().sea_conv_umma_workspace_bytes(C, O)
    if packed is None:
        return torch.empty((nbytes,), dtype=torch.uint8, device=device), weight.data_ptr()
    ws, fresh = packed.get(slot, (src if src is not None else weight,), nbytes, device)
    return ws, (weight.data_ptr() if fresh else None)


def causal_conv3x3_dil2_relu(x, weight, bias, force_simt=False, packed: 'PackedWeights' = None, slot: str = 'conv', src=None):
    """a5: one CausalConv2d(C,O,3,padding=2,dilation=2,causal)+ReLU on channels-last x [N,T,W,C];
    weight fp32 in the reference layout [O,C,5,3].  bf16 with C=O=64 runs the tcgen05 implicit-GEMM kernel,
    everything else the fp32 SIMT kernel."""
    _cuda(x, weight, bias)
    x, weight, bias = _dense(x), _dense(weight, torch.float32), _dense(bias, torch.float32)
    N, T, W, C = x.shape
    O = weight.shape[0]
    y = torch.empty((N, T, W, O), dtype=x.dtype, device=x.device)
    if not force_simt and O == 64 and conv_umma_supported(x.dtype, W, C, O):
        ws, wptr = _packed_conv_ws(packed, slot, weight, src, C, O, x.device)
        _lib.call('sea_causal_conv3x3_dil2_relu_umma', x.data_ptr(), wptr, bias.data_ptr(), y.data_ptr(), ws.data_ptr(),
                  N, T, W, C, O, _stream(), kernels=2 if wptr is not None else 1)
        return y
    _lib.call('sea_causal_conv3x3_dil2_relu', x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), _dtype_code(x),
              N, T, W, C, O, _stream())
    return y


def conv1x1_umma(x, weight, bias, packed: 'PackedWeights' = None, slot: str = 'conv1x1', src=None):
    """1x1 CausalConv2d(64 -> 32) before the upsample, tcgen05: x bf16 [N,T,W,64] -> y fp32 [N,T,W,32]."""
    _cuda(x, weight, bias)
    x, weight, bias = _dense(x), _dense(weight, torch.float32), _dense(bias, torch.float32)
    N, T, W, C = x.shape
    O = weight.shape[0]
    y = torch.empty((N, T, W, O), dtype=torch.float32, device=x.device)
    ws, wptr = _packed_conv_ws(packed, slot, weight, src, C, O, x.device)
    _lib.call('sea_conv1x1_umma', x.data_ptr(), wptr, bias.data_ptr(), y.data_ptr(), ws.data_ptr(), N, T, W, C, O, _stream(),
              kernels=2 if wptr is not None else 1)
    return y


def conv3x3_conv1x1_supported(dtype, W: int, C: int, O: int, O3: int) -> bool:
    return dtype == torch.bfloat16 and bool(_lib.load().sea_conv3x3_conv1x1_umma_supported(_lib.SEA_DTYPE_BF16, W, C, O, O3))


def causal_conv3x3_relu_conv1x1(x, weight, bias, weight3, bias3, packed: 'PackedWeights' = None, slot: str = 'conv', src=None,
                                slot3: str = 'conv1x1', src3=None):
    """a5: the second CausalConv2d(64,64,3,dilation 2)+ReLU and the 1x1 CausalConv2d(64 -> 32) behind it in one tcgen05 kernel:
    x bf16 [N,T,W,64] -> y3 fp32 [N,T,W,32]; the activation between the two convolutions never reaches HBM."""
    _cuda(x, weight, bias, weight3, bias3)
    x, weight, bias, weight3, bias3 = _dense(x), _dense(weight, torch.float32), _dense(bias, torch.float32), _dense(weight3, torch.float32), _dense(bias3, torch.float32)
    N, T, W, C = x.shape
    O, O3 = weight.shape[0], weight3.shape[0]
    y3 = torch.empty((N, T, W, O3), dtype=torch.float32, device=x.device)
    ws, wptr = _packed_conv_ws(packed, slot, weight, src, C, O, x.device)
    ws3, wptr3 = _packed_conv_ws(packed, slot3, weight3, src3, C, O3, x.device)
    _lib.call('sea_causal_conv3x3_dil2_relu_conv1x1_umma', x.data_ptr(), wptr, bias.data_ptr(), ws.data_ptr(),
              wptr3, bias3.data_ptr(), ws3.data_ptr(), y3.data_ptr(), N, T, W, C, O, O3, _stream(),
              kernels=1 + (wptr is not None) + (wptr3 is not None))
    return y3


def performer_state_new(N: int, H: int, D: int, F: int, device) -> torch.Tensor:
    """Zeroed decode state of the causal Performer + running mean (fp32; S | z | vsum per (n, h))."""
    return torch.zeros((int(_lib.load().sea_performer_state_floats(N, H, D, F)),), dtype=torch.float32, device=device)


def performer_state_build(k, v, pos_emb, proj, state: torch.Tensor):
    """Adds the running sums of a whole prompt (k, v [N,H,T,D]) to `state` (in place): the state a prefill leaves for the decode."""
    _cuda(k, v, pos_emb, proj, state)
    N, H, T, D = k.shape
    k, v = _inner_contig(k), _inner_contig(v)
    pos_emb, proj = _dense(pos_emb.reshape(-1, D), torch.float32), _dense(proj, torch.float32)
    if pos_emb.shape[0] < T:
        raise SeaError(f'v_eye_learned_causal holds {pos_emb.shape[0]} positions < T={T}')
    _lib.call('sea_performer_state_build', k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
              pos_emb.data_ptr(), proj.data_ptr(), _dtype_code(k), state.data_ptr(), N, H, T, D, proj.shape[0], _stream())
    return state


def performer_causal_state(q, k, v, pos_emb, proj, state: torch.Tensor, t0: int, want_cumavg=True):
    """a2 + a3 + a13 advanced token by token from `state` (in place): q, k, v [N,H,T_new,D] are the NEW tokens at positions
    t0 .. t0+T_new-1.  Returns ctx [N,H,T_new,2D], cumavg [N,H,T_new,D]."""
    _cuda(q, k, v, pos_emb, proj, state)
    N, H, T_new, D = q.shape
    F = proj.shape[0]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    pos_emb, proj = _dense(pos_emb.reshape(-1, D), torch.float32), _dense(proj, torch.float32)
    if pos_emb.shape[0] < int(t0) + T_new:
        raise SeaError(f'v_eye_learned_causal holds {pos_emb.shape[0]} positions < t0 + T_new = {int(t0) + T_new} '
                       f'(decode past max_position_embeddings)')
    ctx = torch.empty((N, H, T_new, 2 * D), dtype=q.dtype, device=q.device)
    avg = torch.empty((N, H, T_new, D), dtype=q.dtype, device=q.device) if want_cumavg else None
    _lib.call('sea_performer_causal_state_fwd', q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
              v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), pos_emb.data_ptr(), proj.data_ptr(), _dtype_code(q), state.data_ptr(),
              ctx.data_ptr(), _p(avg), N, H, T_new, int(t0), D, F, _stream())
    return ctx, avg


def predictor_tail_topk(y3, bias, ln_w, ln_b, k_per_row, P: int, want_probs=True, want_bits=True, count_k: int = 0):
    """a5 tail + a6 + a7 fused: y3 fp32 [N,T,W,H] (conv1x1_umma) -> probs fp32 [N,H,T,P], top-k bit mask [N,T,H*P/32].
    count_k > 0 additionally fuses pass 1 of a8 (causal prefill): returns int32 crow [N,T+1] holding per-row counts."""
    _cuda(y3, bias, ln_w, ln_b, k_per_row)
    y3, bias, ln_w, ln_b = _dense(y3, torch.float32), _dense(bias, torch.float32), _dense(ln_w, torch.float32), _dense(ln_b, torch.float32)
    N, T, W, H = y3.shape
    probs = torch.empty((N, H, T, P), dtype=torch.float32, device=y3.device) if want_probs else None
    bits = torch.empty((N, T, (H * P) // 32), dtype=torch.int32, device=y3.device) if want_bits else None
    crow = torch.empty((N, T + 1), dtype=torch.int32, device=y3.device) if count_k > 0 else None
    kpr = None if k_per_row is None else k_per_row.reshape(-1).float().contiguous()
    _lib.call('sea_predictor_tail_topk_fwd', y3.data_ptr(), bias.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), _p(kpr), _p(probs),
              _p(bits), _p(crow), int(count_k), N, H, T, W, P, _stream())
    if count_k > 0:
        return probs, bits, crow
    return probs, bits


def predictor_tail(x, weight, bias, ln_w, ln_b, P: int, want_scores=False):
    """a5 tail + a6 -> probs fp32 [N,H,T,P] (and the pre-softmax scores when asked)."""
    _cuda(x, weight)
    x, weight, bias, ln_w, ln_b = _dense(x), _dense(weight, torch.float32), _dense(bias, torch.float32), _dense(ln_w, torch.float32), _dense(ln_b, torch.float32)
    N, T, W, C = x.shape
    H = weight.shape[0]
    probs = torch.empty((N, H, T, P), dtype=torch.float32, device=x.device)
    scores = torch.empty_like(probs) if want_scores else None
    _lib.call('sea_predictor_tail_fwd', x.data_ptr(), _dtype_code(x), weight.data_ptr(), bias.data_ptr(), ln_w.data_ptr(),
              ln_b.data_ptr(), probs.data_ptr(), _p(scores), N, H, T, W, C, P, _stream())
    return probs, scores


def _avg_arg(avg, N, H, T_DST, D):
    """The mix-in average: per-row causal running mean [N,H,T_DST,D] or the BERT per-head mean [N,H,1,D] / [N,H,D]."""
    if avg is None:
        return None, 0, 0
    a = avg.contiguous()
    if a.numel() == N * H * T_DST * D and T_DST > 1:
        return a, T_DST * D, D
    if a.numel() == N * H * D:
        return a, D, 0
    raise SeaError(f'average context has unexpected shape {tuple(avg.shape)}')


def sparse_attention(crow, col, q, k, v, scales, cumavg, use_scaler=True, want_probs=False, head_ptr=None):
    """a9-a14 fused -> context [N,T_DST,H*D] (dtype of q) and, when asked, the probabilities [N,Z] fp32."""
    _cuda(crow, col, q, k, v, scales, cumavg)
    N, H, T_DST, D = q.shape
    T_SRC = k.shape[2]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    Z = col.shape[-1]
    out = torch.empty((N, T_DST, H * D), dtype=q.dtype, device=q.device)
    pv = torch.zeros((N, Z), dtype=torch.float32, device=q.device) if want_probs else None
    sc = scales.float().contiguous()
    ca, avg_sh, avg_st = _avg_arg(cumavg, N, H, T_DST, D)
    _lib.call('sea_sparse_attention_fwd', crow.data_ptr(), col.data_ptr(), _idx64(crow, col), Z,
              q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
              v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), sc.data_ptr(), _p(ca), avg_sh, avg_st, int(bool(use_scaler)), _dtype_code(q),
              out.data_ptr(), _p(pv), _p(head_ptr), N, H, T_DST, T_SRC, D, _stream())
    return out, pv


def sparse_attention_from_bits(bits, q, k, v, scales, cumavg, P: int, k_clamp: int, use_scaler=True, is_causal=True, kernel: str = 'auto'):
    """a8 + a9-a14 fused: attention driven directly by the top-k bit mask (no CSR tensors are materialised).

    kernel = 'auto' picks the tile-skipping block kernel (sea_block_attention_fwd) where the library supports the shape
    (d = 64, no clamped pixel, T_SRC <= 8192) and the per-(row, head) gather kernel otherwise; 'gather' / 'block' force one."""
    _cuda(bits, q, k, v, scales, cumavg)
    N, H, T_DST, D = q.shape
    T_SRC = k.shape[2]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    out = torch.empty((N, T_DST, H * D), dtype=q.dtype, device=q.device)
    sc = scales.float().contiguous()
    ca, avg_sh, avg_st = _avg_arg(cumavg, N, H, T_DST, D)
    if kernel not in ('auto', 'gather', 'block'):
        raise SeaError(f'unknown attention kernel {kernel!r}')
    if kernel == 'auto' and os.environ.get('SEA_ATTN_GATHER'):
        kernel = 'gather'             # development switch for A/B timing
    ws_bytes = 0
    if kernel != 'gather':
        ws_bytes = int(_lib.load().sea_block_attention_workspace_bytes(N, H, T_DST, T_SRC, D, int(P), int(k_clamp), _dtype_code(q)))
        if ws_bytes == 0 and kernel == 'block':
            raise SeaError('sparse_attention_from_bits: the block kernel does not support this shape')
    if ws_bytes > 0:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device)
        _lib.call('sea_block_attention_fwd', bits.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
                  k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
                  sc.data_ptr(), _p(ca), avg_sh, avg_st, int(bool(use_scaler)), _dtype_code(q), out.data_ptr(), N, H, T_DST, T_SRC, D, int(P),
                  int(k_clamp), int(bool(is_causal)), ws.data_ptr(), ws_bytes, _stream(), kernels=1 if ws_bytes <= 16 else 2)
        return out
    _lib.call('sea_sparse_attention_bits_fwd', bits.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
              k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
              sc.data_ptr(), _p(ca), avg_sh, avg_st, int(bool(use_scaler)), _dtype_code(q), out.data_ptr(), N, H, T_DST, T_SRC, D, int(P), int(k_clamp),
              int(bool(is_causal)), _stream())
    return out


def sparse_attention_from_bits_backward(bits, q, k, v, scales, cumavg, dout, P: int, k_clamp: int, use_scaler=True, is_causal=True):
    """Gradient of sparse_attention_from_bits w.r.t. (q, k, v, scales), restricted to the alive pairs of the bit mask
    (SURVEY 8f-1).  Returns fp32 (dq [N,H,T_DST,D], dk, dv [N,H,T_SRC,D], dscales [N,H,T_DST,2]).  The running-mean branch
    (cumavg) contributes to dv; the mask itself carries no gradient."""
    _cuda(bits, q, k, v, scales, dout)
    N, H, T_DST, D = q.shape
    T_SRC = k.shape[2]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    dout = dout.to(q.dtype).contiguous()
    sc = scales.float().contiguous()
    ca, avg_sh, avg_st = _avg_arg(cumavg, N, H, T_DST, D)
    dq = torch.empty((N, H, T_DST, D), dtype=torch.float32, device=q.device)
    dk = torch.empty((N, H, T_SRC, D), dtype=torch.float32, device=q.device)
    dv = torch.empty((N, H, T_SRC, D), dtype=torch.float32, device=q.device)
    dsc = torch.empty((N, H, T_DST, 2), dtype=torch.float32, device=q.device)
    _lib.call('sea_sparse_attention_bits_bwd', bits.data_ptr(), q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
              k.data_ptr(), k.stride(0), k.stride(1), k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
              sc.data_ptr(), _p(ca), avg_sh, avg_st, int(bool(use_scaler)), _dtype_code(q), dout.data_ptr(), dq.data_ptr(), dk.data_ptr(),
              dv.data_ptr(), dsc.data_ptr(), N, H, T_DST, T_SRC, D, int(P), int(k_clamp), int(bool(is_causal)), _stream())
    return dq, dk, dv, dsc


class SparseAttentionFromBits(torch.autograd.Function):
    """autograd wrapper: forward = sparse_attention_from_bits, backward = sea_sparse_attention_bits_bwd.  `cumavg` is
    recomputed as the running mean of v inside the graph by the caller when it must receive gradient; here it is treated
    as the forward's own function of v (the kernel adds its dv term)."""

    @staticmethod
    def forward(ctx, bits, q, k, v, scales, cumavg, P, k_clamp, use_scaler, is_causal):
        out = sparse_attention_from_bits(bits, q, k, v, scales, cumavg, P, k_clamp, use_scaler, is_causal)
        ctx.save_for_backward(bits, q, k, v, scales, cumavg)
        ctx.cfg = (P, k_clamp, use_scaler, is_causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        bits, q, k, v, scales, cumavg = ctx.saved_tensors
        P, k_clamp, use_scaler, is_causal = ctx.cfg
        dq, dk, dv, dsc = sparse_attention_from_bits_backward(bits, q, k, v, scales, cumavg, dout, P, k_clamp, use_scaler, is_causal)
        return None, dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), dsc.to(scales.dtype), None, None, None, None, None


def sparse_attention_from_bits_autograd(bits, q, k, v, scales, cumavg, P: int, k_clamp: int, use_scaler=True, is_causal=True):
    """Differentiable form of sparse_attention_from_bits (q, k, v, scales receive gradient)."""
    return SparseAttentionFromBits.apply(bits, q, k, v, scales, cumavg, int(P), int(k_clamp), bool(use_scaler), bool(is_causal))


def attention_bits_supported(dtype, D: int, P: int) -> bool:
    return dtype in (torch.bfloat16, torch.float16) and D in (32, 64, 80, 96, 128) and P % 32 == 0 and P <= 1024


# --------------------------------------------------------------------------------------------- non-causal (BERT)
def performer_noncausal(q, k, v, proj, lengths=None):
    """a2'+a3': cat(grid-sampled identity, v) + FAVOR+ Performer -> ctx [N,H,T,2D].  lengths: valid tokens per item (right-padded batch)."""
    _cuda(q, k, v, proj)
    N, H, T, D = q.shape
    F = proj.shape[0]
    q, k, v = _inner_contig(q), _inner_contig(k), _inner_contig(v)
    pj = proj.float().contiguous()
    ctx = torch.empty((N, H, T, 2 * D), dtype=q.dtype, device=q.device)
    ws = torch.empty((_lib.load().sea_performer_noncausal_workspace_floats(N, H, T, D, F),), dtype=torch.float32, device=q.device)
    ln = _lengths_arg(lengths, N, q.device)
    _lib.call('sea_performer_noncausal_len_fwd', q.data_ptr(), q.stride(0), q.stride(1), q.stride(2), k.data_ptr(), k.stride(0), k.stride(1),
              k.stride(2), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), pj.data_ptr(), _dtype_code(q), ctx.data_ptr(), ws.data_ptr(), _p(ln),
              N, H, T, D, F, _stream())
    return ctx


def conv3x3_cl(x, weight, bias, stride_t=1, up=1, relu=True):
    """Conv2d(C,O,3,padding=1,stride=(stride_t,1)) on channels-last x [N,Tin,W,C] (rows nearest-upsampled by `up` first)."""
    _cuda(x, weight, bias)
    x, weight, bias = _dense(x), _dense(weight, torch.float32), _dense(bias, torch.float32)
    N, Tin, W, C = x.shape
    O = weight.shape[0]
    Tout = (Tin * up + 2 - 3) // stride_t + 1
    y = torch.empty((N, Tout, W, O), dtype=x.dtype, device=x.device)
    _lib.call('sea_conv3x3_cl', x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), _dtype_code(x), N, Tin, Tout, W, C, O,
              int(stride_t), int(up), int(bool(relu)), _stream())
    return y


def bert_tail(y, T: int, P: int, want_scores=False):
    """bilinear resize of channels-last y [N,Tin,Win,H] to (T,P) + softmax(P) -> probs fp32 [N,H,T,P]."""
    _cuda(y)
    y = _dense(y)
    N, Tin, Win, H = y.shape
    probs = torch.empty((N, H, T, P), dtype=torch.float32, device=y.device)
    scores = torch.empty_like(probs) if want_scores else None
    _lib.call('sea_bert_tail_fwd', y.data_ptr(), _dtype_code(y), probs.data_ptr(), _p(scores), N, H, Tin, Win, T, P, _stream())
    return probs, scores


def topk_mask_bits_batch(probs, k_per_item, single_cta: bool = False, group_heads: int = None):
    """k_flatten_dim='batch' top-k: one group per item over the H*T*P keys (flat order of view(N, H*T*P)).
    group_heads=1: k_flatten_dim='head' (one group per (item, head), k_per_item has N*H entries)."""
    _cuda(probs, k_per_item)
    N, H, T, P = probs.shape
    pr = probs.float().contiguous()
    kp = k_per_item.reshape(-1).float().contiguous()
    bits = torch.empty((N, T, (H * P + 31) // 32), dtype=torch.int32, device=probs.device)
    gh = H if group_heads is None else int(group_heads)
    if single_cta:
        if gh != H:
            raise SeaError('the single-CTA top-k only implements the per-item group')
        _lib.call('sea_topk_mask_bits_batch', pr.data_ptr(), kp.data_ptr(), bits.data_ptr(), N, H, T, P, _stream())
        return bits
    nbytes = int(_lib.load().sea_topk_batch_workspace_bytes(N, H, T, P, gh))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=probs.device)
    _lib.call('sea_topk_mask_bits_batch_ws', pr.data_ptr(), kp.data_ptr(), bits.data_ptr(), ws.data_ptr(), nbytes, N, H, T, P, gh, _stream())
    return bits


def bert_avg(probs, v, lengths=None):
    """a13': probability-weighted mean of v -> [N,H,1,D] (dtype of v).  lengths: valid tokens per item (right-padded batch)."""
    _cuda(probs, v)
    N, H, T, P = probs.shape
    D = v.shape[-1]
    v = _inner_contig(v)
    avg = torch.empty((N, H, 1, D), dtype=v.dtype, device=v.device)
    ln = _lengths_arg(lengths, N, v.device)
    _lib.call('sea_bert_avg_len_fwd', probs.contiguous().data_ptr(), v.data_ptr(), v.stride(0), v.stride(1), v.stride(2), _dtype_code(v), avg.data_ptr(),
              _p(ln), N, H, T, P, D, _stream())
    return avg


# every public op runs on the device of its tensors (see _on_tensor_device)
for _name, _fn in list(globals().items()):
    if callable(_fn) and not _name.startswith('_') and getattr(_fn, '__module__', None) == __name__ and not isinstance(_fn, type):
        globals()[_name] = _on_tensor_device(_fn)
del _name, _fn
