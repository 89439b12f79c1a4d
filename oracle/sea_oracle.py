"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch-CPU float ops + numpy integer ops) of the
reference's per-layer SEA / Perlin attention forward.  It is the checker the CUDA path is compared
against; nothing in the product package imports it.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module.

Parity status
  * pinned against the reference ITSELF run in the build container (oracle/make_golden.py imports
    the unmodified /root/reference, dense torch path on CPU and the Triton kernels through Triton's
    numpy interpreter) -> fixtures in tests/golden/*.npz, checked by tests/test_oracle_golden.py;
  * pinned against the reference's own known-answer vectors: the per-row nnz table of
    src/poc/neko/visualize_ops_causal_resize.ipynb:29-35 and the CausalConv2d eye(8) table of
    src/poc/neko/test_causal_conv.ipynb:42-47,65;
  * Performer arithmetic (third-party performer-pytorch==1.1.4, not vendored): **unpinned** by any
    reference test; restated from the published algorithm, see third_party_restated/.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Tensor names follow the reference: N batch, H heads, T/T_DST/T_SRC tokens, P = T_M =
attention_predictor_length, k, F = performer features, d = head dim.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as TF


# ----------------------------------------------------------------------------- helpers
def fp_min_for(dtype) -> float:
    """attention.py:393-399."""
    if dtype in (torch.float16, torch.bfloat16):
        return torch.finfo(torch.float16).min / 2
    return torch.finfo(torch.float32).min / 2


def causal_additive_mask(T: int, dtype=torch.float32, N: int = 1) -> torch.Tensor:
    """test_perlin_opt_consist.py:112-113 / SURVEY 8d synthetic inputs."""
    m = (torch.arange(T).view(1, T) > torch.arange(T).view(T, 1)) * fp_min_for(dtype)
    return m.view(1, 1, T, T).to(dtype).expand(N, 1, T, T).contiguous()


def round_half_away(x: np.ndarray) -> np.ndarray:
    """libdevice roundf (causal_resize_m_to_t.py:232): half AWAY from zero, exact.
    Done in fp64 so that |x|+0.5 is exact for every fp32 input."""
    x64 = x.astype(np.float64)
    r = np.floor(np.abs(x64) + 0.5)
    return np.where(x64 < 0, -r, r).astype(np.float32)


# ----------------------------------------------------------------------------- a3 Performer
def performer_features_generalized(x: torch.Tensor, proj: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    """phi(x) = relu(d^-1/4 x P^T) + 1e-3  (performer-pytorch generalized_kernel; JAX twin
    performer_attention.py:163-198; ctor args attention.py:159-164)."""
    d = x.shape[-1]
    return torch.relu(torch.einsum('...id,jd->...ij', x * (d ** -0.25), proj.to(x.dtype))) + eps


def performer_causal(q, k, v, proj, eps: float = 1e-6, acc_dtype=torch.float64) -> torch.Tensor:
    """out_t = (phi(q_t) . S_t) / (phi(q_t) . (z_t + eps)),  S_t = sum_{s<=t} phi(k_s) (x) v_s,
    z_t = sum_{s<=t} phi(k_s)  (performer-pytorch causal_linear_attention*, JAX twin
    performer_attention.py:432-515; in-tree twin attention_state.py:80-98).  Prefix sums are
    accumulated in `acc_dtype` (the reference accumulates in the input dtype; fp64 here makes the
    oracle the better-conditioned side of the comparison)."""
    qf = performer_features_generalized(q.float(), proj.float())
    kf = performer_features_generalized(k.float(), proj.float())
    T = q.shape[-2]
    out = torch.empty(q.shape[:-1] + (v.shape[-1],), dtype=torch.float32)
    S = torch.zeros(q.shape[:-2] + (kf.shape[-1], v.shape[-1]), dtype=acc_dtype)
    z = torch.zeros(q.shape[:-2] + (kf.shape[-1],), dtype=acc_dtype)
    CH = 64
    for t0 in range(0, T, CH):
        t1 = min(T, t0 + CH)
        kc = kf[..., t0:t1, :].to(acc_dtype)
        qc = qf[..., t0:t1, :].to(acc_dtype)
        vc = v[..., t0:t1, :].to(acc_dtype)
        A = torch.tril(torch.einsum('...if,...jf->...ij', qc, kc))
        num = torch.einsum('...ij,...je->...ie', A, vc) + torch.einsum('...if,...fe->...ie', qc, S)
        zc = z.unsqueeze(-2) + kc.cumsum(-2)
        den = torch.einsum('...if,...if->...i', qc, zc + eps)
        out[..., t0:t1, :] = (num / den.unsqueeze(-1)).float()
        S = S + torch.einsum('...jf,...je->...fe', kc, vc)
        z = z + kc.sum(-2)
    return out


def performer_noncausal(q, k, v, proj, eps: float = 1e-4) -> torch.Tensor:
    """FAVOR+ softmax features + un-prefixed sums (performer-pytorch softmax_kernel /
    linear_attention; JAX twin performer_attention.py:50-108)."""
    d = q.shape[-1]
    F_ = proj.shape[0]
    norm = d ** -0.25
    ratio = F_ ** -0.5
    qd = torch.einsum('...id,jd->...ij', q.float() * norm, proj.float())
    kd = torch.einsum('...id,jd->...ij', k.float() * norm, proj.float())
    q_diag = (q.float() ** 2).sum(-1, keepdim=True) / 2.0 * norm ** 2
    k_diag = (k.float() ** 2).sum(-1, keepdim=True) / 2.0 * norm ** 2
    qp = ratio * (torch.exp(qd - q_diag - qd.amax(dim=-1, keepdim=True)) + eps)
    kp = ratio * (torch.exp(kd - k_diag - kd.amax(dim=(-1, -2), keepdim=True)) + eps)
    ksum = kp.sum(-2)
    d_inv = 1.0 / torch.einsum('...nf,...f->...n', qp, ksum)
    ctx = torch.einsum('...nf,...ne->...fe', kp, v.float())
    return torch.einsum('...fe,...nf,...n->...ne', ctx, qp, d_inv)


# ----------------------------------------------------------------------------- a4 predictor MLP
def layer_norm(x, w, b, eps: float = 1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def predictor_enc(performer_value, sd: Dict[str, torch.Tensor]):
    """attention.py:190-196,620: Linear(3d,2d) -> LayerNorm -> GELU(erf)."""
    y = performer_value @ sd['attention_predictor_enc.0.weight'].t() + sd['attention_predictor_enc.0.bias']
    y = layer_norm(y, sd['attention_predictor_enc.1.weight'], sd['attention_predictor_enc.1.bias'])
    return gelu_erf(y)


def predictor_dec_row(t_pred, sd, splits: int):
    """attention.py:242-245 (causal) / 203-206 (BERT) + ChannelSplit attention.py:123-131:
    [N,H,T,splits*W] -> [N,H*splits,T,W], channel = h*splits + s."""
    y = t_pred @ sd['attention_predictor_dec_row.0.weight'].t() + sd['attention_predictor_dec_row.0.bias']
    N, H, T, SW = y.shape
    W = SW // splits
    return y.view(N, H, T, splits, W).permute(0, 1, 3, 2, 4).reshape(N, H * splits, T, W)


def predictor_dec_scaler(t_pred, sd):
    """attention.py:289-291,1167."""
    return t_pred @ sd['attention_predictor_dec_scaler.0.weight'].t() + sd['attention_predictor_dec_scaler.0.bias']


# ----------------------------------------------------------------------------- a5 predictor CNN
def causal_conv2d(x, weight, weight_mask, bias, kernel_size: int, dilation: int, w_pad: int, stride: int = 1):
    """modules.py:96-192 restated tap by tap.  `weight` is [O, C, 2*ks-1, ks] with rows >= ks
    masked to zero (:131-142,173); padding = ((ks-1)*dil, w_pad) (:142); so (stride 1)
      y[o,t,w] = b[o] + sum_{i<ks} sum_{j<ks} W[o,c,i,j] * x[c, t-(ks-1)*dil+i*dil, w-w_pad+j*dil]
    (rows only at or before t -> causal along T).  Output width = W + 2*w_pad - (ks-1)*dil."""
    w = weight * (weight_mask != 0)
    N, C, T, W = x.shape
    O = w.shape[0]
    Wo = W + 2 * w_pad - (kernel_size - 1) * dilation
    xp = TF.pad(x, (w_pad, w_pad, (kernel_size - 1) * dilation, 0))
    y = bias.view(1, O, 1, 1).expand(N, O, T, Wo).clone()
    for i in range(kernel_size):
        for j in range(kernel_size):
            xs = xp[:, :, i * dilation:i * dilation + T, j * dilation:j * dilation + Wo]
            y = y + torch.einsum('oc,nctw->notw', w[:, :, i, j], xs)
    # The module pads (ks-1)*dil rows on BOTH sides (modules.py:142,177) and the 2ks-1 tall kernel
    # eats them again, so stride 1 gives exactly T rows; a strided conv (only the notebook KAT
    # uses one) is the stride-1 result subsampled.
    return y[:, :, ::stride, ::stride]


def area_resize_width(x, out_w: int):
    """modules.py:12-31 with mode='area' (chosen when shrinking, :15) == adaptive average pooling:
    out[j] = mean(x[floor(j*W/out_w) : ceil((j+1)*W/out_w)])."""
    W = x.shape[-1]
    if W == out_w:
        return x
    cols = []
    for j in range(out_w):
        s = (j * W) // out_w
        e = -((-(j + 1) * W) // out_w)
        cols.append(x[..., s:e].mean(-1))
    return torch.stack(cols, dim=-1)


def predictor_cnn_causal(x, sd):
    """attention.py:266-281: LN(P/4) -> [CausalConv2d(2H,2H,3,pad 2,dil 2)+ReLU]x2 -> nearest x4 on W
    -> CausalConv2d(2H,H,1,pad 1) (width P+2) -> area resize to P (KeepRes, modules.py:42-55) -> LN(P)."""
    p = 'attention_predictor_cnn.'
    y = layer_norm(x, sd[p + '0.module.weight'], sd[p + '0.module.bias'])
    # PERLIN_HOTFIX_OPT_DEEPER=1 (attention.py:246-263): a third dilated conv + ReLU; the 1x1 conv then sits at index 7 instead of 5
    deeper = (p + '1.module.net.7.module.weight') in sd
    for idx in (('0', '2', '4') if deeper else ('0', '2')):
        q = p + f'1.module.net.{idx}.module.'
        y = torch.relu(causal_conv2d(y, sd[q + 'weight'], sd[q + 'weight_mask'], sd[q + 'bias'], 3, 2, 2))
    y = y.repeat_interleave(4, dim=-1)                       # UpsampleFP32((1,4)) nearest, modules.py:77-92
    q = p + ('1.module.net.7.module.' if deeper else '1.module.net.5.module.')
    y = causal_conv2d(y, sd[q + 'weight'], sd[q + 'weight_mask'], sd[q + 'bias'], 1, 1, 1)
    y = area_resize_width(y, sd[p + '2.module.weight'].shape[0])
    return layer_norm(y, sd[p + '2.module.weight'], sd[p + '2.module.bias'])


def predictor_cnn_bert(x, sd, P: int):
    """attention.py:207-218: Conv2d(4H,4H,3,p1,stride(2,1))+ReLU -> Conv2d(4H,4H,3,p1)+ReLU ->
    nearest (2,1) -> Conv2d(4H,H,3,p1) -> KeepRes interpolate to (T, P) (bilinear when widening)."""
    p = 'attention_predictor_cnn.0.net.'
    T = x.shape[-2]
    y = torch.relu(TF.conv2d(x, sd[p + '0.weight'], sd[p + '0.bias'], stride=(2, 1), padding=1))
    y = torch.relu(TF.conv2d(y, sd[p + '2.weight'], sd[p + '2.bias'], padding=1))
    y = TF.interpolate(y, scale_factor=(2, 1), mode='nearest')
    y = TF.conv2d(y, sd[p + '5.weight'], sd[p + '5.bias'], padding=1)
    if y.shape[-2:] != (T, P):
        mode = 'bilinear' if P >= y.shape[-1] else 'area'
        y = TF.interpolate(y, (T, P), mode=mode)
    return y


# ----------------------------------------------------------------------------- a7 grouped top-k
def per_item_top_k_causal(H: int, k: float, k_oversample: float, P: int, T: int) -> np.ndarray:
    """attention.py:849,856,866: K_t = max(round_half_even(H * (k*os*P / (t+1))), 1), computed by
    torch in fp32 (python float k*os*P -> fp32 scalar; int64 arange promoted to fp32)."""
    tl_ = torch.arange(1, T + 1, dtype=torch.long)
    kt = H * ((k * k_oversample * P) / tl_)
    kt = torch.clamp_min(torch.round(kt), 1)
    assert kt.dtype == torch.float32
    return kt.numpy().astype(np.float32)


def topk_alive_rows(keys: np.ndarray, K: np.ndarray) -> np.ndarray:
    """attention.py:871-917 for one flattened group per row: alive = the K[row] largest keys, rank by
    descending value.  The reference's sort is unstable (ties implementation-defined, :880-885); the
    contract here (SURVEY 8d parity gates) is: among equal keys the LOWER flat index wins.
    keys [R, G] fp32, K [R] -> bool [R, G]."""
    R, G = keys.shape
    order = np.argsort(-keys, axis=-1, kind='stable')
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.broadcast_to(np.arange(G), (R, G)), axis=-1)
    return rank < K.reshape(R, 1)


def topk_mask_causal_batch(probs: torch.Tensor, k, k_oversample: float = 1.0,
                           dst_valid: Optional[torch.Tensor] = None, floor_variant: bool = False) -> torch.Tensor:
    """attention.py:816-849 + 871-947 (`causal_batch`, causal): group = all heads of one query row,
    flat index h*P+m.  Returns the 0/1 float mask [N,H,T,P] of the benchmarking branch (:916-931).
    floor_variant=True is the kernel-test helper causal_topk_masking.py:31."""
    N, H, T, P = probs.shape
    t = probs
    if dst_valid is not None:
        t = t * dst_valid.view(N, 1, T, 1).to(t.dtype)
    keys = t.detach().transpose(1, 2).reshape(N * T, H * P).float().numpy()       # (the mask carries no gradient)
    if floor_variant:
        tl_ = torch.arange(1, T + 1, dtype=torch.long)
        K = torch.clamp(H * torch.floor(k * P / tl_), 1, H * P).numpy().astype(np.float32)
    else:
        K = per_item_top_k_causal(H, k, k_oversample, P, T)
    K = np.tile(K, N)
    alive = topk_alive_rows(keys, K)
    m = torch.from_numpy(alive.reshape(N, T, H, P)).permute(0, 2, 1, 3).float()
    if dst_valid is not None:
        m = m * dst_valid.view(N, 1, T, 1)
    return m.contiguous()


def topk_mask_noncausal(probs: torch.Tensor, k, k_oversample: float, token_length: torch.Tensor,
                        k_flatten_dim: str, dst_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """attention.py:833-853 non-causal groupings: 'batch' (one group per item, H*T*P keys),
    'head' (per head, T*P), 'query' (per (h,t), P), 'causal_batch' (per t, H*P)."""
    N, H, T, P = probs.shape
    t = probs if dst_valid is None else probs * dst_valid.view(N, 1, T, 1).to(probs.dtype)
    tl_ = token_length.view(N).long()
    if k_flatten_dim == 'batch':
        keys = t.reshape(N, H * T * P)
        K = tl_ * H * (k * k_oversample * P / tl_)
    elif k_flatten_dim == 'head':
        keys = t.reshape(N * H, T * P)
        K = (tl_ * (k * k_oversample * P / tl_)).view(N, 1).expand(N, H).reshape(-1)
    elif k_flatten_dim == 'causal_batch':
        keys = t.transpose(1, 2).reshape(N * T, H * P)
        K = (H * (k * k_oversample * P / tl_)).view(N, 1).expand(N, T).reshape(-1)
    elif k_flatten_dim == 'query':
        keys = t.reshape(N * H * T, P)
        K = (k * k_oversample * P / tl_).view(N, 1).expand(N, H * T).reshape(-1)
    else:
        raise ValueError(k_flatten_dim)
    K = torch.clamp_min(torch.round(K.float()), 1).numpy().astype(np.float32)
    alive = topk_alive_rows(keys.float().numpy(), K)
    if k_flatten_dim == 'causal_batch':
        m = torch.from_numpy(alive.reshape(N, T, H, P)).permute(0, 2, 1, 3).float()
    else:
        m = torch.from_numpy(alive.reshape(N, H, T, P)).float()
    if dst_valid is not None and k_flatten_dim in ('causal_batch', 'query'):
        m = m * dst_valid.view(N, 1, T, 1)
    return m.contiguous()


# ----------------------------------------------------------------------------- a8 CSR interpolation
def resize_from_m_to_t_csr(mask: torch.Tensor, k: int, target_width: Optional[int] = None, is_causal: bool = True,
                           scale_mode: str = 'ieee', noncausal_width: Optional[int] = None):
    """causal_resize_m_to_t.py:910-1007 -> scan_col METHOD 1 (:648-762) -> __scan_col_4_compute (:493-572).
    mask [N,H,T_DST,P] 0/1.  Returns (crow [N,T_DST+1] i64, col [N,Z] i64, Z) with Z = max_n nnz_n;
    rows of items with fewer nnz are zero padded at the tail (:669).  Entry order: (t, h, m) then
    descending column inside a pixel (:561-572).

    scale_mode: how `scales = target_width / original_width` (:642, int64 tensor / python int) is evaluated.  'ieee' = fp32(L) / fp32(P),
    what torch does on the CPU (and what the CUDA kernels of this repo do).  'cuda_reciprocal' = fp32(L) * (1.0f / fp32(P)), what
    torch's CUDA true-division kernel does when the divisor is a host scalar (ATen BinaryDivTrueKernel.cu) -- i.e. the reference
    run on a GPU.  The two agree whenever P is a power of two (every shipped config); for other P a few pixel edges move by one."""
    N, H, T_DST, P = mask.shape
    T_SRC = target_width if target_width is not None else T_DST
    if is_causal:
        tw = np.arange(1, T_SRC + 1, dtype=np.int64)[-T_DST:]          # :954
    else:
        # :957 uses T_SRC for every row ("TODO confirm correctness" there: the reference's sparse non-causal path ignores padding in
        # the interpolation).  noncausal_width = the item's token length gives the width its DENSE path uses (resize_m_to_t.py:36-47,
        # the per-item mask cumsum) -- what a padded non-causal batch needs; None keeps the reference's sparse behaviour.
        tw = np.full((T_SRC,), T_SRC if noncausal_width is None else int(noncausal_width), dtype=np.int64)[-T_DST:]
    if scale_mode == 'ieee':
        scales = (torch.from_numpy(tw) / P).numpy()                    # :642 int64 / int -> fp32
    elif scale_mode == 'cuda_reciprocal':
        scales = (tw.astype(np.float32) * (np.float32(1.0) / np.float32(P))).astype(np.float32)
    else:
        raise ValueError(scale_mode)
    assert scales.dtype == np.float32
    b = np.arange(P, dtype=np.int64).reshape(1, P)
    vs = round_half_away((b.astype(np.float32) * scales.reshape(T_DST, 1)).astype(np.float32))        # :654
    ve = round_half_away(((b + 1).astype(np.float32) * scales.reshape(T_DST, 1)).astype(np.float32))  # :655
    x = mask.transpose(1, 2).reshape(N, T_DST, H, P).numpy()
    npx = (ve - vs).reshape(1, T_DST, 1, P).astype(np.int32) * x.astype(np.int32)                     # :657
    npx = np.minimum(npx, k)                                                                           # :659
    pix = npx.reshape(N, -1).astype(np.int64).cumsum(-1)                                              # :664
    row_end = pix.reshape(N, T_DST, -1)[:, :, -1]
    Z = int(row_end.max()) if row_end.size else 0                                                      # :667
    crow = np.zeros((N, T_DST + 1), dtype=np.int64)
    crow[:, 1:] = row_end                                                                              # :672
    col = np.zeros((N, Z), dtype=np.int64)
    vs_i = vs.astype(np.int64)
    ve_i = ve.astype(np.int64)
    for n in range(N):
        flat = npx[n].reshape(-1)
        nz = np.nonzero(flat)[0]                                                                       # :724
        if nz.size == 0:
            continue
        t_idx = nz // (H * P)
        h_idx = (nz % (H * P)) // P
        m_idx = nz % P
        cl = flat[nz].astype(np.int64)
        start = pix[n][nz] - cl
        rs = vs[t_idx, m_idx]
        re = ve[t_idx, m_idx]
        # value = range_end - int(j * ((range_end - range_start) / col_len)) - 1  (:569), fp32 math;
        # Triton lowers `/` to div.full.f32 (approximate) -- IEEE division is used here; the two can
        # only differ when col_len < v_end - v_start, i.e. the clamp to k is active (T/P > k).
        ratio = ((re - rs) / cl.astype(np.float32)).astype(np.float32)
        maxw = int(cl.max())
        j = np.arange(maxw, dtype=np.float32).reshape(1, maxw)
        off = (j * ratio.reshape(-1, 1)).astype(np.float32).astype(np.int64)   # fp->int truncation
        vals = (h_idx * T_SRC + ve_i[t_idx, m_idx]).reshape(-1, 1) - off - 1
        valid = np.arange(maxw).reshape(1, maxw) < cl.reshape(-1, 1)
        pos = start.reshape(-1, 1) + np.arange(maxw).reshape(1, maxw)
        col[n, pos[valid]] = vals[valid]
    return torch.from_numpy(crow), torch.from_numpy(col), Z


def flat_csr_to_dense(crow: torch.Tensor, col: torch.Tensor, values: torch.Tensor, T_SRC: int, H: int) -> torch.Tensor:
    """flat_csr_to_dense.py:3-36 -> [N,H,T_DST,T_SRC]."""
    N, R1 = crow.shape
    T_DST = R1 - 1
    out = torch.zeros((N, H, T_DST, T_SRC), dtype=values.dtype)
    crow_n = crow.numpy()
    for n in range(N):
        nnz = int(crow_n[n, -1])
        if nnz == 0:
            continue
        rows = np.repeat(np.arange(T_DST), np.diff(crow_n[n]))
        c = col[n, :nnz].numpy()
        out[n, torch.from_numpy(c // T_SRC), torch.from_numpy(rows), torch.from_numpy(c % T_SRC)] = values[n, :nnz]
    return out


# ----------------------------------------------------------------------------- a16 dense resize
def resize_from_m_to_t_dense(x, masked_fill_value, attention_mask, target_width=None, is_causal=True,
                             k=None, oversampled=None):
    """resize_m_to_t.py:6-73 (training=False): column j of a row with L valid tokens reads pixel
    floor((j+0.5)/L*P - 1e-4); invalid columns read the pad pixel (fill)."""
    N, H, T1, P = x.shape
    T2 = target_width if target_width is not None else T1
    if not is_causal:
        attention_mask = attention_mask.expand(N, 1, T1, T2)
    mask = (attention_mask > -1).float()
    mask_cs = mask.cumsum(-1)
    token_length = mask_cs[:, :, :, -1].unsqueeze(-1)
    idx = torch.floor(((mask_cs - 1) + 0.5) / token_length * P - 1e-4).to(torch.long) + ((1 - mask) * P).to(torch.long)
    idx = torch.clamp(idx, 0, P).expand(N, H, T1, T2)
    out = TF.pad(x, pad=(0, 1), value=masked_fill_value).gather(dim=-1, index=idx)
    if oversampled is not None:
        xs = torch.arange(0, T2).view(1, 1, 1, T2)
        ws = token_length
        ps = torch.clamp_min(torch.round(token_length / oversampled), 1)
        oys = torch.clamp(token_length, round(k), round(k * oversampled)) / k
        keep = torch.abs(((xs + 1) / ws * ps) - torch.round((xs + 1) / ws * ps)) <= ((1 / oys) * 0.5 + 1e-4)
        out = out.masked_fill(~keep, masked_fill_value)
    return out


# ----------------------------------------------------------------------------- a9-a12 flat-CSR ops
def _rows_of(crow_n: np.ndarray) -> np.ndarray:
    return np.repeat(np.arange(crow_n.shape[0] - 1), np.diff(crow_n))


def flat_csr_masked_bmm(q, k, crow, col):
    """flat_csr_masked_bmm.py:39-125: score[z] = <Q[n,h,row], K[n,h,col]>, fp32, UNSCALED."""
    N, H, T_DST, d = q.shape
    T_SRC = k.shape[2]
    vals = torch.zeros(col.shape, dtype=torch.float32)
    for n in range(N):
        cn = crow[n].numpy()
        nnz = int(cn[-1])
        rows = torch.from_numpy(_rows_of(cn))
        c = col[n, :nnz]
        h = c // T_SRC
        j = c % T_SRC
        vals[n, :nnz] = (q[n, h, rows].float() * k[n, h, j].float()).sum(-1)
    return vals


def flat_csr_softmax(vals, crow, col, H, T_SRC):
    """flat_csr_softmax.py:55-125: softmax inside every (row, head) group of entries."""
    out = torch.zeros_like(vals)
    for n in range(vals.shape[0]):
        cn = crow[n].numpy()
        nnz = int(cn[-1])
        if nnz == 0:
            continue
        rows = torch.from_numpy(_rows_of(cn))
        g = rows * H + col[n, :nnz] // T_SRC
        G = (crow.shape[1] - 1) * H
        v = vals[n, :nnz].double()
        gmax = torch.full((G,), -float('inf'), dtype=torch.float64).scatter_reduce(0, g, v, 'amax')
        e = torch.exp(v - gmax[g])
        s = torch.zeros((G,), dtype=torch.float64).index_add_(0, g, e)
        out[n, :nnz] = (e / s[g]).float()
    return out


def flat_csr_elmul_rowscale(vals, crow, col, row_scaler, T_SRC):
    """flat_csr_elmul.py:42-108 with the stride-0 [N,H,T,T] expansion of attention.py:1170:
    p[z] *= row_scaler[n, h, row]."""
    out = torch.zeros_like(vals)
    for n in range(vals.shape[0]):
        cn = crow[n].numpy()
        nnz = int(cn[-1])
        rows = torch.from_numpy(_rows_of(cn))
        h = col[n, :nnz] // T_SRC
        out[n, :nnz] = vals[n, :nnz] * row_scaler[n, h, rows].float()
    return out


def flat_csr_sdbmm(vals, crow, col, v, H):
    """flat_csr_sdbmm.py:141-313 WITHOUT its MAX_ROW_T truncation (:382-388): out[n,h,row] = sum_z p[z] V[n,h,col]."""
    N, _, T_SRC, d = v.shape
    T_DST = crow.shape[1] - 1
    out = torch.zeros((N, H, T_DST, d), dtype=torch.float32)
    for n in range(N):
        cn = crow[n].numpy()
        nnz = int(cn[-1])
        rows = torch.from_numpy(_rows_of(cn))
        c = col[n, :nnz]
        h = c // T_SRC
        j = c % T_SRC
        contrib = vals[n, :nnz].unsqueeze(-1).double() * v[n, h, j].double()
        flat = torch.zeros((H * T_DST, d), dtype=torch.float64).index_add_(0, h * T_DST + rows, contrib)
        out[n] = flat.view(H, T_DST, d).float()
    return out


# ----------------------------------------------------------------------------- whole layer
def state_dict_of(module) -> Dict[str, torch.Tensor]:
    return {k_: v_.detach().float().cpu() for k_, v_ in module.state_dict().items()}


def perlin_forward_causal(sd: Dict[str, torch.Tensor], q, k, v, *, k_top: int, P: int,
                          k_oversample: float = 1.0, partial_attention_scaler: bool = True,
                          sparse: bool = True, dst_valid: Optional[torch.Tensor] = None,
                          keep_dense: bool = False, query_skips: int = 1, mask_override: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """attention.py:333-1359, causal prefill, `context_output_method='mix'`.
    sparse=True follows the `benchmarking` branch (CSR mask + flat_csr ops, :1036-1042, :1151-1173);
    sparse=False follows the dense branch (:960-962, :1066-1133) the reference itself runs on CPU.
    Returns every stage tensor under the reference's temp-buffer names."""
    q, k, v = q.float(), k.float(), v.float()
    N, H, T, d = q.shape
    buf: Dict[str, torch.Tensor] = {}
    # a2 vmask.cat_fill (:504-508)
    pos = sd['v_eye_learned_causal'][:, :, :T, :].expand(N, H, T, d)
    v_for_atten = torch.cat([pos, v], dim=-1)
    if dst_valid is not None:
        v_for_atten = v_for_atten * dst_valid.view(N, 1, T, 1)
        v = v * dst_valid.view(N, 1, T, 1)
    # a3
    pcl = performer_causal(q, k, v_for_atten, sd['performer.projection_matrix'])
    buf['performer_context_layer'] = pcl
    # a4
    pv = torch.cat([pcl, v], dim=-1)
    if query_skips > 1:                                                                          # :617-619
        assert T % query_skips == 0
        pv = pv[:, :, ::query_skips, :]
    t_pred = predictor_enc(pv, sd)
    dec = predictor_dec_row(t_pred, sd, splits=2)
    # a5 + a6
    score = predictor_cnn_causal(dec, sd)
    if query_skips > 1:                                                                          # :640-644: every result repeated
        score = score.repeat_interleave(query_skips, dim=2)
        t_pred = t_pred.repeat_interleave(query_skips, dim=2)
    buf['t_attention_predictor'] = t_pred
    buf['estimated_attention_score'] = score
    probs = torch.softmax(score, dim=-1)
    buf['estimated_attention_probs'] = probs
    # a7
    # mask_override: a given top-k selection [N,H,T,P] (e.g. the reference's own, whose unstable CPU sort cuts exact ties arbitrarily)
    mask_m = topk_mask_causal_batch(probs, k_top, k_oversample, dst_valid) if mask_override is None else mask_override.float()
    buf['partial_attention_mask_before_interp'] = mask_m
    scales = predictor_dec_scaler(t_pred, sd)
    buf['estimated_scales'] = scales
    if sparse:
        crow, col, Z = resize_from_m_to_t_csr(mask_m, k_top, target_width=T, is_causal=True)
        buf['crow_indices'], buf['col_indices'] = crow, col
        s = flat_csr_masked_bmm(q, k, crow, col)
        p = flat_csr_softmax(s, crow, col, H, T)
        if partial_attention_scaler:
            p = flat_csr_elmul_rowscale(p, crow, col, torch.sigmoid(scales[..., 0]), T)
        buf['partial_attention_probs_values'] = p
        ctx = flat_csr_sdbmm(p, crow, col, v, H)
        if keep_dense:
            buf['partial_attention_mask'] = flat_csr_to_dense(crow, col, torch.ones_like(p), T, H)
    else:
        fmin = fp_min_for(torch.float32)
        cmask = causal_additive_mask(T, torch.float32, N)
        if dst_valid is not None:
            cmask = cmask.masked_fill((dst_valid.view(N, 1, 1, T) == 0), fmin)
        pm = (1.0 - mask_m) * fmin
        pm = resize_from_m_to_t_dense(pm, fmin, cmask, target_width=T, is_causal=True, k=k_top, oversampled=k_oversample)
        pm = pm.masked_fill(cmask < -1, fmin)
        if keep_dense:
            buf['partial_attention_mask'] = (pm > -1).float()
        s = q @ k.transpose(-1, -2) + pm
        p = torch.softmax(s, dim=-1).masked_fill(pm < -1, 0)
        if partial_attention_scaler:
            p = p * torch.sigmoid(scales[..., 0:1])
        ctx = p @ v
    buf['partial_context_layer_1'] = ctx
    # a13 (:1237-1244)
    avg = v.cumsum(-2) / torch.arange(1, T + 1).view(1, 1, T, 1)
    buf['average_context_layer'] = avg
    a = torch.sigmoid(scales[..., 1:2])
    out = ctx * a + (1 - a) * avg
    # a14 (:1279-1282)
    buf['context_layer'] = out.permute(0, 2, 1, 3).reshape(N, T, H * d).contiguous()
    return buf


def perlin_train_forward(sd: Dict[str, torch.Tensor], q, k, v, scores_truth, context_truth, *, k_top: int, P: int,
                         partial_attention_scaler: bool = True, mask_override: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """TRAINING branch, causal, no padding, the 10 % resize jitter off (attention.py:680-765 predictor distillation on the resized
    estimated scores, :1066-1102 the same two terms on the dense q.k^T scores, :1328-1332 MSE of the context).  Built on the dense
    branch of perlin_forward_causal, so it is differentiable by autograd w.r.t. q, k, v and every tensor of `sd` that requires grad:
    the gradient reference of tests/test_training_*.py (the unmodified reference pins the forward values, see make_golden.py)."""
    buf = perlin_forward_causal(sd, q, k, v, k_top=k_top, P=P, sparse=False, keep_dense=True, partial_attention_scaler=partial_attention_scaler,
                                mask_override=mask_override)
    N, H, T, d = q.shape
    fmin = fp_min_for(torch.float32)
    cmask = causal_additive_mask(T, torch.float32, N)
    dead = cmask < -1
    score, probs = buf['estimated_attention_score'], buf['estimated_attention_probs']
    # handle_oversample=False -> oversampled = 1.0 (:691-703): the under-sampling test of resize_m_to_t.py:54-71 keeps every column
    est_probs_resized = resize_from_m_to_t_dense(probs, 0.0, cmask, target_width=T, is_causal=True, k=k_top, oversampled=1.0)
    est_score_resized = resize_from_m_to_t_dense(score, fmin, cmask, target_width=T, is_causal=True, k=k_top, oversampled=1.0)

    def kd(scores):
        inp = TF.log_softmax(scores.masked_fill(dead, fmin), dim=-1).reshape(-1, T)                       # :743, :1087
        target = TF.softmax(scores_truth.float().masked_fill(dead, fmin), dim=-1).reshape(-1, T)           # :745, :1089
        return TF.kl_div(inp, target, reduction='batchmean') * 0.1 + \
            TF.mse_loss(TF.softmax(scores.masked_fill(dead, fmin), dim=-1).reshape(-1, T), target)        # :747-758, :1091-1101

    loss = kd(est_score_resized)
    dense = q.float() @ k.float().transpose(-1, -2)
    loss = loss + kd(dense)
    loss = loss + TF.mse_loss(context_truth.float(), buf['context_layer'])                                 # :1328-1332
    buf['loss'] = loss
    buf['estimated_attention_probs_resized'] = est_probs_resized
    buf['dense_attention_probs'] = torch.softmax(dense.masked_fill(dead, fmin) + cmask, dim=-1)             # :1111-1115
    return buf


# ----------------------------------------------------------------------------- non-causal (BERT) layer
def v_identity_grid(N: int, H: int, T: int, d: int, valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """attention.py:462-495: bilinear grid_sample (align_corners=True) of the d x d identity at x = channel c,
    y = (rank of the token among the valid tokens) / (count - 1) * (d - 1)  ->  [N,H,T,d]; in closed form the sample
    is the hat function max(0, 1 - |y - c|)."""
    if valid is None:
        valid = torch.ones(N, T)
    cs = valid.float().cumsum(-1)
    tot = valid.float().sum(-1, keepdim=True)
    y_norm = (cs - 1.0) / ((tot - 1.0) + 1e-8) * 2 - 1
    ypix = (y_norm + 1) / 2 * (d - 1)                                   # align_corners=True un-normalisation
    c = torch.arange(d, dtype=torch.float32).view(1, 1, d)
    w = torch.clamp(1.0 - (ypix.view(N, T, 1) - c).abs(), min=0.0)
    inside = ((ypix >= 0) & (ypix <= d - 1)).view(N, T, 1)             # zero padding outside the grid
    return (w * inside).view(N, 1, T, d).expand(N, H, T, d).contiguous()


def perlin_forward_noncausal(sd: Dict[str, torch.Tensor], q, k, v, *, k_top: int, P: int, k_flatten_dim: str = 'batch',
                             k_oversample: float = 1.0, partial_attention_scaler: bool = True, sparse: bool = True,
                             keep_dense: bool = False, lengths: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """attention.py:333-1359 with `causal=False` (BERT), context_output_method='mix'.  lengths [N] (optional): right-padded batch, item n
    has lengths[n] real tokens (attention_mask [N,1,1,T] <= -1 on the rest): v and v_for_atten are zeroed on padded tokens (:512-514), the
    identity grid follows the valid-token rank (:482), the probabilities of padded query rows are zeroed (:777-778), per_item_top_k uses
    the token length (:837), the interpolation width is the token length (dense path, resize_m_to_t.py:36-47), the average context
    skips padded tokens (:1209-1219).  Rows of padded queries are not meaningful (the reference leaves them to the caller's masking)."""
    q, k, v = q.float(), k.float(), v.float()
    N, H, T, d = q.shape
    buf: Dict[str, torch.Tensor] = {}
    valid = None
    if lengths is not None:
        valid = (torch.arange(T).view(1, T) < lengths.view(N, 1)).float()
        v = v * valid.view(N, 1, T, 1)
    v_for_atten = torch.cat([v_identity_grid(N, H, T, d, valid), v], dim=-1)                    # :462-502
    if valid is not None:
        v_for_atten = v_for_atten * valid.view(N, 1, T, 1)                                       # :512-514
    pcl = performer_noncausal(q, k, v_for_atten, sd['performer.projection_matrix'])              # :527-534
    buf['performer_context_layer'] = pcl
    t_pred = predictor_enc(torch.cat([pcl, v], dim=-1), sd)                                      # :577-620
    buf['t_attention_predictor'] = t_pred
    dec = predictor_dec_row(t_pred, sd, splits=4)                                                # :203-206
    score = predictor_cnn_bert(dec, sd, P)                                                       # :207-218
    buf['estimated_attention_score'] = score
    probs = torch.softmax(score, dim=-1)                                                         # :670-673
    buf['estimated_attention_probs'] = probs
    token_length = torch.full((N,), T, dtype=torch.long) if lengths is None else lengths.long()
    if valid is not None:
        probs = probs * valid.view(N, 1, T, 1)                                                   # :777-778
    mask_m = topk_mask_noncausal(probs, k_top, k_oversample, token_length, k_flatten_dim,
                                 dst_valid=None if valid is None else valid)                     # :833-947
    buf['partial_attention_mask_before_interp'] = mask_m
    scales = predictor_dec_scaler(t_pred, sd)
    buf['estimated_scales'] = scales
    fmin = fp_min_for(torch.float32)
    amask = torch.zeros(N, 1, 1, T) if valid is None else ((1.0 - valid) * fmin).view(N, 1, 1, T)
    if sparse:
        if lengths is None:
            crow, col, Z = resize_from_m_to_t_csr(mask_m, k_top, target_width=T, is_causal=False)    # :1025-1027
        else:       # per-item interpolation width = token length (see resize_from_m_to_t_csr)
            per = [resize_from_m_to_t_csr(mask_m[n:n + 1], k_top, target_width=T, is_causal=False, noncausal_width=int(lengths[n])) for n in range(N)]
            Z = max(p_[2] for p_ in per)
            crow = torch.cat([p_[0] for p_ in per], dim=0)
            col = torch.zeros((N, Z), dtype=torch.long)
            for n, p_ in enumerate(per):
                col[n, :p_[2]] = p_[1][0]
        buf['crow_indices'], buf['col_indices'] = crow, col
        s = flat_csr_masked_bmm(q, k, crow, col)
        p = flat_csr_softmax(s, crow, col, H, T)
        if partial_attention_scaler:
            p = flat_csr_elmul_rowscale(p, crow, col, torch.sigmoid(scales[..., 0]), T)
        buf['partial_attention_probs_values'] = p
        ctx = flat_csr_sdbmm(p, crow, col, v, H)
        if keep_dense:
            buf['partial_attention_mask'] = flat_csr_to_dense(crow, col, torch.ones_like(p), T, H)
    else:
        pm = resize_from_m_to_t_dense((1.0 - mask_m) * fmin, fmin, amask, target_width=T, is_causal=False, k=k_top, oversampled=k_oversample)
        if keep_dense:
            buf['partial_attention_mask'] = (pm > -1).float()
        s = q @ k.transpose(-1, -2) + pm
        p = torch.softmax(s, dim=-1).masked_fill(pm < -1, 0)
        if partial_attention_scaler:
            p = p * torch.sigmoid(scales[..., 0:1])
        ctx = p @ v
    buf['partial_context_layer_1'] = ctx
    # :1209-1219: probability-weighted mean of v (v is already zero on padded tokens)
    wts = resize_from_m_to_t_dense(probs.mean(-2, keepdim=True), 0.0, amask, target_width=T, is_causal=False)   # [N,H,1,T]
    avg = (v * wts.transpose(-1, -2)).sum(-2, keepdim=True)
    buf['average_context_layer'] = avg
    a = torch.sigmoid(scales[..., 1:2])
    out = ctx * a + (1 - a) * avg
    buf['context_layer'] = out.permute(0, 2, 1, 3).reshape(N, T, H * d).contiguous()
    return buf


def sparse_attention_grads(alive: torch.Tensor, q, k, v, scales, dout, use_scaler: bool = True, with_avg: bool = True):
    """Gradient of the masked attention + scaler + running-mean mix w.r.t. (q, k, v, scales), by autograd in fp64.
    alive: bool [N,H,T_DST,T_SRC] (densified partial_attention_mask).  Follows the reference's dense training expression
    (attention.py:1066-1133 masked softmax, :1166-1171 scaler, :1237-1244 running mean) restricted to the mask -- the
    parity target of sea_sparse_attention_bits_bwd.  dout: [N,T_DST,H*D].  Returns (out, dq, dk, dv, dscales)."""
    N, H, T_DST, D = q.shape
    T_SRC = k.shape[2]
    q_, k_, v_, s_ = (t.detach().double().requires_grad_(True) for t in (q, k, v, scales))
    scores = torch.einsum('nhtd,nhsd->nhts', q_, k_).masked_fill(~alive, float('-inf'))
    has_any = alive.any(-1, keepdim=True)
    probs = torch.softmax(scores.masked_fill(~has_any, 0.0), dim=-1) * alive
    ctx = torch.einsum('nhts,nhsd->nhtd', probs, v_)
    if use_scaler:
        ctx = ctx * torch.sigmoid(s_[..., 0:1])
    if with_avg:
        assert T_SRC == T_DST
        avg = torch.cumsum(v_, dim=2) / torch.arange(1, T_SRC + 1, dtype=torch.float64).view(1, 1, -1, 1)
        a = torch.sigmoid(s_[..., 1:2])
        ctx = ctx * a + (1 - a) * avg
    out = ctx.permute(0, 2, 1, 3).reshape(N, T_DST, H * D)
    out.backward(dout.double())
    return out.detach().float(), q_.grad.float(), k_.grad.float(), v_.grad.float(), s_.grad.float()


# =====================================================================================================================
# a17: the stateful decode ops of attention_state.py, restated (pinned by tests/golden/state_ops.npz, which the unmodified
# reference classes produced -- oracle/make_golden.py::golden_state_ops)
# =====================================================================================================================
class StatefulCausalPerformerOracle:
    """attention_state.py:43-98.  Linear-attention recurrence on the q / k it is handed: running sums in fp64 (:87,:91), cast
    back to the input dtype before every contraction, eps 1e-12 on the normaliser (:89).  The new tokens are cut with
    torch.chunk(min(T_new, 16)) (:84-86) -- i.e. into that many pieces, not pieces of that size."""

    def __init__(self):
        self.k_cumsum = 0
        self.context_cumsum = 0

    def __call__(self, q, k, v):
        T_new = q.shape[-2]
        n_chunks = min(T_new, 16)
        outs = []
        for qc, kc, vc in zip(q.chunk(n_chunks, dim=-2), k[..., -T_new:, :].chunk(n_chunks, dim=-2), v[..., -T_new:, :].chunk(n_chunks, dim=-2)):
            k_cum = self.k_cumsum + kc.cumsum(dim=-2, dtype=torch.float64)
            d_inv = 1.0 / torch.einsum('...nd,...nd->...n', qc, k_cum.type_as(qc) + 1e-12)
            ctx = torch.einsum('...nd,...ne->...nde', kc, vc)
            ctx_cum = self.context_cumsum + ctx.cumsum(dim=-3, dtype=torch.float64)
            outs.append(torch.einsum('...nde,...nd,...n->...ne', ctx_cum.type_as(qc), qc, d_inv))
            self.k_cumsum = k_cum[:, :, -1:]
            self.context_cumsum = ctx_cum[:, :, -1:]
        return torch.cat(outs, dim=-2)


class StatefulCumAvgOracle:
    """attention_state.py:205-224: running mean of v over the absolute position, advanced by q_len tokens per call."""

    def __init__(self):
        self.cumsum = 0
        self.prev_len = 0

    def __call__(self, v, q_len: int):
        cs = v[..., -q_len:, :].cumsum(dim=-2) + self.cumsum
        self.cumsum = cs[..., -1:, :].clone()
        cs = cs / torch.arange(self.prev_len + 1, self.prev_len + 1 + cs.shape[-2]).view(1, 1, -1, 1)
        self.prev_len += q_len
        return cs


class StatefulCausalCNNOracle:
    """attention_state.py:142-187: the CNN is re-run on a window of the most recent rows -- at least max(T_new, 24) of them,
    extended back to the start of the oldest stored piece -- and the last T_new output rows are kept.  Exact as long as the
    CNN's causal receptive field (8 rows for the predictor's two dilated 3x3 convs) fits into the window."""

    def __init__(self, window_size: int = 24):
        self.window_size = window_size
        self.xs = []
        self.xs_len = 0

    def __call__(self, cnn, x, x_len: int):
        self.xs.append(x)
        self.xs_len += x.shape[-2]
        x_start = max(self.xs_len - max(x.shape[-2], self.window_size), 0)
        need = self.xs_len - x_start
        keep, got = [], 0
        for piece in reversed(self.xs):
            if got >= need:
                break
            keep.append(piece)
            got += piece.shape[-2]
        keep.reverse()
        self.xs = keep
        window = torch.cat(keep, dim=-2)
        window = window[..., -min(window.shape[-2], need):, :]
        return cnn(window)[..., -x_len:, :].clone()
