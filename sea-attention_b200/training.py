"""Training branch of the causal layer (SURVEY a16 + 8f-1; reference attention.py:518-534, 595-765, 960-962, 1066-1133, 1237-1250,
1279-1282, 1328-1359 with `benchmarking=False`): the predictor's distillation losses against the teacher's attention and the dense
masked attention, differentiable end to end (q, k, v and every predictor parameter receive gradient).

This branch is O(T^2) by definition -- the teacher's `attention_scores_truth` is a dense [N,H,T,T] tensor and the reference's losses are
dense KL / MSE terms against it -- and it is NOT the hot path of this repo.  It is therefore composed of differentiable torch operations
on the GPU (cuBLAS / cuDNN do the Linear and Conv2d gradients), arranged for the device rather than translated: the causal Performer is
chunk-parallel (per-chunk state sums + an exclusive prefix over chunks, the formulation of csrc/performer_mma.cu) instead of a
T-step scan, each CausalConv2d is ONE dilated convolution over its masked 5x3 weight, the interpolation is a precomputed gather.  The
grouped top-k -- the one stage that must be bit-exact and carries no gradient -- runs on the repo's own kernel (sea_topk_mask_bits).
No CPU path: CUDA tensors only.  Parity: tests/test_training_gpu.py (loss and context against the unmodified reference's training-mode
run, tests/golden/layer_causal_training_h3_t48.npz; gradients against autograd of the CPU oracle's restatement).
"""
import random

import torch
import torch.nn.functional as F

from . import ops
from ._lib import SeaError


def _fp_min(dtype):
    return torch.finfo(torch.float16).min / 2 if dtype in (torch.float16, torch.bfloat16) else torch.finfo(torch.float32).min / 2


def _features(x, proj):
    """generalized (ReLU) Performer features, attention.py:159-164: relu(d^-1/4 x P^T) + 1e-3"""
    return F.relu(F.linear(x * (x.shape[-1] ** -0.25), proj)) + 1e-3


def performer_causal(q, k, v2, proj, chunk: int = 128, eps: float = 1e-6):
    """out_t = phi(q_t) . S_t / phi(q_t) . (z_t + eps), S_t = sum_{s<=t} phi(k_s) (x) v2_s, z_t = sum_{s<=t} phi(k_s); chunk-parallel:
    within a chunk the lower triangle of phi(q) phi(k)^T, across chunks an exclusive prefix of the per-chunk sums."""
    N, H, T, d = q.shape
    E = v2.shape[-1]
    pad = (-T) % chunk
    if pad:         # zero rows at the END only influence (dropped) later rows
        q, k, v2 = (F.pad(t, (0, 0, 0, pad)) for t in (q, k, v2))
    nc = (T + pad) // chunk
    qf = _features(q, proj).view(N, H, nc, chunk, -1)
    kf = _features(k, proj).view(N, H, nc, chunk, -1)
    vc = v2.view(N, H, nc, chunk, E)
    s_c = torch.einsum('nhcjf,nhcje->nhcfe', kf, vc)
    z_c = kf.sum(3)
    s_prev = s_c.cumsum(2) - s_c
    z_prev = z_c.cumsum(2) - z_c
    a = torch.tril(torch.einsum('nhcif,nhcjf->nhcij', qf, kf))
    num = a @ vc + torch.einsum('nhcif,nhcfe->nhcie', qf, s_prev)
    den = a.sum(-1) + torch.einsum('nhcif,nhcf->nhci', qf, z_prev + eps)          # = phi(q_t) . (z_t + eps)
    out = num / den.unsqueeze(-1)
    return out.reshape(N, H, nc * chunk, E)[:, :, :T]


def _resize_index(attention_mask, P: int, jitter: bool):
    """Column -> pixel map of resize_from_m_to_t (resize_m_to_t.py:36-48): column j of a row with L valid columns reads pixel
    floor((rank_j - 0.5) / L * P - 1e-4); masked columns read the pad pixel P.  attention_mask [N,1,T1,T2] additive.
    jitter: the training-time perturbation of the ranks (:40-45), applied by the caller's coin."""
    mask = (attention_mask > -1).float()
    cs = mask.cumsum(-1)
    length = cs[:, :, :, -1:]
    if jitter:      # verbatim bounds of the reference, including its use of the arg-max INDEX of the rank as the upper bound
        cs = torch.clamp(cs + (torch.rand_like(cs) * 1.5 - 0.75), torch.ones((1, 1, 1, 1), device=cs.device),
                         cs.max(dim=-1, keepdim=True)[1].to(cs.dtype))
    idx = torch.floor(((cs - 1) + 0.5) / length * P - 1e-4).to(torch.long) + ((1 - mask) * P).to(torch.long)
    return torch.clamp(idx, 0, P)


def _resize(x, fill: float, idx):
    """x [N,H,T1,P] -> [N,H,T1,T2] through the index map (differentiable w.r.t. x)."""
    N, H, T1, P = x.shape
    return F.pad(x, (0, 1), value=fill).gather(-1, idx.expand(N, H, T1, idx.shape[-1]))


def _kd_loss(scores, truth, dead, fmin):
    """0.1 * KL(batchmean) + MSE between the row softmaxes of `scores` and of the teacher's scores, both causally masked
    (attention.py:740-765, 1084-1102); fp32."""
    T2 = scores.shape[-1]
    s = scores.float().masked_fill(dead, fmin)
    target = F.softmax(truth.float().masked_fill(dead, fmin), dim=-1).view(-1, T2)
    logp = F.log_softmax(s, dim=-1).view(-1, T2)
    return F.kl_div(logp, target, reduction='batchmean') * 0.1 + F.mse_loss(F.softmax(s, dim=-1).view(-1, T2), target)


def forward_train(mod, q, k, v, q_for_atten, k_for_atten, v_for_atten, q_for_score, k_for_score, attention_mask,
                  attention_scores_truth, context_layer_truth, output_cls, topk_mask_fn=None):
    """The `benchmarking=False` forward of the causal layer with losses.  `topk_mask_fn(probs, k_per_row, row_valid) -> 0/1 mask
    [N,H,T,P]` defaults to the CUDA top-k kernel (tests inject a CPU stand-in to check the torch math without a GPU)."""
    pc = mod.pconfig
    if not pc.causal:
        raise SeaError('the training branch is implemented for the causal model')
    N, H, T, d = q.shape
    P = pc.attention_predictor_length
    if attention_mask is None:
        fm = _fp_min(q.dtype)
        attention_mask = ((torch.arange(T, device=q.device).view(1, T) > torch.arange(T, device=q.device).view(T, 1)) * fm).to(q.dtype).view(1, 1, T, T).expand(N, 1, T, T)
    if attention_mask.shape != (N, 1, T, T):
        raise SeaError(f'causal additive mask must be [N,1,T,T], got {tuple(attention_mask.shape)}')
    fmin = _fp_min(q.dtype)
    training = mod.training
    dst_valid = (attention_mask[:, :, :, :1] > -1)                                   # [N,1,T,1]
    padded = not bool(dst_valid.all())
    # a2 (+ :512-514)
    pos = mod.v_eye_learned_causal[:, :, :T, :].to(v.dtype).expand(N, H, T, d)
    v2 = torch.cat([pos, v_for_atten], dim=-1)
    if padded:
        v2 = v2 * dst_valid.to(v2.dtype)
        v = v * dst_valid.to(v.dtype)
    # a3 in fp32 (:521-534)
    ctx_p = performer_causal(q_for_atten.float(), k_for_atten.float(), v2.float(), mod.performer.projection_matrix.float()).to(q.dtype)
    # a4
    enc, dec, scl, cnn = mod.attention_predictor_enc, mod.attention_predictor_dec_row, mod.attention_predictor_dec_scaler, mod.attention_predictor_cnn
    cast = lambda p_: p_.to(q.dtype)
    t_pred = F.gelu(F.layer_norm(F.linear(torch.cat([ctx_p, v], dim=-1), cast(enc[0].weight), cast(enc[0].bias)), (2 * d,),
                                 cast(enc[1].weight), cast(enc[1].bias)))
    S = mod.attention_predictor_dec_row_splits
    W = P // mod.attention_predictor_dec_row_down_scale
    x = F.linear(t_pred, cast(dec[0].weight), cast(dec[0].bias)).view(N, H, T, S, W).permute(0, 1, 3, 2, 4).reshape(N, H * S, T, W)
    scales = F.linear(t_pred, cast(scl[0].weight), cast(scl[0].bias))
    # a5: LN(W) -> [CausalConv2d 3x3 dil 2 + ReLU] x 2 (3) -> nearest x4 -> CausalConv2d 1x1 (pad 1) -> area resize -> LN(P)
    x = F.layer_norm(x, (W,), cast(cnn[0].module.weight), cast(cnn[0].module.bias))
    c3x3, c1x1 = mod._cnn_convs()
    for c in c3x3:
        x = F.relu(F.conv2d(x, cast(c.weight * c.weight_mask), cast(c.bias), padding=(4, 2), dilation=2))
    x = x.repeat_interleave(4, dim=-1)
    x = F.conv2d(x, cast(c1x1.weight * c1x1.weight_mask), cast(c1x1.bias), padding=(0, 1))
    x = F.adaptive_avg_pool2d(x.float(), (T, P)).to(q.dtype)
    score = F.layer_norm(x, (P,), cast(cnn[2].module.weight), cast(cnn[2].module.bias))
    probs = F.softmax(score.float(), dim=-1).to(score.dtype) if training else F.softmax(score, dim=-1)           # softmax_bf16 (:62-72)
    # interpolation maps (the reference draws its 10 % jitter coin once per resize call, :40-41)
    coin = lambda: training and random.random() < 0.1
    dead = attention_mask < -1
    loss = 0
    est_probs_resized = None
    if attention_scores_truth is not None:
        est_probs_resized = _resize(probs, 0.0, _resize_index(attention_mask, P, coin()))
        est_score_resized = _resize(score.float(), fmin, _resize_index(attention_mask, P, coin()))
        loss = loss + _kd_loss(est_score_resized, attention_scores_truth, dead, fmin)
    # a7: grouped top-k (no gradient), then the interpolated additive mask (:960-962)
    probs_k = probs.detach().float()
    if padded:
        probs_k = probs_k * dst_valid.to(probs_k.dtype)
    kpr, _ = mod._shape_consts(H, P, T, T, q.device)
    kpr = kpr.repeat(N) if N > 1 else kpr
    row_valid = dst_valid.view(N, T) if padded else None
    if topk_mask_fn is None:
        bits = ops.topk_mask_bits(probs_k.contiguous(), kpr, 'causal_batch', row_valid=row_valid)
        mask_m = ops.bits_to_mask(bits, H, P)
    else:
        mask_m = topk_mask_fn(probs_k, kpr, row_valid)
    pm = _resize((1.0 - mask_m) * fmin, fmin, _resize_index(attention_mask, P, coin()))
    if pc.k_oversample != 1.0:
        raise SeaError('the training branch does not implement k_oversample != 1 (resize_m_to_t.py:54-71)')
    pm = pm.masked_fill(dead, fmin)
    # a9-a12, dense (:1066-1133)
    dense = torch.matmul(q_for_score, k_for_score.transpose(-1, -2))
    if attention_scores_truth is not None:
        loss = loss + _kd_loss(dense, attention_scores_truth, dead, fmin)
    sm = (lambda t_: F.softmax(t_.float(), dim=-1).to(t_.dtype)) if training else (lambda t_: F.softmax(t_, dim=-1))
    dense_masked = dense.masked_fill(dead, fmin)                 # (the reference writes the causal fill into `dense` in place, :1087)
    dense_probs = sm(dense_masked + attention_mask.to(dense.dtype))
    pp = sm(dense_masked + pm.to(dense.dtype)).masked_fill(pm < -1, 0)
    if pc.partial_attention_scaler:
        pp = pp * torch.sigmoid(scales[..., 0:1])
    ctx = torch.matmul(pp, v)
    # a13, a14
    avg = (v * dst_valid.to(v.dtype)).cumsum(-2) / torch.arange(1, T + 1, device=v.device, dtype=torch.float32).view(1, 1, T, 1)
    a = torch.sigmoid(scales[..., 1:2])
    out = ctx * a + (1 - a) * avg.to(v.dtype)
    context = out.permute(0, 2, 1, 3).reshape(N, T, H * d)
    if context_layer_truth is not None:
        loss = loss + F.mse_loss(context_layer_truth, context)
    return output_cls(loss=loss, context_layer=context, partial_attention_probs=pp, partial_attention_mask=pm,
                      estimated_attention_probs_m=probs, estimated_attention_probs=est_probs_resized, dense_attention_probs=dense_probs,
                      key_for_score=k_for_score, state=None)
