"""GPU tests of the use_cache / decode path.  The reference checks its stateful ops against the stateless
forward (test_perlin_opt_cache.py: cache vs no-cache consistency); the same property is the parity target here: row t of a
token-by-token decode equals row t of the prefill, whose parity with the reference is established in test_layer_gpu.py."""
import pytest
import torch

from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _module(sea, H, d, T, P, k, nbf, seed):
    import transformers
    torch.manual_seed(seed)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV)
    for n_, p_ in mod.named_parameters():
        if p_.ndim == 1:
            p_.data.add_(0.1 * torch.randn_like(p_))
    return mod


def _close_rows(a, b, rtol, atol):
    """fraction of rows (dim 1) of two [N,T,C] tensors that agree within tolerance"""
    ok = ((a - b).abs() <= atol + rtol * b.abs()).all(dim=-1).all(dim=0)
    return float(ok.float().mean())


@pytest.mark.parametrize('N,H,d,T,T0,P,k', [(1, 4, 64, 48, 29, 32, 8), (2, 3, 32, 40, 2, 32, 8)])
def test_decode_matches_prefill_fp32(sea, N, H, d, T, T0, P, k):
    mod = _module(sea, H, d, T, P, k, 8, seed=T + H)
    g = torch.Generator().manual_seed(7)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).to(DEV)
    kk = torch.randn(N, H, T, d, generator=g).to(DEV)
    v = torch.randn(N, H, T, d, generator=g).to(DEV)
    full = mod(q, kk, v, q, kk, v, q, kk, so.causal_additive_mask(T, torch.float32, N).to(DEV), None, None)
    assert full.state is None
    mod.pconfig.use_cache = True
    s = lambda x, a, b: x[:, :, a:b]
    out0 = mod(s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0),
               so.causal_additive_mask(T0, torch.float32, N).to(DEV), None, None)
    assert out0.state is not None and out0.state.t == T0
    torch.testing.assert_close(out0.context_layer, full.context_layer[:, :T0], rtol=1e-4, atol=1e-5)
    state = out0.state
    ctx_rows, prob_rows = [], []
    t = T0
    for step in (1, 1, 3, 1):                      # single tokens and a 3-token chunk
        while t + step <= T and (step == 3 or len(ctx_rows) < 100):
            dummy = torch.zeros(N, 1, step, t + step, device=DEV)
            o = mod(s(q, t, t + step), s(kk, 0, t + step), s(v, 0, t + step), s(q, t, t + step), s(kk, 0, t + step), s(v, 0, t + step),
                    s(q, t, t + step), s(kk, 0, t + step), dummy, None, None, last_state=state)
            assert o.state.t == t + step and state.t == t          # functional update: the caller's state is untouched
            state = o.state
            ctx_rows.append(o.context_layer)
            prob_rows.append(o.estimated_attention_probs)
            t += step
            if step == 3:
                break
        if t >= T:
            break
    while t < T:
        dummy = torch.zeros(N, 1, 1, t + 1, device=DEV)
        o = mod(s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1),
                dummy, None, None, last_state=state)
        state = o.state
        ctx_rows.append(o.context_layer)
        prob_rows.append(o.estimated_attention_probs)
        t += 1
    ctx = torch.cat(ctx_rows, dim=1)
    probs = torch.cat(prob_rows, dim=2)
    torch.testing.assert_close(probs, full.estimated_attention_probs[:, :, T0:], rtol=1e-3, atol=1e-6)
    # a near-tie of the top-k may flip under the (1e-6-level) different summation order of the incremental Performer
    assert _close_rows(ctx, full.context_layer[:, T0:], 1e-3, 1e-4) >= 0.9


def test_decode_bf16_tensor_core_shape(sea):
    """OPT-1.3B-like head count (tcgen05 MLP / conv kernels in the step): probabilities track the prefill within bf16 tolerance."""
    N, H, d, T, T0, P, k = 1, 32, 64, 80, 64, 64, 16
    mod = _module(sea, H, d, T, P, k, 8, seed=3)
    g = torch.Generator().manual_seed(8)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    full = mod(q, kk, v, q, kk, v, q, kk, so.causal_additive_mask(T, torch.bfloat16, N).to(DEV), None, None)
    mod.pconfig.use_cache = True
    s = lambda x, a, b: x[:, :, a:b]
    o = mod(s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0),
            so.causal_additive_mask(T0, torch.bfloat16, N).to(DEV), None, None)
    state = o.state
    rows, prows = [], []
    for t in range(T0, T):
        o = mod(s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1),
                torch.zeros(N, 1, 1, t + 1, device=DEV, dtype=torch.bfloat16), None, None, last_state=state)
        state = o.state
        rows.append(o.context_layer)
        prows.append(o.estimated_attention_probs)
    probs = torch.cat(prows, dim=2)
    torch.testing.assert_close(probs, full.estimated_attention_probs[:, :, T0:], rtol=1e-1, atol=2e-3)
    ctx = torch.cat(rows, dim=1).float()
    assert torch.isfinite(ctx).all()
    assert _close_rows(ctx, full.context_layer[:, T0:].float(), 5e-2, 5e-2) >= 0.5


def test_decode_primitives_match_the_stateful_oracle(sea):
    """a17 against an oracle of attention_state.py (oracle.StatefulCausalPerformerOracle / StatefulCumAvgOracle /
    StatefulCausalCNNOracle, pinned to the unmodified reference classes by tests/golden/state_ops.npz): the incremental Performer
    + running mean kernel and the windowed causal convolutions, advanced over the same token chunks.

    The reference's recurrence is fed what its stateless Performer feeds the same sums: the generalized-attention features of
    q / k (common/performer.py, FastAttention) and v_for_atten = cat(v_eye_learned_causal, v).  (attention_state.py:287-301 hands
    the recurrence the raw q / k instead -- the reference's own 'TODO: fix numerical stability' path -- so the decode of the
    reference is not its prefill; here row t of a decode IS row t of the prefill, the property test_perlin_opt_cache.py wants.)"""
    N, H, d, T, F = 2, 3, 32, 37, 11
    chunks = [5, 1, 1, 3, 1, 16, 1, 9]
    g = torch.Generator().manual_seed(11)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5)
    kk = torch.randn(N, H, T, d, generator=g)
    v = torch.randn(N, H, T, d, generator=g)
    proj = torch.randn(F, d, generator=g)
    pos = torch.randn(T, d, generator=g)
    qf, kf = so.performer_features_generalized(q, proj), so.performer_features_generalized(kk, proj)
    vfa = torch.cat([pos.view(1, 1, T, d).expand(N, H, T, d), v], dim=-1)
    perf, ca = so.StatefulCausalPerformerOracle(), so.StatefulCumAvgOracle()
    state = sea.ops.performer_state_new(N, H, d, F, DEV)
    qd, kd, vd, pd, prd = (x.to(DEV) for x in (q, kk, v, pos, proj))
    t = 0
    for i, c in enumerate(chunks):
        want_ctx = perf(qf[:, :, t:t + c], kf[:, :, :t + c], vfa[:, :, :t + c])
        want_avg = ca(v[:, :, :t + c], c)
        if i == 0:          # the prompt: the state a prefill leaves behind
            sea.ops.performer_state_build(kd[:, :, :c], vd[:, :, :c], pd, prd, state)
            got_ctx, got_avg = sea.ops.performer_causal(qd[:, :, :c], kd[:, :, :c], vd[:, :, :c], pd, prd)
        else:
            got_ctx, got_avg = sea.ops.performer_causal_state(qd[:, :, t:t + c], kd[:, :, t:t + c], vd[:, :, t:t + c], pd, prd, state, t)
        torch.testing.assert_close(got_ctx.cpu(), want_ctx, rtol=2e-4, atol=2e-5)
        torch.testing.assert_close(got_avg.cpu(), want_avg, rtol=1e-5, atol=1e-6)
        t += c
    # windowed CNN: the decode keeps the last 4 rows of each dilated conv's input (5-row windows); the oracle re-runs >= 24 rows
    import numpy as np, os
    fx = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'state_ops.npz'))
    w0, m0, b0 = (torch.from_numpy(fx['cnn.0.' + n]) for n in ('weight', 'weight_mask', 'bias'))
    w2, m2, b2 = (torch.from_numpy(fx['cnn.2.' + n]) for n in ('weight', 'weight_mask', 'bias'))
    x = torch.from_numpy(fx['cnn_x'])                       # [N,C,T,W]
    xcl = x.permute(0, 2, 3, 1).contiguous().to(DEV)         # channels-last [N,T,W,C]
    Nc, C, Tc, Wc = x.shape
    win_x = torch.zeros(Nc, 4, Wc, C, device=DEV)
    win_y = torch.zeros(Nc, 4, Wc, C, device=DEV)
    rows = []
    for tt in range(Tc):
        xw = torch.cat([win_x, xcl[:, tt:tt + 1]], dim=1)
        y1 = sea.ops.causal_conv3x3_dil2_relu(xw, (w0 * m0).to(DEV), b0.to(DEV))[:, 4:5].contiguous()
        yw = torch.cat([win_y, y1], dim=1)
        y2 = sea.ops.causal_conv3x3_dil2_relu(yw, (w2 * m2).to(DEV), b2.to(DEV))[:, 4:5].contiguous()
        rows.append(y2)
        win_x, win_y = xw[:, 1:].contiguous(), yw[:, 1:].contiguous()
    got = torch.cat(rows, dim=1).permute(0, 3, 1, 2).cpu()
    torch.testing.assert_close(got, torch.from_numpy(fx['cnn_out']), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('N,H,d,T,T0,P,k,dtype', [(1, 32, 64, 72, 64, 64, 16, torch.bfloat16), (2, 12, 64, 40, 30, 256, 16, torch.bfloat16),
                                                 (1, 32, 128, 70, 64, 256, 32, torch.bfloat16), (2, 3, 32, 24, 2, 32, 8, torch.float32),
                                                 (1, 4, 64, 300, 290, 32, 8, torch.float32), (1, 8, 80, 50, 45, 128, 8, torch.bfloat16)])
def test_native_decode_step_equals_the_per_op_sequence(sea, N, H, d, T, T0, P, k, dtype):
    """sea_decode_step (one C call per token, csrc/decode_step.cu) against the per-op python sequence: the whole state must come out
    identical, probabilities and context equal up to the summation order of the fused tail kernel -- tensor-core shapes, zero-padded channels (H = 12), other head dims,
    fp32 (CSR attention with the shape-derived nnz bound), N > 1, a prompt shorter than the 4-row CNN window."""
    mod = _module(sea, H, d, T, P, k, 8, seed=T + H + d)
    g = torch.Generator().manual_seed(21)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).to(dtype).to(DEV)
    kk = torch.randn(N, H, T, d, generator=g).to(dtype).to(DEV)
    v = torch.randn(N, H, T, d, generator=g).to(dtype).to(DEV)
    mod.pconfig.use_cache = True
    s = lambda x, a, b: x[:, :, a:b]
    o = mod(s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0), None, None, None)
    st_a = st_b = o.state
    for t in range(T0, T):
        args = (s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), None, None, None)
        mod.decode_native = True
        a = mod(*args, last_state=st_a)
        mod.decode_native = False
        b = mod(*args, last_state=st_b)
        mod.decode_native = True
        assert a.state.t == b.state.t == t + 1
        # the native step runs the prefill's fused tail / top-k kernel (1x1 conv row + tail_topk_reg), the per-op sequence the stand-alone
        # tail and top-k kernels: same probabilities up to summation order, same mask except where two keys tie within that rounding
        torch.testing.assert_close(a.estimated_attention_probs, b.estimated_attention_probs, rtol=1e-4, atol=1e-7)
        ca, cb = a.context_layer.float(), b.context_layer.float()
        assert float(((ca - cb).abs() > 2e-2 * (1 + cb.abs())).float().mean()) < 0.01, t
        assert torch.equal(a.state.performer, b.state.performer) and torch.equal(a.state.cnn_in_win, b.state.cnn_in_win.contiguous())
        assert torch.equal(a.state.conv1_win, b.state.conv1_win.contiguous())
        st_a, st_b = a.state, b.state


@pytest.mark.parametrize('N,H,d,T,T0,P,k', [(1, 32, 64, 80, 64, 64, 16), (2, 12, 64, 40, 30, 256, 16)])
def test_frozen_module_decode_plan_equals_unfrozen(sea, N, H, d, T, T0, P, k):
    """A frozen module resolves the per-module constants of the native step once (the decode plan kept beside the cached weights):
    same kernels, same arguments -- outputs and state identical to the unfrozen module's, and the plan goes away with the cache."""
    mod = _module(sea, H, d, T, P, k, 8, seed=3)
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(N, H, T, d, generator=g) * d ** -0.5).bfloat16().to(DEV)
    kk = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    v = torch.randn(N, H, T, d, generator=g).bfloat16().to(DEV)
    mod.pconfig.use_cache = True
    s = lambda x, a, b: x[:, :, a:b]
    o = mod(s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0), s(v, 0, T0), s(q, 0, T0), s(kk, 0, T0), None, None, None)
    st_a = st_b = o.state
    used_plan = False
    for t in range(T0, T):
        args = (s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), None, None, None)
        mod.freeze_packed_weights(False)
        b = mod(*args, last_state=st_b)
        mod.freeze_packed_weights(True)
        if t > T0:            # keep the frozen-side caches of the previous step: re-freezing dropped them, so run two frozen calls
            mod(*args, last_state=st_a)
            mod(*args, last_state=st_a)
        a = mod(*args, last_state=st_a)
        used_plan = used_plan or (mod._w_cache is not None and '_decode_plan' in mod._w_cache)
        assert torch.equal(a.estimated_attention_probs, b.estimated_attention_probs), t
        assert torch.equal(a.context_layer, b.context_layer), t
        assert torch.equal(a.state.performer, b.state.performer) and torch.equal(a.state.cnn_in_win, b.state.cnn_in_win)
        assert torch.equal(a.state.conv1_win, b.state.conv1_win)
        st_a, st_b = a.state, b.state
    assert used_plan
    mod.invalidate_packed()
    assert mod._w_cache is None
