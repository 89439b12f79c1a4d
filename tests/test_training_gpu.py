"""Training branch on the GPU (SURVEY a16 + 8f-1; BASELINE configs[2] is a fwd+bwd workload): PerlinAttention in train() mode with
teacher tensors, through the public forward, with the CUDA grouped top-k -- loss, context and the gradients w.r.t. q, k, v and the
predictor parameters against autograd of the CPU oracle's restatement (pinned to the unmodified reference's training-mode run by
tests/test_training_cpu.py)."""
import random

import numpy as np
import pytest
import torch
import transformers

from conftest import golden_layer
from oracle import sea_oracle as so

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _module(sea, m, sd):
    cfg = transformers.BertConfig(hidden_size=m['H'] * m['d'], num_attention_heads=m['H'], max_position_embeddings=m['T'])
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=m['nbf'], k=m['k'], attention_predictor_length=m['P'], causal=True))
    assert not mod.load_state_dict(sd, strict=False)[1]
    return mod.to(DEV)


def test_training_forward_backward_matches_oracle_autograd_fp32(sea, monkeypatch):
    g, m, sd = golden_layer('layer_causal_training_h3_t48')
    N, H, T, P, d, k = m['N'], m['H'], m['T'], m['P'], m['d'], m['k']
    monkeypatch.setattr(random, 'random', lambda: 1.0)                       # the 10 % resize jitter off (as in the fixture)
    mod = _module(sea, m, sd).train()
    truth, ctx_truth = torch.from_numpy(g['scores_truth']), torch.from_numpy(g['context_truth'])
    q, kk, v = (torch.from_numpy(g[x]).to(DEV).requires_grad_(True) for x in 'qkv')
    out = mod(q, kk, v, q, kk, v, q, kk, so.causal_additive_mask(T, torch.float32, N).to(DEV), truth.to(DEV), ctx_truth.to(DEV))
    out.loss.backward()
    # oracle on the CPU, its own top-k rule == the CUDA kernel's (lower flat index wins a tie)
    sd_g = {k_: v_.clone().requires_grad_(v_.dtype.is_floating_point) for k_, v_ in sd.items()}
    q2, k2, v2 = (torch.from_numpy(g[x]).clone().requires_grad_(True) for x in 'qkv')
    b = so.perlin_train_forward(sd_g, q2, k2, v2, truth, ctx_truth, k_top=k, P=P)
    b['loss'].backward()
    assert abs(float(out.loss.detach()) - float(b['loss'])) <= 1e-4 * abs(float(b['loss'])) + 1e-6, (float(out.loss), float(b['loss']))
    # and within the tie-induced band of the unmodified reference's loss
    assert abs(float(out.loss.detach()) - float(g['loss'])) <= 2e-3 * abs(float(g['loss']))
    mine = (out.partial_attention_mask.detach().cpu() > -1)
    assert float((mine == (b['partial_attention_mask'] > 0.5)).float().mean()) >= 0.999
    torch.testing.assert_close(out.estimated_attention_probs_m.detach().cpu(), b['estimated_attention_probs'].detach(), rtol=1e-3, atol=2e-6)
    same_rows = (mine == (b['partial_attention_mask'] > 0.5)).all(dim=3).all(dim=1)          # [N,T]
    torch.testing.assert_close(out.context_layer.detach().cpu()[same_rows], b['context_layer'].detach()[same_rows], rtol=1e-3, atol=3e-5)
    if bool(same_rows.all()):
        for a_, b_, name in ((q.grad, q2.grad, 'q'), (kk.grad, k2.grad, 'k'), (v.grad, v2.grad, 'v')):
            torch.testing.assert_close(a_.cpu(), b_, rtol=3e-3, atol=2e-6, msg=lambda s_, n_=name: f'd loss / d {n_}: {s_}')
        checked = 0
        for name, p_ in mod.named_parameters():
            ref = sd_g.get(name)
            if ref is None or ref.grad is None:
                continue
            torch.testing.assert_close(p_.grad.cpu(), ref.grad, rtol=3e-3, atol=3e-6, msg=lambda s_, n_=name: f'gradient of {n_}: {s_}')
            checked += 1
        assert checked >= 18, checked
    else:       # a near-tie flipped between the CPU and the GPU arithmetic: the gradients still have to be close in aggregate
        rel = float((q.grad.cpu() - q2.grad).norm() / q2.grad.norm())
        assert rel < 5e-2, rel


def test_training_step_bf16_and_eval_loss(sea, monkeypatch):
    """bf16 parameters and inputs (how the reference trains: autocast): finite loss close to the fp32 one, gradients on every predictor
    parameter; eval() with teacher tensors reports the same loss without the training-mode extras."""
    g, m, sd = golden_layer('layer_causal_training_h3_t48')
    N, T = m['N'], m['T']
    monkeypatch.setattr(random, 'random', lambda: 1.0)
    truth, ctx_truth = torch.from_numpy(g['scores_truth']).to(DEV), torch.from_numpy(g['context_truth']).to(DEV)
    mod = _module(sea, m, sd).train()
    qf, kf, vf = (torch.from_numpy(g[x]).to(DEV) for x in 'qkv')
    ref = mod(qf, kf, vf, qf, kf, vf, qf, kf, None, truth, ctx_truth)
    mod_b = _module(sea, m, sd).bfloat16().train()
    q, kk, v = (torch.from_numpy(g[x]).to(DEV).bfloat16().requires_grad_(True) for x in 'qkv')
    out = mod_b(q, kk, v, q, kk, v, q, kk, None, truth.bfloat16(), ctx_truth.bfloat16())
    assert torch.isfinite(out.loss) and abs(float(out.loss) - float(ref.loss)) < 0.1 * abs(float(ref.loss))
    out.loss.float().backward()
    assert q.grad is not None and torch.isfinite(q.grad.float()).all()
    missing = [n_ for n_, p_ in mod_b.named_parameters() if p_.grad is None and ('predictor_enc.' in n_ or 'dec_row' in n_ or 'dec_scaler' in n_ or 'predictor_cnn' in n_)]
    assert not missing, missing
    mod.eval()
    with torch.no_grad():
        ev = mod(qf, kf, vf, qf, kf, vf, qf, kf, None, truth, ctx_truth)
    assert abs(float(ev.loss) - float(ref.loss)) <= 1e-4 * abs(float(ref.loss)) + 1e-6


def test_training_step_at_a_model_shape(sea):
    """One fwd+bwd at an OPT-like head shape (H = 12, d = 64, T = 512, P = 128): runs, finite, all gradients present."""
    N, H, d, T, P, k = 2, 12, 64, 512, 128, 32
    torch.manual_seed(0)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=8, k=k, attention_predictor_length=P, causal=True)).to(DEV).train()
    q = (torch.randn(N, H, T, d, device=DEV) * d ** -0.5).requires_grad_(True)
    kk = torch.randn(N, H, T, d, device=DEV, requires_grad=True)
    v = torch.randn(N, H, T, d, device=DEV, requires_grad=True)
    truth = torch.randn(N, H, T, T, device=DEV)
    out = mod(q, kk, v, q, kk, v, q, kk, None, truth, torch.randn(N, T, H * d, device=DEV))
    out.loss.backward()
    assert torch.isfinite(out.loss) and all(torch.isfinite(t_.grad).all() for t_ in (q, kk, v))
    assert out.context_layer.shape == (N, T, H * d) and out.dense_attention_probs.shape == (N, H, T, T)
