"""Training branch (SURVEY a16 + 8f-1), CPU part: (1) the oracle's restatement of the reference's training-mode forward
(oracle.sea_oracle.perlin_train_forward) is pinned to the UNMODIFIED reference run in train() mode with teacher tensors
(tests/golden/layer_causal_training_h3_t48.npz: loss, context, probabilities, masks); (2) the torch math of the product's training
branch (sea-attention_b200/training.py) is held to that oracle -- values and, through autograd on both sides, the gradients w.r.t.
q, k, v and every predictor parameter.  The product refuses CPU tensors at its public entry (PerlinAttention.forward); here
training.forward_train is called directly with the oracle's top-k injected in place of the CUDA kernel, to check the differentiable math
where no GPU exists.  The GPU run of the same comparison (with the CUDA top-k) is tests/test_training_gpu.py."""
import importlib

import numpy as np
import pytest
import torch
import transformers

from conftest import golden_layer
from oracle import sea_oracle as so


def _bits(g, key, shape):
    n = int(np.prod(shape))
    return np.unpackbits(g[key])[:n].reshape(shape)


def test_oracle_training_forward_matches_reference_fixture():
    g, m, sd = golden_layer('layer_causal_training_h3_t48')
    N, H, T, P, d, k = m['N'], m['H'], m['T'], m['P'], m['d'], m['k']
    q, kk, v = (torch.from_numpy(g[x]) for x in 'qkv')
    # on the reference's own top-k selection (its unstable CPU sort cuts the exact ties of the x4-upsampled predictor arbitrarily; the
    # oracle's rule -- lower flat index wins -- picks other members of the same ties): identical loss, masks and context
    ref_mask = torch.from_numpy(_bits(g, 'mask_before_interp_alive', (N, H, T, P)).astype(np.float32))
    b = so.perlin_train_forward(sd, q, kk, v, torch.from_numpy(g['scores_truth']), torch.from_numpy(g['context_truth']), k_top=k, P=P,
                                mask_override=ref_mask)
    assert abs(float(b['loss']) - float(g['loss'])) <= 1e-5 * abs(float(g['loss'])) + 1e-6, (float(b['loss']), float(g['loss']))
    # the oracle's own selection differs from the reference's only inside ties: same number of pixels per row, equal probabilities
    own = so.topk_mask_causal_batch(b['estimated_attention_probs'], k)
    assert torch.equal(own.sum(dim=(1, 3)), ref_mask.sum(dim=(1, 3)))
    pr = b['estimated_attention_probs'].transpose(1, 2).reshape(N * T, -1)
    for r in range(N * T):
        a_, b_ = own.transpose(1, 2).reshape(N * T, -1)[r].bool(), ref_mask.transpose(1, 2).reshape(N * T, -1)[r].bool()
        if not torch.equal(a_, b_):
            assert torch.allclose(pr[r][a_ & ~b_].sort().values, pr[r][b_ & ~a_].sort().values, rtol=0, atol=1e-7)
    torch.testing.assert_close(b['estimated_attention_probs'], torch.from_numpy(g['estimated_attention_probs_m']), rtol=1e-3, atol=2e-6)
    torch.testing.assert_close(b['estimated_attention_probs_resized'], torch.from_numpy(g['estimated_attention_probs']), rtol=1e-3, atol=2e-6)
    torch.testing.assert_close(b['dense_attention_probs'], torch.from_numpy(g['dense_attention_probs']), rtol=1e-3, atol=2e-6)
    alive = _bits(g, 'partial_attention_mask_alive', (N, H, T, T)).astype(bool)
    assert np.array_equal(alive, b['partial_attention_mask'].numpy().astype(bool))
    torch.testing.assert_close(b['context_layer'], torch.from_numpy(g['context_layer']), rtol=1e-3, atol=2e-5)


def _oracle_topk(k):
    def fn(probs, kpr, row_valid):
        dv = None if row_valid is None else row_valid.float()
        return so.topk_mask_causal_batch(probs.cpu(), k, 1.0, dv).to(probs.device)
    return fn


def _grads_case(seed=0):
    sea = importlib.import_module('sea-attention_b200')
    g, m, sd = golden_layer('layer_causal_training_h3_t48')
    N, H, T, P, d, k, nbf = (m[x] for x in ('N', 'H', 'T', 'P', 'd', 'k', 'nbf'))
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True))
    assert not mod.load_state_dict(sd, strict=False)[1]
    return sea, g, m, sd, mod


def test_training_math_matches_oracle_values_and_gradients(monkeypatch):
    import random
    sea, g, m, sd, mod = _grads_case()
    training = importlib.import_module('sea-attention_b200.training')
    N, H, T, P, d, k = m['N'], m['H'], m['T'], m['P'], m['d'], m['k']
    monkeypatch.setattr(random, 'random', lambda: 1.0)                       # the 10 % resize jitter off, as in the fixture
    truth, ctx_truth = torch.from_numpy(g['scores_truth']), torch.from_numpy(g['context_truth'])
    # the reference's own top-k selection (arbitrary inside exact ties, see the test above) is injected on both sides
    ref_mask = torch.from_numpy(_bits(g, 'mask_before_interp_alive', (N, H, T, P)).astype(np.float32))
    # product side
    mod.train()
    q, kk, v = (torch.from_numpy(g[x]).clone().requires_grad_(True) for x in 'qkv')
    own = training.forward_train(mod, q, kk, v, q, kk, v, q, kk, so.causal_additive_mask(T, torch.float32, N), truth, ctx_truth,
                                 sea.PerlinAttentionOutput, topk_mask_fn=_oracle_topk(k))
    b_own = so.perlin_train_forward(sd, q.detach(), kk.detach(), v.detach(), truth, ctx_truth, k_top=k, P=P)
    assert abs(float(own.loss.detach()) - float(b_own['loss'])) <= 1e-5 * abs(float(b_own['loss'])) + 1e-6       # with its own top-k rule
    out = training.forward_train(mod, q, kk, v, q, kk, v, q, kk, so.causal_additive_mask(T, torch.float32, N), truth, ctx_truth,
                                 sea.PerlinAttentionOutput, topk_mask_fn=lambda *_: ref_mask)
    assert abs(float(out.loss.detach()) - float(g['loss'])) <= 1e-5 * abs(float(g['loss'])) + 1e-6, (float(out.loss), float(g['loss']))
    torch.testing.assert_close(out.context_layer.detach(), torch.from_numpy(g['context_layer']), rtol=1e-3, atol=2e-5)
    torch.testing.assert_close(out.estimated_attention_probs.detach(), torch.from_numpy(g['estimated_attention_probs']), rtol=1e-3, atol=2e-6)
    torch.testing.assert_close(out.partial_attention_probs.detach(), torch.from_numpy(g['partial_attention_probs']), rtol=1e-3, atol=2e-6)
    out.loss.backward()
    # oracle side: the same leaves, autograd through the oracle's own restatement
    sd_g = {k_: v_.clone().requires_grad_(v_.dtype.is_floating_point) for k_, v_ in sd.items()}
    q2, k2, v2 = (torch.from_numpy(g[x]).clone().requires_grad_(True) for x in 'qkv')
    b = so.perlin_train_forward(sd_g, q2, k2, v2, truth, ctx_truth, k_top=k, P=P, mask_override=ref_mask)
    b['loss'].backward()
    for a_, b_, name in ((q.grad, q2.grad, 'q'), (kk.grad, k2.grad, 'k'), (v.grad, v2.grad, 'v')):
        torch.testing.assert_close(a_, b_, rtol=2e-3, atol=1e-6, msg=lambda s_, n_=name: f'd loss / d {n_}: {s_}')
    checked = 0
    for name, p_ in mod.named_parameters():
        ref = sd_g.get(name)
        if ref is None or ref.grad is None:
            continue
        if name.endswith('.weight') and 'net.' in name and p_.ndim == 4:       # CausalConv2d: the masked-out taps receive no gradient
            mask = sd[name.replace('.weight', '.weight_mask')]
            assert float((p_.grad * (1 - mask)).abs().max()) == 0.0
        assert p_.grad is not None, name
        torch.testing.assert_close(p_.grad, ref.grad, rtol=2e-3, atol=2e-6, msg=lambda s_, n_=name: f'gradient of {n_}: {s_}')
        checked += 1
    assert checked >= 18, checked


def test_eval_mode_with_teacher_gives_the_same_loss(monkeypatch):
    """eval() + teacher tensors (validation-time loss reporting): same forward without the jitter and with the eval softmax."""
    sea, g, m, sd, mod = _grads_case()
    training = importlib.import_module('sea-attention_b200.training')
    N, T, k = m['N'], m['T'], m['k']
    mod.eval()
    q, kk, v = (torch.from_numpy(g[x]) for x in 'qkv')
    with torch.no_grad():
        out = training.forward_train(mod, q, kk, v, q, kk, v, q, kk, so.causal_additive_mask(T, torch.float32, N), torch.from_numpy(g['scores_truth']),
                                     torch.from_numpy(g['context_truth']), sea.PerlinAttentionOutput, topk_mask_fn=_oracle_topk(k))
        out0 = training.forward_train(mod, q, kk, v, q, kk, v, q, kk, None, None, None, sea.PerlinAttentionOutput, topk_mask_fn=_oracle_topk(k))
    b_own = so.perlin_train_forward(sd, q, kk, v, torch.from_numpy(g['scores_truth']), torch.from_numpy(g['context_truth']), k_top=k, P=m['P'])
    assert abs(float(out.loss) - float(b_own['loss'])) <= 1e-5 * abs(float(b_own['loss'])) + 1e-6
    assert float(out0.loss) == 0.0 and out0.estimated_attention_probs is None
    torch.testing.assert_close(out0.context_layer, out.context_layer)
