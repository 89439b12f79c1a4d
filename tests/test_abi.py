"""CPU: the C-ABI library builds, loads, and exports every entry point include/sea_b200.h declares;
the product path fails loudly (no fallback) without a GPU."""
import ctypes
import os

import pytest
import torch


def test_library_exports_every_declared_symbol(sea):
    lib = sea._lib.load()
    names = sea._lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/sea_b200.h but not exported'
        assert n in sea._lib._SIGNATURES, f'{n} has no ctypes signature'
    assert lib.sea_abi_version() == 1


def test_header_cites_reference_lines(sea):
    text = open(sea._lib.HEADER_PATH).read()
    for needle in ['causal_resize_m_to_t.py:910', 'flat_csr_masked_bmm.py:137', 'flat_csr_softmax.py:127',
                   'flat_csr_elmul.py:110', 'flat_csr_sdbmm.py:323', 'flat_csr_to_dense.py:3', 'resize_m_to_t.py:6', 'attention.py:']:
        assert needle in text


def test_bad_arguments_return_error_codes(sea):
    lib = sea._lib.load()
    rc = lib.sea_csr_count(None, None, 1, 1, 1, 1, 1, 1, 1, 1, None)
    assert rc == -1
    assert b'null pointer' in lib.sea_last_error()


def test_no_cpu_fallback(sea):
    x = torch.zeros(1, 1, 4, 4)
    with pytest.raises(sea.SeaError):
        sea.resize_from_m_to_t_csr(x, 0, 2)
    import transformers
    m = sea.PerlinAttention(transformers.BertConfig(hidden_size=64, num_attention_heads=2, max_position_embeddings=16),
                            sea.PerlinAttentionConfig(performer_nb_factor=8, k=4, attention_predictor_length=8, causal=True)).eval()
    q = torch.zeros(1, 2, 16, 32)
    with pytest.raises(sea.SeaError):
        m(q, q, q, q, q, q, q, q, torch.zeros(1, 1, 16, 16), None, None)


def test_state_dict_keys_match_reference_fixture(sea):
    """Parameter names equal the reference's (fixture holds the reference state_dict minus unused keys)."""
    from conftest import golden_layer
    import transformers
    g, m, sd = golden_layer('layer_causal_h4_t128')
    mod = sea.PerlinAttention(transformers.BertConfig(hidden_size=m['H'] * m['d'], num_attention_heads=m['H'], max_position_embeddings=m['T']),
                              sea.PerlinAttentionConfig(performer_nb_factor=m['nbf'], k=m['k'], attention_predictor_length=m['P'], causal=True))
    own = mod.state_dict()
    for key, val in sd.items():
        assert key in own, key
        assert tuple(own[key].shape) == tuple(val.shape), key
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert not unexpected


def test_caller_block_has_reference_parameter_names_and_no_cpu_path(sea):
    """SeaOPTAttention keeps the reference's parameter names (perlin_opt.py:300-330, self_attention.py:52) and, like
    every other entry of the package, refuses CPU tensors instead of falling back."""
    import pytest
    import torch
    import transformers
    cfg = transformers.BertConfig(hidden_size=64, num_attention_heads=2, max_position_embeddings=16)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=8, k=4, attention_predictor_length=8, causal=True)
    blk = sea.SeaOPTAttention(64, 2, cfg, pc).eval()
    keys = set(blk.state_dict().keys())
    for name in ('q_proj.weight', 'k_proj.bias', 'v_proj.weight', 'out_proj.weight', 'perlin_self_attention.attention.performer.projection_matrix',
                 'perlin_self_attention.attention.attention_predictor_cnn.1.module.net.5.module.weight'):
        assert name in keys, name
    with pytest.raises(sea.SeaError):
        blk(torch.zeros(1, 16, 64))
    with pytest.raises(sea.SeaError):
        sea.SeaOPTAttention(64, 2, cfg, sea.PerlinAttentionConfig(causal=False))
