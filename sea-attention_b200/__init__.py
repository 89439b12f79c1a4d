"""sea-attention_b200: B200-native (sm_100a) implementation of the per-layer SEA / Perlin attention
forward of gmlwns2000/sea-attention, behind the reference's own module / operator interface.

The directory name carries a hyphen (it is the project's name), so import it with
`importlib.import_module('sea-attention_b200')` or through the root-level alias module
`sea_attention_b200`.
"""
from . import _lib, ops
from ._lib import SeaError
from .attention import PerlinAttention, PerlinAttentionOutput, ProjectionUpdater
from .attention_state import PerlinAttentionState
from .config import PerlinAttentionConfig, get_default_config, register_default_config
from .opt_attention import SeaOPTAttention
from .ops import (flat_csr_elmul, flat_csr_masked_bmm, flat_csr_sdbmm, flat_csr_softmax, flat_csr_to_dense,
                  resize_from_m_to_t, resize_from_m_to_t_csr)

__all__ = [
    'PerlinAttention', 'PerlinAttentionOutput', 'PerlinAttentionConfig', 'PerlinAttentionState', 'ProjectionUpdater', 'SeaError', 'SeaOPTAttention',
    'get_default_config', 'register_default_config', 'ops',
    'resize_from_m_to_t', 'resize_from_m_to_t_csr', 'flat_csr_elmul', 'flat_csr_masked_bmm', 'flat_csr_sdbmm',
    'flat_csr_softmax', 'flat_csr_to_dense',
]
