"""ctypes binding of libsea_b200.so (C ABI: include/sea_b200.h).

There is NO fallback: if the library is missing or an entry fails, the call raises.  PyTorch is only
used by the callers for device memory and streams; every pointer crossing this boundary is a raw
`data_ptr()`.
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'lib', 'libsea_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_PKG), 'include', 'sea_b200.h')

SEA_DTYPE_F32, SEA_DTYPE_BF16, SEA_DTYPE_F16 = 0, 1, 2

_lib = None


class SeaError(RuntimeError):
    pass


def declared_symbols():
    """Names of every SEA_API entry point the header declares."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r'SEA_API\s+[\w\s\*]+?\b(sea_\w+)\s*\(', text)))


_c = ctypes
_P, _I, _L, _F = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float

_SIGNATURES = {
    'sea_abi_version': (_I, []),
    'sea_last_error': (_c.c_char_p, []),
    'sea_device_arch': (_I, []),
    'sea_topk_mask_bits': (_I, [_P, _L, _L, _L, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_mask_float_to_bits': (_I, [_P, _L, _L, _L, _P, _I, _I, _I, _I, _P]),
    'sea_mask_bits_to_float': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'sea_csr_count': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_csr_fill': (_I, [_P, _P, _P, _I, _L, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_csr_count_len': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'sea_csr_fill_len': (_I, [_P, _P, _P, _I, _L, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'sea_flat_csr_to_dense': (_I, [_P, _P, _I, _P, _L, _P, _I, _I, _I, _I, _P]),
    'sea_flat_csr_masked_bmm': (_I, [_P, _P, _I, _L, _P, _L, _L, _L, _P, _L, _L, _L, _I, _P, _I, _I, _I, _I, _I, _P]),
    'sea_flat_csr_softmax': (_I, [_P, _P, _I, _L, _P, _P, _I, _I, _I, _I, _P]),
    'sea_flat_csr_elmul': (_I, [_P, _P, _I, _L, _P, _P, _P, _L, _L, _L, _L, _I, _I, _I, _I, _P]),
    'sea_flat_csr_sdbmm': (_I, [_P, _P, _I, _L, _P, _P, _L, _L, _L, _I, _P, _I, _I, _I, _I, _I, _P]),
    'sea_resize_m_to_t_dense': (_I, [_P, _F, _P, _L, _L, _P, _I, _I, _I, _I, _I, _P]),
    'sea_performer_workspace_floats': (_L, [_I, _I, _I, _I, _I]),
    'sea_performer_causal_fwd': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_predictor_mlp_fwd': (_I, [_P, _P, _L, _L, _L, _I] + [_P] * 10 + [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_causal_conv3x3_dil2_relu': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_performer_mma_supported': (_I, [_I, _I, _I]),
    'sea_performer_mma_workspace_floats': (_L, [_I, _I, _I, _I, _I]),
    'sea_performer_causal_mma_fwd': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_performer_mma_state_floats': (_L, [_I, _I, _I, _I]),
    'sea_performer_causal_mma_range': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_predictor_mlp_umma_supported': (_I, [_I, _I, _I, _I, _I]),
    'sea_predictor_mlp_umma_workspace_bytes': (_L, []),
    'sea_predictor_mlp_umma_fwd': (_I, [_P, _P, _L, _L, _L] + [_P] * 10 + [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_predictor_mlp_umma_fwd_ex': (_I, [_P, _P, _L, _L, _L] + [_P] * 10 + [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_predictor_mlp_mma_supported': (_I, [_I, _I, _I, _I, _I]),
    'sea_predictor_mlp_mma_workspace_bytes': (_L, [_I, _I, _I]),
    'sea_predictor_mlp_mma_fwd': (_I, [_P, _P, _L, _L, _L] + [_P] * 10 + [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_conv_umma_supported': (_I, [_I, _I, _I, _I]),
    'sea_conv_umma_workspace_bytes': (_L, [_I, _I]),
    'sea_causal_conv3x3_dil2_relu_umma': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_conv1x1_umma': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_conv3x3_conv1x1_umma_supported': (_I, [_I, _I, _I, _I, _I]),
    'sea_causal_conv3x3_dil2_relu_conv1x1_umma': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_predictor_tail_topk_fwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_crow_scan': (_I, [_P, _I, _I, _I, _P]),
    'sea_predictor_tail_fwd': (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_sparse_attention_bits_fwd': (_I, [_P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _L, _L, _I, _I, _P,
                                           _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_block_attention_workspace_bytes': (_L, [_I, _I, _I, _I, _I, _I, _I, _I]),
    'sea_block_attention_fwd': (_I, [_P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _L, _L, _I, _I, _P,
                                     _I, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P]),
    'sea_sparse_attention_bits_bwd': (_I, [_P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _L, _L, _I, _I,
                                           _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_performer_state_floats': (_L, [_I, _I, _I, _I]),
    'sea_performer_state_build': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    'sea_performer_causal_state_fwd': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_performer_noncausal_workspace_floats': (_L, [_I, _I, _I, _I, _I]),
    'sea_performer_noncausal_fwd': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_performer_noncausal_len_fwd': (_I, [_P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_conv3x3_cl': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sea_bert_tail_fwd': (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sea_topk_mask_bits_batch': (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    'sea_topk_batch_workspace_bytes': (_L, [_I, _I, _I, _I, _I]),
    'sea_topk_mask_bits_batch_ws': (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P]),
    'sea_bert_avg_fwd': (_I, [_P, _P, _L, _L, _L, _I, _P, _I, _I, _I, _I, _I, _P]),
    'sea_debug_attn_trace_read': (_L, [_P, _L]),
    'sea_decode_step_workspace_bytes': (_L, [_I] * 8),
    'sea_decode_step': (_I, [_P, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _I] + [_P] * 20 + [_P, _P, _P, _I, _P] + [_P] * 6 + [_P, _P, _P, _L]
                        + [_I] * 10 + [_P]),
    'sea_bert_avg_len_fwd': (_I, [_P, _P, _L, _L, _L, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sea_sparse_attention_fwd': (_I, [_P, _P, _I, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _L, _L, _I, _I, _P, _P, _P,
                                      _I, _I, _I, _I, _I, _P]),
}


def load():
    """Loads the shared library (once).  Raises SeaError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SeaError(f'{LIB_PATH} not found: build it with `python sea-attention_b200/build.py` '
                       f'(or __graft_entry__.build()); there is no fallback path')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SeaError(f'{LIB_PATH} does not export {name}') from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# Kernel launches each entry enqueues (bench.py reports the sum as `gpu_launches`).
KERNELS_PER_CALL = {
    'sea_topk_mask_bits': 1, 'sea_mask_float_to_bits': 1, 'sea_mask_bits_to_float': 1, 'sea_csr_count': 2, 'sea_csr_fill': 1, 'sea_csr_count_len': 2, 'sea_csr_fill_len': 1, 'sea_performer_noncausal_len_fwd': 4, 'sea_bert_avg_len_fwd': 1, 'sea_decode_step': 9, 'sea_crow_scan': 1,
    'sea_flat_csr_to_dense': 1, 'sea_flat_csr_masked_bmm': 1, 'sea_flat_csr_softmax': 1, 'sea_flat_csr_elmul': 1,
    'sea_flat_csr_sdbmm': 1, 'sea_resize_m_to_t_dense': 1, 'sea_performer_causal_fwd': 3, 'sea_performer_causal_mma_fwd': 3, 'sea_performer_causal_mma_range': 2, 'sea_predictor_mlp_fwd': 1,
    'sea_causal_conv3x3_dil2_relu': 1, 'sea_causal_conv3x3_dil2_relu_umma': 2, 'sea_conv1x1_umma': 2, 'sea_causal_conv3x3_dil2_relu_conv1x1_umma': 3, 'sea_predictor_mlp_umma_fwd': 2, 'sea_predictor_mlp_umma_fwd_ex': 2, 'sea_predictor_mlp_mma_fwd': 2, 'sea_predictor_tail_topk_fwd': 1, 'sea_predictor_tail_fwd': 1, 'sea_sparse_attention_fwd': 1, 'sea_sparse_attention_bits_fwd': 1, 'sea_block_attention_fwd': 1, 'sea_sparse_attention_bits_bwd': 2, 'sea_performer_noncausal_fwd': 4, 'sea_conv3x3_cl': 1, 'sea_bert_tail_fwd': 1,
    'sea_topk_mask_bits_batch': 1, 'sea_topk_mask_bits_batch_ws': 12, 'sea_bert_avg_fwd': 1,
}
LAUNCH_COUNT = 0
TRACE = None   # bench.py sets this to a list to get (entry, start_event, stop_event) per call


def call(name, *args, kernels=None):
    """Calls an int-returning entry; raises SeaError with the library's message on failure.
    `kernels` overrides the entry's default launch count (e.g. a call that reuses a cached weight packing)."""
    global LAUNCH_COUNT
    lib = load()
    LAUNCH_COUNT += KERNELS_PER_CALL.get(name, 1) if kernels is None else kernels
    if TRACE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        TRACE.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.sea_last_error()
        raise SeaError(f'{name} failed (code {rc}): {msg.decode() if msg else ""}')
    return rc
