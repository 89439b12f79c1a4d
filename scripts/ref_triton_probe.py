"""Probe: does the reference's compiled Triton path run on this box?  Prints the full error if not."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_harness as rh
rh.load_reference(interpret_triton=False)
import src.models.perlin_attention.ops as ops
try:
    x = torch.ones(1, 1, 4, 2, device='cuda')
    r = ops.resize_from_m_to_t_csr(x, 0, 2, target_width=4, is_causal=True, oversampled=1.0)
    print('ok', r.crow_indices(), r.col_indices())
except Exception as e:
    traceback.print_exc()
    print('MSG', str(e)[-3000:])
