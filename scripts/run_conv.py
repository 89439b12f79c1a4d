"""Runs only the tcgen05 conv kernel at the north-star shape (for ncu)."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, T, W, C = 1, 4096, 64, 64
torch.manual_seed(0)
x = torch.randn(N, T, W, C, device='cuda').bfloat16()
w = torch.zeros(64, 64, 5, 3, device='cuda'); w[:, :, :3] = torch.randn(64, 64, 3, 3, device='cuda') * 0.05
b = torch.randn(64, device='cuda') * 0.1
for _ in range(3):
    y = sea.ops.causal_conv3x3_dil2_relu(x, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = sea.ops.causal_conv3x3_dil2_relu(x, w, b)
e1.record(); torch.cuda.synchronize()
print('conv umma us/iter', e0.elapsed_time(e1) * 100)
