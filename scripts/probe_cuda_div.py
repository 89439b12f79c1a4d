"""Does torch's CUDA `int tensor / python int` equal IEEE division (CPU) or multiplication by the fp32 reciprocal?"""
import numpy as np, torch
for P in (96, 24, 48, 100, 256):
    L = torch.arange(1, 20001)
    cpu = (L / P).numpy()
    gpu = (L.cuda() / P).cpu().numpy()
    rec = (L.numpy().astype(np.float32) * np.float32(np.float32(1.0) / np.float32(P))).astype(np.float32)
    div = (L.numpy().astype(np.float32) / np.float32(P)).astype(np.float32)
    print(P, 'gpu==cpu', bool((gpu == cpu).all()), 'gpu==ieee_div', bool((gpu == div).all()), 'gpu==L*(1/P)', bool((gpu == rec).all()),
          'cpu==ieee_div', bool((cpu == div).all()))
