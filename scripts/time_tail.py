import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200'); ops = sea.ops
N, H, T, W, P, k = 1, 32, 4096, 64, 256, 64
torch.manual_seed(0)
y3 = torch.randn(N, T, W, H, device='cuda'); bias = torch.randn(H, device='cuda'); lw = torch.ones(P, device='cuda'); lb = torch.zeros(P, device='cuda')
tl_ = torch.arange(1, T + 1, device='cuda'); kpr = torch.clamp_min(torch.round(H * ((k * 1.0 * P) / tl_)), 1).float()
def timeit(name, fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); print(name, round(e0.elapsed_time(e1) / it * 1000, 1), 'us')
timeit('probs only', lambda: ops.predictor_tail_topk(y3, bias, lw, lb, kpr, P, want_probs=True, want_bits=False))
timeit('bits only (no probs store)', lambda: ops.predictor_tail_topk(y3, bias, lw, lb, kpr, P, want_probs=False, want_bits=True))
timeit('probs+bits', lambda: ops.predictor_tail_topk(y3, bias, lw, lb, kpr, P))
timeit('probs+bits+count', lambda: ops.predictor_tail_topk(y3, bias, lw, lb, kpr, P, count_k=k))
probs, bits = ops.predictor_tail_topk(y3, bias, lw, lb, kpr, P)
timeit('standalone topk', lambda: ops.topk_mask_bits(probs, kpr, 'causal_batch'))
x = torch.randn(N, T, W, 64, device='cuda').bfloat16(); w3 = torch.randn(32, 64, device='cuda'); b3 = torch.randn(32, device='cuda')
timeit('conv1x1_umma', lambda: ops.conv1x1_umma(x, w3, b3))
