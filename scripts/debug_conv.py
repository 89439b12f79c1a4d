"""Development aid: per-tap identity weights through the tcgen05 causal conv; reports which (dt, dw) shift each tap actually applies."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, T, W, C = 1, 12, 64, 64
torch.manual_seed(0)
x = torch.rand(N, T, W, C, device='cuda').bfloat16() + 0.5
b = torch.zeros(64, device='cuda')
xf = x.float()
def shifted(dt, dw):
    out = torch.zeros_like(xf)
    ts = slice(max(0, -dt), min(T, T - dt)); td = slice(max(0, dt), min(T, T + dt))
    ws = slice(max(0, -dw), min(W, W - dw)); wd = slice(max(0, dw), min(W, W + dw))
    # out[t, w] = x[t - dt, w - dw]
    out[:, td, wd] = xf[:, ts, ws]
    return out
for i in range(3):
    for j in range(3):
        w = torch.zeros(64, 64, 5, 3, device='cuda')
        w[:, :, i, j] = torch.eye(64, device='cuda')
        y = sea.ops.causal_conv3x3_dil2_relu(x, w, b).float()
        exp_dt, exp_dw = 2 * (2 - i), -2 * (j - 1)
        ok = torch.allclose(y, shifted(exp_dt, exp_dw), atol=1e-2)
        best = None
        if not ok:
            for dt in range(0, 7):
                for dw in range(-6, 7):
                    err = (y - shifted(dt, dw)).abs().mean().item()
                    if best is None or err < best[0]: best = (err, dt, dw)
        print(f'tap i={i} j={j}: expect out[t,w]=x[t-{exp_dt}, w-({exp_dw})] ok={ok} best={best}')
