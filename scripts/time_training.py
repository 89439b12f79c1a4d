"""One layer forward + backward of the TRAINING branch at the BASELINE configs[2] shape (OPT-1.3B: H32 d64, T=2048, batch N per GPU), bf16
parameters and inputs, teacher tensors given: N H d T P k [iters]."""
import importlib, json, os, sys
import torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')


def main(N=8, H=32, d=64, T=2048, P=256, k=64, iters=3):
    dev = 'cuda:0'
    torch.manual_seed(42)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=8, k=k, attention_predictor_length=P, causal=True)).to(dev).bfloat16().train()
    dt = torch.bfloat16
    q = (torch.randn(N, H, T, d, device=dev) * d ** -0.5).to(dt).requires_grad_(True)
    kk = torch.randn(N, H, T, d, device=dev).to(dt).requires_grad_(True)
    v = torch.randn(N, H, T, d, device=dev).to(dt).requires_grad_(True)
    truth = torch.randn(N, H, T, T, device=dev, dtype=dt)
    ctx_truth = torch.randn(N, T, H * d, device=dev, dtype=dt)

    def step():
        for t_ in (q, kk, v):
            t_.grad = None
        mod.zero_grad(set_to_none=True)
        out = mod(q, kk, v, q, kk, v, q, kk, None, truth, ctx_truth)
        out.loss.float().backward()
        return out.loss

    loss = step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({'workload': f'training branch, one layer fwd+bwd, N={N} H={H} d={d} T={T} P={P} k={k}, bf16', 'ms': round(ms, 2),
                      'tokens_per_s': round(N * T / ms * 1e3), 'loss': float(loss), 'finite_grads': bool(torch.isfinite(q.grad.float()).all()),
                      'peak_mem_gb': round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}))


if __name__ == '__main__':
    main(*[int(a) for a in sys.argv[1:]])
