"""CUDA-event time of one PerlinAttention.forward (eager, frozen packings) at a given shape: N H d T P k nbf [iters]."""
import importlib, json, os, sys
import torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')


def main(N=1, H=32, d=128, T=4096, P=256, k=128, nbf=8, iters=5):
    dev = 'cuda:0'
    torch.manual_seed(42)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval().to(dev)
    mod.check_padding = False
    mod.freeze_packed_weights()
    dt = torch.bfloat16
    q = (torch.randn(N, H, T, d, device=dev) * d ** -0.5).to(dt); kk = torch.randn(N, H, T, d, device=dev).to(dt); v = torch.randn(N, H, T, d, device=dev).to(dt)
    trace = []
    with torch.no_grad():
        for _ in range(2):
            out = mod(q, kk, v, q, kk, v, q, kk, None, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = mod(q, kk, v, q, kk, v, q, kk, None, None, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        sea._lib.TRACE = trace
        out = mod(q, kk, v, q, kk, v, q, kk, None, None, None)
        torch.cuda.synchronize()
        sea._lib.TRACE = None
    per = {}
    for name, a, b in trace:
        per[name] = per.get(name, 0.0) + a.elapsed_time(b) * 1000
    print(json.dumps({'shape': [N, H, d, T, P, k, nbf], 'ms': round(ms, 4), 'tok_per_s': round(N * T / ms * 1000), 'finite': bool(torch.isfinite(out.context_layer.float()).all()),
                      'entries_us': {k_: round(v_, 1) for k_, v_ in per.items()}}))


if __name__ == '__main__':
    main(*[int(a) for a in sys.argv[1:]])
