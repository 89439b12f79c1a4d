import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200'); ops = sea.ops
N, H, d, T, P, k, nbf = 1, 32, 64, 4096, 256, 64, 8
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().cuda()
dt = torch.bfloat16
q = (torch.randn(N, H, T, d, device='cuda') * d ** -0.5).to(dt); kk = torch.randn(N, H, T, d, device='cuda').to(dt); v = torch.randn(N, H, T, d, device='cuda').to(dt)
w = mod._weights_fp32()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(name, fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(it):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    print(name, round(tot / it * 1000, 1), 'us (L2 flushed)')
timeit('performer', lambda: ops.performer_causal(q, kk, v, w['pos'], w['proj']))
ctx, avg = ops.performer_causal(q, kk, v, w['pos'], w['proj'])
timeit('mlp', lambda: ops.predictor_mlp(ctx, v, w, 2, P // 4, packed=mod._packed))
