"""BERT-base (non-causal) PerlinAttention layer forward timing: BASELINE configs[0] shape (H12 d64 T512 k64 P128 nbf1)."""
import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, H, d, T, P, k, nbf = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 12, 64, 512, 128, 64, 1
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=False, k_flatten_dim='batch')).eval().cuda()
mod.check_padding = False
for dt in (torch.float32, torch.bfloat16):
    q = (torch.randn(N, H, T, d, device='cuda') * d ** -0.5).to(dt); kk = torch.randn(N, H, T, d, device='cuda').to(dt); v = torch.randn(N, H, T, d, device='cuda').to(dt)
    mask = torch.zeros(N, 1, 1, T, device='cuda', dtype=dt)
    try:
        sea._lib.TRACE = None
        for _ in range(3): out = mod(q, kk, v, q, kk, v, q, kk, mask, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = mod(q, kk, v, q, kk, v, q, kk, mask, None, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        sea._lib.TRACE = []
        out = mod(q, kk, v, q, kk, v, q, kk, mask, None, None)
        torch.cuda.synchronize()
        per = {}
        for name, a, b in sea._lib.TRACE:
            per[name] = per.get(name, 0) + a.elapsed_time(b) * 1000
        sea._lib.TRACE = None
        print(dt, f'{ms*1000:.0f} us/layer-forward ({N*T/ms*1e3:.0f} tok/s)', {k_: round(v_) for k_, v_ in per.items()})
    except Exception as e:
        print(dt, 'error', str(e)[:300])
