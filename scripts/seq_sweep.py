"""Layer forward time vs sequence length (BASELINE configs[3], [4]: linear-in-T check; development aid, not the bench)."""
import importlib, sys, os, json, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
H, d, P, k, nbf = int(sys.argv[1]), int(sys.argv[2]), 256, int(sys.argv[3]), 8
res = []
for T in [int(x) for x in sys.argv[4:]]:
    torch.manual_seed(42)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().cuda()
    mod.check_padding = False
    dt = torch.bfloat16
    q = (torch.randn(1, H, T, d, device='cuda') * d ** -0.5).to(dt); kk = torch.randn(1, H, T, d, device='cuda').to(dt); v = torch.randn(1, H, T, d, device='cuda').to(dt)
    mask = torch.zeros(1, 1, 1, 1, device='cuda', dtype=dt).expand(1, 1, T, T)      # stride-0 view: shape only, the causal structure is implied
    try:
        for _ in range(2): out = mod(q, kk, v, q, kk, v, q, kk, mask, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = mod(q, kk, v, q, kk, v, q, kk, mask, None, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res.append({'T': T, 'ms': round(ms, 3), 'tok_per_s': round(T / ms * 1e3), 'finite': bool(torch.isfinite(out.context_layer.float()).all())})
    except Exception as e:
        res.append({'T': T, 'error': str(e)[:200]})
    del mod, q, kk, v
    torch.cuda.empty_cache()
print(json.dumps({'H': H, 'd': d, 'k': k, 'P': P, 'results': res}))
