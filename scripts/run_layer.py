"""Runs a few full forwards at the north-star shape (for ncu)."""
import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, H, d, T, P, k, nbf = 1, 32, 64, 4096, 256, 64, 8
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().cuda()
mod.check_padding = False
mod.freeze_packed_weights()          # forwards after the first launch exactly the 9 kernels of the layer
dt = torch.bfloat16
q = (torch.randn(N, H, T, d, device='cuda') * d ** -0.5).to(dt); kk = torch.randn(N, H, T, d, device='cuda').to(dt); v = torch.randn(N, H, T, d, device='cuda').to(dt)
mask = torch.zeros(N, 1, T, T, device='cuda', dtype=dt)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(iters):
    out = mod(q, kk, v, q, kk, v, q, kk, mask, None, None)
torch.cuda.synchronize()
print('ok', float(out.context_layer.float().abs().mean()))
