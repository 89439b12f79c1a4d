"""cProfile of the host side of the decode loop (development aid): where the per-token python / ctypes / launch time goes."""
import cProfile, pstats, importlib, sys, os, io, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, H, d, T, P, k, nbf = 1, 32, 64, 1024, 256, 64, 8
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T + 512)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True, use_cache=True)).eval().cuda()
mod.check_padding = False
mod.freeze_packed_weights()
dt = torch.bfloat16
mk = lambda t: torch.randn(N, H, t, d, device='cuda').to(dt)
q, kk, v = mk(T + 512) * d ** -0.5, mk(T + 512), mk(T + 512)
s = lambda x, a, b: x[:, :, a:b]
o = mod(s(q, 0, T), s(kk, 0, T), s(v, 0, T), s(q, 0, T), s(kk, 0, T), s(v, 0, T), s(q, 0, T), s(kk, 0, T), None, None, None)
state = o.state
dm = torch.zeros(N, 1, 1, 1, device='cuda', dtype=dt)
def loop(t0, n):
    global state
    for t in range(t0, t0 + n):
        o = mod(s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), dm, None, None, last_state=state)
        state = o.state
loop(T, 20)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
loop(T + 20, 200)
pr.disable()
torch.cuda.synchronize()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats('tottime').print_stats(18)
print(out.getvalue())
