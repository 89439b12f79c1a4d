"""Per-token latency of the decode (use_cache) step at the north-star head shape (development aid)."""
import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, H, d, T, P, k, nbf = 1, 32, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 256, 64, 8
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T + 64)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True, use_cache=True)).eval().cuda()
mod.check_padding = False
mod.freeze_packed_weights()          # inference loop: weights are constant
dt = torch.bfloat16
mk = lambda t: torch.randn(N, H, t, d, device='cuda').to(dt)
q, kk, v = mk(T + 64) * d ** -0.5, mk(T + 64), mk(T + 64)
s = lambda x, a, b: x[:, :, a:b]
mask = torch.zeros(1, 1, 1, 1, device='cuda', dtype=dt).expand(N, 1, T, T)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    o = mod(s(q, 0, T), s(kk, 0, T), s(v, 0, T), s(q, 0, T), s(kk, 0, T), s(v, 0, T), s(q, 0, T), s(kk, 0, T), mask, None, None)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0.record()
    o = mod(s(q, 0, T), s(kk, 0, T), s(v, 0, T), s(q, 0, T), s(kk, 0, T), s(v, 0, T), s(q, 0, T), s(kk, 0, T), mask, None, None)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f'prefill + state build, T={T}: {best * 1000:.0f} us (best of 5)')
state = o.state
dm = torch.zeros(N, 1, 1, 1, device='cuda', dtype=dt)
for t in range(T, T + 8):
    o = mod(s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), dm, None, None, last_state=state)
    state = o.state
torch.cuda.synchronize()
state_start = state
e0.record()
n = 40
for t in range(T + 8, T + 8 + n):
    o = mod(s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), dm, None, None, last_state=state)
    state = o.state
e1.record()
import time
t_host = time.perf_counter()
torch.cuda.synchronize()
print(f'decode step at context {T}: {e0.elapsed_time(e1) / n * 1000:.0f} us per token per layer (one sea_decode_step call per token)')
# host side alone: the same loop enqueued without waiting (the GPU queue absorbs it while it is shorter than the device time)
torch.cuda.synchronize()
state = state_start
t0_ = time.perf_counter()
for t in range(T + 8, T + 8 + n):
    o = mod(s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), s(v, 0, t + 1), s(q, t, t + 1), s(kk, 0, t + 1), dm, None, None, last_state=state)
    state = o.state
t1_ = time.perf_counter()
torch.cuda.synchronize()
t2_ = time.perf_counter()
print(f'host enqueue {(t1_ - t0_) / n * 1e6:.0f} us per token, drain after the loop {(t2_ - t1_) * 1e6:.0f} us')
