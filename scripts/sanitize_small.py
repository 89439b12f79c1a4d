"""Small end-to-end exercise of every kernel family on ragged shapes, run natively (a plain smoke of the paths: bf16 prefill on the
tcgen05 path, block attention in both variants, backward, decode steps, BERT layer, the other head dims, query blocks, the DEEPER
predictor, the caller block).  compute-sanitizer is not available on the GPU pool, so out-of-bounds WRITES are hunted by
tests/test_guard_bytes_gpu.py instead (outputs carved out of sentinel-filled buffers, guard bytes checked after the launch) and reads by
the comparisons with the CPU oracle on the same ragged shapes."""
import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
DEV = 'cuda:0'
torch.manual_seed(0)
N, H, d, T, P, k, nbf = 1, 32, 64, 200, 64, 16, 8
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T + 8)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV)
mk = lambda t, s=1.0: (torch.randn(N, H, t, d, device=DEV) * s).bfloat16()
q, kk, v = mk(T + 4, d ** -0.5), mk(T + 4), mk(T + 4)
sl = lambda x, a, b: x[:, :, a:b]
am = torch.zeros(N, 1, T, T, device=DEV, dtype=torch.bfloat16)
o = mod(sl(q, 0, T), sl(kk, 0, T), sl(v, 0, T), sl(q, 0, T), sl(kk, 0, T), sl(v, 0, T), sl(q, 0, T), sl(kk, 0, T), am, None, None)
os.environ['SEA_ATTN_MMA_SYNC'] = '1'
qh = sl(q, 0, T).half(); kh = sl(kk, 0, T).half(); vh = sl(v, 0, T).half()
o2 = mod(qh, kh, vh, qh, kh, vh, qh, kh, am.half(), None, None)          # fp16 -> mma.sync block kernel
qg = sl(q, 0, T).clone().requires_grad_(True); kg = sl(kk, 0, T).clone().requires_grad_(True); vg = sl(v, 0, T).clone().requires_grad_(True)
o3 = mod(qg, kg, vg, qg, kg, vg, qg, kg, am, None, None)
o3.context_layer.float().sum().backward()
mod.pconfig.use_cache = True
o4 = mod(sl(q, 0, T), sl(kk, 0, T), sl(v, 0, T), sl(q, 0, T), sl(kk, 0, T), sl(v, 0, T), sl(q, 0, T), sl(kk, 0, T), am, None, None)
st = o4.state
for t in range(T, T + 3):
    o5 = mod(sl(q, t, t + 1), sl(kk, 0, t + 1), sl(v, 0, t + 1), sl(q, t, t + 1), sl(kk, 0, t + 1), sl(v, 0, t + 1), sl(q, t, t + 1), sl(kk, 0, t + 1),
             torch.zeros(N, 1, 1, 1, device=DEV, dtype=torch.bfloat16), None, None, last_state=st)
    st = o5.state
# OPT-125m-like head count (padded channels, partial head slots)
H2 = 12
cfg2 = transformers.BertConfig(hidden_size=H2 * d, num_attention_heads=H2, max_position_embeddings=T)
mod2 = sea.PerlinAttention(cfg2, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV)
q2 = (torch.randn(N, H2, T, d, device=DEV) * d ** -0.5).bfloat16(); k2 = torch.randn(N, H2, T, d, device=DEV).bfloat16(); v2 = torch.randn(N, H2, T, d, device=DEV).bfloat16()
o6 = mod2(q2, k2, v2, q2, k2, v2, q2, k2, am, None, None)
# BERT layer
mod3 = sea.PerlinAttention(cfg2, sea.PerlinAttentionConfig(performer_nb_factor=1, k=k, attention_predictor_length=P, causal=False, k_flatten_dim='batch')).eval().to(DEV)
o7 = mod3(q2, k2, v2, q2, k2, v2, q2, k2, torch.zeros(N, 1, 1, T, device=DEV, dtype=torch.bfloat16), None, None)
# other head dims: slab Performer, mma.sync MLP, gather attention (d = 80: 10 pieces per row), query block, DEEPER predictor, caller block
outs = []
for (H4, d4, T4, P4, k4) in ((32, 80, 150, 128, 16), (32, 128, 140, 256, 32), (8, 96, 70, 128, 8), (4, 32, 90, 128, 8)):
    cfg4 = transformers.BertConfig(hidden_size=H4 * d4, num_attention_heads=H4, max_position_embeddings=T4)
    m4 = sea.PerlinAttention(cfg4, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k4, attention_predictor_length=P4, causal=True)).eval().to(DEV)
    q4 = (torch.randn(N, H4, T4, d4, device=DEV) * d4 ** -0.5).bfloat16(); k4_ = torch.randn(N, H4, T4, d4, device=DEV).bfloat16(); v4 = torch.randn(N, H4, T4, d4, device=DEV).bfloat16()
    outs.append(m4(q4, k4_, v4, q4, k4_, v4, q4, k4_, None, None, None))
    outs.append(m4.forward_query_block(q4, k4_, v4, T4 // 3, T4 - 5))
os.environ['PERLIN_HOTFIX_OPT_DEEPER'] = '1'
m5 = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV)
del os.environ['PERLIN_HOTFIX_OPT_DEEPER']
outs.append(m5(sl(q, 0, T), sl(kk, 0, T), sl(v, 0, T), sl(q, 0, T), sl(kk, 0, T), sl(v, 0, T), sl(q, 0, T), sl(kk, 0, T), None, None, None))
blk = sea.SeaOPTAttention(H * d, H, cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(DEV).bfloat16()
x = torch.randn(1, T, H * d, device=DEV).bfloat16()
y, _, past = blk(x, use_cache=True)
y2, _, past = blk(torch.randn(1, 1, H * d, device=DEV).bfloat16(), past_key_value=past, use_cache=True)
torch.cuda.synchronize()
print('new paths ok', [float(x_.context_layer.float().abs().mean()) for x_ in outs], float(y.float().abs().mean()), float(y2.float().abs().mean()))
torch.cuda.synchronize()
print('ok', [float(x.context_layer.float().abs().mean()) for x in (o, o2, o3, o5, o6, o7)])
