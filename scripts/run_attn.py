"""Runs the attention kernel alone on the real north-star top-k mask (for ncu / A-B timing)."""
import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
ops = sea.ops
N, H, d, T, P, k, nbf = 1, 32, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 256, 64, 8
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().cuda()
dt = torch.bfloat16
q = (torch.randn(N, H, T, d, device='cuda') * d ** -0.5).to(dt); kk = torch.randn(N, H, T, d, device='cuda').to(dt); v = torch.randn(N, H, T, d, device='cuda').to(dt)
w = mod._weights_fp32()
kpr, z_alloc = mod._shape_consts(H, P, T, T, q.device)
ctx, avg = ops.performer_causal(q, kk, v, w['pos'], w['proj'])
cnn_in, scales, _ = ops.predictor_mlp(ctx, v, w, 2, P // 4)
y = ops.causal_conv3x3_dil2_relu(cnn_in, w['conv1_w'], w['conv1_b'])
y = ops.causal_conv3x3_dil2_relu(y, w['conv2_w'], w['conv2_b'])
y3 = ops.conv1x1_umma(y, w['conv3_w'], w['conv3_b'])
probs, bits = ops.predictor_tail_topk(y3, w['conv3_b'], w['out_ln_w'], w['out_ln_b'], kpr, P)
iters = 5
for _ in range(2):
    out = ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    out = ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True)
e1.record(); torch.cuda.synchronize()
print('attn us', e0.elapsed_time(e1) / iters * 1000, 'mean|out|', float(out.float().abs().mean()))
ref = ops.sparse_attention_from_bits(bits, q, kk, v, scales, avg, P, k, True, True, kernel='gather')
dd = (out.float() - ref.float()).abs()
print('vs gather kernel: max abs diff', float(dd.max()), 'mean', float(dd.mean()), 'nan', int(torch.isnan(out.float()).sum()))
