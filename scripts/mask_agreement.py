"""Top-k mask agreement of the bf16 production pipeline against the fp32 pipeline (both on the GPU, same weights / inputs)."""
import importlib, sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
N, H, d, T, P, k, nbf = 1, 32, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 256, 64, 8
torch.manual_seed(42)
cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().cuda()
mod.check_padding = False
mod.output_attentions = True
q32 = torch.randn(N, H, T, d, device='cuda') * d ** -0.5; k32 = torch.randn(N, H, T, d, device='cuda'); v32 = torch.randn(N, H, T, d, device='cuda')
qb, kb, vb = q32.bfloat16(), k32.bfloat16(), v32.bfloat16()
am32 = torch.zeros(1, 1, 1, 1, device='cuda').expand(N, 1, T, T)
qf, kf, vf = qb.float(), kb.float(), vb.float()
o32 = mod(qf, kf, vf, qf, kf, vf, qf, kf, am32, None, None)
ob = mod(qb, kb, vb, qb, kb, vb, qb, kb, am32.bfloat16(), None, None)
def dense(o):
    pm = o.partial_attention_mask
    return sea.ops.flat_csr_to_dense(pm, T, H) > 0
a, b = dense(o32), dense(ob)
causal = torch.tril(torch.ones(T, T, dtype=torch.bool, device='cuda'))
agree_all = float((a == b).float().mean())
agree_causal = float(((a == b) & causal).sum() / (causal.sum() * N * H))
iou = float((a & b).sum() / (a | b).sum())
perr = float((o32.estimated_attention_probs - ob.estimated_attention_probs.float()).abs().max())
print(f'T={T}: mask agreement bf16 vs fp32 pipeline: {agree_all:.5f} of all [N,H,T,T] elements, {agree_causal:.5f} of the causal half, IoU of alive sets {iou:.4f}; max |dprobs| {perr:.2e}')
