"""Per-stage CUDA-event timing of the causal forward at a given shape (development aid)."""
import importlib, sys, os, json
import torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sea = importlib.import_module('sea-attention_b200')
ops = sea.ops

def main(N=1, H=32, d=64, T=4096, P=256, k=64, nbf=8, dtype='bf16', iters=5):
    dt = {'bf16': torch.bfloat16, 'fp32': torch.float32}[dtype]
    dev = 'cuda:0'
    torch.manual_seed(42)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval().to(dev)
    q = (torch.randn(N, H, T, d, device=dev) * d ** -0.5).to(dt); kk = torch.randn(N, H, T, d, device=dev).to(dt); v = torch.randn(N, H, T, d, device=dev).to(dt)
    w = mod._weights_fp32()
    S, W = 2, P // 4
    kpr, z_alloc = mod._shape_consts(H, P, T, T, q.device)
    kpr = kpr.repeat(N)
    res = {}
    def timed(name, fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): out = fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / iters * 1000
        return out
    ctx, avg = timed('performer', lambda: ops.performer_causal(q, kk, v, w['pos'], w['proj']))
    cnn_in, scales, _ = timed('mlp', lambda: ops.predictor_mlp(ctx, v, w, S, W))
    y1 = timed('conv1', lambda: ops.causal_conv3x3_dil2_relu(cnn_in, w['conv1_w'], w['conv1_b']))
    y2 = timed('conv2', lambda: ops.causal_conv3x3_dil2_relu(y1, w['conv2_w'], w['conv2_b']))
    probs, _ = timed('tail', lambda: ops.predictor_tail(y2, w['conv3_w'], w['conv3_b'], w['out_ln_w'], w['out_ln_b'], P))
    bits = timed('topk', lambda: ops.topk_mask_bits(probs, kpr, 'causal_batch'))
    crow, col, Z = timed('csr', lambda: ops.csr_from_bits(bits, H, P, k, T, True, torch.int32, z_alloc))
    out, _ = timed('attn', lambda: ops.sparse_attention(crow, col, q, kk, v, scales, avg, True, False))
    mask = torch.zeros(N, 1, T, T, device=dev, dtype=dt)
    mod.check_padding = False
    timed('module_total', lambda: mod(q, kk, v, q, kk, v, q, kk, mask, None, None))
    res['nnz'] = int(crow[0, -1]); res['z_alloc'] = z_alloc
    print(json.dumps({'shape': [N, H, d, T, P, k, nbf, dtype], 'us': {k_: round(v_, 1) if isinstance(v_, float) else v_ for k_, v_ in res.items()}}))

if __name__ == '__main__':
    args = [int(a) if a.isdigit() else a for a in sys.argv[1:]]
    main(*args)
