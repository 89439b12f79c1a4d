// One decode step (use_cache, T_new = 1) of the causal layer as ONE C-ABI call: the ~10 kernels of a step are enqueued
// back to back from native code, with every intermediate in a caller-provided workspace -- no per-kernel python dispatch, no tensor
// allocation, no torch.cat.  Same kernels and same results as the per-op python sequence (attention.py::_forward_causal_stateful):
//   incremental Performer + running mean of v (StatefulCausalPerformer / StatefulCumAvg, attention_state.py:43-98, 205-224)
//   -> predictor MLP on the new token -> CNN input window (last 4 rows + new row) -> dilated conv 1 on the 5-row window -> its
//   window -> dilated conv 2 (windowed CNN of attention_state.py:142-187) -> 1x1 conv + upsample + area resize + LayerNorm + softmax
//   -> grouped top-k of the one query row (attention.py:774-947) -> sparse attention of that row over the whole KV cache.
// The state is functional like the reference's (the caller's old state is left untouched): state_in buffers are read, state_out
// buffers written.
#include "common.cuh"

namespace sea {
namespace {

// The window bookkeeping of a step is a handful of small strided copies; each group of them is ONE launch (a cudaMemcpy2DAsync per
// copy costs more in launch latency than the copy itself).  A job copies `width` 4-byte words for every (n, r): dst + n d_sn + r d_sr
// <- src + n s_sn + r s_sr (strides in words).
struct CopyJob {
    uint32_t* dst;
    const uint32_t* src;
    int64_t d_sn, d_sr, s_sn, s_sr;
    int N, R, width;
};
struct CopyJobs {
    CopyJob j[3];
    int count;
    float* bcast_dst;           // optional: bcast_dst[0 .. bcast_n) = bcast_src[0]
    const float* bcast_src;
    int bcast_n;
};

__global__ void __launch_bounds__(256)
decode_copy_kernel(CopyJobs jobs) {
    const int64_t tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t) gridDim.x * blockDim.x;
    for (int q = 0; q < jobs.count; ++q) {
        const CopyJob jb = jobs.j[q];
        const int64_t per_n = (int64_t) jb.R * jb.width, total = per_n * jb.N;
        for (int64_t i = tid; i < total; i += nth) {
            const int n = (int) (i / per_n);
            const int64_t rem = i - (int64_t) n * per_n;
            const int r = (int) (rem / jb.width), c = (int) (rem - (int64_t) r * jb.width);
            jb.dst[n * jb.d_sn + r * jb.d_sr + c] = jb.src[n * jb.s_sn + r * jb.s_sr + c];
        }
    }
    if (jobs.bcast_dst != nullptr)
        for (int64_t i = tid; i < jobs.bcast_n; i += nth) jobs.bcast_dst[i] = jobs.bcast_src[0];
}

inline CopyJob make_job(void* dst, const void* src, int64_t d_sn_b, int64_t d_sr_b, int64_t s_sn_b, int64_t s_sr_b, int N, int R, int64_t width_b) {
    CopyJob j;
    j.dst = reinterpret_cast<uint32_t*>(dst); j.src = reinterpret_cast<const uint32_t*>(src);
    j.d_sn = d_sn_b >> 2; j.d_sr = d_sr_b >> 2; j.s_sn = s_sn_b >> 2; j.s_sr = s_sr_b >> 2;
    j.N = N; j.R = R; j.width = (int) (width_b >> 2);
    return j;
}

inline cudaError_t launch_copies(const CopyJobs& jobs, cudaStream_t s) {
    int64_t words = jobs.bcast_n;
    for (int q = 0; q < jobs.count; ++q) words += (int64_t) jobs.j[q].N * jobs.j[q].R * jobs.j[q].width;
    const int blocks = (int) ((words + 1023) / 1024 < 1 ? 1 : ((words + 1023) / 1024 > 296 ? 296 : (words + 1023) / 1024));
    decode_copy_kernel<<<blocks, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t) 255; }

// The 1x1 convolution of the one new row (attention.py:275-280 via modules.py:96-192 with a 1x1 kernel): y3[n][w][h] = bias[h] +
// sum_c x[n][w][c] * weight[h][c], fp32 out -- the input the fused tail / top-k kernel of the prefill takes.  8 positions per CTA; the
// weight goes through shared memory transposed ([c][h], rows padded by one word) so that both its store and its reads are conflict-free.
template <typename T>
__global__ void __launch_bounds__(256)
decode_conv1x1_kernel(const T* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ y3,
                      int W, int C, int H) {
    extern __shared__ float dc_smem[];
    float* wT = dc_smem;                    // [C][H + 1]
    float* xs = dc_smem + C * (H + 1);      // [8][C]
    const int n = blockIdx.y, w0 = blockIdx.x * 8, tid = threadIdx.x;
    for (int i = tid; i < H * C; i += 256) {
        const int h = i / C, c = i - h * C;
        wT[c * (H + 1) + h] = __ldg(weight + i);
    }
    for (int i = tid; i < 8 * C; i += 256) {
        const int wl = i / C, c = i - wl * C;
        xs[i] = w0 + wl < W ? to_f32(x[((int64_t) n * W + w0 + wl) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int o = tid; o < 8 * H; o += 256) {
        const int wl = o / H, h = o - wl * H;
        if (w0 + wl >= W) continue;
        float acc = __ldg(bias + h);
        const float* xr = xs + wl * C;
        for (int c = 0; c < C; ++c) acc = fmaf(xr[c], wT[c * (H + 1) + h], acc);
        y3[((int64_t) n * W + w0 + wl) * H + h] = acc;
    }
}

struct DecodeWs {
    int64_t ctx, cumavg, cnn_row, xwin5, y1full, ywin5, y2full, y2row, y3row, scales, bits, kpr, crow, col, head_ptr, total;
    int64_t z_alloc;
};

DecodeWs decode_ws(int N, int H, int D, int P, int S, int C, int k_clamp, int esz) {
    const int W = P / 4;
    DecodeWs w;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { const int64_t at = o; o += align256(bytes); return at; };
    w.ctx = take((int64_t) N * H * 2 * D * esz);
    w.cumavg = take((int64_t) N * H * D * esz);
    w.cnn_row = take((int64_t) N * W * C * esz);
    w.xwin5 = take((int64_t) N * 5 * W * C * esz);
    w.y1full = take((int64_t) N * 5 * W * C * esz);
    w.ywin5 = take((int64_t) N * 5 * W * C * esz);
    w.y2full = take((int64_t) N * 5 * W * C * esz);
    w.y2row = take((int64_t) N * W * S * H * esz);
    w.y3row = take((int64_t) N * W * H * 4);
    w.scales = take((int64_t) N * H * 2 * 4);
    w.bits = take((int64_t) N * ((H * P + 31) / 32) * 4);
    w.kpr = take((int64_t) N * 4);
    // CSR of one query row per item (only the fp32 path materialises it).  K_t = round(H k P / L) pixels, each min(floor(L/P) + 1, k) wide:
    // nnz <= H k (1 + P / L) <= 2 H k for L >= P, and <= H L < H P for L < P (pixels at most one token wide)
    w.z_alloc = (int64_t) H * (2 * k_clamp > P ? 2 * k_clamp : P) + k_clamp + 64;      // + one pixel (<= k wide) for the rounding of K_t
    w.crow = take((int64_t) N * 2 * 4);
    w.col = take((int64_t) N * w.z_alloc * 4);
    w.head_ptr = take((int64_t) N * (H + 1) * 4);
    w.total = o;
    return w;
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" {

int64_t sea_decode_step_workspace_bytes(int N, int H, int D, int P, int S, int C, int k_clamp, int dtype) {
    if (N <= 0 || H <= 0 || D <= 0 || P <= 0 || S <= 0 || C < S * H || k_clamp <= 0) return 0;
    return decode_ws(N, H, D, P, S, C, k_clamp, dtype == SEA_DTYPE_F32 ? 4 : 2).total;
}

int sea_decode_step(const void* q, int64_t q_sn, int64_t q_sh,
                    const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                    const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype,
                    const float* pos_emb, const float* proj,
                    const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                    const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                    const float* scl_w, const float* scl_b,
                    const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                    const float* conv3_w, const float* conv3_b, const float* out_ln_w, const float* out_ln_b,
                    void* mlp_ws, void* conv1_ws, void* conv2_ws, int repack,
                    const float* k_per_row,
                    const float* perf_in, float* perf_out, const void* xwin_in, void* xwin_out, const void* ywin_in, void* ywin_out,
                    void* context, float* probs, void* workspace, int64_t workspace_bytes,
                    int N, int H, int D, int F, int P, int S, int C, int t, int k_clamp, int use_scaler, void* stream) {
    SEA_CHECK_ARG(q && k && v && pos_emb && proj && enc_w && enc_b && enc_ln_w && enc_ln_b && dec_w && dec_b && cnn_ln_w && cnn_ln_b && scl_w && scl_b &&
                  conv1_w && conv1_b && conv2_w && conv2_b && conv3_w && conv3_b && out_ln_w && out_ln_b && k_per_row && perf_in && perf_out &&
                  xwin_in && xwin_out && ywin_in && ywin_out && context && probs && workspace, "sea_decode_step: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && D > 0 && F > 0 && P > 0 && (P % 4) == 0 && S > 0 && C >= S * H && t >= 0 && k_clamp > 0, "sea_decode_step: bad shape");
    SEA_CHECK_ARG(dtype == SEA_DTYPE_F32 || dtype == SEA_DTYPE_BF16 || dtype == SEA_DTYPE_F16, "sea_decode_step: bad dtype");
    const int esz = dtype == SEA_DTYPE_F32 ? 4 : 2;
    const int W = P / 4, T_SRC = t + 1, SH = S * H;
    const DecodeWs ws = decode_ws(N, H, D, P, S, C, k_clamp, esz);
    SEA_CHECK_ARG(workspace_bytes >= ws.total && (((uintptr_t) workspace) & 255) == 0, "sea_decode_step: workspace too small or misaligned (%lld B needed)", (long long) ws.total);
    cudaStream_t s = (cudaStream_t) stream;
    uint8_t* wb = reinterpret_cast<uint8_t*>(workspace);
    void *ctx = wb + ws.ctx, *cumavg = wb + ws.cumavg, *cnn_row = wb + ws.cnn_row, *xwin5 = wb + ws.xwin5, *y1full = wb + ws.y1full;
    void *ywin5 = wb + ws.ywin5, *y2full = wb + ws.y2full, *y2row = wb + ws.y2row;
    float* scales = reinterpret_cast<float*>(wb + ws.scales);
    uint32_t* bits = reinterpret_cast<uint32_t*>(wb + ws.bits);
    float* kpr = reinterpret_cast<float*>(wb + ws.kpr);
    int rc;

    // ---- a2 + a3 + a13 from the running sums (functional state: sums copied, then advanced in place)
    const int64_t perf_floats = sea_performer_state_floats(N, H, D, F);
    SEA_CUDA_TRY(cudaMemcpyAsync(perf_out, perf_in, (size_t) perf_floats * 4, cudaMemcpyDeviceToDevice, s), "state copy");
    const uint8_t* k_new = reinterpret_cast<const uint8_t*>(k) + (int64_t) t * k_st * esz;
    const uint8_t* v_new = reinterpret_cast<const uint8_t*>(v) + (int64_t) t * v_st * esz;
    rc = sea_performer_causal_state_fwd(q, q_sn, q_sh, D, k_new, k_sn, k_sh, k_st, v_new, v_sn, v_sh, v_st, pos_emb, proj, dtype, perf_out, ctx, cumavg,
                                        N, H, 1, t, D, F, stream);
    if (rc) return rc;

    // ---- a4 on the new token -> one CNN input row [N,1,W,C]
    const bool bf16 = dtype == SEA_DTYPE_BF16;
    const bool aligned_v = ((v_sn | v_sh | v_st) % 8) == 0;
    if (bf16 && mlp_ws && aligned_v && sea_predictor_mlp_umma_supported(dtype, H, D, S, W) && (int64_t) W * C * 2 <= 32 * 1024 && (C % 8) == 0) {
        rc = sea_predictor_mlp_umma_fwd_ex(ctx, v_new, v_sn, v_sh, v_st, repack ? enc_w : nullptr, enc_b, enc_ln_w, enc_ln_b, repack ? dec_w : nullptr, dec_b,
                                           cnn_ln_w, cnn_ln_b, repack ? scl_w : nullptr, scl_b, cnn_row, scales, mlp_ws, N, H, 1, D, S, W, C, stream);
    } else if (bf16 && mlp_ws && aligned_v && sea_predictor_mlp_mma_supported(dtype, H, D, S, W) && (int64_t) W * C * 2 <= 32 * 1024 && (C % 8) == 0) {
        rc = sea_predictor_mlp_mma_fwd(ctx, v_new, v_sn, v_sh, v_st, repack ? enc_w : nullptr, enc_b, enc_ln_w, enc_ln_b, repack ? dec_w : nullptr, dec_b,
                                       cnn_ln_w, cnn_ln_b, repack ? scl_w : nullptr, scl_b, cnn_row, scales, mlp_ws, N, H, 1, D, S, W, C, stream);
    } else {
        SEA_CHECK_ARG(C == SH, "sea_decode_step: zero-padded channels need the tensor-core MLP");
        rc = sea_predictor_mlp_fwd(ctx, v_new, v_sn, v_sh, v_st, dtype, enc_w, enc_b, enc_ln_w, enc_ln_b, dec_w, dec_b, cnn_ln_w, cnn_ln_b, scl_w, scl_b,
                                   cnn_row, scales, nullptr, N, H, 1, D, S, W, stream);
    }
    if (rc) return rc;

    // ---- windowed CNN: rows t-4 .. t of the CNN input, conv 1, rows t-4 .. t of its output, conv 2 (each conv looks 4 rows back)
    const int64_t row_b = (int64_t) W * C * esz;          // bytes of one window row; C = 2H (or 64) is even, so every size below is a multiple of 4
    SEA_CHECK_ARG((C % 2) == 0 && (SH % 2) == 0, "sea_decode_step: odd channel count");
    uint8_t *xw5 = reinterpret_cast<uint8_t*>(xwin5), *yw5 = reinterpret_cast<uint8_t*>(ywin5);
    {   // x window = [last 4 rows | new row]; the window the next step starts from = its rows 1 .. 4; k_per_row broadcast for the top-k
        CopyJobs jobs = {};
        jobs.j[0] = make_job(xw5, xwin_in, 5 * row_b, 0, 4 * row_b, 0, N, 1, 4 * row_b);
        jobs.j[1] = make_job(xw5 + 4 * row_b, cnn_row, 5 * row_b, 0, row_b, 0, N, 1, row_b);
        jobs.count = 2;
        jobs.bcast_dst = kpr; jobs.bcast_src = k_per_row; jobs.bcast_n = N;
        SEA_CUDA_TRY(launch_copies(jobs, s), "decode_copy_kernel");
    }
    const bool conv_tc = bf16 && conv1_ws && conv2_ws && sea_conv_umma_supported(dtype, W, C, C) && C == 64;
    if (conv_tc) rc = sea_causal_conv3x3_dil2_relu_umma(xwin5, repack ? conv1_w : nullptr, conv1_b, y1full, conv1_ws, N, 5, W, C, C, stream);
    else rc = sea_causal_conv3x3_dil2_relu(xwin5, conv1_w, conv1_b, y1full, dtype, N, 5, W, C, C, stream);
    if (rc) return rc;
    {   // y window = [last 4 rows of conv 1's output | its new row]
        CopyJobs jobs = {};
        jobs.j[0] = make_job(yw5, ywin_in, 5 * row_b, 0, 4 * row_b, 0, N, 1, 4 * row_b);
        jobs.j[1] = make_job(yw5 + 4 * row_b, reinterpret_cast<uint8_t*>(y1full) + 4 * row_b, 5 * row_b, 0, 5 * row_b, 0, N, 1, row_b);
        jobs.count = 2;
        SEA_CUDA_TRY(launch_copies(jobs, s), "decode_copy_kernel");
    }
    if (conv_tc) rc = sea_causal_conv3x3_dil2_relu_umma(ywin5, repack ? conv2_w : nullptr, conv2_b, y2full, conv2_ws, N, 5, W, C, C, stream);
    else rc = sea_causal_conv3x3_dil2_relu(ywin5, conv2_w, conv2_b, y2full, dtype, N, 5, W, C, C, stream);
    if (rc) return rc;
    {   // the windows the next step starts from (rows 1 .. 4), and the last row of conv 2 with its real channels only: [N, W, S*H]
        CopyJobs jobs = {};
        jobs.j[0] = make_job(xwin_out, xw5 + row_b, 4 * row_b, 0, 5 * row_b, 0, N, 1, 4 * row_b);
        jobs.j[1] = make_job(ywin_out, yw5 + row_b, 4 * row_b, 0, 5 * row_b, 0, N, 1, 4 * row_b);
        jobs.j[2] = make_job(y2row, reinterpret_cast<uint8_t*>(y2full) + 4 * row_b, (int64_t) W * SH * esz, (int64_t) SH * esz, 5 * row_b, (int64_t) C * esz, N, W,
                             (int64_t) SH * esz);
        jobs.count = 3;
        SEA_CUDA_TRY(launch_copies(jobs, s), "decode_copy_kernel");
    }

    // ---- a5 tail + a6, a7 of the one query row
    // The prefill's fused tail / top-k kernel (one CTA per row, keys in registers) behind a small 1x1 convolution: 15 us where the
    // stand-alone tail (1x1 conv + resize + LayerNorm + softmax of all heads in one serial CTA) and the stand-alone top-k took 70 + 18 us
    // of the ~150 us step.  Shapes the fused kernel does not take keep the two stand-alone kernels.
    const bool fused_tail = (P % 32) == 0 && (P % W) == 0 && P <= 1024 && H <= 64 && ((H + 7) / 8) * (P / 32) <= 32 &&
                            ((int64_t) SH * (H + 1) + 8 * SH) * 4 <= 48 * 1024;
    if (fused_tail) {
        float* y3row = reinterpret_cast<float*>(wb + ws.y3row);
        const size_t smem = ((size_t) SH * (H + 1) + 8 * (size_t) SH) * 4;
        const dim3 grid((unsigned) ((W + 7) / 8), (unsigned) N);
        SEA_DISPATCH_DTYPE(dtype, T_, decode_conv1x1_kernel<T_><<<grid, 256, smem, s>>>(reinterpret_cast<const T_*>(y2row), conv3_w, conv3_b, y3row, W, SH, H));
        SEA_CHECK_LAUNCH("decode_conv1x1_kernel");
        rc = sea_predictor_tail_topk_fwd(y3row, conv3_b, out_ln_w, out_ln_b, kpr, probs, bits, nullptr, 0, N, H, 1, W, P, stream);
        if (rc) return rc;
    } else {
        rc = sea_predictor_tail_fwd(y2row, dtype, conv3_w, conv3_b, out_ln_w, out_ln_b, probs, nullptr, N, H, 1, W, SH, P, stream);
        if (rc) return rc;
        rc = sea_topk_mask_bits(probs, (int64_t) H * P, (int64_t) P, (int64_t) P, kpr, nullptr, bits, N, H, 1, P, 0, stream);
        if (rc) return rc;
    }

    // ---- a8 - a14: the query row against the whole KV cache
    if (dtype != SEA_DTYPE_F32 && (D == 32 || D == 64 || D == 80 || D == 96 || D == 128) && (P % 32) == 0 && P <= 1024 &&
        ((q_sn | q_sh | k_sn | k_sh | k_st | v_sn | v_sh | v_st) % 8) == 0) {
        return sea_sparse_attention_bits_fwd(bits, q, q_sn, q_sh, D, k, k_sn, k_sh, k_st, v, v_sn, v_sh, v_st, scales, cumavg, (int64_t) D, (int64_t) D,
                                             use_scaler, dtype, context, N, H, 1, T_SRC, D, P, k_clamp, 1, stream);
    }
    int32_t* crow = reinterpret_cast<int32_t*>(wb + ws.crow);
    int32_t* col = reinterpret_cast<int32_t*>(wb + ws.col);
    int32_t* head_ptr = (P % 32) == 0 ? reinterpret_cast<int32_t*>(wb + ws.head_ptr) : nullptr;
    rc = sea_csr_count(bits, crow, 0, N, H, 1, P, T_SRC, k_clamp, 1, stream);
    if (rc) return rc;
    rc = sea_csr_fill(bits, crow, col, 0, ws.z_alloc, head_ptr, N, H, 1, P, T_SRC, k_clamp, 1, stream);
    if (rc) return rc;
    return sea_sparse_attention_fwd(crow, col, 0, ws.z_alloc, q, q_sn, q_sh, D, k, k_sn, k_sh, k_st, v, v_sn, v_sh, v_st, scales, cumavg, (int64_t) D, (int64_t) D,
                                    use_scaler, dtype, context, nullptr, head_ptr, N, H, 1, T_SRC, D, stream);
}

}  // extern "C"
