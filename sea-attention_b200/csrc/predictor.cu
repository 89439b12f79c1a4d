// a4-a6: attention predictor, fp32 SIMT implementation (the exactness path; see umma_conv.cu for the
// bf16 tcgen05 implicit-GEMM path).  Activations between the stages are channels-last [N, T, W, C]
// so that a (t, w) pixel's channels are one contiguous 128..256-byte line.
//   predictor_mlp_kernel   : cat(ctx, v) -> Linear+LN+GELU -> {dec_row Linear + ChannelSplit + LN(W), scaler Linear}
//                            (reference attention.py:190-196, 242-245, 267, 289-291, 577-625)
//   causal_conv_kernel     : CausalConv2d(C,O,3,pad 2,dil 2,causal)+ReLU (modules.py:96-192, attention.py:271-274)
//   predictor_tail_kernel  : nearest x4 -> 1x1 CausalConv2d(pad 1) -> area resize -> LN(P) -> softmax(P)
//                            (attention.py:275-280, 670-673; modules.py:12-31, 42-55, 77-92)
#include "common.cuh"
#include "tile_gemm.cuh"

namespace sea {

constexpr int kMlpThreads = 256;
constexpr int kKC = 32;  // K-chunk of the weight staged per step

// C[tok][o] = bias[o] + sum_k A[tok][k] * Wt[o][k]   (Wt in torch layout [OUT][K], global memory)
// A in shared memory [TOK][lda]; result handed to epi(tok, o, value).  Weight K-chunks are staged
// transposed ([kk][OUT+1]) so the inner loop reads are conflict free.
template <int kMaxTiles, class FE>
__device__ __forceinline__ void block_linear(const float* A, int lda, int TOK, const float* __restrict__ Wt,
                                             const float* __restrict__ bias, int OUT, int K, float* wstage, FE epi) {
    const int ntj = (OUT + 3) >> 2, nti = (TOK + 3) >> 2;
    const int ntiles = nti * ntj;
    float acc[kMaxTiles][4][4];
#pragma unroll
    for (int s = 0; s < kMaxTiles; ++s)
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[s][x][y] = 0.f;
    const int ldw = OUT + 4;   // keeps float4 alignment of wstage rows
    for (int k0 = 0; k0 < K; k0 += kKC) {
        const int kc = min(kKC, K - k0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < OUT * kKC; idx += blockDim.x) {
            int o = idx / kKC, kk = idx % kKC;
            wstage[kk * ldw + o] = kk < kc ? Wt[(int64_t) o * K + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < kMaxTiles; ++s) {
            const int tile = threadIdx.x + s * blockDim.x;
            if (tile < ntiles) {
                const int i0 = (tile / ntj) << 2, j0 = (tile % ntj) << 2;
                for (int kk = 0; kk < kc; ++kk) {
                    float bv[4];
                    if (j0 + 4 <= OUT) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&wstage[kk * ldw + j0]);
                        bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
                    } else {
#pragma unroll
                        for (int y = 0; y < 4; ++y) bv[y] = j0 + y < OUT ? wstage[kk * ldw + j0 + y] : 0.f;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float av = i0 + x < TOK ? A[(i0 + x) * lda + k0 + kk] : 0.f;
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[s][x][y] = fmaf(av, bv[y], acc[s][x][y]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < kMaxTiles; ++s) {
        const int tile = threadIdx.x + s * blockDim.x;
        if (tile < ntiles) {
            const int i0 = (tile / ntj) << 2, j0 = (tile % ntj) << 2;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    if (i0 + x < TOK && j0 + y < OUT) epi(i0 + x, j0 + y, acc[s][x][y] + bias[j0 + y]);
        }
    }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

struct MlpDims {
    int N, H, T, D, S, W;
    int TT;     // query rows per CTA
    int TOK;    // TT * H tokens per CTA
};

template <typename T>
__global__ void __launch_bounds__(kMlpThreads)
predictor_mlp_kernel(const T* __restrict__ ctx, const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                     const float* __restrict__ enc_w, const float* __restrict__ enc_b,
                     const float* __restrict__ enc_ln_w, const float* __restrict__ enc_ln_b,
                     const float* __restrict__ dec_w, const float* __restrict__ dec_b,
                     const float* __restrict__ cnn_ln_w, const float* __restrict__ cnn_ln_b,
                     const float* __restrict__ scl_w, const float* __restrict__ scl_b,
                     T* __restrict__ cnn_in, float* __restrict__ scales, T* __restrict__ t_pred, MlpDims dm) {
    extern __shared__ __align__(16) float smem[];
    const int D = dm.D, H = dm.H, S = dm.S, W = dm.W;
    const int D2 = 2 * D, D3 = 3 * D, SW = S * W, C = H * S;
    const int TOK = dm.TOK;
    const int ldx = D3 + 1, ldh = D2 + 1, ldd = SW + 1;
    float* xs = smem;                        // [TOK][ldx]
    float* hs = xs + TOK * ldx;              // [TOK][ldh]
    float* ds = hs + TOK * ldh;              // [TOK][ldd]
    float* wstage = ds + TOK * ldd;          // [kKC][max(D2,SW)+4]
    wstage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wstage) + 15) & ~uintptr_t(15));
    const int tblocks = (dm.T + dm.TT - 1) / dm.TT;
    const int n = blockIdx.x / tblocks;
    const int t0 = (blockIdx.x % tblocks) * dm.TT;
    const int tid = threadIdx.x;
    // token index inside the CTA: tok = tt*H + h
    for (int idx = tid; idx < TOK * D3; idx += kMlpThreads) {
        const int tok = idx / D3, c = idx % D3;
        const int tt = tok / H, h = tok % H, t = t0 + tt;
        float val = 0.f;
        if (t < dm.T) {
            if (c < D2) val = to_f32(ctx[(((int64_t) n * H + h) * dm.T + t) * D2 + c]);
            else val = to_f32(v[(int64_t) n * v_sn + (int64_t) h * v_sh + (int64_t) t * v_st + (c - D2)]);
        }
        xs[tok * ldx + c] = val;
    }
    // enc Linear(3D -> 2D)
    block_linear<4>(xs, ldx, TOK, enc_w, enc_b, D2, D3, wstage, [&](int tok, int o, float val) { hs[tok * ldh + o] = val; });
    __syncthreads();
    // LayerNorm(2D) + GELU, one warp per token
    {
        const int lane = tid & 31, wid = tid >> 5;
        for (int tok = wid; tok < TOK; tok += kMlpThreads / 32) {
            float s = 0.f;
            for (int c = lane; c < D2; c += 32) s += hs[tok * ldh + c];
            const float mean = warp_sum(s) / (float) D2;
            float q = 0.f;
            for (int c = lane; c < D2; c += 32) { float d = hs[tok * ldh + c] - mean; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(warp_sum(q) / (float) D2 + 1e-5f);
            const int tt = tok / H, h = tok % H, t = t0 + tt;
            for (int c = lane; c < D2; c += 32) {
                float y = gelu_erf((hs[tok * ldh + c] - mean) * rstd * enc_ln_w[c] + enc_ln_b[c]);
                hs[tok * ldh + c] = y;
                if (t_pred != nullptr && t < dm.T) t_pred[(((int64_t) n * H + h) * dm.T + t) * D2 + c] = from_f32<T>(y);
            }
        }
    }
    __syncthreads();
    // dec_row Linear(2D -> S*W) and scaler Linear(2D -> 2)
    block_linear<4>(hs, ldh, TOK, dec_w, dec_b, SW, D2, wstage, [&](int tok, int o, float val) { ds[tok * ldd + o] = val; });
    {
        const int lane = tid & 31, wid = tid >> 5;
        for (int pair = wid; pair < TOK * 2; pair += kMlpThreads / 32) {
            const int tok = pair >> 1, o = pair & 1;
            float s = 0.f;
            for (int c = lane; c < D2; c += 32) s = fmaf(hs[tok * ldh + c], scl_w[o * D2 + c], s);
            s = warp_sum(s);
            const int tt = tok / H, h = tok % H, t = t0 + tt;
            if (lane == 0 && t < dm.T) scales[((((int64_t) n * H + h) * dm.T + t) << 1) + o] = s + scl_b[o];
        }
    }
    __syncthreads();
    // first CNN LayerNorm over W for every (tok, s) row; in place (the BERT predictor has none: null weights skip it)
    if (cnn_ln_w != nullptr) {
        const int lane = tid & 31, wid = tid >> 5;
        for (int row = wid; row < TOK * S; row += kMlpThreads / 32) {
            float* r = ds + (row / S) * ldd + (row % S) * W;
            float s = 0.f;
            for (int w = lane; w < W; w += 32) s += r[w];
            const float mean = warp_sum(s) / (float) W;
            float q = 0.f;
            for (int w = lane; w < W; w += 32) { float d = r[w] - mean; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(warp_sum(q) / (float) W + 1e-5f);
            for (int w = lane; w < W; w += 32) r[w] = (r[w] - mean) * rstd * cnn_ln_w[w] + cnn_ln_b[w];
        }
    }
    __syncthreads();
    // channels-last store: cnn_in[n, t, w, c = h*S + s]
    for (int idx = tid; idx < dm.TT * W * C; idx += kMlpThreads) {
        const int c = idx % C, w = (idx / C) % W, tt = idx / (C * W);
        const int t = t0 + tt;
        if (t < dm.T) {
            const int h = c / S, s = c % S;
            cnn_in[(((int64_t) n * dm.T + t) * W + w) * C + c] = from_f32<T>(ds[(tt * H + h) * ldd + s * W + w]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CausalConv2d 3x3, dilation 2, + ReLU on channels-last data.  CTA = kTB consecutive rows t of one
// 64-wide w tile, all output channels; all nine weight slabs stay resident in shared memory.
// thread tile: 4 positions x 4 output channels.
// ------------------------------------------------------------------------------------------------
constexpr int kConvPT = 64;   // positions (w) per CTA tile
constexpr int kConvTB = 8;    // rows (t) per CTA

template <typename T>
__global__ void __launch_bounds__(1024)
causal_conv_kernel(const T* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
                   T* __restrict__ y, int N, int Tn, int W, int C, int O) {
    extern __shared__ __align__(16) float smem[];
    const int Op = (O + 3) & ~3;
    const int ldp = kConvPT + 4 + 4;          // positions -2 .. PT+1 (+ pad keeps rows 16B aligned)
    float* ws = smem;                          // [9][C][Op]
    float* xs = ws + 9 * C * Op;               // [3][C][ldp]
    const int wtiles = (W + kConvPT - 1) / kConvPT;
    const int tblocks = (Tn + kConvTB - 1) / kConvTB;
    int b = blockIdx.x;
    const int wt = b % wtiles; b /= wtiles;
    const int tb = b % tblocks;
    const int n = b / tblocks;
    const int w0 = wt * kConvPT;
    const int tid = threadIdx.x;
    // weight [O][C][5][3] -> ws[tap = i*3+j][c][o]
    for (int idx = tid; idx < 9 * C * Op; idx += blockDim.x) {
        const int o = idx % Op, c = (idx / Op) % C, tap = idx / (Op * C);
        const int i = tap / 3, j = tap % 3;
        ws[idx] = o < O ? weight[(((int64_t) o * C + c) * 5 + i) * 3 + j] : 0.f;
    }
    const int ntj = Op >> 2;                    // output-channel groups
    const int pg = tid / ntj, og = tid % ntj;   // position group (4 positions), channel group
    const bool active = pg < kConvPT / 4;
    for (int t = tb * kConvTB; t < min(Tn, (tb + 1) * kConvTB); ++t) {
        __syncthreads();
        // rows t-4, t-2, t ; positions w0-2 .. w0+PT+1
        for (int idx = tid; idx < 3 * (kConvPT + 4) * C; idx += blockDim.x) {
            const int c = idx % C, p = (idx / C) % (kConvPT + 4), i = idx / (C * (kConvPT + 4));
            const int tr = t - 4 + 2 * i, wc = w0 - 2 + p;
            float val = 0.f;
            if (tr >= 0 && wc >= 0 && wc < W) val = to_f32(x[(((int64_t) n * Tn + tr) * W + wc) * C + c]);
            xs[(i * C + c) * ldp + p] = val;
        }
        __syncthreads();
        if (active) {
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[a][q] = 0.f;
            const int p0 = pg * 4;
            for (int tap = 0; tap < 9; ++tap) {
                const int i = tap / 3, j = tap % 3;
                const float* xrow = xs + (i * C) * ldp + p0 + 2 * j;
                const float* wrow = ws + (tap * C) * Op + og * 4;
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    const float2 xa = *reinterpret_cast<const float2*>(xrow + c * ldp);
                    const float2 xb = *reinterpret_cast<const float2*>(xrow + c * ldp + 2);
                    const float4 wv = *reinterpret_cast<const float4*>(wrow + c * Op);
                    const float xv[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        acc[a][0] = fmaf(xv[a], wv.x, acc[a][0]); acc[a][1] = fmaf(xv[a], wv.y, acc[a][1]);
                        acc[a][2] = fmaf(xv[a], wv.z, acc[a][2]); acc[a][3] = fmaf(xv[a], wv.w, acc[a][3]);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int wc = w0 + p0 + a;
                if (wc < W) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int o = og * 4 + q;
                        if (o < O) y[(((int64_t) n * Tn + t) * W + wc) * O + o] = from_f32<T>(fmaxf(acc[a][q] + bias[o], 0.f));
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tail: CTA per (n, t).  y[h][w] = b[h] + sum_c Wt[h][c] x[w][c]; u[h][0] = u[h][P+1] = b[h] (the zero
// padded columns of the 1x1 conv), u[h][1+4w..4+4w] = y[h][w] (nearest x4 BEFORE the conv commutes with a
// 1x1 conv); area resize P+2 -> P; LayerNorm(P); softmax(P).
// ------------------------------------------------------------------------------------------------
constexpr int kTailThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kTailThreads)
predictor_tail_kernel(const T* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
                      const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                      float* __restrict__ probs, float* __restrict__ scores, int N, int H, int Tn, int W, int C, int P) {
    extern __shared__ __align__(16) float smem[];
    const int ldw_ = W + 4;
    float* xT = smem;                 // [C][W+4]
    float* wt = xT + C * ldw_;        // [H][C+1]
    float* ys = wt + H * (C + 1);     // [H][W+1]
    const int n = blockIdx.x / Tn, t = blockIdx.x % Tn;
    const int tid = threadIdx.x;
    const T* xr = x + ((int64_t) n * Tn + t) * W * C;
    for (int idx = tid; idx < W * C; idx += kTailThreads) {
        const int c = idx % C, w = idx / C;
        xT[c * ldw_ + w] = to_f32(xr[idx]);
    }
    for (int idx = tid; idx < H * C; idx += kTailThreads) wt[(idx / C) * (C + 1) + idx % C] = weight[idx];
    __syncthreads();
    tile_gemm(H, W, C, [&](int h, int c) { return wt[h * (C + 1) + c]; }, [&](int c, int w) { return xT[c * ldw_ + w]; },
              [&](int h, int w, float acc) { ys[h * (W + 1) + w] = acc + bias[h]; });
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5;
    const int up = P / W;             // nearest upsample factor (4)
    const int PW = P + 2;
    constexpr int kMaxPerLane = 32;   // P <= 1024
    for (int h = wid; h < H; h += kTailThreads / 32) {
        float val[kMaxPerLane];
        const float bh = bias[h];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int j = lane + 32 * i;
            float a = 0.f;
            if (j < P) {
                const int st = (int) (((int64_t) j * PW) / P);
                const int en = (int) (((int64_t) (j + 1) * PW + P - 1) / P);
                for (int pcol = st; pcol < en; ++pcol)
                    a += (pcol == 0 || pcol == PW - 1) ? bh : ys[h * (W + 1) + (pcol - 1) / up];
                a /= (float) (en - st);
                s += a;
            }
            val[i] = a;
        }
        const float mean = warp_sum(s) / (float) P;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int j = lane + 32 * i;
            if (j < P) { float d = val[i] - mean; q = fmaf(d, d, q); }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float) P + 1e-5f);
        float mx = -INFINITY;
        float* srow = scores ? scores + (((int64_t) n * H + h) * Tn + t) * P : nullptr;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int j = lane + 32 * i;
            if (j < P) {
                val[i] = (val[i] - mean) * rstd * ln_w[j] + ln_b[j];
                if (srow) srow[j] = val[i];
                mx = fmaxf(mx, val[i]);
            }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int j = lane + 32 * i;
            if (j < P) { val[i] = expf(val[i] - mx); sum += val[i]; }
        }
        const float inv = 1.0f / warp_sum(sum);
        float* prow = probs + (((int64_t) n * H + h) * Tn + t) * P;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            const int j = lane + 32 * i;
            if (j < P) prow[j] = val[i] * inv;
        }
    }
}

}  // namespace sea

using namespace sea;

extern "C" {

int sea_predictor_mlp_fwd(const void* ctx, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype,
                          const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                          const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                          const float* scl_w, const float* scl_b, void* cnn_in, float* scales, void* t_pred,
                          int N, int H, int T, int D, int S, int W, void* stream) {
    SEA_CHECK_ARG(ctx && v && enc_w && enc_b && enc_ln_w && enc_ln_b && dec_w && dec_b && scl_w && scl_b && cnn_in && scales && (!cnn_ln_w == !cnn_ln_b),
                  "sea_predictor_mlp_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && D > 0 && S > 0 && W > 0, "sea_predictor_mlp_fwd: bad shape");
    MlpDims dm;
    dm.N = N; dm.H = H; dm.T = T; dm.D = D; dm.S = S; dm.W = W;
    dm.TT = H >= 64 ? 1 : (64 / H);
    if (dm.TT > T) dm.TT = T;
    dm.TOK = dm.TT * H;
    const int D2 = 2 * D, D3 = 3 * D, SW = S * W;
    const int maxout = D2 > SW ? D2 : SW;
    SEA_CHECK_ARG(((dm.TOK + 3) / 4) * ((maxout + 3) / 4) <= kMlpThreads * 4,
                  "sea_predictor_mlp_fwd: tile budget exceeded (H=%d, 2D=%d, S*W=%d)", H, D2, SW);
    const size_t smem = ((size_t) dm.TOK * (D3 + 1) + (size_t) dm.TOK * (D2 + 1) + (size_t) dm.TOK * (SW + 1) +
                         (size_t) kKC * (maxout + 4) + 8) * sizeof(float);
    SEA_CHECK_ARG(smem <= 227 * 1024, "sea_predictor_mlp_fwd: needs %zu B of shared memory", smem);
    const int tblocks = (T + dm.TT - 1) / dm.TT;
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto kern = predictor_mlp_kernel<T_>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        kern<<<(unsigned) ((int64_t) N * tblocks), kMlpThreads, smem, (cudaStream_t) stream>>>(
            (const T_*) ctx, (const T_*) v, v_sn, v_sh, v_st, enc_w, enc_b, enc_ln_w, enc_ln_b, dec_w, dec_b, cnn_ln_w, cnn_ln_b,
            scl_w, scl_b, (T_*) cnn_in, scales, (T_*) t_pred, dm);
        SEA_CHECK_LAUNCH("predictor_mlp_kernel");
    });
    return SEA_OK;
}

int sea_causal_conv3x3_dil2_relu(const void* x, const float* weight, const float* bias, void* y, int dtype,
                                 int N, int T, int W, int C, int O, void* stream) {
    SEA_CHECK_ARG(x && weight && bias && y, "sea_causal_conv3x3_dil2_relu: null pointer");
    SEA_CHECK_ARG(N > 0 && T > 0 && W > 0 && C > 0 && O > 0, "sea_causal_conv3x3_dil2_relu: bad shape");
    const int Op = (O + 3) & ~3;
    const int threads = (kConvPT / 4) * (Op / 4);
    SEA_CHECK_ARG(threads <= 1024, "sea_causal_conv3x3_dil2_relu: %d output channels unsupported (<= 256)", O);
    const size_t smem = ((size_t) 9 * C * Op + (size_t) 3 * C * (kConvPT + 8)) * sizeof(float);
    SEA_CHECK_ARG(smem <= 227 * 1024, "sea_causal_conv3x3_dil2_relu: C=%d O=%d needs %zu B of shared memory", C, O, smem);
    const int wtiles = (W + kConvPT - 1) / kConvPT, tblocks = (T + kConvTB - 1) / kConvTB;
    const int nthreads = ((threads + 31) / 32) * 32;
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto kern = causal_conv_kernel<T_>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        kern<<<(unsigned) ((int64_t) N * tblocks * wtiles), nthreads, smem, (cudaStream_t) stream>>>(
            (const T_*) x, weight, bias, (T_*) y, N, T, W, C, O);
        SEA_CHECK_LAUNCH("causal_conv_kernel");
    });
    return SEA_OK;
}

int sea_predictor_tail_fwd(const void* x, int dtype, const float* weight, const float* bias, const float* ln_w,
                           const float* ln_b, float* probs, float* scores, int N, int H, int T, int W, int C, int P,
                           void* stream) {
    SEA_CHECK_ARG(x && weight && bias && ln_w && ln_b && probs, "sea_predictor_tail_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && W > 0 && C > 0 && P > 0, "sea_predictor_tail_fwd: bad shape");
    SEA_CHECK_ARG(P % W == 0 && P <= 1024, "sea_predictor_tail_fwd: P=%d must be a multiple of W=%d and <= 1024", P, W);
    const size_t smem = ((size_t) C * (W + 4) + (size_t) H * (C + 1) + (size_t) H * (W + 1)) * sizeof(float);
    SEA_CHECK_ARG(smem <= 227 * 1024, "sea_predictor_tail_fwd: needs %zu B of shared memory", smem);
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto kern = predictor_tail_kernel<T_>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        kern<<<(unsigned) ((int64_t) N * T), kTailThreads, smem, (cudaStream_t) stream>>>(
            (const T_*) x, weight, bias, ln_w, ln_b, probs, scores, N, H, T, W, C, P);
        SEA_CHECK_LAUNCH("predictor_tail_kernel");
    });
    return SEA_OK;
}

}  // extern "C"
