// Non-causal (BERT) variant of the SEA attention path, fp32 SIMT kernels for every dtype.
//   a2'  v_for_atten = cat(bilinear grid-sample of the d x d identity, v)          attention.py:462-502
//   a3'  FAVOR+ softmax-feature Performer, un-prefixed sums                          performer-pytorch softmax_kernel / linear_attention
//   a5'  BERT predictor CNN: Conv2d(s(2,1))+ReLU, Conv2d+ReLU, nearest (2,1), Conv2d, bilinear to (T,P), softmax
//                                                                                    attention.py:207-218, 670-673
//   a7'  'batch' grouped top-k: one group per batch item over H*T*P keys              attention.py:833-837, 871-917
//   a13' probability-weighted mean of v                                              attention.py:1209-1219
// The non-causal CSR interpolation and the sparse attention reuse the causal kernels (is_causal = 0).
#include "common.cuh"
#include "tile_gemm.cuh"
#include "topk.cuh"

namespace sea {

constexpr int kNcThreads = 256;
constexpr int kNcRows = 32;       // token rows per CTA tile

// hat-function value of the grid-sampled identity (closed form of F.grid_sample(eye, bilinear, align_corners=True))
__device__ __forceinline__ float v_identity(int t, int T, int c, int D) {      // T = valid tokens of the item; padded tokens (t >= T) give 0
    if (t >= T) return 0.f;
    const float y_norm = ((float) t / (((float) T - 1.0f) + 1e-8f)) * 2.0f - 1.0f;      // (cumsum-1)/((sum-1)+1e-8)*2-1
    const float ypix = (y_norm + 1.0f) * 0.5f * (float) (D - 1);
    if (ypix < 0.f || ypix > (float) (D - 1)) return 0.f;
    return fmaxf(0.f, 1.0f - fabsf(ypix - (float) c));
}

struct NcPerfDims {
    int N, H, T, D, F, Fp, E, nchunks;
    int64_t slot;      // floats per chunk partial: Fp*E + Fp
    const int32_t* lengths;   // [N] valid tokens per item of a right-padded batch (attention.py:482, 512-514), or nullptr
};

// pass 1: per-(n,h) maximum of d^-1/4 k . P^T over (T, F) -- the stabiliser of the key features
template <typename T>
__global__ void __launch_bounds__(kNcThreads)
nc_kmax_kernel(const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st, const float* __restrict__ proj,
               int* __restrict__ kmax_ord, NcPerfDims dm) {
    extern __shared__ __align__(16) float smem[];
    const int D = dm.D, F = dm.F, Fp = dm.Fp;
    float* projT = smem;                  // [D][Fp]
    float* xk = projT + D * Fp;           // [kNcRows][D]
    __shared__ float red[kNcThreads / 32];
    const int chunk = blockIdx.x, nh = blockIdx.y, n = nh / dm.H, h = nh % dm.H;
    const int r0 = chunk * kNcRows, nv = min(kNcRows, dm.T - r0);
    const float norm = rsqrtf(sqrtf((float) D));
    for (int idx = threadIdx.x; idx < D * Fp; idx += kNcThreads) { int c = idx / Fp, f = idx % Fp; projT[idx] = f < F ? proj[f * D + c] : 0.f; }
    const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
    for (int idx = threadIdx.x; idx < kNcRows * D; idx += kNcThreads) {
        int r = idx / D, c = idx % D;
        xk[idx] = r < nv ? norm * to_f32(kb[(int64_t) (r0 + r) * k_st + c]) : 0.f;
    }
    __syncthreads();
    float mx = -INFINITY;
    tile_gemm(kNcRows, Fp, D, [&](int i, int c) { return xk[i * D + c]; }, [&](int c, int f) { return projT[c * Fp + f]; },
              [&](int i, int f, float acc) { if (i < nv && f < F) mx = fmaxf(mx, acc); });
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kNcThreads / 32; ++w) mx = fmaxf(mx, red[w]);
        const int i = __float_as_int(mx);
        atomicMax(&kmax_ord[nh], i >= 0 ? i : i ^ 0x7fffffff);
    }
}

// pass 2: per-chunk partial sums  ctx_c[f][e] = sum_t k'[t][f] v2[t][e],  ksum_c[f] = sum_t k'[t][f]
template <typename T>
__global__ void __launch_bounds__(kNcThreads)
nc_ksum_kernel(const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
               const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st, const float* __restrict__ proj,
               const int* __restrict__ kmax_ord, float* __restrict__ partial, NcPerfDims dm) {
    extern __shared__ __align__(16) float smem[];
    const int D = dm.D, F = dm.F, Fp = dm.Fp, E = dm.E;
    float* projT = smem;                  // [D][Fp]
    float* xk = projT + D * Fp;           // [kNcRows][D]
    float* kdiag = xk + kNcRows * D;      // [kNcRows]
    float* kpT = kdiag + kNcRows;         // [Fp][kNcRows]
    float* v2 = kpT + Fp * kNcRows;       // [kNcRows][E]
    const int chunk = blockIdx.x, nh = blockIdx.y, n = nh / dm.H, h = nh % dm.H;
    const int r0 = chunk * kNcRows, nv = min(kNcRows, dm.T - r0);
    const float norm = rsqrtf(sqrtf((float) D)), ratio = rsqrtf((float) F);
    const int ko = kmax_ord[nh];
    const float kmax = __int_as_float(ko >= 0 ? ko : ko ^ 0x7fffffff);
    for (int idx = threadIdx.x; idx < D * Fp; idx += kNcThreads) { int c = idx / Fp, f = idx % Fp; projT[idx] = f < F ? proj[f * D + c] : 0.f; }
    const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
    const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;
    for (int idx = threadIdx.x; idx < kNcRows * D; idx += kNcThreads) {
        int r = idx / D, c = idx % D;
        xk[idx] = r < nv ? to_f32(kb[(int64_t) (r0 + r) * k_st + c]) : 0.f;
    }
    for (int idx = threadIdx.x; idx < kNcRows * E; idx += kNcThreads) {
        int r = idx / E, c = idx % E;
        float val = 0.f;
        const int len = dm.lengths ? min(dm.lengths[n], dm.T) : dm.T;      // v_for_atten (identity part AND v part) is zero on padded tokens
        if (r < nv && r0 + r < len) val = c < D ? v_identity(r0 + r, len, c, D) : to_f32(vb[(int64_t) (r0 + r) * v_st + (c - D)]);
        v2[idx] = val;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < kNcRows; r += kNcThreads) {
        float s = 0.f;
        for (int c = 0; c < D; ++c) s = fmaf(xk[r * D + c], xk[r * D + c], s);
        kdiag[r] = s * 0.5f * norm * norm;
    }
    __syncthreads();
    tile_gemm(kNcRows, Fp, D, [&](int i, int c) { return norm * xk[i * D + c]; }, [&](int c, int f) { return projT[c * Fp + f]; },
              [&](int i, int f, float acc) { kpT[f * kNcRows + i] = (i < nv && f < F) ? ratio * (expf(acc - kdiag[i] - kmax) + 1e-4f) : 0.f; });
    __syncthreads();
    float* slot = partial + ((int64_t) nh * dm.nchunks + chunk) * dm.slot;
    tile_gemm(Fp, E, kNcRows, [&](int f, int r) { return kpT[f * kNcRows + r]; }, [&](int r, int e) { return v2[r * E + e]; },
              [&](int f, int e, float acc) { slot[f * E + e] = acc; });
    for (int f = threadIdx.x; f < Fp; f += kNcThreads) {
        float s = 0.f;
        for (int r = 0; r < kNcRows; ++r) s += kpT[f * kNcRows + r];
        slot[Fp * E + f] = s;
    }
}

// deterministic in-order reduction of the chunk partials into slot 0 of every (n, h)
__global__ void __launch_bounds__(256)
nc_reduce_kernel(float* __restrict__ partial, int nchunks, int64_t slot) {
    float* base = partial + (int64_t) blockIdx.y * nchunks * slot;
    for (int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < slot; idx += (int64_t) gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < nchunks; ++c) s += base[(int64_t) c * slot + idx];
        base[idx] = s;
    }
}

// pass 3: out[t] = q'[t] ctx / (q'[t] . ksum)
template <typename T>
__global__ void __launch_bounds__(kNcThreads)
nc_out_kernel(const T* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st, const float* __restrict__ proj,
              const float* __restrict__ partial, T* __restrict__ ctx_out, NcPerfDims dm) {
    extern __shared__ __align__(16) float smem[];
    const int D = dm.D, F = dm.F, Fp = dm.Fp, E = dm.E;
    // region A holds the projection while q' is formed and is then overwritten by the summed state (ctx | ksum): the two
    // are never live together, which keeps F = 266 (BERT-base, nbf = 1) inside the 227 KB of one CTA
    const int regionA = max(D * Fp, Fp * E + Fp);
    float* projT = smem;                  // [D][Fp]
    float* S = smem;                      // [Fp][E] + ksum [Fp]   (after the projection is dead)
    float* xq = smem + regionA;           // [kNcRows][D]
    float* qdiag = xq + kNcRows * D;      // [kNcRows]
    float* qp = qdiag + kNcRows;          // [kNcRows][Fp]
    float* den = qp + kNcRows * Fp;       // [kNcRows]
    const int chunk = blockIdx.x, nh = blockIdx.y, n = nh / dm.H, h = nh % dm.H;
    const int r0 = chunk * kNcRows, nv = min(kNcRows, dm.T - r0);
    const float norm = rsqrtf(sqrtf((float) D)), ratio = rsqrtf((float) F);
    for (int idx = threadIdx.x; idx < D * Fp; idx += kNcThreads) { int c = idx / Fp, f = idx % Fp; projT[idx] = f < F ? proj[f * D + c] : 0.f; }
    const T* qb = q + (int64_t) n * q_sn + (int64_t) h * q_sh;
    for (int idx = threadIdx.x; idx < kNcRows * D; idx += kNcThreads) {
        int r = idx / D, c = idx % D;
        xq[idx] = r < nv ? to_f32(qb[(int64_t) (r0 + r) * q_st + c]) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < kNcRows; r += kNcThreads) {
        float s = 0.f;
        for (int c = 0; c < D; ++c) s = fmaf(xq[r * D + c], xq[r * D + c], s);
        qdiag[r] = s * 0.5f * norm * norm;
    }
    tile_gemm(kNcRows, Fp, D, [&](int i, int c) { return norm * xq[i * D + c]; }, [&](int c, int f) { return projT[c * Fp + f]; },
              [&](int i, int f, float acc) { qp[i * Fp + f] = f < F ? acc : -INFINITY; });
    __syncthreads();
    const float* tot = partial + (int64_t) nh * dm.nchunks * dm.slot;
    for (int idx = threadIdx.x; idx < Fp * E + Fp; idx += kNcThreads) S[idx] = tot[idx];
    __syncthreads();
    {   // row max over F, then q' and the denominator; one warp per row
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const float* ksum = S + Fp * E;
        for (int i = wid; i < kNcRows; i += kNcThreads / 32) {
            float mx = -INFINITY;
            for (int f = lane; f < F; f += 32) mx = fmaxf(mx, qp[i * Fp + f]);
            mx = warp_max(mx);
            float dsum = 0.f;
            for (int f = lane; f < Fp; f += 32) {
                const float val = f < F ? ratio * (expf(qp[i * Fp + f] - qdiag[i] - mx) + 1e-4f) : 0.f;
                qp[i * Fp + f] = val;
                dsum = fmaf(val, ksum[f], dsum);
            }
            dsum = warp_sum(dsum);
            if (lane == 0) den[i] = dsum;
        }
    }
    __syncthreads();
    T* ob = ctx_out + (((int64_t) n * dm.H + h) * dm.T) * E;
    tile_gemm(kNcRows, E, Fp, [&](int i, int f) { return qp[i * Fp + f]; }, [&](int f, int e) { return S[f * E + e]; },
              [&](int i, int e, float acc) { if (i < nv) ob[(int64_t) (r0 + i) * E + e] = from_f32<T>(acc / den[i]); });
}

// ------------------------------------------------------------------------------------------------
// generic 3x3 conv, pad 1, channels-last [N, Tin, W, C] -> [N, Tout, W, O]; stride_t on the token axis; `up` = nearest
// upsample factor applied to the INPUT rows (input row index = virtual row / up).  Thread = (pixel, 4 output channels).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_cl_kernel(const T* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias, T* __restrict__ y,
                  int N, int Tin, int Tout, int W, int C, int O, int stride_t, int up, int relu) {
    const int og = (O + 3) >> 2;
    const int64_t total = (int64_t) N * Tout * W * og;
    const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int o0 = (int) (idx % og) * 4;
    const int w = (int) ((idx / og) % W);
    const int t = (int) ((idx / ((int64_t) og * W)) % Tout);
    const int n = (int) (idx / ((int64_t) og * W * Tout));
    const int Tvirt = Tin * up;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < 3; ++i) {
        const int tv = t * stride_t - 1 + i;
        if (tv < 0 || tv >= Tvirt) continue;
        const int ti = tv / up;
        for (int j = 0; j < 3; ++j) {
            const int wc = w - 1 + j;
            if (wc < 0 || wc >= W) continue;
            const T* xp = x + (((int64_t) n * Tin + ti) * W + wc) * C;
            for (int c = 0; c < C; ++c) {
                const float xv = to_f32(xp[c]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (o0 + u < O) acc[u] = fmaf(xv, __ldg(weight + (((int64_t) (o0 + u) * C + c) * 3 + i) * 3 + j), acc[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (o0 + u < O) {
            float r = acc[u] + bias[o0 + u];
            if (relu) r = fmaxf(r, 0.f);
            y[(((int64_t) n * Tout + t) * W + w) * O + o0 + u] = from_f32<T>(r);
        }
}

// The same convolution with the weights staged once per CTA in shared memory as [tap][c][o] (zero-padded to a multiple of kOT
// output channels): a thread owns one output pixel and kOT output channels, so each input value it loads feeds kOT FMAs whose
// weights arrive as broadcast 16-byte shared-memory loads -- the kernel above issues one global weight load per FMA.
template <typename T, int kOT>
__global__ void __launch_bounds__(256)
conv3x3_cl_smem_kernel(const T* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias, T* __restrict__ y,
                       int N, int Tin, int Tout, int W, int C, int O, int stride_t, int up, int relu) {
    extern __shared__ __align__(16) float cw_sm[];            // [9][C][Opad]
    const int OG = (O + kOT - 1) / kOT, Opad = OG * kOT;
    for (int i = threadIdx.x; i < 9 * C * Opad; i += 256) {
        const int o = i % Opad, c = (i / Opad) % C, tap = i / (Opad * C);
        cw_sm[i] = o < O ? __ldg(weight + ((int64_t) o * C + c) * 9 + tap) : 0.f;
    }
    __syncthreads();
    const int ppc = 256 / OG;                                 // pixels per CTA
    const int og = threadIdx.x % OG, pl = threadIdx.x / OG;
    const int64_t pix = (int64_t) blockIdx.x * ppc + pl;
    if (pl >= ppc || pix >= (int64_t) N * Tout * W) return;
    const int w = (int) (pix % W);
    const int t = (int) ((pix / W) % Tout);
    const int n = (int) (pix / ((int64_t) W * Tout));
    const int Tvirt = Tin * up;
    float acc[kOT];
#pragma unroll
    for (int u = 0; u < kOT; ++u) acc[u] = 0.f;
    for (int i = 0; i < 3; ++i) {
        const int tv = t * stride_t - 1 + i;
        if (tv < 0 || tv >= Tvirt) continue;
        const int ti = tv / up;
        for (int j = 0; j < 3; ++j) {
            const int wc = w - 1 + j;
            if (wc < 0 || wc >= W) continue;
            const T* xp = x + (((int64_t) n * Tin + ti) * W + wc) * C;
            const float* wp = cw_sm + (size_t) ((i * 3 + j) * C) * Opad + og * kOT;
            for (int c = 0; c < C; ++c) {
                const float xv = to_f32(xp[c]);
#pragma unroll
                for (int u4 = 0; u4 < kOT / 4; ++u4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wp + (size_t) c * Opad + 4 * u4);
                    acc[4 * u4] = fmaf(xv, w4.x, acc[4 * u4]); acc[4 * u4 + 1] = fmaf(xv, w4.y, acc[4 * u4 + 1]);
                    acc[4 * u4 + 2] = fmaf(xv, w4.z, acc[4 * u4 + 2]); acc[4 * u4 + 3] = fmaf(xv, w4.w, acc[4 * u4 + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kOT; ++u) {
        const int o = og * kOT + u;
        if (o < O) {
            float r = acc[u] + bias[o];
            if (relu) r = fmaxf(r, 0.f);
            y[pix * O + o] = from_f32<T>(r);
        }
    }
}

// bilinear resize (align_corners=False) of [N, Tin, Win, H] (channels-last) to (T, P), then softmax over P.
// warp per (n, h, t).
template <typename T>
__global__ void __launch_bounds__(256)
bert_tail_kernel(const T* __restrict__ y, float* __restrict__ probs, float* __restrict__ scores, int N, int H, int Tin, int Win, int Tn, int P) {
    const int lane = threadIdx.x & 31;
    const int64_t task = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (task >= (int64_t) N * H * Tn) return;
    const int t = (int) (task % Tn), h = (int) ((task / Tn) % H), n = (int) (task / ((int64_t) Tn * H));
    const float st = (float) Tin / (float) Tn, sw = (float) Win / (float) P;
    const float ft = fmaxf(((float) t + 0.5f) * st - 0.5f, 0.f);
    const int t0 = min((int) ft, Tin - 1), t1 = min(t0 + 1, Tin - 1);
    const float lt = ft - (float) t0;
    constexpr int kMax = 32;      // P <= 1024
    float val[kMax];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMax; ++i) {
        const int p = lane + 32 * i;
        float r = -INFINITY;
        if (p < P) {
            const float fw = fmaxf(((float) p + 0.5f) * sw - 0.5f, 0.f);
            const int w0 = min((int) fw, Win - 1), w1 = min(w0 + 1, Win - 1);
            const float lw = fw - (float) w0;
            auto at = [&](int tt, int ww) { return to_f32(y[(((int64_t) n * Tin + tt) * Win + ww) * H + h]); };
            const float top = at(t0, w0) * (1.f - lw) + at(t0, w1) * lw;
            const float bot = at(t1, w0) * (1.f - lw) + at(t1, w1) * lw;
            r = top * (1.f - lt) + bot * lt;
            if (scores) scores[(((int64_t) n * H + h) * Tn + t) * P + p] = r;
        }
        val[i] = r;
        mx = fmaxf(mx, r);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kMax; ++i) { const int p = lane + 32 * i; if (p < P) { val[i] = expf(val[i] - mx); sum += val[i]; } }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int i = 0; i < kMax; ++i) { const int p = lane + 32 * i; if (p < P) probs[(((int64_t) n * H + h) * Tn + t) * P + p] = val[i] * inv; }
}

// ------------------------------------------------------------------------------------------------
// 'batch' top-k: ONE group per batch item over the H*T*P keys in view(N, H*T*P) order (flat = (h*T + t)*P + m).
// A single 1024-thread CTA per item runs the radix select over global memory (the keys are L2 resident); ties at the
// threshold go to the lower FLAT index, resolved with a chunked block scan in flat order.
// ------------------------------------------------------------------------------------------------
constexpr int kBatchThreads = 1024;

__device__ __forceinline__ int block_excl_scan_1024(int v, int* warp_sums, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int incl = warp_scan_incl_i(v, lane);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int wprefix = 0, tot = 0;
    for (int w = 0; w < kBatchThreads / 32; ++w) { const int s = warp_sums[w]; if (w < wid) wprefix += s; tot += s; }
    __syncthreads();
    total = tot;
    return wprefix + incl - v;
}

__global__ void __launch_bounds__(kBatchThreads)
topk_batch_kernel(const float* __restrict__ keys, const float* __restrict__ k_per_item, uint32_t* __restrict__ mask_bits,
                  int H, int T, int P) {
    __shared__ int hist[256];
    __shared__ int wsum[32];
    __shared__ int piv[2];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int64_t G = (int64_t) H * T * P;
    const float* kb = keys + (int64_t) n * G;
    const int wpr = (H * P + 31) >> 5;
    uint32_t* bits = mask_bits + (int64_t) n * T * wpr;
    for (int64_t i = tid; i < (int64_t) T * wpr; i += kBatchThreads) bits[i] = 0u;
    const float kf = k_per_item[n];
    int64_t K = (int64_t) fminf(ceilf(kf), (float) G);
    uint32_t prefix = 0, mask = 0;
    int64_t remaining = K;
    const bool all_alive = K >= G;
    if (!all_alive) {
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            for (int64_t i = tid; i < G; i += kBatchThreads) {
                const uint32_t u = orderable(kb[i]);
                if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255], 1);
            }
            __syncthreads();
            if (tid == 0) {     // 256 bins: serial suffix walk is negligible next to the pass over G keys
                int64_t above = 0;
                for (int dgt = 255; dgt >= 0; --dgt) {
                    const int c = hist[dgt];
                    if (remaining <= above + c) { piv[0] = dgt; piv[1] = (int) (remaining - above); break; }
                    above += c;
                }
            }
            __syncthreads();
            prefix |= (uint32_t) piv[0] << shift;
            mask |= 0xffu << shift;
            remaining = piv[1];
            __syncthreads();
        }
    }
    const uint32_t thr = prefix;
    // flat order sweep: element i alive iff key > thr, or key == thr and fewer than `remaining` equal keys precede it
    int carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < G; base += kBatchThreads) {
        const int64_t i = base + tid;
        bool gt = false, eq = false;
        if (i < G) {
            if (all_alive) gt = true;
            else { const uint32_t u = orderable(kb[i]); gt = u > thr; eq = u == thr; }
        }
        int tot;
        const int before = carry + block_excl_scan_1024(eq ? 1 : 0, wsum, tot);
        carry += tot;
        if (i < G && (gt || (eq && before < remaining))) {
            const int m = (int) (i % P), t = (int) ((i / P) % T), h = (int) (i / ((int64_t) P * T));
            const int b = h * P + m;
            atomicOr(&bits[(int64_t) t * wpr + (b >> 5)], 1u << (b & 31));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Multi-CTA form of the 'batch' top-k (one CTA per item above leaves 147 SMs idle for ~2 ms at BERT-base).  The same radix
// select, spread over ceil(G / 4096) CTAs per item and a handful of small launches; per-item scratch in a caller workspace:
//   u32 hist[4][256] | u32 state[8] = {prefix, mask, remaining, eq_total, ...} | u32 eq_cnt[chunks]
//   tkb_hist_kernel (x4)   keys matching the prefix found so far -> 8-bit digit histogram (shared memory, then global atomics)
//   tkb_pivot_kernel (x4)  one CTA per item: pivot digit of the pass, remaining count
//   tkb_count_kernel       keys == threshold per chunk;  tkb_scan_kernel: exclusive scan over the chunks of an item
//   tkb_bits_kernel        alive = key > thr, or key == thr and fewer than `remaining` equal keys precede it in flat index order;
//                          a warp owns 32 consecutive keys = one word of the bit mask when P % 32 == 0
// ------------------------------------------------------------------------------------------------
constexpr int kTkbChunk = 4096, kTkbThreads = 256, kTkbPer = kTkbChunk / kTkbThreads;
__host__ __device__ inline int64_t tkb_item_words(int64_t G) { return 4 * 256 + 8 + (G + kTkbChunk - 1) / kTkbChunk; }

__global__ void __launch_bounds__(kTkbThreads)
tkb_hist_kernel(const float* __restrict__ keys, uint32_t* __restrict__ ws, int64_t G, int pass) {
    __shared__ int hist[256];
    const int n = blockIdx.y, tid = threadIdx.x;
    uint32_t* wsn = ws + (int64_t) n * tkb_item_words(G);
    const uint32_t prefix = wsn[1024], mask = wsn[1025];
    if (wsn[1028] != 0u) return;                        // every key is alive: nothing to select
    hist[tid] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const float* kb = keys + (int64_t) n * G;
    const int64_t base = (int64_t) blockIdx.x * kTkbChunk;
#pragma unroll
    for (int u = 0; u < kTkbPer; ++u) {
        const int64_t i = base + u * kTkbThreads + tid;
        if (i < G) {
            const uint32_t key = orderable(__ldg(kb + i));
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
        }
    }
    __syncthreads();
    if (hist[tid] != 0) atomicAdd(&wsn[pass * 256 + tid], (uint32_t) hist[tid]);
}

__global__ void tkb_init_kernel(const float* __restrict__ k_per_item, uint32_t* __restrict__ ws, int64_t G, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    uint32_t* wsn = ws + (int64_t) n * tkb_item_words(G);
    const int64_t K = (int64_t) fminf(ceilf(k_per_item[n]), (float) G);
    wsn[1026] = (uint32_t) K;                           // remaining
    wsn[1028] = K >= G ? 1u : 0u;                       // every key alive
}

__global__ void __launch_bounds__(256)
tkb_pivot_kernel(uint32_t* __restrict__ ws, int64_t G, int pass) {
    __shared__ int wsum[8];
    __shared__ int piv[3];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* wsn = ws + (int64_t) n * tkb_item_words(G);
    if (wsn[1028] != 0u) return;
    const int remaining = (int) wsn[1026];
    const int mine = (int) wsn[pass * 256 + (255 - tid)];          // reversed: exclusive scan = keys with a larger digit
    int incl = warp_scan_incl_i(mine, lane);
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    int above = incl - mine;
    for (int w = 0; w < wid; ++w) above += wsum[w];
    if (above < remaining && remaining <= above + mine) { piv[0] = 255 - tid; piv[1] = remaining - above; piv[2] = mine; }
    __syncthreads();
    if (tid == 0) {
        const int shift = 24 - 8 * pass;
        wsn[1024] |= (uint32_t) piv[0] << shift;
        wsn[1025] |= 0xffu << shift;
        wsn[1026] = (uint32_t) piv[1];
        wsn[1027] = (uint32_t) piv[2];                  // keys equal to the threshold (after the last pass)
    }
}

__global__ void __launch_bounds__(kTkbThreads)
tkb_count_kernel(const float* __restrict__ keys, uint32_t* __restrict__ ws, int64_t G) {
    __shared__ int wsum[kTkbThreads / 32];
    const int n = blockIdx.y, tid = threadIdx.x;
    uint32_t* wsn = ws + (int64_t) n * tkb_item_words(G);
    if (wsn[1028] != 0u) return;
    const uint32_t thr = wsn[1024];
    const float* kb = keys + (int64_t) n * G;
    const int64_t base = (int64_t) blockIdx.x * kTkbChunk;
    int c = 0;
#pragma unroll
    for (int u = 0; u < kTkbPer; ++u) {
        const int64_t i = base + u * kTkbThreads + tid;
        if (i < G && orderable(__ldg(kb + i)) == thr) ++c;
    }
    c = warp_sum_i(c);
    if ((tid & 31) == 0) wsum[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kTkbThreads / 32; ++w) tot += wsum[w];
        wsn[1032 + blockIdx.x] = (uint32_t) tot;
    }
}

__global__ void __launch_bounds__(256)
tkb_scan_kernel(uint32_t* __restrict__ ws, int64_t G, int chunks) {
    __shared__ int wsum[8];
    __shared__ int carry_s;
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* wsn = ws + (int64_t) n * tkb_item_words(G);
    if (wsn[1028] != 0u) return;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < chunks; c0 += 256) {
        const int c = c0 + tid;
        const int v = c < chunks ? (int) wsn[1032 + c] : 0;
        const int incl = warp_scan_incl_i(v, lane);
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        int before = carry_s + incl - v;
        for (int w = 0; w < wid; ++w) before += wsum[w];
        if (c < chunks) wsn[1032 + c] = (uint32_t) before;
        __syncthreads();
        if (tid == 255) carry_s = before + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kTkbThreads)
tkb_bits_kernel(const float* __restrict__ keys, const uint32_t* __restrict__ ws, uint32_t* __restrict__ mask_bits, int64_t G, int H, int T, int P,
                int group_heads) {
    __shared__ int wsum[kTkbThreads / 32];
    __shared__ int carry_s;
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t* wsn = ws + (int64_t) n * tkb_item_words(G);
    const bool all_alive = wsn[1028] != 0u;
    const uint32_t thr = wsn[1024];
    const int remaining = (int) wsn[1026];
    const float* kb = keys + (int64_t) n * G;                  // group n = (item, head group): its group_heads * T * P keys are contiguous
    const int wpr = (H * P + 31) >> 5;
    const int gpi = H / group_heads;                           // groups per item
    const int hb = (n % gpi) * group_heads;                    // first head of the group
    uint32_t* bits = mask_bits + (int64_t) (n / gpi) * T * wpr;
    const int64_t base = (int64_t) blockIdx.x * kTkbChunk;
    if (tid == 0) carry_s = all_alive ? 0 : (int) wsn[1032 + blockIdx.x];
    __syncthreads();
    for (int u = 0; u < kTkbPer; ++u) {                 // consecutive keys per step so that the flat index order is the scan order
        const int64_t i = base + u * kTkbThreads + tid;
        bool gt = false, eq = false;
        if (i < G) {
            if (all_alive) gt = true;
            else { const uint32_t key = orderable(__ldg(kb + i)); gt = key > thr; eq = key == thr; }
        }
        const uint32_t eqb = __ballot_sync(kFull, eq);
        if (lane == 0) wsum[wid] = __popc(eqb);
        __syncthreads();
        int before = carry_s + __popc(eqb & ((1u << lane) - 1u));
        int tot = 0;
        for (int w = 0; w < kTkbThreads / 32; ++w) { const int sct = wsum[w]; if (w < wid) before += sct; tot += sct; }
        const bool alive = i < G && (gt || (eq && before < remaining));
        if ((P & 31) == 0) {
            // 32 consecutive keys of one (h, t) row = one word of the mask
            const uint32_t word = __ballot_sync(kFull, alive);
            const int64_t i0 = i - lane;
            if (lane == 0 && i0 < G) {
                const int m0 = (int) (i0 % P), t = (int) ((i0 / P) % T), h = hb + (int) (i0 / ((int64_t) P * T));
                bits[(int64_t) t * wpr + ((h * P + m0) >> 5)] = word;
            }
        } else if (alive) {
            const int m = (int) (i % P), t = (int) ((i / P) % T), h = hb + (int) (i / ((int64_t) P * T));
            const int b = h * P + m;
            atomicOr(&bits[(int64_t) t * wpr + (b >> 5)], 1u << (b & 31));
        }
        __syncthreads();
        if (tid == 0) carry_s += tot;
        __syncthreads();
    }
}

// a13': avg[n,h,:] = sum_j w_j v[n,h,j,:],  w = dense resize of mean_t probs[n,h,t,:] to T (no padding)
template <typename T>
__global__ void __launch_bounds__(256)
bert_avg_kernel(const float* __restrict__ probs, const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                T* __restrict__ avg, int H, int Tn, int P, int D, const int32_t* __restrict__ lengths) {
    extern __shared__ __align__(16) float smem[];
    float* pm = smem;             // [P]
    float* part = pm + P;         // [8][D]
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const float* pb = probs + (int64_t) nh * Tn * P;
    for (int m = threadIdx.x; m < P; m += 256) {
        float s = 0.f;
        for (int t = 0; t < Tn; ++t) s += pb[(int64_t) t * P + m];
        pm[m] = s / (float) Tn;
    }
    __syncthreads();
    const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int L = lengths ? max(1, min(lengths[n], Tn)) : Tn;      // right-padded batch: columns >= L read the fill value 0 and v is 0 there
    for (int c = lane; c < D; c += 32) {
        float s = 0.f;
        for (int j = wid; j < L; j += 8) {
            // floor(((cs - 1) + 0.5) / L * P - 1e-4), cs = j + 1, L = valid tokens   (resize_m_to_t.py:36-47)
            const float a = __fadd_rn(__fsub_rn((float) (j + 1), 1.0f), 0.5f);
            int idx = (int) floorf(__fsub_rn(__fmul_rn(__fdiv_rn(a, (float) L), (float) P), 1e-4f));
            idx = max(0, min(idx, P - 1));
            s = fmaf(pm[idx], to_f32(vb[(int64_t) j * v_st + c]), s);
        }
        part[wid * D + c] = s;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += 256) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += part[w * D + c];
        avg[(int64_t) nh * D + c] = from_f32<T>(s);
    }
}

static NcPerfDims nc_dims(int N, int H, int T, int D, int F) {
    NcPerfDims dm;
    dm.N = N; dm.H = H; dm.T = T; dm.D = D; dm.F = F;
    dm.Fp = (F + 3) & ~3;
    dm.E = 2 * D;
    dm.nchunks = (T + kNcRows - 1) / kNcRows;
    dm.slot = (int64_t) dm.Fp * dm.E + dm.Fp;
    dm.lengths = nullptr;
    return dm;
}

}  // namespace sea

using namespace sea;

extern "C" {

int64_t sea_performer_noncausal_workspace_floats(int N, int H, int T, int D, int F) {
    if (N <= 0 || H <= 0 || T <= 0 || D <= 0 || F <= 0) return 0;
    NcPerfDims dm = nc_dims(N, H, T, D, F);
    return (int64_t) N * H * dm.nchunks * dm.slot + (int64_t) N * H + 16;
}

int sea_performer_noncausal_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                const float* proj, int dtype, void* ctx, float* workspace,
                                int N, int H, int T, int D, int F, void* stream) {
    return sea_performer_noncausal_len_fwd(q, q_sn, q_sh, q_st, k, k_sn, k_sh, k_st, v, v_sn, v_sh, v_st, proj, dtype, ctx, workspace, nullptr,
                                           N, H, T, D, F, stream);
}

int sea_performer_noncausal_len_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                    const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                    const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                    const float* proj, int dtype, void* ctx, float* workspace, const int32_t* lengths,
                                    int N, int H, int T, int D, int F, void* stream) {
    SEA_CHECK_ARG(q && k && v && proj && ctx && workspace, "sea_performer_noncausal_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && D > 0 && F > 0 && (D & 3) == 0 && (int64_t) N * H <= 65535, "sea_performer_noncausal_fwd: bad shape");
    NcPerfDims dm = nc_dims(N, H, T, D, F);
    dm.lengths = lengths;
    const size_t sm1 = ((size_t) D * dm.Fp + (size_t) kNcRows * D) * 4;
    const size_t sm2 = ((size_t) D * dm.Fp + (size_t) kNcRows * D + kNcRows + (size_t) dm.Fp * kNcRows + (size_t) kNcRows * dm.E) * 4;
    const size_t regionA = (size_t) D * dm.Fp > (size_t) dm.Fp * dm.E + dm.Fp ? (size_t) D * dm.Fp : (size_t) dm.Fp * dm.E + dm.Fp;
    const size_t sm3 = (regionA + (size_t) kNcRows * D + kNcRows + (size_t) kNcRows * dm.Fp + kNcRows) * 4;
    SEA_CHECK_ARG(sm3 <= 227 * 1024 && sm2 <= 227 * 1024,
                  "sea_performer_noncausal_fwd: D=%d F=%d needs %zu B of shared memory (> 227 KB)", D, F, sm3);
    cudaStream_t s = (cudaStream_t) stream;
    float* partial = workspace;
    int* kmax = reinterpret_cast<int*>(workspace + (int64_t) N * H * dm.nchunks * dm.slot);
    // ordered-int encoding of -inf
    SEA_CUDA_TRY(cudaMemsetAsync(kmax, 0x80, (size_t) N * H * sizeof(int), s), "memset");
    dim3 grid(dm.nchunks, N * H);
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto k1 = nc_kmax_kernel<T_>;
        auto k2 = nc_ksum_kernel<T_>;
        auto k3 = nc_out_kernel<T_>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm1), "smem attr");
        SEA_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm2), "smem attr");
        SEA_CUDA_TRY(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm3), "smem attr");
        k1<<<grid, kNcThreads, sm1, s>>>((const T_*) k, k_sn, k_sh, k_st, proj, kmax, dm);
        SEA_CHECK_LAUNCH("nc_kmax_kernel");
        k2<<<grid, kNcThreads, sm2, s>>>((const T_*) k, k_sn, k_sh, k_st, (const T_*) v, v_sn, v_sh, v_st, proj, kmax, partial, dm);
        SEA_CHECK_LAUNCH("nc_ksum_kernel");
        nc_reduce_kernel<<<dim3((unsigned) ((dm.slot + 255) / 256), N * H), 256, 0, s>>>(partial, dm.nchunks, dm.slot);
        SEA_CHECK_LAUNCH("nc_reduce_kernel");
        k3<<<grid, kNcThreads, sm3, s>>>((const T_*) q, q_sn, q_sh, q_st, proj, partial, (T_*) ctx, dm);
        SEA_CHECK_LAUNCH("nc_out_kernel");
    });
    return SEA_OK;
}

int sea_conv3x3_cl(const void* x, const float* weight, const float* bias, void* y, int dtype,
                   int N, int Tin, int Tout, int W, int C, int O, int stride_t, int up, int relu, void* stream) {
    SEA_CHECK_ARG(x && weight && bias && y, "sea_conv3x3_cl: null pointer");
    SEA_CHECK_ARG(N > 0 && Tin > 0 && Tout > 0 && W > 0 && C > 0 && O > 0 && stride_t >= 1 && up >= 1, "sea_conv3x3_cl: bad shape");
    // weights in shared memory when they fit (BERT-base: 9 x 48 x 48 fp32 = 83 KB); kOT = 12 output channels per thread when O allows
    const int ot = (O % 12 == 0) ? 12 : 4;
    const int OG = (O + ot - 1) / ot;
    const size_t wsm = (size_t) 9 * C * OG * ot * sizeof(float);
    if (wsm <= 160 * 1024 && OG <= 256) {
        const int ppc = 256 / OG;
        const int64_t pixels = (int64_t) N * Tout * W;
        SEA_DISPATCH_DTYPE(dtype, T_, {
            if (ot == 12) {
                auto kern = conv3x3_cl_smem_kernel<T_, 12>;
                SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) wsm), "smem attr");
                kern<<<cdiv(pixels, ppc), 256, wsm, (cudaStream_t) stream>>>((const T_*) x, weight, bias, (T_*) y, N, Tin, Tout, W, C, O, stride_t, up, relu);
            } else {
                auto kern = conv3x3_cl_smem_kernel<T_, 4>;
                SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) wsm), "smem attr");
                kern<<<cdiv(pixels, ppc), 256, wsm, (cudaStream_t) stream>>>((const T_*) x, weight, bias, (T_*) y, N, Tin, Tout, W, C, O, stride_t, up, relu);
            }
            SEA_CHECK_LAUNCH("conv3x3_cl_smem_kernel");
        });
        return SEA_OK;
    }
    const int64_t total = (int64_t) N * Tout * W * ((O + 3) / 4);
    SEA_DISPATCH_DTYPE(dtype, T_, {
        conv3x3_cl_kernel<T_><<<cdiv(total, 256), 256, 0, (cudaStream_t) stream>>>((const T_*) x, weight, bias, (T_*) y, N, Tin, Tout, W, C, O,
                                                                                      stride_t, up, relu);
        SEA_CHECK_LAUNCH("conv3x3_cl_kernel");
    });
    return SEA_OK;
}

int sea_bert_tail_fwd(const void* y, int dtype, float* probs, float* scores, int N, int H, int Tin, int Win, int T, int P, void* stream) {
    SEA_CHECK_ARG(y && probs, "sea_bert_tail_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && Tin > 0 && Win > 0 && T > 0 && P > 0 && P <= 1024, "sea_bert_tail_fwd: bad shape");
    const int64_t warps = (int64_t) N * H * T;
    SEA_DISPATCH_DTYPE(dtype, T_, {
        bert_tail_kernel<T_><<<cdiv(warps * 32, 256), 256, 0, (cudaStream_t) stream>>>((const T_*) y, probs, scores, N, H, Tin, Win, T, P);
        SEA_CHECK_LAUNCH("bert_tail_kernel");
    });
    return SEA_OK;
}

int sea_topk_mask_bits_batch(const float* keys, const float* k_per_item, uint32_t* mask_bits, int N, int H, int T, int P, void* stream) {
    SEA_CHECK_ARG(keys && k_per_item && mask_bits && N > 0 && H > 0 && T > 0 && P > 0, "sea_topk_mask_bits_batch: bad argument");
    SEA_CHECK_ARG((int64_t) H * T * P < (1ll << 31), "sea_topk_mask_bits_batch: group too large");
    topk_batch_kernel<<<N, kBatchThreads, 0, (cudaStream_t) stream>>>(keys, k_per_item, mask_bits, H, T, P);
    SEA_CHECK_LAUNCH("topk_batch_kernel");
    return SEA_OK;
}

int64_t sea_topk_batch_workspace_bytes(int N, int H, int T, int P, int group_heads) {
    if (N <= 0 || H <= 0 || T <= 0 || P <= 0 || group_heads <= 0 || H % group_heads != 0) return 0;
    return (int64_t) N * (H / group_heads) * tkb_item_words((int64_t) group_heads * T * P) * 4;
}

int sea_topk_mask_bits_batch_ws(const float* keys, const float* k_per_group, uint32_t* mask_bits, void* workspace, int64_t workspace_bytes,
                                int N, int H, int T, int P, int group_heads, void* stream) {
    SEA_CHECK_ARG(keys && k_per_group && mask_bits && workspace && N > 0 && H > 0 && T > 0 && P > 0, "sea_topk_mask_bits_batch_ws: bad argument");
    SEA_CHECK_ARG(group_heads > 0 && H % group_heads == 0, "sea_topk_mask_bits_batch_ws: group_heads must divide H");
    const int64_t G = (int64_t) group_heads * T * P;           // keys per group
    const int groups = N * (H / group_heads);
    SEA_CHECK_ARG(G < (1ll << 31) && groups <= 65535, "sea_topk_mask_bits_batch_ws: group too large");
    SEA_CHECK_ARG(workspace_bytes >= sea_topk_batch_workspace_bytes(N, H, T, P, group_heads), "sea_topk_mask_bits_batch_ws: workspace too small");
    cudaStream_t s = (cudaStream_t) stream;
    uint32_t* ws = reinterpret_cast<uint32_t*>(workspace);
    const int chunks = (int) ((G + kTkbChunk - 1) / kTkbChunk);
    const int wpr = (H * P + 31) >> 5;
    SEA_CUDA_TRY(cudaMemsetAsync(ws, 0, (size_t) sea_topk_batch_workspace_bytes(N, H, T, P, group_heads), s), "memset workspace");
    if ((P & 31) != 0) SEA_CUDA_TRY(cudaMemsetAsync(mask_bits, 0, (size_t) N * T * wpr * 4, s), "memset bits");
    const dim3 grid((unsigned) chunks, (unsigned) groups);
    tkb_init_kernel<<<(groups + 127) / 128, 128, 0, s>>>(k_per_group, ws, G, groups);
    for (int pass = 0; pass < 4; ++pass) {
        tkb_hist_kernel<<<grid, kTkbThreads, 0, s>>>(keys, ws, G, pass);
        tkb_pivot_kernel<<<groups, 256, 0, s>>>(ws, G, pass);
    }
    tkb_count_kernel<<<grid, kTkbThreads, 0, s>>>(keys, ws, G);
    tkb_scan_kernel<<<groups, 256, 0, s>>>(ws, G, chunks);
    tkb_bits_kernel<<<grid, kTkbThreads, 0, s>>>(keys, ws, mask_bits, G, H, T, P, group_heads);
    SEA_CHECK_LAUNCH("tkb kernels");
    return SEA_OK;
}

int sea_bert_avg_fwd(const float* probs, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype, void* avg,
                     int N, int H, int T, int P, int D, void* stream) {
    return sea_bert_avg_len_fwd(probs, v, v_sn, v_sh, v_st, dtype, avg, nullptr, N, H, T, P, D, stream);
}

int sea_bert_avg_len_fwd(const float* probs, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype, void* avg, const int32_t* lengths,
                         int N, int H, int T, int P, int D, void* stream) {
    SEA_CHECK_ARG(probs && v && avg && N > 0 && H > 0 && T > 0 && P > 0 && D > 0, "sea_bert_avg_fwd: bad argument");
    const size_t smem = ((size_t) P + 8 * (size_t) D) * 4;
    SEA_DISPATCH_DTYPE(dtype, T_, {
        bert_avg_kernel<T_><<<N * H, 256, smem, (cudaStream_t) stream>>>(probs, (const T_*) v, v_sn, v_sh, v_st, (T_*) avg, H, T, P, D, lengths);
        SEA_CHECK_LAUNCH("bert_avg_kernel");
    });
    return SEA_OK;
}

}  // extern "C"
