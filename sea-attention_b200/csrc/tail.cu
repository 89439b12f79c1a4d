// a5 tail + a6 + a7 fused (bf16 path): one CTA per query row (n, t) holds that row's H x P estimated scores on
// chip from the 1x1-conv output to the top-k bit mask:
//   y3 [N,T,W,H] fp32 (1x1 conv evaluated before the x4 nearest upsample, csrc/umma_conv.cu)
//   -> upsample x(P/W), two bias-valued pad columns, area resize P+2 -> P   (attention.py:275-277, modules.py:12-31,42-55)
//   -> LayerNorm(P) (attention.py:280) -> softmax(P) (attention.py:670-673)  -> probs fp32 [N,H,T,P] (optional store)
//   -> grouped top-k over the H*P keys of the row (attention.py:843-917)      -> bit mask [N,T,H*P/32]
// The probabilities are read back from HBM by nobody: the top-k consumes them from shared memory.
#include "common.cuh"
#include "topk.cuh"
#include "csr_common.cuh"

#include <stdlib.h>

namespace sea {

template <int kPerLane, int kUp>
__global__ void __launch_bounds__(kTopkThreads)
tail_topk_kernel(const float* __restrict__ y3, const float* __restrict__ bias, const float* __restrict__ ln_w,
                 const float* __restrict__ ln_b, const float* __restrict__ k_per_row, float* __restrict__ probs,
                 uint32_t* __restrict__ mask_bits, int32_t* __restrict__ crow_counts, int k_clamp,
                 int N, int H, int Tn, int W, int P_rt) {
    // P (= 32 * kPerLane) and the upsample factor are compile-time constants: every / and % below folds into
    // shifts / multiplies (a CTA lives for one row only, so per-thread set-up divisions are not amortised)
    constexpr int P = 32 * kPerLane;
    extern __shared__ __align__(16) uint32_t smem_u[];
    __shared__ int hist[256];
    __shared__ int scratch[16];
    const int G = H * P;
    uint32_t* skeys = smem_u;                                   // [G]
    uint32_t* sbits = skeys + G;                                // [G/32]
    float* ys = reinterpret_cast<float*>(sbits + (G >> 5));     // [H][W+2]: W conv outputs, then the bias (pad columns), then 0
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.x / Tn, t = blockIdx.x % Tn;
    const int ldy = W + 2;
    const float* yr = y3 + ((int64_t) n * Tn + t) * W * H;
    {
        int h = tid % H, w = tid / H;
        const int dh = kTopkThreads % H, dw = kTopkThreads / H;
        for (int idx = tid; idx < W * H; idx += kTopkThreads) {
            ys[h * ldy + w] = yr[idx];
            h += dh; w += dw;
            if (h >= H) { h -= H; ++w; }
        }
    }
    for (int h = tid; h < H; h += kTopkThreads) { ys[h * ldy + W] = bias[h]; ys[h * ldy + W + 1] = 0.f; }
    __syncthreads();
    constexpr int PW = P + 2;
    const int up = kUp > 0 ? kUp : P / W;
    // area-resize window of each of this lane's columns, resolved ONCE into (up to) three slots of the per-head
    // vector: a conv output w, the bias slot W (the zero-padded columns of the 1x1 conv) or the zero slot W+1
    int tap[kPerLane][3];
    float rc[kPerLane], lw[kPerLane], lb[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int j = lane + 32 * i;
        const int st = (j * PW) / P;
        const int cnt = ((j + 1) * PW + P - 1) / P - st;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int pcol = st + c;
            tap[i][c] = c >= cnt ? W + 1 : ((pcol == 0 || pcol == PW - 1) ? W : (pcol - 1) / up);
        }
        rc[i] = 1.0f / (float) cnt;
        lw[i] = ln_w[j];
        lb[i] = ln_b[j];
    }
    const float invP = 1.0f / (float) P;
    for (int h = wid; h < H; h += kTopkThreads / 32) {
        const float* yh = ys + h * ldy;
        float val[kPerLane];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            val[i] = (yh[tap[i][0]] + yh[tap[i][1]] + yh[tap[i][2]]) * rc[i];
            s += val[i];
        }
        const float mean = warp_sum(s) * invP;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { const float d = val[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * invP + 1e-5f);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { val[i] = (val[i] - mean) * rstd * lw[i] + lb[i]; mx = fmaxf(mx, val[i]); }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { val[i] = __expf(val[i] - mx); sum += val[i]; }
        const float inv = 1.0f / warp_sum(sum);
        float* prow = probs ? probs + (((int64_t) n * H + h) * Tn + t) * P : nullptr;
        uint32_t* krow = skeys + h * P;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            const int j = lane + 32 * i;
            const float pr = val[i] * inv;
            if (prow) prow[j] = pr;
            krow[j] = orderable(pr);
        }
    }
    if (mask_bits == nullptr) return;
    __syncthreads();
    const float kf = k_per_row[blockIdx.x];
    const int K = (int) fminf(ceilf(kf), (float) G);
    topk_select_to_bits(skeys, G, K, sbits, hist, scratch);
    __syncthreads();
    uint32_t* out_row = mask_bits + (int64_t) blockIdx.x * (G >> 5);
    int cnt = 0;
    const float sc = __fdiv_rn((float) (t + 1), (float) P);       // causal prefill: source length of row t is t+1
    for (int w = tid; w < (G >> 5); w += kTopkThreads) {
        const uint32_t word = sbits[w];
        out_row[w] = word;
        if (crow_counts != nullptr) cnt += word_width_sum(word, w, P, sc, k_clamp);
    }
    if (crow_counts != nullptr) {
        // a8 pass 1 fused: crow[n, t+1] = entries of this row (sea_crow_scan turns the counts into offsets)
        cnt = warp_sum_i(cnt);
        if (lane == 0) scratch[wid] = cnt;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
#pragma unroll
            for (int i = 0; i < kTopkThreads / 32; ++i) tot += scratch[i];
            crow_counts[(int64_t) n * (Tn + 1) + t + 1] = tot;
            if (t == 0) crow_counts[(int64_t) n * (Tn + 1)] = 0;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Register-resident variant (H % 8 == 0, (H/8) * (P/32) <= 32 keys per thread) -- the production path at the OPT shapes.
// The kernel above streams the H*P keys of a row through shared memory five times (or/and, 4 radix passes, bit build) and
// is issue-bound (ncu: 79 % issue-active, ~45 k warp instructions per row).  Here warp w owns heads w, w+8, ...; a lane
// owns P/32 CONSECUTIVE pixels of each of them, so the keys never leave registers:
//   * LayerNorm / softmax per head with warp reductions, probabilities stored as 16-byte vectors,
//   * radix select with 8-bit digits that START at the highest bit in which the row's keys differ (all digits carry
//     entropy -> no serialised shared-memory atomics on a constant exponent byte),
//   * alive bits assembled with shuffles inside 32/(P/32)-lane groups; ties at the threshold are cut in flat-index order
//     (lower h*P+m wins) with one warp scan per head.
// ------------------------------------------------------------------------------------------------
constexpr float kLog2eT = 1.4426950408889634f;
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr int kMaxCand = 1024;

// Warp totals of R per-lane values at once (every lane ends with all R totals).  A butterfly that halves the number of values a lane
// carries while it still has more than one (lanes with bit 4 set keep the upper heads, ...) needs 9 shuffles for R = 4 and 6 for R = 2
// where R independent butterflies need 5 R.
template <int R>
__device__ __forceinline__ void warp_sum_multi(float (&r)[R], int lane) {
    if constexpr (R == 4) {
        const bool hi = (lane & 16) != 0, h8 = (lane & 8) != 0;
        float k0 = hi ? r[2] : r[0], k1 = hi ? r[3] : r[1];
        k0 += __shfl_xor_sync(kFull, hi ? r[0] : r[2], 16);
        k1 += __shfl_xor_sync(kFull, hi ? r[1] : r[3], 16);
        float kk = h8 ? k1 : k0;
        kk += __shfl_xor_sync(kFull, h8 ? k0 : k1, 8);
        kk += __shfl_xor_sync(kFull, kk, 4);
        kk += __shfl_xor_sync(kFull, kk, 2);
        kk += __shfl_xor_sync(kFull, kk, 1);                  // total of head 2 * [bit 4] + [bit 3]
        const float o8 = __shfl_xor_sync(kFull, kk, 8);
        const float a = h8 ? o8 : kk, b = h8 ? kk : o8;        // heads 2 * [bit 4] + {0, 1}
        const float a16 = __shfl_xor_sync(kFull, a, 16), b16 = __shfl_xor_sync(kFull, b, 16);
        r[0] = hi ? a16 : a; r[1] = hi ? b16 : b; r[2] = hi ? a : a16; r[3] = hi ? b : b16;
    } else if constexpr (R == 2) {
        const bool hi = (lane & 16) != 0;
        float kk = hi ? r[1] : r[0];
        kk += __shfl_xor_sync(kFull, hi ? r[0] : r[1], 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) kk += __shfl_xor_sync(kFull, kk, o);
        const float o16 = __shfl_xor_sync(kFull, kk, 16);
        r[0] = hi ? o16 : kk; r[1] = hi ? kk : o16;
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int hh = 0; hh < R; ++hh) r[hh] += __shfl_xor_sync(kFull, r[hh], o);
        }
    }
}
// warp maximum in one instruction (CREDUX.MAX.F32; sm_100a)
__device__ __forceinline__ float warp_max_redux(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

template <int kPerLane, int kHPW, int kUp, bool kExactH, int kMinBlocks = 3>
__global__ void __launch_bounds__(kTopkThreads, kMinBlocks)
tail_topk_reg_kernel(const float* __restrict__ y3, const float* __restrict__ bias, const float* __restrict__ ln_w,
                     const float* __restrict__ ln_b, const float* __restrict__ k_per_row, float* __restrict__ probs,
                     uint32_t* __restrict__ mask_bits, int32_t* __restrict__ crow_counts, int k_clamp,
                     int N, int Tn, int W, int Hr) {
    // Hmax = 8 * kHPW head slots; Hr <= Hmax real heads (kExactH: Hr == Hmax, everything below folds to constants).  The slots
    // h >= Hr hold no keys: they are skipped in every key loop and their alive bits stay 0.
    constexpr int P = 32 * kPerLane, Hmax = 8 * kHPW;
    static_assert(Hmax <= 32, "one lane per head in the tie-cut scan");
    const int H = kExactH ? Hmax : Hr, G = H * P;
    constexpr int kLanesPerWord = 32 / kPerLane > 0 ? 32 / kPerLane : 1;     // lanes that share one 32-pixel word (kPerLane <= 32)
    extern __shared__ __align__(16) uint32_t smem_u[];
    __shared__ int hist[256];
    __shared__ int scratch[16];
    __shared__ int head_eq[Hmax];
    __shared__ uint32_t cand[kMaxCand];
    __shared__ uint32_t s_orand[2];
    __shared__ int4 stap[P];
    float* ys = reinterpret_cast<float*>(smem_u);               // [H][W+2]: W conv outputs, then the bias (pad columns), then 0
    uint32_t* sbits = smem_u + H * (W + 3);                     // [G/32] (only for the fused row counts)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.x / Tn, t = blockIdx.x % Tn;
    pdl_launch_dependents();
    pdl_wait();
    hist[tid] = 0;
    if (tid == 0) { s_orand[0] = 0u; s_orand[1] = 0xffffffffu; scratch[11] = 0; }       // scratch[11]: candidate counter (used once)
    const int ldy = kUp > 0 ? ((P / (kUp > 0 ? kUp : 1) + 2) | 1) : ((W + 2) | 1);        // compile-time when the upsample factor is (W == P / kUp)
    const float* yr = y3 + ((int64_t) n * Tn + t) * W * H;
    // The first batch of conv outputs (8 loads per thread = the whole row at the north-star shape) and the LayerNorm parameters
    // are requested before the tap table is computed, so that table's integer divisions run under the load latency, and one
    // barrier covers both the row image and the table.
    float tmp0[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int idx = u * kTopkThreads + tid;
        tmp0[u] = idx < W * H ? __ldg(yr + idx) : 0.f;
    }
    float lw[kPerLane], lb[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int j = lane * kPerLane + i;
        lw[i] = __ldg(ln_w + j);
        lb[i] = __ldg(ln_b + j);
    }
    constexpr int PW = P + 2;
    const int up = kUp > 0 ? kUp : P / W;
    // 4x upsampling with a lane owning a multiple of 4 pixels: the area-resize windows are known at compile time.  Column j averages the
    // padded columns {j, j+1} (j < P/2) or {j+1, j+2} (j >= P/2) -- floor(j (P+2) / P) = j + [j >= P/2], two columns everywhere -- and
    // padded column c is conv output (c-1)/4 (the bias for c = 0 and c = P+1).  A lane then needs its own kPerLane/4 conv outputs and ONE
    // neighbour (the previous one in the lower half of the row, the next one in the upper half) instead of 3 table-driven loads per pixel.
    constexpr bool kStaticTaps = kUp == 4 && kPerLane % 4 == 0;
    // area-resize window of every output column, resolved once per CTA (thread = column) into three slots of the per-head
    // vector: a conv output w, the bias slot W (the zero-padded columns of the 1x1 conv) or the zero slot W+1
    if constexpr (!kStaticTaps) {
        for (int j = tid; j < P; j += kTopkThreads) {
            const int st = (j * PW) / P;
            const int cnt = ((j + 1) * PW + P - 1) / P - st;
            int tp[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int pcol = st + c;
                tp[c] = c >= cnt ? W + 1 : ((pcol == 0 || pcol == PW - 1) ? W : (pcol - 1) / up);
            }
            stap[(j % kPerLane) * 32 + j / kPerLane] = make_int4(tp[0], tp[1], tp[2], __float_as_int(1.0f / (float) cnt));   // [i][lane]: conflict-free reads
        }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int idx = u * kTopkThreads + tid;
        if (idx < W * H) ys[(idx % H) * ldy + idx / H] = tmp0[u];
    }
    for (int base = 8 * kTopkThreads; base < W * H; base += 8 * kTopkThreads) {
        // 8 independent loads in flight per thread before the first (transposing) shared-memory store
        float tmp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * kTopkThreads + tid;
            tmp[u] = idx < W * H ? __ldg(yr + idx) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = base + u * kTopkThreads + tid;
            if (idx < W * H) ys[(idx % H) * ldy + idx / H] = tmp[u];
        }
    }
    for (int h = tid; h < H; h += kTopkThreads) { ys[h * ldy + W] = bias[h]; ys[h * ldy + W + 1] = 0.f; }
    __syncthreads();
    int tap[kStaticTaps ? 1 : kPerLane][3];
    float rc[kStaticTaps ? 1 : kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        if constexpr (!kStaticTaps) {
            const int4 tp = stap[i * 32 + lane];
            tap[i][0] = tp.x; tap[i][1] = tp.y; tap[i][2] = tp.z;
            rc[i] = __int_as_float(tp.w);
        }
        lw[i] *= kLog2eT;                   // softmax in the log2 domain: exp(x - max) == exp2(x * log2e - max * log2e)
        lb[i] *= kLog2eT;
    }
    constexpr int kOwn = kPerLane / 4 > 0 ? kPerLane / 4 : 1;       // conv outputs under a lane's pixels (kStaticTaps)
    const bool lower = lane < 16;
    const int own0 = lane * kOwn;
    const int nb_idx = lower ? (lane == 0 ? W : own0 - 1) : (lane == 31 ? W : own0 + kOwn);      // the neighbour; slot W holds the bias
    constexpr float invP = 1.0f / (float) P;
    // The kHPW heads of this warp advance through LayerNorm / softmax stage by stage, so that the kHPW warp reductions of a
    // stage are independent shuffle chains in flight together (a head-by-head loop exposes 4 x 5 dependent shuffles per head).
    float val[kHPW][kPerLane];
    float red[kHPW];
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        const bool hv = kExactH || wid + 8 * hh < H;                // warp-uniform
        const float* yh = ys + (hv ? wid + 8 * hh : 0) * ldy;
        float s = 0.f;
        if constexpr (kStaticTaps) {
            float c[kOwn];
#pragma unroll
            for (int r = 0; r < kOwn; ++r) c[r] = yh[own0 + r];
            const float nb = yh[nb_idx];
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) {
                // (x + x + 0) * 0.5 == x: pixels whose two columns fall on the same conv output copy it, like the table path computes it
                const float vl = i == 0 ? (nb + c[0]) * 0.5f : ((i - 1) / 4 == i / 4 ? c[i / 4] : (c[(i - 1) / 4] + c[i / 4]) * 0.5f);
                const float vu = i == kPerLane - 1 ? (c[kOwn - 1] + nb) * 0.5f : ((i + 1) / 4 == i / 4 ? c[i / 4] : (c[i / 4] + c[(i + 1) / 4 < kOwn ? (i + 1) / 4 : 0]) * 0.5f);
                val[hh][i] = lower ? vl : vu;
                s += val[hh][i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) {
                val[hh][i] = (yh[tap[i][0]] + yh[tap[i][1]] + yh[tap[i][2]]) * rc[i];
                s += val[hh][i];
            }
        }
        red[hh] = s;
    }
    warp_sum_multi<kHPW>(red, lane);
    float mean[kHPW];
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        mean[hh] = red[hh] * invP;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { const float d = val[hh][i] - mean[hh]; q = fmaf(d, d, q); }
        red[hh] = q;
    }
    warp_sum_multi<kHPW>(red, lane);
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        const float rstd = rsqrtf(red[hh] * invP + 1e-5f);
        const float nmr = -mean[hh] * rstd;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { val[hh][i] = fmaf(fmaf(val[hh][i], rstd, nmr), lw[i], lb[i]); mx = fmaxf(mx, val[hh][i]); }
        red[hh] = mx;
    }
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) red[hh] = warp_max_redux(red[hh]);
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        const float mx = red[hh];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { val[hh][i] = ex2_fast(val[hh][i] - mx); sum += val[hh][i]; }
        red[hh] = sum;
    }
    warp_sum_multi<kHPW>(red, lane);
    uint32_t key[kHPW][kPerLane];
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        const int h = wid + 8 * hh;
        const float inv = 1.0f / red[hh];
        const bool hv = kExactH || h < H;
        // probabilities are >= +0, so their bit patterns already order like the values: with every head slot in use the key IS the
        // probability's register (no second 32-register array); with empty slots the sign bit is set to keep real keys above the 0 of a slot
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            val[hh][i] *= inv;
            key[hh][i] = kExactH ? __float_as_uint(val[hh][i]) : (hv ? (__float_as_uint(val[hh][i]) | 0x80000000u) : 0u);
        }
        if (probs && hv) {
            float* prow = probs + (((int64_t) n * H + h) * Tn + t) * P + lane * kPerLane;
            if constexpr (kPerLane % 4 == 0) {
#pragma unroll
                for (int i = 0; i < kPerLane; i += 4)       // streaming store: nothing on the hot path reads the probabilities back, keep them from evicting q / k / v
                    __stcs(reinterpret_cast<float4*>(prow + i), make_float4(val[hh][i], val[hh][i + 1], val[hh][i + 2], val[hh][i + 3]));
            } else {
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) prow[i] = val[hh][i];
            }
        }
    }
    if (mask_bits == nullptr) return;
    const float kf = k_per_row[blockIdx.x];
    const int K = (int) fminf(ceilf(kf), (float) G);

    // ---- K-th largest key: or/and -> highest differing bit -> 8-bit radix passes from there ------------------------
    uint32_t thr = 0u;
    int remaining = K, eq_total = G;
    bool all_alive = K >= G;
    if (!all_alive) {
        uint32_t k_or = 0u, k_and = 0xffffffffu;
#pragma unroll
        for (int hh = 0; hh < kHPW; ++hh)
            if (kExactH || wid + 8 * hh < H) {
#pragma unroll
                for (int i = 0; i < kPerLane; ++i) { k_or |= key[hh][i]; k_and &= key[hh][i]; }
            }
        k_or = __reduce_or_sync(kFull, k_or);
        k_and = __reduce_and_sync(kFull, k_and);
        if (lane == 0) { atomicOr(&s_orand[0], k_or); atomicAnd(&s_orand[1], k_and); }       // initialised before the first barrier
        __syncthreads();
        k_or = s_orand[0]; k_and = s_orand[1];
        const uint32_t diff = k_or ^ k_and;
        thr = k_and;                                             // bits above the first difference are common
        if (diff != 0u) {
            int top = 32 - __clz(diff);                          // bits [0, top) still unresolved
            thr &= top >= 32 ? 0u : (0xffffffffu << top);
            bool first = true;
            while (top > 0) {
                const int wd = min(8, top), sh = top - wd;
                const uint32_t hi_mask = top >= 32 ? 0u : (0xffffffffu << top);
                const uint32_t dmask = (1u << wd) - 1u;
                if (!first) {                                    // (zeroed at kernel start for the first pass)
                    hist[tid] = 0;
                    __syncthreads();
                }
                if (first) {                                     // every key carries the common prefix
#pragma unroll
                    for (int hh = 0; hh < kHPW; ++hh)
                        if (kExactH || wid + 8 * hh < H) {
#pragma unroll
                            for (int i = 0; i < kPerLane; ++i) atomicAdd(&hist[(key[hh][i] >> sh) & dmask], 1);
                        }
                } else {
#pragma unroll
                    for (int hh = 0; hh < kHPW; ++hh)
#pragma unroll
                        for (int i = 0; i < kPerLane; ++i) {
                            const uint32_t u = key[hh][i];
                            if ((u & hi_mask) == thr) atomicAdd(&hist[(u >> sh) & dmask], 1);
                        }
                }
                __syncthreads();
                int pivot_digit;
                {
                    // pivot digit, by every warp for itself (8 shared loads and a scan, instead of one warp + a broadcast barrier):
                    // lane l owns digits 255-8l .. 248-8l (descending), suffix counts by warp scan; exactly one (lane, j) matches
                    int c[8], loc = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; loc += c[j]; }
                    int above = warp_scan_incl_i(loc, lane) - loc;          // keys with a digit above this lane's range
                    int dg = -1, rem = 0, eqt = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (above < remaining && remaining <= above + c[j]) {
                            dg = 255 - 8 * lane - j;
                            rem = remaining - above;
                            eqt = c[j];
                        }
                        above += c[j];
                    }
                    const int src = __ffs(__ballot_sync(kFull, dg >= 0)) - 1;
                    pivot_digit = __shfl_sync(kFull, dg, src);
                    remaining = __shfl_sync(kFull, rem, src);
                    eq_total = __shfl_sync(kFull, eqt, src);
                }
                thr |= (uint32_t) pivot_digit << sh;
                top = sh;
                if (first && top > 0 && eq_total <= kMaxCand) {
                    // few keys share the pivot digit: list them and rank them directly instead of more radix passes
                    // (one counter bump per warp: a per-key atomicAdd serialises the whole CTA on one shared-memory word)
                    const uint32_t dsel = dmask << sh, pdsel = (uint32_t) pivot_digit << sh;
                    int mine = 0;
#pragma unroll
                    for (int hh = 0; hh < kHPW; ++hh)
                        if (kExactH || wid + 8 * hh < H) {
#pragma unroll
                            for (int i = 0; i < kPerLane; ++i) mine += (key[hh][i] & dsel) == pdsel ? 1 : 0;
                        }
                    const int incl = warp_scan_incl_i(mine, lane);
                    int pos = 0;
                    if (lane == 31 && incl > 0) pos = atomicAdd(&scratch[11], incl);
                    pos = __shfl_sync(kFull, pos, 31) + incl - mine;
                    if (mine > 0) {                              // (a vote per key slot instead of the count + scan was measured slower: +6 us)
#pragma unroll
                        for (int hh = 0; hh < kHPW; ++hh)
                            if (kExactH || wid + 8 * hh < H) {
#pragma unroll
                                for (int i = 0; i < kPerLane; ++i) {
                                    const uint32_t u = key[hh][i];
                                    if ((u & dsel) == pdsel) cand[pos++] = u;
                                }
                            }
                    }
                    __syncthreads();
                    const int nc = eq_total;
                    for (int ci = tid; ci < nc; ci += kTopkThreads) {
                        const uint32_t mine_k = cand[ci];
                        int gt = 0, eq = 0;
                        for (int j = 0; j < nc; ++j) { const uint32_t o = cand[j]; gt += o > mine_k ? 1 : 0; eq += o == mine_k ? 1 : 0; }
                        if (gt < remaining && remaining <= gt + eq) {       // every thread holding the threshold value writes the same triple
                            scratch[12] = (int) mine_k;
                            scratch[13] = remaining - gt;
                            scratch[14] = eq;
                        }
                    }
                    __syncthreads();
                    thr = (uint32_t) scratch[12];
                    remaining = scratch[13];
                    eq_total = scratch[14];
                    top = 0;
                }
                first = false;
                if (top > 0) __syncthreads();                    // scratch / hist are reused by the next pass
            }
        }
    }
    // ---- alive bits: key > thr, plus the first `remaining` keys == thr in flat-index order ---------------------------
    const bool cut_ties = !all_alive && remaining < eq_total;
    uint32_t alive[kHPW];
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        uint32_t a = 0u;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) a |= ((all_alive || key[hh][i] >= thr) ? 1u : 0u) << i;
        alive[hh] = (kExactH || wid + 8 * hh < H) ? a : 0u;
    }
    if (cut_ties) {                                              // CTA-uniform branch
        int eqc[kHPW], lane_before[kHPW];
#pragma unroll
        for (int hh = 0; hh < kHPW; ++hh) {
            int c = 0;
#pragma unroll
            for (int i = 0; i < kPerLane; ++i) c += key[hh][i] == thr ? 1 : 0;
            eqc[hh] = c;
            const int incl = warp_scan_incl_i(c, lane);
            lane_before[hh] = incl - c;
            if (lane == 31) head_eq[wid + 8 * hh] = incl;
        }
        __syncthreads();
        // equal keys in the heads before head l (exclusive scan of the per-head counts; Hmax <= 32)
        const int he = lane < H ? head_eq[lane] : 0;
        const int heads_before = warp_scan_incl_i(he, lane) - he;
#pragma unroll
        for (int hh = 0; hh < kHPW; ++hh) {
            const int h = wid + 8 * hh;
            const int before = lane_before[hh] + __shfl_sync(kFull, heads_before, h & 31);
            int take = remaining - before;                       // how many of my equal keys (in index order) stay alive
            if (take < eqc[hh]) {
                uint32_t a = alive[hh];
#pragma unroll
                for (int i = 0; i < kPerLane; ++i)
                    if (key[hh][i] == thr) {
                        if (take <= 0) a &= ~(1u << i);
                        --take;
                    }
                alive[hh] = a;
            }
        }
    }
    // ---- words: kLanesPerWord lanes share one 32-pixel word ---------------------------------------------------------
    uint32_t* out_row = mask_bits + (int64_t) blockIdx.x * (G >> 5);
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
        const int h = wid + 8 * hh;
        uint32_t word = kPerLane >= 32 ? alive[hh] : (alive[hh] << (kPerLane * (lane % kLanesPerWord)));
#pragma unroll
        for (int o = 1; o < kLanesPerWord; o <<= 1) word |= __shfl_xor_sync(kFull, word, o);
        if ((lane % kLanesPerWord) == 0 && (kExactH || h < H)) {
            const int w = h * (P >> 5) + lane / kLanesPerWord;
            out_row[w] = word;
            if (crow_counts != nullptr) sbits[w] = word;
        }
    }
    if (crow_counts != nullptr) {
        // a8 pass 1 fused: crow[n, t+1] = entries of this row (sea_crow_scan turns the counts into offsets)
        __syncthreads();
        int cnt = 0;
        const float sc = __fdiv_rn((float) (t + 1), (float) P);
        for (int w = tid; w < (G >> 5); w += kTopkThreads) cnt += word_width_sum(sbits[w], w, P, sc, k_clamp);
        cnt = warp_sum_i(cnt);
        if (lane == 0) scratch[wid] = cnt;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
#pragma unroll
            for (int i = 0; i < kTopkThreads / 32; ++i) tot += scratch[i];
            crow_counts[(int64_t) n * (Tn + 1) + t + 1] = tot;
            if (t == 0) crow_counts[(int64_t) n * (Tn + 1)] = 0;
        }
    }
}

}  // namespace sea

using namespace sea;

static int tail_topk_impl(const float* y3, const float* bias, const float* ln_w, const float* ln_b,
                          const float* k_per_row, float* probs, uint32_t* mask_bits, int32_t* crow_counts, int k_clamp,
                          int N, int H, int T, int W, int P, void* stream) {
    SEA_CHECK_ARG(y3 && bias && ln_w && ln_b && (probs || mask_bits), "sea_predictor_tail_topk_fwd: null pointer");
    SEA_CHECK_ARG(mask_bits == nullptr || k_per_row != nullptr, "sea_predictor_tail_topk_fwd: k_per_row is required for the top-k");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && W > 0 && P > 0, "sea_predictor_tail_topk_fwd: bad shape");
    SEA_CHECK_ARG(crow_counts == nullptr || (mask_bits != nullptr && k_clamp > 0), "sea_predictor_tail_topk_fwd: crow_counts needs mask_bits and k");
    if (P % 32 != 0 || P % W != 0 || P > 1024) {
        set_error("sea_predictor_tail_topk_fwd: P=%d must be a multiple of 32 and of W=%d, <= 1024", P, W);
        return SEA_ERR_UNSUPPORTED;
    }
    const int G = H * P;
    cudaStream_t s = (cudaStream_t) stream;
    const unsigned grid = (unsigned) ((int64_t) N * T);
    static const bool no_reg = getenv("SEA_TAIL_SMEM") != nullptr;        // development switch for A/B timing
    const int hp_slots = (H + 7) / 8;
    if (!no_reg && hp_slots * (P / 32) <= 32 && P <= 1024 && H <= 64) {
        const size_t smem_r = (size_t) H * (W + 3) * 4 + (size_t) (G >> 5) * 4 + 16;
        SEA_CHECK_ARG(smem_r <= 72 * 1024, "sea_predictor_tail_topk: row image too large for the fused mask expansion");
        bool launched = true;
#define SEA_TAILR_L(PL, HP, UP, EX)                                                                                        \
        {                                                                                                                  \
            auto kern = tail_topk_reg_kernel<PL, HP, UP, EX>;                                                              \
            SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_r), "smem attr"); \
            SEA_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(kTopkThreads), smem_r, s, y3, bias, ln_w, ln_b, k_per_row, probs, mask_bits, crow_counts, k_clamp, N, T, W, H), \
                         "tail_topk_reg_kernel launch"); \
        }
#define SEA_TAILR(PL, HP)                                                                                                  \
        {                                                                                                                  \
            if (P / W == 4) SEA_TAILR_L(PL, HP, 4, true) else SEA_TAILR_L(PL, HP, 0, true)                                 \
        }
#define SEA_TAILR_PART(PL, HP)            /* H < 8 * HP: some head slots are empty */                                      \
        {                                                                                                                  \
            if (P / W == 4) SEA_TAILR_L(PL, HP, 4, false) else SEA_TAILR_L(PL, HP, 0, false)                               \
        }
        const int pl = P / 32, hp = hp_slots;
        static const bool minb3 = getenv("SEA_TAIL_MINB3") != nullptr;       // development switch for A/B timing
        if (H % 8 == 0) {
            // the north-star instantiation runs 4 CTAs per SM (64 registers, 80 bytes of spills): 101 us against 106 us with 3 at N=1, T=4096
            if (pl == 8 && hp == 4 && P / W == 4 && !minb3) {
                auto kern = tail_topk_reg_kernel<8, 4, 4, true, 4>;
                SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_r), "smem attr");
                SEA_CUDA_TRY(launch_pdl(kern, dim3(grid), dim3(kTopkThreads), smem_r, s, y3, bias, ln_w, ln_b, k_per_row, probs, mask_bits, crow_counts, k_clamp, N, T, W, H),
                             "tail_topk_reg_kernel launch");
            } else if (pl == 8 && hp == 4) SEA_TAILR(8, 4)
            else if (pl == 8 && hp == 2) SEA_TAILR(8, 2)
            else if (pl == 8 && hp == 1) SEA_TAILR(8, 1)
            else if (pl == 4 && hp == 4) SEA_TAILR(4, 4)
            else if (pl == 4 && hp == 2) SEA_TAILR(4, 2)
            else if (pl == 2 && hp == 1) SEA_TAILR(2, 1)
            else if (pl == 16 && hp == 2) SEA_TAILR(16, 2)
            else if (pl == 16 && hp == 1) SEA_TAILR(16, 1)
            else if (pl == 2 && hp == 4) SEA_TAILR(2, 4)
            else if (pl == 2 && hp == 2) SEA_TAILR(2, 2)
            else if (pl == 4 && hp == 1) SEA_TAILR(4, 1)
            else launched = false;
        } else {
            if (pl == 8 && hp == 2) SEA_TAILR_PART(8, 2)          // e.g. OPT-125m: H = 12, P = 256
            else if (pl == 4 && hp == 2) SEA_TAILR_PART(4, 2)
            else if (pl == 2 && hp == 2) SEA_TAILR_PART(2, 2)
            else if (pl == 8 && hp == 1) SEA_TAILR_PART(8, 1)
            else if (pl == 2 && hp == 1) SEA_TAILR_PART(2, 1)
            else launched = false;
        }
#undef SEA_TAILR_PART
#undef SEA_TAILR
#undef SEA_TAILR_L
        if (launched) {
            SEA_CHECK_LAUNCH("tail_topk_reg_kernel");
            return SEA_OK;
        }
    }
    const size_t smem = ((size_t) G + (G >> 5)) * 4 + (size_t) H * (W + 2) * 4 + 16;
    SEA_CHECK_ARG(smem <= 220 * 1024, "sea_predictor_tail_topk_fwd: H*P=%d keys do not fit shared memory", G);
#define SEA_TAIL_LAUNCH(PL, UP)                                                                                       \
    {                                                                                                                \
        auto kern = tail_topk_kernel<PL, UP>;                                                                        \
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr"); \
        kern<<<grid, kTopkThreads, smem, s>>>(y3, bias, ln_w, ln_b, k_per_row, probs, mask_bits, crow_counts, k_clamp, N, H, T, W, P);      \
    }
#define SEA_TAIL_CASE(PL)                                                                                            \
    case PL: {                                                                                                       \
        if (P / W == 4) SEA_TAIL_LAUNCH(PL, 4) else SEA_TAIL_LAUNCH(PL, 0)                                           \
        break;                                                                                                       \
    }
    switch (P / 32) {
        SEA_TAIL_CASE(1) SEA_TAIL_CASE(2) SEA_TAIL_CASE(4) SEA_TAIL_CASE(8) SEA_TAIL_CASE(16) SEA_TAIL_CASE(32)
        default:
            set_error("sea_predictor_tail_topk_fwd: P=%d unsupported (32,64,128,256,512,1024)", P);
            return SEA_ERR_UNSUPPORTED;
    }
#undef SEA_TAIL_CASE
#undef SEA_TAIL_LAUNCH
    SEA_CHECK_LAUNCH("tail_topk_kernel");
    return SEA_OK;
}

extern "C" {

int sea_predictor_tail_topk_fwd(const float* y3, const float* bias, const float* ln_w, const float* ln_b,
                                const float* k_per_row, float* probs, uint32_t* mask_bits, int32_t* crow_counts, int k_clamp,
                                int N, int H, int T, int W, int P, void* stream) {
    return tail_topk_impl(y3, bias, ln_w, ln_b, k_per_row, probs, mask_bits, crow_counts, k_clamp, N, H, T, W, P, stream);
}

}  // extern "C"
