// a9-a14 fused: flat_csr_masked_bmm -> flat_csr_softmax -> (* sigmoid(s0)) -> flat_csr_sdbmm ->
// mix with the causal running mean -> [N, T, H*D]   (reference attention.py:1151-1173, 1237-1244, 1279-1282).
//
// CTA = one query row (n, t); its H head segments (entries are head-major, a8 emits them that way) are
// located by H+1 parallel binary searches, then dealt to the CTA's warps.  Per segment: lane-per-entry
// Q.K (K rows are 128-bit gathered, q is a shared-memory broadcast), warp-shuffle online softmax over
// 32-entry chunks, and a lane-per-channel-pair P.V accumulation with coalesced V-row reads.
// Nothing but the output row (and optionally the probabilities) is written: scores never touch HBM.
#include "common.cuh"

namespace sea {

constexpr int kAttnWarps = 8;
constexpr int kMaxPairs = 4;   // channel pairs per lane -> D <= 256

template <typename T>
__device__ __forceinline__ float dot_q(const float* __restrict__ qs, const T* __restrict__ krow, int D);

template <>
__device__ __forceinline__ float dot_q<float>(const float* __restrict__ qs, const float* __restrict__ krow, int D) {
    float acc = 0.f;
    const float4* k4 = reinterpret_cast<const float4*>(krow);
    const float4* q4 = reinterpret_cast<const float4*>(qs);
#pragma unroll 4
    for (int c = 0; c < (D >> 2); ++c) {
        const float4 kv = __ldg(k4 + c);
        const float4 qv = q4[c];
        acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
    }
    return acc;
}

template <typename T16>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi);
template <>
__device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
template <>
__device__ __forceinline__ void unpack2<__half>(uint32_t w, float& lo, float& hi) {
    const __half2 h2 = *reinterpret_cast<const __half2*>(&w);
    lo = __low2float(h2);
    hi = __high2float(h2);
}

template <typename T16>
__device__ __forceinline__ float dot_q16(const float* __restrict__ qs, const T16* __restrict__ krow, int D) {
    float acc = 0.f;
    const uint4* k4 = reinterpret_cast<const uint4*>(krow);
    const float4* q4 = reinterpret_cast<const float4*>(qs);
#pragma unroll 2
    for (int c = 0; c < (D >> 3); ++c) {
        const uint4 kv = __ldg(k4 + c);
        const float4 qa = q4[2 * c], qb = q4[2 * c + 1];
        const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
        const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float lo, hi;
            unpack2<T16>(w[i], lo, hi);
            acc = fmaf(qv[2 * i], lo, acc);
            acc = fmaf(qv[2 * i + 1], hi, acc);
        }
    }
    return acc;
}
template <>
__device__ __forceinline__ float dot_q<__nv_bfloat16>(const float* __restrict__ qs, const __nv_bfloat16* __restrict__ krow, int D) {
    return dot_q16<__nv_bfloat16>(qs, krow, D);
}
template <>
__device__ __forceinline__ float dot_q<__half>(const float* __restrict__ qs, const __half* __restrict__ krow, int D) {
    return dot_q16<__half>(qs, krow, D);
}

template <typename T>
__device__ __forceinline__ float2 ld_pair(const T* p);
template <>
__device__ __forceinline__ float2 ld_pair<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <>
__device__ __forceinline__ float2 ld_pair<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 ld_pair<__half>(const __half* p) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <typename T>
__device__ __forceinline__ void st_pair(T* p, float a, float b);
template <>
__device__ __forceinline__ void st_pair<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <>
__device__ __forceinline__ void st_pair<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}
template <>
__device__ __forceinline__ void st_pair<__half>(__half* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T, typename IdxT>
__global__ void __launch_bounds__(kAttnWarps * 32)
sparse_attention_kernel(const IdxT* __restrict__ crow, const IdxT* __restrict__ col, int64_t Z,
                        const T* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                        const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                        const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                        const float* __restrict__ scales, const T* __restrict__ cumavg, int use_scaler,
                        T* __restrict__ out, float* __restrict__ probs_values,
                        int N, int H, int T_DST, int T_SRC, int D) {
    extern __shared__ __align__(16) float smem[];
    float* qs = smem;                                         // [H][D] fp32
    int64_t* hp = reinterpret_cast<int64_t*>(qs + H * D);      // [H+1] absolute entry offsets
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.x / T_DST, t = blockIdx.x % T_DST;
    const int tq = T_SRC - T_DST + t;                          // absolute position of the query row
    const IdxT* colr = col + (int64_t) n * Z;
    const int64_t rs = (int64_t) crow[(int64_t) n * (T_DST + 1) + t];
    const int64_t re = (int64_t) crow[(int64_t) n * (T_DST + 1) + t + 1];
    for (int idx = tid; idx < H * D; idx += kAttnWarps * 32) {
        const int h = idx / D, c = idx % D;
        qs[idx] = to_f32(q[(int64_t) n * q_sn + (int64_t) h * q_sh + (int64_t) t * q_st + c]);
    }
    for (int h = tid; h <= H; h += kAttnWarps * 32) {
        // first entry whose column id >= h*T_SRC
        const int64_t key = (int64_t) h * T_SRC;
        int64_t lo = rs, hi = re;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t) colr[mid] < key) lo = mid + 1; else hi = mid;
        }
        hp[h] = lo;
    }
    __syncthreads();
    constexpr float kLog2e = 1.4426950408889634f;
    for (int h = wid; h < H; h += kAttnWarps) {
        const int64_t s0 = hp[h], s1 = hp[h + 1];
        const float* qh = qs + h * D;
        const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
        const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;
        float m_run = -INFINITY, l_run = 0.f;
        float acc[2 * kMaxPairs];
#pragma unroll
        for (int i = 0; i < 2 * kMaxPairs; ++i) acc[i] = 0.f;
        for (int64_t base = s0; base < s1; base += 32) {
            const int64_t z = base + lane;
            const bool valid = z < s1;
            int j = 0;
            float sc = -INFINITY;
            if (valid) {
                j = (int) ((int64_t) colr[z] - (int64_t) h * T_SRC);
                sc = dot_q<T>(qh, kb + (int64_t) j * k_st, D);
                if (probs_values != nullptr) probs_values[(int64_t) n * Z + z] = sc;
            }
            const float m_new = fmaxf(m_run, warp_max(sc));
            const float alpha = exp2f((m_run - m_new) * kLog2e);   // m_run = -inf -> 0
            const float p = valid ? exp2f((sc - m_new) * kLog2e) : 0.f;
            l_run = l_run * alpha + warp_sum(p);
#pragma unroll
            for (int i = 0; i < 2 * kMaxPairs; ++i) acc[i] *= alpha;
            m_run = m_new;
            const int nvalid = (int) min((int64_t) 32, s1 - base);
            for (int e = 0; e < nvalid; ++e) {
                const float pe = __shfl_sync(kFull, p, e);
                const int je = __shfl_sync(kFull, j, e);
                const T* vrow = vb + (int64_t) je * v_st;
#pragma unroll
                for (int i = 0; i < kMaxPairs; ++i) {
                    const int d = 2 * lane + 64 * i;
                    if (d < D) {
                        const float2 vv = ld_pair<T>(vrow + d);
                        acc[2 * i] = fmaf(pe, vv.x, acc[2 * i]);
                        acc[2 * i + 1] = fmaf(pe, vv.y, acc[2 * i + 1]);
                    }
                }
            }
        }
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
        const float psc = use_scaler ? sigmoidf_(sp[0]) : 1.0f;
        const float a = sigmoidf_(sp[1]);
        T* orow = out + ((int64_t) n * T_DST + t) * ((int64_t) H * D) + (int64_t) h * D;
        const T* arow = cumavg ? cumavg + (((int64_t) n * H + h) * T_DST + t) * D : nullptr;
#pragma unroll
        for (int i = 0; i < kMaxPairs; ++i) {
            const int d = 2 * lane + 64 * i;
            if (d < D) {
                float c0 = acc[2 * i] * inv * psc, c1 = acc[2 * i + 1] * inv * psc;
                if (arow) {
                    const float2 av = ld_pair<T>(arow + d);
                    c0 = c0 * a + (1.0f - a) * av.x;
                    c1 = c1 * a + (1.0f - a) * av.y;
                }
                st_pair<T>(orow + d, c0, c1);
            }
        }
        if (probs_values != nullptr) {
            for (int64_t z = s0 + lane; z < s1; z += 32) {
                float* pv = probs_values + (int64_t) n * Z + z;
                *pv = exp2f((*pv - m_run) * kLog2e) * inv * psc;
            }
        }
    }
    (void) tq;
}


// ------------------------------------------------------------------------------------------------
// v2 (16-bit activations, D in {32, 64, 128}, head_ptr available): warp per (row, head) segment, no shared memory.
// A K / V row (D x 2 bytes) is covered by LPR = D/8 lanes with ONE 16-byte load each, so a warp-wide load instruction
// fetches EPI = 32/LPR whole rows; the 32/EPI loads of a 32-entry chunk -- for K and for V -- are all issued before any
// is consumed (16 independent 128-bit loads in flight per lane), which is what hides the L2 gather latency.
// Scores are reduced across the LPR lanes with shuffles; every lane of a group then already holds the probability of
// the entries whose V slices it accumulates, so P.V needs no further communication until the final cross-group sum.
// ------------------------------------------------------------------------------------------------
template <typename T16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    unpack2<T16>(u.x, f[0], f[1]);
    unpack2<T16>(u.y, f[2], f[3]);
    unpack2<T16>(u.z, f[4], f[5]);
    unpack2<T16>(u.w, f[6], f[7]);
}
template <typename T16>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 p = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T16, typename IdxT, int D>
__global__ void __launch_bounds__(kAttnWarps * 32, 3)
sparse_attention_v2_kernel(const IdxT* __restrict__ col, int64_t Z, const int32_t* __restrict__ head_ptr,
                           const T16* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                           const T16* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                           const T16* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                           const float* __restrict__ scales, const T16* __restrict__ cumavg, int use_scaler,
                           T16* __restrict__ out, float* __restrict__ probs_values,
                           int N, int H, int T_DST, int T_SRC) {
    constexpr int LPR = D / 8;          // lanes per K/V row
    constexpr int EPI = 32 / LPR;       // entries per warp-wide load
    constexpr int NI = 32 / EPI;        // loads per 32-entry chunk (= LPR)
    const int lane = threadIdx.x & 31;
    // task order (n, h, t) with t fastest: the warps of a CTA work on CONSECUTIVE query rows of ONE head, whose pixel
    // runs overlap, so part of the K/V gathers hits the SM's L1 instead of going to L2 again
    const int64_t task = (int64_t) blockIdx.x * kAttnWarps + (threadIdx.x >> 5);
    if (task >= (int64_t) N * T_DST * H) return;
    const int t = (int) (task % T_DST);
    const int h = (int) ((task / T_DST) % H);
    const int n = (int) (task / ((int64_t) T_DST * H));
    const int64_t row = (int64_t) n * T_DST + t;
    const int sub = lane % LPR;         // which 8-channel slice of the row this lane owns
    const int grp = lane / LPR;         // which entry of a load instruction this lane serves
    const int32_t* hp = head_ptr + row * (H + 1) + h;
    const int s0 = hp[0], s1 = hp[1];
    const IdxT* colr = col + (int64_t) n * Z;
    // 16-byte-vector views; row strides in vectors fit 32 bits (strides are multiples of 8 elements)
    const uint4* kb = reinterpret_cast<const uint4*>(k + (int64_t) n * k_sn + (int64_t) h * k_sh) + sub;
    const uint4* vb = reinterpret_cast<const uint4*>(v + (int64_t) n * v_sn + (int64_t) h * v_sh) + sub;
    const uint32_t k_sv = (uint32_t) (k_st >> 3), v_sv = (uint32_t) (v_st >> 3);
    constexpr float kLog2e = 1.4426950408889634f;
    float qf[8];
    {
        const uint4 qu = __ldg(reinterpret_cast<const uint4*>(q + (int64_t) n * q_sn + (int64_t) h * q_sh + (int64_t) t * q_st) + sub);
        unpack8<T16>(qu, qf);
#pragma unroll
        for (int c = 0; c < 8; ++c) qf[c] *= kLog2e;      // scores live in the log2 domain: softmax = 2^(s - m) / sum
    }
    float m_run = -INFINITY, l_run = 0.f;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const int hbase = h * T_SRC;
    for (int base = s0; base < s1; base += 32) {
        const int cnt = min(32, s1 - base);
        int jmine = 0;
        if (lane < cnt) jmine = (int) colr[base + lane] - hbase;
        uint4 ku[NI], vu[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int e = i * EPI + grp;
            const uint32_t j = (uint32_t) __shfl_sync(kFull, jmine, e);
            ku[i] = make_uint4(0, 0, 0, 0);
            vu[i] = make_uint4(0, 0, 0, 0);
            if (e < cnt) {
                ku[i] = __ldg(kb + j * k_sv);
                vu[i] = __ldg(vb + j * v_sv);
            }
        }
        float sc[NI];
        float cmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            float kf[8];
            unpack8<T16>(ku[i], kf);
            float d = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) d = fmaf(qf[c], kf[c], d);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
            sc[i] = (i * EPI + grp < cnt) ? d : -INFINITY;
            cmax = fmaxf(cmax, sc[i]);
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) cmax = fmaxf(cmax, __shfl_xor_sync(kFull, cmax, o));
        if (probs_values != nullptr && sub == 0) {
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (i * EPI + grp < cnt) probs_values[(int64_t) n * Z + base + i * EPI + grp] = sc[i];
        }
        const float m_new = fmaxf(m_run, cmax);
        const float alpha = ex2_approx(m_run - m_new);
        float psum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= alpha;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const float p = ex2_approx(sc[i] - m_new);            // -inf -> 0 for the padding entries
            psum += p;
            float vf[8];
            unpack8<T16>(vu[i], vf);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(p, vf[c], acc[c]);
        }
        // every lane of a group holds the same p's: sum over groups only
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) psum += __shfl_xor_sync(kFull, psum, o);
        l_run = l_run * alpha + psum;
        m_run = m_new;
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(kFull, acc[c], o);
    }
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
    const float psc = use_scaler ? sigmoidf_(sp[0]) : 1.0f;
    if (grp == 0) {
        const float a = sigmoidf_(sp[1]);
        float o8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o8[c] = acc[c] * inv * psc;
        if (cumavg != nullptr) {
            const uint4 au = __ldg(reinterpret_cast<const uint4*>(cumavg + (((int64_t) n * H + h) * T_DST + t) * D) + sub);
            float af[8];
            unpack8<T16>(au, af);
#pragma unroll
            for (int c = 0; c < 8; ++c) o8[c] = o8[c] * a + (1.0f - a) * af[c];
        }
        uint4 ou;
        ou.x = pack2<T16>(o8[0], o8[1]); ou.y = pack2<T16>(o8[2], o8[3]);
        ou.z = pack2<T16>(o8[4], o8[5]); ou.w = pack2<T16>(o8[6], o8[7]);
        *(reinterpret_cast<uint4*>(out + ((int64_t) n * T_DST + t) * ((int64_t) H * D) + (int64_t) h * D) + sub) = ou;
    }
    if (probs_values != nullptr) {
        for (int z = s0 + lane; z < s1; z += 32) {
            float* pv = probs_values + (int64_t) n * Z + z;
            *pv = ex2_approx(*pv - m_run) * inv * psc;
        }
    }
}

}  // namespace sea

using namespace sea;

extern "C" {

int sea_sparse_attention_fwd(const void* crow, const void* col, int idx64, int64_t Z,
                             const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                             const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                             const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                             const float* scales, const void* cumavg, int use_scaler, int dtype, void* out,
                             float* probs_values, const int32_t* head_ptr, int N, int H, int T_DST, int T_SRC, int D, void* stream) {
    SEA_CHECK_ARG(crow && (col || Z == 0) && q && k && v && scales && out, "sea_sparse_attention_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T_DST > 0 && T_SRC >= T_DST && D > 0, "sea_sparse_attention_fwd: bad shape");
    SEA_CHECK_ARG(D % 8 == 0 && D <= 64 * kMaxPairs, "sea_sparse_attention_fwd: head dim %d unsupported (multiple of 8, <= %d)", D, 64 * kMaxPairs);
    SEA_CHECK_ARG((k_st % 8) == 0 && (k_sh % 8) == 0 && (k_sn % 8) == 0 && (v_st % 2) == 0 && (v_sh % 2) == 0 && (v_sn % 2) == 0,
                  "sea_sparse_attention_fwd: k/v strides must keep rows 16-byte aligned");
    SEA_CHECK_ARG((((uintptr_t) k) & 15) == 0 && (((uintptr_t) v) & 15) == 0 && (((uintptr_t) out) & 15) == 0, "sea_sparse_attention_fwd: k/v/out must be 16-byte aligned");
    if (head_ptr != nullptr && dtype != SEA_DTYPE_F32 && (D == 32 || D == 64 || D == 128) &&
        (q_st % 8) == 0 && (q_sh % 8) == 0 && (q_sn % 8) == 0 && (((uintptr_t) q) & 15) == 0 &&
        (cumavg == nullptr || (((uintptr_t) cumavg) & 15) == 0)) {
        const int64_t tasks = (int64_t) N * T_DST * H;
        const unsigned grid = (unsigned) ((tasks + kAttnWarps - 1) / kAttnWarps);
        cudaStream_t s = (cudaStream_t) stream;
#define SEA_ATTN_V2(TT, II, DD)                                                                                              \
        sparse_attention_v2_kernel<TT, II, DD><<<grid, kAttnWarps * 32, 0, s>>>(                                             \
            (const II*) col, Z, head_ptr, (const TT*) q, q_sn, q_sh, q_st, (const TT*) k, k_sn, k_sh, k_st, (const TT*) v, v_sn, \
            v_sh, v_st, scales, (const TT*) cumavg, use_scaler, (TT*) out, probs_values, N, H, T_DST, T_SRC)
#define SEA_ATTN_V2_D(TT, II)                                                  \
        do {                                                                   \
            if (D == 32) SEA_ATTN_V2(TT, II, 32);                              \
            else if (D == 64) SEA_ATTN_V2(TT, II, 64);                         \
            else SEA_ATTN_V2(TT, II, 128);                                     \
        } while (0)
        if (dtype == SEA_DTYPE_BF16) { if (idx64) SEA_ATTN_V2_D(__nv_bfloat16, int64_t); else SEA_ATTN_V2_D(__nv_bfloat16, int32_t); }
        else { if (idx64) SEA_ATTN_V2_D(__half, int64_t); else SEA_ATTN_V2_D(__half, int32_t); }
#undef SEA_ATTN_V2_D
#undef SEA_ATTN_V2
        SEA_CHECK_LAUNCH("sparse_attention_v2_kernel");
        return SEA_OK;
    }
    const size_t smem = (size_t) H * D * sizeof(float) + (size_t) (H + 1) * sizeof(int64_t) + 16;
    SEA_CHECK_ARG(smem <= 227 * 1024, "sea_sparse_attention_fwd: H*D too large");
    SEA_DISPATCH_DTYPE(dtype, T_, SEA_DISPATCH_IDX(idx64, I, {
        auto kern = sparse_attention_kernel<T_, I>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        kern<<<(unsigned) ((int64_t) N * T_DST), kAttnWarps * 32, smem, (cudaStream_t) stream>>>(
            (const I*) crow, (const I*) col, Z, (const T_*) q, q_sn, q_sh, q_st, (const T_*) k, k_sn, k_sh, k_st,
            (const T_*) v, v_sn, v_sh, v_st, scales, (const T_*) cumavg, use_scaler, (T_*) out, probs_values,
            N, H, T_DST, T_SRC, D);
        SEA_CHECK_LAUNCH("sparse_attention_kernel");
    }));
    return SEA_OK;
}

}  // extern "C"
