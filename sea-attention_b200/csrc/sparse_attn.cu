// a9-a14 fused: flat_csr_masked_bmm -> flat_csr_softmax -> (* sigmoid(s0)) -> flat_csr_sdbmm ->
// mix with the causal running mean -> [N, T, H*D]   (reference attention.py:1151-1173, 1237-1244, 1279-1282).
//
// CTA = one query row (n, t); its H head segments (entries are head-major, a8 emits them that way) are
// located by H+1 parallel binary searches, then dealt to the CTA's warps.  Per segment: lane-per-entry
// Q.K (K rows are 128-bit gathered, q is a shared-memory broadcast), warp-shuffle online softmax over
// 32-entry chunks, and a lane-per-channel-pair P.V accumulation with coalesced V-row reads.
// Nothing but the output row (and optionally the probabilities) is written: scores never touch HBM.
#include "common.cuh"
#include "csr_common.cuh"

#include <stdlib.h>

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
                        const float* __restrict__ scales, const T* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
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
        const T* arow = cumavg ? cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st : nullptr;
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
                           const float* __restrict__ scales, const T16* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
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
            const uint4 au = __ldg(reinterpret_cast<const uint4*>(cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st) + sub);
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


// ------------------------------------------------------------------------------------------------
// Attention straight from the top-k BIT MASK.  The flat-CSR column list is a pure function of the bit mask (a8), so when the
// caller does not ask for the CSR tensors the entries are enumerated on the fly, per (row, head), with the same pixel
// arithmetic (csr_common.cuh) -- the count / scan / fill kernels and the 4 bytes per entry of column ids leave the hot path.
// Entry order inside a (row, head) segment equals the CSR order (pixels ascending, tokens descending inside a pixel), so results
// match the CSR kernels up to fp32 summation order.  (A CUDA-core version of this kernel measured no faster than the CSR path
// and was removed; the tensor-core version below is the one in use.)
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// v4: the per-(row, head) segment on the TENSOR CORES.  A single query row is a degenerate GEMM, but the CUDA-core
// version of v2/v3 spends ~800 instructions per 32 entries on bf16 unpacking, FMAs and shuffle reductions, and is
// issue-bound.  Here the 32 gathered K rows and V rows of a chunk are copied global -> shared with 16-byte cp.async
// (no register staging), and
//   scores  = K_tile [32 x D] . q          as  m16n8k16 MMAs with q in column 0 of the B operand
//   out    += p [1 x 32] . V_tile [32 x D]  as  m16n8k16 MMAs with p in row 0 of the A operand (V via ldmatrix.trans)
// 15/16 of every MMA is padding -- irrelevant: the tensor pipe is idle otherwise, and the instruction count per chunk
// drops ~4x.  fp32 accumulation, fp32 online softmax in the log2 domain; q, K, V, p enter the MMAs as bf16/fp16.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t) __cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t) __cvta_generic_to_shared(p)));
}
template <typename T16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
struct AttnMmaCfg {
    static constexpr int kLd = D + 8;                       // padded row (elements): ldmatrix rows hit distinct banks
    // D <= 64: a K tile and a V tile per warp, both gathered at once (3 CTAs / SM).  Larger head dims: ONE tile buffer per warp, filled
    // with the K rows for the scores and then with the V rows for P.V -- with two tiles a d = 128 CTA needs 139 KB and the SM holds a
    // single CTA of 8 warps, latency-bound on its gathers (and at the 80-register cap of 3 CTAs / SM the 64 accumulator registers spill)
    static constexpr bool kShare = D > 64;
    static constexpr int kMinBlocks = kShare ? 2 : 3;          // (2: 128 registers, no spills with the 64 accumulators of d = 128; at 3 CTAs / SM the d = 80 / 96 instances spill ~400 B)
    static constexpr int kWarpElems = (kShare ? 1 : 2) * 32 * kLd;
    static constexpr int kSmemBytes = kAttnWarps * kWarpElems * 2;
};

// One 32-entry chunk: gathers K/V rows `jmine` (lane e holds the source token of entry e; cnt valid entries) and folds
// them into the running (m, l, acc) online-softmax state of this warp's (row, head).
template <typename T16, int D>
__device__ __forceinline__ void attn_chunk_mma(T16* __restrict__ Ks, T16* __restrict__ Vs, const uint4* __restrict__ kb, const uint4* __restrict__ vb,
                                               uint32_t k_sv, uint32_t v_sv, int jmine, int cnt, const uint32_t (&qb)[D / 16][2],
                                               float& m_run, float& l_run, float (&acc)[D / 8][4], int lane) {
    constexpr int LPR = D / 8, kLd = AttnMmaCfg<D>::kLd;
    constexpr bool kShare = AttnMmaCfg<D>::kShare;
    constexpr float kLog2e = 1.4426950408889634f;
    // gather of the 32 rows `jmine` of one operand (base pointer already offset by the lane's piece, row stride in 16-byte units)
    auto gather = [&](T16* dst, const uint4* base, uint32_t sv) {
        if constexpr (32 % LPR == 0) {
            constexpr int EPI = 32 / LPR, NI = LPR;
            const int sub = lane % LPR, grp = lane / LPR;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int e = i * EPI + grp;
                const uint32_t j = (uint32_t) __shfl_sync(kFull, jmine, e);
                const bool ok = e < cnt;
                cp_async16_zfill(dst + e * kLd + sub * 8, base + (ok ? j * sv : 0u), ok ? 16 : 0);
            }
        } else {
            // head dims whose row is not a power-of-two number of 16-byte pieces (D = 80: 10, D = 96: 12): piece idx -> (entry, piece)
            const uint4* base0 = base - (lane % LPR);
#pragma unroll
            for (int i = 0; i < LPR; ++i) {
                const int idx = i * 32 + lane, e = idx / LPR, sb = idx - e * LPR;
                const uint32_t j = (uint32_t) __shfl_sync(kFull, jmine, e);
                const bool ok = e < cnt;
                cp_async16_zfill(dst + e * kLd + sb * 8, base0 + sb + (ok ? j * sv : 0u), ok ? 16 : 0);
            }
        }
    };
    gather(Ks, kb, k_sv);
    if constexpr (!kShare) gather(Vs, vb, v_sv);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const int g = lane >> 2, tq = lane & 3;
    // scores: two m16 tiles of entries; only column 0 of the n8 tile is real
    float sc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sc[mt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t a[4];
            ldsm4(a, Ks + (mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kLd + ks * 16 + 8 * (lane >> 4));
            mma_16816<T16>(sc[mt], a, qb[ks][0], qb[ks][1]);
        }
    }
    // lanes with tq == 0 hold entries g, g+8, 16+g, 24+g in (sc[0][0], sc[0][2], sc[1][0], sc[1][2])
    float s4[4] = {sc[0][0] * kLog2e, sc[0][2] * kLog2e, sc[1][0] * kLog2e, sc[1][2] * kLog2e};
    const int e4[4] = {g, g + 8, 16 + g, 24 + g};
    float cmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (e4[i] >= cnt) s4[i] = -INFINITY;
        cmax = fmaxf(cmax, s4[i]);
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) cmax = fmaxf(cmax, __shfl_xor_sync(kFull, cmax, o));
    cmax = __shfl_sync(kFull, cmax, lane & ~3);              // take the tq == 0 lane's value (others hold padding columns)
    const float m_new = fmaxf(m_run, cmax);
    const float alpha = ex2_approx(m_run - m_new);
    float p4[4], psum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { p4[i] = ex2_approx(s4[i] - m_new); psum += p4[i]; }
    if (tq != 0) psum = 0.f;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) psum += __shfl_xor_sync(kFull, psum, o);
    psum = __shfl_sync(kFull, psum, lane & ~3);
    l_run = l_run * alpha + psum;
    m_run = m_new;
    if constexpr (kShare) {     // the K rows are consumed (the scores live in registers): the V rows take their place
        __syncwarp();
        gather(Vs, vb, v_sv);
    }
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) { acc[nt][0] *= alpha; acc[nt][1] *= alpha; }
    // p as the A operand (row 0 only): lane (g == 0, tq) needs entries 16ks+2tq, +1, 16ks+8+2tq, +1
    uint32_t pa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const float lo0 = __shfl_sync(kFull, p4[2 * ks], 8 * tq);          // entry 16ks + 2tq      (holder lane 4*(2tq), slot c0)
        const float lo1 = __shfl_sync(kFull, p4[2 * ks], 8 * tq + 4);      // entry 16ks + 2tq + 1
        const float hi0 = __shfl_sync(kFull, p4[2 * ks + 1], 8 * tq);      // entry 16ks + 8 + 2tq  (slot c2)
        const float hi1 = __shfl_sync(kFull, p4[2 * ks + 1], 8 * tq + 4);
        const bool row0 = g == 0;
        pa[ks][0] = row0 ? pack2<T16>(lo0, lo1) : 0u;
        pa[ks][1] = 0u;
        pa[ks][2] = row0 ? pack2<T16>(hi0, hi1) : 0u;
        pa[ks][3] = 0u;
    }
    if constexpr (kShare) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int np = 0; np < D / 16; ++np) {
            uint32_t b[4];
            ldsm4_t(b, Vs + (ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kLd + np * 16 + 8 * (lane >> 4));
            mma_16816<T16>(acc[2 * np], pa[ks], b[0], b[1]);
            mma_16816<T16>(acc[2 * np + 1], pa[ks], b[2], b[3]);
        }
    }
    __syncwarp();       // every lane is done reading the tiles before the next chunk overwrites them
}

template <typename T16, int D>
__global__ void __launch_bounds__(kAttnWarps * 32, AttnMmaCfg<D>::kMinBlocks)
sparse_attention_bits_mma_kernel(const uint32_t* __restrict__ mask_bits,
                                 const T16* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                 const T16* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                 const T16* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                 const float* __restrict__ scales, const T16* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
                                 T16* __restrict__ out, int N, int H, int T_DST, int T_SRC, int P, int k_clamp, int is_causal) {
    extern __shared__ __align__(16) uint8_t attn_smem[];
    constexpr int LPR = D / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T16* Ks = reinterpret_cast<T16*>(attn_smem) + warp * AttnMmaCfg<D>::kWarpElems;
    T16* Vs = AttnMmaCfg<D>::kShare ? Ks : Ks + 32 * AttnMmaCfg<D>::kLd;
    const int64_t task = (int64_t) blockIdx.x * kAttnWarps + warp;
    if (task >= (int64_t) N * T_DST * H) return;
    const int t = (int) (task % T_DST);
    const int h = (int) ((task / T_DST) % H);
    const int n = (int) (task / ((int64_t) T_DST * H));
    const int64_t row = (int64_t) n * T_DST + t;
    const int sub = lane % LPR;
    const int g = lane >> 2, tq = lane & 3;
    const uint4* kb = reinterpret_cast<const uint4*>(k + (int64_t) n * k_sn + (int64_t) h * k_sh) + sub;
    const uint4* vb = reinterpret_cast<const uint4*>(v + (int64_t) n * v_sn + (int64_t) h * v_sh) + sub;
    const uint32_t k_sv = (uint32_t) (k_st >> 3), v_sv = (uint32_t) (v_st >> 3);
    // q as the B operand: column 0 of every n8 tile -> only lanes with g == 0 carry data
    uint32_t qb[D / 16][2];
    {
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(q + (int64_t) n * q_sn + (int64_t) h * q_sh + (int64_t) t * q_st);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            qb[ks][0] = g == 0 ? __ldg(qw + ks * 8 + tq) : 0u;           // dims 16ks + 2tq, +1
            qb[ks][1] = g == 0 ? __ldg(qw + ks * 8 + 4 + tq) : 0u;       // dims 16ks + 8 + 2tq, +1
        }
    }
    float m_run = -INFINITY, l_run = 0.f;
    float acc[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;

    const int nw = P >> 5;
    const int L = is_causal ? (T_SRC - T_DST + t + 1) : T_SRC;
    const float s_scale = __fdiv_rn((float) L, (float) P);
    const uint32_t word = lane < nw ? mask_bits[row * ((int64_t) H * nw) + (int64_t) h * nw + lane] : 0u;
    const int pc = __popc(word);
    const int pc_incl = warp_scan_incl_i(pc, lane);
    const int n_alive = __shfl_sync(kFull, pc_incl, 31);
    for (int r0 = 0; r0 < n_alive; r0 += 32) {
        const int slot = r0 + lane;
        int wi = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int vv = __shfl_sync(kFull, pc_incl, wi + step - 1);
            if (vv <= slot) wi += step;
        }
        wi = min(wi, 31);
        const uint32_t wsel = __shfl_sync(kFull, word, wi);
        const int before = __shfl_sync(kFull, pc_incl - pc, wi);
        int wd = 0, ve_i = 0, span_i = 0;
        if (slot < n_alive) {
            const int bit = __fns(wsel, 0, slot - before + 1);
            float vs, ve;
            pixel_bounds(s_scale, (wi << 5) + bit, vs, ve);
            span_i = (int) __fsub_rn(ve, vs);
            ve_i = (int) ve;
            wd = min(span_i, k_clamp);
        }
        const int incl = warp_scan_incl_i(wd, lane);
        const int excl = incl - wd;
        const int total = __shfl_sync(kFull, incl, 31);
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int cnt = min(32, total - e0);
            const int e = e0 + lane;
            int pl = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int vv = __shfl_sync(kFull, incl, pl + step - 1);
                if (vv <= e) pl += step;
            }
            pl = min(pl, 31);
            const int p_excl = __shfl_sync(kFull, excl, pl);
            const int p_ve = __shfl_sync(kFull, ve_i, pl);
            const int p_wd = __shfl_sync(kFull, wd, pl);
            const int p_span = __shfl_sync(kFull, span_i, pl);
            int jmine = 0;
            if (lane < cnt) {
                const int i = e - p_excl;
                jmine = p_wd == p_span ? p_ve - 1 - i
                                       : p_ve - 1 - (int) __fmul_rn((float) i, __fdiv_rn((float) p_span, (float) p_wd));
            }
            attn_chunk_mma<T16, D>(Ks, Vs, kb, vb, k_sv, v_sv, jmine, cnt, qb, m_run, l_run, acc, lane);
        }
    }
    // row 0 of the accumulator tiles lives in lanes g == 0: dims 8nt + 2tq, +1
    if (g == 0) {
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
        const float psc = use_scaler ? sigmoidf_(sp[0]) : 1.0f;
        const float a = sigmoidf_(sp[1]);
        T16* orow = out + ((int64_t) n * T_DST + t) * ((int64_t) H * D) + (int64_t) h * D;
        const T16* arow = cumavg ? cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st : nullptr;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            const int dd = nt * 8 + 2 * tq;
            float c0 = acc[nt][0] * inv * psc, c1 = acc[nt][1] * inv * psc;
            if (arow) {
                float a0, a1;
                unpack2<T16>(__ldg(reinterpret_cast<const uint32_t*>(arow + dd)), a0, a1);
                c0 = c0 * a + (1.0f - a) * a0;
                c1 = c1 * a + (1.0f - a) * a1;
            }
            *reinterpret_cast<uint32_t*>(orow + dd) = pack2<T16>(c0, c1);
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
                             const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype, void* out,
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
            v_sh, v_st, scales, (const TT*) cumavg, avg_sh, avg_st, use_scaler, (TT*) out, probs_values, N, H, T_DST, T_SRC)
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
            (const T_*) v, v_sn, v_sh, v_st, scales, (const T_*) cumavg, avg_sh, avg_st, use_scaler, (T_*) out, probs_values,
            N, H, T_DST, T_SRC, D);
        SEA_CHECK_LAUNCH("sparse_attention_kernel");
    }));
    return SEA_OK;
}

}  // extern "C"

extern "C" int sea_sparse_attention_bits_fwd(const uint32_t* mask_bits,
                                             const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                             const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                             const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                             const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype, void* out,
                                             int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal, void* stream) {
    SEA_CHECK_ARG(mask_bits && q && k && v && scales && out, "sea_sparse_attention_bits_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T_DST > 0 && T_SRC >= T_DST && k_clamp > 0, "sea_sparse_attention_bits_fwd: bad shape");
    if (dtype == SEA_DTYPE_F32 || !(D == 32 || D == 64 || D == 80 || D == 96 || D == 128) || (P % 32) != 0 || P > 1024) {
        set_error("sea_sparse_attention_bits_fwd: unsupported (needs 16-bit activations, D in {32,64,80,96,128}, P %% 32 == 0, P <= 1024)");
        return SEA_ERR_UNSUPPORTED;
    }
    SEA_CHECK_ARG(((q_sn | q_sh | q_st | k_sn | k_sh | k_st | v_sn | v_sh | v_st) % 8) == 0 &&
                  ((((uintptr_t) q) | ((uintptr_t) k) | ((uintptr_t) v) | ((uintptr_t) out) | ((uintptr_t) cumavg)) & 15) == 0,
                  "sea_sparse_attention_bits_fwd: rows must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t) stream;
    const int64_t tasks = (int64_t) N * T_DST * H;
    const unsigned grid = (unsigned) ((tasks + kAttnWarps - 1) / kAttnWarps);
#define SEA_ATTN_BITS(TT, DD)                                                                                                   \
    do {                                                                                                                        \
        auto kern = sparse_attention_bits_mma_kernel<TT, DD>;                                                                   \
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnMmaCfg<DD>::kSmemBytes), "smem attr"); \
        kern<<<grid, kAttnWarps * 32, AttnMmaCfg<DD>::kSmemBytes, s>>>(mask_bits, (const TT*) q, q_sn, q_sh, q_st, (const TT*) k, k_sn,  \
            k_sh, k_st, (const TT*) v, v_sn, v_sh, v_st, scales, (const TT*) cumavg, avg_sh, avg_st, use_scaler, (TT*) out, N, H, T_DST, T_SRC, P, k_clamp, is_causal); \
    } while (0)
    if (dtype == SEA_DTYPE_BF16) {
        if (D == 32) SEA_ATTN_BITS(__nv_bfloat16, 32); else if (D == 64) SEA_ATTN_BITS(__nv_bfloat16, 64); else if (D == 80) SEA_ATTN_BITS(__nv_bfloat16, 80);
        else if (D == 96) SEA_ATTN_BITS(__nv_bfloat16, 96); else SEA_ATTN_BITS(__nv_bfloat16, 128);
    } else {
        if (D == 32) SEA_ATTN_BITS(__half, 32); else if (D == 64) SEA_ATTN_BITS(__half, 64); else if (D == 80) SEA_ATTN_BITS(__half, 80);
        else if (D == 96) SEA_ATTN_BITS(__half, 96); else SEA_ATTN_BITS(__half, 128);
    }
#undef SEA_ATTN_BITS
    SEA_CHECK_LAUNCH("sparse_attention_bits_mma_kernel");
    return SEA_OK;
}
