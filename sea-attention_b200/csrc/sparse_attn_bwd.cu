// Backward of the fused a8-a14 sparse attention (the reference has no sparse backward kernel -- its training
// path differentiates the dense masked attention of attention.py:1066-1133 with autograd; parity target = that gradient
// restricted to the mask, oracle/sea_oracle.py::sparse_attention_grads).
//
// Forward, per (n, h, query row r) with alive source tokens z (enumerated from the top-k bit mask exactly as the forward
// kernels do):   S_z = q_r . k_z     P = softmax_z(S)     c0 = sum_z P_z v_z     ctx = psc * c0,  psc = sigmoid(s0) (or 1)
//                out_r = a * ctx + (1 - a) * avg_r,  a = sigmoid(s1),  avg_r = mean_{j <= r} v_j   (attention.py:1151-1173, 1237-1244)
// Backward with g_z = dout_r . v_z and E = sum_z P_z g_z (= dout_r . c0):
//   d s1 = (psc * E - dout_r . avg_r) * a (1 - a)          d s0 = a * E * psc (1 - psc)
//   dS_z = a psc P_z (g_z - E)       dq_r = sum_z dS_z k_z       dk_z += dS_z q_r       dv_z += a psc P_z dout_r
//   dv_j += sum_{r >= j} (1 - a_r) dout_r / (r + 1)             (the running-mean branch; separate reverse-scan kernel)
// The top-k mask is piecewise constant, so no gradient flows into the predictor through it (as in the reference, whose
// predictor learns from its own distillation losses).
//
// Kernel: warp per (query row, head), two sweeps over the row's entries (32 per step, lane = entry for the two dot
// products; lane = channel slice for the vector updates):  sweep A = online (max, sum, sum P g);  sweep B = gradients.
// dk / dv are accumulated in fp32 with vector red.global (different query rows hit the same source token); dq is owned.
// This is the first correct version (parity-first); the gathers are the same L2-resident K/V rows as the forward.
#include "common.cuh"
#include "csr_common.cuh"

namespace sea {
namespace {

constexpr int kBwdWarps = 8;

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return to_f32<T>(*p); }

__device__ __forceinline__ float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }

// dot of a D-vector in shared memory (fp32) with a global row
template <typename T>
__device__ __forceinline__ float dot_row(const float* __restrict__ a, const T* __restrict__ row, int D) {
    float acc = 0.f;
    for (int c = 0; c < D; ++c) acc = fmaf(a[c], to_f32<T>(row[c]), acc);
    return acc;
}
template <>
__device__ __forceinline__ float dot_row<__nv_bfloat16>(const float* __restrict__ a, const __nv_bfloat16* __restrict__ row, int D) {
    float acc = 0.f;
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    for (int c = 0; c < (D >> 3); ++c) {
        const uint4 u = __ldg(r4 + c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc = fmaf(a[8 * c + 2 * i], __uint_as_float(w[i] << 16), acc);
            acc = fmaf(a[8 * c + 2 * i + 1], __uint_as_float(w[i] & 0xffff0000u), acc);
        }
    }
    return acc;
}

// Enumerates the alive source tokens of (row, head) from its bit words in the order the forward kernels use (pixel by
// pixel, descending token inside a pixel, a8 clamp / sub-sampling included) and calls f(jmine, cnt) per 32-entry step.
template <typename F>
__device__ __forceinline__ void for_each_entry_chunk(uint32_t word, int lane, float s_scale, int k_clamp, F&& f) {
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
            f(jmine, cnt);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kBwdWarps * 32)
sparse_attention_bits_bwd_kernel(const uint32_t* __restrict__ mask_bits,
                                 const T* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                 const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                 const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                 const float* __restrict__ scales, const T* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
                                 const T* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                                 float* __restrict__ dscales, int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal) {
    extern __shared__ __align__(16) float bw_sm[];                       // per warp: q[D] | dout[D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* qs = bw_sm + warp * 2 * D;
    float* gs = qs + D;
    const int64_t task = (int64_t) blockIdx.x * kBwdWarps + warp;
    if (task >= (int64_t) N * T_DST * H) return;
    const int t = (int) (task % T_DST);
    const int h = (int) ((task / T_DST) % H);
    const int n = (int) (task / ((int64_t) T_DST * H));
    const int64_t row = (int64_t) n * T_DST + t;
    const T* qrow = q + (int64_t) n * q_sn + (int64_t) h * q_sh + (int64_t) t * q_st;
    const T* drow = dout + row * ((int64_t) H * D) + (int64_t) h * D;
    for (int c = lane; c < D; c += 32) { qs[c] = ldf(qrow + c); gs[c] = ldf(drow + c); }
    __syncwarp();
    const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
    const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;

    const int nw = P >> 5;
    const int L = is_causal ? (T_SRC - T_DST + t + 1) : T_SRC;
    const float s_scale = __fdiv_rn((float) L, (float) P);
    const uint32_t word = lane < nw ? mask_bits[row * ((int64_t) H * nw) + (int64_t) h * nw + lane] : 0u;

    // ---- sweep A: m = max S, l = sum exp(S - m), Eacc = sum exp(S - m) g ---------------------------------------------
    float m_run = -INFINITY, l_run = 0.f, e_run = 0.f;
    for_each_entry_chunk(word, lane, s_scale, k_clamp, [&](int jmine, int cnt) {
        float s = -INFINITY, g = 0.f;
        if (lane < cnt) {
            s = dot_row<T>(qs, kb + (int64_t) jmine * k_st, D);
            g = dot_row<T>(gs, vb + (int64_t) jmine * v_st, D);
        }
        const float m_new = fmaxf(m_run, warp_max(s));
        const float alpha = m_run == -INFINITY ? 0.f : __expf(m_run - m_new);
        const float p = lane < cnt ? __expf(s - m_new) : 0.f;
        l_run = l_run * alpha + warp_sum(p);
        e_run = e_run * alpha + warp_sum(p * g);
        m_run = m_new;
    });
    const float inv_l = l_run > 0.f ? 1.0f / l_run : 0.f;
    const float E = e_run * inv_l;

    const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
    const float psc = use_scaler ? sigmoid_(sp[0]) : 1.0f;
    const float a = cumavg ? sigmoid_(sp[1]) : 1.0f;          // without the running-mean branch the forward is out = ctx
    if (dscales != nullptr) {
        float davg = 0.f;
        if (cumavg) {
            const T* arow = cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st;
            for (int c = lane; c < D; c += 32) davg = fmaf(gs[c], ldf(arow + c), davg);
            davg = warp_sum(davg);
        }
        if (lane == 0) {
            float* ds = dscales + ((((int64_t) n * H + h) * T_DST + t) << 1);
            ds[0] = use_scaler ? a * E * psc * (1.0f - psc) : 0.f;
            ds[1] = cumavg ? (psc * E - davg) * a * (1.0f - a) : 0.f;
        }
    }

    // ---- sweep B: dS_z = a psc P_z (g_z - E);  dq += dS k;  dk_z += dS q;  dv_z += a psc P_z dout --------------------
    // vector updates: D/4 lanes cover one row with 4 channels each, so 32/(D/4) entries are processed per step and every
    // dk / dv update is ONE 16-byte vector reduction (red.global.add.v4.f32) instead of four scalar atomics
    const float w_ctx = a * psc;
    const int lpr = D >> 2, epi = 32 / lpr;                  // lanes per row, entries per step
    const int sub = lane % lpr, grp = lane / lpr;
    float4 dq_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 q4 = *reinterpret_cast<const float4*>(qs + 4 * sub), g4 = *reinterpret_cast<const float4*>(gs + 4 * sub);
    float* dkb = dk + (((int64_t) n * H + h) * T_SRC) * D;
    float* dvb = dv + (((int64_t) n * H + h) * T_SRC) * D;
    for_each_entry_chunk(word, lane, s_scale, k_clamp, [&](int jmine, int cnt) {
        float ds = 0.f, pv = 0.f;
        if (lane < cnt) {
            const float s = dot_row<T>(qs, kb + (int64_t) jmine * k_st, D);
            const float g = dot_row<T>(gs, vb + (int64_t) jmine * v_st, D);
            const float p = __expf(s - m_run) * inv_l;
            pv = w_ctx * p;
            ds = pv * (g - E);
        }
        for (int e0 = 0; e0 < cnt; e0 += epi) {
            const int e = e0 + grp;
            const int j = __shfl_sync(kFull, jmine, e & 31);
            const float ds_e = __shfl_sync(kFull, ds, e & 31);
            const float pv_e = __shfl_sync(kFull, pv, e & 31);
            if (e < cnt) {
                const T* krow = kb + (int64_t) j * k_st + 4 * sub;
                dq_acc.x = fmaf(ds_e, ldf(krow), dq_acc.x); dq_acc.y = fmaf(ds_e, ldf(krow + 1), dq_acc.y);
                dq_acc.z = fmaf(ds_e, ldf(krow + 2), dq_acc.z); dq_acc.w = fmaf(ds_e, ldf(krow + 3), dq_acc.w);
                atomicAdd(reinterpret_cast<float4*>(dkb + (int64_t) j * D) + sub, make_float4(ds_e * q4.x, ds_e * q4.y, ds_e * q4.z, ds_e * q4.w));
                atomicAdd(reinterpret_cast<float4*>(dvb + (int64_t) j * D) + sub, make_float4(pv_e * g4.x, pv_e * g4.y, pv_e * g4.z, pv_e * g4.w));
            }
        }
    });
    // combine the entry groups (lanes sub, sub + lpr, ...) and store dq
    for (int o = lpr; o < 32; o <<= 1) {
        dq_acc.x += __shfl_xor_sync(kFull, dq_acc.x, o); dq_acc.y += __shfl_xor_sync(kFull, dq_acc.y, o);
        dq_acc.z += __shfl_xor_sync(kFull, dq_acc.z, o); dq_acc.w += __shfl_xor_sync(kFull, dq_acc.w, o);
    }
    if (grp == 0) *(reinterpret_cast<float4*>(dq + (((int64_t) n * H + h) * T_DST + t) * D) + sub) = dq_acc;
}

// dv_j += sum_{r >= j} (1 - a_r) dout_r / (r + 1): reverse running sum down the query rows, one thread per (n, h, channel).
// Causal prefill only (T_SRC == T_DST): row r averages v_0 .. v_r (attention.py:1237-1241).
template <typename T>
__global__ void cumavg_bwd_kernel(const T* __restrict__ dout, const float* __restrict__ scales, float* __restrict__ dv,
                                  int N, int H, int Tn, int D) {
    const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t) N * H * D) return;
    const int c = (int) (idx % D);
    const int h = (int) ((idx / D) % H);
    const int n = (int) (idx / ((int64_t) D * H));
    float run = 0.f;
    for (int r = Tn - 1; r >= 0; --r) {
        const float a = sigmoid_(scales[((((int64_t) n * H + h) * Tn + r) << 1) + 1]);
        run = fmaf((1.0f - a) / (float) (r + 1), ldf(dout + ((int64_t) n * Tn + r) * ((int64_t) H * D) + (int64_t) h * D + c), run);
        dv[(((int64_t) n * H + h) * Tn + r) * D + c] += run;
    }
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" int sea_sparse_attention_bits_bwd(const uint32_t* mask_bits,
                                             const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                             const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                             const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                             const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype,
                                             const void* dout, float* dq, float* dk, float* dv, float* dscales,
                                             int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal, void* stream) {
    SEA_CHECK_ARG(mask_bits && q && k && v && scales && dout && dq && dk && dv, "sea_sparse_attention_bits_bwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T_DST > 0 && T_SRC >= T_DST && k_clamp > 0, "sea_sparse_attention_bits_bwd: bad shape");
    SEA_CHECK_ARG((D == 32 || D == 64 || D == 128) && (P % 32) == 0 && P <= 1024, "sea_sparse_attention_bits_bwd: needs D in {32, 64, 128}, P %% 32 == 0, P <= 1024");
    SEA_CHECK_ARG(((((uintptr_t) dq) | ((uintptr_t) dk) | ((uintptr_t) dv)) & 15) == 0, "sea_sparse_attention_bits_bwd: dq / dk / dv must be 16-byte aligned");
    SEA_CHECK_ARG(cumavg == nullptr || (is_causal && T_SRC == T_DST), "sea_sparse_attention_bits_bwd: the running-mean branch needs causal prefill (T_SRC == T_DST)");
    SEA_CHECK_ARG(((k_sn | k_sh | k_st | v_sn | v_sh | v_st) % 8) == 0 && ((((uintptr_t) k) | ((uintptr_t) v)) & 15) == 0,
                  "sea_sparse_attention_bits_bwd: k / v rows must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t) stream;
    SEA_CUDA_TRY(cudaMemsetAsync(dk, 0, (size_t) N * H * T_SRC * D * sizeof(float), s), "memset dk");
    SEA_CUDA_TRY(cudaMemsetAsync(dv, 0, (size_t) N * H * T_SRC * D * sizeof(float), s), "memset dv");
    const int64_t tasks = (int64_t) N * T_DST * H;
    const unsigned grid = (unsigned) ((tasks + kBwdWarps - 1) / kBwdWarps);
    const size_t smem = (size_t) kBwdWarps * 2 * D * sizeof(float);
    SEA_DISPATCH_DTYPE(dtype, T_, {
        sparse_attention_bits_bwd_kernel<T_><<<grid, kBwdWarps * 32, smem, s>>>(
            mask_bits, (const T_*) q, q_sn, q_sh, q_st, (const T_*) k, k_sn, k_sh, k_st, (const T_*) v, v_sn, v_sh, v_st, scales,
            (const T_*) cumavg, avg_sh, avg_st, use_scaler, (const T_*) dout, dq, dk, dv, dscales, N, H, T_DST, T_SRC, D, P, k_clamp, is_causal);
        SEA_CHECK_LAUNCH("sparse_attention_bits_bwd_kernel");
        if (cumavg != nullptr) {
            const int64_t cols = (int64_t) N * H * D;
            cumavg_bwd_kernel<T_><<<(unsigned) ((cols + 127) / 128), 128, 0, s>>>((const T_*) dout, scales, dv, N, H, T_DST, D);
            SEA_CHECK_LAUNCH("cumavg_bwd_kernel");
        }
    });
    return SEA_OK;
}
