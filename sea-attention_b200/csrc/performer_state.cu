// Decode / use_cache form of a2 + a3 + a13: the causal Performer estimate and the running mean of v advanced
// token by token from a state, instead of the chunk-parallel scan of performer.cu / performer_mma.cu.
//   state per (n, h):  S [F][2D] = sum_s phi(k_s) (x) v2_s,   z [F] = sum_s phi(k_s),   vsum [D] = sum_s v_s      (fp32)
//   per new token t:   S, z, vsum are advanced with k_t, v2_t = cat(pos_emb[t], v_t);
//                      ctx_t = (phi(q_t) . S) / (phi(q_t) . (z + 1e-6)),   cumavg_t = vsum / (t + 1)
//   phi(x) = relu(d^-1/4 x P^T) + 1e-3   (generalized-attention features, as in the prefill kernels)
// Replaces the reference's StatefulCausalPerformer / StatefulCumAvg (attention_state.py:43-98, 205-224), which keep the same
// running sums (in fp64) and are checked by the reference against the stateless forward (test_perlin_opt_cache.py); here the
// parity target is the prefill path row by row.  CTA = one (n, h); tokens are processed sequentially (T_new = 1 when decoding;
// a whole prompt can be replayed to build the state after a prefill).  fp32 arithmetic throughout.
#include "common.cuh"

namespace sea {
namespace {

constexpr int kStThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kStThreads)
performer_state_kernel(const T* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                       const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                       const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                       const float* __restrict__ pos_emb, const float* __restrict__ proj, float* __restrict__ state,
                       T* __restrict__ ctx, T* __restrict__ cumavg, int H, int T_new, int t0, int D, int F) {
    extern __shared__ float st_sm[];
    const int E = 2 * D;
    float* S = st_sm;                 // [F][E]
    float* z = S + F * E;             // [F]
    float* vs = z + F;                // [D]
    float* pr = vs + D;               // [F][D] projection
    float* xk = pr + F * D;           // [D]
    float* xq = xk + D;               // [D]
    float* v2 = xq + D;               // [E]
    float* fk = v2 + E;               // [F]
    float* fq = fk + F;               // [F]
    float* red = fq + F;              // [1] denominator
    const int nh = blockIdx.x, n = nh / H, h = nh % H;
    const int tid = threadIdx.x;
    const int64_t st_stride = (int64_t) F * E + F + D;
    float* gs = state + (int64_t) nh * st_stride;
    for (int i = tid; i < F * E + F + D; i += kStThreads) st_sm[i] = gs[i];       // S | z | vsum are contiguous in both layouts
    for (int i = tid; i < F * D; i += kStThreads) pr[i] = proj[i];
    __syncthreads();
    const float norm = rsqrtf(sqrtf((float) D));
    const T* qb = q + (int64_t) n * q_sn + (int64_t) h * q_sh;
    const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
    const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;
    for (int tt = 0; tt < T_new; ++tt) {
        const int t = t0 + tt;
        for (int c = tid; c < D; c += kStThreads) {
            xk[c] = to_f32(kb[(int64_t) tt * k_st + c]);
            xq[c] = to_f32(qb[(int64_t) tt * q_st + c]);
            const float vv = to_f32(vb[(int64_t) tt * v_st + c]);
            v2[c] = pos_emb[(int64_t) t * D + c];
            v2[D + c] = vv;
            vs[c] += vv;
        }
        __syncthreads();
        for (int f = tid; f < 2 * F; f += kStThreads) {          // phi(k) and phi(q)
            const float* x = f < F ? xk : xq;
            const float* p = pr + (f < F ? f : f - F) * D;
            float acc = 0.f;
            for (int c = 0; c < D; ++c) acc = fmaf(x[c], p[c], acc);
            (f < F ? fk : fq)[f < F ? f : f - F] = fmaxf(acc * norm, 0.f) + 1e-3f;
        }
        __syncthreads();
        for (int i = tid; i < F * E; i += kStThreads) S[i] = fmaf(fk[i / E], v2[i % E], S[i]);
        for (int f = tid; f < F; f += kStThreads) z[f] += fk[f];
        __syncthreads();
        if (tid < 32) {
            float d = 0.f;
            for (int f = tid; f < F; f += 32) d = fmaf(fq[f], z[f] + 1e-6f, d);
            d = warp_sum(d);
            if (tid == 0) red[0] = d;
        }
        __syncthreads();
        const float inv = 1.0f / red[0];
        for (int e = tid; e < E; e += kStThreads) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(fq[f], S[f * E + e], acc);
            ctx[(((int64_t) n * H + h) * T_new + tt) * E + e] = from_f32<T>(acc * inv);
        }
        if (cumavg != nullptr)
            for (int c = tid; c < D; c += kStThreads)
                cumavg[(((int64_t) n * H + h) * T_new + tt) * D + c] = from_f32<T>(vs[c] / (float) (t + 1));
        __syncthreads();
    }
    for (int i = tid; i < F * E + F + D; i += kStThreads) gs[i] = st_sm[i];
}

// State after a whole prompt, in parallel: every CTA reduces one 128-token chunk of one (n, h) into partial sums and adds them
// to the state with fp32 atomics (the token-by-token kernel above would take T sequential steps).
constexpr int kBuildChunk = 128;

template <typename T>
__global__ void __launch_bounds__(kStThreads)
performer_state_build_kernel(const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                             const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                             const float* __restrict__ pos_emb, const float* __restrict__ proj, float* __restrict__ state,
                             int H, int Tn, int D, int F, int chunk) {
    extern __shared__ float sb_sm[];
    const int E = 2 * D;
    float* pr = sb_sm;                          // [F][D]
    float* xk = pr + F * D;                     // [chunk][D]
    float* v2 = xk + chunk * D;                 // [chunk][E]
    float* fk = v2 + chunk * E;                 // [chunk][F]
    const int nh = blockIdx.y, n = nh / H, h = nh % H;
    const int t0 = blockIdx.x * chunk, nt = min(chunk, Tn - t0);
    const int tid = threadIdx.x;
    const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
    const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;
    for (int i = tid; i < F * D; i += kStThreads) pr[i] = proj[i];
    for (int i = tid; i < nt * D; i += kStThreads) {
        const int r = i / D, c = i - r * D;
        xk[i] = to_f32(kb[(int64_t) (t0 + r) * k_st + c]);
        v2[r * E + c] = pos_emb[(int64_t) (t0 + r) * D + c];
        v2[r * E + D + c] = to_f32(vb[(int64_t) (t0 + r) * v_st + c]);
    }
    __syncthreads();
    const float norm = rsqrtf(sqrtf((float) D));
    for (int i = tid; i < nt * F; i += kStThreads) {
        const int r = i / F, f = i - r * F;
        float acc = 0.f;
        for (int c = 0; c < D; ++c) acc = fmaf(xk[r * D + c], pr[f * D + c], acc);
        fk[i] = fmaxf(acc * norm, 0.f) + 1e-3f;
    }
    __syncthreads();
    float* gs = state + (int64_t) nh * ((int64_t) F * E + F + D);
    for (int i = tid; i < F * E; i += kStThreads) {
        const int f = i / E, e = i - f * E;
        float acc = 0.f;
        for (int r = 0; r < nt; ++r) acc = fmaf(fk[r * F + f], v2[r * E + e], acc);
        atomicAdd(gs + i, acc);
    }
    for (int f = tid; f < F; f += kStThreads) {
        float acc = 0.f;
        for (int r = 0; r < nt; ++r) acc += fk[r * F + f];
        atomicAdd(gs + F * E + f, acc);
    }
    for (int c = tid; c < D; c += kStThreads) {
        float acc = 0.f;
        for (int r = 0; r < nt; ++r) acc += v2[r * E + D + c];
        atomicAdd(gs + F * E + F + c, acc);
    }
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" {

int64_t sea_performer_state_floats(int N, int H, int D, int F) {
    if (N <= 0 || H <= 0 || D <= 0 || F <= 0) return 0;
    return (int64_t) N * H * ((int64_t) F * 2 * D + F + D);
}

int sea_performer_causal_state_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                   const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                   const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                   const float* pos_emb, const float* proj, int dtype, float* state, void* ctx, void* cumavg,
                                   int N, int H, int T_new, int t0, int D, int F, void* stream) {
    SEA_CHECK_ARG(q && k && v && pos_emb && proj && state && ctx, "sea_performer_causal_state_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T_new > 0 && t0 >= 0 && D > 0 && F > 0, "sea_performer_causal_state_fwd: bad shape");
    const size_t smem = ((size_t) F * 2 * D + F + D + (size_t) F * D + 2 * D + 2 * D + 2 * F + 4) * sizeof(float);
    SEA_CHECK_ARG(smem <= 200 * 1024, "sea_performer_causal_state_fwd: F * D too large");
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto kern = performer_state_kernel<T_>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        kern<<<(unsigned) ((int64_t) N * H), kStThreads, smem, (cudaStream_t) stream>>>(
            (const T_*) q, q_sn, q_sh, q_st, (const T_*) k, k_sn, k_sh, k_st, (const T_*) v, v_sn, v_sh, v_st, pos_emb, proj, state,
            (T_*) ctx, (T_*) cumavg, H, T_new, t0, D, F);
        SEA_CHECK_LAUNCH("performer_state_kernel");
    });
    return SEA_OK;
}

int sea_performer_state_build(const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                              const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                              const float* pos_emb, const float* proj, int dtype, float* state,
                              int N, int H, int T, int D, int F, void* stream) {
    SEA_CHECK_ARG(k && v && pos_emb && proj && state, "sea_performer_state_build: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && D > 0 && F > 0 && (int64_t) N * H <= 65535, "sea_performer_state_build: bad shape");
    // rows per CTA: as many as fit 200 KB of shared memory next to the projection (128 at d = 64; 64 at d = 128, F = 77)
    int chunk = kBuildChunk;
    while (chunk > 8 && ((size_t) F * D + (size_t) chunk * (3 * D + F)) * sizeof(float) > 200 * 1024) chunk >>= 1;
    const size_t smem = ((size_t) F * D + (size_t) chunk * (3 * D + F)) * sizeof(float);
    SEA_CHECK_ARG(smem <= 200 * 1024, "sea_performer_state_build: F * D too large");
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto kern = performer_state_build_kernel<T_>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        kern<<<dim3((unsigned) ((T + chunk - 1) / chunk), (unsigned) (N * H)), kStThreads, smem, (cudaStream_t) stream>>>(
            (const T_*) k, k_sn, k_sh, k_st, (const T_*) v, v_sn, v_sh, v_st, pos_emb, proj, state, H, T, D, F, chunk);
        SEA_CHECK_LAUNCH("performer_state_build_kernel");
    });
    return SEA_OK;
}

}  // extern "C"
