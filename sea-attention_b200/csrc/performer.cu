// a2 + a3: causal Performer estimate (FastAttention, generalized ReLU features) and the causal
// running mean of v (a13), as a chunk-parallel scan.
//   pass A: per super-chunk (kRC rows) sums  S_c = sum phi(k_r) (x) v2_r,  z_c = sum phi(k_r),  vs_c = sum v_r
//   pass B: exclusive prefix over the super-chunks of one (n, h)
//   pass C: per super-chunk, 32-row sub-chunks: out = (tril(phi(q) phi(k)^T) v2 + phi(q) S) / den
// v2 = cat(pos_emb[t], v) is formed on the fly (reference attention.py:504-508), never materialised.
// Reference: performer_pytorch.FastAttention (call sites attention.py:159-164, 527-534, 559-572);
// running mean attention.py:1237-1241.
#include "common.cuh"
#include "tile_gemm.cuh"

namespace sea {

constexpr int kRC = 256;   // rows per super-chunk (one CTA)
constexpr int kSC = 32;    // rows per sub-chunk
constexpr int kPerfThreads = 256;

struct PerfDims {
    int N, H, T, D, F, Fp, E;   // Fp = F rounded up to 4, E = 2*D
    int nchunks;
    int64_t ws_stride;          // floats per chunk slot: Fp*E + Fp + D
};

template <typename T>
__device__ __forceinline__ void load_rows_f32(float* dst, int ld, const T* src, int64_t row_stride, int r0, int nrows_valid,
                                              int nrows, int width, float scale) {
    // dst[r][c] = scale * src[(r0+r)*row_stride + c] for r < nrows_valid, 0 for the rest
    for (int idx = threadIdx.x; idx < nrows * width; idx += blockDim.x) {
        int r = idx / width, c = idx % width;
        float val = 0.f;
        if (r < nrows_valid) val = scale * to_f32(src[(int64_t) (r0 + r) * row_stride + c]);
        dst[r * ld + c] = val;
    }
}

template <typename T, bool kSumsOnly>
__global__ void __launch_bounds__(kPerfThreads)
performer_chunk_kernel(const T* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                       const T* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                       const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                       const float* __restrict__ pos_emb, const float* __restrict__ proj,
                       T* __restrict__ ctx, T* __restrict__ cumavg, float* __restrict__ ws, PerfDims dm) {
    extern __shared__ float smem[];
    const int D = dm.D, F = dm.F, Fp = dm.Fp, E = dm.E;
    const int chunk = blockIdx.x;
    const int nh = blockIdx.y;
    const int n = nh / dm.H, h = nh % dm.H;
    const int tid = threadIdx.x;

    // shared-memory carve-up (floats)
    float* projT = smem;                       // [D][Fp]
    float* S = projT + D * Fp;                 // [Fp][E]
    float* z = S + Fp * E;                     // [Fp]
    float* vsum = z + Fp;                      // [D]
    float* xk = vsum + D;                      // [kSC][D]
    float* v2 = xk + kSC * D;                  // [kSC][E]
    float* phikT = v2 + kSC * E;               // [Fp][kSC]
    float* xq = phikT + Fp * kSC;              // [kSC][D]        (pass C only)
    float* phiq = xq + kSC * D;                // [kSC][Fp]
    float* A = phiq + kSC * Fp;                // [kSC][kSC+1]
    float* den = A + kSC * (kSC + 1);          // [kSC]

    for (int idx = tid; idx < D * Fp; idx += kPerfThreads) {
        int c = idx / Fp, f = idx % Fp;
        projT[idx] = f < F ? proj[f * D + c] : 0.f;
    }
    float* slot = ws + ((int64_t) nh * dm.nchunks + chunk) * dm.ws_stride;
    if (kSumsOnly) {
        for (int idx = tid; idx < Fp * E + Fp + D; idx += kPerfThreads) S[idx] = 0.f;   // S, z, vsum are contiguous
    } else {
        for (int idx = tid; idx < Fp * E + Fp + D; idx += kPerfThreads) S[idx] = slot[idx];
    }
    __syncthreads();

    const float norm = rsqrtf(sqrtf((float) D));   // D^-1/4
    const T* qb = q + (int64_t) n * q_sn + (int64_t) h * q_sh;
    const T* kb = k + (int64_t) n * k_sn + (int64_t) h * k_sh;
    const T* vb = v + (int64_t) n * v_sn + (int64_t) h * v_sh;
    const int row_begin = chunk * kRC;
    const int row_end = min(dm.T, row_begin + kRC);

    for (int r0 = row_begin; r0 < row_end; r0 += kSC) {
        const int nv = min(kSC, row_end - r0);
        load_rows_f32<T>(xk, D, kb, k_st, r0, nv, kSC, D, norm);
        if (!kSumsOnly) load_rows_f32<T>(xq, D, qb, q_st, r0, nv, kSC, D, norm);
        for (int idx = tid; idx < kSC * E; idx += kPerfThreads) {
            int r = idx / E, c = idx % E;
            float val = 0.f;
            if (r < nv) val = c < D ? pos_emb[(int64_t) (r0 + r) * D + c] : to_f32(vb[(int64_t) (r0 + r) * v_st + (c - D)]);
            v2[idx] = val;
        }
        __syncthreads();
        // phi(k)^T [Fp][kSC]; padded features and rows beyond T contribute nothing
        tile_gemm(kSC, Fp, D, [&](int i, int c) { return xk[i * D + c]; }, [&](int c, int f) { return projT[c * Fp + f]; },
                  [&](int i, int f, float acc) { phikT[f * kSC + i] = (f < F && i < nv) ? fmaxf(acc, 0.f) + 1e-3f : 0.f; });
        if (!kSumsOnly) {
            tile_gemm(kSC, Fp, D, [&](int i, int c) { return xq[i * D + c]; }, [&](int c, int f) { return projT[c * Fp + f]; },
                      [&](int i, int f, float acc) { phiq[i * Fp + f] = f < F ? fmaxf(acc, 0.f) + 1e-3f : 0.f; });
        }
        __syncthreads();
        if (!kSumsOnly) {
            // A[i][j] = phi(q_i) . phi(k_j), j <= i
            tile_gemm(kSC, kSC, Fp, [&](int i, int f) { return phiq[i * Fp + f]; }, [&](int f, int j) { return phikT[f * kSC + j]; },
                      [&](int i, int j, float acc) { A[i * (kSC + 1) + j] = j <= i ? acc : 0.f; });
            __syncthreads();
            // den[i] = sum_j A[i][j] + phi(q_i) . (z + 1e-6)
            {
                const int lane = tid & 31, wid = tid >> 5;
                for (int i = wid; i < kSC; i += kPerfThreads / 32) {
                    float part = A[i * (kSC + 1) + lane];
                    for (int f = lane; f < F; f += 32) part = fmaf(phiq[i * Fp + f], z[f] + 1e-6f, part);
                    part = warp_sum(part);
                    if (lane == 0) den[i] = part;
                }
            }
            __syncthreads();
            // out[i][e] = (sum_j A[i][j] v2[j][e] + sum_f phi(q_i)[f] S[f][e]) / den[i]
            T* ob = ctx + (((int64_t) n * dm.H + h) * dm.T) * E;
            const int ntj = E >> 2;
            for (int tile = tid; tile < (kSC / 4) * ntj; tile += kPerfThreads) {
                const int i0 = (tile / ntj) << 2, j0 = (tile % ntj) << 2;
                float acc[4][4] = {};
                for (int kk = 0; kk < kSC; ++kk) {
                    const float4 bv = *reinterpret_cast<const float4*>(&v2[kk * E + j0]);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float av = A[(i0 + x) * (kSC + 1) + kk];
                        acc[x][0] = fmaf(av, bv.x, acc[x][0]); acc[x][1] = fmaf(av, bv.y, acc[x][1]);
                        acc[x][2] = fmaf(av, bv.z, acc[x][2]); acc[x][3] = fmaf(av, bv.w, acc[x][3]);
                    }
                }
                for (int f = 0; f < Fp; ++f) {
                    const float4 bv = *reinterpret_cast<const float4*>(&S[f * E + j0]);
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float av = phiq[(i0 + x) * Fp + f];
                        acc[x][0] = fmaf(av, bv.x, acc[x][0]); acc[x][1] = fmaf(av, bv.y, acc[x][1]);
                        acc[x][2] = fmaf(av, bv.z, acc[x][2]); acc[x][3] = fmaf(av, bv.w, acc[x][3]);
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (i0 + x < nv) {
                        const float inv = 1.0f / den[i0 + x];
                        T* o = ob + (int64_t) (r0 + i0 + x) * E + j0;
#pragma unroll
                        for (int y = 0; y < 4; ++y) o[y] = from_f32<T>(acc[x][y] * inv);
                    }
                }
            }
            // running mean of v (attention.py:1237-1241): thread per channel walks the sub-chunk
            if (cumavg != nullptr) {
                T* cb = cumavg + (((int64_t) n * dm.H + h) * dm.T) * D;
                for (int c = tid; c < D; c += kPerfThreads) {
                    float run = vsum[c];
                    for (int r = 0; r < nv; ++r) {
                        run += v2[r * E + D + c];
                        cb[(int64_t) (r0 + r) * D + c] = from_f32<T>(run / (float) (r0 + r + 1));
                    }
                }
            }
            __syncthreads();
        }
        // state update: S += phi(k)^T v2, z += sum_r phi(k_r), vsum += sum_r v_r
        tile_gemm(Fp, E, kSC, [&](int f, int r) { return phikT[f * kSC + r]; }, [&](int r, int e) { return v2[r * E + e]; },
                  [&](int f, int e, float acc) { S[f * E + e] += acc; });
        for (int f = tid; f < Fp; f += kPerfThreads) {
            float a = 0.f;
            for (int r = 0; r < kSC; ++r) a += phikT[f * kSC + r];
            z[f] += a;
        }
        for (int c = tid; c < D; c += kPerfThreads) {
            float a = 0.f;
            for (int r = 0; r < nv; ++r) a += v2[r * E + D + c];
            vsum[c] += a;
        }
        __syncthreads();
    }
    if (kSumsOnly) {
        for (int idx = tid; idx < Fp * E + Fp + D; idx += kPerfThreads) slot[idx] = S[idx];
    }
}

// exclusive prefix over the chunk slots of one (n, h): slot[c] <- sum_{c' < c} slot[c']
__global__ void __launch_bounds__(256)
performer_prefix_kernel(float* __restrict__ ws, int nchunks, int64_t ws_stride) {
    float* base = ws + (int64_t) blockIdx.y * nchunks * ws_stride;
    constexpr int kBatch = 16;      // independent loads issued together, then the serial prefix
    for (int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < ws_stride; idx += (int64_t) gridDim.x * blockDim.x) {
        float run = 0.f;
        for (int c0 = 0; c0 < nchunks; c0 += kBatch) {
            float cur[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) cur[i] = (c0 + i < nchunks) ? __ldcg(base + (int64_t) (c0 + i) * ws_stride + idx) : 0.f;
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                if (c0 + i < nchunks) __stcg(base + (int64_t) (c0 + i) * ws_stride + idx, run);
                run += cur[i];
            }
        }
    }
}

static PerfDims make_dims(int N, int H, int T, int D, int F) {
    PerfDims dm;
    dm.N = N; dm.H = H; dm.T = T; dm.D = D; dm.F = F;
    dm.Fp = (F + 3) & ~3;
    dm.E = 2 * D;
    dm.nchunks = (T + kRC - 1) / kRC;
    dm.ws_stride = (int64_t) dm.Fp * dm.E + dm.Fp + D;
    return dm;
}

static size_t perf_smem_bytes(const PerfDims& dm) {
    size_t fl = (size_t) dm.D * dm.Fp + (size_t) dm.Fp * dm.E + dm.Fp + dm.D + (size_t) kSC * dm.D + (size_t) kSC * dm.E +
                (size_t) dm.Fp * kSC + (size_t) kSC * dm.D + (size_t) kSC * dm.Fp + (size_t) kSC * (kSC + 1) + kSC;
    return fl * sizeof(float);
}

}  // namespace sea

using namespace sea;

extern "C" {

int64_t sea_performer_workspace_floats(int N, int H, int T, int D, int F) {
    if (N <= 0 || H <= 0 || T <= 0 || D <= 0 || F <= 0) return 0;
    PerfDims dm = make_dims(N, H, T, D, F);
    return (int64_t) N * H * dm.nchunks * dm.ws_stride;
}

int sea_performer_causal_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                             const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                             const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                             const float* pos_emb, const float* proj, int dtype, void* ctx, void* cumavg, float* workspace,
                             int N, int H, int T, int D, int F, void* stream) {
    SEA_CHECK_ARG(q && k && v && pos_emb && proj && ctx && workspace, "sea_performer_causal_fwd: null pointer");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && D > 0 && F > 0, "sea_performer_causal_fwd: bad shape");
    SEA_CHECK_ARG((D & 3) == 0, "sea_performer_causal_fwd: head dim %d must be a multiple of 4", D);
    PerfDims dm = make_dims(N, H, T, D, F);
    const size_t smem = perf_smem_bytes(dm);
    SEA_CHECK_ARG(smem <= 227 * 1024, "sea_performer_causal_fwd: D=%d F=%d needs %zu B of shared memory (> 227 KB)", D, F, smem);
    SEA_CHECK_ARG((int64_t) N * H <= 65535, "sea_performer_causal_fwd: N*H too large");
    cudaStream_t s = (cudaStream_t) stream;
    dim3 grid(dm.nchunks, N * H);
    SEA_DISPATCH_DTYPE(dtype, T_, {
        auto ka = performer_chunk_kernel<T_, true>;
        auto kc = performer_chunk_kernel<T_, false>;
        SEA_CUDA_TRY(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        SEA_CUDA_TRY(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        if (dm.nchunks > 1) {
            ka<<<grid, kPerfThreads, smem, s>>>((const T_*) q, q_sn, q_sh, q_st, (const T_*) k, k_sn, k_sh, k_st,
                                                (const T_*) v, v_sn, v_sh, v_st, pos_emb, proj, nullptr, nullptr, workspace, dm);
            SEA_CHECK_LAUNCH("performer_chunk_kernel<sums>");
            dim3 pgrid((unsigned) ((dm.ws_stride + 255) / 256), N * H);
            performer_prefix_kernel<<<pgrid, 256, 0, s>>>(workspace, dm.nchunks, dm.ws_stride);
            SEA_CHECK_LAUNCH("performer_prefix_kernel");
        } else {
            SEA_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t) N * H * dm.ws_stride * sizeof(float), s), "memset");
        }
        kc<<<grid, kPerfThreads, smem, s>>>((const T_*) q, q_sn, q_sh, q_st, (const T_*) k, k_sn, k_sh, k_st,
                                            (const T_*) v, v_sn, v_sh, v_st, pos_emb, proj, (T_*) ctx, (T_*) cumavg, workspace, dm);
        SEA_CHECK_LAUNCH("performer_chunk_kernel<out>");
    });
    return SEA_OK;
}

}  // extern "C"
