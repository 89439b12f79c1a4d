// CTA-cooperative small GEMM on shared-memory operands with 4x4 register tiles (fp32 SIMT).
// Used by the fp32 (exactness) path of the dense stages; the bf16 path runs on tcgen05 (umma_*.cu).
#pragma once
#include "common.cuh"

namespace sea {

// C[i][j] = sum_k a(i,k) * b(k,j) for i<M, j<N; 4x4 tiles are dealt round-robin to the CTA's threads
// with j fastest, so b(k, j0..j0+3) of neighbouring lanes is contiguous and a(i,k) is a broadcast.
// a/b are functors returning 0 outside their logical bounds is NOT required: indices are clamped
// by the guards below.  epi(i, j, value) is called once per valid output.
template <class FA, class FB, class FE>
__device__ __forceinline__ void tile_gemm(int M, int N, int K, FA a, FB b, FE epi) {
    const int ntj = (N + 3) >> 2, nti = (M + 3) >> 2;
    const int ntiles = nti * ntj;
    for (int tile = threadIdx.x; tile < ntiles; tile += blockDim.x) {
        const int i0 = (tile / ntj) << 2, j0 = (tile % ntj) << 2;
        float acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
        const bool full = (i0 + 4 <= M) && (j0 + 4 <= N);
        if (full) {
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                float av[4], bv[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) av[x] = a(i0 + x, k);
#pragma unroll
                for (int y = 0; y < 4; ++y) bv[y] = b(k, j0 + y);
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
            }
        } else {
            for (int k = 0; k < K; ++k) {
                float av[4], bv[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) av[x] = (i0 + x < M) ? a(i0 + x, k) : 0.f;
#pragma unroll
                for (int y = 0; y < 4; ++y) bv[y] = (j0 + y < N) ? b(k, j0 + y) : 0.f;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
            }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
                if (i0 + x < M && j0 + y < N) epi(i0 + x, j0 + y, acc[x][y]);
    }
}

}  // namespace sea
