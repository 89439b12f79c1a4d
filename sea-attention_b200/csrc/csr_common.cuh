// a8 pixel arithmetic shared by the CSR kernels and the fused predictor tail (which emits the per-row entry counts).
#pragma once
#include "common.cuh"

namespace sea {

// ------------------------------------------------------------------------------------------------
// a8 CSR interpolation.  Pixel m of a row whose (causal) source length is L covers source tokens
// [roundf(m*s), roundf((m+1)*s)),  s = fp32(L)/fp32(P)   (un-fused IEEE ops, roundf = half away).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pixel_bounds(float s, int m, float& vs, float& ve) {
    vs = roundf(__fmul_rn((float) m, s));
    ve = roundf(__fmul_rn((float) (m + 1), s));
}
__device__ __forceinline__ int pixel_width(float s, int m, int k) {
    float vs, ve;
    pixel_bounds(s, m, vs, ve);
    return min((int) __fsub_rn(ve, vs), k);
}

// flat pixel index i = h*P + m; P is a power of two in every shipped config, so avoid the runtime division then
__device__ __forceinline__ int pix_m(int i, int P) { return (P & (P - 1)) == 0 ? (i & (P - 1)) : i % P; }
__device__ __forceinline__ int pix_h(int i, int P) { return (P & (P - 1)) == 0 ? (i >> (31 - __clz(P))) : i / P; }

__device__ __forceinline__ int word_width_sum(uint32_t word, int w, int P, float s, int k) {
    int acc = 0;
    for (uint32_t x = word; x; x &= x - 1) acc += pixel_width(s, pix_m((w << 5) + __ffs(x) - 1, P), k);
    return acc;
}

}  // namespace sea
