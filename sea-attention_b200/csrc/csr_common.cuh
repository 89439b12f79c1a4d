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

// First source token of pixel m for a row of source length L (a8: roundf(fp32(m) * (fp32(L) / fp32(P))), un-fused).
// When P = 2^lg and m * L < 2^24 every intermediate is exact, so roundf(m * L / P) == (m * L + P/2) >> lg: integer path.
struct RowScale {
    int L, lg, halfP;
    float s;
    __device__ __forceinline__ int edge(int m) const {
        return lg >= 0 ? (m * L + halfP) >> lg : (int) roundf(__fmul_rn((float) m, s));
    }
};
__host__ __device__ inline int exact_edge_shift(int P, int T_SRC) {     // lg(P) when the integer path is exact, else -1
    if ((P & (P - 1)) != 0 || (int64_t) P * T_SRC > (1 << 24)) return -1;
    int lg = 0;
    while ((1 << lg) < P) ++lg;
    return lg;
}

// Dense bit-packed partial_attention_mask of the short-context attention path (block_attn.cu): one u64 per (head, query row,
// 64-token tile), rows padded to an even word count; tile activity is kept per (head, kMaskRowBlock-row query block).
constexpr int kMaskRowBlock = 128;
constexpr int kMaskTile = 64;
__host__ __device__ inline int mask_row_words(int T_SRC) { return (((T_SRC + kMaskTile - 1) / kMaskTile) + 1) & ~1; }
__host__ __device__ inline int mask_act_words(int T_SRC) { return ((T_SRC + kMaskTile - 1) / kMaskTile + 31) / 32; }
inline int64_t mask_act_bytes(int N, int H, int T_DST, int T_SRC) {
    const int64_t words = (int64_t) N * H * ((T_DST + kMaskRowBlock - 1) / kMaskRowBlock) * mask_act_words(T_SRC);
    return (words * 4 + 15) & ~(int64_t) 15;
}

}  // namespace sea
