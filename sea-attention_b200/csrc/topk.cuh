// a7 grouped top-k as a CTA-level device routine (shared by the standalone kernel and the fused predictor tail).
#pragma once
#include "common.cuh"

namespace sea {

// ------------------------------------------------------------------------------------------------
// a7 top-k: one CTA per group; keys staged once in shared memory as order-preserving u32; 4-pass
// 8-bit radix select finds the K-th largest key; ties at the threshold are resolved in index order
// with a block scan so that the LOWER flat index wins.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t orderable(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u << 1) == 0) return 0x80000000u;  // +-0 compare equal
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kTopkThreads = 256;

__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int& total) {
    // exclusive scan of one int per thread over a 256-thread CTA; `total` = sum over the CTA
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_scan_incl_i(v, lane);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int wprefix = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kTopkThreads / 32; ++w) {
        int s = warp_sums[w];
        if (w < wid) wprefix += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return wprefix + incl - v;
}

// Select over `G` orderable keys resident in shared memory; writes ceil(G/32) words of alive bits.
static __device__ void topk_select_to_bits(const uint32_t* skeys, int G, int K, uint32_t* out_bits,
                                    int* hist /*256*/, int* scratch /*16*/) {
    const int tid = threadIdx.x;
    const int nwords = (G + 31) >> 5;
    if (K >= G) {
        for (int w = tid; w < nwords; w += kTopkThreads) {
            int rem = G - (w << 5);
            out_bits[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        }
        return;
    }
    if (K <= 0) {
        for (int w = tid; w < nwords; w += kTopkThreads) out_bits[w] = 0u;
        return;
    }
    // Probabilities of one row share their leading bits (same sign, nearly the same exponent): a radix pass over those
    // bits would funnel every key into one histogram bin (fully serialised shared-memory atomics).  Find the common
    // prefix with a CTA-wide OR / AND and start the select at the first bit that actually differs.
    uint32_t k_or = 0, k_and = 0xffffffffu;
    for (int i = tid; i < G; i += kTopkThreads) { const uint32_t u = skeys[i]; k_or |= u; k_and &= u; }
    k_or = __reduce_or_sync(kFull, k_or);
    k_and = __reduce_and_sync(kFull, k_and);
    if ((tid & 31) == 0) { hist[tid >> 5] = (int) k_or; hist[8 + (tid >> 5)] = (int) k_and; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kTopkThreads / 32; ++w) { k_or |= (uint32_t) hist[w]; k_and &= (uint32_t) hist[8 + w]; }
    __syncthreads();
    const uint32_t diff = k_or ^ k_and;
    const int npass = diff == 0 ? 0 : ((31 - __clz(diff)) >> 3) + 1;
    uint32_t mask = npass >= 4 ? 0u : (0xffffffffu << (8 * npass));
    uint32_t prefix = k_and & mask;
    int remaining = K;
    for (int pass = 4 - npass; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < G; i += kTopkThreads) {
            uint32_t u = skeys[i];
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255], 1);
        }
        __syncthreads();
        // suffix sums over digits: thread d gets count of keys with digit > d (within the prefix class)
        int mine = hist[255 - tid];  // reversed so that an exclusive scan gives "strictly greater"
        int tot;
        int above = block_excl_scan(mine, scratch, tot);
        // digit d = 255 - tid is the pivot digit iff above < remaining <= above + mine
        if (above < remaining && remaining <= above + mine) {
            scratch[8] = 255 - tid;
            scratch[9] = remaining - above;
        }
        __syncthreads();
        prefix |= (uint32_t) scratch[8] << shift;
        mask |= 0xffu << shift;
        remaining = scratch[9];
        __syncthreads();
    }
    const uint32_t thr = prefix;  // K-th largest key; `remaining` of the keys equal to thr are alive
    int carry = 0;
    for (int w0 = 0; w0 < nwords; w0 += kTopkThreads) {
        const int w = w0 + tid;
        uint32_t gt = 0, eq = 0;
        if (w < nwords) {
            const int base = w << 5;
            const int lim = min(32, G - base);
            // thread = word, so a plain b-loop would be a 32-way bank conflict: rotate by tid
            for (int r = 0; r < 32; ++r) {
                const int b = (r + tid) & 31;
                if (b < lim) {
                    uint32_t u = skeys[base + b];
                    gt |= (u > thr ? 1u : 0u) << b;
                    eq |= (u == thr ? 1u : 0u) << b;
                }
            }
        }
        int neq = __popc(eq);
        int tot;
        int before = carry + block_excl_scan(neq, scratch, tot);
        carry += tot;
        if (w < nwords) {
            int take = remaining - before;  // how many of my equal keys (in index order) are alive
            uint32_t sel = 0;
            if (take >= neq) sel = eq;
            else if (take > 0) {
                uint32_t e = eq;
                for (int c = 0; c < take; ++c) {
                    uint32_t low = e & (~e + 1u);
                    sel |= low;
                    e ^= low;
                }
            }
            out_bits[w] = gt | sel;
        }
    }
}

}  // namespace sea
