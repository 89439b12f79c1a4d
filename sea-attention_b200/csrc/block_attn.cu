// a8-a14 fused, short-context form: the SEA sparse attention as a tile-skipping, element-masked flash attention.
//
// Why a second kernel: at the north-star shape (T 4096, k 64, P 256) every (row, head) attends ~64 source tokens spread
// over ~4..64 interpolated pixels, and neighbouring query rows rarely share pixels, so the gather kernel
// (sparse_attn.cu) moves 2*Z*d*2 = 2.1 GB of K/V rows through L2 per layer and is bound there (~0.38 ms).
// Measured on the real top-k masks of the random-init layer (profiles/r01d_tile_stats.txt): 91 % of the 64x64
// (query block x source block) tiles hold at least one alive element, but only 42 % of their 16x16 sub-blocks do.
// So, two kernels:
//  expand_mask_kernel   a8 in dense bit-packed form: one u64 per (head, query row, 64-token tile), built from the ALIVE
//                       pixels only with the a8 pixel arithmetic (exact integer form when P is a power of two), so the set
//                       of (row, source token) pairs is bit-identical to the CSR the reference builds
//                       (causal_resize_m_to_t.py:648-762) as long as no pixel is clamped (span <= k; checked by shape).
//                       67 MB at the north-star shape -- the reference's dense partial_attention_mask at 1 bit / element.
//  block_attention_bits_kernel   CTA = 128 query rows of one head (8 warps x 16 rows);  K/V tiles of 64 source tokens are
//                       fetched ONCE per CTA by TMA (SWIZZLE_128B boxes straight from the strided [N,H,T,d] tensors, 4-stage
//                       ring: a "full" mbarrier per stage, and the last warp to release a stage refills it, so no warp ever
//                       spins as a producer) and shared by all 128 rows (L2 traffic 2.1 GB -> 0.5 GB);  tiles in which no row
//                       of the block is alive are skipped;  mma.sync m16n8k16 fed by ldmatrix from the swizzled tiles.
// Work is O(active tiles), i.e. O(T^2) in the worst case: callers keep the gather kernel (O(T k)) for long contexts.
//
// Softmax: fp32, online, exponent fused as ex2(s * log2e - m * log2e); the per-row reference maximum moves lazily
// (only when a tile exceeds it by more than 2^8), so the accumulator is almost never rescaled;
// epilogue = * sigmoid(s0), mix with the causal running mean, permuted store [N, T, H*D]
// (reference attention.py:1151-1173, 1237-1244, 1279-1282).
#include "common.cuh"
#include "csr_common.cuh"
#include "umma.cuh"
#include "block_attn.cuh"

#include <stdlib.h>

namespace sea {
namespace {

constexpr int kBM = 128;           // query rows per CTA (8 warps x 16 rows)
constexpr int kBN = 64;            // source tokens per tile
constexpr int kBD = 64;            // head dim: one K / V row = 128 bytes = one SWIZZLE_128B row
constexpr int kBWarps = kBM / 16;
constexpr int kBThreads = kBWarps * 32;
constexpr int kStages = 4;
constexpr int kTileBytes = kBN * kBD * 2;      // 8 KB
constexpr int kMaskBytes = kBM * 16;            // per stage: 128 rows x 2 u64 (the element masks of an aligned tile pair)
constexpr int kStageBytes = 2 * kTileBytes + kMaskBytes;
constexpr int kMaxTileWords = 64;  // activity bitmap words -> T_SRC <= 64 * 32 * 64 = 131072
template <typename T16>
__device__ __forceinline__ uint32_t pack2b(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2b<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}
template <>
__device__ __forceinline__ uint32_t pack2b<__half>(float a, float b) {
    __half2 p = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}
template <typename T16>
__device__ __forceinline__ void unpack2b(uint32_t w, float& lo, float& hi);
template <>
__device__ __forceinline__ void unpack2b<__nv_bfloat16>(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
template <>
__device__ __forceinline__ void unpack2b<__half>(uint32_t w, float& lo, float& hi) {
    const __half2 h2 = *reinterpret_cast<const __half2*>(&w);
    lo = __low2float(h2);
    hi = __high2float(h2);
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ void ldsm(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
template <typename T16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T16>
__global__ void __launch_bounds__(kBThreads, 2)
block_attention_bits_kernel(const uint32_t* __restrict__ tile_act, int act_words,
                            const T16* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                            const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                            const __grid_constant__ CUtensorMap tmap_m,
                            const float* __restrict__ scales, const T16* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
                            T16* __restrict__ out, int N, int H, int T_DST, int T_SRC, int is_causal, int n_row_blocks, int max_tiles) {
    extern __shared__ uint8_t bsm_raw[];
    uint8_t* bsm = bsm_raw + ((1024u - ((uint32_t) __cvta_generic_to_shared(bsm_raw) & 1023u)) & 1023u);      // (offset from the __shared__ array, not an integer round trip: keeps the shared address space -> LDS / STS, not generic LD / ST)   // swizzle atoms are 1 KB
    uint8_t* kv = bsm;                                                               // [stages][K 8 KB | V 8 KB | masks 2 KB]
    uint32_t* sact = reinterpret_cast<uint32_t*>(kv + kStages * kStageBytes);     // [kMaxTileWords]
    uint64_t* full = reinterpret_cast<uint64_t*>(sact + kMaxTileWords);              // [stages]
    int* rel_cnt = reinterpret_cast<int*>(full + kStages);                           // [stages] warps that released the stage
    uint16_t* slist = reinterpret_cast<uint16_t*>(rel_cnt + kStages);                // [max_tiles]
    __shared__ int s_nact;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // heavy (late) row blocks first: a causal row block owns the most tiles
    const int rb = n_row_blocks - 1 - (int) (blockIdx.x / (unsigned) (N * H));
    const int nh = (int) (blockIdx.x % (unsigned) (N * H));
    const int n = nh / H, h = nh % H;
    const int r0 = rb * kBM;

    // ---- set-up: barriers, tile activity of this (head, row block) as left by expand_mask_kernel ---------------------------
    if (tid < kMaxTileWords)
        sact[tid] = tid < act_words ? __ldg(tile_act + (((int64_t) n * H + h) * n_row_blocks + rb) * act_words + tid) : 0u;
    if (tid == 0) {
        umma::prefetch_tensormap(&tmap_k); umma::prefetch_tensormap(&tmap_v); umma::prefetch_tensormap(&tmap_m);
        for (int s = 0; s < kStages; ++s) { umma::mbar_init(&full[s], 1); rel_cnt[s] = 0; }
        umma::fence_barrier_init();
    }
    __syncthreads();
    if (warp == 0) {
        // compact the active tile ids (ascending) into slist
        const int nwords = (max_tiles + 31) >> 5;
        int base = 0;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
            const uint32_t word = (w0 + lane) < nwords ? sact[w0 + lane] : 0u;
            const int pc = __popc(word);
            const int incl = warp_scan_incl_i(pc, lane);
            int pos = base + incl - pc;
            for (uint32_t x = word; x; x &= x - 1) slist[pos++] = (uint16_t) (((w0 + lane) << 5) + __ffs(x) - 1);
            base += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) s_nact = base;
    }
    __syncthreads();
    const int nact = s_nact;

    // K / V tile of 64 source tokens -> one stage, by TMA (rows past T_SRC are zero-filled by the box)
    auto issue_tile = [&](int j) {
        const int s = j % kStages;
        uint8_t* dst = kv + s * kStageBytes;
        const int tile = (int) slist[j];
        umma::mbar_arrive_expect_tx(&full[s], kStageBytes);
        umma::tma_load_4d(dst, &tmap_k, &full[s], 0, tile * kBN, h, n);
        umma::tma_load_4d(dst + kTileBytes, &tmap_v, &full[s], 0, tile * kBN, h, n);
        umma::tma_load_3d(dst + 2 * kTileBytes, &tmap_m, &full[s], (tile & ~1) * 2, r0, n * H + h);    // u32 columns of the tile pair
    };
    if (tid == 0)
        for (int j = 0; j < kStages && j < nact; ++j) issue_tile(j);

    // ---- per-warp state ---------------------------------------------------------------------------------------------
    const int g = lane >> 2, tq = lane & 3;
    // Q as the A operand, straight from global memory (read once per CTA)
    uint32_t qa[kBD / 16][4];
    {
        const int t0 = r0 + warp * 16 + g, t1 = t0 + 8;
        const uint32_t* q0 = reinterpret_cast<const uint32_t*>(q + (int64_t) n * q_sn + (int64_t) h * q_sh + (int64_t) min(t0, T_DST - 1) * q_st);
        const uint32_t* q1 = reinterpret_cast<const uint32_t*>(q + (int64_t) n * q_sn + (int64_t) h * q_sh + (int64_t) min(t1, T_DST - 1) * q_st);
#pragma unroll
        for (int ks = 0; ks < kBD / 16; ++ks) {
            qa[ks][0] = t0 < T_DST ? __ldg(q0 + ks * 8 + tq) : 0u;
            qa[ks][1] = t1 < T_DST ? __ldg(q1 + ks * 8 + tq) : 0u;
            qa[ks][2] = t0 < T_DST ? __ldg(q0 + ks * 8 + 4 + tq) : 0u;
            qa[ks][3] = t1 < T_DST ? __ldg(q1 + ks * 8 + 4 + tq) : 0u;
        }
    }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, nms[2] = {0.f, 0.f};     // nms = -m_run * log2(e)
    float acc[kBD / 8][4];
#pragma unroll
    for (int nt = 0; nt < kBD / 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kLazy = 8.0f / kLog2e;          // 2^8 in raw score units
    const uint32_t kv_base = umma::smem_u32(kv);
    // ldmatrix lane offsets inside a swizzled [64 x 128 B] tile: row * 128 + ((chunk ^ (row & 7)) << 4); row & 7 == lane & 7
    const uint32_t k_row_off = (uint32_t) ((lane & 7) + 8 * (lane >> 4)) * 128u, k_chunk = (uint32_t) ((lane >> 3) & 1);
    const uint32_t v_row_off = (uint32_t) ((lane & 7) + 8 * ((lane >> 3) & 1)) * 128u, v_chunk = (uint32_t) (lane >> 4);
    const uint32_t swz = (uint32_t) (lane & 7);

    const uint32_t full_addr = umma::smem_u32(full);
    const uint32_t mrow_off = (uint32_t) (warp * 16 + g) * 16u;      // this thread's rows g, g + 8 inside a stage's mask block
    int old_pending = -1, pending_it = 0;                            // deferred "was I the last releaser" check (lane 0)

    for (int it = 0; it < nact; ++it) {
        const int stage = it % kStages;
        const int tile = slist[it];
        const uint8_t* st_ptr = kv + stage * kStageBytes;
        const uint32_t ks_addr = kv_base + (uint32_t) stage * (uint32_t) kStageBytes;
        const uint32_t vs_addr = ks_addr + kTileBytes;
        if (lane == 0 && old_pending == kBWarps - 1) {
            // I was the last of the 8 warps to release the stage of tile pending_it: refill it (no warp ever waits to produce)
            rel_cnt[pending_it % kStages] = 0;
            __threadfence_block();
            if (pending_it + kStages < nact) issue_tile(pending_it + kStages);
        }
        old_pending = -1;
        umma::mbar_wait_addr(full_addr + 8u * (uint32_t) stage, (uint32_t) ((it / kStages) & 1));
        const unsigned long long mk0 = *reinterpret_cast<const unsigned long long*>(st_ptr + 2 * kTileBytes + mrow_off + (tile & 1) * 8);
        const unsigned long long mk1 = *reinterpret_cast<const unsigned long long*>(st_ptr + 2 * kTileBytes + mrow_off + 128 + (tile & 1) * 8);
        const uint32_t mlo0 = (uint32_t) mk0, mhi0 = (uint32_t) (mk0 >> 32), mlo1 = (uint32_t) mk1, mhi1 = (uint32_t) (mk1 >> 32);
        const uint32_t anyl = mlo0 | mlo1, anyh = mhi0 | mhi1;
        const uint32_t act = __reduce_or_sync(kFull, ((anyl & 0xffffu) ? 1u : 0u) | ((anyl >> 16) ? 2u : 0u) | ((anyh & 0xffffu) ? 4u : 0u) | ((anyh >> 16) ? 8u : 0u));
        if (act != 0u) {
            float sc[kBN / 8][4];
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
                if (act & (1u << cg)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { sc[2 * cg][i] = 0.f; sc[2 * cg + 1][i] = 0.f; }
#pragma unroll
                    for (int ks = 0; ks < kBD / 16; ++ks) {
                        uint32_t b[4];
                        // rows (source tokens) 16cg + 0..15; matrices: (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
                        ldsm(b, ks_addr + cg * 2048u + k_row_off + (((uint32_t) (ks * 2) + k_chunk) ^ swz) * 16u);
                        mma16816<T16>(sc[2 * cg], qa[ks], b[0], b[1]);
                        mma16816<T16>(sc[2 * cg + 1], qa[ks], b[2], b[3]);
                    }
                }
            }
            // masked raw scores and the tile's row maxima; the words are pre-shifted so that every test uses an immediate
            const uint32_t sh2 = 2u * (uint32_t) tq;
            const uint32_t r0lo = mlo0 >> sh2, r0hi = mhi0 >> sh2, r1lo = mlo1 >> sh2, r1hi = mhi1 >> sh2;
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
                if (act & (1u << cg)) {
                    const uint32_t w0 = cg < 2 ? r0lo : r0hi, w1 = cg < 2 ? r1lo : r1hi;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int nt = 2 * cg + j;
                        constexpr uint32_t one = 1u;
                        const uint32_t b0 = one << ((nt & 3) * 8), b1 = b0 << 1;
                        sc[nt][0] = (w0 & b0) ? sc[nt][0] : -INFINITY;
                        sc[nt][1] = (w0 & b1) ? sc[nt][1] : -INFINITY;
                        sc[nt][2] = (w1 & b0) ? sc[nt][2] : -INFINITY;
                        sc[nt][3] = (w1 & b1) ? sc[nt][3] : -INFINITY;
                        mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
                        mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
                    }
                }
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 2));
            // lazy rescaling: the reference maximum of a row only moves when the tile exceeds it by more than 2^8
            // (probabilities stay <= 256, exact after the final normalisation because the same reference is used throughout)
            if (__any_sync(kFull, mx0 > m_run[0] + kLazy || mx1 > m_run[1] + kLazy)) {
                const float mn0 = mx0 > m_run[0] + kLazy ? mx0 : m_run[0], mn1 = mx1 > m_run[1] + kLazy ? mx1 : m_run[1];
                const float al0 = mn0 == -INFINITY ? 1.f : ex2f((m_run[0] - mn0) * kLog2e);       // m_run = -inf -> 0
                const float al1 = mn1 == -INFINITY ? 1.f : ex2f((m_run[1] - mn1) * kLog2e);
                m_run[0] = mn0; m_run[1] = mn1;
                nms[0] = mn0 == -INFINITY ? 0.f : -mn0 * kLog2e;
                nms[1] = mn1 == -INFINITY ? 0.f : -mn1 * kLog2e;
                l_run[0] *= al0; l_run[1] *= al1;
#pragma unroll
                for (int nt = 0; nt < kBD / 8; ++nt) { acc[nt][0] *= al0; acc[nt][1] *= al0; acc[nt][2] *= al1; acc[nt][3] *= al1; }
            }
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) {
                if (act & (1u << cg)) {
                    uint32_t pa[4];
                    {
                        const float p00 = ex2f(fmaf(sc[2 * cg][0], kLog2e, nms[0])), p01 = ex2f(fmaf(sc[2 * cg][1], kLog2e, nms[0]));
                        const float p02 = ex2f(fmaf(sc[2 * cg][2], kLog2e, nms[1])), p03 = ex2f(fmaf(sc[2 * cg][3], kLog2e, nms[1]));
                        const float p10 = ex2f(fmaf(sc[2 * cg + 1][0], kLog2e, nms[0])), p11 = ex2f(fmaf(sc[2 * cg + 1][1], kLog2e, nms[0]));
                        const float p12 = ex2f(fmaf(sc[2 * cg + 1][2], kLog2e, nms[1])), p13 = ex2f(fmaf(sc[2 * cg + 1][3], kLog2e, nms[1]));
                        ps0 += (p00 + p01) + (p10 + p11);
                        ps1 += (p02 + p03) + (p12 + p13);
                        pa[0] = pack2b<T16>(p00, p01); pa[1] = pack2b<T16>(p02, p03);
                        pa[2] = pack2b<T16>(p10, p11); pa[3] = pack2b<T16>(p12, p13);
                    }
#pragma unroll
                    for (int np = 0; np < kBD / 16; ++np) {
                        uint32_t b[4];
                        ldsm_t(b, vs_addr + cg * 2048u + v_row_off + (((uint32_t) (np * 2) + v_chunk) ^ swz) * 16u);
                        mma16816<T16>(acc[2 * np], pa, b[0], b[1]);
                        mma16816<T16>(acc[2 * np + 1], pa, b[2], b[3]);
                    }
                }
            }
            l_run[0] += ps0;
            l_run[1] += ps1;
        }
        __syncwarp();                                   // every lane is done reading the stage
        if (lane == 0) {
            __threadfence_block();
            old_pending = atomicAdd(&rel_cnt[stage], 1);      // consumed at the top of the next iteration: the round trip is hidden
            pending_it = it;
        }
    }
    if (lane == 0 && old_pending == kBWarps - 1) rel_cnt[pending_it % kStages] = 0;

    // ---- epilogue: normalise, * sigmoid(s0), mix with the running mean, permuted store ----------------------------------
    l_run[0] += __shfl_xor_sync(kFull, l_run[0], 1); l_run[0] += __shfl_xor_sync(kFull, l_run[0], 2);
    l_run[1] += __shfl_xor_sync(kFull, l_run[1], 1); l_run[1] += __shfl_xor_sync(kFull, l_run[1], 2);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int t = r0 + warp * 16 + g + 8 * rr;
        if (t >= T_DST) continue;
        const float inv = l_run[rr] > 0.f ? 1.0f / l_run[rr] : 0.f;
        const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
        const float psc = use_scaler ? sigm(sp[0]) : 1.0f;
        const float a = sigm(sp[1]);
        T16* orow = out + ((int64_t) n * T_DST + t) * ((int64_t) H * kBD) + (int64_t) h * kBD;
        const T16* arow = cumavg ? cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st : nullptr;
#pragma unroll
        for (int nt = 0; nt < kBD / 8; ++nt) {
            const int dd = nt * 8 + 2 * tq;
            float c0 = acc[nt][2 * rr] * inv * psc, c1 = acc[nt][2 * rr + 1] * inv * psc;
            if (arow) {
                float a0, a1;
                unpack2b<T16>(__ldg(reinterpret_cast<const uint32_t*>(arow + dd)), a0, a1);
                c0 = c0 * a + (1.0f - a) * a0;
                c1 = c1 * a + (1.0f - a) * a1;
            }
            *reinterpret_cast<uint32_t*>(orow + dd) = pack2b<T16>(c0, c1);
        }
    }
}

// a8 in dense, bit-packed form: dmask[n][h][t][w] (u64) holds the alive source tokens 64w .. 64w+63 of query row t, i.e.
// the reference's partial_attention_mask (attention.py:1025-1042) at one bit per element.  CTA = one query row with all
// its heads; thread = one 32-pixel word of the top-k bit mask; every alive pixel ORs its token run into the row image in
// shared memory, which is then written out with coalesced stores.  Only the words a 128-row query block can see are
// written (causal).
__global__ void __launch_bounds__(256)
expand_mask_kernel(const uint32_t* __restrict__ mask_bits, unsigned long long* __restrict__ dmask, uint32_t* __restrict__ tile_act, int act_words, int W64,
                   int N, int H, int T_DST, int T_SRC, int P, int p_lg, int is_causal) {
    extern __shared__ uint32_t ex_sm[];                   // [H][2 * W64] row image | [H][act_words] | [8 warps][1024] pixel lists
    pdl_launch_dependents();
    pdl_wait();
    const int nw = P >> 5;
    const int row = blockIdx.x, n = row / T_DST, t = row % T_DST;
    const int src_off = is_causal ? (T_SRC - T_DST) : 0;
    RowScale rs;
    rs.L = is_causal ? (src_off + t + 1) : T_SRC; rs.lg = p_lg; rs.halfP = P >> 1;
    rs.s = __fdiv_rn((float) rs.L, (float) P);
    const int blk_end = is_causal ? src_off + min((t / kBM + 1) * kBM, T_DST) : T_SRC;
    const int wneed = min(W64, (blk_end + 63) >> 6);      // u64 words per head row
    for (int i = threadIdx.x; i < H * 2 * wneed; i += blockDim.x) ex_sm[i] = 0u;
    __syncthreads();
    const uint32_t* brow = mask_bits + (int64_t) row * ((int64_t) H * nw);
    // Alive pixels are first compacted into a per-warp list (scan of the words' popcounts), then processed one pixel per lane:
    // a per-thread loop over the set bits of its own word runs at the pace of the fullest word of the warp (ncu: 7 of 32 lanes
    // active on the edge arithmetic and the atomics, which are most of this kernel's instructions).
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    uint16_t* lst = reinterpret_cast<uint16_t*>(ex_sm + H * 2 * W64 + H * act_words) + wrp * 1024;      // [1024] (head << 10) | pixel  (P <= 1024, H <= 64)
    for (int base = 0; base < H * nw; base += blockDim.x) {
        const int idx = base + threadIdx.x;
        uint32_t x = 0u;
        int h = 0, w = 0;
        if (idx < H * nw) { x = __ldg(brow + idx); h = idx / nw; w = idx - h * nw; }
        const int cnt = __popc(x);
        const int incl = warp_scan_incl_i(cnt, lane);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int off = incl - cnt;
        for (; x; x &= x - 1) lst[off++] = (uint16_t) ((h << 10) | ((w << 5) + __ffs(x) - 1));
        __syncwarp();
        for (int i = lane; i < total; i += 32) {
            const uint32_t e = lst[i];
            const int m = (int) (e & 1023u);
            uint32_t* img = ex_sm + (e >> 10) * 2 * wneed;
            const int a = rs.edge(m), b = rs.edge(m + 1);
            if (b > a)
                for (int wd = a >> 5; wd <= ((b - 1) >> 5); ++wd) {
                    const int lo = max(a - (wd << 5), 0), hi = min(b - (wd << 5), 32);
                    atomicOr(img + wd, (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo);
                }
        }
        __syncwarp();
    }
    __syncthreads();
    // tile activity of the (head, 128-row query block): which 64-token tiles hold an alive element of any of its rows.
    // Collected per CTA in shared memory first, so that a row issues at most H * act_words global atomics.
    uint32_t* act_sm = ex_sm + H * 2 * W64;                                   // [H][act_words]
    for (int i = threadIdx.x; i < H * act_words; i += blockDim.x) act_sm[i] = 0u;
    __syncthreads();
    for (int h = wrp; h < H; h += (int) (blockDim.x >> 5)) {                  // warp per head: no division in the write-out loop
        unsigned long long* drow = dmask + (((int64_t) n * H + h) * T_DST + t) * W64;
        for (int w = lane; w < wneed; w += 32) {
            const uint2 v = *reinterpret_cast<const uint2*>(ex_sm + 2 * (h * wneed + w));
            drow[w] = (unsigned long long) v.x | ((unsigned long long) v.y << 32);
            if ((v.x | v.y) != 0u) atomicOr(act_sm + h * act_words + (w >> 5), 1u << (w & 31));
        }
    }
    __syncthreads();
    uint32_t* act_blk = tile_act + (int64_t) n * H * ((T_DST + kBM - 1) / kBM) * act_words + (int64_t) (t / kBM) * act_words;
    const int64_t act_hs = (int64_t) ((T_DST + kBM - 1) / kBM) * act_words;
    for (int i = threadIdx.x; i < H * act_words; i += blockDim.x) {
        const uint32_t bits = act_sm[i];
        if (bits != 0u) {
            uint32_t* aw = act_blk + (i / act_words) * act_hs + (i % act_words);
            if ((__ldcg(aw) & bits) != bits) atomicOr(aw, bits);
        }
    }
}

}  // namespace

bool block_attention_eligible(int D, int T_SRC, int P, int k_clamp) {
    // no pixel may be clamped (span <= ceil(L/P) + 1 <= k), and the O(T^2)-worst-case tiling must still pay off
    return D == kBD && (P % 32) == 0 && P <= 1024 && (T_SRC + P - 1) / P + 1 <= k_clamp && T_SRC <= 8192;
}

static_assert(kBM == kMaskRowBlock && kBN == kMaskTile, "mask layout constants");

}  // namespace sea

using namespace sea;

extern "C" {

// bf16 up to T_SRC = 4096 runs the tcgen05 kernel, which interpolates the mask itself from the top-k bits (no dense mask, no
// expansion kernel); fp16 and longer rows run the mma.sync kernel over the dense bit mask written by expand_mask_kernel.
static bool use_umma_kernel(int dtype, int T_SRC, int P) {
    static const bool force_mma_sync = getenv("SEA_ATTN_MMA_SYNC") != nullptr;      // A/B timing switch
    // (its in-kernel interpolation uses the exact integer pixel edges: P a power of two, <= 512)
    return dtype == SEA_DTYPE_BF16 && !force_mma_sync && T_SRC <= 4096 && P <= 512 && exact_edge_shift(P, T_SRC) >= 0;      // measured: mma.sync ahead at T = 8192
}

int64_t sea_block_attention_workspace_bytes(int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int dtype) {
    if (dtype != SEA_DTYPE_BF16 && dtype != SEA_DTYPE_F16) return 0;
    if (N <= 0 || H <= 0 || H > 64 || T_DST <= 0 || T_SRC < T_DST || !block_attention_eligible(D, T_SRC, P, k_clamp)) return 0;      // (H <= 64: 6-bit head field of the expansion's pixel list)
    if (use_umma_kernel(dtype, T_SRC, P)) return 16;                                                          // no workspace use; > 0 = "supported"
    return (int64_t) N * H * T_DST * mask_row_words(T_SRC) * 8 + mask_act_bytes(N, H, T_DST, T_SRC);      // dense bit mask + tile activity
}

int sea_block_attention_fwd(const uint32_t* mask_bits,
                            const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                            const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                            const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                            const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype, void* out,
                            int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal,
                            void* workspace, int64_t workspace_bytes, void* stream) {
    SEA_CHECK_ARG(q && k && v && scales && out && workspace, "sea_block_attention_fwd: null pointer");
    const int64_t need = sea_block_attention_workspace_bytes(N, H, T_DST, T_SRC, D, P, k_clamp, dtype);
    if (need == 0) {
        set_error("sea_block_attention_fwd: unsupported shape (needs 16-bit activations, D = 64, P %% 32 == 0, no clamped pixel: "
                  "ceil(T_SRC / P) + 1 <= k, T_SRC <= 8192); use sea_sparse_attention_bits_fwd");
        return SEA_ERR_UNSUPPORTED;
    }
    SEA_CHECK_ARG(workspace_bytes >= need, "sea_block_attention_fwd: workspace too small (%lld < %lld)", (long long) workspace_bytes, (long long) need);
    SEA_CHECK_ARG(((q_sn | q_sh | q_st | k_sn | k_sh | k_st | v_sn | v_sh | v_st) % 8) == 0 &&
                  ((((uintptr_t) q) | ((uintptr_t) k) | ((uintptr_t) v) | ((uintptr_t) out) | ((uintptr_t) cumavg) | ((uintptr_t) workspace)) & 15) == 0,
                  "sea_block_attention_fwd: rows must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t) stream;
    const int p_lg = exact_edge_shift(P, T_SRC);      // integer pixel edges are exact iff P is a power of two and m * L stays below 2^24
    if (use_umma_kernel(dtype, T_SRC, P)) {
        SEA_CHECK_ARG(mask_bits != nullptr, "sea_block_attention_fwd: the tcgen05 kernel needs the top-k bit mask");
        return launch_block_attention_umma(mask_bits, P, p_lg, q, q_sn, q_sh, q_st, k, k_sn, k_sh, k_st, v, v_sn, v_sh, v_st, scales, cumavg,
                                           avg_sh, avg_st, use_scaler, out, N, H, T_DST, T_SRC, is_causal, s);
    }
    const int W64 = mask_row_words(T_SRC);
    unsigned long long* dmask = reinterpret_cast<unsigned long long*>(workspace);
    uint32_t* tile_act = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(workspace) + (int64_t) N * H * T_DST * W64 * 8);
    const int act_words = mask_act_words(T_SRC);
    SEA_CHECK_ARG(mask_bits != nullptr, "sea_block_attention_fwd: null mask_bits");
    {
        SEA_CUDA_TRY(cudaMemsetAsync(tile_act, 0, (size_t) mask_act_bytes(N, H, T_DST, T_SRC), s), "memset tile activity");
        const size_t smem = (size_t) H * W64 * 8 + (size_t) H * act_words * 4 + (size_t) 8 * 1024 * 2;      // row image + tile activity + per-warp pixel lists
        SEA_CHECK_ARG(smem <= 200 * 1024, "sea_block_attention_fwd: H * T_SRC too large for the mask expansion");
        SEA_CUDA_TRY(cudaFuncSetAttribute(expand_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        SEA_CUDA_TRY(launch_pdl(expand_mask_kernel, dim3((unsigned) ((int64_t) N * T_DST)), dim3(256), (size_t) smem, s, mask_bits, dmask, tile_act, act_words, W64, N, H, T_DST,
                                T_SRC, P, p_lg, is_causal), "expand_mask_kernel launch");
        SEA_CHECK_LAUNCH("expand_mask_kernel");
    }
    // fp16, T_SRC > 4096 (and SEA_ATTN_MMA_SYNC=1, for A/B timing): mma.sync over the expanded mask
    const int n_row_blocks = (T_DST + kBM - 1) / kBM;
    const int max_tiles = (T_SRC + kBN - 1) / kBN;
    CUtensorMap t_k, t_v;
    {
        const uint64_t dims[4] = {(uint64_t) kBD, (uint64_t) T_SRC, (uint64_t) H, (uint64_t) N};
        const uint32_t box[4] = {(uint32_t) kBD, (uint32_t) kBN, 1, 1};
        const uint64_t ks[3] = {(uint64_t) k_st * 2, (uint64_t) k_sh * 2, (uint64_t) k_sn * 2};
        const uint64_t vs[3] = {(uint64_t) v_st * 2, (uint64_t) v_sh * 2, (uint64_t) v_sn * 2};
        int rc = make_tmap_bf16_sw128(&t_k, const_cast<void*>(k), 4, dims, ks, box);
        if (rc) return rc;
        rc = make_tmap_bf16_sw128(&t_v, const_cast<void*>(v), 4, dims, vs, box);
        if (rc) return rc;
    }
    CUtensorMap t_m;
    {
        const uint64_t dims[3] = {(uint64_t) W64 * 2, (uint64_t) T_DST, (uint64_t) N * H};
        const uint64_t str[2] = {(uint64_t) W64 * 8, (uint64_t) T_DST * W64 * 8};
        const uint32_t box[3] = {4, (uint32_t) kBM, 1};
        int rc = make_tmap_u32_plain(&t_m, dmask, 3, dims, str, box);
        if (rc) return rc;
    }
    const size_t smem = 1024 + (size_t) kStages * kStageBytes + kMaxTileWords * 4 + kStages * 8 + kStages * 4 + (size_t) ((max_tiles + 7) & ~7) * 2;
    const unsigned grid = (unsigned) ((int64_t) n_row_blocks * N * H);
#define SEA_BLOCK_ATTN(TT)                                                                                                       \
    do {                                                                                                                         \
        auto kern = block_attention_bits_kernel<TT>;                                                                             \
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");          \
        kern<<<grid, kBThreads, smem, s>>>(tile_act, act_words, (const TT*) q, q_sn, q_sh, q_st, t_k, t_v, t_m, scales, (const TT*) cumavg,    \
            avg_sh, avg_st, use_scaler, (TT*) out, N, H, T_DST, T_SRC, is_causal, n_row_blocks, max_tiles);                      \
    } while (0)
    if (dtype == SEA_DTYPE_BF16) SEA_BLOCK_ATTN(__nv_bfloat16); else SEA_BLOCK_ATTN(__half);
#undef SEA_BLOCK_ATTN
    SEA_CHECK_LAUNCH("block_attention_bits_kernel");
    return SEA_OK;
}

int64_t sea_debug_attn_trace_read(uint32_t* host, int64_t max_words) { return attn_trace_read(host, max_words); }

}  // extern "C"
