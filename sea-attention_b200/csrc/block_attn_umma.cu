// a8-a14 of the short-context path on the 5th-generation tensor cores (bf16, d = 64): same algorithm and results as
// block_attention_bits_kernel (block_attn.cu), but the two contractions of a tile are tcgen05.mma instructions with the
// accumulators in TMEM, and the a8 interpolation happens INSIDE the kernel, straight from the top-k pixel bits: no dense mask, no
// expansion kernel -- the only mask bytes the attention reads from HBM are the N*T*H*P/8 top-k bits.
//
// CTA = 256 query rows of one head = two 128-row Q tiles (A, B) that share every K/V tile (halves the L2 -> SM tile traffic of a
// 128-row CTA), 640 threads, 1 CTA / SM:
//   warp 0 (one lane)  TMA producer: both Q tiles once; per source tile one stage = K [64 x 128 B] | V [64 x 128 B] (SWIZZLE_128B
//                      boxes straight from the strided [N,H,T,d] tensors), 4-stage full / empty mbarrier ring
//   warp 1             MMA issuer (warp-converged, elect.sync inside the helpers), per source tile j and Q tile g:
//                        S_g[128 x 64]  = Q_g . K^T    (both K-major, 4 k-steps)                     -> TMEM S_g[j & 1]
//                        O_g[128 x 64] += P_g . V      (A = P K-major from shared memory, B = V as the MN-major operand)
//                        l_g[128 x 16] += P_g . 1      (B = a constant tile of ones): the softmax denominator comes out of the
//                                                       tensor core, consistent with the bf16-rounded P that P.V uses
//                      S of tile j + 1 is issued before P.V of tile j, so it overlaps the softmax of tile j
//   warps 4-19         softmax, TWO threads per query row (32 of the 64 tile columns each; Q tile g = (warp - 4) / 8):
//                        * mask: every 8 tiles ONE of the row's two threads (they alternate) turns the alive pixels of the row that
//                          fall into the next 512 source tokens into eight 64-bit element masks in shared memory (pixel cursor
//                          kept in shared memory; pixels and tokens both ascend; exact a8 edge arithmetic, csr_common.cuh::
//                          RowScale).  The loop is warp-uniform: one pixel (or one cursor step) per lane and iteration.
//                        * no running maximum: the reference exponent of a row is fixed BEFORE the loop from the score of the
//                          row's own position (q_t . k_t, half of the dot product per thread) plus a 2^32 head-room, so p = 2^(s - ref)
//                          needs no per-element max, no per-tile rescale and no per-tile exchange between the two halves of a
//                          row; every alive p is checked for p >= 2 through an OR of the packed bf16 results (bit 14).  Only then
//                          (a score e^22 above the reference: practically never) the tile's true maximum is taken, and at the next
//                          8-tile boundary -- where the two halves of a row meet at a named barrier anyway -- the reference moves
//                          and O / l are rescaled in TMEM (fp32 / bf16 hold p up to 2^127, so deferring is exact).
//                        * 8-column groups with no alive element in the warp's 32 rows are skipped (zeros stored)
//                        * P -> bf16 -> shared memory in the swizzled K-major layout tcgen05 reads
// TMEM (512 columns): S_A[2], S_B[2] (4 x 64), O_A, O_B (2 x 64), l_A, l_B (2 x 16).
// Reference: attention.py:1036-1042 (a8), 1151-1173, 1237-1244, 1279-1282; mask semantics causal_resize_m_to_t.py:648-762.
#include "common.cuh"
#include "csr_common.cuh"
#include "umma.cuh"
#include "block_attn.cuh"
#include <type_traits>

namespace sea {
namespace {

constexpr int kUM = 128;                 // query rows per Q tile
constexpr int kUN = 64;                  // source tokens per tile
constexpr int kUD = 64;                  // head dim
constexpr int kUStages = 4;             // per ring: K tiles and V tiles travel in separate rings (K is released after S, two tiles before V)
constexpr int kUTile = kUN * kUD * 2;    // 8 KB (K or V tile)
constexpr int kUQ = kUM * kUD * 2;       // 16 KB per Q tile
constexpr int kUP = kUM * kUN * 2;       // 16 KB per P buffer
constexpr int kUChunk = 8;               // tiles per mask-generation chunk
// G = Q tiles per CTA.  G = 2: 256 rows share every K/V tile (half the L2 -> SM traffic), 1 CTA / SM (512 TMEM columns, 640 threads).
// G = 1: 128 rows, 2 CTAs / SM (256 TMEM columns, 384 threads): the set-up / epilogue of one CTA runs under the other's main loop.
__host__ __device__ constexpr int urows(int G) { return G * kUM; }
__host__ __device__ constexpr int uthreads(int G) { return 128 + G * 8 * 32; }   // warp 0: TMA; warps 1 .. G: MMA issuer of Q tile g; warps 4 ..: softmax (Q tile, column half, TMEM lane quarter)
constexpr float kUMargin = 32.0f;        // head-room (log2) between a row's reference score and p = 1

template <int G>
struct USmem {
    static constexpr int kUG = G, kURows = G * kUM;
    static constexpr int kQ = 0;
    static constexpr int kKV = kQ + kUG * kUQ;
    static constexpr int kMw = kKV + 2 * kUStages * kUTile;                     // u64 [2 (chunk parity)][kUChunk][256 rows]
    static constexpr int kCur = kMw + 2 * kUChunk * kURows * 8;  // {int word, u32 remaining bits} [256 rows]: pixel cursor
    static constexpr int kXch = kCur + kURows * 8;               // float [2 (chunk parity)][2 halves][256 rows]
    static constexpr int kBar = kXch + 6 * kURows * 4;          // (+ [2 halves][256 rows] for the final row-sum exchange)
    static constexpr int kNumBars = 1 + 4 * kUStages + 4 * kUG * 2;
    static constexpr int kBits = kBar + kNumBars * 8 + 16;       // u32 [256 rows][ubits_stride(P / 32)]
};
__host__ __device__ constexpr int ubits_stride(int nw) { return nw | 1; }       // odd stride: row-strided reads hit distinct banks

// kind::f16 instruction descriptor, D fp32, A = B = bf16, A K-major, B MN-major (bit 16), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16_b_mn(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (bf16 pairs, one 32-bit TMEM column per two K elements, lane = row) never
// touches shared memory.  Warp-converged like umma::mma_bf16_ss_elect.
__device__ __forceinline__ void mma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x2(uint32_t taddr, const uint32_t (&r)[2]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2u(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigu(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float dot8_bf(uint4 a, uint4 b, float acc) {
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(bw[i] << 16), acc);
        acc = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(bw[i] & 0xffff0000u), acc);
    }
    return acc;
}
// explicit shared-window accesses (the dynamic shared memory base is re-aligned by hand, which hides the address space from the
// compiler: plain C++ accesses through it become generic LD / ST)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
constexpr float kLog2e = 1.4426950408889634f;


// eight scores (one 8-column group) -> four packed bf16 pairs of p = 2^(s * log2e + nms), dead elements exactly 0
template <int kBase>
__device__ __forceinline__ void exp_group(const uint32_t (&s)[32], uint32_t mword, float nms, uint4& out, float& l0, float& l1) {
    float p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float x = fmaf(__uint_as_float(s[kBase + i]), kLog2e, nms);
        p[i] = ex2u((mword & (1u << (kBase + i))) ? x : -INFINITY);
    }
    l0 += (p[0] + p[1]) + (p[2] + p[3]);
    l1 += (p[4] + p[5]) + (p[6] + p[7]);
    out = make_uint4(pack_bf(p[0], p[1]), pack_bf(p[2], p[3]), pack_bf(p[4], p[5]), pack_bf(p[6], p[7]));
}

// Development aid (SEA_ATTN_TRACE=1): per-CTA cycle counters of where the softmax and MMA warps spend their time, written by lane 0 of
// every warp into a global buffer [cta][warp][8] (read back with sea_debug_attn_trace_read; scripts/attn_trace.py prints the averages).
// Softmax warps: 0 wait S, 1 tcgen05.ld, 2 exp / pack, 3 wait P buffer, 4 tcgen05.st + hand-off, 5 chunk boundary + mask word, 6 set-up,
// 7 whole loop (6 = mask word + column union).  MMA warps: 0 wait S buffer, 1 issue S, 2 wait V, 3 wait P, 4 issue P.V, 5 wait K, 7 whole loop.
__device__ uint32_t* g_attn_trace = nullptr;
__device__ __forceinline__ uint32_t clock_lo() { return (uint32_t) clock64(); }

template <int kUG, bool kTrace = false>
__global__ void __launch_bounds__(uthreads(kUG), 3 - kUG)
block_attention_umma_kernel(const uint32_t* __restrict__ mask_bits, int P, int p_lg,
                            const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                            const __grid_constant__ CUtensorMap tmap_v,
                            const __nv_bfloat16* __restrict__ qg, int64_t q_sn, int64_t q_sh, int64_t q_st,
                            const __nv_bfloat16* __restrict__ kg, int64_t k_sn, int64_t k_sh, int64_t k_st,
                            const float* __restrict__ scales, const __nv_bfloat16* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
                            __nv_bfloat16* __restrict__ out, int N, int H, int T_DST, int T_SRC, int is_causal, int n_pairs) {
    constexpr int kURows = urows(kUG), kUThreads = uthreads(kUG), kUSoftmaxWarps = 8 * kUG;
    using USmem = USmem<kUG>;
    extern __shared__ uint8_t usm_raw[];
    uint8_t* sm = usm_raw + ((1024u - ((uint32_t) __cvta_generic_to_shared(usm_raw) & 1023u)) & 1023u);      // (offset from the __shared__ array, not an integer round trip: keeps the shared address space -> LDS / STS, not generic LD / ST)
    const uint32_t sm_a = umma::smem_u32(sm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + USmem::kBar);
    uint64_t* q_full = bars;                         // [1]
    uint64_t* k_full = bars + 1;                     // [stages]
    uint64_t* k_empty = k_full + kUStages;           // [stages]  (= S of the tile has completed, for every Q tile)
    uint64_t* v_full = k_empty + kUStages;           // [stages]
    uint64_t* v_empty = v_full + kUStages;           // [stages]  (= P.V of the tile has completed, for every Q tile)
    uint64_t* s_full = v_empty + kUStages;           // [g][b]
    uint64_t* p_full = s_full + kUG * 2;             // [g][b]
    uint64_t* p_empty = p_full + kUG * 2;            // [g][b]   (= P.V of the tile that used the buffer has completed)
    uint64_t* s_empty = p_empty + kUG * 2;           // [g][b]   (= the softmax warps have read S out of TMEM)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_empty + kUG * 2);

    // warp index through a shuffle: provably warp-uniform, which keeps the MMA issue loop in the uniform datapath
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    uint32_t t_entry = 0;
    if constexpr (kTrace) {
        t_entry = clock_lo();
        if (tid == 96 && g_attn_trace) {          // (warp 3, the idle warp, owns the CTA-level slot: SM id, start time in ns)
            uint32_t smid; uint64_t gt;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 0] = smid;
            g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 1] = (uint32_t) gt;
        }
    }
    pdl_launch_dependents();
    pdl_wait();                // the set-up below already reads the top-k bits written by the predecessor
    const int rbp = n_pairs - 1 - (int) (blockIdx.x / (unsigned) (N * H));       // heavy (late) row blocks first
    const int nh = (int) (blockIdx.x % (unsigned) (N * H));
    const int n = nh / H, h = nh % H;
    const int r0 = rbp * kURows;
    const int src_off = is_causal ? (T_SRC - T_DST) : 0;
    // source tiles each Q tile needs (causal: up to its last row's diagonal)
    int ntg[kUG];
#pragma unroll
    for (int g = 0; g < kUG; ++g) {
        const int rf = r0 + g * kUM;
        const int lim = rf >= T_DST ? 0 : (is_causal ? min(src_off + min(rf + kUM, T_DST), T_SRC) : T_SRC);
        ntg[g] = (lim + kUN - 1) / kUN;
    }
    const int nt = max(ntg[0], ntg[kUG - 1]);

    // ---- set-up -------------------------------------------------------------------------------------------------------
    const int nw = P >> 5, bst = ubits_stride(nw);
    // K runs kLag tiles ahead of V: S of tile j + 2 is issued while P.V is still at tile j, so K stages turn over earlier
    constexpr int kLag = 2;
    auto issue_kv = [&](int jj) {            // producer step jj: K tile jj and V tile jj - kLag (thread 0 only)
        if (jj < nt) {
            const int s = jj % kUStages;
            if (jj >= kUStages) umma::mbar_wait_sleep(&k_empty[s], (uint32_t) ((jj / kUStages - 1) & 1), 256);
            umma::mbar_arrive_expect_tx(&k_full[s], kUTile);
            umma::tma_load_4d(sm + USmem::kKV + s * kUTile, &tmap_k, &k_full[s], 0, jj * kUN, h, n);
        }
        const int j = jj - kLag;
        if (j >= 0) {
            const int s = j % kUStages;
            if (j >= kUStages) umma::mbar_wait_sleep(&v_empty[s], (uint32_t) ((j / kUStages - 1) & 1), 256);
            umma::mbar_arrive_expect_tx(&v_full[s], kUTile);
            umma::tma_load_4d(sm + USmem::kKV + (kUStages + s) * kUTile, &tmap_v, &v_full[s], 0, j * kUN, h, n);
        }
    };
    int jj0 = 0;
    if (tid == 0) {
        umma::prefetch_tensormap(&tmap_q); umma::prefetch_tensormap(&tmap_k); umma::prefetch_tensormap(&tmap_v);
        umma::mbar_init(q_full, 1);
        for (int s = 0; s < kUStages; ++s) {
            umma::mbar_init(&k_full[s], 1); umma::mbar_init(&k_empty[s], kUG); umma::mbar_init(&v_full[s], 1); umma::mbar_init(&v_empty[s], kUG);
        }
        for (int b = 0; b < kUG * 2; ++b) {
            umma::mbar_init(&s_full[b], 1); umma::mbar_init(&p_full[b], kUSoftmaxWarps / kUG); umma::mbar_init(&p_empty[b], 1);
            umma::mbar_init(&s_empty[b], kUSoftmaxWarps / kUG);
        }
        umma::fence_barrier_init();
        // The Q tiles and the first ring stages (no stage is re-used below kUStages steps: no waits) are requested right here, by the
        // thread that initialised their barriers, so that they are in flight under the pixel-bit copy, the TMEM allocation and the
        // block barrier below instead of starting after them.
        if (nt > 0) {
            umma::mbar_arrive_expect_tx(q_full, (uint32_t) ((ntg[0] > 0) + (kUG > 1 && ntg[kUG - 1] > 0)) * kUQ);
            if (ntg[0] > 0) umma::tma_load_4d(sm + USmem::kQ, &tmap_q, q_full, 0, r0, h, n);
            if (kUG > 1 && ntg[kUG - 1] > 0) umma::tma_load_4d(sm + USmem::kQ + kUQ, &tmap_q, q_full, 0, r0 + kUM, h, n);
            for (; jj0 < kUStages && jj0 < nt + kLag; ++jj0) issue_kv(jj0);
        }
    }
    // top-k pixel bits of this head's 256 query rows -> shared memory (rows past T_DST: no alive pixel); pixel cursors
    for (int i = tid; i < kURows * nw; i += kUThreads) {
        const int r = i / nw, w = i - r * nw, tr = r0 + r;
        const uint32_t v = tr < T_DST ? __ldg(mask_bits + (((int64_t) n * T_DST + tr) * H + h) * nw + w) : 0u;
        sts32(sm_a + USmem::kBits + (uint32_t) (r * bst + w) * 4, v);
        if (w == 0) sts64(sm_a + USmem::kCur + (uint32_t) r * 8, make_uint2(0u, v));
    }
    if (warp == 1) umma::tmem_alloc(tmem_ptr, 256 * kUG);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
#define SEA_STAMP(i) if constexpr (kTrace) { if (warp == 4 && lane == 0 && g_attn_trace) g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + (i)] = clock_lo() - t_entry; }
    SEA_STAMP(8)                 // bits in shared memory, barriers, TMEM
    // TMEM columns: S_g[b] at 128 g + 64 b, O_g at 128 G + 64 g, P_g[b] (bf16 pairs) at 192 G + 64 g + 32 b

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0 && nt > 0) {
            for (int jj = jj0; jj < nt + kLag; ++jj) issue_kv(jj);       // (steps below jj0 were issued before the block barrier)
        }
    } else if (warp >= 1 && warp <= kUG) {
        // ---------------------------------------------------------------- MMA issuer of Q tile g.  One warp per Q tile: a serial
        // uniform-datapath issue stream costs tens of cycles per MMA (descriptor arithmetic), and with both Q tiles on one warp that
        // stream -- not the tensor core, not the softmax -- set the pace of the whole CTA (measured).
        const int g0 = warp - 1, g1 = g0 + 1;
        const int nt_a = ntg[0], nt_b = ntg[kUG - 1];
        const uint32_t idesc_s = umma::make_idesc_bf16(kUM, kUN);          // A, B K-major
        const uint32_t idesc_o = idesc_bf16_b_mn(kUM, kUD);                // B = V, MN-major
        // descriptors of the k = 0 step; a k-step adds a constant to the 14-bit (address >> 4) field (no carry: smem < 256 KB)
        const uint64_t d_q0 = umma::make_desc_k_sw128(sm_a + USmem::kQ);
        const uint64_t d_kv = umma::make_desc_k_sw128(sm_a + USmem::kKV);
        uint32_t tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c0 = 0, cstart = 0;
#define SEA_TR(i) if constexpr (kTrace) { const uint32_t c1_ = clock_lo(); tr[i] += c1_ - c0; c0 = c1_; }
        auto issue_s = [&](int j) {
            const int s = j % kUStages, b = j & 1;
            umma::mbar_wait(&k_full[s], (uint32_t) ((j / kUStages) & 1));
            SEA_TR(5)
            umma::tc_fence_after();
            const uint64_t d_k = d_kv + (uint64_t) (s * (kUTile >> 4));
            for (int g = g0; g < g1; ++g) {
                if (j < (g ? nt_b : nt_a)) {
                    const uint64_t d_q = d_q0 + (uint64_t) (g * (kUQ >> 4));
#pragma unroll
                    for (int k = 0; k < kUD / 16; ++k)
                        umma::mma_bf16_ss_elect(tmem_base + 128 * g + 64 * b, d_q + 2 * k, d_k + 2 * k, idesc_s, (uint32_t) (k != 0));
                    umma::mma_commit_elect(&s_full[g * 2 + b]);
                }
            }
            // the K stage is free once S of every Q tile on it has completed; a warp whose Q tile does not use tile j still arrives, but
            // only after it has seen the tile land (the wait above), so it can never run a whole ring ahead of the other warp and
            // complete a later phase of the barrier on its own
            umma::mma_commit_elect(&k_empty[s]);
        };
        if (nt > 0) {          // all 32 lanes walk the loop; elect.sync inside the *_elect helpers picks the issuing lane
            umma::mbar_wait(q_full, 0);
            issue_s(0);
            if (nt > 1) issue_s(1);
        }
        if constexpr (kTrace) { cstart = c0 = clock_lo(); tr[5] = 0; }
        for (int j = 0; j < nt; ++j) {
            const int s = j % kUStages, b = j & 1;
            // scores run two tiles ahead: S of tile j + 2 goes into the buffer of tile j as soon as the softmax warps have pulled tile j
            // out of TMEM (s_empty, early in their work on tile j), not after their P of tile j is complete
            if (j + 2 < nt) {
                for (int g = g0; g < g1; ++g)
                    if (j + 2 < (g ? nt_b : nt_a)) umma::mbar_wait(&s_empty[g * 2 + b], (uint32_t) ((j >> 1) & 1));
                SEA_TR(0)
                issue_s(j + 2);
                SEA_TR(1)
            }
            umma::mbar_wait(&v_full[s], (uint32_t) ((j / kUStages) & 1));
            SEA_TR(2)
            for (int g = g0; g < g1; ++g) {
                if (j < (g ? nt_b : nt_a)) {
                    umma::mbar_wait(&p_full[g * 2 + b], (uint32_t) ((j >> 1) & 1));
                    SEA_TR(3)
                    umma::tc_fence_after();
                    const uint64_t d_v = d_kv + (uint64_t) (((kUStages + s) * kUTile) >> 4);
                    const uint32_t tP = tmem_base + 192 * kUG + 64 * g + 32 * b;
#pragma unroll
                    for (int k = 0; k < kUN / 16; ++k)       // 16 source tokens per step: A = 8 TMEM columns of bf16 pairs, B two 8-token groups of V (2048 B)
                        mma_bf16_ts_elect(tmem_base + 128 * kUG + 64 * g, tP + 8 * k, d_v + 128 * k, idesc_o, (uint32_t) ((j | k) != 0));
                    umma::mma_commit_elect(&p_empty[g * 2 + b]);
                }
            }
            umma::mma_commit_elect(&v_empty[s]);          // (same rule as k_empty: v_full was waited on above)
            SEA_TR(4)
        }
        if constexpr (kTrace) {
            tr[7] = clock_lo() - cstart;
            if (lane == 0 && g_attn_trace) for (int i = 0; i < 8; ++i) g_attn_trace[((int64_t) blockIdx.x * 20 + warp) * 16 + i] = tr[i];
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- softmax warps: two threads per query row
        const int g = (warp - 4) >> 3;
        const int half = ((warp - 4) >> 2) & 1;
        const int qd = warp & 3;                                  // TMEM lane quarter this warp may access
        const int row = qd * 32 + lane, grow = g * kUM + row;
        const int t = r0 + grow;
        const int my_nt = g ? ntg[kUG - 1] : ntg[0];
        const int bar_id = 1 + g * 4 + qd;                        // named barrier of the row's two threads' warps (64 threads)
        const uint32_t lane_addr = (uint32_t) (qd * 32) << 16;
        const uint32_t tS = tmem_base + 128 * g + 32 * half + lane_addr, tO = tmem_base + 128 * kUG + 64 * g + 32 * half + lane_addr;
        const uint32_t tP = tmem_base + 192 * kUG + 64 * g + 16 * half + lane_addr;
        const uint32_t a_s_full = umma::smem_u32(s_full + g * 2), a_p_full = umma::smem_u32(p_full + g * 2), a_p_empty = umma::smem_u32(p_empty + g * 2), a_s_empty = umma::smem_u32(s_empty + g * 2);
        // a8 edge arithmetic of this thread's query row
        RowScale rs;
        rs.L = t < T_DST ? (is_causal ? min(src_off + t + 1, T_SRC) : T_SRC) : 1; rs.lg = p_lg; rs.halfP = P >> 1;
        rs.s = __fdiv_rn((float) rs.L, (float) P);
        const uint32_t a_brow = sm_a + USmem::kBits + (uint32_t) (grow * bst) * 4;
        const uint32_t a_mw = sm_a + USmem::kMw + (uint32_t) grow * 8;            // + ((parity * kUChunk + i) * 256) * 8
        const uint32_t a_cur = sm_a + USmem::kCur + (uint32_t) grow * 8;
        const uint32_t a_xch = sm_a + USmem::kXch + (uint32_t) grow * 4;          // + ((parity * 2 + half) * 256) * 4

        // The shortest rows (L <= 320 tokens: up to P alive pixels, most of them empty) go token-parallel instead: lane = source
        // token, its pixel is the largest m with edge(m) <= c, i.e. m = floor(((c + 1) P - P/2 - 1) / L) for the exact integer edges
        // (edge(m) = (m L + P/2) >> lg), and a ballot of the pixels' alive bits IS the element mask word.  The warp walks its 32 rows.
        auto gen_tokens = [&]() {
            const int wrow0 = g * kUM + qd * 32;
            for (int rr = 0; rr < 32; ++rr) {
                const int tr = r0 + wrow0 + rr;
                if (tr >= T_DST) break;
                const int Lr = is_causal ? min(src_off + tr + 1, T_SRC) : T_SRC;
                const float inv_l = 1.0f / (float) Lr;
                const uint32_t a_b = sm_a + USmem::kBits + (uint32_t) ((wrow0 + rr) * bst) * 4;
                const uint32_t a_w0 = sm_a + USmem::kMw + (uint32_t) (wrow0 + rr) * 8;
                for (int w32 = 0; w32 < ((Lr + 31) >> 5); ++w32) {
                    const int cc = (w32 << 5) + lane;
                    const int x1 = (cc + 1) * P - (P >> 1) - 1;
                    int m = (int) ((float) x1 * inv_l);
                    if ((m + 1) * Lr <= x1) ++m;
                    if (m * Lr > x1) --m;
                    m = min(m, P - 1);
                    const bool alive = cc < Lr && ((lds32(a_b + (m >> 5) * 4) >> (m & 31)) & 1u);
                    const uint32_t word = __ballot_sync(kFull, alive);
                    if (lane == 0) sts32(a_w0 + (uint32_t) ((w32 >> 1) * kURows) * 8 + (w32 & 1) * 4, word);
                }
            }
        };
        // ---- small CTAs (<= 16 source tiles): all element masks are generated up front.  The row's two threads split its pixel
        // words (first / second half of the pixels) and OR their runs into the row's words with shared-memory reductions.
        // "small" CTA = one that holds rows too short for the pixel window of the lazy masks (32 P / L + 2 <= 31 bits needs L >= 1.11 P):
        // its first row decides (rows only get longer); such a CTA has at most (1.11 P + 256) / 64 <= 2 kUChunk source tiles for
        // P <= 512, and its masks are generated up front.  (Before the lazy masks every CTA of <= 16 tiles took this path: at the
        // north-star shape that was 4 of 16 row blocks, each spending 9 - 12 us in the up-front generation.)
        const int min_l = is_causal ? src_off + r0 + 1 : T_SRC;
        const bool small_cta = nt <= 2 * kUChunk && ((int64_t) min_l * 29 < (int64_t) 32 * P || P > 512);
        int cstar = -1;                          // source token whose score fixes the row's reference exponent
        if (small_cta && my_nt > 0) {
            const int warp_last = r0 + g * kUM + qd * 32 + 31;
            const bool by_token = p_lg >= 0 && (is_causal ? src_off + warp_last + 1 : T_SRC) <= 96;         // warp-uniform
#pragma unroll
            for (int i = 0; i < kUChunk; ++i) sts64(a_mw + (uint32_t) ((half * kUChunk + i) * kURows) * 8, make_uint2(0u, 0u));
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            if (by_token) {
                if (half == 0) gen_tokens();
            } else {
                const int w_end = half ? nw : (nw >> 1);
                int cur_w = half ? (nw >> 1) : 0;
                uint32_t cur_x = cur_w < w_end ? lds32(a_brow + cur_w * 4) : 0u;
                bool done = cur_w >= w_end;
                while (!done) {
                    if (cur_x == 0u) {
                        if (++cur_w >= w_end) done = true; else cur_x = lds32(a_brow + cur_w * 4);
                    } else {
                        const int m = (cur_w << 5) + __ffs(cur_x) - 1;
                        cur_x &= cur_x - 1;
                        const int pa = rs.edge(m), pb = rs.edge(m + 1);
                        for (int c32 = pa >> 5; c32 <= ((pb - 1) >> 5) && pb > pa; ++c32) {
                            const int l = max(pa - (c32 << 5), 0), hh = min(pb - (c32 << 5), 32);
                            const uint32_t bits = (0xffffffffu >> (32 - (hh - l))) << l;
                            asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a_mw + (uint32_t) ((c32 >> 1) * kURows) * 8 + (c32 & 1) * 4), "r"(bits) : "memory");
                        }
                    }
                }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        }
        // reference of the row's exponent: the SMALLER of its scores against source tokens 1 and 2.  ANY per-row constant is exact -- it
        // cancels in the normalisation, p only has to stay inside the bf16 / fp32 exponent range (alive scores up to 65 nats below and 22
        // above the reference; above that the p >= 2 test below moves it, which is exact but slow; below that p would underflow, which is
        // NOT detected).  Hence two probes and their minimum -- a reference that is too high needs BOTH probe tokens to score 65 nats above
        // every alive one -- and hence not token 0, the usual attention sink.  Needs no search for the first alive token (1.6 us of every
        // CTA's set-up) and no gather: the probe keys are two broadcast rows, q_t is read from the Q tile the TMA put in shared memory.
        if (t < T_DST && my_nt > 0) cstar = min(1, T_SRC - 1);
        SEA_STAMP(9)             // small CTAs: element masks of all tiles
        // reference exponent: score of that element (+ head-room), so that every alive p stays far below 2
        // (each of the row's two threads gathers HALF of the two rows -- the 512 threads' row gathers were 2.8 us of every CTA's set-up --
        // and the halves are added at the rendezvous that opens tile 0)
        float nms = 0.f;
        {
            float acc = 0.f, acc_b = 0.f;
            if (my_nt > 0) umma::mbar_wait(q_full, 0);               // the Q tiles have landed (warp-uniform: my_nt depends on the Q tile only)
            if (cstar >= 0) {
                const uint4* kp = reinterpret_cast<const uint4*>(kg + (int64_t) n * k_sn + (int64_t) h * k_sh + (int64_t) cstar * k_st) + half * (kUD / 16);
                const uint4* kp_b = reinterpret_cast<const uint4*>(kg + (int64_t) n * k_sn + (int64_t) h * k_sh + (int64_t) min(2, T_SRC - 1) * k_st) + half * (kUD / 16);
                // row `row` of Q tile g: 128 bytes, its 16-byte chunk c at position c ^ (row & 7) (SWIZZLE_128B)
                const uint32_t a_q = sm_a + USmem::kQ + (uint32_t) g * kUQ + (uint32_t) row * 128;
#pragma unroll
                for (int i = 0; i < kUD / 16; ++i) {
                    const int c = half * (kUD / 16) + i;
                    uint4 qv;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qv.x), "=r"(qv.y), "=r"(qv.z), "=r"(qv.w) : "r"(a_q + (uint32_t) ((c ^ (row & 7)) << 4)) : "memory");
                    acc = dot8_bf(qv, __ldg(kp + i), acc);
                    acc_b = dot8_bf(qv, __ldg(kp_b + i), acc_b);
                }
            }
            sts32(a_xch + (uint32_t) ((4 + half) * kURows) * 4, __float_as_uint(acc));       // (the slots of the final row-sum exchange)
            sts32(a_xch + (uint32_t) ((2 + half) * kURows) * 4, __float_as_uint(acc_b));     // (the parity-1 slots of the boundary exchange: first used at tile 8)
        }
        SEA_STAMP(10)            // q . k of the first alive element
        // epilogue inputs, requested early
        float sc0 = 0.f, sc1 = 0.f;
        if (t < T_DST) {
            const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
            sc0 = __ldg(sp); sc1 = __ldg(sp + 1);
        }
        // ---- large CTAs (every row longer than 1024 tokens: a pixel is >= 4 tokens wide): the element masks are produced tile by tile by
        // the thread that consumes them, from the row's pixel bitmap.  The first pixel that can reach into the thread's 32 columns of tile j,
        // m(j) = floor(((c0 + 1) P - P/2 - 1) / L), advances by a per-row constant (quotient / remainder of 64 P / L: no division in the
        // loop); a window of the next <= 32 P / L + 2 pixel bits comes out of two bitmap words with one funnel shift.  The window is
        // empty for ~90 % of the (row, tile) pairs (a row has k P / L alive pixels per head: 4 at L = 4096): those cost ~12 non-divergent
        // instructions; an alive pixel costs its two edges and a clipped run.  No shared-memory mask arrays, no generation burst at the
        // 8-tile boundaries (SEA_ATTN_TRACE: the chunked generation + the barrier imbalance it caused took 25 % of the softmax warps'
        // time, reading the mask words back another 9 %).
        int lz_m = 0, lz_r = 0, lz_q = 0, lz_s = 0;
        uint32_t lz_nb = 0u;
        if (my_nt > 0 && !small_cta) {
            const int step = kUN * P, x0 = (32 * half + 1) * P - (P >> 1) - 1;
            lz_q = step / rs.L; lz_s = step - lz_q * rs.L;
            lz_m = x0 / rs.L; lz_r = x0 - lz_m * rs.L;
            const int nb = min(31, (32 * P) / rs.L + 2);
            lz_nb = (1u << nb) - 1u;
        }
        float l_a = 0.f, l_b = 0.f;              // fp32 row sum of this thread's 32 columns (two chains)
        float trig_mx = -INFINITY;              // largest alive score (log2 domain) of the tiles that overflowed the head-room since the last boundary
        uint32_t tr[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c0 = 0, cstart = 0;
        if constexpr (kTrace) {
            cstart = c0 = clock_lo();
            if (warp == 4 && lane == 0 && g_attn_trace) g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 4] = cstart - t_entry;      // set-up
        }

        for (int j = 0; j < my_nt; ++j) {
            const int b = j & 1;
            const int c = j / kUChunk, jc = j % kUChunk;
            if (jc == 0) {
                // ---- 8-tile boundary: the row's two threads meet; masks of chunk c become visible, chunk c + 1 is generated by one
                // of them, and a pending reference move (rare) is applied by both
                sts32(a_xch + (uint32_t) (((c & 1) * 2 + half) * kURows) * 4, __float_as_uint(trig_mx));
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                const float mxb = fmaxf(trig_mx, __uint_as_float(lds32(a_xch + (uint32_t) (((c & 1) * 2 + (half ^ 1)) * kURows) * 4)));
                trig_mx = -INFINITY;
                if (j == 0 && cstar >= 0) {      // reference exponent from the two half dot products (same value in both threads of the row)
                    const float d0 = __uint_as_float(lds32(a_xch + (uint32_t) (4 * kURows) * 4)), d1 = __uint_as_float(lds32(a_xch + (uint32_t) (5 * kURows) * 4));
                    const float e0 = __uint_as_float(lds32(a_xch + (uint32_t) (2 * kURows) * 4)), e1 = __uint_as_float(lds32(a_xch + (uint32_t) (3 * kURows) * 4));
                    nms = -(fminf(d0 + d1, e0 + e1) * kLog2e + kUMargin);
                }
                if (__any_sync(kFull, mxb > -INFINITY)) {
                    const bool mv = mxb > -INFINITY;
                    const float nms_new = mv ? -(mxb + kUMargin) : nms;
                    const float al = mv ? ex2u(nms_new - nms) : 1.0f;
                    nms = nms_new;
                    // j >= kUChunk here; P.V of tile j - 1 must have completed before O / l are touched
                    umma::mbar_wait_addr(a_p_empty + 8u * ((j - 1) & 1), (uint32_t) (((j - 1) >> 1) & 1));
                    umma::tc_fence_after();
                    uint32_t o32[32];
                    umma::tmem_ld_32x32(tO, o32);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * al);
                    tmem_st_32x32(tO, o32);
                    l_a *= al; l_b *= al;
                    tmem_st_wait();
                }
            }
            SEA_TR(5)
            uint32_t mw;
            if (small_cta) {
                mw = lds32(a_mw + (uint32_t) (((c & 1) * kUChunk + jc) * kURows) * 8 + half * 4);
            } else {
                const int c0_ = j * kUN + 32 * half, c1_ = c0_ + 32;           // this thread's columns of the tile
                const int wq = lz_m >> 5;
                const uint32_t lo = wq < nw ? lds32(a_brow + wq * 4) : 0u, hi = wq + 1 < nw ? lds32(a_brow + wq * 4 + 4) : 0u;
                uint32_t win = __funnelshift_r(lo, hi, lz_m & 31) & lz_nb;       // bit i: pixel lz_m + i alive
                mw = 0u;
                while (win) {
                    const int mm = lz_m + __ffs(win) - 1;
                    win &= win - 1;
                    const int l = max(rs.edge(mm), c0_) - c0_, hh = min(rs.edge(mm + 1), c1_) - c0_;
                    if (hh > l) mw |= (0xffffffffu >> (32 - (hh - l))) << l;
                }
                lz_m += lz_q; lz_r += lz_s;                                      // first pixel of the next tile's columns
                if (lz_r >= rs.L) { lz_r -= rs.L; ++lz_m; }
            }
            const uint32_t um = __reduce_or_sync(kFull, mw);       // column union of the warp's 32 rows
            SEA_TR(6)
            umma::mbar_wait_addr(a_s_full + 8u * b, (uint32_t) ((j >> 1) & 1));
            SEA_TR(0)
            umma::tc_fence_after();
            uint4 pk[4];
            uint32_t orv = 0u;
            {
                uint32_t s0[32];
                umma::tmem_ld_32x32(tS + 64 * b, s0);
                umma::tmem_ld_wait();
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(a_s_empty + 8u * b);         // S[b] may be overwritten by the scores of tile j + 2
                SEA_TR(1)
                if (um & 0xffu) { exp_group<0>(s0, mw, nms, pk[0], l_a, l_b); orv |= pk[0].x | pk[0].y; orv |= pk[0].z | pk[0].w; } else pk[0] = make_uint4(0, 0, 0, 0);
                if (um & 0xff00u) { exp_group<8>(s0, mw, nms, pk[1], l_a, l_b); orv |= pk[1].x | pk[1].y; orv |= pk[1].z | pk[1].w; } else pk[1] = make_uint4(0, 0, 0, 0);
                if (um & 0xff0000u) { exp_group<16>(s0, mw, nms, pk[2], l_a, l_b); orv |= pk[2].x | pk[2].y; orv |= pk[2].z | pk[2].w; } else pk[2] = make_uint4(0, 0, 0, 0);
                if (um & 0xff000000u) { exp_group<24>(s0, mw, nms, pk[3], l_a, l_b); orv |= pk[3].x | pk[3].y; orv |= pk[3].z | pk[3].w; } else pk[3] = make_uint4(0, 0, 0, 0);
                const bool grow_ref = (orv & 0x40004000u) != 0u;         // some alive p >= 2 (or inf)
                if (__any_sync(kFull, grow_ref)) {
                    // rare: remember the tile's true maximum; the reference moves at the next boundary (p up to 2^127 is exact meanwhile)
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (mw & (1u << i)) mx = fmaxf(mx, __uint_as_float(s0[i]) * kLog2e);
                    if (grow_ref) trig_mx = fmaxf(trig_mx, mx);
                }
            }
            // P[b] (TMEM) is free once P.V of tile j - 2 has completed
            SEA_TR(2)
            if (j >= 2) umma::mbar_wait_addr(a_p_empty + 8u * b, (uint32_t) (((j >> 1) - 1) & 1));
            SEA_TR(3)
            umma::tc_fence_after();
            {
                const uint32_t p16[16] = {pk[0].x, pk[0].y, pk[0].z, pk[0].w, pk[1].x, pk[1].y, pk[1].z, pk[1].w,
                                          pk[2].x, pk[2].y, pk[2].z, pk[2].w, pk[3].x, pk[3].y, pk[3].z, pk[3].w};
                tmem_st_32x16(tP + 32 * b, p16);
                tmem_st_wait();
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(a_p_full + 8u * b);
            SEA_TR(4)
        }
        if constexpr (kTrace) {
            tr[7] = clock_lo() - cstart;
            if (lane == 0 && g_attn_trace) for (int i = 0; i < 8; ++i) g_attn_trace[((int64_t) blockIdx.x * 20 + warp) * 16 + i] = tr[i];
            if (warp == 4 && lane == 0 && g_attn_trace) g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 5] = clock_lo() - t_entry;           // end of the loop
        }
#undef SEA_TR
        // ---- epilogue: O / l, * sigmoid(s0), mix with the running mean, permuted store (32 channels per thread) -------------
        // row sum: the two halves of a row exchange their partial sums
        sts32(a_xch + (uint32_t) ((4 + half) * kURows) * 4, __float_as_uint(l_a + l_b));
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float l_run = (l_a + l_b) + __uint_as_float(lds32(a_xch + (uint32_t) ((4 + (half ^ 1)) * kURows) * 4));
        SEA_STAMP(11)            // row sums exchanged
        // the running-mean row of the mix is requested BEFORE the wait for the last P.V (four 16-byte loads in flight under it)
        uint4 av4[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        const bool has_avg = cumavg != nullptr && t < T_DST;
        if (has_avg) {
            const uint4* arow0 = reinterpret_cast<const uint4*>(cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st + 32 * half);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) av4[c8] = __ldg(arow0 + c8);
        }
        uint32_t o0[32];
        if (my_nt > 0) {
            umma::mbar_wait_addr(a_p_empty + 8u * ((my_nt - 1) & 1), (uint32_t) (((my_nt - 1) >> 1) & 1));
            umma::tc_fence_after();
            umma::tmem_ld_32x32(tO, o0);
            umma::tmem_ld_wait();
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o0[i] = 0u;
        }
        SEA_STAMP(12)            // last P.V complete, O in registers
        if (t < T_DST) {
            const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
            const float psc = use_scaler ? sigu(sc0) : 1.0f;
            const float a = sigu(sc1);
            const float w = inv * psc;
            __nv_bfloat16* orow = out + ((int64_t) n * T_DST + t) * ((int64_t) H * kUD) + (int64_t) h * kUD + 32 * half;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = l_run > 0.f ? __uint_as_float(o0[c8 * 8 + i]) * w : 0.f;
                if (has_avg) {
                    const uint4 av = av4[c8];
                    const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        x[2 * i] = x[2 * i] * a + (1.0f - a) * __uint_as_float(aw[i] << 16);
                        x[2 * i + 1] = x[2 * i + 1] * a + (1.0f - a) * __uint_as_float(aw[i] & 0xffff0000u);
                    }
                }
                *reinterpret_cast<uint4*>(orow + c8 * 8) = make_uint4(pack_bf(x[0], x[1]), pack_bf(x[2], x[3]), pack_bf(x[4], x[5]), pack_bf(x[6], x[7]));
            }
        }
    }
    if constexpr (kTrace) {
        if (warp == 4 && lane == 0 && g_attn_trace) g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 6] = clock_lo() - t_entry;               // end of the epilogue
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, 256 * kUG);
    }
    if constexpr (kTrace) {
        if (tid == 96 && g_attn_trace) {
            uint64_t gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 2] = (uint32_t) gt;
            g_attn_trace[((int64_t) blockIdx.x * 20 + 3) * 16 + 7] = clock_lo() - t_entry;
        }
    }
}

}  // namespace

static uint32_t* g_trace_buf = nullptr;
static int64_t g_trace_words = 0, g_trace_ctas = 0;

int launch_block_attention_umma(const uint32_t* mask_bits, int P, int p_lg,
                                const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, void* out,
                                int N, int H, int T_DST, int T_SRC, int is_causal, cudaStream_t s) {
    SEA_CHECK_ARG(mask_bits != nullptr && (P % 32) == 0 && P <= 1024, "block attention: needs the top-k bit mask, P %% 32 == 0, P <= 1024");
    CUtensorMap t_q, t_k, t_v;
    {
        const uint64_t qdims[4] = {(uint64_t) kUD, (uint64_t) T_DST, (uint64_t) H, (uint64_t) N};
        const uint64_t dims[4] = {(uint64_t) kUD, (uint64_t) T_SRC, (uint64_t) H, (uint64_t) N};
        const uint32_t qbox[4] = {(uint32_t) kUD, (uint32_t) kUM, 1, 1};
        const uint32_t box[4] = {(uint32_t) kUD, (uint32_t) kUN, 1, 1};
        const uint64_t qs[3] = {(uint64_t) q_st * 2, (uint64_t) q_sh * 2, (uint64_t) q_sn * 2};
        const uint64_t ks[3] = {(uint64_t) k_st * 2, (uint64_t) k_sh * 2, (uint64_t) k_sn * 2};
        const uint64_t vs[3] = {(uint64_t) v_st * 2, (uint64_t) v_sh * 2, (uint64_t) v_sn * 2};
        int rc = make_tmap_bf16_sw128(&t_q, const_cast<void*>(q), 4, qdims, qs, qbox);
        if (rc) return rc;
        rc = make_tmap_bf16_sw128(&t_k, const_cast<void*>(k), 4, dims, ks, box);
        if (rc) return rc;
        rc = make_tmap_bf16_sw128(&t_v, const_cast<void*>(v), 4, dims, vs, box);
        if (rc) return rc;
    }
    static const int g_env = getenv("SEA_ATTN_QTILES") ? atoi(getenv("SEA_ATTN_QTILES")) : 0;       // A/B timing switch
    const int G = g_env == 1 || g_env == 2 ? g_env : (T_DST > kUM ? 2 : 1);       // measured at the north-star shape: 179 vs 186 us
    auto launch = [&](auto kernel, auto g_tag) -> int {
        constexpr int kG = decltype(g_tag)::value;
        const int n_blocks = (T_DST + urows(kG) - 1) / urows(kG);
        const size_t smem = 1024 + (size_t) USmem<kG>::kBits + (size_t) urows(kG) * ubits_stride(P >> 5) * 4;
        const unsigned grid = (unsigned) ((int64_t) n_blocks * N * H);
        SEA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
        SEA_CUDA_TRY(launch_pdl(kernel, dim3(grid), dim3(uthreads(kG)), (size_t) smem, s, mask_bits, P, p_lg, t_q, t_k, t_v,
                                (const __nv_bfloat16*) q, q_sn, q_sh, q_st, (const __nv_bfloat16*) k, k_sn, k_sh, k_st, scales,
                                (const __nv_bfloat16*) cumavg, avg_sh, avg_st, use_scaler, (__nv_bfloat16*) out, N, H, T_DST, T_SRC, is_causal, n_blocks),
                     "block_attention_umma_kernel launch");
        return SEA_OK;
    };
    static const bool trace = getenv("SEA_ATTN_TRACE") != nullptr;
    if (trace) {
        const int64_t words = (int64_t) ((T_DST + urows(G) - 1) / urows(G)) * N * H * 20 * 16;
        if (g_trace_words < words) {
            if (g_trace_buf) cudaFree(g_trace_buf);
            SEA_CUDA_TRY(cudaMalloc(&g_trace_buf, (size_t) words * 4), "trace buffer");
            g_trace_words = words;
            SEA_CUDA_TRY(cudaMemcpyToSymbol(g_attn_trace, &g_trace_buf, sizeof(g_trace_buf)), "trace symbol");
        }
        SEA_CUDA_TRY(cudaMemsetAsync(g_trace_buf, 0, (size_t) words * 4, s), "trace clear");
        g_trace_ctas = words / (20 * 16);
        if (G == 2) return launch(block_attention_umma_kernel<2, true>, std::integral_constant<int, 2>{});
        return launch(block_attention_umma_kernel<1, true>, std::integral_constant<int, 1>{});
    }
    if (G == 2) return launch(block_attention_umma_kernel<2>, std::integral_constant<int, 2>{});
    return launch(block_attention_umma_kernel<1>, std::integral_constant<int, 1>{});
}

// development aid: copies the trace of the last traced launch to the host (synchronises); returns the number of CTAs
int64_t attn_trace_read(uint32_t* host, int64_t max_words) {
    if (!g_trace_buf) return 0;
    cudaDeviceSynchronize();
    const int64_t words = g_trace_ctas * 20 * 16 < max_words ? g_trace_ctas * 20 * 16 : max_words;
    cudaMemcpy(host, g_trace_buf, (size_t) words * 4, cudaMemcpyDeviceToHost);
    return g_trace_ctas;
}

}  // namespace sea
