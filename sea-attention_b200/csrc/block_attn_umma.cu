// a9-a14 of the short-context path on the 5th-generation tensor cores (bf16, d = 64): same algorithm, inputs (dense bit-packed
// mask from expand_mask_kernel) and results as block_attention_bits_kernel (block_attn.cu), but the two contractions of a tile
// are tcgen05.mma instructions with the accumulators in TMEM, so the 128 softmax threads work thread-per-row with no
// ldmatrix / mma.sync / shuffle traffic at all -- the mma.sync kernel is issue-bound on exactly that work.
//
// CTA = 128 query rows of one head, 192 threads:
//   warp 0 (one lane)  TMA producer: Q tile once; per active source tile one stage = K [64 x 128 B], V [64 x 128 B] (both
//                      SWIZZLE_128B boxes straight from the strided [N,H,T,d] tensors) + the tile pair's element masks
//   warp 1 (one lane)  MMA issuer:  S[128 x 64]  = Q . K^T      (A = Q K-major, B = K K-major, 4 k-steps of 16)  -> TMEM S[j & 1]
//                                   O[128 x 64] += P . V        (A = P K-major from shared memory, B = V as the MN-major
//                                                                operand: rows = source tokens, 128 B = 64 channels)
//                      S of tile j + 1 is issued before P.V of tile j, so it overlaps the softmax of tile j
//   warps 2-5          softmax, thread = query row: tcgen05.ld of the row's 64 scores, element mask from the stage's mask
//                      block with immediate bit tests, lazy running maximum (the row reference only moves when a tile exceeds
//                      it by 2^8; the rare O correction is a tcgen05.ld / st round trip), P -> bf16 -> shared memory in the
//                      swizzled K-major layout tcgen05 reads; 16-column groups with no alive element in the warp's 32 rows
//                      are skipped (zeros stored).
// TMEM: 2 x 64 columns of S + 64 columns of O (256 allocated) -> 2 CTAs per SM.  Reference: attention.py:1151-1173, 1237-1244,
// 1279-1282; mask semantics causal_resize_m_to_t.py:648-762 (via the dense mask).
#include "common.cuh"
#include "umma.cuh"
#include "block_attn.cuh"

namespace sea {
namespace {

constexpr int kUM = 128;                 // query rows per CTA
constexpr int kUN = 64;                  // source tokens per tile
constexpr int kUD = 64;                  // head dim
constexpr int kUStages = 3;
constexpr int kUTile = kUN * kUD * 2;    // 8 KB (K or V tile)
constexpr int kUMask = kUM * 16;         // 2 KB
constexpr int kUStage = 2 * kUTile + kUMask;
constexpr int kUQ = kUM * kUD * 2;       // 16 KB
constexpr int kUP = kUM * kUN * 2;       // 16 KB per P buffer
constexpr int kUSoftmaxThreads = 256;          // 2 threads per query row: 32 of the 64 tile columns each
constexpr int kUThreads = 64 + kUSoftmaxThreads;
constexpr int kUMaxTileWords = 64;

struct USmem {
    static constexpr int kQ = 0;
    static constexpr int kKV = kQ + kUQ;
    static constexpr int kP = kKV + kUStages * kUStage;          // kUStage is a multiple of 1024
    static constexpr int kAct = kP + 2 * kUP;
    static constexpr int kXch = kAct + kUMaxTileWords * 4;       // float [2 (tile parity)][2 halves][128 rows]: row maxima / sums exchanged between the halves
    static constexpr int kBar = kXch + 4 * kUM * 4;
    static constexpr int kNumBars = 1 + 2 * kUStages + 2 + 2 + 2 + 2;
    static constexpr int kList = kBar + kNumBars * 8 + 16;
};

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

__global__ void __launch_bounds__(kUThreads, 2)
block_attention_umma_kernel(const uint32_t* __restrict__ tile_act, int act_words,
                            const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                            const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_m,
                            const float* __restrict__ scales, const __nv_bfloat16* __restrict__ cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler,
                            __nv_bfloat16* __restrict__ out, int N, int H, int T_DST, int T_SRC, int is_causal, int n_row_blocks, int max_tiles) {
    extern __shared__ uint8_t usm_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(usm_raw) + 1023) & ~(uintptr_t) 1023);
    uint32_t* sact = reinterpret_cast<uint32_t*>(sm + USmem::kAct);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + USmem::kBar);
    uint64_t* q_full = bars;                         // [1]
    uint64_t* kv_full = bars + 1;                    // [stages]
    uint64_t* kv_empty = kv_full + kUStages;         // [stages]
    uint64_t* s_full = kv_empty + kUStages;          // [2]
    uint64_t* s_empty = s_full + 2;                  // [2]
    uint64_t* p_full = s_empty + 2;                  // [2]
    uint64_t* p_empty = p_full + 2;                  // [2]   (= P.V of the tile that used the buffer has completed)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p_empty + 2);
    uint16_t* slist = reinterpret_cast<uint16_t*>(sm + USmem::kList);
    __shared__ int s_nact;

    // warp index (and below the tile count) through a shuffle: provably warp-uniform, which keeps the MMA issue loop in the uniform datapath
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    pdl_launch_dependents();
    pdl_wait();                // the set-up below already reads the tile activity written by the expansion kernel
    const int rb = n_row_blocks - 1 - (int) (blockIdx.x / (unsigned) (N * H));       // heavy (late) row blocks first
    const int nh = (int) (blockIdx.x % (unsigned) (N * H));
    const int n = nh / H, h = nh % H;
    const int r0 = rb * kUM;

    // ---- set-up -------------------------------------------------------------------------------------------------------
    if (tid < kUMaxTileWords)
        sact[tid] = tid < act_words ? __ldg(tile_act + (((int64_t) n * H + h) * n_row_blocks + rb) * act_words + tid) : 0u;
    if (tid == 0) {
        umma::prefetch_tensormap(&tmap_q); umma::prefetch_tensormap(&tmap_k); umma::prefetch_tensormap(&tmap_v); umma::prefetch_tensormap(&tmap_m);
        umma::mbar_init(q_full, 1);
        for (int s = 0; s < kUStages; ++s) { umma::mbar_init(&kv_full[s], 1); umma::mbar_init(&kv_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            umma::mbar_init(&s_full[b], 1); umma::mbar_init(&s_empty[b], kUSoftmaxThreads);
            umma::mbar_init(&p_full[b], kUSoftmaxThreads); umma::mbar_init(&p_empty[b], 1);
        }
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_ptr, 256);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 0) {
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
    const int nact = __shfl_sync(0xffffffffu, s_nact, 0);
    const uint32_t tS0 = tmem_base, tS1 = tmem_base + 64, tO = tmem_base + 128;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0 && nact > 0) {        // (no active tile: nothing may be left in flight when the CTA exits)
            umma::mbar_arrive_expect_tx(q_full, kUQ);
            umma::tma_load_4d(sm + USmem::kQ, &tmap_q, q_full, 0, r0, h, n);
            for (int j = 0; j < nact; ++j) {
                const int s = j % kUStages;
                if (j >= kUStages) umma::mbar_wait(&kv_empty[s], (uint32_t) ((j / kUStages - 1) & 1));
                uint8_t* dst = sm + USmem::kKV + s * kUStage;
                const int tile = (int) slist[j];
                umma::mbar_arrive_expect_tx(&kv_full[s], kUStage);
                umma::tma_load_4d(dst, &tmap_k, &kv_full[s], 0, tile * kUN, h, n);
                umma::tma_load_4d(dst + kUTile, &tmap_v, &kv_full[s], 0, tile * kUN, h, n);
                umma::tma_load_3d(dst + 2 * kUTile, &tmap_m, &kv_full[s], (tile & ~1) * 2, r0, n * H + h);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (nact > 0) {          // all 32 lanes walk the loop; elect.sync inside the *_elect helpers picks the issuing lane
            const uint32_t idesc_s = umma::make_idesc_bf16(kUM, kUN);          // A, B K-major
            const uint32_t idesc_o = idesc_bf16_b_mn(kUM, kUD);                // B = V, MN-major
            const uint32_t q_addr = umma::smem_u32(sm + USmem::kQ);
            const uint32_t kv_addr = umma::smem_u32(sm + USmem::kKV);
            const uint32_t p_addr = umma::smem_u32(sm + USmem::kP);
            auto issue_s = [&](int j) {
                const int s = j % kUStages, b = j & 1;
                umma::mbar_wait(&kv_full[s], (uint32_t) ((j / kUStages) & 1));
                if (j >= 2) umma::mbar_wait(&s_empty[b], (uint32_t) (((j >> 1) - 1) & 1));
                umma::tc_fence_after();
                const uint32_t ka = kv_addr + (uint32_t) s * kUStage;
#pragma unroll
                for (int k = 0; k < kUD / 16; ++k)
                    umma::mma_bf16_ss_elect(b ? tS1 : tS0, umma::make_desc_k_sw128(q_addr + k * 32), umma::make_desc_k_sw128(ka + k * 32), idesc_s, (uint32_t) (k != 0));
                umma::mma_commit_elect(&s_full[b]);
            };
            umma::mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nact; ++j) {
                const int s = j % kUStages, b = j & 1;
                if (j + 1 < nact) issue_s(j + 1);
                umma::mbar_wait(&p_full[b], (uint32_t) ((j >> 1) & 1));
                umma::tc_fence_after();
                const uint32_t va = kv_addr + (uint32_t) s * kUStage + kUTile;
                const uint32_t pa = p_addr + (uint32_t) b * kUP;
#pragma unroll
                for (int k = 0; k < kUN / 16; ++k)       // 16 source tokens per step: A advances 32 B inside the row, B two 8-token groups
                    umma::mma_bf16_ss_elect(tO, umma::make_desc_k_sw128(pa + k * 32), umma::make_desc_k_sw128(va + k * 2048), idesc_o, (uint32_t) ((j | k) != 0));
                umma::mma_commit_elect(&kv_empty[s]);
                umma::mma_commit_elect(&p_empty[b]);
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax warps: 2 threads per query row
        // warps 2-5 take tile columns 0-31 of their TMEM lane quarter's rows, warps 6-9 columns 32-63 (and the matching halves
        // of P and O); the two halves of a row agree on the row maximum through shared memory + a 64-thread named barrier
        const int qd = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const int t = r0 + row;
        const uint32_t lane_addr = (uint32_t) (qd * 32) << 16;
        float* xch = reinterpret_cast<float*>(sm + USmem::kXch);
        constexpr float kLog2e = 1.4426950408889634f;
        constexpr float kLazy = 8.0f / kLog2e;
        float m_run = -INFINITY, l_run = 0.f, nms = 0.f;
        for (int j = 0; j < nact; ++j) {
            const int s = j % kUStages, b = j & 1;
            const int tile = (int) slist[j];
            umma::mbar_wait(&kv_full[s], (uint32_t) ((j / kUStages) & 1));         // acquire the TMA-written mask block
            const uint32_t mw = *reinterpret_cast<const uint32_t*>(sm + USmem::kKV + s * kUStage + 2 * kUTile + row * 16 + (tile & 1) * 8 + half * 4);
            const uint32_t act = __reduce_or_sync(kFull, ((mw & 0xffffu) ? 1u : 0u) | ((mw >> 16) ? 2u : 0u));
            umma::mbar_wait(&s_full[b], (uint32_t) ((j >> 1) & 1));
            umma::tc_fence_after();
            float sc[32];
            {
                uint32_t r32[32];
                umma::tmem_ld_32x32((b ? tS1 : tS0) + lane_addr + 32u * half, r32);
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) sc[i] = __uint_as_float(r32[i]);
            }
            umma::tc_fence_before();
            umma::mbar_arrive(&s_empty[b]);                        // S[b] may be overwritten by the scores of tile j + 2
            float mx = -INFINITY;
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {
                if (act & (1u << cg)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = cg * 16 + i;
                        sc[c] = (mw & (1u << c)) ? sc[c] : -INFINITY;
                        mx = fmaxf(mx, sc[c]);
                    }
                }
            }
            // row maximum of the tile across both halves
            xch[(b * 2 + half) * kUM + row] = mx;                  // slots alternate with the tile parity: no write-after-read race
            asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
            mx = fmaxf(mx, xch[(b * 2 + (half ^ 1)) * kUM + row]);
            // lazy reference maximum; the (rare) O correction needs P.V of tile j - 1 to have completed
            const bool grow = mx > m_run + kLazy;
            if (__any_sync(kFull, grow)) {
                const float mn = grow ? mx : m_run;
                const float al = mn == -INFINITY ? 1.f : ex2u((m_run - mn) * kLog2e);           // m_run = -inf -> 0
                m_run = mn;
                nms = mn == -INFINITY ? 0.f : -mn * kLog2e;
                l_run *= al;
                if (j > 0) {
                    umma::mbar_wait(&p_empty[(j - 1) & 1], (uint32_t) (((j - 1) >> 1) & 1));
                    umma::tc_fence_after();
                    uint32_t o32[32];
                    umma::tmem_ld_32x32(tO + lane_addr + 32u * half, o32);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * al);
                    tmem_st_32x32(tO + lane_addr + 32u * half, o32);
                    tmem_st_wait();
                }
            }
            // P[b] is free once P.V of tile j - 2 has completed
            if (j >= 2) umma::mbar_wait(&p_empty[b], (uint32_t) (((j >> 1) - 1) & 1));
            uint8_t* prow = sm + USmem::kP + b * kUP + row * 128;
            float ps = 0.f;
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {
                uint4 c0 = make_uint4(0, 0, 0, 0), c1 = make_uint4(0, 0, 0, 0);
                if (act & (1u << cg)) {
                    float p[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { p[i] = ex2u(fmaf(sc[cg * 16 + i], kLog2e, nms)); ps += p[i]; }
                    c0 = make_uint4(pack_bf(p[0], p[1]), pack_bf(p[2], p[3]), pack_bf(p[4], p[5]), pack_bf(p[6], p[7]));
                    c1 = make_uint4(pack_bf(p[8], p[9]), pack_bf(p[10], p[11]), pack_bf(p[12], p[13]), pack_bf(p[14], p[15]));
                }
                const int ch = half * 4 + 2 * cg;                 // 16-byte chunk (8 source tokens) inside the 128-byte row
                *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = c0;
                *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ (row & 7)) << 4)) = c1;
            }
            l_run += ps;
            umma::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
            umma::tc_fence_before();
            umma::mbar_arrive(&p_full[b]);
        }
        // ---- epilogue: O / l, * sigmoid(s0), mix with the running mean, permuted store (32 channels per thread) -------------
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");         // the partner is done with the last exchange slot
        xch[half * kUM + row] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
        l_run += xch[(half ^ 1) * kUM + row];
        float o[32];
        if (nact > 0) {
            umma::mbar_wait(&p_empty[(nact - 1) & 1], (uint32_t) (((nact - 1) >> 1) & 1));
            umma::tc_fence_after();
            uint32_t r32[32];
            umma::tmem_ld_32x32(tO + lane_addr + 32u * half, r32);
            umma::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r32[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0.f;
        }
        if (t < T_DST) {
            const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
            const float* sp = scales + ((((int64_t) n * H + h) * T_DST + t) << 1);
            const float psc = use_scaler ? sigu(sp[0]) : 1.0f;
            const float a = sigu(sp[1]);
            __nv_bfloat16* orow = out + ((int64_t) n * T_DST + t) * ((int64_t) H * kUD) + (int64_t) h * kUD + 32 * half;
            const uint4* arow = cumavg ? reinterpret_cast<const uint4*>(cumavg + ((int64_t) n * H + h) * avg_sh + (int64_t) t * avg_st + 32 * half) : nullptr;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = l_run > 0.f ? o[c8 * 8 + i] * inv * psc : 0.f;
                if (arow) {
                    const uint4 av = __ldg(arow + c8);
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
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

int launch_block_attention_umma(const unsigned long long* dmask, int W64, const uint32_t* tile_act, int act_words,
                                const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, void* out,
                                int N, int H, int T_DST, int T_SRC, int is_causal, cudaStream_t s) {
    const int n_row_blocks = (T_DST + kUM - 1) / kUM;
    const int max_tiles = (T_SRC + kUN - 1) / kUN;
    SEA_CHECK_ARG(max_tiles <= kUMaxTileWords * 32, "block attention: T_SRC too large");
    CUtensorMap t_q, t_k, t_v, t_m;
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
        const uint64_t mdims[3] = {(uint64_t) W64 * 2, (uint64_t) T_DST, (uint64_t) N * H};
        const uint64_t mstr[2] = {(uint64_t) W64 * 8, (uint64_t) T_DST * W64 * 8};
        const uint32_t mbox[3] = {4, (uint32_t) kUM, 1};
        rc = make_tmap_u32_plain(&t_m, const_cast<unsigned long long*>(dmask), 3, mdims, mstr, mbox);
        if (rc) return rc;
    }
    const size_t smem = 1024 + (size_t) USmem::kList + (size_t) ((max_tiles + 7) & ~7) * 2;
    const unsigned grid = (unsigned) ((int64_t) n_row_blocks * N * H);
    SEA_CUDA_TRY(cudaFuncSetAttribute(block_attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
    SEA_CUDA_TRY(launch_pdl(block_attention_umma_kernel, dim3(grid), dim3(kUThreads), (size_t) smem, s, tile_act, act_words, t_q, t_k, t_v, t_m, scales,
                            (const __nv_bfloat16*) cumavg, avg_sh, avg_st, use_scaler, (__nv_bfloat16*) out, N, H, T_DST, T_SRC, is_causal, n_row_blocks, max_tiles),
                 "block_attention_umma_kernel launch");
    return SEA_OK;
}

}  // namespace sea
