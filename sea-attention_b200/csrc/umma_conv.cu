// a5 on the 5th-generation tensor cores: the predictor's CausalConv2d(3x3, dilation 2) + ReLU and its 1x1
// output conv as IMPLICIT GEMM with tcgen05.mma (bf16 in, fp32 accumulate in TMEM), operands staged by TMA.
//
//   activations channels-last [N, T, W, C=64] bf16  ->  one pixel = one 128-byte row = one SWIZZLE_128B K-atom
//   output tile  = 128 pixels (TR = 128/W rows of t) x kNOut channels, accumulator [128 lanes x kNOut cols] in TMEM
//   K loop       = 9 taps x 64 channels: tap (i, j) is the SAME 4-D TMA box shifted by (dt, dw) = (2i-4, 2j-2);
//                  out-of-range rows/columns (the conv's zero padding, incl. the causal top padding) are zero-filled
//                  by TMA, so no im2col buffer and no boundary code exists anywhere.
//   weights      = 9 x [kNOut x 64] bf16 K-major tiles, resident in shared memory for the whole persistent CTA
//   warp roles   = warp 0: TMA producer, warp 1: MMA issuer (one elected lane) + TMEM allocator,
//                  warps 2-5: epilogue (tcgen05.ld -> +bias, ReLU -> bf16 -> swizzled smem -> coalesced 16B stores)
//   pipelines    = kStages-deep smem ring (full/empty mbarriers), 2 TMEM accumulators (tmem_full/empty mbarriers)
// Reference: modules.py:96-192 (CausalConv2d), attention.py:271-276.
#include "common.cuh"
#include "umma.cuh"

#include <mutex>
#include <stdlib.h>

namespace sea {

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tmap_bf16_sw128(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available (driver entry point lookup failed)");
        return SEA_ERR_CUDA;
    }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t) rank, base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int) r);
        return SEA_ERR_CUDA;
    }
    return SEA_OK;
}

int make_tmap_u32_plain(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available (driver entry point lookup failed)");
        return SEA_ERR_CUDA;
    }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, (cuuint32_t) rank, base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (u32) failed with CUresult %d", (int) r);
        return SEA_ERR_CUDA;
    }
    return SEA_OK;
}

namespace {

constexpr int kStages = 8;
constexpr int kATileBytes = 128 * 128;     // 128 pixels x 64 bf16
constexpr int kConvThreads = 192;

template <int kTaps, int kNOut>
struct ConvSmem {
    static constexpr int kWBytes = kTaps * kNOut * 128;
    static constexpr int kStageOff = kWBytes;
    static constexpr int kOutOff = kStageOff + kStages * kATileBytes;
    static constexpr int kOutBytes = 128 * 128;                      // 128 rows x 128 B (64 bf16 or 32 fp32)
    static constexpr int kBarOff = kOutOff + kOutBytes;
    static constexpr int kTotal = kBarOff + 256 + 1024;              // + barriers + alignment slack
};

__global__ void pack_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int C, int taps) {
    // reference layout [O][C][2k-1][k] (k = 3 -> 5x3, k = 1 -> 1x1)  ->  [tap][o][c] bf16
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= taps * O * C) return;
    const int c = idx % C, o = (idx / C) % O, tap = idx / (C * O);
    float val;
    if (taps == 9) val = w[(((int64_t) o * C + c) * 5 + tap / 3) * 3 + tap % 3];
    else val = w[(int64_t) o * C + c];
    out[idx] = __float2bfloat16_rn(val);
}

template <int kTaps, int kNOut, bool kRelu, typename OutT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const float* __restrict__ bias, OutT* __restrict__ y, int N, int T, int W, int TR, int tblocks, int num_tiles) {
    using SM = ConvSmem<kTaps, kNOut>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - ((uint32_t) __cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);      // (offset from the __shared__ array, not an integer round trip: keeps the shared address space -> LDS / STS, not generic LD / ST)
    uint8_t* wsm = smem;
    uint8_t* asm_ = smem + SM::kStageOff;
    uint8_t* osm = smem + SM::kOutOff;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
    uint64_t* empty = full + kStages;
    uint64_t* tmem_full = empty + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* wbar = tmem_empty + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    if (threadIdx.x == 0) {
        umma::prefetch_tensormap(&tmap_x);
        umma::prefetch_tensormap(&tmap_w);
        for (int s = 0; s < kStages; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(&tmem_full[a], 1); umma::mbar_init(&tmem_empty[a], 128); }
        umma::mbar_init(wbar, 1);
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_ptr, 2 * kNOut);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            umma::mbar_arrive_expect_tx(wbar, SM::kWBytes);
            for (int tap = 0; tap < kTaps; ++tap) umma::tma_load_2d(wsm + tap * kNOut * 128, &tmap_w, wbar, 0, tap * kNOut);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n = tile / tblocks, t0 = (tile % tblocks) * TR;
                for (int tap = 0; tap < kTaps; ++tap) {
                    umma::mbar_wait(&empty[stage], phase ^ 1);
                    umma::mbar_arrive_expect_tx(&full[stage], kATileBytes);
                    const int dt = kTaps == 9 ? (tap / 3) * 2 - 4 : 0;
                    const int dw = kTaps == 9 ? (tap % 3) * 2 - 2 : 0;
                    umma::tma_load_4d(asm_ + stage * kATileBytes, &tmap_x, &full[stage], 0, dw, t0 + dt, n);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {   // all 32 lanes walk the loop; elect.sync inside the *_elect helpers picks the issuing lane
            constexpr uint32_t idesc = umma::make_idesc_bf16(128, kNOut);
            umma::mbar_wait(wbar, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                umma::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t) (acc * kNOut);
                for (int tap = 0; tap < kTaps; ++tap) {
                    umma::mbar_wait(&full[stage], phase);
                    umma::tc_fence_after();
                    const uint32_t a_addr = umma::smem_u32(asm_ + stage * kATileBytes);
                    const uint32_t b_addr = umma::smem_u32(wsm + tap * kNOut * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma::mma_bf16_ss_elect(d_tmem, umma::make_desc_k_sw128(a_addr + k * 32), umma::make_desc_k_sw128(b_addr + k * 32),
                                          idesc, (uint32_t) ((tap | k) != 0));
                    umma::mma_commit_elect(&empty[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma::mma_commit_elect(&tmem_full[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                // pixel row inside the tile
        const int et = threadIdx.x - 64;              // 0..127 among the epilogue threads
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n = tile / tblocks, t0 = (tile % tblocks) * TR;
            umma::mbar_wait(&tmem_full[acc], acc_phase);
            umma::tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kNOut; c0 += 32) {
                uint32_t r[32];
                umma::tmem_ld_32x32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (acc * kNOut + c0), r);
                umma::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float val = __uint_as_float(r[i]) + __ldg(bias + c0 + i);
                    f[i] = kRelu ? fmaxf(val, 0.f) : val;
                }
                if (sizeof(OutT) == 2) {
                    // 32 channels -> 64 bytes = chunks (c0/8) .. (c0/8 + 3) of the 128-byte row
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint4 pk;
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(f[ch * 8 + 0], f[ch * 8 + 1]);
                        __nv_bfloat162 p1 = __floats2bfloat162_rn(f[ch * 8 + 2], f[ch * 8 + 3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(f[ch * 8 + 4], f[ch * 8 + 5]);
                        __nv_bfloat162 p3 = __floats2bfloat162_rn(f[ch * 8 + 6], f[ch * 8 + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                        pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                        const int chunk = (c0 >> 3) + ch;
                        *reinterpret_cast<uint4*>(osm + row * 128 + ((chunk ^ (row & 7)) << 4)) = pk;
                    }
                } else {
                    // fp32 out: 32 channels = the whole 128-byte row (kNOut == 32)
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 pk = make_float4(f[ch * 4 + 0], f[ch * 4 + 1], f[ch * 4 + 2], f[ch * 4 + 3]);
                        *reinterpret_cast<float4*>(osm + row * 128 + ((ch ^ (row & 7)) << 4)) = pk;
                    }
                }
            }
            umma::tc_fence_before();
            umma::mbar_arrive(&tmem_empty[acc]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // 128 rows x 128 B are contiguous in global memory: pixel (t0 + r / W, r % W)
            const int valid_rows = min(TR, T - t0) * W;
            uint8_t* gout = reinterpret_cast<uint8_t*>(y) + (((int64_t) n * T + t0) * W) * (int64_t) (kNOut * sizeof(OutT));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int g = i * 128 + et;
                const int rr = g >> 3, ch = g & 7;
                if (rr < valid_rows)
                    *reinterpret_cast<uint4*>(gout + (int64_t) g * 16) = *reinterpret_cast<const uint4*>(osm + rr * 128 + ((ch ^ (rr & 7)) << 4));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, 2 * kNOut);
    }
}

// ------------------------------------------------------------------------------------------------
// Row-pair ring version of the 3x3 / dilation-2 causal convolution for W = 64 (a tile = 2 time rows x 64 columns).
// The kernel above fetches nine shifted 16 KB windows per tile and is bound by that L2 -> SM traffic (ncu: tensor pipe 21 %).
// Here every input row pair is fetched ONCE, 17 KB instead of 144 KB per tile:
//  * time shifts: with dilation 2 and 2-row tiles the three time taps of tile k are exactly the row pairs k-2, k-1, k, so a
//    CTA walks consecutive tiles and keeps the pairs in a ring;
//  * column shifts: the pair is stored column-major interleaved with a zero halo -- shared-memory row r = 2 * (w + 2) + tt for
//    w = -2 .. 65 (one TMA box through a tensor map whose W and T dimensions are swapped; the halo is TMA's out-of-bounds
//    zero fill = the convolution's zero padding).  A shift by 2 columns is then a shift by 4 rows = 512 B of the operand
//    start address, i.e. the three column taps are three UMMA descriptors into the same buffer (start offsets 0 / 512 /
//    1024 B; measured on B200: the 128B swizzle is applied to absolute shared-memory address bits, so the tap that starts
//    half-way into a 1024-B atom needs no descriptor base-offset -- setting it gives wrong results).
// The MMA order is pair-stationary: when pair p lands, its 36 MMAs go to three TMEM accumulators -- the last taps of tile p
// (which completes it), the middle taps of tile p+1, the first taps of tile p+2 -- and the ring position is released.
// Four accumulators of 64 columns: three in flight plus one being drained by the epilogue warps.  TMEM lane r of a tile is
// pixel (t0 + (r & 1), r >> 1).
// No output staging buffer: every epilogue thread owns 32 channels of one pixel = 64 contiguous bytes of the output.
// ------------------------------------------------------------------------------------------------
constexpr int kRingPos = 4, kRingAcc = 4;
constexpr int kPairRows = 2 * (64 + 4), kPairBytes = kPairRows * 128, kPairSlot = 18 * 1024;   // haloed pair, slot 1024-aligned
constexpr int kRingEpiWarps = 8, kRingThreads = 64 + 32 * kRingEpiWarps;      // TMA warp, MMA warp, 8 epilogue warps
struct RingSmem {
    static constexpr int kWBytes = 9 * 64 * 128;
    static constexpr int kRingOff = kWBytes;
    static constexpr int kW3Off = kRingOff + kRingPos * kPairSlot;          // fused 1x1: [32 out][64 in] bf16, K-major SW128 (4 KB)
    static constexpr int kA2Off = kW3Off + 4096;                            // fused 1x1: two 128 x 64 bf16 activation tiles
    static constexpr int kBarOff = kA2Off + 2 * kATileBytes;
    static constexpr int kBiasOff = kBarOff + 256;
    static constexpr int kTotal = kBiasOff + 512 + 1024;      // 64 + 32 bias floats
};

// kFuse1x1: the 1x1 convolution that follows the second 3x3 convolution (64 -> 32 channels, no activation) runs in the same
// kernel: the epilogue warps write the ReLU'd bf16 tile into shared memory in the K-major SW128 layout instead of global
// memory, the MMA warp multiplies it with the 32 x 64 weight (4 MMAs, N = 32) into a second TMEM accumulator one tile behind
// the main chain, and a second epilogue pass stores fp32 [pixel][32].  The 64-channel activation never reaches HBM.
template <bool kRelu, bool kFuse1x1>
__global__ void __launch_bounds__(kRingThreads, 1)
conv_ring_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_w3, const float* __restrict__ bias, const float* __restrict__ bias3,
                      __nv_bfloat16* __restrict__ y, float* __restrict__ y3, int N, int T, int PT, int pairs_per_cta) {
    constexpr int kNOut = 64, W = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - ((uint32_t) __cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);      // (offset from the __shared__ array, not an integer round trip: keeps the shared address space -> LDS / STS, not generic LD / ST)
    uint8_t* wsm = smem;
    uint8_t* ring = smem + RingSmem::kRingOff;                     // [pos][136 rows x 128 B], row = 2 * (w + 2) + tt
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + RingSmem::kBarOff);     // [pos]
    uint64_t* empty = full + kRingPos;                             // [pos]
    uint64_t* tmem_full = empty + kRingPos;                        // [kRingAcc]
    uint64_t* tmem_empty = tmem_full + kRingAcc;                   // [kRingAcc]
    uint64_t* wbar = tmem_empty + kRingAcc;
    uint64_t* a2_full = wbar + 1;                                  // [2]  (fused 1x1) activation tile written by the epilogue warps
    uint64_t* acc2_full = a2_full + 2;                             // [2]  (fused 1x1) second accumulator complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc2_full + 2);
    uint8_t* w3sm = smem + RingSmem::kW3Off;
    uint8_t* a2sm = smem + RingSmem::kA2Off;
    float* s_bias = reinterpret_cast<float*>(smem + RingSmem::kBiasOff);
    if (threadIdx.x < kNOut) s_bias[threadIdx.x] = bias[threadIdx.x];
    if (kFuse1x1 && threadIdx.x >= 64 && threadIdx.x < 96) s_bias[64 + threadIdx.x - 64] = bias3[threadIdx.x - 64];
    // the shuffle makes the warp index provably warp-uniform, so the role branches below stay in the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int G = N * PT;
    const int g_begin = blockIdx.x * pairs_per_cta, g_end = min(G, g_begin + pairs_per_cta);
    umma::pdl_launch_dependents();                 // the next kernel may run its prologue under this one
    if (threadIdx.x == 0) {
        umma::prefetch_tensormap(&tmap_x);
        umma::prefetch_tensormap(&tmap_w);
        for (int p = 0; p < kRingPos; ++p) { umma::mbar_init(&full[p], 1); umma::mbar_init(&empty[p], 1); }
        for (int a = 0; a < kRingAcc; ++a) { umma::mbar_init(&tmem_full[a], 1); umma::mbar_init(&tmem_empty[a], 32 * kRingEpiWarps); }
        umma::mbar_init(wbar, 1);
        for (int a = 0; a < 2; ++a) { umma::mbar_init(&a2_full[a], kRingEpiWarps); umma::mbar_init(&acc2_full[a], 1); }
        umma::fence_barrier_init();
    }
    constexpr uint32_t kTmemCols = kFuse1x1 ? 512 : kRingAcc * kNOut;       // + 2 x 32 columns for the 1x1 accumulators
    if (warp == 1) umma::tmem_alloc(tmem_ptr, kTmemCols);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0 && g_begin < g_end) {
            umma::mbar_arrive_expect_tx(wbar, RingSmem::kWBytes + (kFuse1x1 ? 4096 : 0));
            for (int tap = 0; tap < 9; ++tap) umma::tma_load_2d(wsm + tap * kNOut * 128, &tmap_w, wbar, 0, tap * kNOut);
            if (kFuse1x1) umma::tma_load_2d(w3sm, &tmap_w3, wbar, 0, 0);
            umma::pdl_wait();                             // weights are parameters; the activations need the previous kernel
            int pos = 0, round = 0;                      // ring position / how many times the ring has wrapped
            for (int g0 = g_begin; g0 < g_end;) {
                const int n = g0 / PT, k0 = g0 % PT;     // one run: tiles [k0, k1) of batch item n
                const int k1 = min(PT, k0 + (g_end - g0));
                for (int pk = k0 - 2; pk < k1; ++pk) {   // row pairs of the run (pk < 0: all zeros, the causal padding)
                    if (round > 0) umma::mbar_wait(&empty[pos], (uint32_t) ((round - 1) & 1));
                    umma::mbar_arrive_expect_tx(&full[pos], kPairBytes);
                    umma::tma_load_4d(ring + pos * kPairSlot, &tmap_x, &full[pos], 0, 2 * pk, -2, n);     // dims (C, T, W, N)
                    if (++pos == kRingPos) { pos = 0; ++round; }
                }
                g0 += k1 - k0;
            }
        }
    } else if (warp == 1) {
        if (g_begin < g_end) {
            // The whole warp walks this loop (warp-uniform control flow and operands); elect.sync inside the helpers picks the
            // issuing lane.  Descriptors are base + constant offsets in units of 16 B: ring slot 1152, column tap 32, weight tap 512,
            // k-step 2.
            constexpr uint32_t idesc = umma::make_idesc_bf16(128, kNOut);
            const uint64_t a_base = umma::make_desc_k_sw128(umma::smem_u32(ring));
            const uint64_t b_base = umma::make_desc_k_sw128(umma::smem_u32(wsm));
            constexpr uint32_t idesc3 = umma::make_idesc_bf16(128, 32);
            const uint64_t a2_base = umma::make_desc_k_sw128(umma::smem_u32(a2sm));
            const uint64_t w3_base = umma::make_desc_k_sw128(umma::smem_u32(w3sm));
            int done2 = 0;                                // tiles whose 1x1 MMAs have been issued
            auto issue_1x1 = [&](int c) {                 // D2[128 x 32] = relu(conv2 tile c)[128 x 64] . W3^T
                const int b2 = c & 1;
                umma::mbar_wait(&a2_full[b2], (uint32_t) ((c >> 1) & 1));
                umma::tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma::mma_bf16_ss_elect(tmem_base + (uint32_t) (kRingAcc * kNOut + 32 * b2), a2_base + (uint64_t) (b2 * (kATileBytes >> 4) + kk * 2),
                                            w3_base + (uint64_t) (kk * 2), idesc3, (uint32_t) (kk != 0));
                umma::mma_commit_elect(&acc2_full[b2]);
            };
            umma::mbar_wait(wbar, 0);
            int pos = 0, round = 0, c_base = 0;           // c_base: tiles of this CTA that belong to earlier runs
            for (int g0 = g_begin; g0 < g_end;) {
                const int k0 = g0 % PT;
                const int k1 = min(PT, k0 + (g_end - g0));
                for (int pk = k0 - 2; pk < k1; ++pk) {
                    if (kFuse1x1) {                       // one tile behind the main chain: its activation tile is long written
                        const int completed = c_base + max(pk - k0, 0);
                        while (done2 < completed - 1) issue_1x1(done2++);
                    }
                    umma::mbar_wait(&full[pos], (uint32_t) (round & 1));
                    umma::tc_fence_after();
                    const uint64_t a_pos = a_base + (uint64_t) (pos * (kPairSlot >> 4));
#pragma unroll
                    for (int ii = 0; ii < 3; ++ii) {      // oldest tile first: tile pk takes this pair with its last taps (i = 2)
                        const int i = 2 - ii;
                        const int t = pk + ii;
                        if (t >= k0 && t < k1) {
                            const int c = c_base + (t - k0);
                            const int slot = c & (kRingAcc - 1);
                            if (i == 0) {                  // first taps of tile t: its accumulator must have been drained
                                umma::mbar_wait(&tmem_empty[slot], (uint32_t) (((c / kRingAcc) & 1) ^ 1));
                                umma::tc_fence_after();
                            }
                            const uint32_t d_tmem = tmem_base + (uint32_t) (slot * kNOut);
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma::mma_bf16_ss_elect(d_tmem, a_pos + (uint64_t) (j * (512 >> 4) + kk * 2),
                                                            b_base + (uint64_t) ((i * 3 + j) * (kNOut * 128 >> 4) + kk * 2), idesc,
                                                            (uint32_t) ((i | j | kk) != 0));
                            }
                            if (i == 2) umma::mma_commit_elect(&tmem_full[slot]);
                        }
                    }
                    umma::mma_commit_elect(&empty[pos]);
                    if (++pos == kRingPos) { pos = 0; ++round; }
                }
                c_base += k1 - k0;
                g0 += k1 - k0;
            }
            if (kFuse1x1) while (done2 < c_base) issue_1x1(done2++);
        }
    } else {
        // Eight epilogue warps: warp % 4 is the TMEM lane quarter the hardware lets a warp read, (warp - 2) / 4 the column
        // half.  Every thread owns 32 channels of one pixel = 64 contiguous bytes of the channels-last output, written with
        // two 256-bit stores (full 32-byte sectors).  Bias comes from shared memory as broadcast float4 reads.
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;                // TMEM lane = pixel (t0 + (row & 1), row >> 1)
        const int c0 = half * 32;
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
        umma::pdl_wait();                                 // before the first store: the output buffer may alias a predecessor's input
        auto finish_1x1 = [&](int c) {                // second epilogue pass of tile c: 16 of the 32 output channels per thread
            const int gg = g_begin + c, n2 = gg / PT, t2 = (gg % PT) * 2 + (row & 1);
            umma::mbar_wait(&acc2_full[c & 1], (uint32_t) ((c >> 1) & 1));
            umma::tc_fence_after();
            uint32_t r2[16];
            umma::tmem_ld_32x16(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (kRingAcc * kNOut + 32 * (c & 1) + 16 * half), r2);
            umma::tmem_ld_wait();
            const float4* b3 = reinterpret_cast<const float4*>(s_bias + 64 + 16 * half);
            float* dst = y3 + ((((int64_t) n2 * T + t2) * W) + (row >> 1)) * 32 + 16 * half;
            if (t2 < T) {
#pragma unroll
                for (int hv = 0; hv < 2; ++hv) {
                    const float4 ba = b3[2 * hv], bb = b3[2 * hv + 1];
                    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8 * hv),
                                 "f"(__uint_as_float(r2[8 * hv]) + ba.x), "f"(__uint_as_float(r2[8 * hv + 1]) + ba.y),
                                 "f"(__uint_as_float(r2[8 * hv + 2]) + ba.z), "f"(__uint_as_float(r2[8 * hv + 3]) + ba.w),
                                 "f"(__uint_as_float(r2[8 * hv + 4]) + bb.x), "f"(__uint_as_float(r2[8 * hv + 5]) + bb.y),
                                 "f"(__uint_as_float(r2[8 * hv + 6]) + bb.z), "f"(__uint_as_float(r2[8 * hv + 7]) + bb.w)
                                 : "memory");
                }
            }
        };
        for (int g = g_begin; g < g_end; ++g) {
            const int n = g / PT, t0 = (g % PT) * 2;
            const int c = g - g_begin, acc = c & (kRingAcc - 1);
            umma::mbar_wait(&tmem_full[acc], (uint32_t) ((c / kRingAcc) & 1));
            umma::tc_fence_after();
            const bool ok = t0 + (row & 1) < T;
            uint8_t* gout = reinterpret_cast<uint8_t*>(y) + ((((int64_t) n * T + t0 + (row & 1)) * W) + (row >> 1)) * (int64_t) (kNOut * 2) + c0 * 2;
            uint32_t r[32];
            umma::tmem_ld_32x32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (acc * kNOut + c0), r);
            umma::tmem_ld_wait();
            umma::tc_fence_before();
            umma::mbar_arrive(&tmem_empty[acc]);       // the accumulator is in registers: hand it back before the stores
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = b4[i];
                float v0 = __uint_as_float(r[4 * i + 0]) + b.x, v1 = __uint_as_float(r[4 * i + 1]) + b.y;
                float v2 = __uint_as_float(r[4 * i + 2]) + b.z, v3 = __uint_as_float(r[4 * i + 3]) + b.w;
                if (kRelu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1), p1 = __floats2bfloat162_rn(v2, v3);
                o[2 * i] = *reinterpret_cast<uint32_t*>(&p0);
                o[2 * i + 1] = *reinterpret_cast<uint32_t*>(&p1);
            }
            if (!kFuse1x1) {
                if (ok) {
#pragma unroll
                    for (int hv = 0; hv < 2; ++hv)
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(gout + hv * 32), "r"(o[8 * hv]), "r"(o[8 * hv + 1]),
                                     "r"(o[8 * hv + 2]), "r"(o[8 * hv + 3]), "r"(o[8 * hv + 4]), "r"(o[8 * hv + 5]), "r"(o[8 * hv + 6]), "r"(o[8 * hv + 7])
                                     : "memory");
                }
            } else {
                // activation tile c -> shared memory, K-major SW128: row = TMEM lane, 16-byte chunk ch at (ch ^ (row & 7)); this
                // buffer was last read by the 1x1 MMAs of tile c-2, whose completion this warp awaited in finish_1x1(c-2)
                uint8_t* arow = a2sm + (c & 1) * kATileBytes + row * 128;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    *reinterpret_cast<uint4*>(arow + (((half * 4 + ch) ^ (row & 7)) << 4)) = make_uint4(o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
                umma::fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&a2_full[c & 1]);
                if (c > 0) finish_1x1(c - 1);
            }
        }
        if (kFuse1x1 && g_begin < g_end) finish_1x1(g_end - g_begin - 1);
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, kTmemCols);
    }
}

struct Fuse1x1 { const float* w3; const float* b3; void* ws3; float* y3; };      // the 1x1 conv fused behind a 3x3 conv (ring kernel only)

template <int kTaps, int kNOut, bool kRelu, typename OutT>
int launch_conv_umma(const void* x, const float* weight, const float* bias, void* y, void* ws, int N, int T, int W, int C, cudaStream_t s,
                     const Fuse1x1* fuse = nullptr) {
    using SM = ConvSmem<kTaps, kNOut>;
    __nv_bfloat16* wpack = reinterpret_cast<__nv_bfloat16*>(ws);
    const int nel = kTaps * kNOut * C;
    if (weight != nullptr)          // nullptr: `workspace` still holds the packing of an earlier call with the same weights
        pack_conv_weights_kernel<<<(nel + 255) / 256, 256, 0, s>>>(weight, wpack, kNOut, C, kTaps);
    SEA_CHECK_LAUNCH("pack_conv_weights_kernel");
    const int TR = 128 / W;
    CUtensorMap tx, tw;
    {
        const uint64_t dims[4] = {(uint64_t) C, (uint64_t) W, (uint64_t) T, (uint64_t) N};
        const uint64_t strides[3] = {(uint64_t) C * 2, (uint64_t) W * C * 2, (uint64_t) T * W * C * 2};
        const uint32_t box[4] = {64, (uint32_t) W, (uint32_t) TR, 1};
        int rc = make_tmap_bf16_sw128(&tx, const_cast<void*>(x), 4, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t) C, (uint64_t) kTaps * kNOut};
        const uint64_t strides[1] = {(uint64_t) C * 2};
        const uint32_t box[2] = {64, (uint32_t) kNOut};
        int rc = make_tmap_bf16_sw128(&tw, wpack, 2, dims, strides, box);
        if (rc) return rc;
    }
    const int tblocks = (T + TR - 1) / TR;
    const int num_tiles = N * tblocks;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if constexpr (kTaps == 9 && kNOut == 64 && sizeof(OutT) == 2) {
        static const bool no_ring = getenv("SEA_CONV_NO_RING") != nullptr;       // development switch for A/B timing
        if (W == 64 && !no_ring) {
            // row-pair ring: CTAs walk runs of consecutive tiles; one run per CTA
            const int grid_r = num_tiles < sms ? num_tiles : sms;
            const int per = (num_tiles + grid_r - 1) / grid_r;
            const int grid_used = (num_tiles + per - 1) / per;
            CUtensorMap txp;                      // the same tensor with W and T swapped: a box is (64 ch, 2 rows, 68 columns)
            {
                const uint64_t dims[4] = {(uint64_t) C, (uint64_t) T, (uint64_t) W, (uint64_t) N};
                const uint64_t strides[3] = {(uint64_t) W * C * 2, (uint64_t) C * 2, (uint64_t) T * W * C * 2};
                const uint32_t box[4] = {64, 2, (uint32_t) (W + 4), 1};
                int rc = make_tmap_bf16_sw128(&txp, const_cast<void*>(x), 4, dims, strides, box);
                if (rc) return rc;
            }
            // programmatic dependent launch: the prologue (barriers, TMEM, 74-78 KB of weights) overlaps the previous kernel's tail
            static const bool no_pdl = getenv("SEA_NO_PDL") != nullptr;          // development switch for A/B timing
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned) grid_used); cfg.blockDim = dim3(kRingThreads); cfg.dynamicSmemBytes = RingSmem::kTotal; cfg.stream = s;
            cudaLaunchAttribute pdl_attr[1];
            pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
            // (not right behind a weight-packing kernel: the packed weights are loaded before pdl_wait())
            const bool fresh_pack = weight != nullptr || (fuse != nullptr && fuse->w3 != nullptr);
            cfg.attrs = pdl_attr; cfg.numAttrs = (no_pdl || fresh_pack) ? 0 : 1;
            if (fuse == nullptr) {
                auto kr = conv_ring_umma_kernel<kRelu, false>;
                SEA_CUDA_TRY(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, RingSmem::kTotal), "smem attr");
                SEA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kr, txp, tw, tw, bias, (const float*) nullptr, reinterpret_cast<__nv_bfloat16*>(y), (float*) nullptr, N, T,
                                                tblocks, per), "conv_ring_umma_kernel launch");
            } else {
                // + the 1x1 convolution 64 -> 32 (fp32 output), its bf16 weight packing [32][64] lives in fuse->ws3
                __nv_bfloat16* w3pack = reinterpret_cast<__nv_bfloat16*>(fuse->ws3);
                if (fuse->w3 != nullptr) pack_conv_weights_kernel<<<(32 * 64 + 255) / 256, 256, 0, s>>>(fuse->w3, w3pack, 32, C, 1);
                SEA_CHECK_LAUNCH("pack_conv_weights_kernel");
                CUtensorMap tw3;
                const uint64_t dims[2] = {(uint64_t) C, 32};
                const uint64_t strides[1] = {(uint64_t) C * 2};
                const uint32_t box[2] = {64, 32};
                int rc = make_tmap_bf16_sw128(&tw3, w3pack, 2, dims, strides, box);
                if (rc) return rc;
                auto kr = conv_ring_umma_kernel<kRelu, true>;
                SEA_CUDA_TRY(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, RingSmem::kTotal), "smem attr");
                SEA_CUDA_TRY(cudaLaunchKernelEx(&cfg, kr, txp, tw, tw3, bias, fuse->b3, (__nv_bfloat16*) nullptr, fuse->y3, N, T, tblocks, per),
                             "conv_ring_umma_kernel launch");
            }
            SEA_CHECK_LAUNCH("conv_ring_umma_kernel");
            return SEA_OK;
        }
    }
    if (fuse != nullptr) {
        set_error("conv3x3 + conv1x1 fusion needs the row-pair ring kernel (W = 64)");
        return SEA_ERR_UNSUPPORTED;
    }
    auto kern = conv_umma_kernel<kTaps, kNOut, kRelu, OutT>;
    SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal), "smem attr");
    const int grid = num_tiles < sms ? num_tiles : sms;
    kern<<<grid, kConvThreads, SM::kTotal, s>>>(tx, tw, bias, reinterpret_cast<OutT*>(y), N, T, W, TR, tblocks, num_tiles);
    SEA_CHECK_LAUNCH("conv_umma_kernel");
    return SEA_OK;
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" {

int sea_conv_umma_supported(int dtype, int W, int C, int O) {
    return dtype == SEA_DTYPE_BF16 && C == 64 && (O == 64 || O == 32) && W >= 1 && W <= 128 && (128 % W) == 0;
}

int64_t sea_conv_umma_workspace_bytes(int C, int O) { return (int64_t) 9 * O * C * 2 + 1024; }

int sea_causal_conv3x3_dil2_relu_umma(const void* x, const float* weight, const float* bias, void* y, void* workspace,
                                      int N, int T, int W, int C, int O, void* stream) {
    SEA_CHECK_ARG(x && bias && y && workspace, "sea_causal_conv3x3_dil2_relu_umma: null pointer");
    SEA_CHECK_ARG(N > 0 && T > 0, "sea_causal_conv3x3_dil2_relu_umma: bad shape");
    if (!sea_conv_umma_supported(SEA_DTYPE_BF16, W, C, O) || O != 64) {
        set_error("sea_causal_conv3x3_dil2_relu_umma: unsupported shape W=%d C=%d O=%d (need C=O=64, W | 128)", W, C, O);
        return SEA_ERR_UNSUPPORTED;
    }
    SEA_CHECK_ARG((((uintptr_t) x) & 127) == 0 && (((uintptr_t) y) & 15) == 0 && (((uintptr_t) workspace) & 127) == 0,
                  "sea_causal_conv3x3_dil2_relu_umma: misaligned pointer");
    return launch_conv_umma<9, 64, true, __nv_bfloat16>(x, weight, bias, y, workspace, N, T, W, C, (cudaStream_t) stream);
}

int sea_conv3x3_conv1x1_umma_supported(int dtype, int W, int C, int O, int O3) {
    return dtype == SEA_DTYPE_BF16 && W == 64 && C == 64 && O == 64 && O3 == 32 && getenv("SEA_CONV_NO_RING") == nullptr;
}

int sea_causal_conv3x3_dil2_relu_conv1x1_umma(const void* x, const float* weight, const float* bias, void* workspace,
                                              const float* weight3, const float* bias3, void* workspace3, float* y3,
                                              int N, int T, int W, int C, int O, int O3, void* stream) {
    SEA_CHECK_ARG(x && bias && workspace && bias3 && workspace3 && y3, "sea_causal_conv3x3_dil2_relu_conv1x1_umma: null pointer");
    SEA_CHECK_ARG(N > 0 && T > 0, "sea_causal_conv3x3_dil2_relu_conv1x1_umma: bad shape");
    if (!sea_conv3x3_conv1x1_umma_supported(SEA_DTYPE_BF16, W, C, O, O3)) {
        set_error("sea_causal_conv3x3_dil2_relu_conv1x1_umma: unsupported shape W=%d C=%d O=%d O3=%d (need W=C=O=64, O3=32)", W, C, O, O3);
        return SEA_ERR_UNSUPPORTED;
    }
    SEA_CHECK_ARG((((uintptr_t) x) & 127) == 0 && (((uintptr_t) y3) & 31) == 0 && (((uintptr_t) workspace) & 127) == 0 && (((uintptr_t) workspace3) & 127) == 0,
                  "sea_causal_conv3x3_dil2_relu_conv1x1_umma: misaligned pointer");
    const Fuse1x1 fuse{weight3, bias3, workspace3, y3};
    return launch_conv_umma<9, 64, true, __nv_bfloat16>(x, weight, bias, nullptr, workspace, N, T, W, C, (cudaStream_t) stream, &fuse);
}

int sea_conv1x1_umma(const void* x, const float* weight, const float* bias, float* y, void* workspace,
                     int N, int T, int W, int C, int O, void* stream) {
    SEA_CHECK_ARG(x && bias && y && workspace, "sea_conv1x1_umma: null pointer");
    if (!sea_conv_umma_supported(SEA_DTYPE_BF16, W, C, O) || O != 32) {
        set_error("sea_conv1x1_umma: unsupported shape W=%d C=%d O=%d (need C=64, O=32, W | 128)", W, C, O);
        return SEA_ERR_UNSUPPORTED;
    }
    SEA_CHECK_ARG((((uintptr_t) x) & 127) == 0 && (((uintptr_t) y) & 15) == 0 && (((uintptr_t) workspace) & 127) == 0, "sea_conv1x1_umma: misaligned pointer");
    return launch_conv_umma<1, 32, false, float>(x, weight, bias, y, workspace, N, T, W, C, (cudaStream_t) stream);
}

}  // extern "C"
