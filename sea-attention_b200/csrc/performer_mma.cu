// a2 + a3 (+ a13 running mean) for bf16: the causal Performer estimate as chunk-parallel tensor-core GEMMs.  Written for D = 64
// (one slab); any head dim that is a multiple of 16 (OPT-2.7B D = 80, the long-context sweep D = 128) runs the same kernels once per
// 128-column SLAB of V2 = [pos_emb | v] (blockIdx.z): the output columns of linear attention are independent given phi(q), phi(k), so
// a slab only recomputes the two feature maps (K dimension D) and carries its own ones column for the denominator.
//
// Everything the stage needs per 128-row chunk is a small GEMM, chained through REGISTERS (accumulator fragments are
// re-packed as the next MMA's A fragments, flash-attention style), so nothing but q, k, v is read and nothing but
// ctx / cumavg is written:
//   phi(k) = relu(d^-1/4 K P^T)+1e-3, phi(q) likewise                [128 x 64] x [64 x Fp]
//   pass A : S_c = phi_ext(k)^T V2ext                                [Fp x 128] x [128 x 144]   per chunk -> workspace
//   prefix : exclusive prefix of S_c over the chunks of one (n, h)   (performer.cu: performer_prefix_kernel)
//   pass C : O = tril(phi(q) phi(k)^T) V2ext + phi(q) S_prev         [128 x 128] x [128 x 144] + [128 x Fp] x [Fp x 144]
//            ctx = O[:, :128] / O[:, 128] ;  cumavg = (tril(1) V + vsum_prev) / (t+1)
// with V2ext = [pos_emb | v | 1 | 0-pad] (the `1` column carries the denominator: O[:,128] = sum_j A_ij + phi(q).z) and
// phi_ext(k) = [phi(k) | 1 | 0-pad] (the `1` feature makes row F of S the column sums of V2ext = running sum of v).
// Warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) + ldmatrix: the FLOPs here are ~13 GF per layer, the stage is
// bound by HBM and launch latency, not by the tensor pipe; tcgen05 would not change the roofline.
// Reference: performer_pytorch.FastAttention call sites attention.py:159-164, 504-508, 527-534; running mean :1237-1241.
#include <stdlib.h>
#include "common.cuh"

namespace sea {
namespace {

constexpr int kCh = 128;            // rows per chunk
constexpr int kThreads = 256;       // 8 warps x 16 rows
constexpr int kE = 128;             // v2 slab width
constexpr int kEx = 144;            // v2ext slab width (ones column at 128)
constexpr int kLdV = kEx + 8;       // smem row strides (elements); +8 keeps ldmatrix rows on distinct banks
__host__ __device__ constexpr int ld_q(int D) { return D + 8; }
__host__ __device__ constexpr int n_slabs(int D) { return (2 * D + kE - 1) / kE; }

// which columns of slab z of V2 = [pos_emb (D) | v (D)] come from where (all bounds multiples of 16)
struct Slab {
    int pos_hi;      // slab columns [0, pos_hi) = pos_emb[:, pos_off + c]
    int pos_off;
    int v_lo, v_hi;  // slab columns [v_lo, v_hi) = v[:, v_off + c - v_lo]
    int v_off;
    int width;       // real columns of the slab (the rest up to 128 is zero)
};
__device__ __forceinline__ Slab make_slab(int D, int z) {
    const int g0 = z * kE;
    Slab s;
    s.pos_hi = max(0, min(kE, D - g0));
    s.pos_off = g0;
    s.v_lo = max(0, D - g0);
    s.v_hi = max(s.v_lo, min(kE, 2 * D - g0));
    s.v_off = max(0, g0 - D);
    s.width = max(s.pos_hi, s.v_hi);
    return s;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

template <int kFp, int kDm>
struct PerfSmem {
    static constexpr int kLdQ = ld_q(kDm);
    static constexpr int kP = 0;                                // [kFp][kLdQ]
    static constexpr int kQ = kP + kFp * kLdQ;                  // [kCh][kLdQ]
    static constexpr int kK = kQ + kCh * kLdQ;                  // [kCh][kLdQ]
    static constexpr int kPhiK = kK + kCh * kLdQ;               // [kCh][kFp + 8]
    static constexpr int kV = kPhiK + kCh * (kFp + 8);          // [kCh][kLdV]
    static constexpr int kS = kV + kCh * kLdV;                  // [kFp][kLdV]
    static constexpr int kElems = kS + kFp * kLdV;
    static constexpr int kBytes = kElems * 2 + kE * 4;          // + vsum_prev fp32 [<= 128]
    // pass A needs neither Q nor S_prev, and its phi(k) tile may overwrite the K tile it was computed from: [P | K = PhiK | V]
    // (64 KB at Fp = 48, D = 64 -> three CTAs per SM instead of two)
    static constexpr int kSumsK = kFp * kLdQ;
    static constexpr int kSumsV = kSumsK + kCh * kLdQ;
    static constexpr int kSumsBytes = (kSumsV + kCh * kLdV) * 2;
    static_assert(kFp + 8 <= kLdQ, "phi(k) rows must fit the K rows they replace");
};

// ---- cooperative loads -------------------------------------------------------------------------------------------
// The fp32 operands (projection, positional embedding, prefixed state) go through registers to be rounded to bf16.  Their
// loads are split into an "issue" and a "commit" half so that a kernel can put ALL of its global loads in flight before it
// consumes the first one: one DRAM round trip per CTA instead of one per operand (the CTAs are latency-bound at 2 per SM).
template <int kFp, int kDm>
struct ProjRegs { static constexpr int kVec = kFp * kDm / 4, kIt = (kVec + kThreads - 1) / kThreads; float4 p[kIt]; };
template <int kFp, int kDm>
__device__ __forceinline__ void issue_proj(ProjRegs<kFp, kDm>& r, const float* __restrict__ proj, int F) {
#pragma unroll
    for (int it = 0; it < ProjRegs<kFp, kDm>::kIt; ++it) {
        const int idx = threadIdx.x + it * kThreads, f = idx / (kDm / 4);
        r.p[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < ProjRegs<kFp, kDm>::kVec && f < F) r.p[it] = __ldg(reinterpret_cast<const float4*>(proj) + idx);
    }
}
template <int kFp, int kDm>
__device__ __forceinline__ void commit_proj(__nv_bfloat16* Ps, const ProjRegs<kFp, kDm>& r) {      // [F x D] fp32 -> bf16 [kFp x kLdQ]
#pragma unroll
    for (int it = 0; it < ProjRegs<kFp, kDm>::kIt; ++it) {
        const int idx = threadIdx.x + it * kThreads, f = idx / (kDm / 4), c4 = idx % (kDm / 4);
        if (idx < ProjRegs<kFp, kDm>::kVec)
            *reinterpret_cast<uint2*>(Ps + f * ld_q(kDm) + c4 * 4) = make_uint2(pack_bf16(r.p[it].x, r.p[it].y), pack_bf16(r.p[it].z, r.p[it].w));
    }
}
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int kDm>
__device__ __forceinline__ void load_rows_bf16(__nv_bfloat16* dst, int ld, const __nv_bfloat16* __restrict__ src, int64_t row_stride,
                                               int r0, int nvalid) {
    // 128 rows x D bf16 as 16-byte async copies: no registers are held while the data is in flight, so the q, k and v
    // tiles of a chunk are all requested at once; rows beyond nvalid are zero-filled
    constexpr int kPieces = kDm / 8, kIt = (kCh * kPieces + kThreads - 1) / kThreads;
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int idx = threadIdx.x + it * kThreads, r = idx / kPieces, c8 = idx % kPieces;
        const bool ok = r < nvalid;
        if (idx < kCh * kPieces) cp_async16(dst + r * ld + c8 * 8, src + (int64_t) (r0 + (ok ? r : 0)) * row_stride + c8 * 8, ok ? 16 : 0);
    }
}
// the v part of a slab: `pieces` 16-byte pieces per row (run-time: depends on the slab)
__device__ __forceinline__ void load_rows_bf16_n(__nv_bfloat16* dst, int ld, const __nv_bfloat16* __restrict__ src, int64_t row_stride,
                                                 int r0, int nvalid, int pieces) {
    for (int idx = threadIdx.x; idx < kCh * pieces; idx += kThreads) {
        const int r = idx / pieces, c8 = idx - r * pieces;
        const bool ok = r < nvalid;
        cp_async16(dst + r * ld + c8 * 8, src + (int64_t) (r0 + (ok ? r : 0)) * row_stride + c8 * 8, ok ? 16 : 0);
    }
}
template <int kDm>
struct PosRegs { static constexpr int kPosVec = (kDm < kE ? kDm : kE) / 4, kIt = kCh * kPosVec / kThreads; float4 p[kIt]; };
// V2ext slab = [pos_emb part | v part | 0 ... | 1 0 ...]: v -> its columns by cp.async, pos_emb (fp32) -> registers
template <int kDm>
__device__ __forceinline__ void issue_v2ext(PosRegs<kDm>& r, __nv_bfloat16* Vs, const __nv_bfloat16* __restrict__ v, int64_t v_st,
                                            const float* __restrict__ pos, int r0, int nvalid, const Slab& sl) {
    if (sl.v_hi > sl.v_lo) load_rows_bf16_n(Vs + sl.v_lo, kLdV, v + sl.v_off, v_st, r0, nvalid, (sl.v_hi - sl.v_lo) >> 3);
#pragma unroll
    for (int it = 0; it < PosRegs<kDm>::kIt; ++it) {
        const int idx = threadIdx.x + it * kThreads, row = idx / PosRegs<kDm>::kPosVec, c4 = idx % PosRegs<kDm>::kPosVec;
        r.p[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nvalid && c4 * 4 < sl.pos_hi) r.p[it] = __ldg(reinterpret_cast<const float4*>(pos + (int64_t) (r0 + row) * kDm + sl.pos_off) + c4);
    }
}
template <int kDm>
__device__ __forceinline__ void commit_v2ext(__nv_bfloat16* Vs, const PosRegs<kDm>& r, int nvalid, const Slab& sl) {
#pragma unroll
    for (int it = 0; it < PosRegs<kDm>::kIt; ++it) {                                // pos_emb -> cols 0 .. pos_hi
        const int idx = threadIdx.x + it * kThreads, row = idx / PosRegs<kDm>::kPosVec, c4 = idx % PosRegs<kDm>::kPosVec;
        if (c4 * 4 < sl.pos_hi)
            *reinterpret_cast<uint2*>(Vs + row * kLdV + c4 * 4) = make_uint2(pack_bf16(r.p[it].x, r.p[it].y), pack_bf16(r.p[it].z, r.p[it].w));
    }
    if (sl.width < kE) {                                                         // zero columns of a partly filled slab
        const int zp = (kE - sl.width) >> 3;
        for (int idx = threadIdx.x; idx < kCh * zp; idx += kThreads) {
            const int row = idx / zp, part = idx - row * zp;
            *reinterpret_cast<uint4*>(Vs + row * kLdV + sl.width + part * 8) = make_uint4(0, 0, 0, 0);
        }
    }
    for (int idx = threadIdx.x; idx < kCh * 3; idx += kThreads) {             // cols 128..151: ones column then zeros
        const int row = idx / 3, part = idx % 3;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (part == 0 && row < nvalid) val.x = 0x00003F80u;                    // bf16(1.0) at column 128
        *reinterpret_cast<uint4*>(Vs + row * kLdV + kE + part * 8) = val;
    }
}

// phi of this warp's 16 rows: acc[kFp/8][4] = X[16 x D] . P^T, then relu(norm * acc) + 1e-3 on the valid features
template <int kFp, int kDm>
__device__ __forceinline__ void phi_rows(float (&acc)[kFp / 8][4], const __nv_bfloat16* Xs, const __nv_bfloat16* Ps, int warp, int lane) {
#pragma unroll
    for (int nt = 0; nt < kFp / 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const int arow = 16 * warp + (lane & 7) + 8 * ((lane >> 3) & 1), acol = 8 * (lane >> 4);
    const int brow = (lane & 7) + 8 * (lane >> 4), bcol = 8 * ((lane >> 3) & 1);       // B from [n][k] storage
    constexpr int kLdQ = ld_q(kDm);
#pragma unroll
    for (int ks = 0; ks < kDm / 16; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, smem_u32(Xs + arow * kLdQ + ks * 16 + acol));
#pragma unroll
        for (int np = 0; np < kFp / 16; ++np) {
            uint32_t b[4];      // b[0],b[1]: n-tile 2np ; b[2],b[3]: n-tile 2np+1
            ldsm_x4(b, smem_u32(Ps + (np * 16 + brow) * kLdQ + ks * 16 + bcol));
            mma16816(acc[2 * np], a, b[0], b[1]);
            mma16816(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// ---- pass A: per-chunk state sums -----------------------------------------------------------------------------
template <int kFp, int kDm>
__global__ void __launch_bounds__(kThreads)
performer_sums_mma_kernel(const __nv_bfloat16* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                          const __nv_bfloat16* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                          const float* __restrict__ pos_emb, const float* __restrict__ proj, float* __restrict__ ws,
                          int H, int T, int F, int nchunks) {
    using SM = PerfSmem<kFp, kDm>;
    constexpr int kLdQ = SM::kLdQ;
    extern __shared__ __align__(16) __nv_bfloat16 sm[];
    __nv_bfloat16 *Ps = sm, *Ks = sm + SM::kSumsK, *PhiK = Ks, *Vs = sm + SM::kSumsV;
    pdl_launch_dependents();
    pdl_wait();
    const int chunk = blockIdx.x, nh = blockIdx.y, n = nh / H, h = nh % H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int r0 = chunk * kCh, nvalid = min(kCh, T - r0);
    const Slab sl = make_slab(kDm, (int) blockIdx.z);
    {   // all global loads in flight first (cp.async tiles + both register batches), then the bf16 roundings
        ProjRegs<kFp, kDm> pr;
        PosRegs<kDm> po;
        load_rows_bf16<kDm>(Ks, kLdQ, k + (int64_t) n * k_sn + (int64_t) h * k_sh, k_st, r0, nvalid);
        issue_v2ext<kDm>(po, Vs, v + (int64_t) n * v_sn + (int64_t) h * v_sh, v_st, pos_emb, r0, nvalid, sl);
        issue_proj<kFp, kDm>(pr, proj, F);
        commit_v2ext<kDm>(Vs, po, nvalid, sl);
        commit_proj<kFp, kDm>(Ps, pr);
    }
    cp_async_wait_all();
    __syncthreads();
    const float norm = rsqrtf(sqrtf((float) kDm));
    {
        float acc[kFp / 8][4];
        phi_rows<kFp, kDm>(acc, Ks, Ps, warp, lane);
        __syncthreads();           // phi(k) is written over the K tile: every warp must have read its K rows
#pragma unroll
        for (int nt = 0; nt < kFp / 8; ++nt) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = 16 * warp + g + 8 * half;
                float o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int f = nt * 8 + 2 * tq + e;
                    float val = 0.f;
                    if (row < nvalid) val = f < F ? fmaxf(acc[nt][2 * half + e] * norm, 0.f) + 1e-3f : (f == F ? 1.0f : 0.f);
                    o[e] = val;
                }
                *reinterpret_cast<uint32_t*>(PhiK + row * (kFp + 8) + nt * 8 + 2 * tq) = pack_bf16(o[0], o[1]);
            }
        }
    }
    __syncthreads();
    // S_c[f][e] = sum_r PhiK[r][f] * V2ext[r][e]: warp w owns n-tiles {w, w+8, w+16}
    constexpr int kMT = kFp / 16;
    float acc[kMT][3][4];
#pragma unroll
    for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.f;
    const int ar = (lane & 7) + 8 * (lane >> 4), af = 8 * ((lane >> 3) & 1);        // A^T from [r][f] storage (trans)
    const int br = (lane & 7) + 8 * ((lane >> 3) & 1);                               // B from [k][n] storage (trans), x2 per n-tile
#pragma unroll
    for (int ks = 0; ks < kCh / 16; ++ks) {
        uint32_t a[kMT][4];
#pragma unroll
        for (int mt = 0; mt < kMT; ++mt) ldsm_x4_t(a[mt], smem_u32(PhiK + (ks * 16 + ar) * (kFp + 8) + mt * 16 + af));
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int nt = warp + 8 * j;
            if (nt < kEx / 8) {
                uint32_t b[4];
                // x4.trans: matrices (k 0-7, n 0-7), (k 8-15, n 0-7), and the same for the next 8 columns (unused half ok)
                ldsm_x4_t(b, smem_u32(Vs + (ks * 16 + br) * kLdV + nt * 8 + 8 * (lane >> 4)));
#pragma unroll
                for (int mt = 0; mt < kMT; ++mt) mma16816(acc[mt][j], a[mt], b[0], b[1]);
            }
        }
    }
    float* slot = ws + (((int64_t) nh * gridDim.z + blockIdx.z) * nchunks + chunk) * (kFp * kEx);
#pragma unroll
    for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int nt = warp + 8 * j;
            if (nt < kEx / 8) {
                // (rows above the ones feature F are identically zero: never written, scanned or read back)
                const int f0 = mt * 16 + g, e0 = nt * 8 + 2 * tq;
                if (f0 <= F) *reinterpret_cast<float2*>(slot + f0 * kEx + e0) = make_float2(acc[mt][j][0], acc[mt][j][1]);
                if (f0 + 8 <= F) *reinterpret_cast<float2*>(slot + (f0 + 8) * kEx + e0) = make_float2(acc[mt][j][2], acc[mt][j][3]);
            }
        }
}

// ---- pass C: outputs ------------------------------------------------------------------------------------------
template <int kFp, int kDm>
__global__ void __launch_bounds__(kThreads, (kDm <= 64 ? 2 : 1))
performer_out_mma_kernel(const __nv_bfloat16* __restrict__ q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                         const __nv_bfloat16* __restrict__ k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                         const __nv_bfloat16* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                         const float* __restrict__ pos_emb, const float* __restrict__ proj, const float* __restrict__ ws,
                         __nv_bfloat16* __restrict__ ctx, __nv_bfloat16* __restrict__ cumavg, int H, int T, int F, int nchunks, int t_off) {
    using SM = PerfSmem<kFp, kDm>;
    constexpr int kLdQ = SM::kLdQ;
    extern __shared__ __align__(16) __nv_bfloat16 sm[];
    __nv_bfloat16 *Ps = sm + SM::kP, *Qs = sm + SM::kQ, *Ks = sm + SM::kK, *PhiK = sm + SM::kPhiK, *Vs = sm + SM::kV, *Ss = sm + SM::kS;
    float* vprev_s = reinterpret_cast<float*>(sm + SM::kElems);          // [v columns of the slab] fp32
    pdl_launch_dependents();
    pdl_wait();
    const int chunk = blockIdx.x, nh = blockIdx.y, n = nh / H, h = nh % H;
    const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    // Row tile of this rt.  The causal loops below run rt + 1 steps, and warps w and w + 4 share an SM sub-partition (its tensor pipe):
    // tiles {s, 7 - s} on sub-partition s make 9 steps everywhere instead of 5 .. 11.
    const int rt = (threadIdx.x >> 5) < 4 ? (threadIdx.x >> 5) : 11 - (threadIdx.x >> 5);
    const int r0 = chunk * kCh, nvalid = min(kCh, T - r0);
    const Slab sl = make_slab(kDm, (int) blockIdx.z);
    {   // all global loads in flight first: q/k/v tiles by cp.async, then the three fp32 register batches (S_prev, pos_emb,
        // projection); only then the bf16 roundings -- one DRAM round trip per CTA instead of three
        load_rows_bf16<kDm>(Qs, kLdQ, q + (int64_t) n * q_sn + (int64_t) h * q_sh, q_st, r0, nvalid);
        load_rows_bf16<kDm>(Ks, kLdQ, k + (int64_t) n * k_sn + (int64_t) h * k_sh, k_st, r0, nvalid);
        PosRegs<kDm> po;
        issue_v2ext<kDm>(po, Vs, v + (int64_t) n * v_sn + (int64_t) h * v_sh, v_st, pos_emb, r0, nvalid, sl);
        // S_prev (exclusive prefix, fp32) -> bf16; the z column gets the +1e-6 of the reference's denominator
        const float* slot = ws + (((int64_t) nh * gridDim.z + blockIdx.z) * nchunks + chunk) * (kFp * kEx);
        constexpr int kVec = kFp * kEx / 4, kIt = (kVec + kThreads - 1) / kThreads;      // kEx % 4 == 0: a float4 never straddles rows
        float4 sv[kIt];
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const int idx = threadIdx.x + it * kThreads;
            sv[it] = (idx < kVec && idx / (kEx / 4) <= F) ? __ldcg(reinterpret_cast<const float4*>(slot) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ProjRegs<kFp, kDm> pr;
        issue_proj<kFp, kDm>(pr, proj, F);
        commit_v2ext<kDm>(Vs, po, nvalid, sl);
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const int idx = threadIdx.x + it * kThreads;
            if (idx < kVec) {
                const int f = idx / (kEx / 4), e = (idx % (kEx / 4)) * 4;
                if (e == kE && f < F) sv[it].x += 1e-6f;
                if (f == F && e >= sl.v_lo && e < sl.v_hi) *reinterpret_cast<float4*>(vprev_s + (e - sl.v_lo)) = sv[it];     // vsum of the earlier chunks, kept in fp32
                *reinterpret_cast<uint2*>(Ss + f * kLdV + e) = make_uint2(pack_bf16(sv[it].x, sv[it].y), pack_bf16(sv[it].z, sv[it].w));
            }
        }
        for (int idx = threadIdx.x; idx < kFp; idx += kThreads) *reinterpret_cast<uint4*>(Ss + idx * kLdV + kEx) = make_uint4(0, 0, 0, 0);
        commit_proj<kFp, kDm>(Ps, pr);
    }
    cp_async_wait_all();
    __syncthreads();
    const float norm = rsqrtf(sqrtf((float) kDm));
    uint32_t aq[kFp / 16][4];     // phi(q) of this rt's rows as A fragments
    {
        float acc[kFp / 8][4];
        phi_rows<kFp, kDm>(acc, Ks, Ps, rt, lane);
#pragma unroll
        for (int nt = 0; nt < kFp / 8; ++nt)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = 16 * rt + g + 8 * half;
                float o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int f = nt * 8 + 2 * tq + e;
                    o[e] = (row < nvalid && f < F) ? fmaxf(acc[nt][2 * half + e] * norm, 0.f) + 1e-3f : 0.f;
                }
                *reinterpret_cast<uint32_t*>(PhiK + row * (kFp + 8) + nt * 8 + 2 * tq) = pack_bf16(o[0], o[1]);
            }
        phi_rows<kFp, kDm>(acc, Qs, Ps, rt, lane);
#pragma unroll
        for (int ks = 0; ks < kFp / 16; ++ks) {
            float o[2][4];
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = (2 * ks + t2) * 8 + 2 * tq + (i & 1);
                    o[t2][i] = f < F ? fmaxf(acc[2 * ks + t2][i] * norm, 0.f) + 1e-3f : 0.f;
                }
            aq[ks][0] = pack_bf16(o[0][0], o[0][1]);
            aq[ks][1] = pack_bf16(o[0][2], o[0][3]);
            aq[ks][2] = pack_bf16(o[1][0], o[1][1]);
            aq[ks][3] = pack_bf16(o[1][2], o[1][3]);
        }
    }
    __syncthreads();
    float O[kEx / 8][4];        // [16 x 144]
#pragma unroll
    for (int nt = 0; nt < kEx / 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) O[nt][i] = 0.f;
    const int brow = (lane & 7) + 8 * (lane >> 4), bcol = 8 * ((lane >> 3) & 1);    // B from [n][k] storage (PhiK)
    const int vr = (lane & 7) + 8 * ((lane >> 3) & 1), vc = 8 * (lane >> 4);        // B from [k][n] storage (Vs, Ss), trans
    for (int jt = 0; jt <= rt; ++jt) {
        // P tile [16 x 16] = phi(q) . phi(k_j)^T for source rows 16jt .. 16jt+15
        float p[2][4];
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
            for (int i = 0; i < 4; ++i) p[t2][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < kFp / 16; ++ks) {
            uint32_t b[4];
            ldsm_x4(b, smem_u32(PhiK + (jt * 16 + brow) * (kFp + 8) + ks * 16 + bcol));
            mma16816(p[0], aq[ks], b[0], b[1]);
            mma16816(p[1], aq[ks], b[2], b[3]);
        }
        uint32_t pa[4];
        if (jt == rt) {       // diagonal tile: keep source j <= query i
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int col = t2 * 8 + 2 * tq + (i & 1), row = g + 8 * (i >> 1);
                    if (col > row) p[t2][i] = 0.f;
                }
        }
        pa[0] = pack_bf16(p[0][0], p[0][1]); pa[1] = pack_bf16(p[0][2], p[0][3]);
        pa[2] = pack_bf16(p[1][0], p[1][1]); pa[3] = pack_bf16(p[1][2], p[1][3]);
#pragma unroll
        for (int np = 0; np < kEx / 16; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, smem_u32(Vs + (jt * 16 + vr) * kLdV + np * 16 + vc));
            mma16816(O[2 * np], pa, b[0], b[1]);
            mma16816(O[2 * np + 1], pa, b[2], b[3]);
        }
    }
    // + phi(q) . S_prev
#pragma unroll
    for (int ks = 0; ks < kFp / 16; ++ks)
#pragma unroll
        for (int np = 0; np < kEx / 16; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, smem_u32(Ss + (ks * 16 + vr) * kLdV + np * 16 + vc));
            mma16816(O[2 * np], aq[ks], b[0], b[1]);
            mma16816(O[2 * np + 1], aq[ks], b[2], b[3]);
        }
    // denominators live in column 128 (n-tile 16, element 0 of the quad leader)
    const float den_lo = __shfl_sync(kFull, O[kE / 8][0], lane & ~3);
    const float den_hi = __shfl_sync(kFull, O[kE / 8][2], lane & ~3);
    const float inv_lo = 1.0f / den_lo, inv_hi = 1.0f / den_hi;
    const int row_lo = 16 * rt + g, row_hi = row_lo + 8;
    constexpr int kCtxW = 2 * kDm;                  // ctx row = [pos part (D) | v part (D)]; this slab owns columns 128 z ..
    __nv_bfloat16* cb = ctx + (((int64_t) n * H + h) * T + r0) * kCtxW + (int) blockIdx.z * kE;
#pragma unroll
    for (int nt = 0; nt < kE / 8; ++nt) {
        const int e0 = nt * 8 + 2 * tq;
        if (nt * 8 < sl.width) {
            if (row_lo < nvalid) *reinterpret_cast<uint32_t*>(cb + (int64_t) row_lo * kCtxW + e0) = pack_bf16(O[nt][0] * inv_lo, O[nt][1] * inv_lo);
            if (row_hi < nvalid) *reinterpret_cast<uint32_t*>(cb + (int64_t) row_hi * kCtxW + e0) = pack_bf16(O[nt][2] * inv_hi, O[nt][3] * inv_hi);
        }
    }
    // a13 running mean of v (attention.py:1237-1241), fused: cumavg[t] = (vsum_prev(chunk) + sum_{j <= t in chunk} v_j) / (t + 1).
    // The in-chunk cumulative sum is L . V with L the lower-triangular ones matrix: the same V fragments as above against an
    // all-ones A fragment (triangular on the diagonal tile); bf16 ones x bf16 v accumulate exactly in fp32.  vsum_prev is row F
    // (the ones feature) of the prefixed state, columns 64..127, kept in fp32 in shared memory by the S_prev load above.
    if (cumavg != nullptr && sl.v_hi > sl.v_lo) {
        constexpr int kCT = (kDm < kE ? kDm : kE) / 8;          // n-tiles of v columns a slab can hold
        const int vtiles = (sl.v_hi - sl.v_lo) >> 3;           // ... and holds (even)
        float C[kCT][4];
#pragma unroll
        for (int nt = 0; nt < kCT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) C[nt][i] = 0.f;
        constexpr uint32_t kOne2 = 0x3F803F80u;                    // two bf16 ones
        for (int jt = 0; jt <= rt; ++jt) {
            uint32_t la[4] = {kOne2, kOne2, kOne2, kOne2};
            if (jt == rt) {                                      // diagonal tile: source column <= query row
                const uint32_t tri = ((2 * tq <= g) ? 0x00003F80u : 0u) | ((2 * tq + 1 <= g) ? 0x3F800000u : 0u);
                la[0] = tri; la[1] = kOne2; la[2] = 0u; la[3] = tri;
            }
#pragma unroll
            for (int np = 0; np < kCT / 2; ++np) {
                if (2 * np < vtiles) {
                    uint32_t b[4];
                    ldsm_x4_t(b, smem_u32(Vs + (jt * 16 + vr) * kLdV + sl.v_lo + np * 16 + vc));
                    mma16816(C[2 * np], la, b[0], b[1]);
                    mma16816(C[2 * np + 1], la, b[2], b[3]);
                }
            }
        }
        // (t_off: absolute position of row 0 when the call covers a later range of the sequence)
        const float r_lo = 1.0f / (float) (t_off + r0 + row_lo + 1), r_hi = 1.0f / (float) (t_off + r0 + row_hi + 1);
        __nv_bfloat16* ab = cumavg + (((int64_t) n * H + h) * T + r0) * kDm + sl.v_off;
#pragma unroll
        for (int nt = 0; nt < kCT; ++nt) {
            const int e0 = nt * 8 + 2 * tq;
            if (nt < vtiles) {
                const float2 pv = *reinterpret_cast<const float2*>(vprev_s + e0);
                if (row_lo < nvalid) *reinterpret_cast<uint32_t*>(ab + (int64_t) row_lo * kDm + e0) = pack_bf16((C[nt][0] + pv.x) * r_lo, (C[nt][1] + pv.y) * r_lo);
                if (row_hi < nvalid) *reinterpret_cast<uint32_t*>(ab + (int64_t) row_hi * kDm + e0) = pack_bf16((C[nt][2] + pv.x) * r_hi, (C[nt][3] + pv.y) * r_hi);
            }
        }
    }
}

// exclusive prefix over the chunk slots of one (n, h).  The loads of a batch of chunks are issued together (they are
// independent; a naive load/store loop serialises on aliasing and costs one L2 round trip per chunk).
__global__ void __launch_bounds__(256)
prefix_chunks_kernel(float* __restrict__ ws, int nchunks, int64_t stride, int64_t used4, const float* __restrict__ init, float* __restrict__ total, int write_prefix) {
    // init (nullable, [gridDim.y][stride]): the state BEFORE the first chunk (a rank's share of a sequence sharded over ranks starts from
    // the sums of the ranks before it); total (nullable): init + all chunks; write_prefix = 0: only `total` is produced (the chunk slots
    // keep their per-chunk sums for a later prefix pass with the exchanged init)
    pdl_launch_dependents();
    pdl_wait();
    // 16-byte accesses (stride = kFp * kEx is a multiple of 4): a quarter of the load/store instructions, same bytes in flight
    float4* base = reinterpret_cast<float4*>(ws + (int64_t) blockIdx.y * nchunks * stride);
    const int64_t stride4 = stride >> 2;
    constexpr int kBatch = 32;
    // used4: float4s of a slot that carry data (feature rows 0 .. F; the padding rows up to kFp are never written)
    for (int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; idx < used4; idx += (int64_t) gridDim.x * blockDim.x) {
        float4 run = init ? __ldg(reinterpret_cast<const float4*>(init) + (int64_t) blockIdx.y * stride4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = 0; c0 < nchunks; c0 += kBatch) {
            float4 cur[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) cur[i] = (c0 + i < nchunks) ? __ldcg(base + (int64_t) (c0 + i) * stride4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                if (write_prefix && c0 + i < nchunks) __stcg(base + (int64_t) (c0 + i) * stride4 + idx, run);
                run.x += cur[i].x; run.y += cur[i].y; run.z += cur[i].z; run.w += cur[i].w;
            }
        }
        if (total) reinterpret_cast<float4*>(total)[(int64_t) blockIdx.y * stride4 + idx] = run;
    }
}

// phase 0: the whole stage (sums, prefix from zero, outputs).  Sequence sharded over ranks: phase 1 = per-chunk sums of this
// rank's range + their total (-> exchanged between the ranks), phase 2 = prefix starting from `init` (the sums of everything before the
// range) + outputs; `ws` carries the per-chunk sums from phase 1 to phase 2.
template <int kFp, int kDm>
int launch_performer_mma(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st, const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                         const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, const float* pos_emb, const float* proj,
                         void* ctx, void* cumavg, float* ws, int N, int H, int T, int F, cudaStream_t s,
                         int phase = 0, const float* init = nullptr, float* total = nullptr, int t_off = 0) {
    using SM = PerfSmem<kFp, kDm>;
    const int nchunks = (T + kCh - 1) / kCh;
    constexpr int kSlabs = n_slabs(kDm);
    dim3 grid(nchunks, N * H, kSlabs);
    auto ka = performer_sums_mma_kernel<kFp, kDm>;
    auto kc = performer_out_mma_kernel<kFp, kDm>;
    SEA_CUDA_TRY(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kSumsBytes), "smem attr");
    SEA_CUDA_TRY(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes), "smem attr");
    using B = __nv_bfloat16;
    const int64_t stride = (int64_t) kFp * kEx;
    const int64_t used4 = (int64_t) (F + 1) * (kEx / 4);
    const dim3 pgrid((unsigned) ((used4 + 255) / 256), N * H * kSlabs);
    if (phase == 0 || phase == 1)
        SEA_CUDA_TRY(launch_pdl(ka, grid, dim3(kThreads), (size_t) SM::kSumsBytes, s, (const B*) k, k_sn, k_sh, k_st, (const B*) v, v_sn, v_sh, v_st, pos_emb, proj, ws, H, T, F,
                                nchunks), "performer_sums_mma_kernel launch");
    if (phase == 1) {
        SEA_CUDA_TRY(launch_pdl(prefix_chunks_kernel, pgrid, dim3(256), (size_t) 0, s, ws, nchunks, stride, used4, (const float*) nullptr, total, 0), "prefix_chunks_kernel launch");
        return SEA_OK;
    }
    // one exclusive prefix per (n, h, slab): the slabs' chunk slots are laid out [nh][slab][chunk]
    SEA_CUDA_TRY(launch_pdl(prefix_chunks_kernel, pgrid, dim3(256), (size_t) 0, s, ws, nchunks, stride, used4, init, total, 1), "prefix_chunks_kernel launch");
    SEA_CUDA_TRY(launch_pdl(kc, grid, dim3(kThreads), (size_t) SM::kBytes, s, (const B*) q, q_sn, q_sh, q_st, (const B*) k, k_sn, k_sh, k_st, (const B*) v, v_sn, v_sh, v_st,
                            pos_emb, proj, (const float*) ws, (B*) ctx, (B*) cumavg, H, T, F, nchunks, t_off), "performer_out_mma_kernel launch");
    return SEA_OK;
}

}  // namespace
}  // namespace sea

using namespace sea;

extern "C" {

// feature counts the kernels are instantiated for, per head dim (Fp = F + 1 rounded up to 16; F = int(D ln D / nb_factor))
int sea_performer_mma_supported(int dtype, int D, int F) {
    if (dtype != SEA_DTYPE_BF16 || F < 1) return 0;
    const int Fp = ((F + 1) + 15) & ~15;
    switch (D) {
        case 32: return Fp <= 32;
        case 64: return Fp <= 64;
        case 80: return Fp >= 32 && Fp <= 64;
        case 96: return Fp == 64;
        case 128: return Fp >= 48 && Fp <= 80;
    }
    return 0;
}

int64_t sea_performer_mma_workspace_floats(int N, int H, int T, int D, int F) {
    if (N <= 0 || H <= 0 || T <= 0 || F <= 0 || D <= 0) return 0;
    const int Fp = ((F + 1) + 15) & ~15;
    return (int64_t) N * H * ((2 * D + kE - 1) / kE) * ((T + kCh - 1) / kCh) * Fp * kEx;
}

int sea_performer_causal_mma_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                 const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                 const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                 const float* pos_emb, const float* proj, void* ctx, void* cumavg, float* workspace,
                                 int N, int H, int T, int D, int F, void* stream) {
    return sea_performer_causal_mma_range(q, q_sn, q_sh, q_st, k, k_sn, k_sh, k_st, v, v_sn, v_sh, v_st, pos_emb, proj, ctx, cumavg, workspace, nullptr, nullptr,
                                          N, H, T, D, F, 0, 0, stream);
}

int64_t sea_performer_mma_state_floats(int N, int H, int D, int F) {
    if (N <= 0 || H <= 0 || F <= 0 || D <= 0) return 0;
    const int Fp = ((F + 1) + 15) & ~15;
    return (int64_t) N * H * ((2 * D + kE - 1) / kE) * Fp * kEx;
}

int sea_performer_causal_mma_range(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                   const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                   const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                   const float* pos_emb, const float* proj, void* ctx, void* cumavg, float* workspace,
                                   const float* init, float* total, int N, int H, int T, int D, int F, int t_off, int phase, void* stream) {
    SEA_CHECK_ARG(k && v && pos_emb && proj && workspace && (phase == 1 || (q && ctx)) && phase >= 0 && phase <= 2 && t_off >= 0 &&
                  (phase != 1 || total), "sea_performer_causal_mma_fwd: null pointer");
    if (phase == 1) { q = k; q_sn = k_sn; q_sh = k_sh; q_st = k_st; }
    if (!sea_performer_mma_supported(SEA_DTYPE_BF16, D, F)) {
        set_error("sea_performer_causal_mma_fwd: unsupported shape D=%d F=%d", D, F);
        return SEA_ERR_UNSUPPORTED;
    }
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && (int64_t) N * H <= 65535, "sea_performer_causal_mma_fwd: bad shape");
    SEA_CHECK_ARG(((q_sn | q_sh | q_st | k_sn | k_sh | k_st | v_sn | v_sh | v_st) % 8) == 0 &&
                  ((((uintptr_t) q) | ((uintptr_t) k) | ((uintptr_t) v) | ((uintptr_t) pos_emb) | ((uintptr_t) ctx) | ((uintptr_t) proj) | ((uintptr_t) workspace)) & 15) == 0,
                  "sea_performer_causal_mma_fwd: q/k/v rows must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t) stream;
    const int Fp = ((F + 1) + 15) & ~15;
#define SEA_PF_CASE(FF, DD)                                                                                                                   \
    if (Fp == FF && D == DD)                                                                                                                  \
        return launch_performer_mma<FF, DD>(q, q_sn, q_sh, q_st, k, k_sn, k_sh, k_st, v, v_sn, v_sh, v_st, pos_emb, proj, ctx, cumavg, workspace, N, H, T, F, s, \
                                            phase, init, total, t_off)
    SEA_PF_CASE(16, 32); SEA_PF_CASE(32, 32);
    SEA_PF_CASE(16, 64); SEA_PF_CASE(32, 64); SEA_PF_CASE(48, 64); SEA_PF_CASE(64, 64);
    SEA_PF_CASE(32, 80); SEA_PF_CASE(48, 80); SEA_PF_CASE(64, 80);
    SEA_PF_CASE(64, 96);
    SEA_PF_CASE(48, 128); SEA_PF_CASE(64, 128); SEA_PF_CASE(80, 128);
#undef SEA_PF_CASE
    set_error("sea_performer_causal_mma_fwd: no kernel for D=%d, padded feature count %d", D, Fp);
    return SEA_ERR_UNSUPPORTED;
}

}  // extern "C"
