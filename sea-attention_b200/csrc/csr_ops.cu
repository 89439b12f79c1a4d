// Integer / index stages of the SEA hot path and the standalone flat-CSR operators.
//   a7  grouped top-k            (reference attention.py:774-947)
//   a8  resize_from_m_to_t_csr   (reference ops/kernels/causal_resize_m_to_t.py:493-572, 648-762, 910-1007)
//   a9-a12 flat_csr_{masked_bmm,softmax,elmul,sdbmm}, flat_csr_to_dense, dense resize_from_m_to_t
// All of these are HBM/L2-bound gather or bit work: warp-per-row kernels, shuffle scans, 128-bit loads
// where rows are contiguous.  No tensor cores on purpose.
#include "common.cuh"
#include "topk.cuh"
#include "csr_common.cuh"

#include <stdarg.h>

namespace sea {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

__global__ void __launch_bounds__(kTopkThreads)
topk_mask_bits_kernel(const float* __restrict__ keys, int64_t sn, int64_t sh, int64_t st,
                      const float* __restrict__ k_per_group, const uint8_t* __restrict__ row_valid,
                      uint32_t* __restrict__ mask_bits, int N, int H, int T, int P, int group_mode) {
    extern __shared__ uint32_t smem_u32[];
    __shared__ int hist[256];
    __shared__ int scratch[16];
    const int tid = threadIdx.x;
    int n, t, h0, G;
    float kf;
    if (group_mode == 0) {
        n = blockIdx.x / T; t = blockIdx.x % T; h0 = 0; G = H * P;
        kf = k_per_group[blockIdx.x];
    } else {
        n = blockIdx.x / (H * T); h0 = (blockIdx.x / T) % H; t = blockIdx.x % T; G = P;
        kf = k_per_group[n];
    }
    const int words_per_row = (H * P + 31) >> 5;
    uint32_t* out_row = mask_bits + ((int64_t) n * T + t) * words_per_row;
    const bool valid = row_valid == nullptr || row_valid[(int64_t) n * T + t] != 0;
    uint32_t* skeys = smem_u32;
    uint32_t* sbits = smem_u32 + ((G + 31) & ~31);
    const float* base = keys + (int64_t) n * sn + (int64_t) t * st;
    for (int i = tid; i < G; i += kTopkThreads) {
        int h = h0 + i / P, m = i % P;
        skeys[i] = orderable(base[(int64_t) h * sh + m]);
    }
    __syncthreads();
    // rank < K with K fp32 integer valued (attention.py:916): alive count = min(ceil(K), G)
    int K = valid ? (int) fminf(ceilf(kf), (float) G) : 0;
    topk_select_to_bits(skeys, G, K, sbits, hist, scratch);
    __syncthreads();
    const int nwords = (G + 31) >> 5;
    if (group_mode == 0) {
        for (int w = tid; w < nwords; w += kTopkThreads) out_row[w] = sbits[w];
    } else {
        // P keys of head h0 land at bit offset h0*P of the row: merge with atomics when P % 32 != 0
        const int bit0 = h0 * P;
        if ((P & 31) == 0) {
            for (int w = tid; w < nwords; w += kTopkThreads) out_row[(bit0 >> 5) + w] = sbits[w];
        } else {
            for (int i = tid; i < G; i += kTopkThreads) {
                if ((sbits[i >> 5] >> (i & 31)) & 1u) atomicOr(&out_row[(bit0 + i) >> 5], 1u << ((bit0 + i) & 31));
            }
        }
    }
}

__global__ void mask_float_to_bits_kernel(const float* __restrict__ mask, int64_t sn, int64_t sh, int64_t st,
                                          uint32_t* __restrict__ bits, int N, int H, int T, int P, int words_per_row) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total = (int64_t) N * T * words_per_row;
    if (warp >= total) return;
    const int w = (int) (warp % words_per_row);
    const int64_t row = warp / words_per_row;
    const int n = (int) (row / T), t = (int) (row % T);
    const int i = (w << 5) + lane;
    bool on = false;
    if (i < H * P) {
        int h = i / P, m = i % P;
        on = ((int) mask[(int64_t) n * sn + (int64_t) h * sh + (int64_t) t * st + m]) != 0;
    }
    uint32_t word = __ballot_sync(kFull, on);
    if (lane == 0) bits[warp] = word;
}

__global__ void mask_bits_to_float_kernel(const uint32_t* __restrict__ bits, float* __restrict__ mask,
                                          int N, int H, int T, int P, int words_per_row) {
    const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t) N * H * T * P;
    if (idx >= total) return;
    const int m = (int) (idx % P);
    const int t = (int) ((idx / P) % T);
    const int h = (int) ((idx / ((int64_t) P * T)) % H);
    const int n = (int) (idx / ((int64_t) P * T * H));
    const int i = h * P + m;
    uint32_t word = bits[((int64_t) n * T + t) * words_per_row + (i >> 5)];
    mask[idx] = (float) ((word >> (i & 31)) & 1u);
}

constexpr int kCsrWarps = 8;
constexpr int kCsrThreads = 256;

// CTA per query row, thread per 32-pixel word: the early causal rows keep all H*P pixels alive, so a warp-per-row
// mapping leaves one warp with thousands of serial pixels on the critical path.
template <typename IdxT>
__global__ void __launch_bounds__(kCsrThreads)
csr_count_kernel(const uint32_t* __restrict__ bits, IdxT* __restrict__ crow,
                 int N, int H, int T_DST, int P, int T_SRC, int k, int is_causal, int words_per_row, const int32_t* __restrict__ lengths) {
    __shared__ int wsum[kCsrThreads / 32];
    const int64_t row = blockIdx.x;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    // non-causal: the interpolation width is T_SRC, or the item's token length for a right-padded batch (resize_m_to_t.py:36-47)
    const int L = is_causal ? (T_SRC - T_DST + t + 1) : (lengths ? max(1, min(lengths[n], T_SRC)) : T_SRC);
    const float s = __fdiv_rn((float) L, (float) P);
    const uint32_t* rb = bits + row * words_per_row;
    int cnt = 0;
    for (int w = threadIdx.x; w < words_per_row; w += kCsrThreads) cnt += word_width_sum(rb[w], w, P, s, k);
    cnt = warp_sum_i(cnt);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int i = 0; i < kCsrThreads / 32; ++i) tot += wsum[i];
        crow[(int64_t) n * (T_DST + 1) + t + 1] = (IdxT) tot;
        if (t == 0) crow[(int64_t) n * (T_DST + 1)] = 0;
    }
}

// in-place inclusive scan of crow[n, 1..T_DST] (one CTA per batch item)
template <typename IdxT>
__global__ void __launch_bounds__(1024)
crow_scan_kernel(IdxT* __restrict__ crow, int T_DST) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry_s;
    IdxT* c = crow + (int64_t) blockIdx.x * (T_DST + 1) + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < T_DST; base += 1024) {
        int i = base + threadIdx.x;
        long long v = i < T_DST ? (long long) c[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long nb = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += nb;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        long long wprefix = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            long long sw = warp_sums[w];
            if (w < wid) wprefix += sw;
            tot += sw;
        }
        long long carry = carry_s;
        if (i < T_DST) c[i] = (IdxT) (carry + wprefix + incl);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
}

// CTA per query row: block-wide exclusive scan of the per-word entry counts gives every thread the offset of its
// word's first entry.  Optionally emits head_ptr[n, t, 0..H] (absolute entry offset where head h starts in the row),
// which the fused attention kernel uses instead of binary searches (only when P % 32 == 0: a head starts on a word).
template <typename IdxT>
__global__ void __launch_bounds__(kCsrThreads)
csr_fill_kernel(const uint32_t* __restrict__ bits, const IdxT* __restrict__ crow, IdxT* __restrict__ col, int64_t Z,
                int32_t* __restrict__ head_ptr, int N, int H, int T_DST, int P, int T_SRC, int k, int is_causal, int words_per_row,
                const int32_t* __restrict__ lengths) {
    __shared__ int wsum[kCsrThreads / 32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t row = blockIdx.x;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    const int L = is_causal ? (T_SRC - T_DST + t + 1) : (lengths ? max(1, min(lengths[n], T_SRC)) : T_SRC);
    const float s = __fdiv_rn((float) L, (float) P);
    const uint32_t* rb = bits + row * words_per_row;
    IdxT* out = col + (int64_t) n * Z;
    const int64_t row_start = (int64_t) crow[(int64_t) n * (T_DST + 1) + t];
    const int words_per_head = P >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int w0 = 0; w0 < words_per_row; w0 += kCsrThreads) {
        const int w = w0 + threadIdx.x;
        const uint32_t word = w < words_per_row ? rb[w] : 0u;
        const int mine = word_width_sum(word, w, P, s, k);
        const int incl = warp_scan_incl_i(mine, lane);
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        int wprefix = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < kCsrThreads / 32; ++i) { const int sw = wsum[i]; if (i < wid) wprefix += sw; tot += sw; }
        const int carry = carry_s;
        int64_t p = row_start + carry + wprefix + incl - mine;
        if (head_ptr != nullptr && w < words_per_row && (w % words_per_head) == 0)
            head_ptr[row * (H + 1) + w / words_per_head] = (int32_t) p;
        for (uint32_t x = word; x; x &= x - 1) {
            const int i = (w << 5) + __ffs(x) - 1;
            const int h = pix_h(i, P), m = pix_m(i, P);
            float vs, ve;
            pixel_bounds(s, m, vs, ve);
            const float span = __fsub_rn(ve, vs);
            const int wd = min((int) span, k);
            if (wd <= 0) continue;
            const int64_t top = (int64_t) h * T_SRC + (int64_t) ve - 1;
            if (p + wd > Z) continue;   // caller under-allocated col: never write out of bounds
            if (wd == (int) span) {
                for (int j = 0; j < wd; ++j) out[p + j] = (IdxT) (top - j);
            } else {
                // clamped to k: sub-sample the span (reference :569); the reference divides with the
                // approximate div.full.f32, IEEE division here (= its numpy-interpreted run).
                const float ratio = __fdiv_rn(span, (float) wd);
                for (int j = 0; j < wd; ++j) out[p + j] = (IdxT) (top - (int) __fmul_rn((float) j, ratio));
            }
            p += wd;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (head_ptr != nullptr && threadIdx.x == 0) head_ptr[row * (H + 1) + H] = (int32_t) (row_start + carry_s);
    // zero the tail [crow[n, T_DST], Z) (the reference allocates col with torch.zeros, :669); every CTA of the batch
    // item takes a strided slice so that an over-allocated col costs one coalesced pass
    const int64_t tail_begin = (int64_t) crow[(int64_t) n * (T_DST + 1) + T_DST];
    for (int64_t z = tail_begin + (int64_t) t * kCsrThreads + threadIdx.x; z < Z; z += (int64_t) T_DST * kCsrThreads) out[z] = 0;
}

template <typename IdxT>
__global__ void __launch_bounds__(kCsrWarps * 32)
csr_to_dense_kernel(const IdxT* __restrict__ crow, const IdxT* __restrict__ col, const float* __restrict__ values,
                    int64_t Z, float* __restrict__ out, int N, int H, int T_DST, int T_SRC) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t) blockIdx.x * kCsrWarps + (threadIdx.x >> 5);
    if (row >= (int64_t) N * T_DST) return;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    const int64_t s = crow[(int64_t) n * (T_DST + 1) + t], e = crow[(int64_t) n * (T_DST + 1) + t + 1];
    for (int64_t z = s + lane; z < e; z += 32) {
        int64_t c = col[(int64_t) n * Z + z];
        int h = (int) (c / T_SRC), j = (int) (c % T_SRC);
        out[(((int64_t) n * H + h) * T_DST + t) * T_SRC + j] = values ? values[(int64_t) n * Z + z] : 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// standalone flat-CSR operators (a9-a12).  Warp per row, lane per entry.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float dot_rows(const T* __restrict__ a, const T* __restrict__ b, int D) {
    float acc = 0.f;
    for (int c = 0; c < D; ++c) acc = fmaf(to_f32(a[c]), to_f32(b[c]), acc);
    return acc;
}
template <>
__device__ __forceinline__ float dot_rows<float>(const float* __restrict__ a, const float* __restrict__ b, int D) {
    float acc = 0.f;
    if ((D & 3) == 0 && ((((uintptr_t) a) | ((uintptr_t) b)) & 15) == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (int c = 0; c < (D >> 2); ++c) {
            float4 x = a4[c], y = b4[c];
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
        }
    } else {
        for (int c = 0; c < D; ++c) acc = fmaf(a[c], b[c], acc);
    }
    return acc;
}
template <>
__device__ __forceinline__ float dot_rows<__nv_bfloat16>(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int D) {
    float acc = 0.f;
    if ((D & 7) == 0 && ((((uintptr_t) a) | ((uintptr_t) b)) & 15) == 0) {
        const uint4* a4 = reinterpret_cast<const uint4*>(a);
        const uint4* b4 = reinterpret_cast<const uint4*>(b);
        for (int c = 0; c < (D >> 3); ++c) {
            uint4 x = a4[c], y = b4[c];
            const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc = fmaf(__uint_as_float(xs[q] << 16), __uint_as_float(ys[q] << 16), acc);
                acc = fmaf(__uint_as_float(xs[q] & 0xffff0000u), __uint_as_float(ys[q] & 0xffff0000u), acc);
            }
        }
    } else {
        for (int c = 0; c < D; ++c) acc = fmaf(__bfloat162float(a[c]), __bfloat162float(b[c]), acc);
    }
    return acc;
}

template <typename T, typename IdxT>
__global__ void __launch_bounds__(kCsrWarps * 32)
csr_masked_bmm_kernel(const IdxT* __restrict__ crow, const IdxT* __restrict__ col, int64_t Z,
                      const T* __restrict__ a, int64_t a_sn, int64_t a_sh, int64_t a_st,
                      const T* __restrict__ b, int64_t b_sn, int64_t b_sh, int64_t b_st,
                      float* __restrict__ out, int N, int H, int T_DST, int T_SRC, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t) blockIdx.x * kCsrWarps + (threadIdx.x >> 5);
    if (row >= (int64_t) N * T_DST) return;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    const int64_t s = crow[(int64_t) n * (T_DST + 1) + t], e = crow[(int64_t) n * (T_DST + 1) + t + 1];
    for (int64_t z = s + lane; z < e; z += 32) {
        int64_t c = col[(int64_t) n * Z + z];
        int h = (int) (c / T_SRC), j = (int) (c % T_SRC);
        const T* ap = a + (int64_t) n * a_sn + (int64_t) h * a_sh + (int64_t) t * a_st;
        const T* bp = b + (int64_t) n * b_sn + (int64_t) h * b_sh + (int64_t) j * b_st;
        out[(int64_t) n * Z + z] = dot_rows<T>(ap, bp, D);
    }
}

__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// CTA per row; per-head max / sum in shared memory (entry order inside the row is NOT assumed).
template <typename IdxT>
__global__ void __launch_bounds__(128)
csr_softmax_kernel(const IdxT* __restrict__ crow, const IdxT* __restrict__ col, int64_t Z,
                   const float* __restrict__ in, float* __restrict__ out, int N, int H, int T_DST, int T_SRC) {
    extern __shared__ int smem_i[];
    int* hmax = smem_i;
    float* hsum = reinterpret_cast<float*>(smem_i + H);
    const int64_t row = blockIdx.x;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    const int64_t s = crow[(int64_t) n * (T_DST + 1) + t], e = crow[(int64_t) n * (T_DST + 1) + t + 1];
    for (int h = threadIdx.x; h < H; h += blockDim.x) { hmax[h] = f2ord(-INFINITY); hsum[h] = 0.f; }
    __syncthreads();
    for (int64_t z = s + threadIdx.x; z < e; z += blockDim.x) {
        int h = (int) (col[(int64_t) n * Z + z] / T_SRC);
        atomicMax(&hmax[h], f2ord(in[(int64_t) n * Z + z]));
    }
    __syncthreads();
    for (int64_t z = s + threadIdx.x; z < e; z += blockDim.x) {
        int h = (int) (col[(int64_t) n * Z + z] / T_SRC);
        atomicAdd(&hsum[h], expf(in[(int64_t) n * Z + z] - ord2f(hmax[h])));
    }
    __syncthreads();
    for (int64_t z = s + threadIdx.x; z < e; z += blockDim.x) {
        int h = (int) (col[(int64_t) n * Z + z] / T_SRC);
        out[(int64_t) n * Z + z] = expf(in[(int64_t) n * Z + z] - ord2f(hmax[h])) / hsum[h];
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(kCsrWarps * 32)
csr_elmul_kernel(const IdxT* __restrict__ crow, const IdxT* __restrict__ col, int64_t Z,
                 const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ dense,
                 int64_t d_sn, int64_t d_sh, int64_t d_st, int64_t d_sj, int N, int H, int T_DST, int T_SRC) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t) blockIdx.x * kCsrWarps + (threadIdx.x >> 5);
    if (row >= (int64_t) N * T_DST) return;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    const int64_t s = crow[(int64_t) n * (T_DST + 1) + t], e = crow[(int64_t) n * (T_DST + 1) + t + 1];
    for (int64_t z = s + lane; z < e; z += 32) {
        int64_t c = col[(int64_t) n * Z + z];
        int h = (int) (c / T_SRC), j = (int) (c % T_SRC);
        out[(int64_t) n * Z + z] = in[(int64_t) n * Z + z] * dense[n * d_sn + h * d_sh + t * d_st + j * d_sj];
    }
}

// warp per row; lanes own output dims; entries walked in order, flushed on head change (+= so a head
// split into several segments is still summed correctly).
template <typename T, typename IdxT>
__global__ void __launch_bounds__(kCsrWarps * 32)
csr_sdbmm_kernel(const IdxT* __restrict__ crow, const IdxT* __restrict__ col, int64_t Z, const float* __restrict__ values,
                 const T* __restrict__ v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                 float* __restrict__ out, int N, int H, int T_DST, int T_SRC, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t) blockIdx.x * kCsrWarps + (threadIdx.x >> 5);
    if (row >= (int64_t) N * T_DST) return;
    const int n = (int) (row / T_DST), t = (int) (row % T_DST);
    const int64_t s = crow[(int64_t) n * (T_DST + 1) + t], e = crow[(int64_t) n * (T_DST + 1) + t + 1];
    constexpr int kMaxPerLane = 8;  // D <= 256
    float acc[kMaxPerLane];
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) acc[i] = 0.f;
    int cur_h = -1;
    auto flush = [&](int h) {
        if (h < 0) return;
        float* o = out + (((int64_t) n * H + h) * T_DST + t) * D;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            int dcol = lane + 32 * i;
            if (dcol < D) { o[dcol] += acc[i]; acc[i] = 0.f; }
        }
    };
    for (int64_t z = s; z < e; ++z) {
        int64_t c = col[(int64_t) n * Z + z];
        int h = (int) (c / T_SRC), j = (int) (c % T_SRC);
        if (h != cur_h) { flush(cur_h); cur_h = h; }
        const float p = values[(int64_t) n * Z + z];
        const T* vp = v + (int64_t) n * v_sn + (int64_t) h * v_sh + (int64_t) j * v_st;
#pragma unroll
        for (int i = 0; i < kMaxPerLane; ++i) {
            int dcol = lane + 32 * i;
            if (dcol < D) acc[i] = fmaf(p, to_f32(vp[dcol]), acc[i]);
        }
    }
    flush(cur_h);
}

// a16 dense resize: CTA per (n, t) row of the additive mask; block scan of the valid flags gives the
// 1-based rank cs of each source token, then every (h, j) gathers one compressed pixel.
__global__ void __launch_bounds__(256)
resize_dense_kernel(const float* __restrict__ x, float fill, const float* __restrict__ amask, int64_t m_sn, int64_t m_st,
                    float* __restrict__ out, int N, int H, int T1, int P, int T2) {
    extern __shared__ int s_idx[];  // [T2] pixel index per source token (P = pad pixel)
    __shared__ int warp_sums[8];
    __shared__ int carry_s, total_s;
    const int n = blockIdx.x / T1, t = blockIdx.x % T1;
    const float* mrow = amask + (int64_t) n * m_sn + (int64_t) t * m_st;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < T2; base += 256) {
        int j = base + threadIdx.x;
        int valid = (j < T2 && mrow[j] > -1.0f) ? 1 : 0;
        int incl = warp_scan_incl_i(valid, lane);
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        int wprefix = 0, tot = 0;
        for (int w = 0; w < 8; ++w) { int sw = warp_sums[w]; if (w < wid) wprefix += sw; tot += sw; }
        int cs = carry_s + wprefix + incl;
        if (j < T2) s_idx[j] = valid ? cs : -1;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) total_s = carry_s;
    __syncthreads();
    const float Lf = (float) total_s;
    for (int j = threadIdx.x; j < T2; j += 256) {
        int cs = s_idx[j];
        int idx;
        if (cs < 0) idx = P;
        else {
            // floor(((cs - 1) + 0.5) / L * P - 1e-4) in un-fused fp32 (resize_m_to_t.py:46)
            float a = __fadd_rn(__fsub_rn((float) cs, 1.0f), 0.5f);
            float b = __fmul_rn(__fdiv_rn(a, Lf), (float) P);
            idx = (int) floorf(__fsub_rn(b, 1e-4f));
            idx = max(0, min(idx, P));
        }
        s_idx[j] = idx;
    }
    __syncthreads();
    for (int h = 0; h < H; ++h) {
        const float* xr = x + (((int64_t) n * H + h) * T1 + t) * P;
        float* orow = out + (((int64_t) n * H + h) * T1 + t) * T2;
        for (int j = threadIdx.x; j < T2; j += 256) {
            int idx = s_idx[j];
            orow[j] = idx >= P ? fill : xr[idx];
        }
    }
}

}  // namespace sea

using namespace sea;

extern "C" {

int sea_abi_version(void) { return SEA_ABI_VERSION; }
const char* sea_last_error(void) { return g_err; }

int sea_device_arch(void) {
    int dev = 0;
    cudaDeviceProp prop;
    SEA_CUDA_TRY(cudaGetDevice(&dev), "cudaGetDevice");
    SEA_CUDA_TRY(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    return prop.major * 10 + prop.minor;
}

int sea_topk_mask_bits(const float* keys, int64_t sn, int64_t sh, int64_t st, const float* k_per_group,
                       const uint8_t* row_valid, uint32_t* mask_bits, int N, int H, int T, int P, int group_mode,
                       void* stream) {
    SEA_CHECK_ARG(keys && k_per_group && mask_bits, "sea_topk_mask_bits: null pointer");
    SEA_CHECK_ARG(group_mode == 0 || group_mode == 1, "sea_topk_mask_bits: group_mode %d unsupported", group_mode);
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && P > 0, "sea_topk_mask_bits: bad shape");
    const int G = group_mode == 0 ? H * P : P;
    const size_t smem = ((size_t) ((G + 31) & ~31) + ((G + 31) >> 5)) * sizeof(uint32_t);
    SEA_CHECK_ARG(smem <= 200 * 1024, "sea_topk_mask_bits: group of %d keys does not fit shared memory", G);
    cudaStream_t s = (cudaStream_t) stream;
    const int words_per_row = (H * P + 31) >> 5;
    if (group_mode == 1 && (P & 31) != 0)
        SEA_CUDA_TRY(cudaMemsetAsync(mask_bits, 0, (size_t) N * T * words_per_row * 4, s), "memset");
    SEA_CUDA_TRY(cudaFuncSetAttribute(topk_mask_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
    const int64_t groups = group_mode == 0 ? (int64_t) N * T : (int64_t) N * H * T;
    topk_mask_bits_kernel<<<(unsigned) groups, kTopkThreads, smem, s>>>(keys, sn, sh, st, k_per_group, row_valid,
                                                                         mask_bits, N, H, T, P, group_mode);
    SEA_CHECK_LAUNCH("topk_mask_bits_kernel");
    return SEA_OK;
}

int sea_mask_float_to_bits(const float* mask, int64_t sn, int64_t sh, int64_t st, uint32_t* mask_bits,
                           int N, int H, int T, int P, void* stream) {
    SEA_CHECK_ARG(mask && mask_bits && N > 0 && H > 0 && T > 0 && P > 0, "sea_mask_float_to_bits: bad argument");
    const int wpr = (H * P + 31) >> 5;
    const int64_t warps = (int64_t) N * T * wpr;
    mask_float_to_bits_kernel<<<cdiv(warps * 32, 256), 256, 0, (cudaStream_t) stream>>>(mask, sn, sh, st, mask_bits, N, H, T, P, wpr);
    SEA_CHECK_LAUNCH("mask_float_to_bits_kernel");
    return SEA_OK;
}

int sea_mask_bits_to_float(const uint32_t* mask_bits, float* mask, int N, int H, int T, int P, void* stream) {
    SEA_CHECK_ARG(mask && mask_bits && N > 0 && H > 0 && T > 0 && P > 0, "sea_mask_bits_to_float: bad argument");
    const int wpr = (H * P + 31) >> 5;
    const int64_t total = (int64_t) N * H * T * P;
    mask_bits_to_float_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t) stream>>>(mask_bits, mask, N, H, T, P, wpr);
    SEA_CHECK_LAUNCH("mask_bits_to_float_kernel");
    return SEA_OK;
}

int sea_csr_count(const uint32_t* mask_bits, void* crow, int idx64, int N, int H, int T_DST, int P, int T_SRC, int k,
                  int is_causal, void* stream) {
    return sea_csr_count_len(mask_bits, crow, idx64, N, H, T_DST, P, T_SRC, k, is_causal, nullptr, stream);
}

int sea_csr_count_len(const uint32_t* mask_bits, void* crow, int idx64, int N, int H, int T_DST, int P, int T_SRC, int k,
                      int is_causal, const int32_t* lengths, void* stream) {
    SEA_CHECK_ARG(mask_bits && crow, "sea_csr_count: null pointer");
    SEA_CHECK_ARG(lengths == nullptr || !is_causal, "sea_csr_count: per-item lengths belong to the non-causal interpolation");
    SEA_CHECK_ARG(N > 0 && H > 0 && T_DST > 0 && P > 0 && T_SRC >= T_DST && k > 0, "sea_csr_count: bad shape");
    SEA_CHECK_ARG((int64_t) H * T_SRC < (1ll << 24), "sea_csr_count: H*T_SRC must stay below 2^24 (fp32-exact column ids, as in the reference)");
    const int wpr = (H * P + 31) >> 5;
    const int64_t rows = (int64_t) N * T_DST;
    cudaStream_t s = (cudaStream_t) stream;
    SEA_DISPATCH_IDX(idx64, I, {
        csr_count_kernel<I><<<(unsigned) rows, kCsrThreads, 0, s>>>(mask_bits, (I*) crow, N, H, T_DST, P, T_SRC, k, is_causal, wpr, lengths);
        SEA_CHECK_LAUNCH("csr_count_kernel");
        crow_scan_kernel<I><<<N, 1024, 0, s>>>((I*) crow, T_DST);
        SEA_CHECK_LAUNCH("crow_scan_kernel");
    });
    return SEA_OK;
}

int sea_crow_scan(void* crow, int idx64, int N, int T_DST, void* stream) {
    SEA_CHECK_ARG(crow && N > 0 && T_DST > 0, "sea_crow_scan: bad argument");
    SEA_DISPATCH_IDX(idx64, I, {
        crow_scan_kernel<I><<<N, 1024, 0, (cudaStream_t) stream>>>((I*) crow, T_DST);
        SEA_CHECK_LAUNCH("crow_scan_kernel");
    });
    return SEA_OK;
}

int sea_csr_fill(const uint32_t* mask_bits, const void* crow, void* col, int idx64, int64_t Z, int32_t* head_ptr, int N, int H, int T_DST,
                 int P, int T_SRC, int k, int is_causal, void* stream) {
    return sea_csr_fill_len(mask_bits, crow, col, idx64, Z, head_ptr, N, H, T_DST, P, T_SRC, k, is_causal, nullptr, stream);
}

int sea_csr_fill_len(const uint32_t* mask_bits, const void* crow, void* col, int idx64, int64_t Z, int32_t* head_ptr, int N, int H, int T_DST,
                     int P, int T_SRC, int k, int is_causal, const int32_t* lengths, void* stream) {
    SEA_CHECK_ARG(mask_bits && crow && (col || Z == 0), "sea_csr_fill: null pointer");
    SEA_CHECK_ARG(lengths == nullptr || !is_causal, "sea_csr_fill: per-item lengths belong to the non-causal interpolation");
    SEA_CHECK_ARG(N > 0 && H > 0 && T_DST > 0 && P > 0 && T_SRC >= T_DST && k > 0 && Z >= 0, "sea_csr_fill: bad shape");
    SEA_CHECK_ARG(head_ptr == nullptr || (P % 32) == 0, "sea_csr_fill: head_ptr needs P %% 32 == 0");
    const int wpr = (H * P + 31) >> 5;
    const int64_t rows = (int64_t) N * T_DST;
    SEA_DISPATCH_IDX(idx64, I, {
        csr_fill_kernel<I><<<(unsigned) rows, kCsrThreads, 0, (cudaStream_t) stream>>>(
            mask_bits, (const I*) crow, (I*) col, Z, head_ptr, N, H, T_DST, P, T_SRC, k, is_causal, wpr, lengths);
        SEA_CHECK_LAUNCH("csr_fill_kernel");
    });
    return SEA_OK;
}

int sea_flat_csr_to_dense(const void* crow, const void* col, int idx64, const float* values, int64_t Z, float* out,
                          int N, int H, int T_DST, int T_SRC, void* stream) {
    SEA_CHECK_ARG(crow && (col || Z == 0) && out, "sea_flat_csr_to_dense: null pointer");
    cudaStream_t s = (cudaStream_t) stream;
    SEA_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t) N * H * T_DST * T_SRC * sizeof(float), s), "memset");
    const int64_t rows = (int64_t) N * T_DST;
    SEA_DISPATCH_IDX(idx64, I, {
        csr_to_dense_kernel<I><<<cdiv(rows, kCsrWarps), kCsrWarps * 32, 0, s>>>((const I*) crow, (const I*) col, values, Z, out, N, H, T_DST, T_SRC);
        SEA_CHECK_LAUNCH("csr_to_dense_kernel");
    });
    return SEA_OK;
}

int sea_flat_csr_masked_bmm(const void* crow, const void* col, int idx64, int64_t Z,
                            const void* a, int64_t a_sn, int64_t a_sh, int64_t a_st,
                            const void* b, int64_t b_sn, int64_t b_sh, int64_t b_st, int dtype, float* out_values,
                            int N, int H, int T_DST, int T_SRC, int D, void* stream) {
    SEA_CHECK_ARG(crow && (col || Z == 0) && a && b && (out_values || Z == 0), "sea_flat_csr_masked_bmm: null pointer");
    const int64_t rows = (int64_t) N * T_DST;
    SEA_DISPATCH_DTYPE(dtype, T, SEA_DISPATCH_IDX(idx64, I, {
        csr_masked_bmm_kernel<T, I><<<cdiv(rows, kCsrWarps), kCsrWarps * 32, 0, (cudaStream_t) stream>>>(
            (const I*) crow, (const I*) col, Z, (const T*) a, a_sn, a_sh, a_st, (const T*) b, b_sn, b_sh, b_st,
            out_values, N, H, T_DST, T_SRC, D);
        SEA_CHECK_LAUNCH("csr_masked_bmm_kernel");
    }));
    return SEA_OK;
}

int sea_flat_csr_softmax(const void* crow, const void* col, int idx64, int64_t Z, const float* in_values,
                         float* out_values, int N, int H, int T_DST, int T_SRC, void* stream) {
    SEA_CHECK_ARG(crow && (col || Z == 0) && (in_values || Z == 0) && (out_values || Z == 0), "sea_flat_csr_softmax: null pointer");
    SEA_CHECK_ARG(H * 8 <= 48 * 1024, "sea_flat_csr_softmax: too many heads");
    SEA_DISPATCH_IDX(idx64, I, {
        csr_softmax_kernel<I><<<(unsigned) ((int64_t) N * T_DST), 128, (size_t) H * 8, (cudaStream_t) stream>>>(
            (const I*) crow, (const I*) col, Z, in_values, out_values, N, H, T_DST, T_SRC);
        SEA_CHECK_LAUNCH("csr_softmax_kernel");
    });
    return SEA_OK;
}

int sea_flat_csr_elmul(const void* crow, const void* col, int idx64, int64_t Z, const float* in_values, float* out_values,
                       const float* dense, int64_t d_sn, int64_t d_sh, int64_t d_st, int64_t d_sj,
                       int N, int H, int T_DST, int T_SRC, void* stream) {
    SEA_CHECK_ARG(crow && (col || Z == 0) && dense, "sea_flat_csr_elmul: null pointer");
    const int64_t rows = (int64_t) N * T_DST;
    SEA_DISPATCH_IDX(idx64, I, {
        csr_elmul_kernel<I><<<cdiv(rows, kCsrWarps), kCsrWarps * 32, 0, (cudaStream_t) stream>>>(
            (const I*) crow, (const I*) col, Z, in_values, out_values, dense, d_sn, d_sh, d_st, d_sj, N, H, T_DST, T_SRC);
        SEA_CHECK_LAUNCH("csr_elmul_kernel");
    });
    return SEA_OK;
}

int sea_flat_csr_sdbmm(const void* crow, const void* col, int idx64, int64_t Z, const float* values,
                       const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype, float* out,
                       int N, int H, int T_DST, int T_SRC, int D, void* stream) {
    SEA_CHECK_ARG(crow && (col || Z == 0) && v && out, "sea_flat_csr_sdbmm: null pointer");
    SEA_CHECK_ARG(D > 0 && D <= 256, "sea_flat_csr_sdbmm: head dim %d unsupported (1..256)", D);
    cudaStream_t s = (cudaStream_t) stream;
    SEA_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t) N * H * T_DST * D * sizeof(float), s), "memset");
    const int64_t rows = (int64_t) N * T_DST;
    SEA_DISPATCH_DTYPE(dtype, T, SEA_DISPATCH_IDX(idx64, I, {
        csr_sdbmm_kernel<T, I><<<cdiv(rows, kCsrWarps), kCsrWarps * 32, 0, s>>>(
            (const I*) crow, (const I*) col, Z, values, (const T*) v, v_sn, v_sh, v_st, out, N, H, T_DST, T_SRC, D);
        SEA_CHECK_LAUNCH("csr_sdbmm_kernel");
    }));
    return SEA_OK;
}

int sea_resize_m_to_t_dense(const float* x, float fill, const float* attention_mask, int64_t m_sn, int64_t m_st,
                            float* out, int N, int H, int T1, int P, int T2, void* stream) {
    SEA_CHECK_ARG(x && attention_mask && out, "sea_resize_m_to_t_dense: null pointer");
    SEA_CHECK_ARG((size_t) T2 * 4 <= 200 * 1024, "sea_resize_m_to_t_dense: T2 %d too large for one CTA", T2);
    const size_t smem = (size_t) T2 * sizeof(int);
    SEA_CUDA_TRY(cudaFuncSetAttribute(resize_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr");
    resize_dense_kernel<<<(unsigned) ((int64_t) N * T1), 256, smem, (cudaStream_t) stream>>>(x, fill, attention_mask, m_sn, m_st, out, N, H, T1, P, T2);
    SEA_CHECK_LAUNCH("resize_dense_kernel");
    return SEA_OK;
}

}  // extern "C"
