// a5 tail + a6 + a7 fused (bf16 path): one CTA per query row (n, t) holds that row's H x P estimated scores on
// chip from the 1x1-conv output to the top-k bit mask:
//   y3 [N,T,W,H] fp32 (1x1 conv evaluated before the x4 nearest upsample, csrc/umma_conv.cu)
//   -> upsample x(P/W), two bias-valued pad columns, area resize P+2 -> P   (attention.py:275-277, modules.py:12-31,42-55)
//   -> LayerNorm(P) (attention.py:280) -> softmax(P) (attention.py:670-673)  -> probs fp32 [N,H,T,P] (optional store)
//   -> grouped top-k over the H*P keys of the row (attention.py:843-917)      -> bit mask [N,T,H*P/32]
// The probabilities are read back from HBM by nobody: the top-k consumes them from shared memory.
#include "common.cuh"
#include "topk.cuh"
#include "csr_common.cuh"

namespace sea {

template <int kPerLane, int kUp>
__global__ void __launch_bounds__(kTopkThreads)
tail_topk_kernel(const float* __restrict__ y3, const float* __restrict__ bias, const float* __restrict__ ln_w,
                 const float* __restrict__ ln_b, const float* __restrict__ k_per_row, float* __restrict__ probs,
                 uint32_t* __restrict__ mask_bits, int32_t* __restrict__ crow_counts, int k_clamp,
                 int N, int H, int Tn, int W, int P_rt) {
    // P (= 32 * kPerLane) and the upsample factor are compile-time constants: every / and % below folds into
    // shifts / multiplies (a CTA lives for one row only, so per-thread set-up divisions are not amortised)
    constexpr int P = 32 * kPerLane;
    extern __shared__ __align__(16) uint32_t smem_u[];
    __shared__ int hist[256];
    __shared__ int scratch[16];
    const int G = H * P;
    uint32_t* skeys = smem_u;                                   // [G]
    uint32_t* sbits = skeys + G;                                // [G/32]
    float* ys = reinterpret_cast<float*>(sbits + (G >> 5));     // [H][W+2]: W conv outputs, then the bias (pad columns), then 0
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = blockIdx.x / Tn, t = blockIdx.x % Tn;
    const int ldy = W + 2;
    const float* yr = y3 + ((int64_t) n * Tn + t) * W * H;
    {
        int h = tid % H, w = tid / H;
        const int dh = kTopkThreads % H, dw = kTopkThreads / H;
        for (int idx = tid; idx < W * H; idx += kTopkThreads) {
            ys[h * ldy + w] = yr[idx];
            h += dh; w += dw;
            if (h >= H) { h -= H; ++w; }
        }
    }
    for (int h = tid; h < H; h += kTopkThreads) { ys[h * ldy + W] = bias[h]; ys[h * ldy + W + 1] = 0.f; }
    __syncthreads();
    constexpr int PW = P + 2;
    const int up = kUp > 0 ? kUp : P / W;
    // area-resize window of each of this lane's columns, resolved ONCE into (up to) three slots of the per-head
    // vector: a conv output w, the bias slot W (the zero-padded columns of the 1x1 conv) or the zero slot W+1
    int tap[kPerLane][3];
    float rc[kPerLane], lw[kPerLane], lb[kPerLane];
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int j = lane + 32 * i;
        const int st = (j * PW) / P;
        const int cnt = ((j + 1) * PW + P - 1) / P - st;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int pcol = st + c;
            tap[i][c] = c >= cnt ? W + 1 : ((pcol == 0 || pcol == PW - 1) ? W : (pcol - 1) / up);
        }
        rc[i] = 1.0f / (float) cnt;
        lw[i] = ln_w[j];
        lb[i] = ln_b[j];
    }
    const float invP = 1.0f / (float) P;
    for (int h = wid; h < H; h += kTopkThreads / 32) {
        const float* yh = ys + h * ldy;
        float val[kPerLane];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            val[i] = (yh[tap[i][0]] + yh[tap[i][1]] + yh[tap[i][2]]) * rc[i];
            s += val[i];
        }
        const float mean = warp_sum(s) * invP;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { const float d = val[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * invP + 1e-5f);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { val[i] = (val[i] - mean) * rstd * lw[i] + lb[i]; mx = fmaxf(mx, val[i]); }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) { val[i] = __expf(val[i] - mx); sum += val[i]; }
        const float inv = 1.0f / warp_sum(sum);
        float* prow = probs ? probs + (((int64_t) n * H + h) * Tn + t) * P : nullptr;
        uint32_t* krow = skeys + h * P;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            const int j = lane + 32 * i;
            const float pr = val[i] * inv;
            if (prow) prow[j] = pr;
            krow[j] = orderable(pr);
        }
    }
    if (mask_bits == nullptr) return;
    __syncthreads();
    const float kf = k_per_row[blockIdx.x];
    const int K = (int) fminf(ceilf(kf), (float) G);
    topk_select_to_bits(skeys, G, K, sbits, hist, scratch);
    __syncthreads();
    uint32_t* out_row = mask_bits + (int64_t) blockIdx.x * (G >> 5);
    int cnt = 0;
    const float sc = __fdiv_rn((float) (t + 1), (float) P);       // causal prefill: source length of row t is t+1
    for (int w = tid; w < (G >> 5); w += kTopkThreads) {
        const uint32_t word = sbits[w];
        out_row[w] = word;
        if (crow_counts != nullptr) cnt += word_width_sum(word, w, P, sc, k_clamp);
    }
    if (crow_counts != nullptr) {
        // a8 pass 1 fused: crow[n, t+1] = entries of this row (sea_crow_scan turns the counts into offsets)
        cnt = warp_sum_i(cnt);
        if (lane == 0) scratch[wid] = cnt;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
#pragma unroll
            for (int i = 0; i < kTopkThreads / 32; ++i) tot += scratch[i];
            crow_counts[(int64_t) n * (Tn + 1) + t + 1] = tot;
            if (t == 0) crow_counts[(int64_t) n * (Tn + 1)] = 0;
        }
    }
}

}  // namespace sea

using namespace sea;

extern "C" {

int sea_predictor_tail_topk_fwd(const float* y3, const float* bias, const float* ln_w, const float* ln_b,
                                const float* k_per_row, float* probs, uint32_t* mask_bits, int32_t* crow_counts, int k_clamp,
                                int N, int H, int T, int W, int P, void* stream) {
    SEA_CHECK_ARG(y3 && bias && ln_w && ln_b && (probs || mask_bits), "sea_predictor_tail_topk_fwd: null pointer");
    SEA_CHECK_ARG(mask_bits == nullptr || k_per_row != nullptr, "sea_predictor_tail_topk_fwd: k_per_row is required for the top-k");
    SEA_CHECK_ARG(N > 0 && H > 0 && T > 0 && W > 0 && P > 0, "sea_predictor_tail_topk_fwd: bad shape");
    SEA_CHECK_ARG(crow_counts == nullptr || (mask_bits != nullptr && k_clamp > 0), "sea_predictor_tail_topk_fwd: crow_counts needs mask_bits and k");
    if (P % 32 != 0 || P % W != 0 || P > 1024) {
        set_error("sea_predictor_tail_topk_fwd: P=%d must be a multiple of 32 and of W=%d, <= 1024", P, W);
        return SEA_ERR_UNSUPPORTED;
    }
    const int G = H * P;
    const size_t smem = ((size_t) G + (G >> 5)) * 4 + (size_t) H * (W + 2) * 4 + 16;
    SEA_CHECK_ARG(smem <= 220 * 1024, "sea_predictor_tail_topk_fwd: H*P=%d keys do not fit shared memory", G);
    cudaStream_t s = (cudaStream_t) stream;
    const unsigned grid = (unsigned) ((int64_t) N * T);
#define SEA_TAIL_LAUNCH(PL, UP)                                                                                       \
    {                                                                                                                \
        auto kern = tail_topk_kernel<PL, UP>;                                                                        \
        SEA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem), "smem attr"); \
        kern<<<grid, kTopkThreads, smem, s>>>(y3, bias, ln_w, ln_b, k_per_row, probs, mask_bits, crow_counts, k_clamp, N, H, T, W, P);      \
    }
#define SEA_TAIL_CASE(PL)                                                                                            \
    case PL: {                                                                                                       \
        if (P / W == 4) SEA_TAIL_LAUNCH(PL, 4) else SEA_TAIL_LAUNCH(PL, 0)                                           \
        break;                                                                                                       \
    }
    switch (P / 32) {
        SEA_TAIL_CASE(1) SEA_TAIL_CASE(2) SEA_TAIL_CASE(4) SEA_TAIL_CASE(8) SEA_TAIL_CASE(16) SEA_TAIL_CASE(32)
        default:
            set_error("sea_predictor_tail_topk_fwd: P=%d unsupported (32,64,128,256,512,1024)", P);
            return SEA_ERR_UNSUPPORTED;
    }
#undef SEA_TAIL_CASE
#undef SEA_TAIL_LAUNCH
    SEA_CHECK_LAUNCH("tail_topk_kernel");
    return SEA_OK;
}

}  // extern "C"
