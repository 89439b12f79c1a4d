// Internal (non-ABI) launchers of the short-context attention path.
#pragma once
#include "common.cuh"

namespace sea {

// block_attn_umma.cu: tcgen05 / TMEM version (bf16, d = 64) of the masked block attention, driven directly by the top-k pixel
// bits [N][T_DST][H * P / 32] (the a8 interpolation happens inside the kernel; p_lg = exact_edge_shift(P, T_SRC)).
int launch_block_attention_umma(const uint32_t* mask_bits, int P, int p_lg,
                                const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, void* out,
                                int N, int H, int T_DST, int T_SRC, int is_causal, cudaStream_t s);

// development aid (SEA_ATTN_TRACE=1): per-warp cycle counters of the last traced launch -> host; returns the number of CTAs
int64_t attn_trace_read(uint32_t* host, int64_t max_words);

}  // namespace sea
