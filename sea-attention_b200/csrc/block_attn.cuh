// Internal (non-ABI) launchers of the short-context attention path.
#pragma once
#include "common.cuh"

namespace sea {

// block_attn_umma.cu: tcgen05 / TMEM version (bf16, d = 64) of the masked block attention over the dense bit-packed mask.
int launch_block_attention_umma(const unsigned long long* dmask, int W64, const uint32_t* tile_act, int act_words,
                                const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, void* out,
                                int N, int H, int T_DST, int T_SRC, int is_causal, cudaStream_t s);

}  // namespace sea
