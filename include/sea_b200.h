/*
 * sea_b200.h -- C ABI of libsea_b200.so: the B200 (sm_100a) implementation of the per-layer
 * SEA / Perlin attention forward of gmlwns2000/sea-attention.
 *
 * Every entry point replaces one python-level operator of the reference (cited as file:line,
 * relative to the reference root).  Conventions, shared by all entries:
 *   - all tensor pointers are DEVICE pointers owned by the caller; the library never allocates,
 *     frees, or synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - tensors are dense row-major with the innermost dimension contiguous; where the reference
 *     accepts strided views the entry takes explicit element strides;
 *   - `dtype` is one of SEA_DTYPE_*: it is the storage type of activations (q, k, v, outputs);
 *     all arithmetic accumulates in fp32;
 *   - return value: 0 on success, a negative SEA_ERR_* code otherwise; sea_last_error() returns a
 *     thread-local human readable message for the last failure;
 *   - flat-CSR container (the reference's batched torch.sparse_csr_tensor of shape
 *     [N, T_DST, H*T_SRC], causal_resize_m_to_t.py:757-762): crow [N, T_DST+1], col [N, Z];
 *     a column id encodes (head, source token) as h*T_SRC + j; inside a row entries are head-major
 *     and, inside one compressed pixel, descending in j.  `idx64` selects int64 (torch boundary)
 *     or int32 (internal) indices for BOTH crow and col.
 */
#ifndef SEA_B200_H_
#define SEA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEA_DTYPE_F32 0
#define SEA_DTYPE_BF16 1
#define SEA_DTYPE_F16 2

#define SEA_OK 0
#define SEA_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define SEA_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define SEA_ERR_UNSUPPORTED (-3)

#define SEA_ABI_VERSION 1

#if defined(__GNUC__)
#define SEA_API __attribute__((visibility("default")))
#else
#define SEA_API
#endif

SEA_API int sea_abi_version(void);
SEA_API const char* sea_last_error(void);
/* Compute capability major*10+minor of the current device, or a negative error. */
SEA_API int sea_device_arch(void);

/* ---------------------------------------------------------------------------------------------
 * a7  grouped top-k  (attention.py:774-947; helper variant ops/kernels/causal_topk_masking.py:31)
 * keys: fp32, logical shape [N, H, T, P] with element strides (sn, sh, st), P contiguous.
 * group_mode 0 ('causal_batch', attention.py:843-849): one group per (n, t) = all heads, flat index
 *              h*P+m;  k_per_group [N*T] fp32 = the reference's per_item_top_k (already rounded and
 *              clamped >= 1, attention.py:856,866).
 * group_mode 1 ('query', :850-853): one group per (n, h, t), k_per_group [N] (per item).
 * alive  <=>  rank by descending key < k_per_group[group]; equal keys: LOWER flat index wins.
 * row_valid (nullable, u8 [N, T]): 0 marks a padded query row -> all dead (attention.py:928-931).
 * mask_bits out: u32 words, [N, T, ceil(H*P/32)], bit (h*P+m) of row (n, t).
 */
SEA_API int sea_topk_mask_bits(const float* keys, int64_t sn, int64_t sh, int64_t st,
                       const float* k_per_group, const uint8_t* row_valid,
                       uint32_t* mask_bits, int N, int H, int T, int P, int group_mode, void* stream);

/* 0/1 float mask [N,H,T,P] (the reference's partial_attention_mask of the benchmarking branch,
 * attention.py:916-917) <-> bit mask. */
SEA_API int sea_mask_float_to_bits(const float* mask, int64_t sn, int64_t sh, int64_t st, uint32_t* mask_bits,
                           int N, int H, int T, int P, void* stream);
SEA_API int sea_mask_bits_to_float(const uint32_t* mask_bits, float* mask, int N, int H, int T, int P, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8  resize_from_m_to_t_csr  (causal_resize_m_to_t.py:910-1007 -> scan_col METHOD 1 :648-762 ->
 *     __scan_col_4_compute :493-572, triton_round :214-264).
 * Pass 1 (count): crow[n, 0] = 0, crow[n, t+1] = sum over rows <= t of min(int(ve-vs)*alive, k).
 * Pass 2 (fill):  col[n, crow[n,t] ...] per SURVEY appendix C.3; tail col[n, crow[n,T_DST]:Z) = 0.
 * T_DST query rows are the LAST T_DST of T_SRC tokens (target_width[-T_DST:], :954-957).
 */
SEA_API int sea_csr_count(const uint32_t* mask_bits, void* crow, int idx64,
                  int N, int H, int T_DST, int P, int T_SRC, int k, int is_causal, void* stream);
/* head_ptr (nullable, int32 [N, T_DST, H+1], needs P % 32 == 0): absolute entry offset (inside batch item n) at which
 * head h starts in row t, head_ptr[..., H] = end of the row; an auxiliary index for sea_sparse_attention_fwd. */
SEA_API int sea_csr_fill(const uint32_t* mask_bits, const void* crow, void* col, int idx64, int64_t Z, int32_t* head_ptr,
                 int N, int H, int T_DST, int P, int T_SRC, int k, int is_causal, void* stream);
/* The same two passes for a RIGHT-PADDED non-causal batch: lengths (int32 [N], nullable) = valid tokens of item n, used as the
 * interpolation width of every row of that item (what the reference's dense path does through the mask cumsum, resize_m_to_t.py:36-47;
 * its sparse path uses T_SRC regardless, causal_resize_m_to_t.py:955-957), so no column of a padded token is produced. */
SEA_API int sea_csr_count_len(const uint32_t* mask_bits, void* crow, int idx64,
                  int N, int H, int T_DST, int P, int T_SRC, int k, int is_causal, const int32_t* lengths, void* stream);
SEA_API int sea_csr_fill_len(const uint32_t* mask_bits, const void* crow, void* col, int idx64, int64_t Z, int32_t* head_ptr,
                 int N, int H, int T_DST, int P, int T_SRC, int k, int is_causal, const int32_t* lengths, void* stream);

/* flat_csr_to_dense (ops/kernels/flat_csr_to_dense.py:3-36): out [N,H,T_DST,T_SRC] fp32, zero filled
 * then out[n,h,t,j] = values[n,z] (values == NULL -> 1.0). */
SEA_API int sea_flat_csr_to_dense(const void* crow, const void* col, int idx64, const float* values, int64_t Z,
                          float* out, int N, int H, int T_DST, int T_SRC, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9  flat_csr_masked_bmm (ops/kernels/flat_csr_masked_bmm.py:137-195): SDDMM, unscaled,
 *     out_values[n,z] = <a[n,h,row,:], b[n,h,j,:]> accumulated in fp32.
 * a, b: `dtype`, element strides (n, h, t), d contiguous.
 */
SEA_API int sea_flat_csr_masked_bmm(const void* crow, const void* col, int idx64, int64_t Z,
                            const void* a, int64_t a_sn, int64_t a_sh, int64_t a_st,
                            const void* b, int64_t b_sn, int64_t b_sh, int64_t b_st,
                            int dtype, float* out_values,
                            int N, int H, int T_DST, int T_SRC, int D, void* stream);

/* a10 flat_csr_softmax (ops/kernels/flat_csr_softmax.py:127-176): softmax inside each (row, head). */
SEA_API int sea_flat_csr_softmax(const void* crow, const void* col, int idx64, int64_t Z,
                         const float* in_values, float* out_values,
                         int N, int H, int T_DST, int T_SRC, void* stream);

/* a11 flat_csr_elmul (ops/kernels/flat_csr_elmul.py:110-162): out[z] = in[z] * dense[n,h,row,j];
 * dense fp32 with element strides (n,h,tdst,tsrc); stride 0 allowed (attention.py:1170). */
SEA_API int sea_flat_csr_elmul(const void* crow, const void* col, int idx64, int64_t Z,
                       const float* in_values, float* out_values,
                       const float* dense, int64_t d_sn, int64_t d_sh, int64_t d_st, int64_t d_sj,
                       int N, int H, int T_DST, int T_SRC, void* stream);

/* a12 flat_csr_sdbmm (ops/kernels/flat_csr_sdbmm.py:323-439): out[n,h,row,:] = sum_z p[z] V[n,h,j,:],
 * out fp32 [N,H,T_DST,D] contiguous (zeroed by the call).  Never truncates (the reference silently
 * drops entries beyond MAX_ROW_T, :382-388). */
SEA_API int sea_flat_csr_sdbmm(const void* crow, const void* col, int idx64, int64_t Z, const float* values,
                       const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype,
                       float* out, int N, int H, int T_DST, int T_SRC, int D, void* stream);

/* a16 resize_from_m_to_t (ops/kernels/resize_m_to_t.py:6-73), training=False, oversampled=None|1.0:
 * out[n,h,t,j] = x[n,h,t, floor((cs-0.5)/L*P - 1e-4)] for valid j (cs = 1-based rank of j among the
 * valid source tokens of row t, L = their count), `fill` elsewhere.  attention_mask additive fp32,
 * causal: [N,1,T1,T2]; non causal: [N,1,1,T2] (row stride 0). */
SEA_API int sea_resize_m_to_t_dense(const float* x, float fill, const float* attention_mask, int64_t m_sn, int64_t m_st,
                            float* out, int N, int H, int T1, int P, int T2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a2+a3  causal Performer estimate  (performer_pytorch.FastAttention, causal + generalized ReLU
 *        features; call sites attention.py:159-164, 504-508, 527-534, 559-572).
 *   v_for_atten = cat(pos_emb[t], v[n,h,t])            (width 2*D, attention.py:504-508)
 *   phi(x) = relu(D^-1/4 x P^T) + 1e-3 ; out_t = phi(q_t) S_t / (phi(q_t) . (z_t + 1e-6))
 * Also emits the causal running mean of v (attention.py:1237-1241) from the same pass.
 * q,k,v: `dtype`, strides (n,h,t); pos_emb fp32 [>=T, D]; proj fp32 [F, D];
 * ctx out: `dtype` [N,H,T,2D]; cumavg out (nullable): `dtype` [N,H,T,D];
 * workspace: fp32, at least sea_performer_workspace_floats(N,H,T,D,F) elements.
 */
SEA_API int64_t sea_performer_workspace_floats(int N, int H, int T, int D, int F);
SEA_API int sea_performer_causal_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                             const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                             const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                             const float* pos_emb, const float* proj, int dtype,
                             void* ctx, void* cumavg, float* workspace,
                             int N, int H, int T, int D, int F, void* stream);

/* a2+a3 on the tensor cores (bf16, D = 64, F <= 63; csrc/performer_mma.cu): same contract as
 * sea_performer_causal_fwd, chunk-parallel GEMMs chained through registers.  workspace: fp32,
 * >= sea_performer_mma_workspace_floats(...) elements. */
SEA_API int sea_performer_mma_supported(int dtype, int D, int F);
SEA_API int64_t sea_performer_mma_workspace_floats(int N, int H, int T, int D, int F);
SEA_API int sea_performer_causal_mma_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                         const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                         const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                         const float* pos_emb, const float* proj, void* ctx, void* cumavg, float* workspace,
                                         int N, int H, int T, int D, int F, void* stream);
/* The same stage for a RANGE of a sequence that is sharded over ranks by query block: q / k / v / pos_emb point at the first
 * row of the range (T rows), t_off = its absolute position.  phase 1: per-chunk sums of the range into `workspace` and their total
 * (`total`, sea_performer_mma_state_floats floats: S, z and the running sum of v, per (n, h, slab)) -- what the ranks exchange;
 * phase 2: exclusive prefix starting from `init` (the summed totals of everything before the range; NULL = zero) + the outputs, reusing
 * the workspace of phase 1; phase 0: both, from zero.  q and ctx may be NULL in phase 1. */
SEA_API int64_t sea_performer_mma_state_floats(int N, int H, int D, int F);
SEA_API int sea_performer_causal_mma_range(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                           const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                           const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                           const float* pos_emb, const float* proj, void* ctx, void* cumavg, float* workspace,
                                           const float* init, float* total, int N, int H, int T, int D, int F, int t_off, int phase, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4  predictor MLP (attention.py:190-196,242-245,289-291,577-625) for the causal predictor:
 *   x = cat(ctx[2D], v[D]) -> Linear(3D,2D) -> LayerNorm -> GELU = t_pred
 *   dec = Linear(2D, S*W)(t_pred), S = 2 splits, W = P/4; ChannelSplit -> channel c = h*S+s
 *   first CNN LayerNorm(W) (attention.py:267) is applied here, per (c, t), over W
 *   scales = Linear(2D, 2)(t_pred)
 * outputs: cnn_in channels-last [N, T, W, C=H*S] (`dtype`), scales fp32 [N,H,T,2],
 *          t_pred (nullable, `dtype`, [N,H,T,2D]).
 * All weights fp32, torch layout ([out, in]).
 */
SEA_API int sea_predictor_mlp_fwd(const void* ctx, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype,
                          const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                          const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                          const float* scl_w, const float* scl_b,
                          void* cnn_in, float* scales, void* t_pred,
                          int N, int H, int T, int D, int S, int W, void* stream);

/* a4 on the tensor cores (bf16 only; csrc/umma_mlp.cu): same computation as sea_predictor_mlp_fwd with both Linear
 * layers as chained tcgen05 GEMMs per 128-token tile (intermediates stay in TMEM / shared memory).  Shapes: D = 64,
 * S = 2, H <= 128, W in {16,32,64}; ctx contiguous [N,H,T,2D] bf16; cnn_in bf16 [N,T,W,2H]; no t_pred output.
 * workspace: >= sea_predictor_mlp_umma_workspace_bytes() bytes, 128-byte aligned.  The call re-packs enc_w / dec_w / scl_w
 * into it; passing all three as NULL reuses the packing a previous call left in `workspace` (inference: weights are constant). */
SEA_API int sea_predictor_mlp_umma_supported(int dtype, int H, int D, int S, int W);
SEA_API int64_t sea_predictor_mlp_umma_workspace_bytes(void);
SEA_API int sea_predictor_mlp_umma_fwd(const void* ctx, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                       const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                                       const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                                       const float* scl_w, const float* scl_b, void* cnn_in, float* scales, void* workspace,
                                       int N, int H, int T, int D, int S, int W, void* stream);
/* Same with a padded output: cnn_in is [N,T,W,Cout], Cout >= 2H (channels 2H.. are written as zeros), so that models whose 2H != 64
 * (e.g. OPT-125m, H = 12) can feed the 64-channel tcgen05 convolutions.  H need not divide 128. */
SEA_API int sea_predictor_mlp_umma_fwd_ex(const void* ctx, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                       const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                                       const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                                       const float* scl_w, const float* scl_b, void* cnn_in, float* scales, void* workspace,
                                       int N, int H, int T, int D, int S, int W, int Cout, void* stream);

/* a4 on the tensor cores for any head dim (bf16; csrc/mlp_mma.cu): the same computation as warp-level mma.sync GEMMs chained
 * through registers, the weights streamed through shared memory in K chunks -- for the head dims whose weights do not fit the
 * tcgen05 kernel's shared memory (OPT-2.7B D = 80, the long-context sweep D = 128; reference: attention.py:190-196, 242-245, 267,
 * 289-291, 599-625 take any attention_head_size).  Shapes: D in {32,64,80,96,128}, S = 2, H <= 128, W in {32,64}; ctx contiguous
 * [N,H,T,2D] bf16; cnn_in bf16 [N,T,W,Cout], Cout >= 2H (channels 2H.. are written as zeros).
 * workspace: >= sea_predictor_mlp_mma_workspace_bytes(D,S,W) bytes, 16-byte aligned; weights NULL = reuse the packing in it. */
SEA_API int sea_predictor_mlp_mma_supported(int dtype, int H, int D, int S, int W);
SEA_API int64_t sea_predictor_mlp_mma_workspace_bytes(int D, int S, int W);
SEA_API int sea_predictor_mlp_mma_fwd(const void* ctx, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                      const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                                      const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                                      const float* scl_w, const float* scl_b, void* cnn_in, float* scales, void* workspace,
                                      int N, int H, int T, int D, int S, int W, int Cout, void* stream);

/* a5  one CausalConv2d(C,C,3,padding=2,dilation=2,causal) + ReLU (modules.py:96-192;
 *     attention.py:271-274) on channels-last activations [N,T,W,C]:
 *   y[t,w,o] = relu(b[o] + sum_{i,j<3} sum_c Wt[o,c,i,j] x[t-4+2i, w-2+2j, c])   (zero outside)
 * weight fp32 in the reference layout [O, C, 5, 3] (rows 3..4 are the masked-out taps). */
SEA_API int sea_causal_conv3x3_dil2_relu(const void* x, const float* weight, const float* bias, void* y, int dtype,
                                 int N, int T, int W, int C, int O, void* stream);

/* a5 on the tensor cores (bf16 only): the same CausalConv2d + ReLU as an implicit GEMM with tcgen05.mma / TMEM /
 * TMA (csrc/umma_conv.cu).  Shapes: C = O = 64, W a divisor of 128; sea_conv_umma_supported() tells.
 * workspace: >= sea_conv_umma_workspace_bytes(C, O) bytes, 128-byte aligned (bf16 re-packed weights); weight == NULL
 * reuses the packing a previous call with the same weights left in `workspace`.
 * sea_conv1x1_umma: the 1x1 CausalConv2d(C=64 -> O=32) of attention.py:276 evaluated BEFORE the nearest x4
 * upsample (they commute): y [N,T,W,O] fp32 = x . Wt^T + b. */
SEA_API int sea_conv_umma_supported(int dtype, int W, int C, int O);
SEA_API int64_t sea_conv_umma_workspace_bytes(int C, int O);
SEA_API int sea_causal_conv3x3_dil2_relu_umma(const void* x, const float* weight, const float* bias, void* y, void* workspace,
                                              int N, int T, int W, int C, int O, void* stream);
SEA_API int sea_conv1x1_umma(const void* x, const float* weight, const float* bias, float* y, void* workspace,
                             int N, int T, int W, int C, int O, void* stream);
/* The second 3x3 convolution and the 1x1 convolution behind it (attention.py:274-276: CausalConv2d + ReLU, then -- commuted
 * with the upsample -- CausalConv2d(C, H, 1)) in ONE kernel: y3 [N,T,W,O3] fp32 = conv1x1(relu(conv3x3(x))), the 64-channel
 * activation between them stays in shared memory / TMEM.  Shapes: W = C = O = 64, O3 = 32 (..._supported() tells).
 * workspace / workspace3: bf16 weight packings as for the two separate entries (sea_conv_umma_workspace_bytes(C, O) and
 * (C, O3) bytes); weight == NULL / weight3 == NULL reuse the packing left there by an earlier call. */
SEA_API int sea_conv3x3_conv1x1_umma_supported(int dtype, int W, int C, int O, int O3);
SEA_API int sea_causal_conv3x3_dil2_relu_conv1x1_umma(const void* x, const float* weight, const float* bias, void* workspace,
                                                      const float* weight3, const float* bias3, void* workspace3, float* y3,
                                                      int N, int T, int W, int C, int O, int O3, void* stream);

/* a5 tail + a6  (attention.py:275-280, 670-673): nearest x4 along W, CausalConv2d(C,H,1,padding=1)
 * (width P+2, the two pad columns equal the bias), area-resize to P, LayerNorm(P), softmax(P).
 * x channels-last [N,T,W,C]; weight fp32 [H, C]; probs out fp32 [N,H,T,P]; scores out nullable. */
SEA_API int sea_predictor_tail_fwd(const void* x, int dtype, const float* weight, const float* bias,
                           const float* ln_w, const float* ln_b, float* probs, float* scores,
                           int N, int H, int T, int W, int C, int P, void* stream);

/* a5 tail + a6 + a7 fused (bf16 path): y3 fp32 [N,T,W,H] = output of sea_conv1x1_umma -> upsample, pad, area resize,
 * LayerNorm(P), softmax(P) (attention.py:275-280, 670-673) -> probs fp32 [N,H,T,P] (nullable) and the 'causal_batch'
 * top-k bit mask [N,T,H*P/32] (nullable; k_per_row [N*T] as in sea_topk_mask_bits).  P in {32,...,1024}, W | P. */
SEA_API int sea_predictor_tail_topk_fwd(const float* y3, const float* bias, const float* ln_w, const float* ln_b,
                                        const float* k_per_row, float* probs, uint32_t* mask_bits,
                                        int32_t* crow_counts, int k_clamp,
                                        int N, int H, int T, int W, int P, void* stream);
/* crow_counts (nullable, int32 [N, T+1]): when given, the kernel also performs pass 1 of a8 for the causal prefill
 * (T_DST = T_SRC = T): crow_counts[n, t+1] = entries of row t, crow_counts[n, 0] = 0; sea_crow_scan() then turns the
 * counts into row offsets in place (the second half of sea_csr_count). */
SEA_API int sea_crow_scan(void* crow, int idx64, int N, int T_DST, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9-a14 fused sparse attention of the benchmarking branch (attention.py:1151-1173, 1237-1244,
 * 1279-1282): per (row, head) scores -> softmax -> * sigmoid(scales[...,0]) (if use_scaler) ->
 * sum p V -> out = ctx*sigmoid(scales[...,1]) + (1-sigmoid(.))*cumavg -> [N, T_DST, H*D].
 * probs_values (nullable, fp32 [N,Z]) receives the scaled probabilities (the reference's
 * partial_attention_probs.values()).  Requires head-major entries inside a row (what a8 emits).
 * cumavg nullable (then out = ctx, layout still [N,T_DST,H*D]); its element strides are (avg_sh per (n,h), avg_st per row):
 * causal running mean [N,H,T_DST,D] -> (T_DST*D, D); the BERT probability-weighted mean [N,H,1,D] -> (D, 0).
 * head_ptr (nullable): the index sea_csr_fill emits; with it and 16-bit activations (D in {32,64,128}) the call runs the
 * warp-per-(row, head) kernel with batched 128-bit gathers, otherwise a CTA-per-row kernel that finds the head
 * segments by binary search.
 */
SEA_API int sea_sparse_attention_fwd(const void* crow, const void* col, int idx64, int64_t Z,
                             const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                             const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                             const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                             const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype,
                             void* out, float* probs_values, const int32_t* head_ptr,
                             int N, int H, int T_DST, int T_SRC, int D, void* stream);

/* a8 + a9-a14 fused: the same attention as sea_sparse_attention_fwd, driven directly by the top-k bit mask (the CSR
 * column list is a pure function of it, sea_csr_fill); used when the caller does not need the CSR tensors.
 * 16-bit activations, D in {32,64,128}, P % 32 == 0, P <= 1024.  k_clamp = pconfig.k (pixel width clamp of a8). */
SEA_API int sea_sparse_attention_bits_fwd(const uint32_t* mask_bits,
                                          const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                          const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                          const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                          const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype, void* out,
                                          int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal, void* stream);

/* a8 + a9-a14 fused, short-context form (csrc/block_attn_umma.cu, csrc/block_attn.cu): the same result as
 * sea_sparse_attention_bits_fwd, computed as a tile-skipping, element-masked flash attention in which 128-row query blocks share
 * TMA-staged K/V tiles.  Replaces resize_from_m_to_t_csr -> flat_csr_masked_bmm -> flat_csr_softmax -> flat_csr_elmul ->
 * flat_csr_sdbmm (reference attention.py:1036-1042, 1159-1173) when the caller does not need the CSR tensors.
 *   bf16, T_SRC <= 4096: tcgen05 kernel; the a8 interpolation (causal_resize_m_to_t.py:648-762) runs inside it, straight from
 *     the top-k bit mask -- `workspace` is not touched (sea_block_attention_workspace_bytes returns a token 16 bytes);
 *   fp16 or T_SRC <= 8192: mma.sync kernel over the dense bit-packed partial_attention_mask (one u64 per head / query row /
 *     64-token tile) that an expansion kernel first writes into `workspace`.
 * Supported iff sea_block_attention_workspace_bytes(...) > 0: 16-bit activations, D == 64, P % 32 == 0, P <= 1024,
 * no clamped pixel (ceil(T_SRC / P) + 1 <= k_clamp) and T_SRC <= 8192 (work is O(T^2) in the worst case).
 * `workspace` must be 16-byte aligned; it is overwritten.  mask_bits must not be NULL. */
SEA_API int64_t sea_block_attention_workspace_bytes(int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int dtype);
SEA_API int sea_block_attention_fwd(const uint32_t* mask_bits,
                                    const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                    const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                    const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                    const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype, void* out,
                                    int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal,
                                    void* workspace, int64_t workspace_bytes, void* stream);

/* Backward of sea_sparse_attention_bits_fwd / sea_block_attention_fwd (csrc/sparse_attn_bwd.cu).  The reference has
 * no sparse backward kernel: its training path lets autograd differentiate the dense masked attention
 * (attention.py:1066-1133); this entry returns that gradient restricted to the alive (row, source token) pairs of the bit mask.
 *   dout [N,T_DST,H*D] (`dtype`) = grad of the forward output;  dq [N,H,T_DST,D], dk, dv [N,H,T_SRC,D] fp32 contiguous
 *   (dk / dv are zeroed here and accumulated with atomics);  dscales fp32 [N,H,T_DST,2] (nullable) = grad of the two scaler
 *   logits (attention.py:1166-1171, 1242-1244).  cumavg non-NULL adds the running-mean branch (causal prefill only).
 * The top-k mask is piecewise constant: no gradient flows through it.  D in {32, 64, 128}, P % 32 == 0, P <= 1024. */
SEA_API int sea_sparse_attention_bits_bwd(const uint32_t* mask_bits,
                                          const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                          const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                          const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                          const float* scales, const void* cumavg, int64_t avg_sh, int64_t avg_st, int use_scaler, int dtype,
                                          const void* dout, float* dq, float* dk, float* dv, float* dscales,
                                          int N, int H, int T_DST, int T_SRC, int D, int P, int k_clamp, int is_causal, void* stream);

/* Decode / use_cache form of a2 + a3 + a13 (csrc/performer_state.cu): replaces the reference's StatefulCausalPerformer /
 * StatefulCumAvg (attention_state.py:43-98, 205-224).  `state` holds per (n, h) the fp32 running sums S [F][2D] | z [F] | vsum [D]
 * (sea_performer_state_floats() floats, zero-initialised by the caller before the first token); the call advances it with the
 * T_new tokens at positions t0 .. t0+T_new-1 (q, k, v [N,H,T_new,D] strided like the other entries, pos_emb rows indexed by the
 * absolute position) and writes ctx [N,H,T_new,2D] and cumavg [N,H,T_new,D] (nullable) in `dtype`. */
SEA_API int64_t sea_performer_state_floats(int N, int H, int D, int F);
/* state += the sums of a whole prompt (tokens 0 .. T-1), chunk-parallel with fp32 atomics: what the prefill leaves for the decode. */
SEA_API int sea_performer_state_build(const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                      const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                      const float* pos_emb, const float* proj, int dtype, float* state,
                                      int N, int H, int T, int D, int F, void* stream);
SEA_API int sea_performer_causal_state_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                           const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                           const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                           const float* pos_emb, const float* proj, int dtype, float* state, void* ctx, void* cumavg,
                                           int N, int H, int T_new, int t0, int D, int F, void* stream);

/* Development aid: with SEA_ATTN_TRACE=1 in the environment sea_block_attention_fwd launches an instrumented instance of the tcgen05
 * attention kernel that records, per CTA and warp, eight cycle counters (where the softmax and MMA warps wait); this call copies the
 * counters of the last such launch to `host` ([ctas][20][8] uint32) and returns the number of CTAs (0: nothing traced). Synchronises. */
SEA_API int64_t sea_debug_attn_trace_read(uint32_t* host, int64_t max_words);

/* One decode step (use_cache, one new token per item) of the causal layer in ONE call (csrc/decode_step.cu; reference:
 * attention_state.py:43-236 + attention.py:559-572, 627-639, 774-947, 1151-1173, 1222-1244): incremental Performer + running mean,
 * predictor MLP on the token, the two dilated convs on 5-row windows, tail + softmax, top-k of the row, sparse attention of the row over
 * the KV cache k / v [N,H,t+1,D].  q [N,H,1,D].  All weights fp32 in the reference layouts (conv weights [C,C,5,3]; C > S*H = zero-padded
 * channels for the 64-channel tensor-core convs; conv3_w [H, S*H]); mlp_ws / conv1_ws / conv2_ws: packed-weight workspaces of the tensor-core
 * kernels (NULL -> SIMT kernels), re-packed when `repack` != 0.  k_per_row: device pointer to per_item_top_k of position t.
 * State (functional): perf_in -> perf_out (sea_performer_state_floats floats), xwin / ywin [N,4,W,C] = last 4 rows of the CNN input and of
 * conv 1's output.  Outputs: context [N,1,H*D] (`dtype`), probs fp32 [N,H,1,P].  workspace: sea_decode_step_workspace_bytes, 256-B aligned. */
SEA_API int64_t sea_decode_step_workspace_bytes(int N, int H, int D, int P, int S, int C, int k_clamp, int dtype);
SEA_API int sea_decode_step(const void* q, int64_t q_sn, int64_t q_sh,
                            const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                            const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype,
                            const float* pos_emb, const float* proj,
                            const float* enc_w, const float* enc_b, const float* enc_ln_w, const float* enc_ln_b,
                            const float* dec_w, const float* dec_b, const float* cnn_ln_w, const float* cnn_ln_b,
                            const float* scl_w, const float* scl_b,
                            const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                            const float* conv3_w, const float* conv3_b, const float* out_ln_w, const float* out_ln_b,
                            void* mlp_ws, void* conv1_ws, void* conv2_ws, int repack,
                            const float* k_per_row,
                            const float* perf_in, float* perf_out, const void* xwin_in, void* xwin_out, const void* ywin_in, void* ywin_out,
                            void* context, float* probs, void* workspace, int64_t workspace_bytes,
                            int N, int H, int D, int F, int P, int S, int C, int t, int k_clamp, int use_scaler, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Non-causal (BERT) variant, csrc/noncausal.cu.  No padding.
 * sea_performer_noncausal_fwd: v_for_atten = cat(grid-sampled identity, v) (attention.py:462-502) and the FAVOR+
 *   softmax-feature Performer with un-prefixed sums (performer-pytorch softmax_kernel / linear_attention):
 *   ctx [N,H,T,2D] (`dtype`); workspace fp32 >= sea_performer_noncausal_workspace_floats().
 * sea_conv3x3_cl: Conv2d(C,O,3,padding=1, stride (stride_t,1)) on channels-last [N,Tin,W,C] -> [N,Tout,W,O], input rows
 *   nearest-upsampled by `up` first (attention.py:209-215); weight fp32 [O,C,3,3].
 * sea_bert_tail_fwd: bilinear (align_corners=False) resize of [N,Tin,Win,H] to (T,P) (KeepRes, modules.py:42-55) +
 *   softmax(P) -> probs fp32 [N,H,T,P].
 * sea_topk_mask_bits_batch: k_flatten_dim='batch' (attention.py:833-837): one top-k group per item over H*T*P keys in
 *   view(N, H*T*P) order; k_per_item [N] fp32; ties to the lower flat index; bit layout as sea_topk_mask_bits.
 * sea_bert_avg_fwd: probability-weighted mean of v (attention.py:1209-1219) -> avg [N,H,D] (`dtype`). */
SEA_API int64_t sea_performer_noncausal_workspace_floats(int N, int H, int T, int D, int F);
SEA_API int sea_performer_noncausal_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                        const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                        const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                        const float* proj, int dtype, void* ctx, float* workspace,
                                        int N, int H, int T, int D, int F, void* stream);
/* right-padded batch (attention.py:447-449, 482, 512-514): lengths (int32 [N], nullable) = valid tokens of item n; the identity grid
 * follows the rank among the valid tokens and v_for_atten is zero on the padded ones (k is NOT masked, like the reference). */
SEA_API int sea_performer_noncausal_len_fwd(const void* q, int64_t q_sn, int64_t q_sh, int64_t q_st,
                                            const void* k, int64_t k_sn, int64_t k_sh, int64_t k_st,
                                            const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st,
                                            const float* proj, int dtype, void* ctx, float* workspace, const int32_t* lengths,
                                            int N, int H, int T, int D, int F, void* stream);
SEA_API int sea_conv3x3_cl(const void* x, const float* weight, const float* bias, void* y, int dtype,
                           int N, int Tin, int Tout, int W, int C, int O, int stride_t, int up, int relu, void* stream);
SEA_API int sea_bert_tail_fwd(const void* y, int dtype, float* probs, float* scores, int N, int H, int Tin, int Win, int T, int P, void* stream);
SEA_API int sea_topk_mask_bits_batch(const float* keys, const float* k_per_item, uint32_t* mask_bits, int N, int H, int T, int P, void* stream);
/* the same selection spread over ceil(keys / 4096) CTAs per group (a dozen small launches); scratch in `workspace`.
 * group_heads = H: k_flatten_dim='batch' (one group per item); group_heads = 1: k_flatten_dim='head' (attention.py:838-842, one group per
 * (item, head) over its T*P keys); k_per_group has N * H / group_heads entries. */
SEA_API int64_t sea_topk_batch_workspace_bytes(int N, int H, int T, int P, int group_heads);
SEA_API int sea_topk_mask_bits_batch_ws(const float* keys, const float* k_per_group, uint32_t* mask_bits, void* workspace, int64_t workspace_bytes,
                                        int N, int H, int T, int P, int group_heads, void* stream);
SEA_API int sea_bert_avg_fwd(const float* probs, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype, void* avg,
                             int N, int H, int T, int P, int D, void* stream);
/* the same with per-item token lengths (right-padded batch): the resize width is lengths[n], padded tokens weigh 0 */
SEA_API int sea_bert_avg_len_fwd(const float* probs, const void* v, int64_t v_sn, int64_t v_sh, int64_t v_st, int dtype, void* avg,
                                 const int32_t* lengths, int N, int H, int T, int P, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEA_B200_H_ */
