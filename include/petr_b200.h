/*
 * petr_b200.h -- C ABI of the B200-native PETR detection-head hot path (libpetr_b200.so).
 *
 * Drop-in boundary for sty61010/PETR's head hot path.  The reference has no native code and no
 * FFI of its own: the hot path is straight-line PyTorch inside
 *   projects/mmdet3d_plugin/models/dense_heads/petr_head.py      (PETRHead.forward :366-468,
 *                                                                  position_embeding :286-334)
 *   projects/mmdet3d_plugin/models/dense_heads/petrv2_head.py    (PETRv2Head.forward :429-540)
 *   projects/mmdet3d_plugin/models/utils/petr_transformer.py     (PETRMultiheadAttention :280-367,
 *                                                                  decoder :423-447)
 *   projects/mmdet3d_plugin/models/utils/positional_encoding.py  (SinePositionalEncoding3D :58-100)
 * so each entry point below cites the reference lines whose library calls it replaces.  The
 * binding a maintainer adds is a ctypes stub (see INTEGRATION.md and petr_b200/_lib.py).
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the parameter name starts with `h_` (host, read
 *    during the call).  The caller allocates every input, output and workspace; the library
 *    never allocates, frees or retains device memory.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), does not
 *    synchronise, and is CUDA-graph capturable.  Device = the calling thread's current device.
 *  - Return value: 0 = success; < 0 = argument/shape/alignment error found on the host before
 *    any launch (PETR_E_*); > 0 = a cudaError_t from the launch.  petr_last_error_string()
 *    returns a thread-local description of the last non-zero return.  Nothing throws or exits.
 *  - "tokens" are the position-aware image keys, flattened (b, n, h, w) with w fastest:
 *    T = B*N*H*W rows, Lk = N*H*W per sample (order of petr_transformer.py:90).
 *  - Query-side rows are flattened (b, q): row = b*Lq + q.
 */
#ifndef PETR_B200_H_
#define PETR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PETR_B200_ABI_VERSION 3

#define PETR_E_BADARG   (-1)   /* null pointer / non-positive dimension */
#define PETR_E_SHAPE    (-2)   /* unsupported shape (see each function) */
#define PETR_E_ALIGN    (-3)   /* pointer or leading dimension not aligned as required */
#define PETR_E_ARCH     (-4)   /* device is not sm_100 */

#define PETR_DTYPE_F32  0
#define PETR_DTYPE_BF16 1

#define PETR_ACT_NONE         0
#define PETR_ACT_RELU         1
#define PETR_ACT_SIGMOID_GATE 2   /* out = gate[m,n] * sigmoid(acc + bias)  (PETRv2 SELayer) */

#define PETR_A_ROWMAJOR 0         /* A[m*lda + k] */
#define PETR_A_NCHW     1         /* A[(m/hw)*K*hw + k*hw + (m%hw)]: conv1x1 over (G,K,h,w) maps */

int petr_abi_version(void);
const char* petr_last_error_string(void);
/* number of kernels this library has launched on the calling thread since the last reset */
int64_t petr_launch_count(void);
void petr_launch_count_reset(void);
/* Explicit, process-wide tuning switches (A/B measurements and tests; the library reads no
 * environment variables).  Keys: "flash_exact_only" (0/1: skip the max-free fast pass of
 * petr_cross_attention_bf16), "flash_poly" (0..4: exponential pairs per 4 evaluated on the FMA pipe in
 * the fast pass), "gemm_staged" / "gemm_coalesce" / "gemm_skinny" / "gemm_band" (0/1: epilogue and tile-order variants of
 * petr_gemm_bf16_tc), "gemm_debug" (diagnostic, 0 = off: 1 = out_mode 2 skips its global stores, 2 = every epilogue
 * reduced to its TMEM loads -- the MMA-bound floor that tools/kproj_store_bound.py and tools/qside_epilogue_floor.py
 * measure; results are then garbage by design), "pdl" (0/1, default 0: launch the GEMM, LayerNorm, attention and merge
 * kernels with programmatic stream serialization -- griddepcontrol.wait inside the kernels orders the data -- so that a
 * kernel's launch and prologue overlap its predecessor's tail; measured SLOWER inside the captured graph, 1.324 vs
 * 1.287 ms at B = 1 and 15.22 vs 15.09 ms at B = 32, hence off).  Returns 0, or PETR_E_BADARG for an unknown key / out-of-range value;
 * petr_get_tuning returns the current value or -1. */
int petr_set_tuning(const char* key, int value);
int petr_get_tuning(const char* key);
/* Diagnostic: device buffer of int64[4][64][8] that CTA 0 of the tcgen05 GEMM fills with clock64() stamps of
 * its pipeline events (tools/gemm_trace.py); NULL (default) disables it. */
void petr_debug_set_trace(void* device_buffer);

/* ---- a4: key-padding masks.  petr_head.py:383-394 (ones(pad) / zero img_shape /
 * F.interpolate nearest / .to(bool)).  img_hw: int32 [B*N*2] = (img_h, img_w) per view.
 * mask: uint8 [B,N,H,W], 1 = padded key.  Bit-exact. */
int petr_key_padding_mask(const int32_t* img_hw, int B, int N, int pad_h, int pad_w,
                          int H, int W, uint8_t* mask, void* stream);

/* ---- a1 (first half): frustum -> img2lidar -> normalise -> coords_mask -> inverse_sigmoid.
 * petr_head.py:286-331 (== petrv2_head.py:348-394).  img2lidar: fp32 [B,N,4,4], the float64
 * host inverse of lidar2img rounded once to fp32 (petr_head.py:308-315; done by the caller).
 * h_range_lo[i] = (float)position_range[i], h_range_span[i] = (float)(position_range[3+i] -
 * position_range[i]), bin_size = (float)bin_size_double (petr_head.py:296/:301).
 * feat: [T, 3*D] (channel = 3*d + c), fp32 or bf16.  masks (nullable): uint8 [B,N,H,W] OR'ed
 * into coords_mask (nullable, uint8 [B,N,H,W]).  coords3d (nullable): fp32 [B,N,H,W,D,3], the
 * normalised coordinates before inverse_sigmoid, for bit-exactness tests.
 * Grid, depth bins, transform and normalisation use separately rounded fp32 mul/add/div (no FMA
 * contraction) and are bit-exact against the CPU reference; log is within 2 ulp. */
int petr_frustum_features(const float* img2lidar, const uint8_t* masks,
                          int B, int N, int H, int W, int D, int pad_h, int pad_w,
                          int lid, float depth_start, float bin_size,
                          const float* h_range_lo, const float* h_range_span,
                          void* feat, int feat_dtype, uint8_t* coords_mask, float* coords3d,
                          void* stream);

/* ---- a3 (first half): SinePositionalEncoding3D features.  positional_encoding.py:58-100.
 * mask uint8 [B,N,H,W]; dim_t fp32 [num_feats] (= temperature ** (2*(i//2)/num_feats), computed
 * by the caller exactly as positional_encoding.py:81-83).  out: [T, 3*num_feats] with the
 * reference's channel order (n | y | x, each = [sin of even dims | cos of odd dims]). */
int petr_sine3d_features(const uint8_t* mask, const float* dim_t, int B, int N, int H, int W,
                         int num_feats, int normalize, float scale, float eps, float offset,
                         void* out, int out_dtype, void* stream);

/* ---- a6: pos2posemb3d.  petr_head.py:31-43.  ref fp32 [Q,3] -> out fp32 [Q, 3*num_feats]
 * in (y, x, z) order, interleaved sin/cos. */
int petr_pos2posemb3d(const float* ref, const float* dim_t, int Q, int num_feats, float* out,
                      void* stream);

/* ---- fp32 GEMM with fused prologue/epilogue: every nn.Linear / 1x1 Conv2d on the path
 * (petr_head.py:219-274 layers; FFN; MultiheadAttention in/out projections).
 *   C[m,n] = epi( sum_k (A[m,k] + A2[m % a2_rows, k]) * W[n,k] + bias[n] ) + residual[m,n]
 * A2/bias/residual/gate/C_bf16 nullable.  residual, gate and C share ldc.  C_bf16 (nullable) receives
 * a bf16 copy of C with the same ldc.  W is [N,K] row-major (nn.Linear / conv weight layout).
 * Exact fp32 FFMA accumulation (fp32-parity path: <= 1e-6 relative to tensor scale). */
int petr_gemm_f32(const float* A, int lda, int a_layout, int hw,
                  const float* A2, int a2_rows,
                  const float* W, const float* bias, const float* residual, const float* gate,
                  int act, float* C, void* C_bf16, int ldc, int M, int N, int K, void* stream);

/* ---- a9/a10: multi-head attention core (softmax(Q K^T * scale + mask) V), fp32, flash-style
 * streaming over keys; no score matrix is materialised.  torch.nn.MultiheadAttention as called
 * by petr_transformer.py:357-362 (packed projections are separate petr_gemm_* calls).
 * Q rows b*Lq+q (ldq), K/V rows b*Lk+k (ldk/ldv), head h occupies columns [h*head_dim, ...).
 * key_mask nullable uint8 [B,Lk], 1 = ignore.  A row whose keys are all masked yields NaN,
 * like the reference (zeroed later by nan_to_num, petr_head.py:435).  head_dim must be 32.
 * O (fp32, nullable) and/or O_split (nullable: bf16 rows [hi | lo] of width 2*heads*head_dim, the
 * A operand of a bf16x3 petr_gemm_bf16_tc) receive the result.
 * attn_bias (nullable): the `attn_mask` of torch.nn.MultiheadAttention (petr_transformer.py:357-362) as an additive
 * fp32 array [Lq, Lk] (bias_per_head 0) or [B*heads, Lq, Lk] (bias_per_head 1); -inf entries mask.
 * Training (section 8f row 3): lse (nullable) fp32 [B, heads, Lq] receives log2-sum-exp2 of the scaled scores for
 * petr_attention_f32_bwd; drop_p > 0 applies attention-probability dropout (nn.MultiheadAttention's `dropout`) decided
 * by a counter-based hash of (seed, row, key) that the backward kernel re-evaluates.
 * ksplit > 1: the key axis is split over ksplit blocks per (row block, head, sample) -- one thread owns a query row, and
 * 900 rows x 8 heads do not fill the chip at small batch -- through `partial` (ksplit * petr_attention_partial_floats
 * floats, caller-allocated) and an internal petr_attention_merge; ksplit <= 1: single pass, partial may be null. */
int petr_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                       const uint8_t* key_mask, const float* attn_bias, int bias_per_head,
                       float* O, int ldo, void* O_split, float* lse, float drop_p, uint64_t seed,
                       float* partial, int ksplit,
                       int B, int Lq, int Lk, int heads, int head_dim, float scale, void* stream);

/* ---- nn.LayerNorm over the last dim (eps 1e-5 on this path), optional fused ReLU
 * (cls_branches, petr_head.py:226-231) and nan_to_num (petr_head.py:435).  Y may alias X. */
int petr_layernorm_f32(const float* X, const float* gamma, const float* beta, float* Y,
                       int rows, int E, float eps, int relu, int nan_to_num, void* stream);

/* Fused LayerNorm of the decoder residual stream (E = 256).  Optional outputs: Y = LN(X) fp32;
 * split(Y) and split(Y + qpos[row % qpos_rows]) as bf16 [hi | lo] rows (bf16x3 GEMM inputs:
 * query+query_pos of petr_transformer.py:341); Z = LN2(Y) with nan_to_num (the decoder's
 * post_norm on a copy, petr_transformer.py:444 + petr_head.py:435) as fp32 and/or split bf16.
 * relu != 0 applies ReLU to Y (cls_branches, petr_head.py:226-231).  Every pointer must be 16-byte
 * aligned (one warp per row, 16-byte loads and stores).  nparts > 1: X is the sum of nparts arrays part_stride floats
 * apart -- the split-K partial products of the preceding petr_gemm_bf16_tc (ksplit), added in a fixed order. */
int petr_layernorm_multi_f32(const float* X, const float* gamma, const float* beta, float eps,
                             const float* qpos, int qpos_rows, float* Y, void* Y_split,
                             void* Ypos_split, const float* gamma2, const float* beta2,
                             float* Z, void* Z_split, int rows, int E, int relu, int nparts,
                             int64_t part_stride, void* stream);

/* fp32 rows -> bf16 [hi | lo] rows of width 2*E: split(relu?(X + add[row % add_rows])). */
int petr_split_rows_bf16(const float* X, int ldx, const float* add, int add_rows, void* out,
                         int rows, int E, int relu, void* stream);

/* fp32 weight W[N,K] -> bf16 W operand of a bf16x3 GEMM: out[N, 3K] = [hi | hi | lo] of W (transpose 0), or
 * out[K, 3N] = the same of W^T (transpose 1; the operand of the input-gradient GEMM dA = dY W). */
int petr_x3_weight_bf16(const float* W, int N, int K, int transpose, void* out, void* stream);

/* Copies nbytes (multiple of 4) from PINNED host memory (UVA-mapped; the one host pointer this ABI
 * reads from the device) to device memory with a tiny kernel instead of the copy engine, so the
 * per-call meta data (img2lidar, img_shape, timestamps) never queues behind a concurrent bulk
 * feature upload on another stream. */
int petr_stage_small(const void* h_pinned, void* dst, int nbytes, void* stream);

/* out[r, :] = bf16(X[r, :] + add[r % add_rows, :]); E % 4 == 0. */
int petr_add_rows_bf16(const float* X, const float* add, int add_rows, void* out, int rows, int E,
                       void* stream);

/* ---- a12: reference-point add, sigmoid and pc_range affine.  petr_head.py:440-460
 * (+ PETRv2 with_time petrv2_head.py:499-505,:520-521).  reg: fp32 [L*B*Q, code_size] in place.
 * ref: fp32 [Q,3] (reference_points.weight).  h_pc_range: host float[6].  time_div nullable
 * fp32 [B]: channels 8.. are divided by time_div[b]. */
int petr_decode_boxes(float* reg, const float* ref, const float* h_pc_range,
                      const float* time_div, int L, int B, int Q, int code_size, void* stream);

/* ---- f1: NMS-free box decoding.  NMSFreeCoder.decode_single
 * (core/bbox/coders/nms_free_coder.py:48-101) for every sample of the batch in one launch: sigmoid ->
 * top-max_num over the Q*C (query, class) scores -> label = idx %% C, query = idx / C -> gather the query's box ->
 * denormalize_bbox (core/bbox/util.py:63-87: (cx, cy, log w, log l, cz, log h, sin, cos[, vx, vy]) ->
 * (cx, cy, cz, w, l, h, atan2(sin, cos)[, vx, vy])) -> keep = centre inside h_post_center_range[6] (host)
 * [&& score > score_threshold when has_score_threshold]; bottom_center != 0 additionally applies
 * PETRHead.get_bboxes' z -= h / 2 (petr_head.py:744) after the range test.
 *   cls_scores fp32 [B, Q, C] (logits of the last decoder layer), bbox_preds fp32 [B, Q, code], code 8 or 10.
 * Outputs are padded to max_num rows per sample, sorted by descending score (ties: lower flat index first):
 *   out_boxes fp32 [B, max_num, code > 8 ? 9 : 7], out_scores fp32 [B, max_num], out_labels / out_index int64
 *   [B, max_num] (class, query), out_mask uint8 [B, max_num] (1 = kept), out_count int32 [B] (kept rows).
 * Labels / indices are integer-exact (ranking on the logits); Q*C <= 16384. */
int petr_nms_free_decode(const float* cls_scores, const float* bbox_preds, int B, int Q, int C, int code,
                         int max_num, const float* h_post_center_range, int has_score_threshold,
                         float score_threshold, int bottom_center, float* out_boxes, float* out_scores,
                         int64_t* out_labels, int64_t* out_index, uint8_t* out_mask, int32_t* out_count,
                         void* stream);

/* ---- tcgen05 tier (sm_100a tensor cores, TMA-fed, TMEM accumulators) ----------------- */

/* NCHW feature maps x[G][C][hw] (in_dtype PETR_DTYPE_F32, or PETR_DTYPE_BF16 for maps handed over in half
 * precision) -> bf16 token-major out[G*hw][C] (A operand of input_proj, petr_head.py:390). */
int petr_nchw_to_tokens_bf16(const void* x, int in_dtype, void* out, int G, int C, int hw, void* stream);

/* bf16 x bf16 -> fp32 GEMM on tcgen05 (persistent, 128x256 tiles, TMA-fed, TMEM accumulators):
 *   C = alpha * act(A[M,K] * W[N,K]^T + bias) (+ residual)
 * A, W: bf16 row-major, K % 64 == 0, 16-byte aligned rows.  a_k_wrap (0 = K): A has only
 * a_k_wrap columns and column k of the product reads A[:, k % a_k_wrap] -- with A = [hi | lo]
 * (2K0 columns), W = [W_hi | W_hi | W_lo] (3K0 columns) and a_k_wrap = 2K0 this is the bf16x3
 * product A_hi W_hi + A_lo W_hi + A_hi W_lo, accurate to ~2^-16 (fp32-class query-side GEMMs).
 * out_mode selects the store:
 *   0: fp32 C[m*ldc+n];  1: bf16 C[m*ldc+n];
 *   2: bf16 head-split layout C[n / 256][b][(n % 256) / 32][key][32] with b = m / kv_Lk,
 *      key = m % kv_Lk -- the Q/K/V layout petr_cross_attention_bf16 reads;
 *   3: bf16 split C[m*2N + n] = hi, C[m*2N + N + n] = lo, plus (C2 nullable) fp32 C2[m*ldc+n];
 *   4: head-split [hi | lo] pair: C = hi array, C2 = lo array, both in the layout of mode 2 (the
 *      operands of petr_attention_x3_tc);
 *   5: two results from one accumulator: C = bf16 [m*N + n] of alpha*act(acc + bias) WITHOUT the residual and
 *      C2 = fp32 [m*ldc + n] WITH residual[(m %% kv_Lk)*ldc + n] added -- input_proj emitting the V-projection
 *      operand and "memory + batch-invariant sine embedding" (the K-side residual of the position MLP) at once.
 * residual (nullable, fp32, ldc) is added after alpha.  act PETR_ACT_SIGMOID_GATE (out_mode 0) multiplies
 * sigmoid(acc + bias) by gate[m*ldc + n] (fp32) before the residual: the PETRv2 SELayer
 * (petrv2_head.py:48-60) as a GEMM epilogue.  Used for the K/V projection of all
 * decoder layers in one launch (petr_transformer.py:341-344,:357-360), the position_encoder /
 * adapt_pos3d MLPs, input_proj and the query-side projections / FFN.
 * ksplit > 1 (out_mode 0 only): split-K for GEMMs with few output tiles (M = B * 900 rows at small batch leave most SMs
 * idle): part s of the K range writes its partial product to C + s * M * ldc (bias and residual go into part 0);
 * the consumer (petr_layernorm_multi_f32 nparts) adds the parts in a fixed order, so results stay deterministic. */
int petr_gemm_bf16_tc(const void* A, int lda, int a_k_wrap, const void* W, int ldw,
                      const float* bias, const float* residual, const float* gate, int act, float alpha,
                      void* C, void* C2, int ldc, int out_mode, int M, int N, int K,
                      int kv_Lk, int ksplit, void* stream);

/* 3xTF32 GEMM on tcgen05 (kind::tf32; same persistent kernel, fp32 operands): the fp32-PARITY path of the 1x1-conv
 * MLPs (petr_head.py:332 position_encoder, :249-254 adapt_pos3d) and the fp32 K / V projections on the tensor cores.
 *   C = act(A[M,K] * W[N,K]^T + bias) (+ residual),  K % 32 == 0, fp32 row-major operands with 16-byte aligned rows.
 * With A = [hi | lo] (2K0 columns; petr_split_rows_tf32 or out_mode 6 of a previous call: hi = value rounded to
 * tf32, lo = value - hi), W = [W_hi | W_hi | W_lo] (3K0 columns) and a_k_wrap = 2K0 the product is
 * hi*hi + lo*hi + hi*lo: ~2^-21 relative, i.e. fp32-class (the 1e-5 position-embedding tolerance has 20x margin).
 * out_mode 0: fp32 C[m*ldc + n];  6: fp32 split rows C[m*2N + n] = hi, C[m*2N + N + n] = lo (ldc = residual stride). */
int petr_gemm_tf32x3_tc(const float* A, int lda, int a_k_wrap, const float* W, int ldw, const float* bias,
                        const float* residual, int act, float* C, int ldc, int out_mode, int M, int N, int K,
                        void* stream);
/* fp32 rows X[rows, E] (row stride ldx) (+ add[r % add_rows], nullable) -> fp32 [hi | lo] rows out[rows, 2E]. */
int petr_split_rows_tf32(const float* X, int ldx, const float* add, int add_rows, float* out, int rows, int E,
                         void* stream);

/* ---- a1, fused (K1): the whole position embedding of petr_head.py:286-334 in one kernel per
 * 128-token tile: frustum -> img2lidar -> normalise -> inverse_sigmoid -> bf16 A operand written
 * straight into 128B-swizzled shared memory -> tcgen05 GEMM (192 -> hidden, TMEM) -> bias + ReLU ->
 * bf16 shared-memory operand -> tcgen05 GEMM (hidden -> embed, TMEM) -> + bias (+ residual) -> fp32
 * tokens out[T, embed].  The coordinate tensor, the MLP input and the hidden activation never
 * reach HBM; W1 [hidden,192] / W2 [embed,hidden] (bf16) stream through TMA from L2.
 * Specialised for depth_num 64, hidden 1024, embed 256 (every shipped config).  Arguments as
 * petr_frustum_features; residual (nullable) fp32 [T, embed]; residual2 (nullable) fp32
 * [res2_rows, embed] added with row t % res2_rows (the batch-invariant sine embedding); the result
 * goes to out (fp32, nullable) and/or out_bf16 (nullable) -- with both residuals the bf16 output
 * is directly the K-projection input mem + pos_embed of petr_transformer.py:343. */
int petr_position_mlp_fused(const float* img2lidar, int B, int N, int H, int W, int D,
                            int pad_h, int pad_w, int lid, float depth_start, float bin_size,
                            const float* h_range_lo, const float* h_range_span,
                            const void* W1, const float* b1, const void* W2, const float* b2,
                            int hidden, int embed, const float* residual, const float* residual2,
                            int res2_rows, float* out, void* out_bf16, void* stream);

/* fp32 Q [B*Lq, heads*32] (row stride ldq) -> bf16 [B, heads, Lq, 32] multiplied by `scale`
 * (head_dim^-1/2 * log2 e: the softmax of petr_cross_attention_bf16 runs in the exp2 domain). */
int petr_pack_q_bf16(const float* Q, int ldq, void* out, int B, int Lq, int heads, float scale,
                     void* stream);

/* Flash cross-attention on tcgen05 (QK^T -> TMEM, softmax in registers, P -> bf16, PV -> TMEM),
 * petr_transformer.py:357-362 for one decoder layer.
 *   Qs : bf16 [B, heads, Lq, 32]  (petr_pack_q_bf16)
 *   Kh : bf16 [B, heads, Lk, 32]  (out_mode 2 of petr_gemm_bf16_tc)
 *   Vh : bf16 [B, heads, Lk, 32]
 *   key_mask nullable uint8 [B, Lk], 1 = ignore.
 * nsplit key splits run as separate CTAs (split-KV for small B).  write_partial == 0: O fp32
 * [B*Lq, heads*32] (row stride ldo, nullable) and/or O_split (bf16 [hi | lo] rows of width
 * 2*heads*32, nullable) receive the normalised result (partials of nsplit > 1 are
 * merged internally; `partial` is then the scratch, petr_attention_partial_floats * nsplit).
 * write_partial == 1: only the un-normalised partials (O, m, l) are written to `partial`
 * (layout below) for a later petr_attention_merge -- the key-split-across-GPUs path.
 * workspace (nullable): petr_cross_attention_workspace_bytes(...) bytes of device scratch.  With
 * it the call runs a fast "max-free" softmax pass (p = 2^s, no running maximum; valid while every
 * row sum stays inside [2^-80, 2^100], which the pass verifies per row) followed by the exact
 * online-softmax pass restricted to the CTAs whose check failed (fully masked rows, extreme
 * logits); without it only the exact pass runs.  Results are within bf16 rounding either way. */
int petr_cross_attention_bf16(const void* Qs, const void* Kh, const void* Vh,
                              const uint8_t* key_mask, float* O, int ldo, void* O_split,
                              float* partial, void* workspace, int B, int Lq, int Lk, int heads,
                              int nsplit, int write_partial, void* stream);
int64_t petr_cross_attention_workspace_bytes(int B, int Lq, int Lk, int heads, int nsplit);
/* Diagnostic (not a product path): the fast pass of petr_cross_attention_bf16 with clock64() stamps of
 * CTA 0's pipeline events written to trace[7 roles][64 key tiles][8 events] (int64); the partials go
 * to `partial` (nsplit * petr_attention_partial_floats).  See tools/flash_trace.py. */
int petr_cross_attention_trace(const void* Qs, const void* Kh, const void* Vh, float* partial,
                               void* workspace, long long* trace, int B, int Lq, int Lk, int heads,
                               int nsplit, void* stream);

/* fp32-class attention on tcgen05 (the decoder's self-attention; mmcv MultiheadAttention called by
 * BaseTransformerLayer with q = k = query + query_pos, v = query).  Every operand is a bf16
 * [hi | lo] pair (out_mode 4 of petr_gemm_bf16_tc), all bf16 [B, heads, L, 32]:
 *   S = Qh Kh^T + Ql Kh^T + Qh Kl^T,  P = Ph + Pl,  O = Ph Vh + Pl Vh + Ph Vl   (~2^-16 accuracy).
 * `scale` (head_dim^-1/2) is applied inside the softmax.  Outputs / splits / masks / workspace (fast max-free pass +
 * exact pass on flagged CTAs; petr_attention_x3_workspace_bytes) as petr_cross_attention_bf16. */
int petr_attention_x3_tc(const void* Qh, const void* Ql, const void* Kh, const void* Kl,
                         const void* Vh, const void* Vl, const uint8_t* key_mask, float* O,
                         int ldo, void* O_split, float* partial, void* workspace, int B, int Lq, int Lk,
                         int heads, float scale, int nsplit, int write_partial, void* stream);
int64_t petr_attention_x3_workspace_bytes(int B, int Lq, int Lk, int heads, int nsplit);

/* Merge `nparts` partial attention results (on-GPU split-KV or key-split across GPUs after an
 * all-gather):  m = max m_r;  O = sum e^{m_r-m} O_r / sum e^{m_r-m} l_r.
 * partial: fp32 [nparts][ rows*32 (O) | rows*2 (m, l) ], rows = B*heads*Lq ordered (b, head, q);
 * m is in the log2 domain.  lse (nullable): fp32 [B, heads, Lq] log2-sum-exp2 of the merged rows. */
int petr_attention_merge(const float* partial, int nparts, float* O, int ldo, void* O_split, float* lse,
                         int B, int Lq, int heads, void* stream);

int64_t petr_attention_partial_floats(int B, int Lq, int heads);

/* ---- Either side of the main path (reference branches no shipped config takes) ---- */

/* mmdet SinePositionalEncoding (mmdet 2.24.1, un-vendored; called per camera by the with_multiview=False branch,
 * petr_head.py:403-409,:414-420).  mask uint8 [G,H,W] -> out [G*H*W, 2*num_feats] = [pos_y | pos_x]; inside a half,
 * even channel j = sin(v / dim_t[j]), odd = cos(v / dim_t[j]). */
int petr_sine2d_features(const uint8_t* mask, const float* dim_t, int G, int H, int W, int num_feats,
                         int normalize, float scale, float eps, float offset, void* out, int out_dtype,
                         void* stream);

/* LearnedPositionalEncoding3D.forward (positional_encoding.py:137-160): out fp32 [G*H*W, 2*num_feats] =
 * [col_embed[w] | row_embed[h]]. */
int petr_learned_pos_features(const float* row_embed, int row_num, const float* col_embed, int col_num,
                              int G, int H, int W, int num_feats, float* out, void* stream);

/* token-major fp32 [G*hw, C] -> NCHW fp32 [G, C, hw] (the layout the reference's modules return). */
int petr_tokens_to_nchw_f32(const float* x, float* out, int G, int C, int hw, void* stream);

/* CPFPN top-down step (cp_fpn.py:176-184, F.interpolate nearest + add) on token-major maps:
 * dst[g,y,x,:] += src[g, floor(y*h/H), floor(x*w/W), :].  dst fp32 [G,H,W,C], src fp32 [G,h,w,C]. */
int petr_upsample_add_tokens(float* dst, const float* src, int G, int H, int W, int h, int w, int C,
                             void* stream);

/* im2col of CPFPN's level-0 3x3 / pad 1 conv (cp_fpn.py:126-136,:188-190) over a token-major fp32 map [G,H,W,C]:
 * out[t, tap*C + c], tap = 3*ky + kx; out_mode 0 fp32, 1 bf16, 2 bf16 [hi | lo] halves (bf16x3 GEMM operand).  The
 * conv then is one petr_gemm_* with the weight repacked to [Cout][tap][Cin]. */
int petr_im2col3x3_tokens(const float* x, void* out, int out_mode, int G, int H, int W, int C, void* stream);

/* mmdet NormedLinear operand scaling (petr_head.py:231-232): out = x * mult / (||x||^power + eps), the norm taken
 * over the query axis of [B,Q,E] (along_queries 1: what `x.norm(dim=1)` does to the head's 3-D activations) or
 * over each row of [B*Q, E] (along_queries 0: the weight). */
int petr_norm_scale_f32(const float* x, float* out, int B, int Q, int E, int along_queries, float power,
                        float eps, float mult, void* stream);

/* ---- Training path: backward of the fp32 forward kernels above, the loss of
 * petr_head.py:578-728 and the matching cost of hungarian_assigner_3d.py:117-123.  Exact fp32 arithmetic, except that the host side
 * routes the large Linear layers through petr_gemm_bf16_tc in bf16x3 mode (16 mantissa bits per operand). ---- */

/* dW[N,K] += dY[M,N]^T (A[M,K] + A2[m % a2_rows]) -- weight gradient of petr_gemm_f32 (A layouts as there).  dW must
 * hold zeros or a running gradient (partial sums over row chunks are added atomically). */
int petr_gemm_f32_tn(const float* dY, int ldy, const float* A, int lda, int a_layout, int hw,
                     const float* A2, int a2_rows, float* dW, int M, int N, int K, void* stream);
/* Operands of the weight-gradient GEMM on tcgen05 (dW = dY^T (A + A2) is K-major over m once both factors are
 * transposed): fp32 X[M, C] (+ add[m % add_rows, C]) -> bf16 out[C, w * Mpad], w3 0: [hi | lo] (A operand of a bf16x3
 * petr_gemm_bf16_tc), w3 1: [hi | hi | lo] (its W operand); Mpad >= M, a multiple of 64 for the GEMM; columns >= M are 0. */
int petr_split_transpose_bf16(const float* X, int ldx, const float* add, int add_rows, int M, int C, int Mpad, int w3,
                              void* out, void* stream);
/* out[i] = sum over p < nparts of parts[p * n + i] (split-K parts of a GEMM; n a multiple of 4). */
int petr_sum_parts_f32(const float* parts, int nparts, int64_t n, float* out, void* stream);
/* out[N] += column sums of X[M,N] (bias gradient). */
int petr_colsum_f32(const float* X, int ldx, int M, int N, float* out, void* stream);
/* out[period_rows, E] += sum_j X[j*period_rows + r, :] (gradient of the row-periodic A2 / query_pos add). */
int petr_rowsum_periodic_f32(const float* X, int rows, int E, int period_rows, float* out, void* stream);
/* dX = dY where Y > 0 else 0 (ReLU epilogue). */
int petr_relu_bwd_f32(const float* dY, const float* Y, float* dX, int64_t n, void* stream);
/* out = gate * sigmoid(z) (PETR_ACT_SIGMOID_GATE): dZ = dOut * gate * s(1-s), dGate = dOut * s with s = sigmoid(Zpre). */
int petr_sigmoid_gate_bwd_f32(const float* dOut, const float* Zpre, const float* gate, float* dZ, float* dGate,
                              int64_t n, void* stream);
/* nn.Dropout in training, forward and backward alike: Y[i] = (keep(seed, i) ? X[i] / (1 - p) : 0) + residual[i]
 * (residual nullable: the `identity + dropout(out)` of petr_transformer.py:367 in one pass). */
int petr_dropout_f32(const float* X, const float* residual, float* Y, int64_t n, float p, uint64_t seed,
                     void* stream);
/* Backward of petr_norm_scale_f32 (NormedLinear operand scaling): dX from dOut and the forward input x; same group
 * layout switch.  A zero-norm group passes dOut * mult / eps through (x is zero there). */
int petr_norm_scale_bwd_f32(const float* x, const float* dOut, float* dX, int B, int Q, int E, int along_queries,
                            float power, float eps, float mult, void* stream);
/* Backward of petr_learned_pos_features: token gradients [G*H*W, 2*num_feats] summed into d_col_embed[W, num_feats]
 * and d_row_embed[H, num_feats] (the first W / H rows of the embedding tables; overwritten, not accumulated). */
int petr_learned_pos_bwd(const float* dOut, int G, int H, int W, int num_feats, float* d_row_embed,
                         float* d_col_embed, void* stream);
/* nn.LayerNorm backward (E = 256): dX, and dgamma / dbeta (nullable) accumulated atomically.  relu / nan_to_num as in
 * petr_layernorm_f32 (their masks are re-derived from the recomputed output). */
int petr_layernorm_bwd_f32(const float* X, const float* gamma, const float* beta, const float* dY, float* dX,
                           float* dgamma, float* dbeta, int rows, int E, float eps, int relu, int nan_to_num,
                           void* stream);
/* Backward of petr_attention_f32 (same operands, lse / drop_p / seed from the forward call): dQ, dK, dV in the
 * layouts of Q, K, V.  Nothing of size Lq x Lk is stored; P is recomputed from lse. */
int petr_attention_f32_bwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                           const float* O, int ldo, const float* dO, int lddo, const float* lse,
                           const uint8_t* key_mask, const float* attn_bias, int bias_per_head,
                           float* dQ, int lddq, float* dK, int lddk, float* dV, int lddv,
                           float drop_p, uint64_t seed, int B, int Lq, int Lk, int heads, int head_dim,
                           float scale, void* stream);
/* d reference_points += backward of petr_pos2posemb3d. */
int petr_pos2posemb3d_bwd(const float* ref, const float* dim_t, const float* dOut, int Q, int num_feats,
                          float* dref, void* stream);
/* Backward of petr_decode_boxes given its OUTPUT: dReg [L*B*Q, code_size]; dref (nullable) += gradient through
 * inverse_sigmoid(reference_points). */
int petr_decode_boxes_bwd(const float* out, const float* dOut, const float* ref, const float* h_pc_range,
                          const float* time_div, float* dReg, float* dref, int L, int B, int Q, int code_size,
                          void* stream);
/* normalize_bbox (core/bbox/util.py:38-61): gt [n, code_in] (cx, cy, cz, w, l, h, rot[, vx, vy]) -> [n, code_in + 1]. */
int petr_normalize_bbox(const float* gt, int n, int code_in, float* out, void* stream);
/* HungarianAssigner3D cost matrix [Q, G] = w_cls * FocalLossCost + w_reg * BBox3DL1Cost over the first 8 box
 * channels, nan_to_num'ed as hungarian_assigner_3d.py:131; the assignment itself is scipy's linear_sum_assignment on
 * the host, as in the reference. */
int petr_hungarian_cost(const float* cls, const float* pred, const int64_t* gt_labels, const float* ngt,
                        int Q, int G, int C, int code, int gcode, float w_cls, float w_reg, float alpha,
                        float gamma, float eps, float* cost, void* stream);
/* Sigmoid focal loss (mmdet FocalLoss -> mmcv sigmoid_focal_loss) with per-row weights: *loss_sum += sum; dX
 * (nullable) = grad_scale * d loss / dX.  labels == C means background. */
int petr_focal_loss(const float* X, const int64_t* labels, const float* row_weight, int64_t rows, int C,
                    float alpha, float gamma, float grad_scale, float* loss_sum, float* dX, void* stream);
/* Weighted L1 over the rows whose target is finite (petr_head.py:634-640): *loss_sum += sum w |X - T|; dX (nullable). */
int petr_l1_loss(const float* X, const float* T, const float* W, int64_t rows, int code, float grad_scale,
                 float* loss_sum, float* dX, void* stream);

/* ---- Key-split exchange over NVLink peer memory (second row) without a library collective ----
 * Every rank owns one peer-mapped ("symmetric") allocation: gathered[2 parities][world][slot_floats] fp32 followed by
 * flags[2][world] uint32.  d_peer_bufs / d_peer_flags are DEVICE arrays of `world` pointers to the ranks' gathered /
 * flags bases.  Per decoder layer (seq = layer index, parity = seq & 1, seqs_per_epoch = layers per forward):
 *   petr_cross_attention_bf16(write_partial) into the rank's OWN slot gathered[parity][rank]  ->
 *   petr_p2p_push_partial: stores that slot into gathered[parity][rank] of every peer and, from the last CTA, releases
 *     flags[parity][rank] = epoch * seqs_per_epoch + seq + 1 at system scope (also on the own rank)  ->
 *   petr_attention_merge_p2p: acquires flags[parity][0..world) >= that value, then merges the nparts = world * nsplit
 *     parts like petr_attention_merge (a slot may hold the nsplit split-KV parts of its rank: slot_floats = nsplit *
 *     petr_attention_partial_floats).
 * petr_p2p_epoch_inc bumps the device-side forward counter once per forward, so a captured CUDA graph replays with
 * fresh flag values.  d_counters: uint32[world] zero-initialised scratch; d_err: set non-zero if a flag did not arrive
 * within ~2 s (the kernel then proceeds instead of hanging).  All ranks must issue the same sequence of layers. */
int petr_p2p_push_partial(const float* src, void* const* d_peer_bufs, void* const* d_peer_flags, int world, int rank,
                          int64_t slot_floats, int parity, const uint32_t* d_epoch, int seq, int seqs_per_epoch,
                          uint32_t* d_counters, void* stream);
int petr_p2p_epoch_inc(uint32_t* d_epoch, void* stream);
int petr_attention_merge_p2p(const float* gathered, int nparts, int world, const uint32_t* d_flags, const uint32_t* d_epoch,
                             int seq, int seqs_per_epoch, uint32_t* d_err, float* O, int ldo, void* O_split,
                             int B, int Lq, int heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PETR_B200_H_ */
