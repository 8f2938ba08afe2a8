/*
 * erv_b200.h -- C ABI of the B200-native attention hot path of efficient-rpe-vit.
 *
 * The reference (alemassaad/efficient-rpe-vit) is pure Python: its "FFI" for this path is the
 * plugin interface of models/attention/ and models/rpe/ (ATTENTION_REGISTRY, RPE_REGISTRY).  The
 * host-side mirror of those classes lives in efficient-rpe-vit_b200/erv_b200/ and binds the entry
 * points below with ctypes (erv_b200/_capi.py).  Each entry point names the reference code it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns an int status: 0 ok, ERV_E_* otherwise; erv_last_error() gives a
 *     thread-local message.  No exceptions, no allocation: outputs and workspaces are caller owned.
 *   - all pointers are DEVICE pointers unless the name says host; `stream` is a cudaStream_t.
 *   - `dtype` selects the element type of the activation tensors (qkv, out and their gradients):
 *     ERV_F32 or ERV_BF16.  Parameters, tables, feature tensors and statistics are always fp32,
 *     and all arithmetic accumulates in fp32.
 *   - packed layout: qkv is the output of the reference's qkv Linear, [B, N, 3, H, Dh] contiguous
 *     (favor_plus.py:174-176); q/k/v are read in place.  out is [B, N, H, Dh] == [B, N, C], i.e. the
 *     layout after the reference's transpose(1,2).reshape (favor_plus.py:263).  dqkv mirrors qkv.
 *   - re-entrant; the only global state is an atomic launch counter and per-function attributes.
 */
#ifndef ERV_B200_H
#define ERV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERV_ABI_VERSION 2  /* 2: activation dtype in the erv_block_* calls, erv_kerple_set_fft */

enum { ERV_OK = 0, ERV_E_INVALID = 1, ERV_E_UNSUPPORTED = 2, ERV_E_CUDA = 3, ERV_E_WORKSPACE = 4 };
enum { ERV_F32 = 0, ERV_BF16 = 1 };
enum { ERV_FEAT_FAVOR = 0, ERV_FEAT_RELU = 1 };                 /* favor_plus.py:112 / relu.py:116 */
enum { ERV_ROT_NONE = 0, ERV_ROT_ROPE = 1, ERV_ROT_CIRCULANT = 2 }; /* rope.py:70 / circulant_string.py:297 */
enum { ERV_PREP_SCALE = 0, ERV_PREP_L2NORM = 1, ERV_PREP_NONE = 2 }; /* favor_plus.py:187-209 */

int erv_abi_version(void);
const char* erv_last_error(void);
/* number of kernels this library has launched since the last reset (bench.py's gpu_launches) */
uint64_t erv_launch_count(void);
void erv_reset_launch_count(void);

/* ---- tables ------------------------------------------------------------------------------ */

/* RoPE cos/sin caches [num_patches, Dh/2]; replaces RoPE.__init__ (rope.py:53-68). */
int erv_rope_table(float theta, int num_patches, int head_dim, float* cos_out, float* sin_out, void* stream);

/* Circulant-STRING rotation table.  For head h, token n (n=0 is CLS -> identity) the rotation
 * of circulant_string.py:234-295 equals a circular convolution with g = Re IFFT(exp(i*theta)),
 * theta[h,n,:] = 2*sum_k pos[n-1,k]*Im FFT(coeffs[h,k,:]).  g_out is [H, N, Dh].
 * coeffs [H, coord_dim, Dh], positions [N-1, coord_dim]. */
int erv_circulant_table_fwd(const float* coeffs, const float* positions, int H, int N, int head_dim,
                            int coord_dim, float* g_out, void* stream);
/* d coeffs from d g.  dg_part is [H, slots, N, Dh] (partial sums the attention backward kernels
 * produce, one slot per CTA); they are summed here in a fixed order, in place into slot 0 (dg_part is
 * clobbered).  scratch: erv_circulant_table_bwd_scratch() bytes, 8-byte aligned. */
size_t erv_circulant_table_bwd_scratch(int H, int N, int head_dim, int coord_dim);
int erv_circulant_table_bwd(const float* coeffs, const float* positions, float* dg_part, int slots, int H,
                            int N, int head_dim, int coord_dim, float* dcoeffs, void* scratch,
                            size_t scratch_bytes, void* stream);

/* Stand-alone rotation of a [B, H, N, Dh] fp32 contiguous tensor: RoPE.apply_rotary_emb
 * (rope.py:70-137; tab_a = cos, tab_b = sin, each [N, Dh/2], already gathered at `positions`) or
 * CirculantStringRPE.apply_circulant_string (circulant_string.py:297-341; tab_a = g table).
 * inverse != 0 applies the transpose (the backward of the rotation).  dg_accum (circulant,
 * optional, [H, N, Dh], pre-zeroed) receives sum_b of the table gradient given x_raw. */
int erv_rotate(const float* x, float* y, int B, int H, int N, int head_dim, int rot,
               const float* tab_a, const float* tab_b, int inverse, void* stream);
int erv_rotate_table_grad(const float* x_raw, const float* dy, int B, int H, int N, int head_dim,
                          float* dg_accum, void* stream);

/* ---- random-feature maps ----------------------------------------------------------------- */

/* phi(x) for x [B, H, N, Dh] fp32 contiguous, omega [H, Dh, M]: FAVORPlusAttention._compute_phi_positive
 * (favor_plus.py:112-140, per-token max subtracted) or ReLUAttention._compute_relu_features
 * (relu.py:116-138).  phi_out [B, H, N, M].  workspace: erv_feature_map_workspace() bytes. */
size_t erv_feature_map_workspace(int H, int head_dim, int M);
int erv_feature_map_fwd(const float* x, const float* omega, int B, int H, int N, int head_dim, int M,
                        int kind, float* phi_out, void* workspace, size_t workspace_bytes, void* stream);
/* dx given dphi (the max is a constant, favor_plus.py:131). */
int erv_feature_map_bwd(const float* x, const float* omega, const float* dphi, int B, int H, int N,
                        int head_dim, int M, int kind, float* dx, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- attention cores --------------------------------------------------------------------- */

/* Workspace (bytes) needed by the *_fwd / *_bwd calls below for these shapes. */
size_t erv_linear_attention_workspace(int B, int N, int H, int head_dim, int M, int rot, int backward);
size_t erv_kerple_attention_workspace(int B, int N, int H, int head_dim, int M, int backward);
size_t erv_softmax_attention_workspace(int B, int N, int H, int head_dim, int rot, int backward);
/* number of circulant partial-gradient slots the backward kernels write (see circulant_table_bwd) */
int erv_circulant_slots(int B, int H);

/* FAVOR+/ReLU linear attention without KERPLE (favor_plus.py:179-260 else-branch, relu.py:177-258):
 *   x = rot(q|k) * Dh^-1/4 ; phi ; out = phi(q) (phi(k)^T v) / (phi(q) sum_n phi(k) + 1e-6).
 * omega [H, Dh, M].  rot tables as in erv_rotate (rope tables cover positions 0..N-1). */
int erv_linear_attention_fwd(const void* qkv, void* out, const float* omega, int B, int N, int H,
                             int head_dim, int M, int kind, int rot, const float* tab_a,
                             const float* tab_b, int dtype, float* kv_state, void* workspace,
                             size_t workspace_bytes, void* stream);
/* Floats of the optional kv_state buffer: the finished [phi(k)^T v | sum_n phi(k)] of every (batch, head)
 * pair, which the forward writes and the backward reads instead of rebuilding it (the reference's autograd
 * saves the same tensors, favor_plus.py:250-257).  0 = these shapes do not use it; pass NULL then.  Passing
 * NULL to both calls is always valid (the backward recomputes). */
size_t erv_linear_attention_state_floats(int B, int N, int H, int head_dim, int M);
/* Backward (replaces autograd over the ops above).  `out` is the saved forward
 * output.  dg_part ([H, erv_circulant_slots, N, Dh], circulant only) is overwritten. */
int erv_linear_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv,
                             const float* omega, int B, int N, int H, int head_dim, int M, int kind,
                             int rot, const float* tab_a, const float* tab_b, float* dg_part, int dtype,
                             const float* kv_state, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the block around the attention core, dim 32 / MLP width 64 only -------- */

/* 1 when the fused block kernels below cover these dims (the reference's: configs/datasets/mnist.py:20-24). */
int erv_block_supported(int dim, int mlp_dim);
/* Kernel family behind the four block calls: 1 = tcgen05 tiles (default), 0 = fp32 FFMA2 register tiles,
 * -1 = default / ERV_DISABLE_BLOCK_TC environment variable.  Both compute the same functions. */
void erv_block_set_tensor_core(int mode);
/* act_dtype (ERV_F32 / ERV_BF16) in the four calls below is the element type of the activations exchanged with the
 * attention core: qkv, dqkv, attn_out, d_attn_out.  ERV_BF16 is what a bf16-autocast nn.Linear hands the core
 * (favor_plus.py:174 under torch.autocast); the kernels convert in registers, so no cast kernels run around the core.
 * bf16 needs the tcgen05 family (erv_block_act_bf16_supported() == 1); x, y, dx, dy and all parameters stay fp32. */
int erv_block_act_bf16_supported(void);
/* qkv [rows, 3*dim] = LayerNorm(x; ln_w, ln_b, eps) w_qkv^T (+ b_qkv, may be NULL): norm1 + attention.qkv
 * (unified_transformer.py:75-83, favor_plus.py:174).  x [rows, dim] fp32. */
int erv_block_ln_qkv_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w_qkv,
                         const float* b_qkv, void* qkv, int act_dtype, int rows, int dim, float eps, void* stream);
/* Backward: dx = dres (may be NULL) + d/dx ; dparams [erv_block_ln_qkv_params()] = dW_qkv [3 dim, dim] |
 * db_qkv [3 dim] | dln_w [dim] | dln_b [dim].  Recomputes the LayerNorm from x.  With grad_accum (4 pointers
 * in that order, entries may be NULL) the parameter gradients are ADDED to those buffers instead and dparams
 * may be NULL (fused accumulation into .grad). */
int erv_block_ln_qkv_params(void);
size_t erv_block_ln_qkv_bwd_workspace(int rows);
int erv_block_ln_qkv_bwd(const float* x, const void* dqkv, int act_dtype, const float* dres, const float* ln_w,
                         const float* ln_b, const float* w_qkv, float* dx, float* dparams,
                         float* const* grad_accum, int rows, int dim, float eps, void* workspace,
                         size_t workspace_bytes, void* stream);
/* y = x1 + drop(fc2(drop(gelu(fc1(LayerNorm(x1)))))),  x1 = x + drop(attn_out w_proj^T + b_proj):
 * attention.proj + proj_dropout + residual + norm2 + mlp + residual (favor_plus.py:263-265,
 * unified_transformer.py:85-88).  params = {w_proj, b_proj, ln_w, ln_b, w_fc1, b_fc1, w_fc2, b_fc2}
 * (nn.Linear layouts).  Dropout masks are a counter hash of (*seed, salt, element); seed is a device
 * pointer (CUDA-graph friendly) and may be NULL when p_drop == 0. */
int erv_block_mlp_fwd(const void* attn_out, int act_dtype, const float* x, const float* const* params, float* y,
                      int rows, int dim, int mlp_dim, float eps, float p_drop, const long long* seed, int salt,
                      void* stream);
/* Backward: recomputes the forward from (attn_out, x, seed).  d_attn_out, dx1 [rows, dim] (dx1 is the
 * gradient of the residual stream, i.e. of x); dparams [erv_block_mlp_params()] = dW_proj | db_proj |
 * dln_w | dln_b | dW_fc1 | db_fc1 | dW_fc2 | db_fc2; grad_accum (8 pointers, same order) as above. */
int erv_block_mlp_params(void);
size_t erv_block_mlp_bwd_workspace(int rows);
int erv_block_mlp_bwd(const void* attn_out, int act_dtype, const float* x, const float* dy,
                      const float* const* params, void* d_attn_out, float* dx1, float* dparams,
                      float* const* grad_accum, int rows, int dim, int mlp_dim, float eps, float p_drop,
                      const long long* seed, int salt, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the two ends of the ViT, model dim 32 ------------------------------------ */

/* x [B, N, 32] = patch embedding + CLS + position embedding (base_vit.py:190-223): x[b,0] = cls + pos[0];
 * x[b,1+p] = patch(b,p) w^T + bias + pos[1+p], patches in raster order, patch element k = c P^2 + i P + j.
 * images [B, Cin, S, S] fp32, w [32, Cin P^2], N = (S/P)^2 + 1. */
int erv_embed_supported(int dim, int patch_dim);
int erv_embed_fwd(const float* images, const float* w, const float* b, const float* cls, const float* pos,
                  float* out, int B, int Cin, int S, int P, void* stream);
size_t erv_embed_bwd_workspace(int B, int Cin, int S, int P);
/* Gradients of the four parameters from dout [B, N, 32]; accumulate != 0 adds to the buffers' contents. */
int erv_embed_bwd(const float* images, const float* dout, float* dw, float* db, float* dcls, float* dpos,
                  int accumulate, int B, int Cin, int S, int P, void* workspace, size_t workspace_bytes,
                  void* stream);
/* loss = mean_b CrossEntropy(Linear(LayerNorm(x[b, 0])), labels[b]) (base_vit.py:230-233 + the training loop's
 * criterion, training.py:57-60).  x [B, N, dim], w [K, dim], labels int64 [B], loss: one float on the device. */
size_t erv_head_loss_workspace(int B, int K);
int erv_head_loss_fwd(const float* x, const float* ln_w, const float* ln_b, const float* w, const float* b,
                      const long long* labels, float* loss, int B, int N, int dim, int K, float eps, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Backward: dx [B, N, dim] is fully written (zero except token 0); dparams = dW [K, dim] | db [K] | dln_w | dln_b,
 * or added to the 4 grad_accum buffers.  dloss: one float on the device. */
int erv_head_loss_bwd(const float* x, const float* ln_w, const float* ln_b, const float* w, const float* b,
                      const long long* labels, const float* dloss, float* dx, float* dparams,
                      float* const* grad_accum, int B, int N, int dim, int K, float eps, void* workspace,
                      size_t workspace_bytes, void* stream);

/* KERPLE linear attention (favor_plus.py:197-245 + kerple.py:99-344 + fft_utils.py:112-172), evaluated
 * as Toeplitz-masked attention: A = (phi(q) phi(k)^T) * exp(bias[j-i+N-1]); out = A v / (A 1 + 1e-6)
 * with q, k L2-normalised.  rel_pos_bias [H, 2N-1].  den_out [B, H, N] is saved for the backward. */
int erv_kerple_attention_fwd(const void* qkv, void* out, float* den_out, const float* omega,
                             const float* rel_pos_bias, int B, int N, int H, int head_dim, int M,
                             int kind, int dtype, void* workspace, size_t workspace_bytes, void* stream);
/* Route of erv_kerple_attention_fwd: 1 = the reference's FFT route (kerple.py:252-270 -> fft_utils.py:142-170: Toeplitz
 * product by FFT along the patch axis, here a shared-memory radix-16 transform of 8192 points fused with the phi(q)
 * read-out) whenever N - 1 <= 4096, 0 = Toeplitz-masked tiles always, -1 = default (FFT for N - 1 > 2048 and (M > 64 or B*H >= 16),
 * the measured crossover, or the ERV_KERPLE_FFT environment variable).  Both routes compute the same function; the workspace query follows the mode. */
void erv_kerple_set_fft(int mode);
int erv_kerple_attention_bwd(const void* qkv, const void* out, const float* den, const void* dout,
                             void* dqkv, float* dbias, const float* omega, const float* rel_pos_bias,
                             int B, int N, int H, int head_dim, int M, int kind, int dtype,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Softmax attention (softmax.py:86-115): out = dropout(softmax(rot(q) rot(k)^T / sqrt(Dh) + mask)) v.
 * mask: optional uint8 [B, N, N] (0 = masked, softmax.py:104-108), may be NULL.
 * dropout_p in [0,1): attention-probability dropout (softmax.py:112) from a counter-based generator
 * keyed by (seed ^ *seed_dev, b, h, i, j) so the backward regenerates it; seed_dev (optional, may be NULL) is a
 * device-resident 64-bit value, so a captured CUDA graph draws new masks on every replay.
 * lse_out [B, H, N] (log-sum-exp, saved for backward).  attn_out: optional [B, H, N, N] fp32 dump of the
 * (post-dropout) probabilities for return_attention=True (softmax.py:122-123), may be NULL. */
int erv_softmax_attention_fwd(const void* qkv, void* out, float* lse_out, float* attn_out,
                              const uint8_t* mask, int B, int N, int H, int head_dim, int rot,
                              const float* tab_a, const float* tab_b, float dropout_p, uint64_t seed,
                              const long long* seed_dev, int dtype, void* workspace, size_t workspace_bytes,
                              void* stream);
int erv_softmax_attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout,
                              void* dqkv, const uint8_t* mask, int B, int N, int H, int head_dim, int rot,
                              const float* tab_a, const float* tab_b, float* dg_part, float dropout_p,
                              uint64_t seed, const long long* seed_dev, int dtype, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- Toeplitz product -------------------------------------------------------------------- */

/* y[p] = T(c[p % c_count]) x[p], T[i,j] = c[j-i+n-1]; replaces fft_toeplitz_matmul
 * (fft_utils.py:17-172).  x, y: [P, n, d] fp32; c: [c_count, 2n-1] with P % c_count == 0 and batch p
 * using coefficient row p % c_count when c_per_batch == 0 (shared/per-head rows), or row p when
 * c_count == P. */
int erv_toeplitz_matmul_fwd(const float* c, const float* x, float* y, int P, int c_count, int n, int d,
                            void* stream);
/* dx = T^T dy ; dc[r, delta] = sum_{p: row r} sum_{j-i=delta} dy[p,i,:] . x[p,j,:] */
int erv_toeplitz_matmul_bwd(const float* c, const float* x, const float* dy, float* dx, float* dc,
                            int P, int c_count, int n, int d, void* stream);

/* ---- training-loop helpers ------------------------------------ */

/* Fused Adam/AdamW over one flat fp32 parameter buffer (torch.optim.Adam semantics,
 * experiments/utils/training.py:304-309).  grad_scale multiplies the gradient first (1/world_size after
 * the data-parallel allreduce).  step is the 1-based step count, read from the DEVICE pointer step_dev
 * (int64) if non-NULL -- so a captured CUDA graph can advance it -- else from the argument. */
int erv_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int decoupled_wd,
                  float grad_scale, int64_t step, const int64_t* step_dev, void* stream);
/* The same step with the hyper-parameters read from device memory at run time: hyper = [lr, beta1, beta2, eps,
 * weight_decay] (5 floats) and the step counter *step_dev.  A CUDA graph that recorded this call follows a learning-rate
 * schedule (experiments/train.py:216-286 create_lr_scheduler) by rewriting hyper[0] between replays. */
int erv_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, const float* hyper,
                      int decoupled_wd, float grad_scale, const int64_t* step_dev, void* stream);

/* ---- the rest of the block, for the reference's small model dims ------------- */

/* Weight and bias gradient of a Linear layer (base_vit.py:85, unified_transformer.py:53-59, favor_plus.py:59-62):
 * dw[O, I] = sum_r dy[r, O] x[r, I], db[O] = sum_r dy[r, O] (db may be NULL), reduction split over the R token rows.
 * dy, x: fp32 or bf16 (dtype), contiguous; dw, db: fp32.  erv_linear_wgrad_supported() tells whether the shape is
 * handled (small O*I, many rows); other shapes belong to the library GEMM. */
int erv_linear_wgrad_supported(int R, int O, int I);
size_t erv_linear_wgrad_workspace(int R, int O, int I);
int erv_linear_wgrad(const void* dy, const void* x, float* dw, float* db, int R, int O, int I, int dtype,
                     void* workspace, size_t workspace_bytes, void* stream);

/* LayerNorm over the last dimension C of x [R, C] fp32 (unified_transformer.py:61-62, base_vit.py:105), biased variance,
 * one warp per row.  mean / rstd [R] are saved for the backward. */
int erv_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                      int R, int C, float eps, void* stream);
size_t erv_layernorm_bwd_workspace(int R, int C);
int erv_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      float* dx, float* dgamma, float* dbeta, int R, int C, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---- data-parallel gradient reduction --------------------------------- */

/* One-shot sum all-reduce over peer memory: peer_bufs (HOST array of `world` device pointers) are every rank's symmetric
 * buffer as mapped in this process (torch.distributed._symmetric_memory); each holds n floats of data and, at
 * flag_offset_floats, erv_allreduce_flag_floats() flag words that start at zero.  On return (stream order) the local buffer
 * holds the sum over ranks, added in rank order (bit-identical on every rank); scratch is n floats of local memory.
 * epoch_dev: one device uint32 per rank, starts at zero, owned by this call.  Every rank must make the same sequence of
 * calls.  Replaces the NCCL all-reduce of the flat gradient (the reference has no multi-GPU path). */
int erv_allreduce_flag_floats(void);
int erv_allreduce_oneshot(const void* const* peer_bufs, size_t n, size_t flag_offset_floats, float* scratch, int rank,
                          int world, uint32_t* epoch_dev, void* stream);

/* ---- diagnostics -------------------------------------------------------------------------- */

/* D[128, N] = A[128, K] * B[N, K]^T on tcgen05 (TF32 or BF16 operands, fp32 accumulate in TMEM) using the shared-
 * memory operand layouts of the fused kernels; *_mn_major selects MN-major instead of K-major.  Used by the GPU tests
 * to pin the descriptor encodings. */
int erv_debug_umma_gemm(const float* A, const float* B, float* D, int N, int K, int a_mn_major, int b_mn_major,
                        int bf16, void* stream);
/* SM clocks for `iters` back-to-back M=128 MMAs of width N issued by one thread (cycles: device int64). */
/* Optional phase trace of the tensor-core backward: CTA 0 writes (tag, clock64) pairs into this device buffer of at
 * least 2000 int64 (NULL disables). */
void erv_debug_set_trace(long long* device_buffer);
int erv_debug_umma_timing(int N, int bf16, int a_mn_major, int b_mn_major, int iters, long long* cycles,
                          void* stream);
/* The same product with the A operand in tensor memory (written with tcgen05.st, lane = row): pins the TMEM operand
 * layout of the kernels that keep their feature tiles out of shared memory. */
int erv_debug_umma_gemm_ts(const float* A, const float* B, float* D, int N, int K, int b_mn_major, int bf16,
                           void* stream);
/* MMA cost with `nacc` independent accumulators, M = 64 or 128, A from shared (0) or tensor (1) memory.
 * elected: 1 = issued by the elected lane of a converged warp (elect.sync), 0 = by a divergent `tid == 0`.
 * cycles[0] = first issue -> completion, cycles[1] = issue loop alone (device int64[2]). */
int erv_debug_umma_timing2(int N, int M, int bf16, int a_tmem, int b_mn_major, int nacc, int iters, int elected,
                           long long* cycles, void* stream);
/* One-SM throughput probes: mode 0 tcgen05.ld, 1 tcgen05.st, 2 cvt.rn.bf16x2.f32, 3 ex2.approx, 4 both, 5 FFMA.
 * cycles[0] = SM clocks for `iters` instructions per thread; sink: >= 512 floats or NULL. */
int erv_debug_unit_probe(int mode, int threads, int iters, long long* cycles, float* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ERV_B200_H */
