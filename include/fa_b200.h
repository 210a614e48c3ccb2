/* fa_b200.h -- C ABI of libfa_b200.so: B200 (sm_100a) flash-attention forward + backward.
 *
 * This is the drop-in boundary for the flash-attention path of pdrebello/flash-attention-minitorch.
 * Two surfaces:
 *
 *  1. The reference's own host-pointer entry points, bit-for-bit the same signatures its ctypes
 *     loader binds (reference minitorch/cuda_kernel_ops.py:549-576 and :613-647):
 *         launch_flash_attn_fw   replaces  src/flash_attn_fw.cu:302-359  (FA-1: l = sum exp, m = row max)
 *                                and       src/flash_attn2_fw.cu:312-372 (FA-2: l = logsumexp, m untouched)
 *         launch_flash_attn_bw   replaces  src/flash_attn_bw.cu:277-365  and src/flash_attn2_bw.cu:279-369
 *     They are exported -- under exactly those names -- by the six shim libraries that
 *     compile_cuda.sh writes to minitorch/cuda_kernels/ (flash_attn_fw.so, flash_attn_bw.so,
 *     flash_attn_causal_fw.so, flash_attn_causal_bw.so, flash_attn2_fw.so, flash_attn2_bw.so; reference
 *     loader: minitorch/cuda_kernel_ops.py:27-35).  Each shim forwards to fa_b200_host_fwd /
 *     fa_b200_host_bwd below with the (l, m) convention of the library it replaces.
 *
 *  2. A device-resident API (fa_b200_fwd / fa_b200_bwd) that takes device pointers, supports bf16
 *     storage, and returns error codes.  The reference has no equivalent: every one of its calls
 *     stages through host memory.  This is the surface the throughput numbers are measured on.
 *
 * All tensors are dense row-major (batch_heads, T, d); batch_heads = B*H flattened exactly as the
 * reference does (cuda_kernel_ops.py:542-547).  Plain pointers and sizes only; no framework types.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define FA_OK 0
#define FA_ERR_BAD_ARG 1      /* null pointer, non-positive size, unsupported d/dtype/precision */
#define FA_ERR_CUDA 2         /* a CUDA runtime call or kernel launch failed; see fa_b200_last_error() */
#define FA_ERR_WORKSPACE 3    /* workspace pointer null or too small */
#define FA_ERR_UNSUPPORTED 4  /* shape not covered by the requested precision */

/* ---- storage dtypes of q/k/v/o/dO/dq/dk/dv in the device API ---- */
#define FA_DTYPE_F32 0
#define FA_DTYPE_BF16 1

/* ---- arithmetic used for the two contractions ----
 * AUTO      : bf16 storage -> BF16_TC; fp32 storage -> TF32_TC when the shape is covered (forward and backward:
 *             d in {32,64}; the backward without dropout), else FP32_SIMT.  The host-pointer API (fp32 storage) resolves AUTO the
 *             same way: tf32 tensor cores wherever a tcgen05 kernel exists, which meets the 1e-3 max-abs bar of the
 *             fp32 path; FA_B200_HOST_PRECISION=fp32 / fa_b200_set_host_precision(FA_PREC_FP32_SIMT) keeps exact fp32.
 * FP32_SIMT : fp32 FMA on CUDA cores (any d <= 128 -- the reference's own limit, src/flash_attn_fw.cu:43 -- and any T); exact-fp32 cross-check path
 * TF32_TC   : tcgen05.mma kind::tf32 on fp32 storage read directly by TMA (no conversion pass; the backward rounds its
 *             operands to tf32 in its prep pass); shapes it does not cover run FP32_SIMT
 * BF16_TC   : tcgen05.mma kind::f16 (bf16 inputs, fp32 accumulate); fp32 storage is converted on device */
#define FA_PREC_AUTO 0
#define FA_PREC_FP32_SIMT 1
#define FA_PREC_TF32_TC 2
#define FA_PREC_BF16_TC 3

/* ---- (l, m) conventions of the reference's two kernel families ---- */
#define FA_CONV_FA1 1 /* l = sum_j exp(s_j - m), m = max_j s_j           (src/flash_attn_fw.cu:262-284) */
#define FA_CONV_FA2 2 /* l = m + log(sum_j exp(s_j - m)), m left as is   (src/flash_attn2_fw.cu:279-294) */

int fa_b200_version(void);
/* Thread-local, NUL-terminated description of the last failure on this thread ("" if none). */
const char* fa_b200_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all entry points). */
unsigned long long fa_b200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Device-resident API.  All pointers are device pointers valid on the current device; `stream`
 * is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous w.r.t. the host.
 *
 * fa_b200_fwd:  O = softmax(tau * Q K^T [+ causal mask]) V,  tau = 1/sqrt(d)
 *   q,k,v : [batch_heads, T, d] of `dtype`            o : same shape/dtype (output)
 *   lse   : [batch_heads, T] fp32 logsumexp (natural log) of the scaled scores -- the FA-2 "L" (may be NULL)
 *   m_out, l_out : optional FA-1 statistics (row max, sum exp); NULL to skip
 * ------------------------------------------------------------------------------------------ */
int fa_b200_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* m_out, float* l_out,
                int dtype, int precision, int batch_heads, int T, int d, int causal, void* stream);

/* Scratch bytes fa_b200_bwd needs for this problem (0 is a valid answer). */
size_t fa_b200_bwd_workspace_bytes(int dtype, int precision, int batch_heads, int T, int d);

/* fa_b200_bwd: dQ, dK, dV of the above given dO, recomputing P from `lse`.
 *   accumulate_dq != 0 : dq += result (reference semantics, src/flash_attn2_bw.cu:228); else dq = result.
 *   dk, dv are always overwritten. */
int fa_b200_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_out, const float* lse,
                void* dq, void* dk, void* dv, int dtype, int precision, int batch_heads, int T, int d, int causal,
                int accumulate_dq, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention dropout fused into the kernels (device-resident API).  The reference applies no dropout on its
 * flash path; the convention is that of its Dropout module (minitorch/modules_basic.py:77-80): keep ~
 * Bernoulli(1 - p), kept probabilities scaled by 1/(1 - p), applied to P after the softmax normalisation
 * (lse is the logsumexp of the undropped scores).  The mask is a pure function of (seed, head, query, key)
 * regenerated inside the backward -- nothing is stored.  p is quantised to 1/128: p_eff = round(128 p)/128.
 *   fa_b200_fwd_dropout / fa_b200_bwd_dropout : as fa_b200_fwd / fa_b200_bwd with O = (P o Z) V
 *   fa_b200_dropout_mask      : keep[n_heads][T][T] (1 = kept) for heads [head_begin, head_begin + n_heads) of the
 *                               flattened batch*heads axis, exactly as the kernels draw it (debug / oracle)
 *   fa_b200_dropout_keep_scale: the 1/(1 - p_eff) the kernels apply;  fa_b200_dropout_effective_p: p_eff
 * ------------------------------------------------------------------------------------------ */
int fa_b200_fwd_dropout(const void* q, const void* k, const void* v, void* o, float* lse, int dtype, int precision,
                        int batch_heads, int T, int d, int causal, float p_drop, unsigned long long seed,
                        void* stream);
int fa_b200_bwd_dropout(const void* q, const void* k, const void* v, const void* o, const void* d_out, const float* lse,
                        void* dq, void* dk, void* dv, int dtype, int precision, int batch_heads, int T, int d,
                        int causal, int accumulate_dq, float p_drop, unsigned long long seed, void* workspace,
                        size_t workspace_bytes, void* stream);
int fa_b200_dropout_mask(unsigned char* keep, int head_begin, int n_heads, int T, float p_drop,
                         unsigned long long seed, void* stream);
float fa_b200_dropout_keep_scale(float p_drop);
float fa_b200_dropout_effective_p(float p_drop);

/* ------------------------------------------------------------------------------------------
 * Host-pointer API: the reference ABI with an explicit convention and an error code.
 * All pointers are HOST fp32 buffers (pageable or pinned).  The call stages through a process-global
 * device arena, runs on `stream`, and returns after the outputs have landed in host memory.
 * ------------------------------------------------------------------------------------------ */
int fa_b200_host_fwd(const float* q, const float* k, const float* v, float* out, float* l, float* m, int batch,
                     int N, int d, int causal, int convention, void* stream);
int fa_b200_host_bwd(const float* q, const float* k, const float* v, const float* out, const float* out_grad,
                     float* q_grad, float* k_grad, float* v_grad, const float* l, const float* m, int batch, int N,
                     int d, int causal, int convention, void* stream);
/* Arithmetic used by the host-pointer API (FA_PREC_*).  Default FA_PREC_AUTO; also settable with the
 * environment variable FA_B200_HOST_PRECISION = auto | fp32 | tf32 | bf16 read at first use. */
void fa_b200_set_host_precision(int precision);
int fa_b200_get_host_precision(void);
/* on = 1: fa_b200_host_bwd may assume that q_grad holds zeros on entry -- what every caller of the reference passes
 * (minitorch/cuda_kernel_ops.py:609-611) -- and skips its upload (11 % of the H2D bytes); q_grad is then overwritten.
 * Default 0 (the reference kernels accumulate, src/flash_attn2_bw.cu:228); environment: FA_B200_HOST_ZERO_QGRAD=1. */
void fa_b200_set_host_zero_qgrad(int on);
/* on = 1: the bf16 forward at d = 128 runs the CTA-pair kernel (tcgen05.mma.cta_group::2, each CTA stages half of every
 * K/V tile, P through shared memory) instead of the single-CTA kernel.  Same results (parity-tested); measured slower on
 * B200, hence off by default (environment: FA_B200_FWD_PAIR=1).  Kept as a tuning knob and as the record of that design. */
void fa_b200_set_fwd_pair(int on);
int fa_b200_get_fwd_pair(void);
/* The bf16 tensor-core backward (d = 64, 128) is a PERSISTENT kernel: one CTA per SM draws (head, K/V tile) work items from
 * a counter and keeps its TMA / MMA / softmax-gradient pipelines running across items.  on = 0 launches one CTA per work item
 * instead (the round-1 schedule; same arithmetic, dK / dV bit-identical, dQ equal up to the order of its fp32 reduce-adds);
 * measured 4-6 % slower at config 3.  Default 1 (environment: FA_B200_BWD_PERSISTENT=0).  A tuning / test knob. */
void fa_b200_set_bwd_persistent(int on);
int fa_b200_get_bwd_persistent(void);
/* The same for the tensor-core forward (bf16 d = 64, 128; tf32 d = 32, 64): persistent by default, work item = one 256-row
 * query block of one head, the work-item counter lives in a library-owned slot that the kernel leaves zeroed.  on = 0
 * launches one CTA per block (bit-identical results: no cross-CTA arithmetic).  Environment: FA_B200_FWD_PERSISTENT=0. */
void fa_b200_set_fwd_persistent(int on);
int fa_b200_get_fwd_persistent(void);
/* Release the device arena held by the host-pointer API. */
void fa_b200_release_arena(void);

/* ------------------------------------------------------------------------------------------
 * Operators of the decoder-only LM around the attention path ("next" rows), device pointers,
 * fp32.  They replace, for the model path of project/run_machine_translation.py, the reference's host-staged
 * kernels of src/combine.cu (MatrixMultiplyKernel :150-253, reduceKernel :315-387).
 *   fa_b200_gemm_tf32 : C[M x N] (+)= A[M x K] B[K x N] (+ bias[N]) on tcgen05 kind::tf32, fp32 accumulate.
 *       a_mn_major = 0: A stored [M][K];  1: A stored [K][M].   b_mn_major = 0: B stored [N][K];  1: B stored [K][N].
 *       Linear (W is (in, out)):  Y = X W: (0, 1);  dX = dY W^T: (0, 0) with B = W;  dW = X^T dY: (1, 1).
 *       M, N, K multiples of 4.
 *   fa_b200_colsum_f32: out[N] (+)= column sums of x[M][N]  (bias gradients)
 * ------------------------------------------------------------------------------------------ */
int fa_b200_gemm_tf32(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, int a_mn_major,
                      int b_mn_major, int accumulate, void* stream);
int fa_b200_colsum_f32(const float* x, float* out, int M, int N, int accumulate, void* stream);
/* The same GEMM with "head layout" operands: a logical [B*T][H*d] matrix stored as (B, H, T, d), the layout of the
 * attention kernels.  The TMA tensor maps do the (B,T,H,d) <-> (B,H,T,d) permutation, so the q/k/v projections write
 * straight into attention layout and the output projection / the projection gradients read straight out of it --
 * the three .contiguous() copies per layer of minitorch/modules_transfomer.py:79-89,115-117 disappear.
 * a_heads / b_heads / c_heads flag which operands use it (b_heads needs b_mn_major); T % 128 == 0, d % 32 == 0.
 * fa_b200_colsum_heads_f32: out[h*d + dd] (+)= sum_{b,t} x[b][h][t][dd]. */
int fa_b200_gemm_tf32_heads(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, int a_mn_major,
                            int b_mn_major, int accumulate, int a_heads, int b_heads, int c_heads, int T, int H, int d,
                            void* stream);
int fa_b200_colsum_heads_f32(const float* x, float* out, int B, int H, int T, int d, int accumulate, void* stream);
/* A head-layout operand of fa_b200_gemm_tf32_heads may also be `parts` such tensors back to back, (parts, B, H, T, d) --
 * q | k | v of one fused projection: its logical width is then parts * H * d (N = 3 H d for the fused forward / weight
 * gradient, K = 3 H d for the fused input gradient).  fa_b200_colsum_heads_parts_f32: out[part][h*d + dd] (+)= sum_{b,t}. */
int fa_b200_colsum_heads_parts_f32(const float* x, float* out, int parts, int B, int H, int T, int d, int accumulate, void* stream);
/* One fused, device-resident pass per operator of the model path (reference formulas in parentheses); fp32, row-major,
 * inner dimensions and element counts multiples of 4.
 *   layernorm : y = w (x - mean) / sqrt(var + eps) + b, biased variance (minitorch/modules_basic.py:166-175);
 *               fwd saves mean / rstd per row; bwd overwrites dx, dw, db
 *   gelu      : tanh approximation (minitorch/nn.py:208-214)
 *   dropout   : y = x * keep / (1 - p_eff), keep = f(seed, element index) -- call it again on dy for the backward
 *               (minitorch/modules_basic.py:77-80; p quantised as in fa_b200_fwd_dropout).  seed_offset_dev (optional,
 *               device memory) is a step counter mixed into the seed at run time: a captured CUDA graph then draws
 *               a fresh mask on every replay
 *   embedding : y[i] = w[idx[i]]  (one_hot @ weights, minitorch/modules_basic.py:46-53); bwd overwrites dw[vocab][dim]
 *   softmax_loss : loss[i] = logsumexp(logits[i]) - logits[i][target[i]] (minitorch/nn.py:256-274); lse saved;
 *               bwd: dlogits[i][j] = gscale[i] (softmax(logits[i])[j] - [j == target[i]])
 *   adam_step : m = b1 m + (1-b1) g; v = b2 v + v_coef g^2; p -= step_size m / (sqrt(v) + eps) over one flat buffer
 *               (minitorch/optim.py:44-76, whose v_coef is 1 - beta1; step_size = lr sqrt(1-b2^t)/(1-b1^t)) */
int fa_b200_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, int rows,
                          int dim, float eps, void* stream);
int fa_b200_layernorm_bwd(const float* x, const float* w, const float* dy, const float* mean, const float* rstd, float* dx,
                          float* dw, float* db, int rows, int dim, void* stream);
/* the same with dw / db ACCUMULATED into (a parameter's gradient buffer) instead of overwritten */
int fa_b200_layernorm_bwd_accum(const float* x, const float* w, const float* dy, const float* mean, const float* rstd, float* dx,
                                float* dw, float* db, int rows, int dim, void* stream);
/* The residual stream fused into LayerNorm (TransformerLayer.forward, minitorch/modules_transfomer.py: x + attn(ln_1(x)),
 * x + ff(ln_2(x))): forward sum = a + b (b, sum optional) and y = LN(sum) in one pass; backward dx = LN'(dy) + res
 * (res optional: the gradient arriving at the same tensor along the residual branch), dw / db overwritten or
 * accumulated.  dcol (optional) (+)= the column sums of dx -- the bias gradient of the linear layer that produced b
 * (out_projection / ff.linear_out), whose dy is this dx.  b / res / dcol need dim in {128, 256, 512}
 * (fa_b200_layernorm_fused_width), where the row stays in registers and the backward is a single launch. */
int fa_b200_add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, float* sum, float* y,
                              float* mean, float* rstd, int rows, int dim, float eps, void* stream);
int fa_b200_layernorm_bwd_res(const float* x, const float* w, const float* dy, const float* mean, const float* rstd,
                              const float* res, float* dx, float* dw, float* db, float* dcol, int rows, int dim, int accumulate,
                              void* stream);
/* FeedForward's gelu -> dropout pair (minitorch/modules_transfomer.py FeedForward.forward) as one pass each way:
 * y = drop(gelu(x));  dx = drop(dy) gelu'(x), with exactly the mask of fa_b200_dropout_apply(p_drop, seed, seed_offset_dev). */
int fa_b200_gelu_dropout_fwd(const float* x, float* y, size_t n, float p_drop, unsigned long long seed,
                             const unsigned long long* seed_offset_dev, void* stream);
int fa_b200_gelu_dropout_bwd(const float* x, const float* dy, float* dx, size_t n, float p_drop, unsigned long long seed,
                             const unsigned long long* seed_offset_dev, void* stream);
/* gelu_dropout_bwd for x (rows, dim), dim in {128, 256, 512}, that also reduces dcol[c] (+)= sum_rows dx[.][c]: the bias
 * gradient of the linear layer that produced x (ff.linear_in).  p_drop = 0: plain gelu backward. */
int fa_b200_gelu_dropout_bwd_colsum(const float* x, const float* dy, float* dx, float* dcol, int rows, int dim, float p_drop,
                                    unsigned long long seed, const unsigned long long* seed_offset_dev, int accumulate, void* stream);
int fa_b200_layernorm_fused_width(int dim);
int fa_b200_gelu_fwd(const float* x, float* y, size_t n, void* stream);
int fa_b200_gelu_bwd(const float* x, const float* dy, float* dx, size_t n, void* stream);
int fa_b200_dropout_apply(const float* x, float* y, size_t n, float p_drop, unsigned long long seed,
                          const unsigned long long* seed_offset_dev, void* stream);
int fa_b200_embedding_fwd(const long long* idx, const float* w, float* y, size_t n_idx, int dim, void* stream);
int fa_b200_embedding_bwd(const long long* idx, const float* dy, float* dw, size_t n_idx, int dim, size_t vocab, void* stream);
int fa_b200_embedding_bwd_accum(const long long* idx, const float* dy, float* dw, size_t n_idx, int dim, size_t vocab, void* stream);
/* The model's input stage (DecoderLM.forward, minitorch/modules_transfomer.py: token_embeddings(idx) +
 * position_embeddings(arange(T)), then dropout) as one pass each way: y[r] = drop(tok_w[idx[r]] + pos_w[r % T]) for the
 * n_idx = B*T row-major indices; bwd scatter-adds drop(dy[r]) onto dtok[idx[r]] and dpos[r % T] (both tables zeroed first
 * unless accumulate).  Mask = fa_b200_dropout_apply's on the flattened (n_idx, dim) tensor; p_drop = 0: no dropout. */
int fa_b200_embed_pos_dropout_fwd(const long long* idx, const float* tok_w, const float* pos_w, float* y, size_t n_idx, int T, int dim,
                                  float p_drop, unsigned long long seed, const unsigned long long* seed_offset_dev, void* stream);
int fa_b200_embed_pos_dropout_bwd(const long long* idx, const float* dy, float* dtok, float* dpos, size_t n_idx, int T, int dim,
                                  size_t tok_rows, size_t pos_rows, float p_drop, unsigned long long seed,
                                  const unsigned long long* seed_offset_dev, int accumulate, void* stream);
int fa_b200_softmax_loss_fwd(const float* logits, const long long* target, float* loss, float* lse, int rows, int n_classes,
                             void* stream);
int fa_b200_softmax_loss_bwd(const float* logits, const long long* target, const float* lse, const float* gscale,
                             float* dlogits, int rows, int n_classes, void* stream);
/* softmax_loss_bwd that also reduces dcol[j] (+)= sum_i dlogits[i][j] -- the bias gradient of the linear layer that
 * produced the logits (lm_head), in the same pass; n_classes <= 49152 (fa_b200_softmax_loss_colsum_width). */
int fa_b200_softmax_loss_bwd_colsum(const float* logits, const long long* target, const float* lse, const float* gscale,
                                    float* dlogits, float* dcol, int rows, int n_classes, int accumulate, void* stream);
int fa_b200_softmax_loss_colsum_width(int n_classes);
int fa_b200_adam_step(float* p, const float* g, float* m, float* v, size_t n, float step_size, float beta1, float beta2,
                      float v_coef, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Gradient all-reduce of the data-parallel machine-translation run (BASELINE.json configs[4]; the reference itself has
 * no distributed backend, its training loop is single-process: project/run_machine_translation.py:196-221).  One
 * process per GPU; NCCL over NVLink 5 / NVSwitch, resolved with dlopen at first use (libfa_b200.so does not link it).
 * Rank 0 calls fa_b200_nccl_unique_id and hands the 128 bytes to every rank (fa_b200/mt/comm.py: TCP rendezvous on
 * MASTER_ADDR / MASTER_PORT); every rank then calls fa_b200_nccl_init on its CUDA device.  fa_b200_nccl_allreduce_f32
 * sums (average = 0) or averages (average = 1) `count` floats in place on `stream`; it may be captured into a CUDA graph.
 * All return FA_OK or an error code with the text in fa_b200_nccl_last_error().
 * ------------------------------------------------------------------------------------------ */
#define FA_B200_NCCL_UNIQUE_ID_BYTES 128
int fa_b200_nccl_unique_id(void* id_out /* 128 bytes */);
int fa_b200_nccl_init(const void* id /* 128 bytes */, int rank, int world);
int fa_b200_nccl_allreduce_f32(float* buf, size_t count, int average, void* stream);
int fa_b200_nccl_world(void);   /* 0 before init */
int fa_b200_nccl_rank(void);    /* -1 before init */
int fa_b200_nccl_version(void); /* NCCL_VERSION_CODE of the loaded library, -1 if libnccl cannot be loaded */
int fa_b200_nccl_destroy(void);
const char* fa_b200_nccl_last_error(void);

/* ------------------------------------------------------------------------------------------
 * The reference entry points themselves.  Exported by the six shim libraries (not by libfa_b200.so,
 * where the two conventions would collide on one symbol name).  `stream` is a cudaStream_t.
 * On failure they print to stderr and exit(EXIT_FAILURE), as the reference does
 * (src/flash_attn2_fw.cu:356-362).
 * ------------------------------------------------------------------------------------------ */
#ifdef FA_B200_DECLARE_REFERENCE_ABI
void launch_flash_attn_fw(float* q, float* k, float* v, float* out, float* l, float* m, int batch, int N, int d,
                          bool causal_mask, void* stream);
void launch_flash_attn_bw(float* q, float* k, float* v, float* out, float* out_grad, float* q_grad, float* k_grad,
                          float* v_grad, float* l, float* m, int batch, int N, int d, bool causal_mask,
                          void* stream);
#endif

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
