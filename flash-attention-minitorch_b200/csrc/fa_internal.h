// fa_internal.h -- declarations shared between the translation units of libfa_b200.so.
// Nothing here is part of the C ABI (see include/fa_b200.h for that).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "fa_dropout.cuh"

namespace fa {

constexpr Dropout kNoDropout = {0u, 1.0f, 0u, 0u};

// ---- fp32 CUDA-core path (fa_simt.cu) -------------------------------------------------------
cudaError_t simt_fwd(const float* q, const float* k, const float* v, float* o, float* lse, float* m_out,
                     float* l_out, int BH, int T, int d, bool causal, cudaStream_t st, Dropout drop = kNoDropout);
cudaError_t simt_bwd_prep(const float* o, const float* dO, float* delta, const float* l_in, const float* m_in,
                          float* lse_out, size_t rows, int d, cudaStream_t st);
cudaError_t simt_bwd(const float* q, const float* k, const float* v, const float* dO, const float* lse,
                     const float* delta, float* dq, float* dk, float* dv, int BH, int T, int d, bool causal,
                     bool accumulate_dq, cudaStream_t st, Dropout drop = kNoDropout);

// ---- tcgen05 path (fa_fwd_sm100.cu / fa_bwd_sm100.cu) ---------------------------------------
// bf16 Q/K/V laid out (BH, T, d) contiguous.  `o` is bf16 or fp32 depending on o_is_f32.
// lse = natural-log logsumexp of tau*QK^T per row (the FA-2 "L"); m_out/l_out (optional) give the
// FA-1 pair (row max of tau*S, sum exp(tau*S - m)).
bool tc_fwd_supported(int d);
void tc_fwd_set_persistent(int on);  // A/B switch: 0 = one CTA per 256-row block (default: persistent, FA_B200_FWD_PERSISTENT)
int tc_fwd_get_persistent();
cudaError_t tc_fwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, void* o, bool o_is_f32,
                   float* lse, float* m_out, float* l_out, int BH, int T, int d, bool causal, cudaStream_t st,
                   Dropout drop = kNoDropout);

// the same at d = 128 on CTA pairs (tcgen05.mma.cta_group::2, P through shared memory): fa_fwd_pair_sm100.cu
bool fwd_pair_enabled();
void set_fwd_pair(bool on);
cudaError_t tc_fwd_pair(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, void* o, bool o_is_f32, float* lse,
                        float* m_out, float* l_out, int BH, int T, bool causal, cudaStream_t st, Dropout drop = kNoDropout);

// fp32 storage read directly by TMA, kind::tf32 MMAs, fp32 output; d in {32, 64}
bool tc_fwd_tf32_supported(int d);
cudaError_t tc_fwd_tf32(const float* q, const float* k, const float* v, float* o, float* lse, float* m_out,
                        float* l_out, int BH, int T, int d, bool causal, cudaStream_t st, Dropout drop = kNoDropout);

bool tc_bwd_supported(int d);
int tc_bwd_num_launches(bool grads_f32);
void tc_fwd_set_trace(long long* dev_buf);  // same for the forward (CTA (0,0): the heaviest causal block)
void tc_bwd_set_trace(long long* dev_buf);  // developer timeline buffer (256 int64), nullptr = off
void tc_bwd_set_trace_start(int t0);        // first tile pair (running count of CTA 0) the timeline records
void tc_bwd_set_persistent(int on);         // A/B switch: 0 = one CTA per work item (default: persistent, FA_B200_BWD_PERSISTENT)
int tc_bwd_get_persistent();
size_t tc_bwd_workspace_bytes(int BH, int T, int d);
// dq/dk/dv are bf16 or fp32 (grads_f32).  workspace: tc_bwd_workspace_bytes() bytes of device memory.
cudaError_t tc_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* o,
                   const __nv_bfloat16* dO, const float* lse, void* dq, void* dk, void* dv, bool grads_f32,
                   bool accumulate_dq, void* workspace, int BH, int T, int d, bool causal, cudaStream_t st,
                   Dropout drop = kNoDropout);

// fp32 storage, tf32 tensor cores, head dim 32 or 64 (fa_bwd_tf32_sm100.cu): fp32 gradients, dq accumulated into when asked
bool tc_bwd_tf32_supported(int d);
size_t tc_bwd_tf32_workspace_bytes(int BH, int T, int d);
cudaError_t tc_bwd_tf32(const float* q, const float* k, const float* v, const float* o, const float* dO, const float* lse,
                        float* dq, float* dk, float* dv, bool accumulate_dq, void* workspace, int BH, int T, int d, bool causal,
                        cudaStream_t st);

// ---- elementwise helpers (fa_util.cu) -------------------------------------------------------
cudaError_t cvt_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t st);
cudaError_t cvt_bf16_to_f32(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t st);
// lse = m + log(l)   (FA-1 convention -> FA-2 convention)
cudaError_t lse_from_lm(const float* l, const float* m, float* lse, size_t n, cudaStream_t st);
// keep[h - head_begin][i][j] = 1/0 for heads [head_begin, head_begin + n_heads): the mask every kernel regenerates
cudaError_t dropout_mask(unsigned char* keep, int head_begin, int n_heads, int T, Dropout drop, cudaStream_t st);

// ---- decoder-LM operators around the attention path (fa_gemm_sm100.cu, fa_lm_kernels.cu) --------
// C[M x N] (+)= A * B (+ bias[N]); fp32 storage, tf32 tensor cores.  a_mn_major: A stored [K][M] (M contiguous) instead
// of [M][K]; b_mn_major: B stored [K][N] (N contiguous) instead of [N][K].  M, N, K multiples of 4.
// HeadLayout: which operands are logical [B*T][H*d] matrices stored as (B, H, T, d) (T % 128 == 0, d % 32 == 0)
struct HeadLayout {
  int a, b, c;  // flags
  int T, H, d;
};
constexpr HeadLayout kNoHeads = {0, 0, 0, 0, 0, 0};
cudaError_t gemm_tf32(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, bool a_mn_major,
                      bool b_mn_major, bool accumulate, cudaStream_t st, HeadLayout hl = kNoHeads);
cudaError_t colsum_heads_f32(const float* x, float* out, int B, int H, int T, int d, bool accumulate, cudaStream_t st,
                             int parts = 1);  // x = (parts, B, H, T, d), out = [parts][H * d]
cudaError_t colsum_f32(const float* x, float* out, int M, int N, bool accumulate, cudaStream_t st);
// fa_lm_kernels.cu: one fused pass per operator (dims multiples of 4; dw/db/embedding-dw are zeroed inside)
cudaError_t layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, int rows,
                          int dim, float eps, cudaStream_t st);
// a (+ b -> sum, the residual sum) then LayerNorm; b / sum may be null.  b needs dim in {128, 256, 512}
cudaError_t add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, float* sum, float* y,
                              float* mean, float* rstd, int rows, int dim, float eps, cudaStream_t st);
bool layernorm_fused_width(int dim);
// res, dcol (optional, fused widths only): dx = LN'(dy) + res;  dcol (+)= column sums of dx
cudaError_t layernorm_bwd(const float* x, const float* w, const float* dy, const float* mean, const float* rstd, float* dx,
                          float* dw, float* db, int rows, int dim, cudaStream_t st, bool accumulate = false,
                          const float* res = nullptr, float* dcol = nullptr);
cudaError_t gelu_dropout(const float* x, const float* dy, float* y, size_t n, bool backward, Dropout drop,
                         const unsigned long long* seed_offset, cudaStream_t st);
cudaError_t gelu(const float* x, const float* dy, float* y, size_t n, bool backward, cudaStream_t st);
// backward of drop(gelu(x)), x (rows, dim in {128, 256, 512}), plus dcol (+)= column sums of dx
cudaError_t gelu_dropout_bwd_colsum(const float* x, const float* dy, float* dx, float* dcol, int rows, int dim, Dropout drop,
                                    const unsigned long long* seed_offset, bool accumulate, cudaStream_t st);
cudaError_t dropout_apply(const float* x, float* y, size_t n, Dropout drop, const unsigned long long* seed_offset,
                          cudaStream_t st);
cudaError_t embedding_fwd(const long long* idx, const float* w, float* y, size_t n_idx, int dim, cudaStream_t st);
cudaError_t embedding_bwd(const long long* idx, const float* dy, float* dw, size_t n_idx, int dim, size_t vocab, cudaStream_t st,
                          bool accumulate = false);
// y[r] = drop(tok[idx[r]] + pos[r % T]) and its backward (scatter-add onto both tables; zeroed first unless accumulate)
cudaError_t embed_pos_dropout_fwd(const long long* idx, const float* tok, const float* pos, float* y, size_t n_idx, int T, int dim,
                                  Dropout drop, const unsigned long long* seed_offset, cudaStream_t st);
cudaError_t embed_pos_dropout_bwd(const long long* idx, const float* dy, float* dtok, float* dpos, size_t n_idx, int T, int dim,
                                  size_t tok_rows, size_t pos_rows, Dropout drop, const unsigned long long* seed_offset,
                                  bool accumulate, cudaStream_t st);
cudaError_t softmax_loss_fwd(const float* logits, const long long* target, float* loss, float* lse, int rows, int C,
                             cudaStream_t st);
cudaError_t softmax_loss_bwd(const float* logits, const long long* target, const float* lse, const float* gscale,
                             float* dlogits, int rows, int C, cudaStream_t st);
// + dcol (+)= column sums of dlogits (the logits layer's bias gradient); C <= 49152
bool softmax_loss_colsum_width(int C);
cudaError_t softmax_loss_bwd_colsum(const float* logits, const long long* target, const float* lse, const float* gscale,
                                    float* dlogits, float* dcol, int rows, int C, bool accumulate, cudaStream_t st);
cudaError_t adam_step(float* p, const float* g, float* m, float* v, size_t n, float step_size, float b1, float b2,
                      float v_coef, float eps, cudaStream_t st);

}  // namespace fa
