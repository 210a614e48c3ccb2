// fa_capi.cu -- the C ABI of libfa_b200.so (see include/fa_b200.h).
//
// Device-resident entry points dispatch to the tcgen05 kernels (bf16 / tf32) or the fp32 CUDA-core
// kernels.  Host-pointer entry points reproduce the reference launchers
// (src/flash_attn2_fw.cu:312-372, src/flash_attn2_bw.cu:279-369, src/flash_attn_fw.cu:302-359,
// src/flash_attn_bw.cu:277-365): stage in, run, stage out, synchronise -- but through a persistent
// grow-only device arena instead of 6-10 cudaMalloc/cudaFree pairs per call.
#include <stdarg.h>
#include <stdio.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/fa_b200.h"
#include "fa_internal.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define FA_CUDA(expr, nlaunch)                                                                  \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(FA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    g_launches += (nlaunch);                                                                    \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- persistent device arena for the host-pointer API ---------------------------------------
struct Arena {
  void* base = nullptr;
  size_t cap = 0;
  int device = -1;
  size_t used = 0;
};
Arena g_arena;
std::mutex g_arena_mu;

int arena_reset(size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(FA_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (g_arena.base && (g_arena.device != dev || g_arena.cap < bytes)) {
    cudaFree(g_arena.base);
    g_arena = Arena();
  }
  if (!g_arena.base) {
    size_t cap = align_up(bytes + (bytes >> 3), (size_t)2 << 20);
    e = cudaMalloc(&g_arena.base, cap);
    if (e != cudaSuccess) {
      g_arena = Arena();
      return fail(FA_ERR_CUDA, "cudaMalloc(%zu) for the staging arena: %s", cap, cudaGetErrorString(e));
    }
    g_arena.cap = cap;
    g_arena.device = dev;
  }
  g_arena.used = 0;
  return FA_OK;
}
template <typename T>
T* arena_take(size_t count) {
  size_t off = align_up(g_arena.used, 256);
  g_arena.used = off + count * sizeof(T);
  return reinterpret_cast<T*>(static_cast<char*>(g_arena.base) + off);
}

// ---- copy/compute pipeline of the host-pointer API -------------------------------------------
constexpr int kPipeBufs = 3;
struct HostPipe {
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kPipeBufs] = {}, ev_done[kPipeBufs] = {}, ev_free[kPipeBufs] = {};
  int device = -1;
  void destroy() {
    if (device < 0) return;
    for (int i = 0; i < kPipeBufs; ++i) {
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
      if (ev_free[i]) cudaEventDestroy(ev_free[i]);
      ev_in[i] = ev_done[i] = ev_free[i] = nullptr;
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    s_in = s_out = nullptr;
    device = -1;
  }
  int ensure() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(FA_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    if (device == dev) return FA_OK;
    destroy();
    if ((e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) != cudaSuccess)
      return fail(FA_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    for (int i = 0; i < kPipeBufs; ++i)
      if ((e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming)) != cudaSuccess)
        return fail(FA_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    device = dev;
    return FA_OK;
  }
};
HostPipe g_pipe;

// heads per pipeline chunk: ~64 MB of fp32 per tensor (a few ms of PCIe), whole heads only
int chunk_heads(int batch, size_t elems_per_head) {
  size_t target = (size_t)16 << 20;  // elements
  if (const char* e = getenv("FA_B200_HOST_CHUNK_ELEMS")) {  // test hook: force many small chunks
    const long long v = atoll(e);
    if (v > 0) target = (size_t)v;
  }
  size_t hc = target / (elems_per_head ? elems_per_head : 1);
  if (hc < 1) hc = 1;
  if (hc > (size_t)batch) hc = (size_t)batch;
  return (int)hc;
}

int g_host_precision = -1;  // -1: not yet initialised from the environment
int host_precision() {
  if (g_host_precision < 0) {
    const char* s = getenv("FA_B200_HOST_PRECISION");
    int v = FA_PREC_AUTO;
    if (s) {
      if (!strcmp(s, "fp32") || !strcmp(s, "simt")) v = FA_PREC_FP32_SIMT;
      else if (!strcmp(s, "tf32")) v = FA_PREC_TF32_TC;
      else if (!strcmp(s, "bf16")) v = FA_PREC_BF16_TC;
    }
    g_host_precision = v;
  }
  return g_host_precision;
}
// "auto" on the reference ABI = tf32 tensor cores wherever a tcgen05 kernel covers the shape (resolve_precision falls
// back to the exact-fp32 CUDA-core kernels for the rest): it meets the fp32-path bar of north_star (max-abs 1e-3
// against the reference, |dO| <= 1: forward 4.9e-4, backward 7.4e-4 measured) and is what a user who only drops the
// six .so files in gets.  FA_B200_HOST_PRECISION=fp32 (or fa_b200_set_host_precision) keeps exact fp32 arithmetic.
int host_effective_precision() { return host_precision() == FA_PREC_AUTO ? FA_PREC_TF32_TC : host_precision(); }

// FA_B200_HOST_ZERO_QGRAD=1: the caller promises that q_grad holds zeros on entry -- what every caller of the reference
// passes (minitorch/cuda_kernel_ops.py:609-611) -- so the backward skips its upload (11 % of the H2D bytes) and
// overwrites instead of accumulating.  Off by default: the reference kernels accumulate (src/flash_attn2_bw.cu:228).
int g_host_zero_qgrad = -1;  // -1: not yet initialised from the environment
bool host_zero_qgrad() {
  if (g_host_zero_qgrad < 0) {
    const char* s = getenv("FA_B200_HOST_ZERO_QGRAD");
    g_host_zero_qgrad = (s && s[0] == '1') ? 1 : 0;
  }
  return g_host_zero_qgrad == 1;
}

bool valid_shape(int bh, int T, int d) { return bh > 0 && T > 0 && d > 0 && d <= 128; }

// Which arithmetic actually runs for (dtype, requested precision, d).  Returns <0 if unsupported.
int resolve_precision(int dtype, int precision, int d, bool backward) {
  const bool tc_ok = backward ? fa::tc_bwd_supported(d) : fa::tc_fwd_supported(d);
  if (dtype == FA_DTYPE_BF16) {
    if (precision == FA_PREC_AUTO || precision == FA_PREC_BF16_TC) return tc_ok ? FA_PREC_BF16_TC : FA_PREC_FP32_SIMT;
    if (precision == FA_PREC_FP32_SIMT) return FA_PREC_FP32_SIMT;
    return -1;
  }
  if (dtype == FA_DTYPE_F32) {
    // tf32 tensor-core kernels exist for the forward and the backward at d in {32, 64}; everything else on fp32
    // storage runs the exact-fp32 CUDA-core kernels, which satisfy any tf32 tolerance
    const bool tf32_ok = backward ? fa::tc_bwd_tf32_supported(d) : fa::tc_fwd_tf32_supported(d);
    if (precision == FA_PREC_AUTO || precision == FA_PREC_TF32_TC) return tf32_ok ? FA_PREC_TF32_TC : FA_PREC_FP32_SIMT;
    if (precision == FA_PREC_FP32_SIMT) return FA_PREC_FP32_SIMT;
    if (precision == FA_PREC_BF16_TC) return tc_ok ? FA_PREC_BF16_TC : FA_PREC_FP32_SIMT;
  }
  return -1;
}

}  // namespace

extern "C" {

int fa_b200_version(void) { return 100; }
const char* fa_b200_last_error(void) { return g_err; }
unsigned long long fa_b200_launch_count(void) { return g_launches.load(); }

}  // extern "C"

namespace {

int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, float* m_out, float* l_out, int dtype,
             int precision, int batch_heads, int T, int d, int causal, fa::Dropout drop, void* stream) {
  if (!q || !k || !v || !o) return fail(FA_ERR_BAD_ARG, "fa_b200_fwd: null tensor pointer");
  if (!valid_shape(batch_heads, T, d)) return fail(FA_ERR_BAD_ARG, "fa_b200_fwd: bad shape BH=%d T=%d d=%d", batch_heads, T, d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int prec = resolve_precision(dtype, precision, d, false);
  if (prec < 0) return fail(FA_ERR_BAD_ARG, "fa_b200_fwd: dtype %d / precision %d not supported", dtype, precision);
  if (dtype == FA_DTYPE_BF16) {
    if (prec != FA_PREC_BF16_TC)
      return fail(FA_ERR_UNSUPPORTED, "fa_b200_fwd: bf16 storage needs d in {64,128} (got %d)", d);
    FA_CUDA(fa::tc_fwd((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, o, false, lse, m_out,
                       l_out, batch_heads, T, d, causal != 0, st, drop), 1);
    return FA_OK;
  }
  if (prec == FA_PREC_TF32_TC) {
    FA_CUDA(fa::tc_fwd_tf32((const float*)q, (const float*)k, (const float*)v, (float*)o, lse, m_out, l_out, batch_heads,
                            T, d, causal != 0, st, drop), 1);
    return FA_OK;
  }
  if (prec != FA_PREC_FP32_SIMT)
    return fail(FA_ERR_UNSUPPORTED, "fa_b200_fwd: fp32 storage with bf16 tensor-core precision needs the host API");
  FA_CUDA(fa::simt_fwd((const float*)q, (const float*)k, (const float*)v, (float*)o, lse, m_out, l_out, batch_heads, T,
                       d, causal != 0, st, drop), 1);
  return FA_OK;
}

// p_drop in [0, 1): returns nonzero (and sets the error string) when out of range
int check_p_drop(float p_drop, const char* who) {
  if (!(p_drop >= 0.f) || !(p_drop < 1.f)) return fail(FA_ERR_BAD_ARG, "%s: p_drop %g outside [0, 1)", who, (double)p_drop);
  return FA_OK;
}

}  // namespace

extern "C" {

int fa_b200_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* m_out, float* l_out,
                int dtype, int precision, int batch_heads, int T, int d, int causal, void* stream) {
  return fwd_impl(q, k, v, o, lse, m_out, l_out, dtype, precision, batch_heads, T, d, causal, fa::kNoDropout, stream);
}

int fa_b200_fwd_dropout(const void* q, const void* k, const void* v, void* o, float* lse, int dtype, int precision,
                        int batch_heads, int T, int d, int causal, float p_drop, unsigned long long seed,
                        void* stream) {
  if (int rc = check_p_drop(p_drop, "fa_b200_fwd_dropout")) return rc;
  return fwd_impl(q, k, v, o, lse, nullptr, nullptr, dtype, precision, batch_heads, T, d, causal,
                  fa::make_dropout(p_drop, seed), stream);
}

float fa_b200_dropout_keep_scale(float p_drop) { return fa::make_dropout(p_drop, 0).keep_scale; }
float fa_b200_dropout_effective_p(float p_drop) { return (float)fa::make_dropout(p_drop, 0).thr / 128.0f; }

int fa_b200_dropout_mask(unsigned char* keep, int head_begin, int n_heads, int T, float p_drop,
                         unsigned long long seed, void* stream) {
  if (!keep || head_begin < 0 || n_heads <= 0 || n_heads > 65535 || T <= 0)
    return fail(FA_ERR_BAD_ARG, "fa_b200_dropout_mask: bad argument");
  if (int rc = check_p_drop(p_drop, "fa_b200_dropout_mask")) return rc;
  FA_CUDA(fa::dropout_mask(keep, head_begin, n_heads, T, fa::make_dropout(p_drop, seed), static_cast<cudaStream_t>(stream)), 1);
  return FA_OK;
}

size_t fa_b200_bwd_workspace_bytes(int dtype, int precision, int batch_heads, int T, int d) {
  if (!valid_shape(batch_heads, T, d)) return 0;
  const int prec = resolve_precision(dtype, precision, d, true);
  const size_t rows = (size_t)batch_heads * T, n = rows * d;
  if (prec == FA_PREC_BF16_TC && dtype == FA_DTYPE_BF16) return fa::tc_bwd_workspace_bytes(batch_heads, T, d);
  size_t bytes = align_up(rows * sizeof(float), 256);            // delta
  if (dtype == FA_DTYPE_BF16) bytes += 8 * align_up(n * sizeof(float), 256);  // fp32 images of 5 inputs + 3 grads
  // the tf32 backward (padded L / delta) -- the larger of the two, so that the dropout fallback to the CUDA-core kernels fits
  if (prec == FA_PREC_TF32_TC) bytes = std::max(bytes, fa::tc_bwd_tf32_workspace_bytes(batch_heads, T, d));
  return bytes;
}

}  // extern "C"

namespace {

int bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* d_out, const float* lse, void* dq,
             void* dk, void* dv, int dtype, int precision, int batch_heads, int T, int d, int causal, int accumulate_dq,
             fa::Dropout drop, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !o || !d_out || !lse || !dq || !dk || !dv)
    return fail(FA_ERR_BAD_ARG, "fa_b200_bwd: null tensor pointer");
  if (!valid_shape(batch_heads, T, d)) return fail(FA_ERR_BAD_ARG, "fa_b200_bwd: bad shape BH=%d T=%d d=%d", batch_heads, T, d);
  const int prec = resolve_precision(dtype, precision, d, true);
  if (prec < 0) return fail(FA_ERR_BAD_ARG, "fa_b200_bwd: dtype %d / precision %d not supported", dtype, precision);
  const size_t need = fa_b200_bwd_workspace_bytes(dtype, precision, batch_heads, T, d);
  if (need && (!workspace || workspace_bytes < need))
    return fail(FA_ERR_WORKSPACE, "fa_b200_bwd: workspace %zu bytes < required %zu", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t rows = (size_t)batch_heads * T, n = rows * d;
  char* ws = static_cast<char*>(workspace);

  if (dtype == FA_DTYPE_BF16 && prec == FA_PREC_BF16_TC) {
    FA_CUDA(fa::tc_bwd((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                       (const __nv_bfloat16*)o, (const __nv_bfloat16*)d_out, lse, dq, dk, dv, false, accumulate_dq != 0,
                       workspace, batch_heads, T, d, causal != 0, st, drop), fa::tc_bwd_num_launches(false));
    return FA_OK;
  }
  if (dtype == FA_DTYPE_F32 && prec == FA_PREC_TF32_TC && !drop.thr) {
    FA_CUDA(fa::tc_bwd_tf32((const float*)q, (const float*)k, (const float*)v, (const float*)o, (const float*)d_out, lse,
                            (float*)dq, (float*)dk, (float*)dv, accumulate_dq != 0, workspace, batch_heads, T, d, causal != 0, st), 2);
    return FA_OK;
  }
  float* delta = reinterpret_cast<float*>(ws);
  ws += align_up(rows * sizeof(float), 256);
  if (dtype == FA_DTYPE_F32) {
    FA_CUDA(fa::simt_bwd_prep((const float*)o, (const float*)d_out, delta, nullptr, nullptr, nullptr, rows, d, st), 1);
    FA_CUDA(fa::simt_bwd((const float*)q, (const float*)k, (const float*)v, (const float*)d_out, lse, delta, (float*)dq,
                         (float*)dk, (float*)dv, batch_heads, T, d, causal != 0, accumulate_dq != 0, st, drop), 2);
    return FA_OK;
  }
  // bf16 storage, head dim outside the tensor-core set: widen to fp32, run the CUDA-core kernels, narrow back
  float* f[8];
  for (int i = 0; i < 8; ++i) {
    f[i] = reinterpret_cast<float*>(ws);
    ws += align_up(n * sizeof(float), 256);
  }
  const void* src[5] = {q, k, v, o, d_out};
  for (int i = 0; i < 5; ++i) FA_CUDA(fa::cvt_bf16_to_f32((const __nv_bfloat16*)src[i], f[i], n, st), 1);
  if (accumulate_dq) FA_CUDA(fa::cvt_bf16_to_f32((const __nv_bfloat16*)dq, f[5], n, st), 1);
  FA_CUDA(fa::simt_bwd_prep(f[3], f[4], delta, nullptr, nullptr, nullptr, rows, d, st), 1);
  FA_CUDA(fa::simt_bwd(f[0], f[1], f[2], f[4], lse, delta, f[5], f[6], f[7], batch_heads, T, d, causal != 0,
                       accumulate_dq != 0, st, drop), 2);
  void* dst[3] = {dq, dk, dv};
  for (int i = 0; i < 3; ++i) FA_CUDA(fa::cvt_f32_to_bf16(f[5 + i], (__nv_bfloat16*)dst[i], n, st), 1);
  return FA_OK;
}

}  // namespace

extern "C" {

int fa_b200_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_out, const float* lse,
                void* dq, void* dk, void* dv, int dtype, int precision, int batch_heads, int T, int d, int causal,
                int accumulate_dq, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_impl(q, k, v, o, d_out, lse, dq, dk, dv, dtype, precision, batch_heads, T, d, causal, accumulate_dq,
                  fa::kNoDropout, workspace, workspace_bytes, stream);
}

int fa_b200_bwd_dropout(const void* q, const void* k, const void* v, const void* o, const void* d_out, const float* lse,
                        void* dq, void* dk, void* dv, int dtype, int precision, int batch_heads, int T, int d,
                        int causal, int accumulate_dq, float p_drop, unsigned long long seed, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (int rc = check_p_drop(p_drop, "fa_b200_bwd_dropout")) return rc;
  return bwd_impl(q, k, v, o, d_out, lse, dq, dk, dv, dtype, precision, batch_heads, T, d, causal, accumulate_dq,
                  fa::make_dropout(p_drop, seed), workspace, workspace_bytes, stream);
}

// ---- decoder-LM operators (device pointers, fp32) ------------------------------------------
int fa_b200_gemm_tf32(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, int a_mn_major,
                      int b_mn_major, int accumulate, void* stream) {
  if (!a || !b || !c) return fail(FA_ERR_BAD_ARG, "fa_b200_gemm_tf32: null pointer");
  if (M <= 0 || N <= 0 || K <= 0 || (M & 3) || (N & 3) || (K & 3))
    return fail(FA_ERR_BAD_ARG, "fa_b200_gemm_tf32: M, N, K must be positive multiples of 4 (got %d %d %d)", M, N, K);
  FA_CUDA(fa::gemm_tf32(a, b, c, bias, M, N, K, a_mn_major != 0, b_mn_major != 0, accumulate != 0,
                        static_cast<cudaStream_t>(stream)), 1);
  return FA_OK;
}

int fa_b200_gemm_tf32_heads(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, int a_mn_major,
                            int b_mn_major, int accumulate, int a_heads, int b_heads, int c_heads, int T, int H, int d,
                            void* stream) {
  if (!a || !b || !c) return fail(FA_ERR_BAD_ARG, "fa_b200_gemm_tf32_heads: null pointer");
  const fa::HeadLayout hl = {a_heads != 0, b_heads != 0, c_heads != 0, T, H, d};
  cudaError_t e = fa::gemm_tf32(a, b, c, bias, M, N, K, a_mn_major != 0, b_mn_major != 0, accumulate != 0,
                                static_cast<cudaStream_t>(stream), hl);
  if (e == cudaErrorInvalidValue)
    return fail(FA_ERR_BAD_ARG, "fa_b200_gemm_tf32_heads: unsupported shape M=%d N=%d K=%d T=%d H=%d d=%d (needs multiples of 4, "
                "T %% 128 == 0, d %% 32 == 0, head-layout operand width a multiple of H*d)", M, N, K, T, H, d);
  FA_CUDA(e, 1);
  return FA_OK;
}

int fa_b200_colsum_heads_f32(const float* x, float* out, int B, int H, int T, int d, int accumulate, void* stream) {
  if (!x || !out) return fail(FA_ERR_BAD_ARG, "fa_b200_colsum_heads_f32: null pointer");
  FA_CUDA(fa::colsum_heads_f32(x, out, B, H, T, d, accumulate != 0, static_cast<cudaStream_t>(stream)), 1);
  return FA_OK;
}

int fa_b200_colsum_heads_parts_f32(const float* x, float* out, int parts, int B, int H, int T, int d, int accumulate, void* stream) {
  if (!x || !out) return fail(FA_ERR_BAD_ARG, "fa_b200_colsum_heads_parts_f32: null pointer");
  FA_CUDA(fa::colsum_heads_f32(x, out, B, H, T, d, accumulate != 0, static_cast<cudaStream_t>(stream), parts), 1);
  return FA_OK;
}

int fa_b200_colsum_f32(const float* x, float* out, int M, int N, int accumulate, void* stream) {
  if (!x || !out || M <= 0 || N <= 0) return fail(FA_ERR_BAD_ARG, "fa_b200_colsum_f32: bad argument");
  FA_CUDA(fa::colsum_f32(x, out, M, N, accumulate != 0, static_cast<cudaStream_t>(stream)), 1);
  return FA_OK;
}

#define FA_ST static_cast<cudaStream_t>(stream)
int fa_b200_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, int rows,
                          int dim, float eps, void* stream) {
  if (!x || !w || !b || !y || !mean || !rstd) return fail(FA_ERR_BAD_ARG, "fa_b200_layernorm_fwd: null pointer");
  FA_CUDA(fa::layernorm_fwd(x, w, b, y, mean, rstd, rows, dim, eps, FA_ST), 1);
  return FA_OK;
}
int fa_b200_layernorm_bwd(const float* x, const float* w, const float* dy, const float* mean, const float* rstd, float* dx,
                          float* dw, float* db, int rows, int dim, void* stream) {
  if (!x || !w || !dy || !mean || !rstd || !dx || !dw || !db) return fail(FA_ERR_BAD_ARG, "fa_b200_layernorm_bwd: null pointer");
  FA_CUDA(fa::layernorm_bwd(x, w, dy, mean, rstd, dx, dw, db, rows, dim, FA_ST), fa::layernorm_fused_width(dim) ? 1 : 2);
  return FA_OK;
}
int fa_b200_layernorm_bwd_accum(const float* x, const float* w, const float* dy, const float* mean, const float* rstd, float* dx,
                                float* dw, float* db, int rows, int dim, void* stream) {
  if (!x || !w || !dy || !mean || !rstd || !dx || !dw || !db) return fail(FA_ERR_BAD_ARG, "fa_b200_layernorm_bwd_accum: null pointer");
  FA_CUDA(fa::layernorm_bwd(x, w, dy, mean, rstd, dx, dw, db, rows, dim, FA_ST, true), fa::layernorm_fused_width(dim) ? 1 : 2);
  return FA_OK;
}
int fa_b200_add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, float* sum, float* y,
                              float* mean, float* rstd, int rows, int dim, float eps, void* stream) {
  if (!a || !w || !bias || !y || !mean || !rstd || (b && !sum)) return fail(FA_ERR_BAD_ARG, "fa_b200_add_layernorm_fwd: null pointer");
  if (b && !fa::layernorm_fused_width(dim)) return fail(FA_ERR_UNSUPPORTED, "fa_b200_add_layernorm_fwd: the fused add needs dim in {128, 256, 512}");
  FA_CUDA(fa::add_layernorm_fwd(a, b, w, bias, sum, y, mean, rstd, rows, dim, eps, FA_ST), 1);
  return FA_OK;
}
int fa_b200_layernorm_bwd_res(const float* x, const float* w, const float* dy, const float* mean, const float* rstd,
                              const float* res, float* dx, float* dw, float* db, float* dcol, int rows, int dim, int accumulate,
                              void* stream) {
  if (!x || !w || !dy || !mean || !rstd || !dx || !dw || !db) return fail(FA_ERR_BAD_ARG, "fa_b200_layernorm_bwd_res: null pointer");
  if ((res || dcol) && !fa::layernorm_fused_width(dim)) return fail(FA_ERR_UNSUPPORTED, "fa_b200_layernorm_bwd_res: res / dcol need dim in {128, 256, 512}");
  FA_CUDA(fa::layernorm_bwd(x, w, dy, mean, rstd, dx, dw, db, rows, dim, FA_ST, accumulate != 0, res, dcol), fa::layernorm_fused_width(dim) ? 1 : 2);
  return FA_OK;
}
int fa_b200_gelu_dropout_fwd(const float* x, float* y, size_t n, float p_drop, unsigned long long seed,
                             const unsigned long long* seed_offset_dev, void* stream) {
  if (!x || !y) return fail(FA_ERR_BAD_ARG, "fa_b200_gelu_dropout_fwd: null pointer");
  if (int rc = check_p_drop(p_drop, "fa_b200_gelu_dropout_fwd")) return rc;
  FA_CUDA(fa::gelu_dropout(x, nullptr, y, n, false, fa::make_dropout(p_drop, seed), seed_offset_dev, FA_ST), 1);
  return FA_OK;
}
int fa_b200_gelu_dropout_bwd(const float* x, const float* dy, float* dx, size_t n, float p_drop, unsigned long long seed,
                             const unsigned long long* seed_offset_dev, void* stream) {
  if (!x || !dy || !dx) return fail(FA_ERR_BAD_ARG, "fa_b200_gelu_dropout_bwd: null pointer");
  if (int rc = check_p_drop(p_drop, "fa_b200_gelu_dropout_bwd")) return rc;
  FA_CUDA(fa::gelu_dropout(x, dy, dx, n, true, fa::make_dropout(p_drop, seed), seed_offset_dev, FA_ST), 1);
  return FA_OK;
}
int fa_b200_layernorm_fused_width(int dim) { return fa::layernorm_fused_width(dim) ? 1 : 0; }
int fa_b200_gelu_fwd(const float* x, float* y, size_t n, void* stream) {
  if (!x || !y) return fail(FA_ERR_BAD_ARG, "fa_b200_gelu_fwd: null pointer");
  FA_CUDA(fa::gelu(x, nullptr, y, n, false, FA_ST), 1);
  return FA_OK;
}
int fa_b200_gelu_bwd(const float* x, const float* dy, float* dx, size_t n, void* stream) {
  if (!x || !dy || !dx) return fail(FA_ERR_BAD_ARG, "fa_b200_gelu_bwd: null pointer");
  FA_CUDA(fa::gelu(x, dy, dx, n, true, FA_ST), 1);
  return FA_OK;
}
int fa_b200_gelu_dropout_bwd_colsum(const float* x, const float* dy, float* dx, float* dcol, int rows, int dim, float p_drop,
                                    unsigned long long seed, const unsigned long long* seed_offset_dev, int accumulate, void* stream) {
  if (!x || !dy || !dx || !dcol) return fail(FA_ERR_BAD_ARG, "fa_b200_gelu_dropout_bwd_colsum: null pointer");
  if (!fa::layernorm_fused_width(dim)) return fail(FA_ERR_UNSUPPORTED, "fa_b200_gelu_dropout_bwd_colsum: dim must be 128, 256 or 512");
  if (int rc = check_p_drop(p_drop, "fa_b200_gelu_dropout_bwd_colsum")) return rc;
  FA_CUDA(fa::gelu_dropout_bwd_colsum(x, dy, dx, dcol, rows, dim, fa::make_dropout(p_drop, seed), seed_offset_dev, accumulate != 0, FA_ST), 1);
  return FA_OK;
}
int fa_b200_dropout_apply(const float* x, float* y, size_t n, float p_drop, unsigned long long seed,
                          const unsigned long long* seed_offset_dev, void* stream) {
  if (!x || !y) return fail(FA_ERR_BAD_ARG, "fa_b200_dropout_apply: null pointer");
  if (int rc = check_p_drop(p_drop, "fa_b200_dropout_apply")) return rc;
  FA_CUDA(fa::dropout_apply(x, y, n, fa::make_dropout(p_drop, seed), seed_offset_dev, FA_ST), 1);
  return FA_OK;
}
int fa_b200_embedding_fwd(const long long* idx, const float* w, float* y, size_t n_idx, int dim, void* stream) {
  if (!idx || !w || !y) return fail(FA_ERR_BAD_ARG, "fa_b200_embedding_fwd: null pointer");
  FA_CUDA(fa::embedding_fwd(idx, w, y, n_idx, dim, FA_ST), 1);
  return FA_OK;
}
int fa_b200_embedding_bwd(const long long* idx, const float* dy, float* dw, size_t n_idx, int dim, size_t vocab, void* stream) {
  if (!idx || !dy || !dw) return fail(FA_ERR_BAD_ARG, "fa_b200_embedding_bwd: null pointer");
  FA_CUDA(fa::embedding_bwd(idx, dy, dw, n_idx, dim, vocab, FA_ST), 1);
  return FA_OK;
}
int fa_b200_embedding_bwd_accum(const long long* idx, const float* dy, float* dw, size_t n_idx, int dim, size_t vocab, void* stream) {
  if (!idx || !dy || !dw) return fail(FA_ERR_BAD_ARG, "fa_b200_embedding_bwd_accum: null pointer");
  FA_CUDA(fa::embedding_bwd(idx, dy, dw, n_idx, dim, vocab, FA_ST, true), 1);
  return FA_OK;
}
int fa_b200_embed_pos_dropout_fwd(const long long* idx, const float* tok_w, const float* pos_w, float* y, size_t n_idx, int T, int dim,
                                  float p_drop, unsigned long long seed, const unsigned long long* seed_offset_dev, void* stream) {
  if (!idx || !tok_w || !pos_w || !y) return fail(FA_ERR_BAD_ARG, "fa_b200_embed_pos_dropout_fwd: null pointer");
  if (int rc = check_p_drop(p_drop, "fa_b200_embed_pos_dropout_fwd")) return rc;
  FA_CUDA(fa::embed_pos_dropout_fwd(idx, tok_w, pos_w, y, n_idx, T, dim, fa::make_dropout(p_drop, seed), seed_offset_dev, FA_ST), 1);
  return FA_OK;
}
int fa_b200_embed_pos_dropout_bwd(const long long* idx, const float* dy, float* dtok, float* dpos, size_t n_idx, int T, int dim,
                                  size_t tok_rows, size_t pos_rows, float p_drop, unsigned long long seed,
                                  const unsigned long long* seed_offset_dev, int accumulate, void* stream) {
  if (!idx || !dy || !dtok || !dpos) return fail(FA_ERR_BAD_ARG, "fa_b200_embed_pos_dropout_bwd: null pointer");
  if (int rc = check_p_drop(p_drop, "fa_b200_embed_pos_dropout_bwd")) return rc;
  FA_CUDA(fa::embed_pos_dropout_bwd(idx, dy, dtok, dpos, n_idx, T, dim, tok_rows, pos_rows, fa::make_dropout(p_drop, seed),
                                    seed_offset_dev, accumulate != 0, FA_ST), 1);
  return FA_OK;
}
int fa_b200_softmax_loss_fwd(const float* logits, const long long* target, float* loss, float* lse, int rows, int n_classes,
                             void* stream) {
  if (!logits || !target || !loss || !lse) return fail(FA_ERR_BAD_ARG, "fa_b200_softmax_loss_fwd: null pointer");
  FA_CUDA(fa::softmax_loss_fwd(logits, target, loss, lse, rows, n_classes, FA_ST), 1);
  return FA_OK;
}
int fa_b200_softmax_loss_bwd(const float* logits, const long long* target, const float* lse, const float* gscale,
                             float* dlogits, int rows, int n_classes, void* stream) {
  if (!logits || !target || !lse || !gscale || !dlogits) return fail(FA_ERR_BAD_ARG, "fa_b200_softmax_loss_bwd: null pointer");
  FA_CUDA(fa::softmax_loss_bwd(logits, target, lse, gscale, dlogits, rows, n_classes, FA_ST), 1);
  return FA_OK;
}
int fa_b200_softmax_loss_bwd_colsum(const float* logits, const long long* target, const float* lse, const float* gscale,
                                    float* dlogits, float* dcol, int rows, int n_classes, int accumulate, void* stream) {
  if (!logits || !target || !lse || !gscale || !dlogits || !dcol) return fail(FA_ERR_BAD_ARG, "fa_b200_softmax_loss_bwd_colsum: null pointer");
  if (!fa::softmax_loss_colsum_width(n_classes)) return fail(FA_ERR_UNSUPPORTED, "fa_b200_softmax_loss_bwd_colsum: n_classes must be a multiple of 4, at most 49152");
  FA_CUDA(fa::softmax_loss_bwd_colsum(logits, target, lse, gscale, dlogits, dcol, rows, n_classes, accumulate != 0, FA_ST), 1);
  return FA_OK;
}
int fa_b200_softmax_loss_colsum_width(int n_classes) { return fa::softmax_loss_colsum_width(n_classes) ? 1 : 0; }
int fa_b200_adam_step(float* p, const float* g, float* m, float* v, size_t n, float step_size, float beta1, float beta2,
                      float v_coef, float eps, void* stream) {
  if (!p || !g || !m || !v) return fail(FA_ERR_BAD_ARG, "fa_b200_adam_step: null pointer");
  FA_CUDA(fa::adam_step(p, g, m, v, n, step_size, beta1, beta2, v_coef, eps, FA_ST), 1);
  return FA_OK;
}
#undef FA_ST

// developer hook (not declared in fa_b200.h): clock64 timeline of one backward CTA, see tools/fa_check.cu
void fa_b200_debug_bwd_trace(void* dev_buf) { fa::tc_bwd_set_trace(static_cast<long long*>(dev_buf)); }
void fa_b200_debug_bwd_trace_start(int t0) { fa::tc_bwd_set_trace_start(t0); }
void fa_b200_debug_fwd_trace(void* dev_buf) { fa::tc_fwd_set_trace(static_cast<long long*>(dev_buf)); }

void fa_b200_set_host_precision(int precision) { g_host_precision = precision; }
void fa_b200_set_host_zero_qgrad(int on) { g_host_zero_qgrad = on ? 1 : 0; }
void fa_b200_set_fwd_pair(int on) { fa::set_fwd_pair(on != 0); }
int fa_b200_get_fwd_pair(void) { return fa::fwd_pair_enabled() ? 1 : 0; }
void fa_b200_set_fwd_persistent(int on) { fa::tc_fwd_set_persistent(on); }
int fa_b200_get_fwd_persistent(void) { return fa::tc_fwd_get_persistent(); }
void fa_b200_set_bwd_persistent(int on) { fa::tc_bwd_set_persistent(on); }
int fa_b200_get_bwd_persistent(void) { return fa::tc_bwd_get_persistent(); }
int fa_b200_get_host_precision(void) { return host_precision(); }
void fa_b200_release_arena(void) {
  std::lock_guard<std::mutex> lk(g_arena_mu);
  if (g_arena.base) cudaFree(g_arena.base);
  g_arena = Arena();
  g_pipe.destroy();
}

// ---- host-pointer API -------------------------------------------------------------------------
// The batch*heads axis is cut into chunks of whole heads (independent problems) that flow through a three-stage
// pipeline on three streams -- H2D (copy-in stream), kernels (the caller's stream), D2H (copy-out stream) -- over
// kPipeBufs rotating device buffers, so PCIe runs full duplex and the kernels hide under the copies.  The reference
// launchers move everything in, run, move everything out (src/flash_attn2_fw.cu:331-352).
int fa_b200_host_fwd(const float* q, const float* k, const float* v, float* out, float* l, float* m, int batch,
                     int N, int d, int causal, int convention, void* stream) {
  if (!q || !k || !v || !out || !l || !m) return fail(FA_ERR_BAD_ARG, "fa_b200_host_fwd: null pointer");
  if (!valid_shape(batch, N, d)) return fail(FA_ERR_BAD_ARG, "fa_b200_host_fwd: bad shape batch=%d N=%d d=%d", batch, N, d);
  if (convention != FA_CONV_FA1 && convention != FA_CONV_FA2)
    return fail(FA_ERR_BAD_ARG, "fa_b200_host_fwd: unknown convention %d", convention);
  std::lock_guard<std::mutex> lk(g_arena_mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int prec = resolve_precision(FA_DTYPE_F32, host_effective_precision(), d, false);
  if (prec < 0) return fail(FA_ERR_BAD_ARG, "fa_b200_host_fwd: bad precision setting %d", host_precision());
  const bool tc = prec == FA_PREC_BF16_TC;
  const bool fa1 = convention == FA_CONV_FA1;
  const size_t per_head = (size_t)N * d;
  const int hc = chunk_heads(batch, per_head);
  const int nchunks = (batch + hc - 1) / hc;
  const int nbuf = nchunks < kPipeBufs ? nchunks : kPipeBufs;
  const size_t cn = (size_t)hc * per_head, crows = (size_t)hc * N;
  const size_t buf_bytes = 4 * align_up(cn * 4, 256) + 3 * align_up(crows * 4, 256) + (tc ? 3 * align_up(cn * 2, 256) : 0);
  if (int rc = arena_reset(nbuf * buf_bytes + 4096)) return rc;
  if (int rc = g_pipe.ensure()) return rc;
  struct Buf {
    float *q, *k, *v, *o, *lse, *m, *l;
    __nv_bfloat16 *bq, *bk, *bv;
  } bufs[kPipeBufs];
  for (int b = 0; b < nbuf; ++b) {
    Buf& B = bufs[b];
    B.q = arena_take<float>(cn); B.k = arena_take<float>(cn); B.v = arena_take<float>(cn); B.o = arena_take<float>(cn);
    B.lse = arena_take<float>(crows); B.m = arena_take<float>(crows); B.l = arena_take<float>(crows);
    B.bq = B.bk = B.bv = nullptr;
    if (tc) { B.bq = arena_take<__nv_bfloat16>(cn); B.bk = arena_take<__nv_bfloat16>(cn); B.bv = arena_take<__nv_bfloat16>(cn); }
  }
  for (int c = 0; c < nchunks; ++c) {
    const int b = c % nbuf, h0 = c * hc, hn = (batch - h0 < hc) ? batch - h0 : hc;
    const size_t off = (size_t)h0 * per_head, roff = (size_t)h0 * N, n = (size_t)hn * per_head, rows = (size_t)hn * N;
    Buf& B = bufs[b];
    if (c >= nbuf) FA_CUDA(cudaStreamWaitEvent(g_pipe.s_in, g_pipe.ev_free[b], 0), 0);  // previous tenant copied out
    FA_CUDA(cudaMemcpyAsync(B.q, q + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.k, k + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.v, v + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaEventRecord(g_pipe.ev_in[b], g_pipe.s_in), 0);
    FA_CUDA(cudaStreamWaitEvent(st, g_pipe.ev_in[b], 0), 0);
    if (tc) {
      FA_CUDA(fa::cvt_f32_to_bf16(B.q, B.bq, n, st), 1);
      FA_CUDA(fa::cvt_f32_to_bf16(B.k, B.bk, n, st), 1);
      FA_CUDA(fa::cvt_f32_to_bf16(B.v, B.bv, n, st), 1);
      FA_CUDA(fa::tc_fwd(B.bq, B.bk, B.bv, B.o, true, B.lse, fa1 ? B.m : nullptr, fa1 ? B.l : nullptr, hn, N, d, causal != 0, st), 1);
    } else if (prec == FA_PREC_TF32_TC) {
      FA_CUDA(fa::tc_fwd_tf32(B.q, B.k, B.v, B.o, B.lse, fa1 ? B.m : nullptr, fa1 ? B.l : nullptr, hn, N, d, causal != 0, st), 1);
    } else {
      FA_CUDA(fa::simt_fwd(B.q, B.k, B.v, B.o, B.lse, fa1 ? B.m : nullptr, fa1 ? B.l : nullptr, hn, N, d, causal != 0, st), 1);
    }
    FA_CUDA(cudaEventRecord(g_pipe.ev_done[b], st), 0);
    FA_CUDA(cudaStreamWaitEvent(g_pipe.s_out, g_pipe.ev_done[b], 0), 0);
    FA_CUDA(cudaMemcpyAsync(out + off, B.o, n * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
    if (fa1) {
      FA_CUDA(cudaMemcpyAsync(l + roff, B.l, rows * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
      FA_CUDA(cudaMemcpyAsync(m + roff, B.m, rows * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
    } else {
      // FA-2 libraries return the logsumexp in `l` and never write `m` (src/flash_attn2_fw.cu:283-289)
      FA_CUDA(cudaMemcpyAsync(l + roff, B.lse, rows * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
    }
    FA_CUDA(cudaEventRecord(g_pipe.ev_free[b], g_pipe.s_out), 0);
  }
  // the caller's stream "owns" the result: later work on it sees the outputs, and the call returns complete
  FA_CUDA(cudaStreamWaitEvent(st, g_pipe.ev_free[(nchunks - 1) % nbuf], 0), 0);
  FA_CUDA(cudaStreamSynchronize(g_pipe.s_out), 0);
  FA_CUDA(cudaStreamSynchronize(st), 0);
  return FA_OK;
}

int fa_b200_host_bwd(const float* q, const float* k, const float* v, const float* out, const float* out_grad,
                     float* q_grad, float* k_grad, float* v_grad, const float* l, const float* m, int batch, int N,
                     int d, int causal, int convention, void* stream) {
  if (!q || !k || !v || !out || !out_grad || !q_grad || !k_grad || !v_grad || !l || !m)
    return fail(FA_ERR_BAD_ARG, "fa_b200_host_bwd: null pointer");
  if (!valid_shape(batch, N, d)) return fail(FA_ERR_BAD_ARG, "fa_b200_host_bwd: bad shape batch=%d N=%d d=%d", batch, N, d);
  if (convention != FA_CONV_FA1 && convention != FA_CONV_FA2)
    return fail(FA_ERR_BAD_ARG, "fa_b200_host_bwd: unknown convention %d", convention);
  std::lock_guard<std::mutex> lk(g_arena_mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int prec = resolve_precision(FA_DTYPE_F32, host_effective_precision(), d, true);
  if (prec < 0) return fail(FA_ERR_BAD_ARG, "fa_b200_host_bwd: bad precision setting %d", host_precision());
  const bool tc = prec == FA_PREC_BF16_TC;
  const bool fa1 = convention == FA_CONV_FA1;
  const size_t per_head = (size_t)N * d;
  const int hc = chunk_heads(batch, per_head);
  const int nchunks = (batch + hc - 1) / hc;
  const int nbuf = nchunks < kPipeBufs ? nchunks : kPipeBufs;
  const size_t cn = (size_t)hc * per_head, crows = (size_t)hc * N;
  const bool tf32 = prec == FA_PREC_TF32_TC;
  const bool zero_qgrad = host_zero_qgrad();
  const size_t ws_bytes = tc ? fa::tc_bwd_workspace_bytes(hc, N, d) : tf32 ? fa::tc_bwd_tf32_workspace_bytes(hc, N, d) : 0;
  const size_t buf_bytes = 8 * align_up(cn * 4, 256) + 4 * align_up(crows * 4, 256) + (tc ? 5 * align_up(cn * 2, 256) : 0) +
                           align_up(ws_bytes, 256) + 256;
  if (int rc = arena_reset(nbuf * buf_bytes + 4096)) return rc;
  if (int rc = g_pipe.ensure()) return rc;
  struct Buf {
    float *q, *k, *v, *o, *g, *gq, *gk, *gv, *l, *m, *lse, *delta;
    __nv_bfloat16* b[5];
    void* ws;
  } bufs[kPipeBufs];
  for (int i = 0; i < nbuf; ++i) {
    Buf& B = bufs[i];
    B.q = arena_take<float>(cn); B.k = arena_take<float>(cn); B.v = arena_take<float>(cn); B.o = arena_take<float>(cn);
    B.g = arena_take<float>(cn); B.gq = arena_take<float>(cn); B.gk = arena_take<float>(cn); B.gv = arena_take<float>(cn);
    B.l = arena_take<float>(crows); B.m = arena_take<float>(crows); B.lse = arena_take<float>(crows);
    B.delta = arena_take<float>(crows);
    for (int j = 0; j < 5; ++j) B.b[j] = tc ? arena_take<__nv_bfloat16>(cn) : nullptr;
    B.ws = ws_bytes ? arena_take<char>(ws_bytes) : nullptr;
  }
  for (int c = 0; c < nchunks; ++c) {
    const int b = c % nbuf, h0 = c * hc, hn = (batch - h0 < hc) ? batch - h0 : hc;
    const size_t off = (size_t)h0 * per_head, roff = (size_t)h0 * N, n = (size_t)hn * per_head, rows = (size_t)hn * N;
    Buf& B = bufs[b];
    if (c >= nbuf) FA_CUDA(cudaStreamWaitEvent(g_pipe.s_in, g_pipe.ev_free[b], 0), 0);
    FA_CUDA(cudaMemcpyAsync(B.q, q + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.k, k + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.v, v + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.o, out + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.g, out_grad + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    // the reference accumulates into whatever the caller put in q_grad (src/flash_attn2_bw.cu:228)
    if (!zero_qgrad) FA_CUDA(cudaMemcpyAsync(B.gq, q_grad + off, n * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaMemcpyAsync(B.l, l + roff, rows * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    if (fa1) FA_CUDA(cudaMemcpyAsync(B.m, m + roff, rows * 4, cudaMemcpyHostToDevice, g_pipe.s_in), 0);
    FA_CUDA(cudaEventRecord(g_pipe.ev_in[b], g_pipe.s_in), 0);
    FA_CUDA(cudaStreamWaitEvent(st, g_pipe.ev_in[b], 0), 0);
    const float* lse = B.l;
    if (fa1) {
      FA_CUDA(fa::lse_from_lm(B.l, B.m, B.lse, rows, st), 1);
      lse = B.lse;
    }
    if (tc) {
      const float* src[5] = {B.q, B.k, B.v, B.o, B.g};
      for (int j = 0; j < 5; ++j) FA_CUDA(fa::cvt_f32_to_bf16(src[j], B.b[j], n, st), 1);
      FA_CUDA(fa::tc_bwd(B.b[0], B.b[1], B.b[2], B.b[3], B.b[4], lse, B.gq, B.gk, B.gv, true, !zero_qgrad, B.ws, hn, N, d,
                         causal != 0, st), 2);
    } else if (tf32) {
      FA_CUDA(fa::tc_bwd_tf32(B.q, B.k, B.v, B.o, B.g, lse, B.gq, B.gk, B.gv, !zero_qgrad, B.ws, hn, N, d, causal != 0, st), 2);
    } else {
      FA_CUDA(fa::simt_bwd_prep(B.o, B.g, B.delta, nullptr, nullptr, nullptr, rows, d, st), 1);
      FA_CUDA(fa::simt_bwd(B.q, B.k, B.v, B.g, lse, B.delta, B.gq, B.gk, B.gv, hn, N, d, causal != 0, !zero_qgrad, st), 2);
    }
    FA_CUDA(cudaEventRecord(g_pipe.ev_done[b], st), 0);
    FA_CUDA(cudaStreamWaitEvent(g_pipe.s_out, g_pipe.ev_done[b], 0), 0);
    FA_CUDA(cudaMemcpyAsync(q_grad + off, B.gq, n * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
    FA_CUDA(cudaMemcpyAsync(k_grad + off, B.gk, n * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
    FA_CUDA(cudaMemcpyAsync(v_grad + off, B.gv, n * 4, cudaMemcpyDeviceToHost, g_pipe.s_out), 0);
    FA_CUDA(cudaEventRecord(g_pipe.ev_free[b], g_pipe.s_out), 0);
  }
  FA_CUDA(cudaStreamWaitEvent(st, g_pipe.ev_free[(nchunks - 1) % nbuf], 0), 0);
  FA_CUDA(cudaStreamSynchronize(g_pipe.s_out), 0);
  FA_CUDA(cudaStreamSynchronize(st), 0);
  return FA_OK;
}

}  // extern "C"
