// fa_lm_kernels.cu -- HBM-bound operators of the decoder-LM slice: LayerNorm,
// GELU, dropout, embedding, softmax cross-entropy and the optimizer step, device-resident and fused per operator.
//
// What they replace in the reference (each of which stages through host numpy storage, one elementwise
// map/zip/reduce launch per arithmetic op, minitorch/cuda_kernel_ops.py:60-200 over src/combine.cu:150-734):
//   LayerNorm1d.forward     mean, (x-mean)^2 mean, +eps, ^0.5, /, *w, +b  (8 launches)   minitorch/modules_basic.py:166-175
//   GELU                    tanh approximation (9 launches)                              minitorch/nn.py:208-214
//   Dropout.forward         host np.random.binomial mask, 2 launches                     minitorch/modules_basic.py:77-80
//   Embedding.forward       one_hot (host) @ weights                                     minitorch/modules_basic.py:46-53
//   softmax_loss            logsumexp - sum(one_hot * logits)                            minitorch/nn.py:217-274
//   Adam.step               ~12 launches per parameter                                   minitorch/optim.py:44-76
// All are memory-bound: one pass over the data, 16-byte accesses where the layout allows, grids sized in multiples
// of the SM count.  Row reductions use one warp (LayerNorm, dim <= 8192) or one CTA (loss, any class count) per row.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fa_dropout.cuh"
#include "fa_internal.h"

namespace fa {

namespace {

constexpr int kSMs = 148;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// ---------------------------------------------------------------- LayerNorm (one warp per row)
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                     float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                     int rows, int dim, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * dim;
  float s = 0.f;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / dim;
  float q = 0.f;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a = v.x - mean, bb = v.y - mean, cc = v.z - mean, d = v.w - mean;
    q += (a * a + bb * bb) + (cc * cc + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / dim + eps);  // biased variance, as ((x - mean) ** 2).mean()
  float* yr = y + (size_t)row * dim;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    const float4 bv = *reinterpret_cast<const float4*>(b + c);
    *reinterpret_cast<float4*>(yr + c) = make_float4(ww.x * (v.x - mean) * rstd + bv.x, ww.y * (v.y - mean) * rstd + bv.y,
                                                     ww.z * (v.z - mean) * rstd + bv.z, ww.w * (v.w - mean) * rstd + bv.w);
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// dx = rstd * (w dy - mean(w dy) - xhat mean(w dy xhat));  one warp per row
__global__ void layernorm_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ dy,
                                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                        float* __restrict__ dx, int rows, int dim) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * dim;
  const float* gr = dy + (size_t)row * dim;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(gr + c);
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    const float g0 = g.x * ww.x, g1 = g.y * ww.y, g2 = g.z * ww.z, g3 = g.w * ww.w;
    s1 += (g0 + g1) + (g2 + g3);
    s2 += (g0 * (v.x - mean) + g1 * (v.y - mean) + g2 * (v.z - mean) + g3 * (v.w - mean)) * rstd;
  }
  s1 = warp_sum(s1) / dim;
  s2 = warp_sum(s2) / dim;
  float* dr = dx + (size_t)row * dim;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(gr + c);
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    *reinterpret_cast<float4*>(dr + c) = make_float4(rstd * (g.x * ww.x - s1 - (v.x - mean) * rstd * s2),
                                                     rstd * (g.y * ww.y - s1 - (v.y - mean) * rstd * s2),
                                                     rstd * (g.z * ww.z - s1 - (v.z - mean) * rstd * s2),
                                                     rstd * (g.w * ww.w - s1 - (v.w - mean) * rstd * s2));
  }
}

// dw[c] += sum_rows dy xhat, db[c] += sum_rows dy: thread = column, block = strip of rows, one atomic per column per block
__global__ void layernorm_bwd_dwdb_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                          const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                          float* __restrict__ dw, float* __restrict__ db, int rows, int dim, int rows_per_block) {
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < dim; c += gridDim.x * blockDim.x) {
    float aw = 0.f, ab = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float g = dy[(size_t)r * dim + c];
      aw += g * (x[(size_t)r * dim + c] - mean_in[r]) * rstd_in[r];
      ab += g;
    }
    atomicAdd(dw + c, aw);
    atomicAdd(db + c, ab);
  }
}

// Register-resident variants for dim = 128 * kVec (the model's 256): every lane keeps its kVec float4s of the row in
// registers, so the row is read once.
//   forward : s = a (+ b) -- the residual sum of the layer before, written out when b is given -- and y = LN(s)
//   backward: dx = LN'(dy) (+ res, the gradient that reaches the same tensor along the residual branch), and the CTA's
//             share of dw / db reduced through shared memory, one atomic per column per CTA: ONE launch instead of
//             dx kernel + dw/db kernel (+ autograd's add of the two branch gradients)
template <int kVec>
__global__ void __launch_bounds__(256) add_layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                float* __restrict__ sum_out, float* __restrict__ y,
                                                                float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                int rows, float eps) {
  constexpr int dim = 128 * kVec;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t base = (size_t)row * dim + lane * 4;
  float4 v[kVec];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    v[j] = *reinterpret_cast<const float4*>(a + base + j * 128);
    if (b != nullptr) {
      const float4 u = *reinterpret_cast<const float4*>(b + base + j * 128);
      v[j] = make_float4(v[j].x + u.x, v[j].y + u.y, v[j].z + u.z, v[j].w + u.w);
      *reinterpret_cast<float4*>(sum_out + base + j * 128) = v[j];
    }
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) / dim;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const float c0 = v[j].x - mean, c1 = v[j].y - mean, c2 = v[j].z - mean, c3 = v[j].w - mean;
    q += (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
  }
  const float rstd = rsqrtf(warp_sum(q) / dim + eps);
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const float4 ww = *reinterpret_cast<const float4*>(w + lane * 4 + j * 128);
    const float4 bv = *reinterpret_cast<const float4*>(bias + lane * 4 + j * 128);
    *reinterpret_cast<float4*>(y + base + j * 128) =
        make_float4(ww.x * (v[j].x - mean) * rstd + bv.x, ww.y * (v[j].y - mean) * rstd + bv.y,
                    ww.z * (v[j].z - mean) * rstd + bv.z, ww.w * (v[j].w - mean) * rstd + bv.w);
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

template <int kVec>
__global__ void __launch_bounds__(256) layernorm_bwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ dy, const float* __restrict__ mean_in,
                                                                  const float* __restrict__ rstd_in, const float* __restrict__ res,
                                                                  float* __restrict__ dx, float* __restrict__ dw,
                                                                  float* __restrict__ db, float* __restrict__ dcol, int rows) {
  // dcol (optional): += column sums of dx -- the bias gradient of the linear layer whose output was the `b` of the
  // forward sum (its dy IS this dx), saving that layer's separate bias reduction pass over the same tensor
  constexpr int dim = 128 * kVec;
  __shared__ float4 red[3][8][32 * kVec];  // [dw | db | dcol][warp][lane's float4s]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = gridDim.x * 8;
  float4 ww[kVec], aw[kVec], ab[kVec], ac[kVec];
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    ww[j] = *reinterpret_cast<const float4*>(w + lane * 4 + j * 128);
    aw[j] = ab[j] = ac[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += stride) {
    const size_t base = (size_t)row * dim + lane * 4;
    float4 xv[kVec], gv[kVec], rv[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      xv[j] = *reinterpret_cast<const float4*>(x + base + j * 128);
      gv[j] = *reinterpret_cast<const float4*>(dy + base + j * 128);
      if (res != nullptr) rv[j] = *reinterpret_cast<const float4*>(res + base + j * 128);
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      // xv <- xhat;  gw = dy * w kept implicit (recomputed below: 4 FMULs are cheaper than 4 more live registers)
      xv[j] = make_float4((xv[j].x - mean) * rstd, (xv[j].y - mean) * rstd, (xv[j].z - mean) * rstd, (xv[j].w - mean) * rstd);
      const float g0 = gv[j].x * ww[j].x, g1 = gv[j].y * ww[j].y, g2 = gv[j].z * ww[j].z, g3 = gv[j].w * ww[j].w;
      s1 += (g0 + g1) + (g2 + g3);
      s2 += (g0 * xv[j].x + g1 * xv[j].y) + (g2 * xv[j].z + g3 * xv[j].w);
      aw[j] = make_float4(aw[j].x + gv[j].x * xv[j].x, aw[j].y + gv[j].y * xv[j].y, aw[j].z + gv[j].z * xv[j].z,
                          aw[j].w + gv[j].w * xv[j].w);
      ab[j] = make_float4(ab[j].x + gv[j].x, ab[j].y + gv[j].y, ab[j].z + gv[j].z, ab[j].w + gv[j].w);
    }
    s1 = warp_sum(s1) / dim;
    s2 = warp_sum(s2) / dim;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      float4 o = make_float4(rstd * (gv[j].x * ww[j].x - s1 - xv[j].x * s2), rstd * (gv[j].y * ww[j].y - s1 - xv[j].y * s2),
                             rstd * (gv[j].z * ww[j].z - s1 - xv[j].z * s2), rstd * (gv[j].w * ww[j].w - s1 - xv[j].w * s2));
      if (res != nullptr) o = make_float4(o.x + rv[j].x, o.y + rv[j].y, o.z + rv[j].z, o.w + rv[j].w);
      ac[j] = make_float4(ac[j].x + o.x, ac[j].y + o.y, ac[j].z + o.z, ac[j].w + o.w);
      *reinterpret_cast<float4*>(dx + base + j * 128) = o;
    }
  }
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    red[0][warp][j * 32 + lane] = aw[j];
    red[1][warp][j * 32 + lane] = ab[j];
    red[2][warp][j * 32 + lane] = ac[j];
  }
  __syncthreads();
  // 2 (3) * dim column sums over the 8 warps: thread t takes scalar column t, t + 256, ...
  const float* rf = reinterpret_cast<const float*>(red);
  for (int i = threadIdx.x; i < (dcol != nullptr ? 3 : 2) * dim; i += 256) {
    const int which = i / dim, c = i % dim;  // c = (j * 32 + lane) * 4 + e  <->  column j * 128 + lane * 4 + e: the same index
    float t = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) t += rf[(which * 8 + wp) * dim + c];
    atomicAdd((which == 0 ? dw : which == 1 ? db : dcol) + c, t);
  }
}

// ---------------------------------------------------------------- GELU (tanh approximation)
__device__ __forceinline__ float gelu_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
}
// mode 0: y = drop(gelu(x));  mode 1: y = drop(dy) * gelu'(x).  drop = the dropout_kernel mask below (same key, same
// element index), or the identity when kDrop is false: FeedForward's gelu -> dropout pair as one pass each way.
template <bool kDrop>
__global__ void gelu_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ y, size_t n4,
                            int mode, uint32_t key, uint32_t thr, float scale, const unsigned long long* __restrict__ seed_offset) {
  if (kDrop && seed_offset) {
    const unsigned long long o = *seed_offset;
    key = drop_mix(key ^ drop_mix((uint32_t)o + 0x9E3779B9u * (uint32_t)(o >> 32)));
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 o;
    if (mode == 0) {
      o = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
    } else {
      const float4 g = reinterpret_cast<const float4*>(dy)[i];
      o = make_float4(g.x * gelu_grad_f(v.x), g.y * gelu_grad_f(v.y), g.z * gelu_grad_f(v.z), g.w * gelu_grad_f(v.w));
    }
    if (kDrop) {
      const uint32_t w = drop_keep_word(drop_mix(key ^ ((uint32_t)i * kDropRowMul) ^ ((uint32_t)(i >> 32) * kDropColMul)), thr);
      o = make_float4((w & 0x80u) ? o.x * scale : 0.f, (w & 0x8000u) ? o.y * scale : 0.f, (w & 0x800000u) ? o.z * scale : 0.f,
                      (w & 0x80000000u) ? o.w * scale : 0.f);
    }
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

// The backward of drop(gelu(x)) for x = (rows, 128 * kVec) that also reduces dcol[c] += sum_rows dx[.][c] -- the bias
// gradient of the linear layer that produced x (ff.linear_in).  One warp per row, lanes own columns (as the LayerNorm
// backward above): sums in registers, CTA reduction in shared memory, one atomic per column per CTA.
template <int kVec>
__global__ void __launch_bounds__(256) gelu_dropout_bwd_colsum_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                      float* __restrict__ dx, float* __restrict__ dcol, int rows,
                                                                      uint32_t key, uint32_t thr, float scale,
                                                                      const unsigned long long* __restrict__ seed_offset) {
  constexpr int dim = 128 * kVec;
  __shared__ float4 red[8][32 * kVec];
  if (seed_offset) {
    const unsigned long long o = *seed_offset;
    key = drop_mix(key ^ drop_mix((uint32_t)o + 0x9E3779B9u * (uint32_t)(o >> 32)));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ac[kVec];
#pragma unroll
  for (int j = 0; j < kVec; ++j) ac[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float4 v[kVec], g[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const size_t i = (size_t)row * (dim / 4) + j * 32 + lane;
      v[j] = reinterpret_cast<const float4*>(x)[i];
      g[j] = reinterpret_cast<const float4*>(dy)[i];
    }
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const size_t i = (size_t)row * (dim / 4) + j * 32 + lane;  // the float4 index dropout_kernel hashes
      const uint32_t w = drop_keep_word(drop_mix(key ^ ((uint32_t)i * kDropRowMul) ^ ((uint32_t)(i >> 32) * kDropColMul)), thr);
      const float4 o = make_float4((w & 0x80u) ? g[j].x * gelu_grad_f(v[j].x) * scale : 0.f, (w & 0x8000u) ? g[j].y * gelu_grad_f(v[j].y) * scale : 0.f,
                                   (w & 0x800000u) ? g[j].z * gelu_grad_f(v[j].z) * scale : 0.f,
                                   (w & 0x80000000u) ? g[j].w * gelu_grad_f(v[j].w) * scale : 0.f);
      reinterpret_cast<float4*>(dx)[i] = o;
      ac[j] = make_float4(ac[j].x + o.x, ac[j].y + o.y, ac[j].z + o.z, ac[j].w + o.w);
    }
  }
#pragma unroll
  for (int j = 0; j < kVec; ++j) red[warp][j * 32 + lane] = ac[j];
  __syncthreads();
  const float* rf = reinterpret_cast<const float*>(red);
  for (int c = threadIdx.x; c < dim; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) t += rf[wp * dim + c];
    atomicAdd(dcol + c, t);
  }
}

// ---------------------------------------------------------------- dropout: y = x * keep / (1 - p), keep = f(seed, index)
// The same call serves forward and backward (the op is linear in x): nothing is stored between them.
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n4, uint32_t key, uint32_t thr,
                               float scale, const unsigned long long* __restrict__ seed_offset) {
  // seed_offset (optional, device memory): a step counter mixed into the key, so that a captured CUDA graph draws a
  // fresh mask on every replay without any host-side parameter change
  if (seed_offset) {
    const unsigned long long o = *seed_offset;
    key = drop_mix(key ^ drop_mix((uint32_t)o + 0x9E3779B9u * (uint32_t)(o >> 32)));
  }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    // one hash word per four elements: byte e (7 bits) >= thr keeps element e
    const uint32_t w = drop_keep_word(drop_mix(key ^ ((uint32_t)i * kDropRowMul) ^ ((uint32_t)(i >> 32) * kDropColMul)), thr);
    reinterpret_cast<float4*>(y)[i] = make_float4((w & 0x80u) ? v.x * scale : 0.f, (w & 0x8000u) ? v.y * scale : 0.f,
                                                  (w & 0x800000u) ? v.z * scale : 0.f, (w & 0x80000000u) ? v.w * scale : 0.f);
  }
}

// ---------------------------------------------------------------- embedding
__global__ void embedding_fwd_kernel(const long long* __restrict__ idx, const float* __restrict__ w, float* __restrict__ y,
                                     size_t n_idx, int dim4) {
  const size_t total = n_idx * dim4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / dim4;
    const int c = (int)(i - r * dim4);
    reinterpret_cast<float4*>(y)[i] = reinterpret_cast<const float4*>(w)[(size_t)idx[r] * dim4 + c];
  }
}
__global__ void embedding_bwd_kernel(const long long* __restrict__ idx, const float* __restrict__ dy, float* __restrict__ dw,
                                     size_t n_idx, int dim) {
  const size_t total = n_idx * dim, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / dim;
    const int c = (int)(i - r * dim);
    atomicAdd(dw + (size_t)idx[r] * dim + c, dy[i]);
  }
}

// The model's input stage as one pass each way (DecoderLM.forward, minitorch/modules_transfomer.py: token embedding +
// position embedding, then dropout): y[r] = drop(tok[idx[r]] + pos[r % T]);  backward: g = drop(dy[r]) scattered onto
// dtok[idx[r]] and dpos[r % T].  Same mask as dropout_kernel on the flattened (rows, dim) tensor (thr = 0: identity).
__global__ void embed_pos_dropout_fwd_kernel(const long long* __restrict__ idx, const float* __restrict__ tok,
                                             const float* __restrict__ pos, float* __restrict__ y, size_t n_idx, int T, int dim4,
                                             uint32_t key, uint32_t thr, float scale, const unsigned long long* __restrict__ seed_offset) {
  if (seed_offset) {
    const unsigned long long o = *seed_offset;
    key = drop_mix(key ^ drop_mix((uint32_t)o + 0x9E3779B9u * (uint32_t)(o >> 32)));
  }
  const size_t total = n_idx * dim4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / dim4;
    const int c = (int)(i - r * dim4);
    const float4 a = reinterpret_cast<const float4*>(tok)[(size_t)idx[r] * dim4 + c];
    const float4 b = reinterpret_cast<const float4*>(pos)[(size_t)(r % T) * dim4 + c];
    const uint32_t w = drop_keep_word(drop_mix(key ^ ((uint32_t)i * kDropRowMul) ^ ((uint32_t)(i >> 32) * kDropColMul)), thr);
    reinterpret_cast<float4*>(y)[i] =
        make_float4((w & 0x80u) ? (a.x + b.x) * scale : 0.f, (w & 0x8000u) ? (a.y + b.y) * scale : 0.f,
                    (w & 0x800000u) ? (a.z + b.z) * scale : 0.f, (w & 0x80000000u) ? (a.w + b.w) * scale : 0.f);
  }
}
__global__ void embed_pos_dropout_bwd_kernel(const long long* __restrict__ idx, const float* __restrict__ dy, float* __restrict__ dtok,
                                             float* __restrict__ dpos, size_t n_idx, int T, int dim4, uint32_t key, uint32_t thr,
                                             float scale, const unsigned long long* __restrict__ seed_offset) {
  if (seed_offset) {
    const unsigned long long o = *seed_offset;
    key = drop_mix(key ^ drop_mix((uint32_t)o + 0x9E3779B9u * (uint32_t)(o >> 32)));
  }
  const size_t total = n_idx * dim4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const size_t r = i / dim4;
    const int c = (int)(i - r * dim4);
    const float4 g = reinterpret_cast<const float4*>(dy)[i];
    const uint32_t w = drop_keep_word(drop_mix(key ^ ((uint32_t)i * kDropRowMul) ^ ((uint32_t)(i >> 32) * kDropColMul)), thr);
    const float g0 = (w & 0x80u) ? g.x * scale : 0.f, g1 = (w & 0x8000u) ? g.y * scale : 0.f;
    const float g2 = (w & 0x800000u) ? g.z * scale : 0.f, g3 = (w & 0x80000000u) ? g.w * scale : 0.f;
    float* a = dtok + ((size_t)idx[r] * dim4 + c) * 4;
    float* b = dpos + ((size_t)(r % T) * dim4 + c) * 4;
    atomicAdd(a, g0); atomicAdd(a + 1, g1); atomicAdd(a + 2, g2); atomicAdd(a + 3, g3);
    atomicAdd(b, g0); atomicAdd(b + 1, g1); atomicAdd(b + 2, g2); atomicAdd(b + 3, g3);
  }
}

// ---------------------------------------------------------------- softmax cross-entropy (one CTA per row)
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}
// loss[i] = logsumexp(logits[i]) - logits[i][target[i]];  lse saved for the backward
__global__ void softmax_loss_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                        float* __restrict__ loss, float* __restrict__ lse_out, int C) {
  __shared__ float sh[32];
  const float* row = logits + (size_t)blockIdx.x * C;
  // one pass: every thread keeps a running (max, sum of exp relative to it); rescaled to the row maximum at the end
  float m = -INFINITY, s = 0.f;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (mx > m) {
      s *= __expf(m - mx);  // (m = -inf on the first visit: s = 0 stays 0)
      m = mx;
    }
    s += (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
  }
  const float mrow = block_reduce(m, sh, true);
  s = block_reduce(m == -INFINITY ? 0.f : s * __expf(m - mrow), sh, false);
  if (threadIdx.x == 0) {
    const float lse = mrow + logf(s);
    lse_out[blockIdx.x] = lse;
    loss[blockIdx.x] = lse - row[target[blockIdx.x]];
  }
}
// dlogits[i][j] = gscale[i] * (exp(logits[i][j] - lse[i]) - [j == target[i]])
__global__ void softmax_loss_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                        const float* __restrict__ lse, const float* __restrict__ gscale,
                                        float* __restrict__ dlogits, int C) {
  const float* row = logits + (size_t)blockIdx.x * C;
  float* out = dlogits + (size_t)blockIdx.x * C;
  const float L = lse[blockIdx.x], g = gscale[blockIdx.x];
  const int t = (int)target[blockIdx.x];
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    float4 o = make_float4(g * __expf(v.x - L), g * __expf(v.y - L), g * __expf(v.z - L), g * __expf(v.w - L));
    if (t >= c && t < c + 4) (&o.x)[t - c] -= g;
    *reinterpret_cast<float4*>(out + c) = o;
  }
}

// The same plus dcol[j] += sum_i dlogits[i][j] -- the bias gradient of the linear layer that produced the logits -- in
// the same pass: a CTA takes a strip of rows and a chunk of 4 * 256 * kLossQuads columns (blockIdx.y), thread t owns
// column quads t, t + 256, ... of the chunk, keeps their sums in registers and issues one atomic per column at the end.
// (6 quads per thread: ~70 registers, three CTAs per SM -- 12 quads / two CTAs left the loads idle during the exps.)
constexpr int kLossQuads = 6;
__global__ void __launch_bounds__(256, 3) softmax_loss_bwd_colsum_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                                      const float* __restrict__ lse, const float* __restrict__ gscale,
                                                                      float* __restrict__ dlogits, float* __restrict__ dcol, int rows,
                                                                      int C, int rows_per_block) {
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 acc[kLossQuads];
#pragma unroll
  for (int k = 0; k < kLossQuads; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < r1; ++r) {
    const float* row = logits + (size_t)r * C;
    float* out = dlogits + (size_t)r * C;
    const float L = lse[r], g = gscale[r];
    const int t = (int)target[r];
    float4 v[kLossQuads];  // the whole row slice first: kLossQuads independent 16-byte loads in flight per thread
#pragma unroll
    for (int k = 0; k < kLossQuads; ++k) {
      const int c = ((blockIdx.y * kLossQuads + k) * 256 + threadIdx.x) * 4;
      if (c < C) v[k] = __ldcs(reinterpret_cast<const float4*>(row + c));
    }
#pragma unroll
    for (int k = 0; k < kLossQuads; ++k) {
      const int c = ((blockIdx.y * kLossQuads + k) * 256 + threadIdx.x) * 4;
      if (c < C) {
        float4 o = make_float4(g * __expf(v[k].x - L), g * __expf(v[k].y - L), g * __expf(v[k].z - L), g * __expf(v[k].w - L));
        if (t >= c && t < c + 4) (&o.x)[t - c] -= g;
        *reinterpret_cast<float4*>(out + c) = o;
        acc[k] = make_float4(acc[k].x + o.x, acc[k].y + o.y, acc[k].z + o.z, acc[k].w + o.w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kLossQuads; ++k) {
    const int c = ((blockIdx.y * kLossQuads + k) * 256 + threadIdx.x) * 4;
    if (c < C) {
      atomicAdd(dcol + c, acc[k].x); atomicAdd(dcol + c + 1, acc[k].y);
      atomicAdd(dcol + c + 2, acc[k].z); atomicAdd(dcol + c + 3, acc[k].w);
    }
  }
}

// ---------------------------------------------------------------- Adam over one flat parameter buffer
// m = b1 m + (1-b1) g;  v = b2 v + (1-b1) g^2  (sic: the reference uses 1-beta1 here, minitorch/optim.py:68);
// p -= step_size * m / (sqrt(v) + eps),  step_size = lr sqrt(1-b2^t) / (1-b1^t) computed by the host.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n4, float step_size, float b1, float b2, float v_coef, float eps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define FA_ADAM1(c)                                    \
  mm.c = b1 * mm.c + (1.f - b1) * gg.c;                \
  vv.c = b2 * vv.c + v_coef * gg.c * gg.c;             \
  pp.c -= step_size * mm.c / (sqrtf(vv.c) + eps);
    FA_ADAM1(x) FA_ADAM1(y) FA_ADAM1(z) FA_ADAM1(w)
#undef FA_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

inline unsigned ew_grid(size_t items, int threads) {
  size_t b = (items + threads - 1) / threads;
  if (b > (size_t)kSMs * 16) b = (size_t)kSMs * 16;
  return (unsigned)(b ? b : 1);
}

}  // namespace

cudaError_t add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, float* sum, float* y,
                              float* mean, float* rstd, int rows, int dim, float eps, cudaStream_t st) {
  if (rows <= 0 || dim <= 0 || (dim & 3) || (b != nullptr && sum == nullptr)) return cudaErrorInvalidValue;
  const unsigned grid = (rows + 7) / 8;
  switch (dim) {
    case 128: add_layernorm_fwd_kernel<1><<<grid, 256, 0, st>>>(a, b, w, bias, sum, y, mean, rstd, rows, eps); break;
    case 256: add_layernorm_fwd_kernel<2><<<grid, 256, 0, st>>>(a, b, w, bias, sum, y, mean, rstd, rows, eps); break;
    case 512: add_layernorm_fwd_kernel<4><<<grid, 256, 0, st>>>(a, b, w, bias, sum, y, mean, rstd, rows, eps); break;
    default:
      if (b != nullptr) return cudaErrorInvalidValue;  // (callers add separately for other widths)
      layernorm_fwd_kernel<<<grid, 256, 0, st>>>(a, w, bias, y, mean, rstd, rows, dim, eps);
  }
  return cudaGetLastError();
}
cudaError_t layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd, int rows,
                          int dim, float eps, cudaStream_t st) {
  return add_layernorm_fwd(x, nullptr, w, b, nullptr, y, mean, rstd, rows, dim, eps, st);
}
bool layernorm_fused_width(int dim) { return dim == 128 || dim == 256 || dim == 512; }  // (1024: 64 KB of reduction smem)
cudaError_t layernorm_bwd(const float* x, const float* w, const float* dy, const float* mean, const float* rstd, float* dx,
                          float* dw, float* db, int rows, int dim, cudaStream_t st, bool accumulate, const float* res, float* dcol) {
  if (rows <= 0 || dim <= 0 || (dim & 3)) return cudaErrorInvalidValue;
  if ((res != nullptr || dcol != nullptr) && !layernorm_fused_width(dim)) return cudaErrorInvalidValue;
  cudaError_t e;
  if (!accumulate) {  // (accumulate: dw / db are a parameter's gradient buffer, the strips atomically add onto it)
    if ((e = cudaMemsetAsync(dw, 0, (size_t)dim * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(db, 0, (size_t)dim * 4, st)) != cudaSuccess) return e;
    if (dcol != nullptr && (e = cudaMemsetAsync(dcol, 0, (size_t)dim * 4, st)) != cudaSuccess) return e;
  }
  if (layernorm_fused_width(dim)) {
    // two rows per warp (5632 rows: 352 CTAs, all resident at 3 CTAs of 80 registers x 256 threads per SM, so no
    // second wave and half the atomics of one row per warp); more rows than 3 CTAs per SM can take loop
    unsigned grid = (rows + 15) / 16;
    if (grid > (unsigned)kSMs * 3) grid = kSMs * 3;
    switch (dim) {
      case 128: layernorm_bwd_fused_kernel<1><<<grid, 256, 0, st>>>(x, w, dy, mean, rstd, res, dx, dw, db, dcol, rows); break;
      case 256: layernorm_bwd_fused_kernel<2><<<grid, 256, 0, st>>>(x, w, dy, mean, rstd, res, dx, dw, db, dcol, rows); break;
      default: layernorm_bwd_fused_kernel<4><<<grid, 256, 0, st>>>(x, w, dy, mean, rstd, res, dx, dw, db, dcol, rows); break;
    }
    return cudaGetLastError();
  }
  layernorm_bwd_dx_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, w, dy, mean, rstd, dx, rows, dim);
  const int col_blocks = (dim + 255) / 256;
  int strips = (kSMs * 2) / col_blocks;
  if (strips < 1) strips = 1;
  if (strips > rows) strips = rows;
  const int rpb = (rows + strips - 1) / strips;
  layernorm_bwd_dwdb_kernel<<<dim3(col_blocks, (rows + rpb - 1) / rpb), 256, 0, st>>>(x, dy, mean, rstd, dw, db, rows, dim, rpb);
  return cudaGetLastError();
}
cudaError_t gelu(const float* x, const float* dy, float* y, size_t n, bool backward, cudaStream_t st) {
  if (n & 3) return cudaErrorInvalidValue;
  gelu_kernel<false><<<ew_grid(n / 4, 256), 256, 0, st>>>(x, dy, y, n / 4, backward ? 1 : 0, 0u, 0u, 1.f, nullptr);
  return cudaGetLastError();
}
cudaError_t gelu_dropout(const float* x, const float* dy, float* y, size_t n, bool backward, Dropout drop,
                         const unsigned long long* seed_offset, cudaStream_t st) {
  if (n & 3) return cudaErrorInvalidValue;
  const uint32_t key = drop_head_key(drop.seed_lo, drop.seed_hi, 0xE1E7u);  // the key of dropout_apply: same mask
  gelu_kernel<true><<<ew_grid(n / 4, 256), 256, 0, st>>>(x, dy, y, n / 4, backward ? 1 : 0, key, drop.thr, drop.keep_scale, seed_offset);
  return cudaGetLastError();
}
cudaError_t gelu_dropout_bwd_colsum(const float* x, const float* dy, float* dx, float* dcol, int rows, int dim, Dropout drop,
                                    const unsigned long long* seed_offset, bool accumulate, cudaStream_t st) {
  if (rows <= 0 || !layernorm_fused_width(dim)) return cudaErrorInvalidValue;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dcol, 0, (size_t)dim * 4, st);
    if (e != cudaSuccess) return e;
  }
  const uint32_t key = drop_head_key(drop.seed_lo, drop.seed_hi, 0xE1E7u);
  unsigned grid = (rows + 15) / 16;  // two rows per warp
  if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
  switch (dim) {
    case 128: gelu_dropout_bwd_colsum_kernel<1><<<grid, 256, 0, st>>>(x, dy, dx, dcol, rows, key, drop.thr, drop.keep_scale, seed_offset); break;
    case 256: gelu_dropout_bwd_colsum_kernel<2><<<grid, 256, 0, st>>>(x, dy, dx, dcol, rows, key, drop.thr, drop.keep_scale, seed_offset); break;
    default: gelu_dropout_bwd_colsum_kernel<4><<<grid, 256, 0, st>>>(x, dy, dx, dcol, rows, key, drop.thr, drop.keep_scale, seed_offset); break;
  }
  return cudaGetLastError();
}
cudaError_t dropout_apply(const float* x, float* y, size_t n, Dropout drop, const unsigned long long* seed_offset,
                          cudaStream_t st) {
  if (n & 3) return cudaErrorInvalidValue;
  const uint32_t key = drop_head_key(drop.seed_lo, drop.seed_hi, 0xE1E7u);
  dropout_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>(x, y, n / 4, key, drop.thr, drop.keep_scale, seed_offset);
  return cudaGetLastError();
}
cudaError_t embedding_fwd(const long long* idx, const float* w, float* y, size_t n_idx, int dim, cudaStream_t st) {
  if (dim <= 0 || (dim & 3)) return cudaErrorInvalidValue;
  embedding_fwd_kernel<<<ew_grid(n_idx * (dim / 4), 256), 256, 0, st>>>(idx, w, y, n_idx, dim / 4);
  return cudaGetLastError();
}
cudaError_t embedding_bwd(const long long* idx, const float* dy, float* dw, size_t n_idx, int dim, size_t vocab, cudaStream_t st,
                          bool accumulate) {
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dw, 0, vocab * dim * 4, st);
    if (e != cudaSuccess) return e;
  }
  embedding_bwd_kernel<<<ew_grid(n_idx * dim, 256), 256, 0, st>>>(idx, dy, dw, n_idx, dim);
  return cudaGetLastError();
}
cudaError_t embed_pos_dropout_fwd(const long long* idx, const float* tok, const float* pos, float* y, size_t n_idx, int T, int dim,
                                  Dropout drop, const unsigned long long* seed_offset, cudaStream_t st) {
  if (dim <= 0 || (dim & 3) || T <= 0) return cudaErrorInvalidValue;
  const uint32_t key = drop_head_key(drop.seed_lo, drop.seed_hi, 0xE1E7u);
  embed_pos_dropout_fwd_kernel<<<ew_grid(n_idx * (dim / 4), 256), 256, 0, st>>>(idx, tok, pos, y, n_idx, T, dim / 4, key, drop.thr,
                                                                                 drop.keep_scale, seed_offset);
  return cudaGetLastError();
}
cudaError_t embed_pos_dropout_bwd(const long long* idx, const float* dy, float* dtok, float* dpos, size_t n_idx, int T, int dim,
                                  size_t tok_rows, size_t pos_rows, Dropout drop, const unsigned long long* seed_offset,
                                  bool accumulate, cudaStream_t st) {
  if (dim <= 0 || (dim & 3) || T <= 0 || pos_rows < (size_t)T) return cudaErrorInvalidValue;
  if (!accumulate) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(dtok, 0, tok_rows * dim * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(dpos, 0, pos_rows * dim * 4, st)) != cudaSuccess) return e;
  }
  const uint32_t key = drop_head_key(drop.seed_lo, drop.seed_hi, 0xE1E7u);
  embed_pos_dropout_bwd_kernel<<<ew_grid(n_idx * (dim / 4), 256), 256, 0, st>>>(idx, dy, dtok, dpos, n_idx, T, dim / 4, key, drop.thr,
                                                                                 drop.keep_scale, seed_offset);
  return cudaGetLastError();
}
cudaError_t softmax_loss_fwd(const float* logits, const long long* target, float* loss, float* lse, int rows, int C,
                             cudaStream_t st) {
  if (rows <= 0 || C <= 0 || (C & 3)) return cudaErrorInvalidValue;
  softmax_loss_fwd_kernel<<<rows, 256, 0, st>>>(logits, target, loss, lse, C);
  return cudaGetLastError();
}
cudaError_t softmax_loss_bwd(const float* logits, const long long* target, const float* lse, const float* gscale,
                             float* dlogits, int rows, int C, cudaStream_t st) {
  if (rows <= 0 || C <= 0 || (C & 3)) return cudaErrorInvalidValue;
  softmax_loss_bwd_kernel<<<rows, 256, 0, st>>>(logits, target, lse, gscale, dlogits, C);
  return cudaGetLastError();
}
bool softmax_loss_colsum_width(int C) { return C > 0 && !(C & 3) && C <= 4 * 256 * kLossQuads * 8; }
cudaError_t softmax_loss_bwd_colsum(const float* logits, const long long* target, const float* lse, const float* gscale,
                                    float* dlogits, float* dcol, int rows, int C, bool accumulate, cudaStream_t st) {
  if (rows <= 0 || !softmax_loss_colsum_width(C)) return cudaErrorInvalidValue;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(dcol, 0, (size_t)C * 4, st);
    if (e != cudaSuccess) return e;
  }
  const int chunks = (C + 4 * 256 * kLossQuads - 1) / (4 * 256 * kLossQuads);
  int strips = (kSMs * 3 + chunks - 1) / chunks;  // three CTAs per SM, each thread with up to kLossQuads 16-byte loads in flight
  if (strips > rows) strips = rows;
  const int rpb = (rows + strips - 1) / strips;
  softmax_loss_bwd_colsum_kernel<<<dim3((rows + rpb - 1) / rpb, chunks), 256, 0, st>>>(logits, target, lse, gscale, dlogits, dcol, rows, C, rpb);
  return cudaGetLastError();
}
cudaError_t adam_step(float* p, const float* g, float* m, float* v, size_t n, float step_size, float b1, float b2,
                      float v_coef, float eps, cudaStream_t st) {
  if (n & 3) return cudaErrorInvalidValue;
  adam_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>(p, g, m, v, n / 4, step_size, b1, b2, v_coef, eps);
  return cudaGetLastError();
}

}  // namespace fa
