// fa_util.cu -- small HBM-bound helper kernels: storage conversion and (l, m) -> logsumexp.
// Grid sized as a multiple of the SM count; 16-byte vector accesses.
#include "fa_internal.h"

namespace fa {

namespace {

__global__ void cvt_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    __nv_bfloat162 r0 = __floats2bfloat162_rn(a.x, a.y), r1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 r2 = __floats2bfloat162_rn(b.x, b.y), r3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&r0);
    o.y = *reinterpret_cast<uint32_t*>(&r1);
    o.z = *reinterpret_cast<uint32_t*>(&r2);
    o.w = *reinterpret_cast<uint32_t*>(&r3);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void cvt_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 a = reinterpret_cast<const uint4*>(src)[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
    float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(f0.x, f0.y, f1.x, f1.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
  }
  for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __bfloat162float(src[i]);
}

__global__ void lse_from_lm_kernel(const float* __restrict__ l, const float* __restrict__ m, float* __restrict__ lse,
                                   size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lse[i] = m[i] + logf(l[i]);
}

// keep[bh][i][j] for the whole (T x T) square of every head (causal or not: the mask does not depend on it)
__global__ void dropout_mask_kernel(unsigned char* __restrict__ keep, int head_begin, int T, Dropout drop) {
  const uint32_t bh = blockIdx.y;
  const uint32_t hkey = drop_head_key(drop.seed_lo, drop.seed_hi, head_begin + bh);
  const size_t n = (size_t)T * T;
  unsigned char* out = keep + (size_t)bh * n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const uint32_t i = (uint32_t)(idx / T), j = (uint32_t)(idx - (size_t)i * T);
    out[idx] = drop_keep(hkey, i, j, drop.thr) ? 1 : 0;
  }
}

inline unsigned grid_for(size_t work_items, int threads) {
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

cudaError_t cvt_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t st) {
  cvt_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t cvt_bf16_to_f32(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t st) {
  cvt_bf16_f32_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t lse_from_lm(const float* l, const float* m, float* lse, size_t n, cudaStream_t st) {
  lse_from_lm_kernel<<<grid_for(n, 256), 256, 0, st>>>(l, m, lse, n);
  return cudaGetLastError();
}

cudaError_t dropout_mask(unsigned char* keep, int head_begin, int n_heads, int T, Dropout drop, cudaStream_t st) {
  dim3 grid(grid_for((size_t)T * T, 256), n_heads);
  dropout_mask_kernel<<<grid, 256, 0, st>>>(keep, head_begin, T, drop);
  return cudaGetLastError();
}

}  // namespace fa
