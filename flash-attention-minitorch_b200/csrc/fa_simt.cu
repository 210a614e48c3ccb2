// fa_simt.cu -- fp32 CUDA-core attention forward/backward for shapes the tensor-core kernels do
// not cover (head dims other than 32/64/128, e.g. the reference's d=34 test) and as an exact-fp32
// device-side cross-check of the tcgen05 kernels at full problem sizes.
//
// Math (reference src/flash_attn2_fw.cu:152-294, src/flash_attn2_bw.cu:143-245):
//   S = tau * Q K^T, causal: S_ij = -inf for j > i, P = exp(S - L), O = P V, L = m + log(l)
//   D_i = sum_k dO_ik O_ik, dV = P^T dO, dP = dO V^T, dS = P o (dP - D), dQ = tau dS K, dK = tau dS^T Q
//
// Thread mapping: each row (query row in fwd/dQ, key row in dK/dV) is owned by NSPLIT adjacent
// lanes; lane h of the group owns head-dim columns {c*NSPLIT + h}.  Dot products are finished
// with xor-shuffles, K/V (or Q/dO) tiles are staged in shared memory and read as broadcasts.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fa_internal.h"
#include "fa_dropout.cuh"

namespace fa {

namespace {

constexpr int kRowsPerBlock = 64;
constexpr int kTile = 32;  // rows of the streamed operand per shared-memory tile

template <int NSPLIT>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = 1; o < NSPLIT; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// cooperative load of `rows` rows x d cols (row stride d) into smem [kTile][DPAD], zero padded
template <int DPAD>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int row0, int T,
                                          int d, int nthreads) {
  for (int idx = threadIdx.x; idx < kTile * DPAD; idx += nthreads) {
    int r = idx / DPAD, c = idx - r * DPAD;
    int gr = row0 + r;
    dst[idx] = (gr < T && c < d) ? src[(size_t)gr * d + c] : 0.f;
  }
}

template <int DPAD, int NSPLIT>
__global__ void __launch_bounds__(kRowsPerBlock* NSPLIT)
simt_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                float* __restrict__ o, float* __restrict__ lse, float* __restrict__ m_out,
                float* __restrict__ l_out, int T, int d, int causal, float tau, Dropout drop) {
  constexpr int DC = DPAD / NSPLIT;
  constexpr int NT = kRowsPerBlock * NSPLIT;
  __shared__ float sK[kTile * DPAD];
  __shared__ float sV[kTile * DPAD];
  const int bh = blockIdx.y;
  const int h = threadIdx.x % NSPLIT;
  const int row = blockIdx.x * kRowsPerBlock + threadIdx.x / NSPLIT;
  const bool row_ok = row < T;
  const float* qb = q + (size_t)bh * T * d;
  const float* kb = k + (size_t)bh * T * d;
  const float* vb = v + (size_t)bh * T * d;

  float qr[DC], acc[DC];
#pragma unroll
  for (int c = 0; c < DC; ++c) {
    int col = c * NSPLIT + h;
    qr[c] = (row_ok && col < d) ? qb[(size_t)row * d + col] : 0.f;
    acc[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const int last_row = min(T, (int)(blockIdx.x + 1) * kRowsPerBlock) - 1;
  const int kv_end = causal ? min(T, last_row + 1) : T;
  const uint32_t hkey = drop_head_key(drop.seed_lo, drop.seed_hi, bh);

  for (int j0 = 0; j0 < kv_end; j0 += kTile) {
    __syncthreads();
    load_tile<DPAD>(sK, kb, j0, T, d, NT);
    load_tile<DPAD>(sV, vb, j0, T, d, NT);
    __syncthreads();
    float s[kTile];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < kTile; ++jj) {
      float part = 0.f;
#pragma unroll
      for (int c = 0; c < DC; ++c) part = fmaf(qr[c], sK[jj * DPAD + c * NSPLIT + h], part);
      float sv = group_sum<NSPLIT>(part) * tau;
      int col = j0 + jj;
      bool ok = col < T && (!causal || col <= row);
      s[jj] = ok ? sv : -INFINITY;
      tmax = fmaxf(tmax, s[jj]);
    }
    float m_new = fmaxf(m, tmax);
    if (m_new == -INFINITY) continue;  // nothing visible yet for this row (uniform within the lane group)
    float corr = expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int c = 0; c < DC; ++c) acc[c] *= corr;
#pragma unroll
    for (int jj = 0; jj < kTile; ++jj) {
      float p = expf(s[jj] - m_new);
      l += p;  // the softmax statistics are those of the undropped scores
      if (drop.thr && !drop_keep(hkey, row, j0 + jj, drop.thr)) p = 0.f;
#pragma unroll
      for (int c = 0; c < DC; ++c) acc[c] = fmaf(p, sV[jj * DPAD + c * NSPLIT + h], acc[c]);
    }
    m = m_new;
  }
  if (row_ok) {
    float inv = (drop.thr ? drop.keep_scale : 1.f) / l;
    float* ob = o + (size_t)bh * T * d + (size_t)row * d;
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      int col = c * NSPLIT + h;
      if (col < d) ob[col] = acc[c] * inv;
    }
    if (h == 0) {
      size_t idx = (size_t)bh * T + row;
      if (lse) lse[idx] = m + logf(l);
      if (m_out) m_out[idx] = m;
      if (l_out) l_out[idx] = l;
    }
  }
}

// delta_i = sum_k dO_ik * O_ik ; optionally lse_i = m_i + log(l_i) for the FA-1 (l, m) convention
__global__ void simt_bwd_prep_kernel(const float* __restrict__ o, const float* __restrict__ dO,
                                     float* __restrict__ delta, const float* __restrict__ l_in,
                                     const float* __restrict__ m_in, float* __restrict__ lse_out, size_t rows, int d) {
  size_t row = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int c = lane; c < d; c += 32) acc = fmaf(o[row * d + c], dO[row * d + c], acc);
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    delta[row] = acc;
    if (lse_out) lse_out[row] = m_in[row] + logf(l_in[row]);
  }
}

template <int DPAD, int NSPLIT>
__global__ void __launch_bounds__(kRowsPerBlock* NSPLIT)
simt_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                   const float* __restrict__ dO, const float* __restrict__ lse, const float* __restrict__ delta,
                   float* __restrict__ dq, int T, int d, int causal, float tau, int accumulate, Dropout drop) {
  constexpr int DC = DPAD / NSPLIT;
  constexpr int NT = kRowsPerBlock * NSPLIT;
  __shared__ float sK[kTile * DPAD];
  __shared__ float sV[kTile * DPAD];
  const int bh = blockIdx.y;
  const int h = threadIdx.x % NSPLIT;
  const int row = blockIdx.x * kRowsPerBlock + threadIdx.x / NSPLIT;
  const bool row_ok = row < T;
  const size_t base = (size_t)bh * T * d;
  float qr[DC], dor[DC], acc[DC];
#pragma unroll
  for (int c = 0; c < DC; ++c) {
    int col = c * NSPLIT + h;
    bool ok = row_ok && col < d;
    qr[c] = ok ? q[base + (size_t)row * d + col] : 0.f;
    dor[c] = ok ? dO[base + (size_t)row * d + col] : 0.f;
    acc[c] = 0.f;
  }
  const float L = row_ok ? lse[(size_t)bh * T + row] : 0.f;
  const float Dl = row_ok ? delta[(size_t)bh * T + row] : 0.f;
  const uint32_t hkey = drop_head_key(drop.seed_lo, drop.seed_hi, bh);
  const int last_row = min(T, (int)(blockIdx.x + 1) * kRowsPerBlock) - 1;
  const int kv_end = causal ? min(T, last_row + 1) : T;
  for (int j0 = 0; j0 < kv_end; j0 += kTile) {
    __syncthreads();
    load_tile<DPAD>(sK, k + base, j0, T, d, NT);
    load_tile<DPAD>(sV, v + base, j0, T, d, NT);
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < kTile; ++jj) {
      float ps = 0.f, pd = 0.f;
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        ps = fmaf(qr[c], sK[jj * DPAD + c * NSPLIT + h], ps);
        pd = fmaf(dor[c], sV[jj * DPAD + c * NSPLIT + h], pd);
      }
      float sv = group_sum<NSPLIT>(ps) * tau;
      float dp = group_sum<NSPLIT>(pd);
      int col = j0 + jj;
      bool ok = row_ok && col < T && (!causal || col <= row);
      float p = ok ? expf(sv - L) : 0.f;
      if (drop.thr) dp = drop_keep(hkey, row, col, drop.thr) ? dp * drop.keep_scale : 0.f;  // dP o Z
      float ds = p * (dp - Dl);
#pragma unroll
      for (int c = 0; c < DC; ++c) acc[c] = fmaf(ds, sK[jj * DPAD + c * NSPLIT + h], acc[c]);
    }
  }
  if (row_ok) {
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      int col = c * NSPLIT + h;
      if (col < d) {
        size_t idx = base + (size_t)row * d + col;
        dq[idx] = (accumulate ? dq[idx] : 0.f) + acc[c] * tau;
      }
    }
  }
}

template <int DPAD, int NSPLIT>
__global__ void __launch_bounds__(kRowsPerBlock* NSPLIT)
simt_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                    const float* __restrict__ dO, const float* __restrict__ lse, const float* __restrict__ delta,
                    float* __restrict__ dk, float* __restrict__ dv, int T, int d, int causal, float tau, Dropout drop) {
  constexpr int DC = DPAD / NSPLIT;
  constexpr int NT = kRowsPerBlock * NSPLIT;
  __shared__ float sQ[kTile * DPAD];
  __shared__ float sdO[kTile * DPAD];
  __shared__ float sL[kTile], sD[kTile];
  const int bh = blockIdx.y;
  const int h = threadIdx.x % NSPLIT;
  const int row = blockIdx.x * kRowsPerBlock + threadIdx.x / NSPLIT;  // key row
  const bool row_ok = row < T;
  const size_t base = (size_t)bh * T * d;
  float kr[DC], vr[DC], dkacc[DC], dvacc[DC];
#pragma unroll
  for (int c = 0; c < DC; ++c) {
    int col = c * NSPLIT + h;
    bool ok = row_ok && col < d;
    kr[c] = ok ? k[base + (size_t)row * d + col] : 0.f;
    vr[c] = ok ? v[base + (size_t)row * d + col] : 0.f;
    dkacc[c] = 0.f;
    dvacc[c] = 0.f;
  }
  const uint32_t hkey = drop_head_key(drop.seed_lo, drop.seed_hi, bh);
  const int first_key = blockIdx.x * kRowsPerBlock;
  const int i_begin = causal ? (first_key / kTile) * kTile : 0;
  for (int i0 = i_begin; i0 < T; i0 += kTile) {
    __syncthreads();
    load_tile<DPAD>(sQ, q + base, i0, T, d, NT);
    load_tile<DPAD>(sdO, dO + base, i0, T, d, NT);
    if (threadIdx.x < kTile) {
      int gi = i0 + threadIdx.x;
      sL[threadIdx.x] = gi < T ? lse[(size_t)bh * T + gi] : 0.f;
      sD[threadIdx.x] = gi < T ? delta[(size_t)bh * T + gi] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ii = 0; ii < kTile; ++ii) {
      float ps = 0.f, pd = 0.f;
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        ps = fmaf(kr[c], sQ[ii * DPAD + c * NSPLIT + h], ps);
        pd = fmaf(vr[c], sdO[ii * DPAD + c * NSPLIT + h], pd);
      }
      float sv = group_sum<NSPLIT>(ps) * tau;
      float dp = group_sum<NSPLIT>(pd);
      int qi = i0 + ii;
      bool ok = row_ok && qi < T && (!causal || row <= qi);
      float p = ok ? expf(sv - sL[ii]) : 0.f;
      float pz = p;  // P o Z feeds dV; the undropped P multiplies (dP o Z - delta)
      if (drop.thr) {
        const float z = drop_keep(hkey, qi, row, drop.thr) ? drop.keep_scale : 0.f;
        pz = p * z;
        dp *= z;
      }
      float ds = p * (dp - sD[ii]);
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        dvacc[c] = fmaf(pz, sdO[ii * DPAD + c * NSPLIT + h], dvacc[c]);
        dkacc[c] = fmaf(ds, sQ[ii * DPAD + c * NSPLIT + h], dkacc[c]);
      }
    }
  }
  if (row_ok) {
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      int col = c * NSPLIT + h;
      if (col < d) {
        dk[base + (size_t)row * d + col] = dkacc[c] * tau;
        dv[base + (size_t)row * d + col] = dvacc[c];
      }
    }
  }
}

template <int DPAD, int NSPLIT>
cudaError_t fwd_launch(const float* q, const float* k, const float* v, float* o, float* lse, float* m_out,
                       float* l_out, int BH, int T, int d, bool causal, Dropout drop, cudaStream_t st) {
  dim3 grid((T + kRowsPerBlock - 1) / kRowsPerBlock, BH);
  simt_fwd_kernel<DPAD, NSPLIT><<<grid, kRowsPerBlock * NSPLIT, 0, st>>>(q, k, v, o, lse, m_out, l_out, T, d,
                                                                          causal ? 1 : 0, sqrtf(1.0f / d), drop);
  return cudaGetLastError();
}
template <int DPAD, int NSPLIT>
cudaError_t bwd_launch(const float* q, const float* k, const float* v, const float* dO, const float* lse,
                       const float* delta, float* dq, float* dk, float* dv, int BH, int T, int d, bool causal,
                       bool accumulate_dq, Dropout drop, cudaStream_t st) {
  dim3 grid((T + kRowsPerBlock - 1) / kRowsPerBlock, BH);
  float tau = sqrtf(1.0f / d);
  simt_bwd_dq_kernel<DPAD, NSPLIT><<<grid, kRowsPerBlock * NSPLIT, 0, st>>>(q, k, v, dO, lse, delta, dq, T, d,
                                                                             causal ? 1 : 0, tau, accumulate_dq, drop);
  simt_bwd_dkv_kernel<DPAD, NSPLIT><<<grid, kRowsPerBlock * NSPLIT, 0, st>>>(q, k, v, dO, lse, delta, dk, dv, T, d,
                                                                              causal ? 1 : 0, tau, drop);
  return cudaGetLastError();
}

}  // namespace

cudaError_t simt_fwd(const float* q, const float* k, const float* v, float* o, float* lse, float* m_out,
                     float* l_out, int BH, int T, int d, bool causal, cudaStream_t st, Dropout drop) {
  if (d <= 32) return fwd_launch<32, 1>(q, k, v, o, lse, m_out, l_out, BH, T, d, causal, drop, st);
  if (d <= 64) return fwd_launch<64, 2>(q, k, v, o, lse, m_out, l_out, BH, T, d, causal, drop, st);
  if (d <= 128) return fwd_launch<128, 4>(q, k, v, o, lse, m_out, l_out, BH, T, d, causal, drop, st);
  return cudaErrorInvalidValue;
}

cudaError_t simt_bwd_prep(const float* o, const float* dO, float* delta, const float* l_in, const float* m_in,
                          float* lse_out, size_t rows, int d, cudaStream_t st) {
  const int warps = 8;
  unsigned blocks = (unsigned)((rows + warps - 1) / warps);
  simt_bwd_prep_kernel<<<blocks, warps * 32, 0, st>>>(o, dO, delta, l_in, m_in, lse_out, rows, d);
  return cudaGetLastError();
}

cudaError_t simt_bwd(const float* q, const float* k, const float* v, const float* dO, const float* lse,
                     const float* delta, float* dq, float* dk, float* dv, int BH, int T, int d, bool causal,
                     bool accumulate_dq, cudaStream_t st, Dropout drop) {
  // four lanes per row also at d <= 32: one lane per row leaves the SMs with ~10 resident warps at the MT model's sizes
  if (d <= 32) return bwd_launch<32, 4>(q, k, v, dO, lse, delta, dq, dk, dv, BH, T, d, causal, accumulate_dq, drop, st);
  if (d <= 64) return bwd_launch<64, 2>(q, k, v, dO, lse, delta, dq, dk, dv, BH, T, d, causal, accumulate_dq, drop, st);
  if (d <= 128) return bwd_launch<128, 4>(q, k, v, dO, lse, delta, dq, dk, dv, BH, T, d, causal, accumulate_dq, drop, st);
  return cudaErrorInvalidValue;
}

}  // namespace fa
