// fa_gemm_sm100.cu -- fp32-storage / tf32 tensor-core GEMM for the decoder-LM linear layers.
//
// Replaces, for the Linear layers of the model path, the reference's MatrixMultiplyKernel (src/combine.cu:150-253, 467-535:
// a 32 x 32 shared-memory-tiled fp32 FMA kernel driven through host staging, minitorch/cuda_kernel_ops.py:302-380):
//     C[M x N] = A[M x K] * B[K x N]  (+ bias[N])         fp32 in HBM, tcgen05.mma kind::tf32, fp32 accumulate in TMEM
// Each operand may be stored with either dimension contiguous, which covers the three products of a Linear layer
// without any transposition pass (W is (in, out) row-major as in minitorch/modules_basic.py:82-122):
//     forward   Y  = X  W        A = X  (K contiguous)        B = W  (N contiguous)
//     backward  dX = dY W^T      A = dY (K contiguous)        B = W  viewed (N' = in, K' = out): K contiguous
//               dW = X^T dY      A = X  viewed (M' = in, K' = tokens): M contiguous;  B = dY: N contiguous
// "K contiguous" operands are K-major UMMA operands (128-byte swizzle); "M/N contiguous" operands are MN-major and use
// the 128B swizzle with 32-byte atomicity that UMMA requires for 32-bit MN-major data (verified by tools/umma_probe).
//
// One CTA computes one 128 x 128 tile: warp 4 = TMA producer (6-stage ring of 128 x 32 operand slabs), warp 5 = MMA
// issuer (4 k-steps of 8 per slab), warps 0-3 = epilogue: TMEM -> registers (+ bias) -> 128B-swizzled shared memory ->
// TMA store (or TMA reduce-add for C += / split-K) of 128 x 32 column slabs, double buffered.  A thread owns one output
// row, so direct global stores would touch 32 different lines per instruction; the TMA engine writes full lines and
// clips the tile at the matrix edge.  TMA zero-fills out-of-range operand rows / columns, so M, N, K only need to be
// multiples of 4 (16-byte global strides).
#include <cuda.h>
#include <math.h>

#include "fa_internal.h"
#include "fa_ptx.cuh"
#include "fa_tma_host.h"

namespace fa {

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;  // tile; kBK fp32 = one 128-byte swizzle row
constexpr int kSlabBytes = 128 * 128;           // 128 x 32 fp32, either orientation
constexpr int kGemmThreads = 192;
// operand ring + two output staging slabs.  Six stages (225 KB, one CTA per SM) for long reductions; two stages (97 KB,
// two CTAs per SM, so one CTA's epilogue overlaps the other's main loop) when a CTA has at most 16 k-blocks -- the
// model's forward GEMMs have K = 256, i.e. 8 k-blocks followed by a 64 KB epilogue.
constexpr int gemm_smem(int stages) { return (2 * stages + 2) * kSlabBytes + 1024; }
constexpr int kMaxStages = 6;

struct GemmParams {
  int M, N, K;
  float* c;
  const float* bias;  // nullptr = none
  int accumulate;     // C += A B instead of C = A B
  int k_blocks_per_split;  // split-K: CTA z reduces k-blocks [z * this, (z + 1) * this) and red.adds into a zeroed C
  // "Head layout" operands: a logical [tokens = B*T][E = H*d] matrix that physically lives as (B, H, T, d) -- the layout
  // the attention kernels consume and produce.  Its tensor map is {d, T, B*H}; a 32-column slab is (part of) one head and a
  // row tile stays inside one batch (T % 128 == 0), so the TMA engine does the (B,T,H,d) <-> (B,H,T,d) permutation for free:
  // the projections write Q/K/V straight into attention layout and read O / dQ / dK / dV straight out of it.
  // A head-layout operand may be `parts` such tensors back to back ((parts, B, H, T, d): q | k | v of one fused
  // projection): its logical width is parts * H * d and part w of it starts B*H heads further.
  int a_heads, b_heads, c_heads;
  int T, H, d, E, BH;  // E = H * d, BH = B * H
};

// coordinates of the 32-wide slab starting at column `e` / row `m` of a head-layout matrix in its {d, T, B*H} tensor map
struct HeadCoord {
  int c0, c1, c2;
};
// the coordinate splits into a part that depends on the feature index e only and one that depends on the token index m
// only: the producer loop hoists whichever is fixed for the CTA (five runtime-divisor divisions per box otherwise --
// 40 per k-block with two MN-major head operands, against ~260 cycles of MMA work per k-block)
struct HeadPart {
  int c, head;
};
__device__ __forceinline__ HeadPart head_e(const GemmParams& p, int e) {
  const int part = e / p.E, el = e - part * p.E;
  return {el % p.d, part * p.BH + el / p.d};
}
__device__ __forceinline__ HeadPart head_m(const GemmParams& p, int m) { return {m % p.T, (m / p.T) * p.H}; }
__device__ __forceinline__ HeadCoord head_coord(const GemmParams& p, int e, int m) {
  const HeadPart pe = head_e(p, e), pm = head_m(p, m);
  return {pe.c, pm.c, pe.head + pm.head};
}

struct GemmBarriers {
  uint64_t full[kMaxStages], empty[kMaxStages], acc_full;
};

// kAMn / kBMn: operand stored with its M / N dimension contiguous (MN-major) instead of K contiguous (K-major)
// kHeads: some operand is in head layout (a separate instantiation: the plain GEMM's producer loop must not carry the
// head-coordinate branches -- measured 131 -> 185 us on the MN x MN weight-gradient GEMM when it did)
template <bool kAMn, bool kBMn, int kStages, bool kHeads>
__global__ void __launch_bounds__(kGemmThreads, kStages <= 2 ? 2 : 1)
fa_gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kSlabBytes;
  uint8_t* sC = smem + 2 * kStages * kSlabBytes;
  __shared__ GemmBarriers bars;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int nk_total = (p.K + kBK - 1) / kBK;
  const int kb_begin = blockIdx.z * p.k_blocks_per_split;
  const int nk = min(p.k_blocks_per_split, nk_total - kb_begin);  // >= 1 by construction of the grid
  const bool split = gridDim.z > 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars.full[i], 1);
      mbar_init(&bars.empty[i], 1);
    }
    mbar_init(&bars.acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(&tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
      // head-layout operands: the parts of the box coordinates that are fixed for this CTA
      HeadPart ae[4] = {}, be[4] = {}, am = {};
      if constexpr (kHeads) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          if (kAMn && p.a_heads) ae[x] = head_e(p, m0 + 32 * x);
          if (kBMn && p.b_heads) be[x] = head_e(p, n0 + 32 * x);
        }
        if (!kAMn && p.a_heads) am = head_m(p, m0);
      }
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % kStages;
        const int k0 = (kb_begin + kb) * kBK;
        HeadPart km = {};  // token part of the MN-major head operands' coordinates (tokens = the reduction dimension)
        if (kHeads && ((kAMn && p.a_heads) || (kBMn && p.b_heads))) km = head_m(p, k0);
        mbar_wait(&bars.empty[st], ((kb / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.full[st], 2 * kSlabBytes);
        uint8_t* a = sA + st * kSlabBytes;
        uint8_t* b = sB + st * kSlabBytes;
        if constexpr (kAMn) {  // global [K][M]: four boxes of (32 m) x (32 k rows)
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            if (kHeads && p.a_heads) {  // [tokens][E] in head layout, tokens = the reduction dimension
              tma_load_3d(a + x * 4096, &map_a, &bars.full[st], ae[x].c, km.c, ae[x].head + km.head, kEvictNormal);
            } else {
              tma_load_3d(a + x * 4096, &map_a, &bars.full[st], m0 + 32 * x, k0, 0, kEvictNormal);
            }
          }
        } else {               // global [M][K]: one box of (32 k) x (128 m rows)
          if (kHeads && p.a_heads) {
            const HeadPart ke = head_e(p, k0);
            tma_load_3d(a, &map_a, &bars.full[st], ke.c, am.c, ke.head + am.head, kEvictNormal);
          } else {
            tma_load_3d(a, &map_a, &bars.full[st], k0, m0, 0, kEvictNormal);
          }
        }
        if constexpr (kBMn) {
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            if (kHeads && p.b_heads) {
              tma_load_3d(b + x * 4096, &map_b, &bars.full[st], be[x].c, km.c, be[x].head + km.head, kEvictNormal);
            } else {
              tma_load_3d(b + x * 4096, &map_b, &bars.full[st], n0 + 32 * x, k0, 0, kEvictNormal);
            }
          }
        } else {
          tma_load_3d(b, &map_b, &bars.full[st], k0, n0, 0, kEvictNormal);
        }
      }
    }
  } else if (warp == 5) {
    // the whole warp runs the issue code (warp-uniform descriptors, one elected lane issues: `_w` forms of fa_ptx.cuh)
    {
      // REDUX: makes the TMEM base provably warp-uniform for ptxas (see the MMA issuer of fa_bwd_sm100.cu)
      const uint32_t tmem = __reduce_max_sync(0xffffffffu, tmem_slot);
      constexpr uint32_t idesc = make_idesc(kFmtTF32, kBM, kBN, kAMn ? 1 : 0, kBMn ? 1 : 0);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % kStages;
        mbar_wait_w(&bars.full[st], (kb / kStages) & 1);
        tc_fence_after();
        // K-major: 8 tf32 = 32 bytes along the swizzle row, 8-row groups 1024 B apart.  MN-major: 8 k-rows = 1 KB,
        // 4-row groups 512 B apart, 32-element MN atoms one 4 KB box apart, 32-byte-atom swizzle.
        const uint32_t a = a0 + st * kSlabBytes, b = b0 + st * kSlabBytes;
        const uint64_t da0 = kAMn ? make_smem_desc(a, 4096, 512, kDescSw128Base32) : make_smem_desc(a, 16, 1024);
        const uint64_t db0 = kBMn ? make_smem_desc(b, 4096, 512, kDescSw128Base32) : make_smem_desc(b, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk)
          umma_ss_w<true>(tmem, desc_advance(da0, kk * (kAMn ? 1024 : 32)), desc_advance(db0, kk * (kBMn ? 1024 : 32)), idesc,
                          (kb > 0 || kk > 0) ? 1u : 0u);
        umma_commit_w(&bars.empty[st]);
      }
      umma_commit_w(&bars.acc_full);
    }
  } else {
    // epilogue: thread = one output row (TMEM lane); four 32-column slabs through two staging buffers
    mbar_wait(&bars.acc_full, 0);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const uint32_t tacc = tmem + (uint32_t(warp * 32) << 16);
    const bool leader = threadIdx.x == 0;
    const bool reduce = split || p.accumulate;
    const bool add_bias = p.bias && blockIdx.z == 0;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      if (n0 + c * 32 >= p.N) break;  // slab entirely outside the matrix (uniform across the CTA)
      uint8_t* buf = sC + (c & 1) * kSlabBytes;
      uint32_t rr[32];
      tmem_ld_x32(tacc + c * 32, rr);
      tmem_wait_ld();
      if (add_bias) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int col = n0 + c * 32 + i;
          if (col < p.N) {  // N is a multiple of 4: a float4 is entirely inside or outside
            const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col);
            rr[i] = __float_as_uint(__uint_as_float(rr[i]) + b4.x);
            rr[i + 1] = __float_as_uint(__uint_as_float(rr[i + 1]) + b4.y);
            rr[i + 2] = __float_as_uint(__uint_as_float(rr[i + 2]) + b4.z);
            rr[i + 3] = __float_as_uint(__uint_as_float(rr[i + 3]) + b4.w);
          }
        }
      }
      if (leader && c >= 2) tma_store_wait_read<1>();  // the buffer's previous store has been read out
      named_bar_sync(1, 128);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(buf + r * 128 + ((ch ^ (r & 7)) << 4)) =
            make_uint4(rr[ch * 4], rr[ch * 4 + 1], rr[ch * 4 + 2], rr[ch * 4 + 3]);
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        HeadCoord h = {n0 + c * 32, m0, 0};
        if (kHeads && p.c_heads) h = head_coord(p, n0 + c * 32, m0);
        if (reduce) tma_reduce_add_3d(&map_c, buf, h.c0, h.c1, h.c2);
        else tma_store_3d(&map_c, buf, h.c0, h.c1, h.c2);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 128);
}

template <bool kAMn, bool kBMn>
cudaError_t launch_gemm(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, bool accumulate,
                        const HeadLayout& hl, cudaStream_t st) {
  CUtensorMap ma, mb;
  // K-major operand [rows][K]: box (32 k, 128 rows), 128B swizzle.  MN-major operand [K][cols]: box (32 cols, 32 k rows),
  // 128B swizzle with 32-byte atoms.  Head-layout operands: {d, T, B*H} with the same boxes.
  // token count and number of parts ((parts, B, H, T, d)) of each head-layout operand
  const int E = hl.H * hl.d > 0 ? hl.H * hl.d : 1;
  const uint64_t tokens = hl.a ? (kAMn ? (uint64_t)K : (uint64_t)M) : hl.b ? (uint64_t)K : (uint64_t)M;
  const uint64_t BH = (hl.a || hl.b || hl.c) ? tokens / hl.T * hl.H : 0;
  const uint64_t parts_a = (kAMn ? M : K) / E, parts_b = N / E, parts_c = N / E;
  if (hl.a ? make_tmap_3d(&ma, a, 4, hl.d, hl.T, parts_a * BH, 32, kAMn ? 32 : 128, kAMn ? 2 : 1)
           : (kAMn ? make_tmap_3d(&ma, a, 4, M, K, 1, 32, 32, 2) : make_tmap_3d(&ma, a, 4, K, M, 1, 32, 128, 1)))
    return cudaErrorInvalidValue;
  if (hl.b ? make_tmap_3d(&mb, b, 4, hl.d, hl.T, parts_b * BH, 32, 32, 2)  // only the MN-major B (dW = X^T dY_heads)
           : (kBMn ? make_tmap_3d(&mb, b, 4, N, K, 1, 32, 32, 2) : make_tmap_3d(&mb, b, 4, K, N, 1, 32, 128, 1)))
    return cudaErrorInvalidValue;
  CUtensorMap mc;  // C [M][N]: slabs of (32 columns, 128 rows), 128B swizzle, clipped at the edges by the TMA engine
  if (hl.c ? make_tmap_3d(&mc, c, 4, hl.d, hl.T, parts_c * BH, 32, 128, 1) : make_tmap_3d(&mc, c, 4, N, M, 1, 32, 128, 1))
    return cudaErrorInvalidValue;
  cudaError_t e;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.c = c; p.bias = bias; p.accumulate = accumulate ? 1 : 0;
  const bool heads = hl.a || hl.b || hl.c;
  p.a_heads = hl.a; p.b_heads = hl.b; p.c_heads = hl.c; p.T = hl.T > 0 ? hl.T : 1; p.H = hl.H; p.d = hl.d > 0 ? hl.d : 1;
  p.E = E; p.BH = (int)BH;
  // split-K when the tiles do not fill whole waves of the 148 SMs -- weight gradients (256 x 256 outputs reduced over
  // thousands of tokens: 4 tiles), but also 88 tiles (60 % of the SMs) or 158 (a second wave for 10 tiles).  Cost model:
  // waves x (k-blocks per split + ~6 k-blocks of prologue / epilogue per CTA); each split keeps at least 8 k-blocks;
  // partial tiles are red.added into a zeroed (or the accumulated-into) C.
  const int tiles = ((N + kBN - 1) / kBN) * ((M + kBM - 1) / kBM);
  const int nk_total = (K + kBK - 1) / kBK;
  int splits = 1;
  if (tiles <= 2 * 148) {
    long long best = -1;
    for (int sp = 1; sp <= 40 && (sp == 1 || nk_total / sp >= 8); ++sp) {
      const long long waves = ((long long)tiles * sp + 147) / 148;
      const long long cost = waves * ((nk_total + sp - 1) / sp + 6);
      if (best < 0 || cost < best) {
        best = cost;
        splits = sp;
      }
    }
  }
  p.k_blocks_per_split = (nk_total + splits - 1) / splits;
  splits = (nk_total + p.k_blocks_per_split - 1) / p.k_blocks_per_split;  // no empty split
  if (splits > 1 && !accumulate) {
    e = cudaMemsetAsync(c, 0, (size_t)M * N * sizeof(float), st);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((N + kBN - 1) / kBN, (M + kBM - 1) / kBM, splits);
  // Short reductions on a grid that oversubscribes the SMs: the 2-stage variant (two CTAs per SM hide each other's
  // prologue / epilogue).  A grid of at most one CTA per SM gains nothing from that and needs the deep ring instead:
  // with 2 stages every k-block waits a full TMA latency (the 5632 x 256 x 256 layers: ~10 us for 0.7 us of math).
  if (p.k_blocks_per_split <= 16 && (long long)grid.x * grid.y * grid.z > 148) {
    auto kern = heads ? fa_gemm_tf32_kernel<kAMn, kBMn, 2, true> : fa_gemm_tf32_kernel<kAMn, kBMn, 2, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(2));
    if (e != cudaSuccess) return e;
    kern<<<grid, kGemmThreads, gemm_smem(2), st>>>(ma, mb, mc, p);
  } else {
    auto kern = heads ? fa_gemm_tf32_kernel<kAMn, kBMn, 6, true> : fa_gemm_tf32_kernel<kAMn, kBMn, 6, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(6));
    if (e != cudaSuccess) return e;
    kern<<<grid, kGemmThreads, gemm_smem(6), st>>>(ma, mb, mc, p);
  }
  return cudaGetLastError();
}

// out[n] (+)= sum_m x[m][n]   (bias gradient).  Block = 32 column-quads x 8 row lanes (float4 loads), grid.y strips of
// rows; the 8 row lanes meet in shared memory and each block issues one atomic per column.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N) {
  const int cq = blockIdx.x * 32 + (threadIdx.x & 31);  // column quad
  const int w = threadIdx.x >> 5;
  __shared__ float4 part[8][32];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cq * 4 < N) {
    const int step = 8 * gridDim.y;
    int m = blockIdx.y * 8 + w;
    for (; m + 3 * step < M; m += 4 * step) {  // four independent 16-byte loads in flight
      const float4 v0 = *reinterpret_cast<const float4*>(x + (size_t)m * N + cq * 4);
      const float4 v1 = *reinterpret_cast<const float4*>(x + (size_t)(m + step) * N + cq * 4);
      const float4 v2 = *reinterpret_cast<const float4*>(x + (size_t)(m + 2 * step) * N + cq * 4);
      const float4 v3 = *reinterpret_cast<const float4*>(x + (size_t)(m + 3 * step) * N + cq * 4);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; m < M; m += step) {
      const float4 v = *reinterpret_cast<const float4*>(x + (size_t)m * N + cq * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && cq * 4 < N) {
    float4 s = part[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) {
      const float4 v = part[i][threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = out + cq * 4;
    atomicAdd(o, s.x); atomicAdd(o + 1, s.y); atomicAdd(o + 2, s.z); atomicAdd(o + 3, s.w);
  }
}

// out[part * H * d + h * d + dd] += sum over (b, t) of x[part][b][h][t][dd]   (bias gradient of a projection whose output is in
// head layout; parts = 3 for a fused q | k | v projection)
__global__ void __launch_bounds__(256) colsum_heads_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H,
                                                          int T, int d) {
  // thread = (row lane r, column quad q): float4 loads, four independent rows in flight per thread
  const int part = blockIdx.x / H, h = blockIdx.x - part * H;
  const int dq = d >> 2, lanes = blockDim.x / dq;
  const int q = threadIdx.x % dq, r = threadIdx.x / dq;
  x += (size_t)part * B * H * T * d + (size_t)h * T * d + q * 4;
  __shared__ float4 part_sum[256];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int BT = B * T, step = lanes * gridDim.y;
  auto at = [&](int row) {
    const int b = row / T, t = row - b * T;
    return *reinterpret_cast<const float4*>(x + ((size_t)b * H * T + t) * d);
  };
  if (r < lanes) {
    int row = blockIdx.y * lanes + r;
    for (; row + 3 * step < BT; row += 4 * step) {
      const float4 v0 = at(row), v1 = at(row + step), v2 = at(row + 2 * step), v3 = at(row + 3 * step);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; row < BT; row += step) {
      const float4 v = at(row);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part_sum[threadIdx.x] = acc;  // (threads past lanes * dq hold zeros)
  __syncthreads();
  if (threadIdx.x < dq) {
    float4 s = part_sum[threadIdx.x];
    for (int i = 1; i < lanes; ++i) {
      const float4 v = part_sum[i * dq + threadIdx.x];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = out + (size_t)part * H * d + h * d + threadIdx.x * 4;
    atomicAdd(o, s.x); atomicAdd(o + 1, s.y); atomicAdd(o + 2, s.z); atomicAdd(o + 3, s.w);
  }
}

}  // namespace

cudaError_t gemm_tf32(const float* a, const float* b, float* c, const float* bias, int M, int N, int K, bool a_mn_major,
                      bool b_mn_major, bool accumulate, cudaStream_t st, HeadLayout hl) {
  if (M <= 0 || N <= 0 || K <= 0 || (M & 3) || (N & 3) || (K & 3)) return cudaErrorInvalidValue;
  if (hl.a || hl.b || hl.c) {
    // head-layout operands: whole 32-column slabs per head, row tiles inside one batch
    if (hl.T <= 0 || hl.H <= 0 || hl.d <= 0 || hl.d % 32 || hl.T % 128) return cudaErrorInvalidValue;
    if (hl.b && !b_mn_major) return cudaErrorInvalidValue;
    const int E = hl.H * hl.d;
    // the head-layout dimension is one or several (q | k | v) full widths E
    if (hl.a && (a_mn_major ? (M % E || K % hl.T) : (K % E || M % hl.T))) return cudaErrorInvalidValue;
    if (hl.b && (N % E || K % hl.T)) return cudaErrorInvalidValue;
    if (hl.c && (N % E || M % hl.T)) return cudaErrorInvalidValue;
  }
  if (a_mn_major) return b_mn_major ? launch_gemm<true, true>(a, b, c, bias, M, N, K, accumulate, hl, st)
                                    : launch_gemm<true, false>(a, b, c, bias, M, N, K, accumulate, hl, st);
  return b_mn_major ? launch_gemm<false, true>(a, b, c, bias, M, N, K, accumulate, hl, st)
                    : launch_gemm<false, false>(a, b, c, bias, M, N, K, accumulate, hl, st);
}

cudaError_t colsum_heads_f32(const float* x, float* out, int B, int H, int T, int d, bool accumulate, cudaStream_t st, int parts) {
  if (B <= 0 || H <= 0 || T <= 0 || d <= 0 || d > 256 || (d & 3) || parts <= 0) return cudaErrorInvalidValue;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)parts * H * d * 4, st);
    if (e != cudaSuccess) return e;
  }
  int strips = (148 * 4 + parts * H - 1) / (parts * H);
  const int lanes = 256 / (d / 4);
  // at least four rows per thread (the unrolled loop), else the strips only add atomics
  if (strips > (B * T + 4 * lanes - 1) / (4 * lanes)) strips = (B * T + 4 * lanes - 1) / (4 * lanes);
  colsum_heads_kernel<<<dim3(parts * H, strips), 256, 0, st>>>(x, out, B, H, T, d);
  return cudaGetLastError();
}

cudaError_t colsum_f32(const float* x, float* out, int M, int N, bool accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0 || (N & 3)) return cudaErrorInvalidValue;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * 4, st);
    if (e != cudaSuccess) return e;
  }
  const int col_blocks = (N / 4 + 31) / 32;
  int strips = (148 * 8 + col_blocks - 1) / col_blocks;
  // at least 8 rows per row lane: narrow matrices (N = 256: two column blocks) otherwise end up with one row per thread
  // and ~600 atomics per column (18 us for 5.8 MB)
  if (strips > (M + 63) / 64) strips = (M + 63) / 64;
  colsum_kernel<<<dim3(col_blocks, strips), 256, 0, st>>>(x, out, M, N);
  return cudaGetLastError();
}

}  // namespace fa
