// umma_rate.cu -- what one SM's tensor pipe actually delivers per tcgen05.mma for every operand form the attention
// kernels use, alone and next to the traffic the other warps of those kernels generate (shared-memory stores,
// tcgen05.ld), plus the cta_group::2 forms (CTA pair, M = 256): correctness against a CPU product and rate.
//
//   umma_rate            -> table: cycles per 128 x N x 16 bf16 MMA  (nominal: N/2 cycles per SM)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I flash-attention-minitorch_b200/csrc tools/umma_rate.cu -o build/umma_rate
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fa_ptx.cuh"
#include "fa_tma_host.h"

using namespace fa;

enum Form { F_SS_KK = 0, F_SS_KMN = 1, F_SS_MNMN = 2, F_TS_MN = 3, F_TS_K = 4, F_NONE = 5 };
static const char* kFormName[] = {"SS  A=K-major  B=K-major  (S^T, dP^T)", "SS  A=K-major  B=MN-major (dK)",
                                  "SS  A=MN-major B=MN-major (dQ)", "TS  A=TMEM     B=MN-major (dV, PV)",
                                  "TS  A=TMEM     B=K-major", "no MMA"};

struct RateOut {
  long long mma_cycles;
  long long lsu_iters, ldtm_iters, bg_cycles;
};

constexpr int kRateThreads = 512;
constexpr int kRounds = 128;  // x 8 k-steps = 1024 MMAs per measurement


// elect-predicated forms: executed by the whole (converged) warp, one lane issues.  The operands are then warp-uniform
// values, ptxas keeps them in uniform registers and emits bare UTCHMMAs; with `if (lane == 0)` it cannot prove uniformity
// and wraps every UTCHMMA into R2UR moves and an ELECT / BRA.U.ANY loop.
__device__ __forceinline__ void umma_ss_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int kForm, bool kElect>
__device__ __forceinline__ void issue_rounds(uint32_t tmem, uint32_t sA, uint32_t sB, int N) {
  const int b_rows = N;  // K-major B: boxes of N rows x 128 B
  auto kmaj = [](uint32_t base, int kk, int rows) { return make_smem_desc(base + (kk >> 2) * rows * 128 + (kk & 3) * 32, 16, 1024); };
  auto mnmaj = [](uint32_t base, int kk) { return make_smem_desc(base + kk * 2048, 16384, 1024); };
  const uint32_t id_kk = make_idesc(kFmtBF16, 128, N, 0, 0), id_kmn = make_idesc(kFmtBF16, 128, N, 0, 1),
                 id_mnmn = make_idesc(kFmtBF16, 128, N, 1, 1);
  for (int r = 0; r < kRounds; ++r) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t acc = (r | kk) ? 1u : 0u;
      if constexpr (kElect) {
        if constexpr (kForm == F_SS_KK) umma_ss_elect(tmem, kmaj(sA, kk, 128), kmaj(sB, kk, b_rows), id_kk, acc);
        if constexpr (kForm == F_SS_KMN) umma_ss_elect(tmem, kmaj(sA, kk, 128), mnmaj(sB, kk), id_kmn, acc);
        if constexpr (kForm == F_SS_MNMN) umma_ss_elect(tmem, mnmaj(sA, kk), mnmaj(sB, kk), id_mnmn, acc);
        if constexpr (kForm == F_TS_MN) umma_ts_elect(tmem, tmem + 384 + kk * 8, mnmaj(sB, kk), id_kmn, acc);
        if constexpr (kForm == F_TS_K) umma_ts_elect(tmem, tmem + 384 + kk * 8, kmaj(sB, kk, b_rows), id_kk, acc);
      } else {
        if constexpr (kForm == F_SS_KK) umma_ss(tmem, kmaj(sA, kk, 128), kmaj(sB, kk, b_rows), id_kk, acc);
        if constexpr (kForm == F_SS_KMN) umma_ss(tmem, kmaj(sA, kk, 128), mnmaj(sB, kk), id_kmn, acc);
        if constexpr (kForm == F_SS_MNMN) umma_ss(tmem, mnmaj(sA, kk), mnmaj(sB, kk), id_mnmn, acc);
        if constexpr (kForm == F_TS_MN) umma_ts(tmem, tmem + 384 + kk * 8, mnmaj(sB, kk), id_kmn, acc);
        if constexpr (kForm == F_TS_K) umma_ts(tmem, tmem + 384 + kk * 8, kmaj(sB, kk, b_rows), id_kk, acc);
      }
    }
  }
}
template <bool kElect>
__device__ __forceinline__ void issue_form(int form, uint32_t tmem, uint32_t sA, uint32_t sB, int N) {
  switch (form) {
    case F_SS_KK: issue_rounds<F_SS_KK, kElect>(tmem, sA, sB, N); break;
    case F_SS_KMN: issue_rounds<F_SS_KMN, kElect>(tmem, sA, sB, N); break;
    case F_SS_MNMN: issue_rounds<F_SS_MNMN, kElect>(tmem, sA, sB, N); break;
    case F_TS_MN: issue_rounds<F_TS_MN, kElect>(tmem, sA, sB, N); break;
    case F_TS_K: issue_rounds<F_TS_K, kElect>(tmem, sA, sB, N); break;
  }
}

// mode bit 0: warps 0-7 stream st.shared.v4 (conflict-free, 4 wavefronts per instruction)
// mode bit 1: warps 8-11 stream tcgen05.ld.32x32b.x32 (128 lanes x 32 columns x 4 B = 16 KB per round of 4 warps)
// mode bit 2: warps 0-7 stream tcgen05.ld.32x32b.x32 instead of stores
__global__ void __launch_bounds__(kRateThreads, 1) rate_kernel(RateOut* out, int form, int N, int mode, int elect) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 16; i += kRateThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 13) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
  if (warp == 13) {
    if (elect) {
      // the whole warp stays converged; one elected lane issues each MMA / the commit
      const long long t_start = clock64();
      while (clock64() - t_start < 4000) {}
      __syncwarp();
      const long long t0 = clock64();
      issue_form<true>(form, tmem, sA, sB, N);
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (lane == 0) {
        out->mma_cycles = t1 - t0;
        stop = 1;
      }
    } else if (lane == 0) {
      // let the background warps reach steady state
      const long long t_start = clock64();
      while (clock64() - t_start < 4000) {}
      const long long t0 = clock64();
      if (form != F_NONE) {
        issue_form<false>(form, tmem, sA, sB, N);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
      } else {
        while (clock64() - t0 < 30000) {}
      }
      const long long t1 = clock64();
      out->mma_cycles = t1 - t0;
      stop = 1;
    }
  } else if (warp < 12) {
    const bool do_lsu = (mode & 1) && warp < 8;
    const bool do_ldtm = ((mode & 2) && warp >= 8) || ((mode & 4) && warp < 8);
    long long iters = 0;
    const long long t0 = clock64();
    if (do_lsu) {
      uint8_t* scratch = smem + 163840 + (warp & 7) * 4096;  // 8 x 4 KB
      const int r = lane;
      uint4 v = make_uint4((uint32_t)warp, 1, 2, 3);
      while (!stop) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(scratch + r * 128 + ((ch ^ (r & 7)) << 4))), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        ++iters;
      }
    } else if (do_ldtm) {
      const uint32_t t = tmem + 256 + (uint32_t((warp & 3) * 32) << 16);
      uint32_t acc = 0;
      while (!stop) {
        uint32_t rr[32];
        tmem_ld_x32(t + (iters & 3) * 32, rr);
        tmem_wait_ld();
        acc += rr[0] + rr[31];
        ++iters;
      }
      if (acc == 0x12345) out->bg_cycles = 1;
    }
    const long long t1 = clock64();
    if (lane == 0 && do_lsu && warp == 0) { out->lsu_iters = iters; out->bg_cycles = t1 - t0; }
    if (lane == 0 && do_ldtm && (warp == 8 || (warp == 0 && (mode & 4)))) { out->ldtm_iters = iters; out->bg_cycles = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Latency of one dependent step of the attention pipelines: issue 8 MMAs (one 128 x 128 x 128 product) + tcgen05.commit,
// then wait on the mbarrier -- either in the issuing warp or in ANOTHER warp that answers through a second mbarrier
// (the MMA -> consumer -> MMA round trip of the S -> softmax -> PV and dS -> dQ -> drain -> dP chains).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) latency_kernel(long long* out, int groups, int via_consumer, int consumer_ld) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_back;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_back, 4);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sA = smem_u32(smem), sB = sA + 32768;
  if (warp == 5) {
    const uint64_t a0 = make_smem_desc(sA, 16, 1024), b0 = make_smem_desc(sB, 16, 1024);
    const uint32_t id = make_idesc(kFmtBF16, 128, 128, 0, 0);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_ss_w(tmem, desc_advance(a0, (kk >> 2) * 16384 + (kk & 3) * 32), desc_advance(b0, (kk >> 2) * 16384 + (kk & 3) * 32), id, kk > 0);
      umma_commit_w(&bar_mma);
      if (via_consumer) mbar_wait_w(&bar_back, g & 1);
      else mbar_wait_w(&bar_mma, g & 1);
      tc_fence_after();
    }
    if (lane == 0) out[0] = clock64() - t0;
  } else if (warp < 4 && via_consumer) {
    const uint32_t t = tmem + (uint32_t(warp * 32) << 16);
    uint32_t acc = 0;
    for (int g = 0; g < groups; ++g) {
      mbar_wait(&bar_mma, g & 1);
      tc_fence_after();
      if (consumer_ld) {
        uint32_t rr[32];
        tmem_ld_x32(t, rr);
        tmem_wait_ld();
        acc += rr[3];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_back);
    }
    if (acc == 0x1234567) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

static void run_latency() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* names[3] = {"issuing warp waits on its own commit", "consumer warps wait, arrive back (no TMEM read)",
                          "consumer warps wait, tcgen05.ld 32 columns, arrive back"};
  for (int v = 0; v < 3; ++v) {
    const int groups = 256;
    latency_kernel<<<1, 256, 65536>>>(d, groups, v > 0, v > 1);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("dependent step (8 MMAs 128x128x16 + commit + wait): %-58s %7.1f cycles per step (8 MMAs back to back: ~520)  %s\n", names[v],
           double(c) / groups, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(d);
}

// ---------------------------------------------------------------------------------------------------------------
// Chip-wide throughput of the backward's dQ path: every SM streams 128-row x 32-column fp32 slabs (16 KB, the staging
// unit of fa_bwd_kernel) from shared memory into a [rows][128] fp32 tensor with cp.reduce.async.bulk.tensor (.add) or
// plain TMA stores, two staging buffers and cp.async.bulk.wait_group.read as in the kernel.
// ---------------------------------------------------------------------------------------------------------------
template <int kInFlight>
__global__ void __launch_bounds__(128, 1) reduce_kernel(const __grid_constant__ CUtensorMap map, int n_tiles, int iters, int do_reduce, int cols,
                                                        long long* cycles, float* gbase) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int slab_bytes = cols * 4 * 128;
    const long long t0 = clock64();
    uint32_t chunk = 0;
    for (int it = 0; it < iters; ++it) {
      const int tile = (blockIdx.x * 7 + it) % n_tiles;  // CTAs walk the tensor at different offsets, every tile is hit by many CTAs
      for (int c = 0; c < 128 / cols; ++c, ++chunk) {
        if (chunk >= kInFlight) tma_store_wait_read<kInFlight - 1>();
        uint8_t* buf = smem + (chunk % kInFlight) * slab_bytes;
        if (do_reduce == 3) {  // 1-D bulk reduce of 16 KB = 32 whole rows of the tile (contiguous in global memory)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gbase + ((size_t)tile * 128 + c * 32) * 128),
                       "r"(smem_u32(buf)), "r"(16384)
                       : "memory");
        } else if (do_reduce == 2) {  // tensor form, box {128 columns, 32 rows}: rows of 512 B
          tma_reduce_add_3d(&map, buf, 0, tile * 128 + c * 32, 0);
        } else if (do_reduce) tma_reduce_add_3d(&map, buf, c * cols, tile * 128, 0);
        else tma_store_3d(&map, buf, c * cols, tile * 128, 0);
        tma_store_commit();
      }
    }
    tma_store_wait_all<0>();
    cycles[blockIdx.x] = clock64() - t0;
  }
}

static void run_reduce() {
  const int n_tiles = 8192;  // x 64 KB = 512 MB
  float* buf;
  long long* cyc;
  cudaMalloc(&buf, (size_t)n_tiles * 128 * 128 * 4);
  cudaMemset(buf, 0, (size_t)n_tiles * 128 * 128 * 4);
  cudaMalloc(&cyc, 148 * 8);
  auto launch = [&](int in_flight, const CUtensorMap& map, int grid, int iters, int do_reduce, int cols) {
    if (in_flight == 2) reduce_kernel<2><<<grid, 128, 65536>>>(map, n_tiles, iters, do_reduce, cols, cyc, buf);
    if (in_flight == 4) reduce_kernel<4><<<grid, 128, 65536>>>(map, n_tiles, iters, do_reduce, cols, cyc, buf);
    if (in_flight == 8) reduce_kernel<8><<<grid, 128, 65536>>>(map, n_tiles, iters, do_reduce, cols, cyc, buf);
  };
  cudaFuncSetAttribute(reduce_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(reduce_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(reduce_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  printf("dQ path: 128-row fp32 slabs from shared memory into a [rows][128] fp32 tensor, one issuing thread per SM, 148 CTAs\n");
  for (int do_reduce = 0; do_reduce < 2; ++do_reduce)
    for (int cols : {32, 16, 8})
      for (int in_flight : {2, 4, 8}) {
        if (cols * 512 * in_flight > 65536) continue;
        CUtensorMap map;
        if (make_tmap_3d(&map, buf, 4, 128, (uint64_t)n_tiles * 128, 1, cols, 128, 0)) return;
        const int iters = 1000, grid = 148;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        launch(in_flight, map, grid, 50, do_reduce, cols);
        cudaEventRecord(e0);
        launch(in_flight, map, grid, iters, do_reduce, cols);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        long long c0 = 0;
        cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = double(grid) * iters * 65536.0;
        printf("TMA %-10s slab %2d KB x %d in flight (%2d KB): %7.1f GB/s chip-wide, %6.1f B/clk per SM, %5.0f cycles per 64 KB tile  %s\n",
               do_reduce ? "reduce-add" : "store", cols / 2, in_flight, cols / 2 * in_flight, bytes / ms * 1e-6, 65536.0 * iters / double(c0),
               double(c0) / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  for (int mode : {2, 3}) {  // wide rows
    CUtensorMap map;
    if (make_tmap_3d(&map, buf, 4, 128, (uint64_t)n_tiles * 128, 1, 128, 32, 0)) return;
    const int iters = 1000, grid = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(2, map, grid, 50, mode, 32);
    cudaEventRecord(e0);
    launch(2, map, grid, iters, mode, 32);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c0 = 0;
    cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
    printf("TMA reduce-add, 16 KB = 32 rows x 512 B, %-22s x 2 in flight: %7.1f GB/s chip-wide, %6.1f B/clk per SM, %5.0f cycles per 64 KB tile  %s\n",
           mode == 2 ? "tensor box {128, 32}" : "1-D bulk (contiguous)", double(grid) * iters * 65536.0 / ms * 1e-6, 65536.0 * iters / double(c0),
           double(c0) / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(buf); cudaFree(cyc);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA pair: D[256 x N] = A[256 x K] B[N x K]^T with tcgen05.mma.cta_group::2.  CTA r holds rows [128 r, +128) of A and
// rows [N/2 r, +N/2) of B (both K-major, 128-B swizzle, written by hand with st.shared) and receives rows [128 r, +128)
// of D in its own TMEM (lanes 0-127, columns [0, N)).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// mode 0: SS, B K-major.  mode 1: SS, B MN-major (CTA r holds N columns [N/2 r, +N/2) of B^T = [K][N]).
// mode 2: TS (A from TMEM, written with tcgen05.st by each CTA for its own 128 rows), B MN-major.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int N, int K,
            int mode, long long* cycles, int rate_rounds) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* sA = smem;           // [128 rows][K] K-major: K/64 boxes of 128 rows x 128 B
  uint8_t* sB = smem + 65536;   // K-major: K/64 boxes of N/2 rows x 128 B;  MN-major: (N/2)/64 boxes of K rows x 128 B
  const int half = N / 2;
  for (int idx = threadIdx.x; idx < 128 * (K / 8); idx += 128) {
    const int r = idx / (K / 8), c = idx % (K / 8);  // row, 16-byte chunk
    const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + r) * K + c * 8);
    *reinterpret_cast<uint4*>(sA + (c >> 3) * 128 * 128 + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = v;
  }
  if (mode == 0) {
    for (int idx = threadIdx.x; idx < half * (K / 8); idx += 128) {
      const int r = idx / (K / 8), c = idx % (K / 8);
      const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)(rank * half + r) * K + c * 8);
      *reinterpret_cast<uint4*>(sB + (c >> 3) * half * 128 + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = v;
    }
  } else {
    // MN-major: element (k, n) of this CTA's half lives in box n / 64, row k, 16-byte chunk (n % 64) / 8
    for (int idx = threadIdx.x; idx < K * (half / 8); idx += 128) {
      const int k = idx / (half / 8), c = idx % (half / 8);
      __nv_bfloat16 tmp[8];
      for (int e = 0; e < 8; ++e) tmp[e] = B[(size_t)(rank * half + c * 8 + e) * K + k];
      *reinterpret_cast<uint4*>(sB + (c >> 3) * K * 128 + k * 128 + (((c & 7) ^ (k & 7)) << 4)) = *reinterpret_cast<uint4*>(tmp);
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0 && rate_rounds == 1) cycles[1 + rank] = tmem;  // TMEM base each CTA of the pair was given
  if (mode == 2) {  // A into TMEM columns [128, 128 + K/2)
    const int row = threadIdx.x;
    const uint32_t lane_addr = tmem + 128 + (uint32_t(warp * 32) << 16);
    for (int c0 = 0; c0 < K; c0 += 64) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(A + (size_t)(rank * 128 + row) * K + c0 + 2 * i);
        r[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st_x32(lane_addr + c0 / 2, r);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(kFmtBF16, 256, N, 0, mode == 0 ? 0 : 1);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int rr = 0; rr < rate_rounds; ++rr)
      for (int kk = 0; kk < K / 16; ++kk) {
        const uint64_t ad = make_smem_desc(a0 + (kk >> 2) * 128 * 128 + (kk & 3) * 32, 16, 1024);
        const uint64_t bd = mode == 0 ? make_smem_desc(b0 + (kk >> 2) * half * 128 + (kk & 3) * 32, 16, 1024)
                                      : make_smem_desc(b0 + kk * 2048, K * 128, 1024);
        const uint32_t acc = (rr | kk) ? 1u : 0u;
        if (mode == 2) umma_ts_2cta(tmem, tmem + 128 + kk * 8, bd, idesc, acc);
        else umma_ss_2cta(tmem, ad, bd, idesc, acc);
      }
    umma_commit_2cta(&bar, 3);
    mbar_wait(&bar, 0);
    *cycles = clock64() - t0;
  } else {
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  tc_fence_after();
  {
    const int row = rank * 128 + threadIdx.x;
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(lane_addr + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  (void)lane;
}

static int run_pair(int N, int K, int mode) {
  const int M = 256;
  std::vector<__nv_bfloat16> A(M * K), B(N * K);
  std::vector<float> Af(M * K), Bf(N * K);
  srand(7 + N + K + mode);
  for (int i = 0; i < M * K; ++i) { Af[i] = float(rand() % 17 - 8) / 8.f; A[i] = __float2bfloat16(Af[i]); }
  for (int i = 0; i < N * K; ++i) { Bf[i] = float(rand() % 17 - 8) / 8.f; B[i] = __float2bfloat16(Bf[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dC, 24);
  cudaMemcpy(dA, A.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), N * K * 2, cudaMemcpyHostToDevice);
  const int smem_bytes = 65536 * 2;
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const char* mname[] = {"SS B=K-major", "SS B=MN-major", "TS B=MN-major"};
  // correctness
  cudaMemset(dD, 0xff, M * N * 4);
  pair_kernel<<<2, 128, smem_bytes>>>(dA, dB, dD, N, K, mode, dC, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("pair M=256 N=%3d K=%3d %-14s CUDA ERROR %s\n", N, K, mname[mode], cudaGetErrorString(e));
    return 1;
  }
  long long bases[3] = {0, 0, 0};
  cudaMemcpy(bases, dC, 24, cudaMemcpyDeviceToHost);
  std::vector<float> Dh(M * N);
  cudaMemcpy(Dh.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad_rows0 = 0, bad_rows1 = 0, bad_cols_lo = 0, bad_cols_hi = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += double(Af[m * K + k]) * double(Bf[n * K + k]);
      const double err = fabs(ref - Dh[m * N + n]);
      if (!(err <= 1e-3)) {
        (m < 128 ? bad_rows0 : bad_rows1)++;
        (n < N / 2 ? bad_cols_lo : bad_cols_hi)++;
      }
      if (!(err <= maxerr)) maxerr = err;
    }
  // rate
  const int rounds = 256 / (K / 16);
  pair_kernel<<<2, 128, smem_bytes>>>(dA, dB, dD, N, K, mode, dC, rounds);
  e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  printf("pair M=256 N=%3d K=%3d %-14s max_abs_err=%.3e %s (bad: rows<128 %d, rows>=128 %d, cols lo %d, cols hi %d; tmem base %llx / %llx)   %6.1f cycles per MMA (%s)\n",
         N, K, mname[mode], maxerr, maxerr < 1e-3 ? "OK" : "FAIL", bad_rows0, bad_rows1, bad_cols_lo, bad_cols_hi, bases[1], bases[2],
         double(cyc) / 256.0, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return maxerr < 1e-3 ? 0 : 1;
}

int main(int argc, char** argv) {
  RateOut* d;
  cudaMalloc(&d, sizeof(RateOut));
  const int smem_bytes = 196608;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int only_pair = argc > 1 && (!strcmp(argv[1], "pair") || !strcmp(argv[1], "lat"));
  if (!only_pair) {
    printf("cycles per tcgen05.mma (M=128, K=16, bf16), 256 back-to-back MMAs on one SM; background = other warps of the CTA\n");
    printf("%-40s %4s | %8s %14s %14s %14s %14s\n", "form", "N", "alone", "+st.shared x8w", "+tc.ld x4w", "+tc.ld x12w", "+st x8w,ld x4w");
    const int modes[] = {0, 1, 2, 6, 3};
    for (int elect = 0; elect < 2; ++elect)
    for (int N : {128, 256, 64}) {
      if (N == 128) printf("-- issue: %s\n", elect ? "converged warp, elect.sync-predicated MMAs (uniform-register descriptors)" : "if (lane == 0) { ... } (divergent: R2UR + ELECT loop per MMA)");
      for (int form = 0; form <= F_NONE; ++form) {
        if ((N != 128 || elect) && form == F_NONE) continue;
        printf("%-40s %4d |", kFormName[form], N);
        for (int mode : modes) {
          if (form == F_NONE && mode == 0) { printf(" %8s", "-"); continue; }
          cudaMemset(d, 0, sizeof(RateOut));
          rate_kernel<<<1, kRateThreads, smem_bytes>>>(d, form, N, mode, elect);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf(" ERROR %s\n", cudaGetErrorString(e));
            return 1;
          }
          RateOut h;
          cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
          char buf[64];
          if (form == F_NONE) {
            // background throughput alone: st.shared bytes/clk (8 warps x 4 KB per iteration), tcgen05.ld bytes/clk
            const double lsu = h.bg_cycles ? double(h.lsu_iters) * 8 * 4096 / h.bg_cycles : 0;
            const double nld = (mode & 4) ? 12 : 4;
            const double ld = h.bg_cycles ? double(h.ldtm_iters) * nld * 4096 / h.bg_cycles : 0;
            snprintf(buf, sizeof buf, "%s%.0fB/clk", (mode & 1) && !(mode & 2) ? "st " : (mode & 1) ? "st+ld(ld) " : "ld ",
                     (mode & 1) && !(mode & 2) ? lsu : ld);
          } else {
            // MMA cycles, and what the background warps achieved next to the MMAs (st.shared B/clk | tcgen05.ld B/clk)
            const double lsu = h.bg_cycles ? double(h.lsu_iters) * 8 * 4096 / h.bg_cycles : 0;
            const double nld = (mode & 4) ? 12 : 4;
            const double ld = h.bg_cycles ? double(h.ldtm_iters) * nld * 4096 / h.bg_cycles : 0;
            if (mode == 0) snprintf(buf, sizeof buf, "%.1f", double(h.mma_cycles) / (kRounds * 8));
            else if ((mode & 1) && !(mode & 2)) snprintf(buf, sizeof buf, "%.1f st%.0f", double(h.mma_cycles) / (kRounds * 8), lsu);
            else snprintf(buf, sizeof buf, "%.1f ld%.0f", double(h.mma_cycles) / (kRounds * 8), ld);
          }
          printf(mode == 0 ? " %8s" : " %14s", buf);
        }
        printf("\n");
      }
    }
  }
  run_latency();
  run_reduce();
  int fails = 0;
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {128, 256})
      for (int K : {64, 128}) {
        if (mode == 2 && N == 256) continue;  // 256 TMEM columns allocated: D [0, N) + A [128, 192)
        fails += run_pair(N, K, mode);
      }
  printf("umma_rate: %d failing pair cases\n", fails);
  return 0;
}
