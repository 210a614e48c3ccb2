// tma_mix.cu -- do the three data streams of the backward kernel get in each other's way inside one SM?
//   stream R: fp32 TMA reduce-adds (cp.reduce.async.bulk.tensor .add) of 16 KB slabs, two in flight   (the dQ partial sums)
//   stream L: TMA tile loads of 32 KB (two 64-column bf16 boxes), one outstanding, latency measured      (Q / dO tiles)
//   stream A: red.global.add.v4.f32 from registers, coalesced 512 B per warp instruction                  (an LSU-side dQ path)
// Every SM runs the selected streams concurrently (one warp each); the report gives, per stream, bytes per clock per SM and
// -- for the loads -- the mean latency from issue to mbarrier completion.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -I flash-attention-minitorch_b200/csrc tools/tma_mix.cu -o build/tma_mix
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "fa_ptx.cuh"
#include "fa_tma_host.h"

using namespace fa;

struct Out {
  long long r_cycles, l_cycles, l_lat_sum, a_cycles;
};

__global__ void __launch_bounds__(128, 1)
mix_kernel(const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_in, float* red_base, int n_tiles_out,
           int n_tiles_in, int iters, int streams, int inflight_loads, Out* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  uint8_t* stage = smem;                 // 2 x 16 KB reduce staging
  uint8_t* tiles = smem + 32768;         // up to 4 x 32 KB load targets
  for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(stage)[i] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0 && (streams & 1)) {
    const long long t0 = clock64();
    uint32_t chunk = 0;
    for (int it = 0; it < iters; ++it) {
      const int tile = (blockIdx.x * 7 + it) % n_tiles_out;
      for (int c = 0; c < 4; ++c, ++chunk) {
        if (chunk >= 2) tma_store_wait_read<1>();
        tma_reduce_add_3d(&map_out, stage + (chunk & 1) * 16384, c * 32, tile * 128, 0);
        tma_store_commit();
      }
    }
    tma_store_wait_all<0>();
    out[blockIdx.x].r_cycles = clock64() - t0;
  }
  if (warp == 1 && lane == 0 && (streams & 2)) {
    const long long t0 = clock64();
    long long lat = 0;
    // `inflight_loads` tiles outstanding: issue times kept in a small ring
    long long issued[4];
    for (int it = 0; it < iters + inflight_loads; ++it) {
      if (it >= inflight_loads) {
        const int s = (it - inflight_loads) % inflight_loads;
        mbar_wait(&bar[s], ((it - inflight_loads) / inflight_loads) & 1);
        lat += clock64() - issued[s];
      }
      if (it < iters) {
        const int s = it % inflight_loads;
        const int tile = (blockIdx.x / 32 * 32 + (blockIdx.x + it) % 32 + it / 32 * 37) % n_tiles_in;  // neighbours share tiles (L2 hits)
        mbar_arrive_expect_tx(&bar[s], 32768);
        issued[s] = clock64();
        tma_load_3d(tiles + s * 32768, &map_in, &bar[s], 0, tile * 128, 0, kEvictLast);
        tma_load_3d(tiles + s * 32768 + 16384, &map_in, &bar[s], 64, tile * 128, 0, kEvictLast);
      }
    }
    out[blockIdx.x].l_cycles = clock64() - t0;
    out[blockIdx.x].l_lat_sum = lat;
  }
  if (warp >= 2 && (streams & 4)) {
    const long long t0 = clock64();
    // each warp instruction: 32 lanes x 16 B contiguous = one 512-B row of a [rows][128] fp32 tile
    for (int it = 0; it < iters; ++it) {
      const int tile = (blockIdx.x * 5 + it) % n_tiles_out;
      float* base = red_base + (size_t)tile * 128 * 128;
#pragma unroll 4
      for (int row = warp - 2; row < 128; row += 2) {
        float* p = base + row * 128 + lane * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
      }
    }
    __threadfence();
    if (lane == 0 && warp == 2) out[blockIdx.x].a_cycles = clock64() - t0;
  }
}

// Same streams inside CTA pairs (clusters of two): each CTA requests HALF of every tile (one 64-column box) and multicasts it
// to both CTAs of the pair, i.e. both receive whole tiles while the L2 serves each byte once.  Stage reuse is guarded by a
// pair of "empty" barriers that both CTAs arrive on (local + remote arrive) after they have seen the tile.
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask,
                                               uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6, %7;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask),
        "l"(hint)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
mix_pair_kernel(const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_in, int n_tiles_out, int n_tiles_in,
                int iters, int streams, int inflight_loads, Out* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4], empty[4];
  uint8_t* stage = smem;
  uint8_t* tiles = smem + 32768;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(stage)[i] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 2);  // both CTAs of the pair have consumed the stage
    }
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0 && (streams & 1)) {
    const long long t0 = clock64();
    uint32_t chunk = 0;
    for (int it = 0; it < iters; ++it) {
      const int tile = (blockIdx.x * 7 + it) % n_tiles_out;
      for (int c = 0; c < 4; ++c, ++chunk) {
        if (chunk >= 2) tma_store_wait_read<1>();
        tma_reduce_add_3d(&map_out, stage + (chunk & 1) * 16384, c * 32, tile * 128, 0);
        tma_store_commit();
      }
    }
    tma_store_wait_all<0>();
    out[blockIdx.x].r_cycles = clock64() - t0;
  }
  if (warp == 1 && lane == 0 && (streams & 2)) {
    const long long t0 = clock64();
    long long lat = 0;
    long long issued[4];
    const int pair = blockIdx.x >> 1;
    for (int it = 0; it < iters + inflight_loads; ++it) {
      if (it >= inflight_loads) {
        const int k = it - inflight_loads, s = k % inflight_loads;
        mbar_wait(&full[s], (k / inflight_loads) & 1);
        lat += clock64() - issued[s];
        // release the stage in both CTAs
        mbar_arrive(&empty[s]);
        mbar_arrive_remote(mapa_shared(smem_u32(&empty[s]), rank ^ 1));
      }
      if (it < iters) {
        const int s = it % inflight_loads;
        if (it >= inflight_loads) mbar_wait_cluster(&empty[s], ((it / inflight_loads) - 1) & 1);
        const int tile = (pair / 16 * 16 + (pair + it) % 16 + it / 16 * 37) % n_tiles_in;
        mbar_arrive_expect_tx(&full[s], 32768);
        issued[s] = clock64();
        tma_load_3d_mc(tiles + s * 32768 + rank * 16384, &map_in, &full[s], rank * 64, tile * 128, 0, 3, kEvictLast);
      }
    }
    out[blockIdx.x].l_cycles = clock64() - t0;
    out[blockIdx.x].l_lat_sum = lat;
  }
  cluster_sync_all();
}

int main(int argc, char** argv) {
  const int n_out = 4096, n_in = 2048;  // 256 MB fp32 reduce target, 64 MB bf16 load source
  float* dq;
  __nv_bfloat16* src;
  Out* out;
  cudaMalloc(&dq, (size_t)n_out * 128 * 128 * 4);
  cudaMemset(dq, 0, (size_t)n_out * 128 * 128 * 4);
  cudaMalloc(&src, (size_t)n_in * 128 * 128 * 2);
  cudaMemset(src, 0, (size_t)n_in * 128 * 128 * 2);
  cudaMalloc(&out, 148 * sizeof(Out));
  CUtensorMap mo, mi;
  if (make_tmap_3d(&mo, dq, 4, 128, (uint64_t)n_out * 128, 1, 32, 128, 0)) return 1;
  if (make_tmap_3d(&mi, src, 2, 128, (uint64_t)n_in * 128, 1, 64, 128, 1)) return 1;
  const int smem = 32768 + 4 * 32768;
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  printf("streams: R = TMA reduce-add fp32 (64 KB per tile), L = TMA loads (32 KB bf16 tiles), A = red.global.add.v4.f32 (64 KB per tile); 148 CTAs, %d tiles each\n", iters);
  for (int inflight : {1, 2}) {
    for (int streams : {1, 2, 4, 3, 5, 6, 7}) {
      cudaMemset(out, 0, 148 * sizeof(Out));
      mix_kernel<<<148, 128, smem>>>(mo, mi, dq, n_out, n_in, 20, streams, inflight, out);
      mix_kernel<<<148, 128, smem>>>(mo, mi, dq, n_out, n_in, iters, streams, inflight, out);
      cudaError_t e = cudaDeviceSynchronize();
      Out h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double r = 0, l = 0, lat = 0, a = 0;
      for (int i = 0; i < 148; ++i) {
        r += h[i].r_cycles;
        l += h[i].l_cycles;
        lat += h[i].l_lat_sum;
        a += h[i].a_cycles;
      }
      r /= 148; l /= 148; lat /= 148; a /= 148;
      printf("loads in flight %d  streams %c%c%c :", inflight, streams & 1 ? 'R' : '-', streams & 2 ? 'L' : '-', streams & 4 ? 'A' : '-');
      if (streams & 1) printf("  R %5.1f B/clk (%5.0f cyc / 64 KB)", 65536.0 * iters / r, r / iters);
      if (streams & 2) printf("  L %5.1f B/clk, latency %5.0f cyc / 32 KB tile", 32768.0 * iters / l, lat / iters);
      if (streams & 4) printf("  A %5.1f B/clk (%5.0f cyc / 64 KB)", 65536.0 * iters / a, a / iters);
      printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  cudaFuncSetAttribute(mix_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("CTA pairs, every tile half-requested by each CTA and multicast to both (L2 serves each byte once):\n");
  for (int inflight : {1, 2}) {
    for (int streams : {2, 3}) {
      cudaMemset(out, 0, 148 * sizeof(Out));
      mix_pair_kernel<<<148, 128, smem>>>(mo, mi, n_out, n_in, 20, streams, inflight, out);
      mix_pair_kernel<<<148, 128, smem>>>(mo, mi, n_out, n_in, iters, streams, inflight, out);
      cudaError_t e = cudaDeviceSynchronize();
      Out h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double r = 0, l = 0, lat = 0;
      for (int i = 0; i < 148; ++i) {
        r += h[i].r_cycles;
        l += h[i].l_cycles;
        lat += h[i].l_lat_sum;
      }
      r /= 148; l /= 148; lat /= 148;
      printf("pair multicast, loads in flight %d  streams %c%c :", inflight, streams & 1 ? 'R' : '-', streams & 2 ? 'L' : '-');
      if (streams & 1) printf("  R %5.1f B/clk (%5.0f cyc / 64 KB)", 65536.0 * iters / r, r / iters);
      printf("  L %5.1f B/clk received per SM, latency %5.0f cyc / 32 KB tile", 32768.0 * iters / l, lat / iters);
      printf("  %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
