// fa_ptx.cuh -- thin inline-PTX layer for sm_100a: mbarrier, TMA, TMEM, tcgen05.
//
// Everything the tensor-core attention kernels need from the Blackwell ISA lives here so the
// kernels themselves read as pipelines, not as asm.  No CUTLASS/CuTe types are used; bit layouts
// of the UMMA shared-memory and instruction descriptors are hand-encoded below and exercised by
// tools/umma_probe.cu on real hardware.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fa {

// ---------------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store / UMMA smem operands)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait for the phase with the given parity to complete.  A lost arrival (pipeline bug) would
// otherwise hang the GPU until the driver watchdog fires; trap after ~4 s instead so the process
// gets a launch error it can report.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Thread-block clusters: rank, cluster barrier, distributed shared memory (DSMEM) stores and remote mbarrier arrives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// wait with cluster-scope acquire: the data behind the barrier was written by another CTA of the cluster
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 3-D tiled load: coordinates are (innermost, middle, outermost) in elements.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "l"(hint)
      : "memory");
}
// L2 prefetch of a tile (no shared-memory destination, no completion): warms the L2 for a tile that is loaded later
__device__ __forceinline__ void tma_prefetch_l2_3d(const void* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 1-D bulk copy global -> shared (size multiple of 16 B, both addresses 16-B aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const void* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM allocation
// ---------------------------------------------------------------------------------------------
// One full warp calls alloc; the base address is written to *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs of this thread retire.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4   [46,48) version = 1 on sm_100
//   [49,52) base offset (0: tiles are 1024-B aligned)   [61,64) swizzle: 2 = 128 B
//
// K-major, 128-B swizzle (rows of 128 B, 8-row groups 1024 B apart):  LBO unused (1), SBO = 1024.
// MN-major, 128-B swizzle (64-element MN atoms of 128 B; K rows 128 B apart, 8-row K groups SBO
// apart, successive MN atoms LBO apart).
constexpr uint64_t kDescSw128 = (uint64_t(1) << 46) | (uint64_t(2) << 61);
// 128-B swizzle with 32-B atomicity (XOR of address bits [5,7) with [7,9)); MN-major 32-bit only.
constexpr uint64_t kDescSw128Base32 = (uint64_t(1) << 46) | (uint64_t(1) << 61);

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint64_t layout = kDescSw128) {
  return layout | uint64_t((saddr >> 4) & 0x3fff) | (uint64_t((lbo_bytes >> 4) & 0x3fff) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3fff) << 32);
}

// Descriptor of the same tile advanced by `bytes` (a multiple of 16 that stays inside the 256 KB window of the 14-bit
// address field): one 32-bit add on the low word instead of re-encoding -- what the MMA issue loops use per k-step.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) {
  return (desc & 0xffffffff00000000ull) | uint64_t(uint32_t(desc) + (bytes >> 4));
}

// Instruction descriptor (32 bit) for kind::f16 / kind::tf32 with fp32 accumulation.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt  [10,13) B fmt  (0 f16, 1 bf16, 2 tf32)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N>>3  [24,29) M>>4
enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2 };
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t m, uint32_t n, uint32_t a_mn_major,
                                                  uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
template <bool kTF32 = false>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem]
template <bool kTF32 = false>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-collective issue (the `_w` forms): called by ALL 32 lanes of a converged warp, one lane elected inside the asm
// issues the instruction.  Why: every operand of UTCHMMA / UTCBAR lives in a uniform register.  Inside `if (lane == 0)`
// ptxas cannot prove the descriptors warp-uniform, so it builds them in vector registers and wraps EVERY tcgen05.mma
// into R2UR moves plus an ELECT / BRA.U.ANY loop -- measured 115-140 cycles per 128x128x16 MMA in the attention
// kernels against 66-68 for the same MMAs issued from uniform registers (tools/umma_rate.cu).  With the whole warp
// executing the issue code, the descriptor arithmetic runs on the uniform datapath and the MMAs are emitted back to
// back.  elect.sync picks the same lane on every call of a converged warp, which is what tcgen05.commit needs (it
// tracks the MMAs of the issuing thread).
// ---------------------------------------------------------------------------------------------
template <bool kTF32 = false>
__device__ __forceinline__ void umma_ss_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
template <bool kTF32 = false>
__device__ __forceinline__ void umma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// mbarrier wait by a whole warp that must stay converged afterwards: the loop condition is a warp vote, i.e. uniform,
// so ptxas keeps treating the code after the wait as convergent (uniform-datapath descriptor arithmetic, bare UTCHMMAs).
__device__ __forceinline__ void mbar_wait_w(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if ((++spins & 0x3ff) == 0) {  // same watchdog as mbar_wait: trap after ~4 s instead of hanging the GPU
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMEM <-> registers.  Shape 32x32b: lane i of the warp reads/writes TMEM lane (base_lane + i),
// N consecutive 32-bit columns.  A warp may only touch lanes [32*(warp_id%4), +32).
// ---------------------------------------------------------------------------------------------
#define FA_R4(a, i) "=r"(a[i]), "=r"(a[i + 1]), "=r"(a[i + 2]), "=r"(a[i + 3])
#define FA_W4(a, i) "r"(a[i]), "r"(a[i + 1]), "r"(a[i + 2]), "r"(a[i + 3])

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : FA_R4(r, 0), FA_R4(r, 4)
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : FA_R4(r, 0), FA_R4(r, 4), FA_R4(r, 8), FA_R4(r, 12)
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : FA_R4(r, 0), FA_R4(r, 4), FA_R4(r, 8), FA_R4(r, 12), FA_R4(r, 16), FA_R4(r, 20), FA_R4(r, 24), FA_R4(r, 28)
      : "r"(taddr));
}
// Shape 16x256b (the classic accumulator-fragment layout): one instruction reads 16 TMEM lanes starting at the
// lane in taddr; repeat n covers columns [8n, 8n+8).  Thread t gets, per repeat, {row t/4: cols 2(t%4), +1} in
// regs 4n, 4n+1 and {row t/4 + 8: same cols} in regs 4n+2, 4n+3 -- i.e. each quad owns 32 contiguous bytes of a row.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : FA_R4(r, 0), FA_R4(r, 4), FA_R4(r, 8), FA_R4(r, 12)
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : FA_R4(r, 0), FA_R4(r, 4), FA_R4(r, 8), FA_R4(r, 12), FA_R4(r, 16), FA_R4(r, 20), FA_R4(r, 24), FA_R4(r, 28)
      : "r"(taddr));
}
// fp32 pair reduction into global memory (no return value): RED.ADD.F32x2
__device__ __forceinline__ void red_add_f32x2(float* gptr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gptr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), FA_W4(r, 0),
               FA_W4(r, 4)
               : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      FA_W4(r, 0), FA_W4(r, 4), FA_W4(r, 8), FA_W4(r, 12)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      FA_W4(r, 0), FA_W4(r, 4), FA_W4(r, 8), FA_W4(r, 12), FA_W4(r, 16), FA_W4(r, 20), FA_W4(r, 24), FA_W4(r, 28)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// misc math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, rel. error ~1e-4, far below the
// bf16 rounding of P).  The MUFU unit (16 ex2/clk/SM) is as scarce as the tensor pipe at head dim 128, so a
// fixed fraction of the exponentials of every tile is evaluated here instead.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  float xi;  // floor(x) + 1.5 * 2^23 : the integer part sits in the low mantissa bits
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(xi) : "f"(x), "f"(12582912.f));
  const float f = x - (xi - 12582912.f);  // fractional part in [0, 1)
  const float p = fmaf(fmaf(fmaf(0.07711909f, f, 0.22756439f), f, 0.69514614f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xi) << 23));  // scale by 2^floor(x)
}
#ifndef FA_EX2_POLY2_PERIOD
#define FA_EX2_POLY2_PERIOD 0  // n > 0: of every n groups of four scores, one pair goes to ex2_poly2 (fraction 1/(2n))
#endif
#ifndef FA_EX2_POLY_MASK
#define FA_EX2_POLY_MASK 0x00  // bit i set: elements with index % 8 == i use the polynomial.  Measured on B200:
                               // 0x00 (MUFU only) is fastest once the softmax warps have enough registers.
#endif
// element `idx` (compile-time after unrolling) of a tile row: MUFU or polynomial by a fixed interleave
__device__ __forceinline__ float ex2_mixed(float x, int idx) {
  return ((FA_EX2_POLY_MASK >> (idx & 7)) & 1) ? ex2_poly(x) : ex2_approx(x);
}
// packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2): one issue slot for two lanes
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return d;
}
// Packed variant: two exponentials per issue slot on the FMA pipe.  x = j + f with j = round(x) (magic-number add),
// f in [-1/2, 1/2]; 2^f by a near-minimax degree-3 polynomial (max rel. error 7.5e-5, far below bf16's 2^-9);
// 2^j by adding j to the exponent field.  6 packed FP ops + 2 clamps + 4 integer ops per pair.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);  // 1.5 * 2^23: the integer lands in the low mantissa bits
  const float2 jm = fadd2(x, magic);
  const float2 j = fadd2(jm, make_float2(-12582912.f, -12582912.f));
  const float2 f = fadd2(x, make_float2(-j.x, -j.y));
  float2 p = ffma2(make_float2(0.05517164f, 0.05517164f), f, make_float2(0.24261113f, 0.24261113f));
  p = ffma2(p, f, make_float2(0.69326097f, 0.69326097f));
  p = ffma2(p, f, make_float2(0.99992806f, 0.99992806f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(jm.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(jm.y) << 23)));
}
__device__ __forceinline__ float round_tf32(float x) {  // round-to-nearest (ties away) fp32 -> tf32, kept in fp32 bits
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

}  // namespace fa
