// fa_fwd_pair_sm100.cu -- flash-attention forward for sm_100a on CTA PAIRS: tcgen05.mma.cta_group::2 + TMEM + TMA.
//
// Replaces the reference kernel src/flash_attn2_fw.cu:23-297 (and the FA-1 forward src/flash_attn_fw.cu:22-287: same O,
// different saved statistics) at head dim 128 / bf16, the configuration the headline metric is quoted on.
//
// Why a pair.  The single-CTA kernel (fa_fwd_sm100.cu) keeps P in TMEM over the first columns of S, so the next
// S = Q K^T of a Q tile cannot be issued before P V has consumed P: every K/V step of a Q tile is the dependency chain
// S -> softmax -> P -> PV -> QK -> S (3200 cycles against 2048 tensor cycles for two tiles).  Breaking it needs P outside
// the S columns -- in shared memory -- and that needs 64 KB more than one SM has next to two Q tiles and a K/V ring.
// Two CTAs of a cluster sharing every K/V tile have it: with cta_group::2 each CTA stages only HALF of each K/V tile
// (K: 64 of the 128 keys, the B operand of S split along N; V: 64 of the 128 head-dim columns), 32 KB per step
// instead of 64 KB.  Per CTA: Q 64 KB | P 64 KB | K ring 3 x 16 KB | V ring 3 x 16 KB = 224 KB.
//
// One cluster = 512 consecutive query rows of one (batch, head): "pair tile" t (t = 0, 1) is 256 rows, CTA r of the
// pair owns rows [128 r, +128) of it -- its own A operand (Q, then P), its own 128 TMEM lanes of S_t and O_t, its own
// softmax warps.  The leader CTA (rank 0) issues every MMA (M = 256) for both.
//   S_t  = Q_t K_j^T    SS   A = Q_t (own rows, K-major)          B = K_j   (keys [64 r, +64) per CTA, K-major)
//   O_t += P_t V_j      SS   A = P_t (own rows, K-major, written by the softmax warps)   B = V_j (d-columns [64 r, +64), MN-major)
// As soon as the softmax warps of BOTH CTAs hold S_t(j) in registers (s_free) the leader issues S_t(j+1); P_t(j) goes out
// when both CTAs have stored it (p_full) -- the chain is gone, what remains is MUFU and tensor throughput.
// Cross-CTA signalling: TMA loads of the peer complete on the LEADER's mbarriers (cp.async.bulk.tensor ... .cta_group::2),
// the peer's softmax warps arrive remotely (mbarrier.arrive.release.cluster), tcgen05.commit multicasts "S ready", "O ready
// / P consumed", "K / V stage free" to both CTAs.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "fa_internal.h"
#include "fa_ptx.cuh"
#include "fa_tma_host.h"

#ifndef FA_PAIR_PINGPONG
#define FA_PAIR_PINGPONG 1
#endif

namespace fa {

extern long long* g_fwd_trace;  // developer timeline buffer (fa_fwd_sm100.cu, tools/fa_check tracefwd)

namespace {

constexpr int kD = 128;
constexpr int kRows = 128;     // rows per CTA tile (= TMEM lanes)
constexpr int kKeys = 128;     // keys per K/V step
constexpr int kThreads = 384;  // warps 0-3 / 4-7 softmax of tile 0 / 1, 8 TMA, 9 MMA (leader), 10-11 complete the warpgroup
constexpr int kStages = 3;
constexpr int kTileBytes = 128 * 128 * 2;   // Q or P tile of one CTA: 128 rows x 128 bf16 = two 16 KB boxes
constexpr int kBox = 16384;
constexpr int kHalfBytes = 16384;           // this CTA's half of a K or V tile
constexpr int kOffQ = 0, kOffP = 2 * kTileBytes, kOffK = 4 * kTileBytes, kOffV = kOffK + kStages * kHalfBytes;
constexpr int kSmemBytes = kOffV + kStages * kHalfBytes + 1024;

struct PairParams {
  int T, causal;
  float tau, scale_log2;
  void* o;
  float* lse;
  float* m_out;
  float* l_out;
  Dropout drop;
  long long* trace;  // clock64 stamps of the leader CTA of cluster (0, 0), 16 slots per K/V step
};

struct PairBarriers {
  uint64_t q_full;                                   // leader: Q tiles of both CTAs have landed
  uint64_t k_full[kStages], v_full[kStages];         // leader: both halves of the stage have landed
  uint64_t k_empty[kStages], v_empty[kStages];       // each CTA (commit multicast): stage consumed
  uint64_t s_full[2], o_full[2];                     // each CTA (commit multicast): S_t(j) in TMEM / PV_t(j) retired
  uint64_t s_free[2], p_full[2];                     // leader: 8 warps (4 per CTA) hold S_t(j) in registers / have stored P_t(j)
};

// ---- the 2-CTA forms of the PTX wrappers (fa_ptx.cuh has the cta_group::1 ones) ----------------------------------
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem: each CTA holds half of N]; one elected lane issues
__device__ __forceinline__ void umma2_ss_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all MMAs issued so far have retired
__device__ __forceinline__ void umma2_commit_w(uint64_t* bar) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// TMA tile load into THIS CTA's shared memory whose completion is counted on an mbarrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const void* map, uint32_t leader_bar, int c0, int c1, int c2, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
// whole-warp wait with cluster-scope acquire (the arrivals come from the other CTA as well); keeps the warp converged
__device__ __forceinline__ void mbar_wait_cluster_w(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait_cluster(bar, parity))) {
    if ((++spins & 0x3ff) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}

template <typename OutT>
__device__ __forceinline__ void store_out32(OutT* dst, const float* x);
template <>
__device__ __forceinline__ void store_out32<float>(float* dst, const float* x) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
}
template <>
__device__ __forceinline__ void store_out32<__nv_bfloat16>(__nv_bfloat16* dst, const float* x) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    uint4 v;
    v.x = pack_bf16x2(x[i], x[i + 1]);
    v.y = pack_bf16x2(x[i + 2], x[i + 3]);
    v.z = pack_bf16x2(x[i + 4], x[i + 5]);
    v.w = pack_bf16x2(x[i + 6], x[i + 7]);
    *reinterpret_cast<uint4*>(dst + i) = v;
  }
}

template <typename OutT, bool kDrop>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
fa_fwd_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  // the 1024-byte alignment must yield the SAME offsets in both CTAs (one descriptor serves both): it does, the dynamic
  // shared-memory window starts at the same address in every CTA of a launch
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sP = smem + kOffP;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  __shared__ PairBarriers bars;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int T = p.T;
  const int bh = blockIdx.y;
  const int npairs = gridDim.x >> 1;
  const int pr = p.causal ? npairs - 1 - (int)(blockIdx.x >> 1) : (int)(blockIdx.x >> 1);  // heavy causal pairs first
  const int nkv = (T + kKeys - 1) / kKeys;
  // pair tile t = CTA tiles 4 pr + 2 t (rank 0) and 4 pr + 2 t + 1 (rank 1); both run the K/V steps the later one needs
  const int n0 = p.causal ? min(4 * pr + 2, nkv) : nkv;
  const int n1 = p.causal ? min(4 * pr + 4, nkv) : nkv;
  long long* const tr = (blockIdx.x == 0 && blockIdx.y == 0) ? p.trace : nullptr;
#define FA_TRP(j_, e_) do { if (tr && (j_) < 16 && lane == 0) tr[(j_) * 16 + (e_)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    mbar_init(&bars.q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars.k_full[i], 1);
      mbar_init(&bars.v_full[i], 1);
      mbar_init(&bars.k_empty[i], 1);
      mbar_init(&bars.v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars.s_full[t], 1);
      mbar_init(&bars.o_full[t], 1);
      mbar_init(&bars.s_free[t], 8);  // 4 softmax warps of each CTA
      mbar_init(&bars.p_full[t], 8);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc2(&tmem_base_smem, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised and both TMEM allocations exist before anything crosses over
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;

  if (warp >= 8) {
    reg_dealloc<80>();
    if (warp == 8) {
      // ------------------------------------------------------------------ TMA producer (one lane per CTA)
      if (lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        const uint32_t lq = mapa_shared(smem_u32(&bars.q_full), 0);
        if (rank == 0) mbar_arrive_expect_tx(&bars.q_full, 2 * 2 * kTileBytes);  // both CTAs' Q tiles
        for (int t = 0; t < 2; ++t)
          for (int b = 0; b < 2; ++b)
            tma_load_3d_pair(sQ + t * kTileBytes + b * kBox, &map_q, lq, b * 64, (4 * pr + 2 * t + (int)rank) * kRows, bh, kEvictFirst);
        for (int j = 0; j < n1; ++j) {
          const int st = j % kStages;
          const uint32_t ph = (j / kStages) & 1;
          mbar_wait(&bars.k_empty[st], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&bars.k_full[st], 2 * kHalfBytes);
          const uint32_t lk = mapa_shared(smem_u32(&bars.k_full[st]), 0);
          for (int b = 0; b < 2; ++b)  // keys [64 rank, +64) of the tile: two boxes of 64 rows x 64 columns
            tma_load_3d_pair(sK + st * kHalfBytes + b * 8192, &map_k, lk, b * 64, j * kKeys + 64 * (int)rank, bh, kEvictLast);
          mbar_wait(&bars.v_empty[st], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&bars.v_full[st], 2 * kHalfBytes);
          const uint32_t lv = mapa_shared(smem_u32(&bars.v_full[st]), 0);
          // head-dim columns [64 rank, +64) of all 128 keys: one box
          tma_load_3d_pair(sV + st * kHalfBytes, &map_v, lv, 64 * (int)rank, j * kKeys, bh, kEvictLast);
        }
      }
    } else if (warp == 9 && rank == 0) {
      // ------------------------------------------------------------------ MMA issuer: leader CTA only, whole warp (`_w` forms)
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, 256, kKeys, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, 256, kD, 0, 1);
      const uint32_t sQa = smem_u32(sQ), sPa = smem_u32(sP), sKa = smem_u32(sK), sVa = smem_u32(sV);
      auto issue_qk = [&](int t, int st) {
        const uint64_t q0 = make_smem_desc(sQa + t * kTileBytes, 16, 1024);
        const uint64_t k0 = make_smem_desc(sKa + st * kHalfBytes, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)  // 4 k-steps of 32 B per swizzle row; A boxes 16 KB apart, B (64-row) boxes 8 KB apart
          umma2_ss_w(tmem + t * 128, desc_advance(q0, (kk >> 2) * kBox + (kk & 3) * 32), desc_advance(k0, (kk >> 2) * 8192 + (kk & 3) * 32),
                     idesc_qk, kk > 0);
      };
      auto issue_pv = [&](int t, int st, bool acc) {
        const uint64_t p0 = make_smem_desc(sPa + t * kTileBytes, 16, 1024);
        const uint64_t v0 = make_smem_desc(sVa + st * kHalfBytes, kBox, 1024);  // MN-major: 16 key rows (2 KB) per k-step, one 64-column atom
#pragma unroll
        for (int kk = 0; kk < kKeys / 16; ++kk)
          umma2_ss_w(tmem + 256 + t * 128, desc_advance(p0, (kk >> 2) * kBox + (kk & 3) * 32), desc_advance(v0, kk * 2048), idesc_pv,
                     (acc || kk > 0) ? 1u : 0u);
      };
      mbar_wait_cluster_w(&bars.q_full, 0);
      mbar_wait_cluster_w(&bars.k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      umma2_commit_w(&bars.s_full[0]);
      issue_qk(1, 0);
      umma2_commit_w(&bars.s_full[1]);
      umma2_commit_w(&bars.k_empty[0]);
      // Issue order  QK0(j+1) PV1(j-1) QK1(j+1) PV0(j):  the two tiles of a CTA run half a step apart, so that one tile's
      // exponentials (MUFU: 1024 cycles per tile and step at full rate) fall into the other tile's store / hand-over phase
      // instead of both softmaxes sharing the 16 ex2/clk/SM at the same time.
      auto pv1 = [&](int jj) {  // O_1 += P_1(jj) V(jj); releases V(jj), whose other reader PV0(jj) was issued before
        const int st = jj % kStages;
        mbar_wait_cluster_w(&bars.v_full[st], (jj / kStages) & 1);
        mbar_wait_cluster_w(&bars.p_full[1], jj & 1);
        tc_fence_after();
        FA_TRP(jj, 12);
        issue_pv(1, st, jj > 0);
        umma2_commit_w(&bars.o_full[1]);
        umma2_commit_w(&bars.v_empty[st]);
        FA_TRP(jj, 13);
      };
      for (int j = 0; j < n1; ++j) {
        const int st = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        const int jn = j + 1, stn = jn % kStages;
        const uint32_t phn = (jn / kStages) & 1;
        if (jn < n1) mbar_wait_cluster_w(&bars.k_full[stn], phn);
        if (jn < n0) {  // next scores of tile 0: they only need the softmax warps' registers to hold the current ones
          mbar_wait_cluster_w(&bars.s_free[0], j & 1);
          tc_fence_after();
          FA_TRP(j, 8);
          issue_qk(0, stn);
          umma2_commit_w(&bars.s_full[0]);
        }
        if (j > 0) pv1(j - 1);
        if (jn < n1) {
          mbar_wait_cluster_w(&bars.s_free[1], j & 1);
          tc_fence_after();
          issue_qk(1, stn);
          umma2_commit_w(&bars.s_full[1]);
          umma2_commit_w(&bars.k_empty[stn]);
          FA_TRP(j, 9);
        }
        if (j < n0) {
          mbar_wait_cluster_w(&bars.v_full[st], ph);
          mbar_wait_cluster_w(&bars.p_full[0], j & 1);
          tc_fence_after();
          FA_TRP(j, 10);
          issue_pv(0, st, j > 0);
          umma2_commit_w(&bars.o_full[0]);
          FA_TRP(j, 11);
        }
      }
      pv1(n1 - 1);
    }
  } else {
    // ------------------------------------------------------------------ softmax groups (both CTAs)
    reg_alloc<208>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + t * 128 + lane_base;
    const uint32_t tO = tmem + 256 + t * 128 + lane_base;
    const int qt = 4 * pr + 2 * t + (int)rank;  // this CTA's 128-row Q tile
    const int nt = t == 0 ? n0 : n1;
    const int row_g = qt * kRows + r;
    const float c2 = p.scale_log2;
    uint8_t* p_row = sP + t * kTileBytes + r * 128;  // this row inside each 64-key box of P_t
    const uint32_t s_free_leader = mapa_shared(smem_u32(&bars.s_free[t]), 0);
    const uint32_t p_full_leader = mapa_shared(smem_u32(&bars.p_full[t]), 0);
    float m_ref = -INFINITY, m_true = -INFINITY, l = 0.f;
    const uint32_t drop_rowkey = drop_row_key(drop_head_key(p.drop.seed_lo, p.drop.seed_hi, bh), uint32_t(row_g) >> 1);
    const uint32_t drop_thr = p.drop.thr;
    const uint32_t drop_b0 = (row_g & 1) * 2;
    const uint32_t drop_sel2 = drop_sel_pair(drop_b0, drop_b0 + 1);

    // Ping-pong of the two softmax groups of a CTA (named barriers 2 / 3 = "tile 0 / tile 1 may use the MUFU pipe"): with
    // the chain gone both tiles always find their next scores ready, and left to themselves they run in phase -- both load
    // and reduce (MUFU idle), then both exponentiate (16 ex2/clk/SM shared: 2048 cycles), then both store.  Alternating
    // them puts one tile's load / max / store / hand-over under the other tile's exponentials.  Tile 1 has n1 >= n0 steps:
    // its last steps run ungated.
    if (FA_PAIR_PINGPONG && t == 1 && n0 > 0) named_bar_arrive(2, 256);  // tile 0 goes first
    for (int j = 0; j < nt; ++j) {
      mbar_wait(&bars.s_full[t], j & 1);
      tc_fence_after();
      if ((warp & 3) == 0) FA_TRP(j, 4 * t + 0);
      float s[128];
      {
        uint32_t* sr = reinterpret_cast<uint32_t*>(s);
        tmem_ld_x32(tS + 0, sr + 0);
        tmem_ld_x32(tS + 32, sr + 32);
        tmem_ld_x32(tS + 64, sr + 64);
        tmem_ld_x32(tS + 96, sr + 96);
        tmem_wait_ld();
      }
      // the scores are in registers: the leader may overwrite S_t with the next step's
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(s_free_leader);
      const bool dead = p.causal && j > qt;  // rank 0 rides along on the step that only rank 1's tile needs: P = 0
      const bool diag = p.causal && (j == qt);
      const bool tail = (j == nkv - 1) && (T - j * kKeys < kKeys);
      if (diag || tail) {
        int limit = 127;
        if (diag) limit = r;
        if (tail) limit = min(limit, T - 1 - j * kKeys);
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c > limit) s[c] = -INFINITY;
      }
      if (!dead) {
        float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
        for (int c = 4; c < 128; c += 4) {
          mx0 = fmaxf(mx0, s[c]);
          mx1 = fmaxf(mx1, s[c + 1]);
          mx2 = fmaxf(mx2, s[c + 2]);
          mx3 = fmaxf(mx3, s[c + 3]);
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        m_true = fmaxf(m_true, mx);
        if (j == 0) m_ref = mx;
      }
      if (!dead && j > 0) {
        const bool need = (m_true - m_ref) * c2 > 8.f;  // (m_true >= this tile's max; earlier tiles never exceeded the bound)
        if (__any_sync(0xffffffffu, need)) {
          // O_t must be quiescent: PV_t(j-1) retired, PV_t(j) not yet issued (it waits for our P)
          mbar_wait(&bars.o_full[t], (j - 1) & 1);
          tc_fence_after();
          const float m_new = fmaxf(m_ref, m_true);
          const float f = ex2_approx((m_ref - m_new) * c2);
          l *= f;
          m_ref = m_new;
#pragma unroll
          for (int c0 = 0; c0 < kD; c0 += 32) {
            uint32_t orr[32];
            tmem_ld_x32(tO + c0, orr);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
            tmem_st_x32(tO + c0, orr);
          }
          tmem_wait_st();
        }
      }
      // K-major A operand, 128-B swizzle: box b holds keys [64 b, +64); 16-byte chunk c of the row sits at c ^ (r & 7).
      // PV_t(j-1) must have retired before P_t is overwritten: it was issued when P_t(j-1) was complete, i.e. about 1100
      // cycles before it retires -- so the first half of the exponentials goes into registers, then comes the wait (which
      // then does not block), then the stores; the second half is stored as it is produced, under the MUFU time.
      const float nm = -m_ref * c2;
      const float2 c2v = make_float2(c2, c2), nmv = make_float2(nm, nm);
      float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
      auto exp_chunk = [&](const int q4, uint32_t* pk) {  // 32 score columns -> 16 packed words
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const int cc = q4 * 32 + c;
          const float2 x0 = ffma2(make_float2(s[cc], s[cc + 1]), c2v, nmv);
          const float2 x1 = ffma2(make_float2(s[cc + 2], s[cc + 3]), c2v, nmv);
          const float2 p0 = make_float2(ex2_approx(x0.x), ex2_approx(x0.y));
          float2 p1;
          if (FA_EX2_POLY2_PERIOD > 0 && ((cc >> 2) % (FA_EX2_POLY2_PERIOD > 0 ? FA_EX2_POLY2_PERIOD : 1)) == 0) p1 = ex2_poly2(x1);
          else p1 = make_float2(ex2_approx(x1.x), ex2_approx(x1.y));
          la = fadd2(la, p0);
          lb = fadd2(lb, p1);
          pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
          pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
          if constexpr (kDrop) {  // la / lb are the undropped row sum; only P fed to the PV product is thinned
            const uint32_t j2 = (uint32_t(j * kKeys + cc) >> 1) + (lane & 1);
            const uint32_t mine = drop_keep_word(drop_mix(drop_rowkey ^ (j2 * kDropColMul)), drop_thr);
            const uint32_t theirs = __shfl_xor_sync(0xffffffffu, mine, 1);
            const uint32_t w0 = (lane & 1) ? theirs : mine, w1 = (lane & 1) ? mine : theirs;  // columns cc,cc+1 | cc+2,cc+3
            pk[c >> 1] &= drop_prmt(w0, drop_sel2);
            pk[(c >> 1) + 1] &= drop_prmt(w1, drop_sel2);
          }
        }
      };
      auto store_chunk = [&](const int q4, const uint32_t* pk) {
        uint8_t* dst = p_row + (q4 >> 1) * kBox;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(dst + (((((q4 & 1) << 2) + i) ^ (r & 7)) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      };
      if (FA_PAIR_PINGPONG && j < n0) named_bar_sync(2 + t, 256);  // my turn on the MUFU pipe
      if (dead) {  // (kept out of the loops: a branch per element would serialise the exponentials)
        if (j > 0) mbar_wait(&bars.o_full[t], (j - 1) & 1);
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(p_row + b * kBox + (c << 4)) = make_uint4(0u, 0u, 0u, 0u);
      } else {
        uint32_t pk0[16], pk1[16];
        exp_chunk(0, pk0);
        exp_chunk(1, pk1);
        if ((warp & 3) == 0) FA_TRP(j, 4 * t + 1);
        if (j > 0) mbar_wait(&bars.o_full[t], (j - 1) & 1);
        store_chunk(0, pk0);
        store_chunk(1, pk1);
#pragma unroll
        for (int q4 = 2; q4 < 4; ++q4) {
          uint32_t pk[16];
          exp_chunk(q4, pk);
          store_chunk(q4, pk);
        }
        l += (la.x + la.y) + (lb.x + lb.y);
      }
      if (FA_PAIR_PINGPONG) {  // hand the MUFU pipe over: tile 0 -> tile 1 for the same j; tile 1 -> tile 0 for j + 1 (if it has one)
        if (t == 0) named_bar_arrive(3, 256);
        else if (j + 1 < n0) named_bar_arrive(2, 256);
      }
      if ((warp & 3) == 0) FA_TRP(j, 4 * t + 2);
      fence_proxy_async_smem();  // the P stores must be visible to the tensor core (async proxy) of this SM
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(p_full_leader);
      if ((warp & 3) == 0) FA_TRP(j, 4 * t + 3);
    }

    // epilogue: O_t / l -> global, statistics
    mbar_wait(&bars.o_full[t], (nt - 1) & 1);
    tc_fence_after();
    const float inv_l = (kDrop ? p.drop.keep_scale : 1.f) / l;
    const bool row_ok = row_g < T;
    OutT* orow = reinterpret_cast<OutT*>(p.o) + ((size_t)bh * T + (row_ok ? row_g : 0)) * kD;
#pragma unroll
    for (int c0 = 0; c0 < kD; c0 += 32) {
      uint32_t orr[32];
      tmem_ld_x32(tO + c0, orr);
      tmem_wait_ld();
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(orr[i]) * inv_l;
      if (row_ok) store_out32<OutT>(orow + c0, x);
    }
    if (row_ok) {
      const size_t idx = (size_t)bh * T + row_g;
      if (p.lse) p.lse[idx] = m_ref * p.tau + logf(l);
      if (p.m_out) p.m_out[idx] = m_true * p.tau;
      if (p.l_out) p.l_out[idx] = l * expf((m_ref - m_true) * p.tau);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA of the pair leaves (or frees TMEM) while the other may still be served by shared MMAs
  if (warp == 9) tmem_dealloc2(tmem, 512);
}

template <typename OutT, bool kDrop>
cudaError_t launch_pair(const void* q, const void* k, const void* v, void* o, float* lse, float* m_out, float* l_out, int BH, int T,
                        bool causal, Dropout drop, cudaStream_t st) {
  CUtensorMap mq, mk, mv;
  if (make_tmap_3d(&mq, q, 2, kD, T, BH, 64, kRows)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mk, k, 2, kD, T, BH, 64, 64)) return cudaErrorInvalidValue;     // half a K tile: 64 keys x 64 columns per box
  if (make_tmap_3d(&mv, v, 2, kD, T, BH, 64, kKeys)) return cudaErrorInvalidValue;  // half a V tile: 128 keys x 64 columns
  cudaError_t e = cudaFuncSetAttribute(fa_fwd_pair_kernel<OutT, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  PairParams p;
  p.T = T;
  p.causal = causal ? 1 : 0;
  p.tau = sqrtf(1.0f / (float)kD);
  p.scale_log2 = p.tau * 1.4426950408889634f;
  p.o = o;
  p.lse = lse;
  p.m_out = m_out;
  p.l_out = l_out;
  p.drop = drop;
  p.trace = g_fwd_trace;
  const int npairs = (T + 4 * kRows - 1) / (4 * kRows);
  dim3 grid(2 * npairs, BH);
  fa_fwd_pair_kernel<OutT, kDrop><<<grid, kThreads, kSmemBytes, st>>>(mq, mk, mv, p);
  return cudaGetLastError();
}

}  // namespace

// bf16, head dim 128; `o` bf16 or fp32
cudaError_t tc_fwd_pair(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, void* o, bool o_is_f32, float* lse,
                        float* m_out, float* l_out, int BH, int T, bool causal, cudaStream_t st, Dropout drop) {
  if (BH > 65535) return cudaErrorInvalidValue;
  if (o_is_f32)
    return drop.thr ? launch_pair<float, true>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st)
                    : launch_pair<float, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
  return drop.thr ? launch_pair<__nv_bfloat16, true>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st)
                  : launch_pair<__nv_bfloat16, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
}

}  // namespace fa
