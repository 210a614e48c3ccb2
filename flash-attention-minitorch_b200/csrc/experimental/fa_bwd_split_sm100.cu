// EXPERIMENTAL, NOT BUILT: two-kernel (dQ | dK/dV) backward with every on-chip operand in TMEM.  Correct on B200
// (fa_check bwd green) but measured 7.84 ms vs 7.19 ms for the fused kernel at B=32 H=16 T=4096 d=128 causal;
// kept as the documented alternative discussed in DESIGN.md ("backward: fused vs split").

// fa_bwd_sm100.cu -- flash-attention backward for sm_100a: tcgen05.mma + TMEM + TMA.
//
// Replaces the reference kernels src/flash_attn2_bw.cu:21-263 (FA-2 backward: P = exp(S - L), atomicAdd
// into dQ) and src/flash_attn_bw.cu:21-261 (FA-1 backward; its (l, m) pair is folded into L up front).
//
// Design note (measured on B200): a first version did all five contractions of a tile pair
// in ONE kernel (CTA per K/V tile, dQ leaving through an smem staging tile and fp32 TMA reduce-add).  It ran
// at 764 TFLOP/s and its clock64 timeline showed the tensor pipe starved by SHARED-MEMORY bandwidth
// (128 B/clk/SM): every smem x smem MMA instruction streams 8 KB of operands for 64 cycles of math, and the
// dQ staging added another 128 KB per tile pair.  The backward is therefore split into two kernels in which
// every operand that is constant for a CTA, and every operand produced on chip, lives in TMEM:
//
//   dQ kernel   (CTA = one 128-row Q tile i, streams K_j / V_j):        TMEM: Q | dO | S | dP(dS) | dQ
//       S  = Q_i K_j^T   (TMEM x smem)      P  = exp2(c S - L_i)                      [registers]
//       dP = dO_i V_j^T  (TMEM x smem)      dS = tau P o (dP - D_i)  -> TMEM (bf16)    [registers]
//       dQ_i += dS K_j   (TMEM x smem, accumulates over j; written once, no atomics, deterministic)
//   dK/dV kernel (CTA = one 128-row K/V tile j, streams Q_i / dO_i), scores held TRANSPOSED (lanes = keys):
//       S^T  = K_j Q_i^T  (smem x smem)     P^T  = exp2(c S^T - L_i)       -> TMEM (bf16)
//       dP^T = V_j dO_i^T (smem x smem)     dS^T = tau P^T o (dP^T - D_i)  -> TMEM (bf16)
//       dV_j += P^T dO_i  (TMEM x smem)     dK_j += dS^T Q_i  (TMEM x smem)            TMEM: S | dP | dV | dK
//
// Seven MMAs per tile pair instead of five, but 416 KB instead of 512 KB of shared-memory traffic, no fp32
// dQ accumulator in HBM (no memset, no atomics, no conversion pass) and bit-reproducible gradients.
// One shared-memory tile serves two roles through the descriptor's major-ness bit: K_j is the K-major B of
// S and the MN-major B of dQ; Q_i / dO_i are K-major B operands of S^T / dP^T and MN-major B operands of
// dK / dV.  Nothing is transposed in memory.
//
// Warp roles in both kernels (384 threads): warps 0-7 compute (thread = one TMEM lane x 64 columns),
// warp 8 TMA producer, warp 9 MMA issuer.
#include <cuda.h>
#include <math.h>

#include "fa_internal.h"
#include "fa_ptx.cuh"
#include "fa_tma_host.h"

namespace fa {

namespace {

constexpr int kTile = 128;
constexpr int kBwdThreads = 384;
constexpr int kBoxBytes = 128 * 128;  // 128 rows x 128 B

struct BwdParams {
  int T, Tpad, causal;
  float tau, scale_log2;
  const float* lse_pad;    // [BH][Tpad]  lse * log2(e); +inf on padding rows (=> P = 0 there)
  const float* delta_pad;  // [BH][Tpad]  rowsum(dO o O); 0 on padding rows
  const __nv_bfloat16* q;  // dQ kernel reads its Q / dO tile straight into TMEM
  const __nv_bfloat16* dO;
  void* dq;
  void* dk;
  void* dv;
  int accumulate_dq;
  long long* trace;  // developer timeline (clock64 stamps of CTA (0,0)); nullptr in production
};

long long* g_bwd_trace = nullptr;
#define FA_TRACE(e, it)                                                                                   \
  do {                                                                                                    \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (it) < 16) p.trace[(it) * 16 + (e)] = clock64(); \
  } while (0)

// [128 x D] bf16 tile made of D/64 swizzle boxes.  K-major view: k-step kk of 16 elements.
__device__ __forceinline__ uint64_t kmaj(uint32_t base, int kk) {
  return make_smem_desc(base + (kk >> 2) * kBoxBytes + (kk & 3) * 32, 16, 1024);
}
// MN-major view of the same tile: 16 rows per k-step, 64-element MN atoms one box apart.
__device__ __forceinline__ uint64_t mnmaj(uint32_t base, int kk) {
  return make_smem_desc(base + kk * 2048, kBoxBytes, 1024);
}
// TMEM column of k-step kk of a bf16 A operand whose two 64-wide K halves were written by the two compute
// groups into the first 32 columns of "their" 64-column half of a 128-column region (see kernels).
__device__ __forceinline__ uint32_t split_a_col(int kk) { return (kk >> 2) * 64 + (kk & 3) * 8; }

template <typename OutT>
__device__ __forceinline__ void store_chunk32(OutT* dst, const float* x, bool accumulate);
template <>
__device__ __forceinline__ void store_chunk32<float>(float* dst, const float* x, bool accumulate) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float4 v = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
    if (accumulate) {
      const float4 o = *reinterpret_cast<const float4*>(dst + i);
      v.x += o.x;
      v.y += o.y;
      v.z += o.z;
      v.w += o.w;
    }
    *reinterpret_cast<float4*>(dst + i) = v;
  }
}
template <>
__device__ __forceinline__ void store_chunk32<__nv_bfloat16>(__nv_bfloat16* dst, const float* x, bool accumulate) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = x[i + e];
    if (accumulate) {
      const uint4 o = *reinterpret_cast<const uint4*>(dst + i);
      const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(oh[e]);
        y[2 * e] += f.x;
        y[2 * e + 1] += f.y;
      }
    }
    uint4 v;
    v.x = pack_bf16x2(y[0], y[1]);
    v.y = pack_bf16x2(y[2], y[3]);
    v.z = pack_bf16x2(y[4], y[5]);
    v.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(dst + i) = v;
  }
}

// =================================================================================================
// dQ kernel
// =================================================================================================
template <int D>
struct DqCfg {
  static constexpr int kBoxes = D / 64;
  static constexpr int kTileBytes = kBoxes * kBoxBytes;
  static constexpr int kStages = D >= 128 ? 3 : 4;  // K_j / V_j ring
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kStages * kTileBytes;
  static constexpr int kOffBars = 2 * kStages * kTileBytes;
  static constexpr int kSmemBytes = kOffBars + 256;
  // TMEM columns: Q (D/2) | dO (D/2) | S 128 | dP 128 (dS^bf16 overlays its first 32 columns per half) | dQ D
  static constexpr uint32_t kColQ = 0, kColDO = D / 2, kColS = 128, kColDP = 256, kColDQ = 384;
};

struct DqBarriers {
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint64_t s_full, s_free, dp_full, ds_full, dq_full;
};

template <int D, typename OutT>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_dq_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const BwdParams p) {
  using Cfg = DqCfg<D>;
  constexpr int S = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + Cfg::kOffK;
  uint8_t* sV = smem + Cfg::kOffV;
  DqBarriers& bars = *reinterpret_cast<DqBarriers*>(smem + Cfg::kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kOffBars + 200);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T;
  const int bh = blockIdx.y;
  const int nq = (T + kTile - 1) / kTile;
  const int i = p.causal ? (nq - 1 - (int)blockIdx.x) : (int)blockIdx.x;  // long (late) rows first when causal
  const int nj = p.causal ? min(i + 1, nq) : nq;                          // K/V tiles this Q tile attends to

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&bars.k_full[s], 1);
      mbar_init(&bars.k_empty[s], 1);
      mbar_init(&bars.v_full[s], 1);
      mbar_init(&bars.v_empty[s], 1);
    }
    mbar_init(&bars.s_full, 1);
    mbar_init(&bars.s_free, 8);
    mbar_init(&bars.dp_full, 1);
    mbar_init(&bars.ds_full, 8);
    mbar_init(&bars.dq_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      for (int j = 0; j < nj; ++j) {
        const int st = j % S;
        const uint32_t ph = (j / S) & 1;
        mbar_wait(&bars.k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&bars.k_full[st], Cfg::kTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_3d(sK + st * Cfg::kTileBytes + b * kBoxBytes, &map_k, &bars.k_full[st], b * 64, j * kTile, bh,
                      kEvictLast);
        mbar_wait(&bars.v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&bars.v_full[st], Cfg::kTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_3d(sV + st * Cfg::kTileBytes + b * kBoxBytes, &map_v, &bars.v_full[st], b * 64, j * kTile, bh,
                      kEvictLast);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, 128, 128, 0, 0);  // A: TMEM, B: K-major
      constexpr uint32_t idesc_dq = make_idesc(kFmtBF16, 128, D, 0, 1);   // A: TMEM, B: MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      auto issue_s = [&](uint32_t a_col, uint32_t b_tile, uint32_t d_col) {  // S = Q K^T or dP = dO V^T
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_ts(tmem + d_col, tmem + a_col + kk * 8, kmaj(b_tile, kk), idesc_s, kk > 0);
      };
      auto issue_dq = [&](uint32_t k_tile, bool acc) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tmem + Cfg::kColDQ, tmem + Cfg::kColDP + split_a_col(kk), mnmaj(k_tile, kk), idesc_dq,
                  (acc || kk > 0) ? 1u : 0u);
      };
      // Q / dO of this tile are written to TMEM by the compute warps before the first MMA:
      // phase 0 of ds_full doubles as "Q and dO are in TMEM"
      mbar_wait(&bars.ds_full, 0);
      mbar_wait(&bars.k_full[0], 0);
      tc_fence_after();
      issue_s(Cfg::kColQ, aK, Cfg::kColS);
      umma_commit(&bars.s_full);
      mbar_wait(&bars.v_full[0], 0);
      tc_fence_after();
      issue_s(Cfg::kColDO, aV, Cfg::kColDP);
      umma_commit(&bars.dp_full);
      umma_commit(&bars.v_empty[0]);
      for (int j = 0; j < nj; ++j) {
        const int st = j % S;
        const int jn = j + 1, stn = jn % S;
        const uint32_t phn = (jn / S) & 1;
        if (jn < nj) {
          mbar_wait(&bars.s_free, j & 1);  // compute warps hold S(j) in registers
          mbar_wait(&bars.k_full[stn], phn);
          tc_fence_after();
          FA_TRACE(7, j);
          issue_s(Cfg::kColQ, aK + stn * Cfg::kTileBytes, Cfg::kColS);
          umma_commit(&bars.s_full);
        }
        mbar_wait(&bars.ds_full, (j + 1) & 1);  // dS(j) is in TMEM (phase 0 was the Q/dO hand-off)
        tc_fence_after();
        FA_TRACE(8, j);
        issue_dq(aK + st * Cfg::kTileBytes, j > 0);
        umma_commit(&bars.k_empty[st]);
        if (jn < nj) {
          mbar_wait(&bars.v_full[stn], phn);
          tc_fence_after();
          FA_TRACE(9, j);
          issue_s(Cfg::kColDO, aV + stn * Cfg::kTileBytes, Cfg::kColDP);  // overwrites dS(j): after dQ(j), in order
          umma_commit(&bars.dp_full);
          umma_commit(&bars.v_empty[stn]);
        }
      }
      umma_commit(&bars.dq_full);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ compute warps
    const int g = warp >> 2;               // which half of the 128 key columns
    const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const int qrow = i * kTile + r;
    const bool row_ok = qrow < T;
    // ---- stage this tile's Q (group 0) / dO (group 1) rows into TMEM as bf16 A operands
    {
      const __nv_bfloat16* src = (g == 0 ? p.q : p.dO) + ((size_t)bh * T + (row_ok ? qrow : 0)) * D;
      const uint32_t dst = tmem + (g == 0 ? Cfg::kColQ : Cfg::kColDO) + lane_base;
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 64) {
        uint32_t w[32];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          uint4 x = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok) x = *reinterpret_cast<const uint4*>(src + c0 + v4 * 8);
          w[v4 * 4 + 0] = x.x;
          w[v4 * 4 + 1] = x.y;
          w[v4 * 4 + 2] = x.z;
          w[v4 * 4 + 3] = x.w;
        }
        tmem_st_x32(dst + c0 / 2, w);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ds_full);  // phase 0: operands ready
    }
    const size_t sidx = (size_t)bh * p.Tpad + (size_t)i * kTile + r;
    const float Lr = p.lse_pad[sidx];  // +inf on padding rows
    const float Dr = p.delta_pad[sidx];
    const float c2 = p.scale_log2;
    const uint32_t tS = tmem + Cfg::kColS + lane_base + g * 64;
    const uint32_t tDP = tmem + Cfg::kColDP + lane_base + g * 64;
    for (int j = 0; j < nj; ++j) {
      mbar_wait(&bars.s_full, j & 1);
      tc_fence_after();
      if (threadIdx.x == 0) FA_TRACE(0, j);
      float pr[64];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(pr));
      tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(pr) + 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.s_free);
#pragma unroll
      for (int c = 0; c < 64; ++c) pr[c] = ex2_mixed(fmaf(pr[c], c2, -Lr), c) * p.tau;
      if (p.causal && j == i) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (g * 64 + c > r) pr[c] = 0.f;  // key index > query index
      }
      if (threadIdx.x == 0) FA_TRACE(1, j);
      mbar_wait(&bars.dp_full, j & 1);
      tc_fence_after();
      if (threadIdx.x == 0) FA_TRACE(2, j);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float dp[32];
        tmem_ld_x32(tDP + hf * 32, reinterpret_cast<uint32_t*>(dp));
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2)
          pk[c >> 1] = pack_bf16x2(pr[hf * 32 + c] * (dp[c] - Dr), pr[hf * 32 + c + 1] * (dp[c + 1] - Dr));
        // dS (bf16) for key columns [64g+32hf, +32) -> TMEM columns [64g+16hf, +16) of the dP region: columns
        // this thread has already read, so the two groups never overwrite each other's unread dP
        tmem_st_x16(tDP + hf * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ds_full);
      if (threadIdx.x == 0) FA_TRACE(3, j);
    }
    // ---- epilogue: group g stores dQ columns [g*D/2, (g+1)*D/2)
    mbar_wait(&bars.dq_full, 0);
    tc_fence_after();
    OutT* out = reinterpret_cast<OutT*>(p.dq) + ((size_t)bh * T + (row_ok ? qrow : 0)) * D + g * (D / 2);
    const uint32_t tDQ = tmem + Cfg::kColDQ + lane_base + g * (D / 2);
#pragma unroll
    for (int c0 = 0; c0 < D / 2; c0 += 32) {
      float x[32];
      tmem_ld_x32(tDQ + c0, reinterpret_cast<uint32_t*>(x));
      tmem_wait_ld();
      if (row_ok) store_chunk32<OutT>(out + c0, x, p.accumulate_dq != 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// =================================================================================================
// dK / dV kernel
// =================================================================================================
template <int D>
struct DkvCfg {
  static constexpr int kBoxes = D / 64;
  static constexpr int kTileBytes = kBoxes * kBoxBytes;
  static constexpr int kStages = D >= 128 ? 2 : 4;  // Q_i / dO_i rings
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kTileBytes;
  static constexpr int kOffQ = kOffV + kTileBytes;
  static constexpr int kOffDO = kOffQ + kStages * kTileBytes;
  static constexpr int kOffStats = kOffDO + kStages * kTileBytes;  // [stage][L, delta][128] fp32
  static constexpr int kOffBars = kOffStats + kStages * 2 * 128 * 4;
  static constexpr int kSmemBytes = kOffBars + 256;
  static constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 256 + D;
};

struct DkvBarriers {
  uint64_t kv_full;
  uint64_t q_full[4], q_empty[4], do_full[4], do_empty[4];
  uint64_t s_full, p_full, dp_full, ds_full, dkv_full;
};

template <int D, typename OutT>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                  const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                  const BwdParams p) {
  using Cfg = DkvCfg<D>;
  constexpr int S = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + Cfg::kOffK;
  uint8_t* sV = smem + Cfg::kOffV;
  uint8_t* sQ = smem + Cfg::kOffQ;
  uint8_t* sDO = smem + Cfg::kOffDO;
  float* sStats = reinterpret_cast<float*>(smem + Cfg::kOffStats);
  DkvBarriers& bars = *reinterpret_cast<DkvBarriers*>(smem + Cfg::kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kOffBars + 200);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T;
  const int bh = blockIdx.y;
  const int j = blockIdx.x;  // K/V tile; when causal, tile 0 has the most work and is scheduled first
  const int nq = (T + kTile - 1) / kTile;
  const int i0 = p.causal ? j : 0;
  const int n = nq - i0;  // number of Q tiles this CTA visits (>= 1)

  if (threadIdx.x == 0) {
    mbar_init(&bars.kv_full, 1);
    for (int s = 0; s < S; ++s) {
      mbar_init(&bars.q_full[s], 1);
      mbar_init(&bars.q_empty[s], 1);
      mbar_init(&bars.do_full[s], 1);
      mbar_init(&bars.do_empty[s], 1);
    }
    mbar_init(&bars.s_full, 1);
    mbar_init(&bars.p_full, 8);
    mbar_init(&bars.dp_full, 1);
    mbar_init(&bars.ds_full, 8);
    mbar_init(&bars.dkv_full, 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      tma_prefetch_desc(&map_do);
      mbar_arrive_expect_tx(&bars.kv_full, 2 * Cfg::kTileBytes);
      for (int b = 0; b < Cfg::kBoxes; ++b) {
        tma_load_3d(sK + b * kBoxBytes, &map_k, &bars.kv_full, b * 64, j * kTile, bh, kEvictFirst);
        tma_load_3d(sV + b * kBoxBytes, &map_v, &bars.kv_full, b * 64, j * kTile, bh, kEvictFirst);
      }
      for (int il = 0; il < n; ++il) {
        const int i = i0 + il;
        const int s = il % S;
        const uint32_t ph = (il / S) & 1;
        mbar_wait(&bars.q_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bars.q_full[s], Cfg::kTileBytes + 1024);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_3d(sQ + s * Cfg::kTileBytes + b * kBoxBytes, &map_q, &bars.q_full[s], b * 64, i * kTile, bh,
                      kEvictLast);
        const size_t soff = (size_t)bh * p.Tpad + (size_t)i * kTile;
        bulk_load_1d(sStats + s * 256, p.lse_pad + soff, 512, &bars.q_full[s]);
        bulk_load_1d(sStats + s * 256 + 128, p.delta_pad + soff, 512, &bars.q_full[s]);
        mbar_wait(&bars.do_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bars.do_full[s], Cfg::kTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_3d(sDO + s * Cfg::kTileBytes + b * kBoxBytes, &map_do, &bars.do_full[s], b * 64, i * kTile, bh,
                      kEvictLast);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_sT = make_idesc(kFmtBF16, 128, 128, 0, 0);  // K-major x K-major
      constexpr uint32_t idesc_acc = make_idesc(kFmtBF16, 128, D, 0, 1);   // TMEM x MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aDO = smem_u32(sDO);
      auto issue_sT = [&](uint32_t a_tile, uint32_t b_tile, uint32_t col) {  // S^T or dP^T
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_ss(tmem + col, kmaj(a_tile, kk), kmaj(b_tile, kk), idesc_sT, kk > 0);
      };
      auto issue_acc = [&](uint32_t d_col, uint32_t a_col, uint32_t b_tile, bool acc) {  // dV += P^T dO, dK += dS^T Q
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tmem + d_col, tmem + a_col + split_a_col(kk), mnmaj(b_tile, kk), idesc_acc,
                  (acc || kk > 0) ? 1u : 0u);
      };
      mbar_wait(&bars.kv_full, 0);
      mbar_wait(&bars.q_full[0], 0);
      tc_fence_after();
      issue_sT(aK, aQ, Cfg::kColS);
      umma_commit(&bars.s_full);
      mbar_wait(&bars.do_full[0], 0);
      tc_fence_after();
      issue_sT(aV, aDO, Cfg::kColDP);
      umma_commit(&bars.dp_full);
      for (int il = 0; il < n; ++il) {
        const int s = il % S;
        const int nx = il + 1, sn = nx % S;
        const uint32_t phn = (nx / S) & 1;
        mbar_wait(&bars.p_full, il & 1);
        tc_fence_after();
        FA_TRACE(7, il);
        issue_acc(Cfg::kColDV, Cfg::kColS, aDO + s * Cfg::kTileBytes, il > 0);  // dV += P^T dO_i
        umma_commit(&bars.do_empty[s]);                                         // last reader of dO_i
        if (nx < n) {
          mbar_wait(&bars.q_full[sn], phn);
          tc_fence_after();
          FA_TRACE(8, il);
          issue_sT(aK, aQ + sn * Cfg::kTileBytes, Cfg::kColS);  // S^T(i+1) overwrites P^T(i): after dV(i), in order
          umma_commit(&bars.s_full);
        }
        mbar_wait(&bars.ds_full, il & 1);
        tc_fence_after();
        FA_TRACE(9, il);
        issue_acc(Cfg::kColDK, Cfg::kColDP, aQ + s * Cfg::kTileBytes, il > 0);  // dK += dS^T Q_i
        umma_commit(&bars.q_empty[s]);
        if (nx < n) {
          mbar_wait(&bars.do_full[sn], phn);
          tc_fence_after();
          FA_TRACE(10, il);
          issue_sT(aV, aDO + sn * Cfg::kTileBytes, Cfg::kColDP);  // dP^T(i+1) overwrites dS^T(i): after dK(i)
          umma_commit(&bars.dp_full);
        }
      }
      umma_commit(&bars.dkv_full);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ compute warps
    const int g = warp >> 2;               // which half of the 128 query columns
    const int r = (warp & 3) * 32 + lane;  // key row within the tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + Cfg::kColS + lane_base + g * 64;
    const uint32_t tDP = tmem + Cfg::kColDP + lane_base + g * 64;
    const float c2 = p.scale_log2;
    for (int il = 0; il < n; ++il) {
      const int i = i0 + il;
      const int s = il % S;
      const float* sL = sStats + s * 256 + g * 64;
      const float* sD = sL + 128;
      // ---- P^T = exp2(c * S^T - L)
      mbar_wait(&bars.q_full[s], (il / S) & 1);  // L / delta of this Q tile have landed
      mbar_wait(&bars.s_full, il & 1);
      tc_fence_after();
      if (threadIdx.x == 0) FA_TRACE(0, il);
      float pr[64];
      tmem_ld_x32(tS, reinterpret_cast<uint32_t*>(pr));
      tmem_ld_x32(tS + 32, reinterpret_cast<uint32_t*>(pr) + 32);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
        const float4 L4 = *reinterpret_cast<const float4*>(sL + c);
        pr[c + 0] = ex2_mixed(fmaf(pr[c + 0], c2, -L4.x), c + 0);
        pr[c + 1] = ex2_mixed(fmaf(pr[c + 1], c2, -L4.y), c + 1);
        pr[c + 2] = ex2_mixed(fmaf(pr[c + 2], c2, -L4.z), c + 2);
        pr[c + 3] = ex2_mixed(fmaf(pr[c + 3], c2, -L4.w), c + 3);
      }
      if (p.causal && i == j) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (g * 64 + c < r) pr[c] = 0.f;  // key index > query index
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 2) pk[c >> 1] = pack_bf16x2(pr[c], pr[c + 1]);
        // P^T of query columns [64g, 64g+64) -> TMEM columns [64g, 64g+32): inside the S columns this group
        // has already consumed, so the two groups never touch each other's unread scores
        tmem_st_x32(tS, pk);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.p_full);
      if (threadIdx.x == 0) FA_TRACE(1, il);
      // ---- dS^T = tau * P^T o (dP^T - delta)
      mbar_wait(&bars.dp_full, il & 1);
      tc_fence_after();
      if (threadIdx.x == 0) FA_TRACE(2, il);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float dp[32];
        tmem_ld_x32(tDP + hf * 32, reinterpret_cast<uint32_t*>(dp));
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 D4 = *reinterpret_cast<const float4*>(sD + hf * 32 + c);
          const float a0 = pr[hf * 32 + c + 0] * p.tau * (dp[c + 0] - D4.x);
          const float a1 = pr[hf * 32 + c + 1] * p.tau * (dp[c + 1] - D4.y);
          const float a2 = pr[hf * 32 + c + 2] * p.tau * (dp[c + 2] - D4.z);
          const float a3 = pr[hf * 32 + c + 3] * p.tau * (dp[c + 3] - D4.w);
          pk[(c >> 1) + 0] = pack_bf16x2(a0, a1);
          pk[(c >> 1) + 1] = pack_bf16x2(a2, a3);
        }
        tmem_st_x16(tDP + hf * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ds_full);
      if (threadIdx.x == 0) FA_TRACE(3, il);
    }
    // ---- epilogue: group 0 stores dV_j, group 1 stores dK_j
    mbar_wait(&bars.dkv_full, 0);
    tc_fence_after();
    const int kv_row = j * kTile + r;
    const bool row_ok = kv_row < T;
    OutT* out = reinterpret_cast<OutT*>(g == 0 ? p.dv : p.dk) + ((size_t)bh * T + (row_ok ? kv_row : 0)) * D;
    const uint32_t tAcc = tmem + (g == 0 ? Cfg::kColDV : Cfg::kColDK) + lane_base;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      float x[32];
      tmem_ld_x32(tAcc + c0, reinterpret_cast<uint32_t*>(x));
      tmem_wait_ld();
      if (row_ok) store_chunk32<OutT>(out + c0, x, false);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// delta = rowsum(dO o O), lse -> log2 domain, both padded to a multiple of 128 rows per head
template <int D>
__global__ void fa_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO,
                                   const float* __restrict__ lse, float* __restrict__ lse_pad,
                                   float* __restrict__ delta_pad, int T, int Tpad, size_t total_rows_pad) {
  const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total_rows_pad) return;
  const size_t bh = row / Tpad;
  const int t = (int)(row - bh * Tpad);
  if (t >= T) {
    if (lane == 0) {
      lse_pad[row] = INFINITY;
      delta_pad[row] = 0.f;
    }
    return;
  }
  const size_t base = (bh * T + t) * D;
  float acc = 0.f;
  if constexpr (D == 128) {
    const uint2 a = *reinterpret_cast<const uint2*>(o + base + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(dO + base + lane * 4);
    const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* bh2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(bh2[e]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  } else {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(o + base + lane * 2);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(dO + base + lane * 2);
    const float2 x = __bfloat1622float2(a), y = __bfloat1622float2(b);
    acc = fmaf(x.x, y.x, fmaf(x.y, y.y, 0.f));
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    delta_pad[row] = acc;
    lse_pad[row] = lse[bh * T + t] * 1.4426950408889634f;
  }
}

size_t pad_rows(int T) { return (size_t)((T + kTile - 1) / kTile) * kTile; }
size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <int D, typename OutT>
cudaError_t launch_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* o,
                       const __nv_bfloat16* dO, const float* lse, void* dq, void* dk, void* dv, bool accumulate_dq,
                       void* workspace, int BH, int T, bool causal, cudaStream_t st) {
  const size_t Tpad = pad_rows(T);
  char* ws = static_cast<char*>(workspace);
  float* lse_pad = reinterpret_cast<float*>(ws);
  ws += align256((size_t)BH * Tpad * 4);
  float* delta_pad = reinterpret_cast<float*>(ws);
  cudaError_t e;
  {
    const size_t rows = (size_t)BH * Tpad;
    const int warps = 8;
    fa_bwd_prep_kernel<D><<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, st>>>(o, dO, lse, lse_pad, delta_pad,
                                                                                          T, (int)Tpad, rows);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  CUtensorMap mq, mk, mv, mdo;
  if (make_tmap_3d(&mq, q, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mk, k, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mv, v, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mdo, dO, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  e = cudaFuncSetAttribute(fa_bwd_dq_kernel<D, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, DqCfg<D>::kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fa_bwd_dkv_kernel<D, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           DkvCfg<D>::kSmemBytes);
  if (e != cudaSuccess) return e;
  BwdParams p;
  p.T = T;
  p.Tpad = (int)Tpad;
  p.causal = causal ? 1 : 0;
  p.tau = sqrtf(1.0f / (float)D);
  p.scale_log2 = p.tau * 1.4426950408889634f;
  p.lse_pad = lse_pad;
  p.delta_pad = delta_pad;
  p.q = q;
  p.dO = dO;
  p.dq = dq;
  p.dk = dk;
  p.dv = dv;
  p.accumulate_dq = accumulate_dq ? 1 : 0;
  p.trace = g_bwd_trace;
  dim3 grid((unsigned)(Tpad / kTile), BH);
  fa_bwd_dkv_kernel<D, OutT><<<grid, kBwdThreads, DkvCfg<D>::kSmemBytes, st>>>(mq, mk, mv, mdo, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (p.trace) p.trace += 256;
  fa_bwd_dq_kernel<D, OutT><<<grid, kBwdThreads, DqCfg<D>::kSmemBytes, st>>>(mk, mv, p);
  return cudaGetLastError();
}

}  // namespace

bool tc_bwd_supported(int d) { return d == 64 || d == 128; }
void tc_bwd_set_trace(long long* dev_buf) { g_bwd_trace = dev_buf; }

size_t tc_bwd_workspace_bytes(int BH, int T, int d) {
  (void)d;
  return 2 * align256((size_t)BH * pad_rows(T) * 4);  // lse (log2 domain) and delta, padded to 128-row tiles
}

cudaError_t tc_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* o,
                   const __nv_bfloat16* dO, const float* lse, void* dq, void* dk, void* dv, bool grads_f32,
                   bool accumulate_dq, void* workspace, int BH, int T, int d, bool causal, cudaStream_t st) {
  if (BH > 65535 || !workspace) return cudaErrorInvalidValue;
  if (d == 128) {
    return grads_f32
               ? launch_bwd<128, float>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, st)
               : launch_bwd<128, __nv_bfloat16>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, st);
  }
  if (d == 64) {
    return grads_f32
               ? launch_bwd<64, float>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, st)
               : launch_bwd<64, __nv_bfloat16>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace fa
