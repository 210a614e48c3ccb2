// fa_bwd_tf32_sm100.cu -- flash-attention backward on fp32 storage with tf32 tensor cores (head dim 32 and 64), sm_100a.
//
// The fp32 / TF32 path of the backward (north_star: max-abs 1e-3 against the reference): replaces the CUDA-core
// kernels of fa_simt.cu for the head dimension of the reference's machine-translation model (n_embd 256, 8 heads: d = 32),
// where they were 42 % of a training step, and for the head dimension of the reference's own kernel tests (d = 64,
// kernel_tests/test_flashattn_bw.py:205-210).  Same algorithm and pipeline as fa_bwd_sm100.cu (one CTA = one 128-row K/V
// tile streaming the Q/dO tiles, scores kept transposed so that P^T / dS^T are TMEM A operands), with what tf32 changes:
//   * kind::tf32, UMMA_K = 8; fp32 tiles are read by TMA as they are (no conversion pass);
//   * a 32-bit operand can be K-major in the 128B swizzle but MN-major only in the 128B swizzle with 32-byte atoms,
//     and a K-major operand faults in that layout (tools/umma_probe) -- so every tile that is used in both major-nesses
//     (Q, dO, K) is loaded TWICE, once per layout;
//   * dS^T goes to TMEM (tf32, A operand of dK) AND to shared memory in the MN-major 32-byte-atom layout (A operand of
//     dQ = dS K, whose M dimension is the query index); TMEM has room for every matrix
//     (S | dP | dV | dK | dQ | dS^T = 480 columns at d = 32, 384 at d = 64), so nothing aliases;
//   * fp32 gradients: dK / dV straight from TMEM, dQ reduce-added by TMA into the caller's buffer -- no conversion pass.
// Tile shapes (template <kD, kTQ>): d = 32 streams 128-query tiles (a tile copy is 16 KB, everything double-buffered);
// d = 64 streams 64-query tiles (twice the bytes per row: 128-query tiles would need 288 KB of shared memory), the
// dQ product then has M = 64 -- its accumulator sits in TMEM lanes {0-15, 32-47, 64-79, 96-111} (16 rows per
// sub-partition), which is what the drain warps read.
#include <cuda.h>
#include <math.h>

#include "fa_internal.h"
#include "fa_ptx.cuh"
#include "fa_tma_host.h"

namespace fa {

namespace {

constexpr int kTile = 128;     // key rows per CTA (= TMEM lanes)
constexpr int kThreads = 512;
constexpr int kBoxKV = 128 * 128;  // one K / V box: 128 rows x 32 fp32

template <int kD, int kTQ>
struct Cfg {
  static_assert((kD == 32 && kTQ == 128) || (kD == 64 && kTQ == 64), "supported: d = 32 with 128-query tiles, d = 64 with 64-query tiles");
  static constexpr int kBoxesD = kD / 32;               // 128-byte boxes per tile row
  static constexpr int kKVBytes = kBoxesD * kBoxKV;     // one K / V tile copy
  static constexpr int kBoxQ = kTQ * 128;               // one Q / dO box: kTQ rows x 32 fp32
  static constexpr int kQBytes = kBoxesD * kBoxQ;       // one Q / dO tile copy (16 KB in both configurations)
  static constexpr int kQStages = kD == 32 ? 2 : 1;
  static constexpr int kDSBoxes = kTQ / 32;             // dS^T [128 keys][kTQ queries] fp32: boxes of 32 queries, 16 KB each
  static constexpr int kCT = kTQ / 2;                   // query columns per compute thread (two warp groups)
  // shared-memory map (bytes); every tile 1024-B aligned
  static constexpr int kOffKk = 0;                                  // K, K-major   (A of S^T)
  static constexpr int kOffKm = kOffKk + kKVBytes;                  // K, MN-major  (B of dQ)
  static constexpr int kOffVk = kOffKm + kKVBytes;                  // V, K-major   (A of dP^T)
  static constexpr int kOffQk = kOffVk + kKVBytes;                  // Q, K-major   (B of S^T), kQStages
  static constexpr int kOffQm = kOffQk + kQStages * kQBytes;        // Q, MN-major  (B of dK), kQStages
  static constexpr int kOffDOk = kOffQm + kQStages * kQBytes;       // dO, K-major  (B of dP^T)
  static constexpr int kOffDOm = kOffDOk + kQBytes;                 // dO, MN-major (B of dV)
  static constexpr int kOffDS = kOffDOm + kQBytes;                  // dS^T, MN-major A of dQ
  static constexpr int kOffStage = kOffDS + kDSBoxes * kBoxKV;      // dQ staging: kTQ rows x kD fp32 (kBoxesD boxes)
  static constexpr int kOffStats = kOffStage + kBoxesD * kBoxQ;     // [2 stages][L, -delta][kTQ] fp32
  static constexpr int kOffBars = kOffStats + 2 * 2 * kTQ * 4;
  static constexpr int kSmemBytes = kOffBars + 256;
  static_assert(kSmemBytes <= 232448, "shared-memory budget");
  // TMEM columns
  static constexpr uint32_t kColS = 0, kColDP = kTQ, kColDV = 2 * kTQ, kColDK = kColDV + kD, kColDQ = kColDK + kD, kColDS = kColDQ + kD;
  static_assert(kColDS + kTQ <= 512, "TMEM budget");
};

struct Params {
  int T, Tpad, causal;
  float tau, scale_log2;
  const float* lse_pad;    // [BH][Tpad]  lse * log2(e); +inf on padding rows (=> P = 0 there)
  const float* delta_pad;  // [BH][Tpad]  -rowsum(dO o O); 0 on padding rows
  float* dk;
  float* dv;
};

struct Barriers {
  uint64_t kv_full;
  uint64_t q_full[2], q_empty[2], do_full, do_empty;
  uint64_t s_full, p_full, dp_full, ds_full, ds_empty, dq_full, dq_drained, dkv_full;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

template <int kD, int kTQ>
__global__ void __launch_bounds__(kThreads, 1)
fa_bwd_tf32_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qm,
                   const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_km,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_dom, const __grid_constant__ CUtensorMap map_dq, const Params p) {
  using C = Cfg<kD, kTQ>;
  constexpr int QS = C::kQStages;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* sStats = reinterpret_cast<float*>(smem + C::kOffStats);
  uint8_t* sDS = smem + C::kOffDS;
  uint8_t* sStage = smem + C::kOffStage;
  Barriers& bars = *reinterpret_cast<Barriers*>(smem + C::kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kOffBars + 192);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T;
  const int bh = blockIdx.y;
  const int j = blockIdx.x;
  const int nq = (T + kTQ - 1) / kTQ;
  const int i0 = p.causal ? (j * kTile) / kTQ : 0;  // first Q tile with a query >= the first key of this CTA
  const int n = nq - i0;

  if (threadIdx.x == 0) {
    mbar_init(&bars.kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.q_full[s], 1);
      mbar_init(&bars.q_empty[s], 1);
    }
    mbar_init(&bars.do_full, 1);
    mbar_init(&bars.do_empty, 1);
    mbar_init(&bars.s_full, 1);
    mbar_init(&bars.p_full, 8);
    mbar_init(&bars.dp_full, 1);
    mbar_init(&bars.ds_full, 8);
    mbar_init(&bars.ds_empty, 1);
    mbar_init(&bars.dq_full, 1);
    mbar_init(&bars.dq_drained, 4);
    mbar_init(&bars.dkv_full, 1);
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 12) reg_dealloc<88>();
  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_qm);
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_km);
      tma_prefetch_desc(&map_v);
      tma_prefetch_desc(&map_do);
      tma_prefetch_desc(&map_dom);
      mbar_arrive_expect_tx(&bars.kv_full, 3 * C::kKVBytes);
      for (int b = 0; b < C::kBoxesD; ++b) {
        tma_load_3d(smem + C::kOffKk + b * kBoxKV, &map_k, &bars.kv_full, b * 32, j * kTile, bh, kEvictFirst);
        tma_load_3d(smem + C::kOffKm + b * kBoxKV, &map_km, &bars.kv_full, b * 32, j * kTile, bh, kEvictFirst);
        tma_load_3d(smem + C::kOffVk + b * kBoxKV, &map_v, &bars.kv_full, b * 32, j * kTile, bh, kEvictFirst);
      }
      // Q(il) (both layouts + L / -delta) and dO(il).  Two Q stages (d = 32): Q(il) then dO(il), so that Q(il+1) -- which
      // S(il+1) wants right after dV(il) -- is already under way when dO(il+1) has to wait for dV(il).  One Q stage (d = 64):
      // the stage is only released at the END of a step (after dK(il)) while dO(il+1) is needed earlier in that same step
      // (dP(il+1) is issued before dK(il)), so dO(il) must be requested BEFORE Q(il): the other order deadlocks the
      // producer against the MMA warp.
      auto load_q = [&](int il) {
        const int i = i0 + il, s = il % QS;
        mbar_wait(&bars.q_empty[s], ((il / QS) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.q_full[s], 2 * C::kQBytes + 2 * kTQ * 4);
        for (int b = 0; b < C::kBoxesD; ++b) {
          tma_load_3d(smem + C::kOffQk + s * C::kQBytes + b * C::kBoxQ, &map_q, &bars.q_full[s], b * 32, i * kTQ, bh, kEvictLast);
          tma_load_3d(smem + C::kOffQm + s * C::kQBytes + b * C::kBoxQ, &map_qm, &bars.q_full[s], b * 32, i * kTQ, bh, kEvictLast);
        }
        // L / -delta ride on the Q barrier; their two slots alternate with il (a Q stage is released by dK(il - QS), i.e.
        // after the compute warps have finished with the statistics that lived in this slot)
        const size_t soff = (size_t)bh * p.Tpad + (size_t)i * kTQ;
        bulk_load_1d(sStats + (il & 1) * 2 * kTQ, p.lse_pad + soff, kTQ * 4, &bars.q_full[s]);
        bulk_load_1d(sStats + (il & 1) * 2 * kTQ + kTQ, p.delta_pad + soff, kTQ * 4, &bars.q_full[s]);
      };
      auto load_do = [&](int il) {
        const int i = i0 + il;
        mbar_wait(&bars.do_empty, (il & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.do_full, 2 * C::kQBytes);
        for (int b = 0; b < C::kBoxesD; ++b) {
          tma_load_3d(smem + C::kOffDOk + b * C::kBoxQ, &map_do, &bars.do_full, b * 32, i * kTQ, bh, kEvictLast);
          tma_load_3d(smem + C::kOffDOm + b * C::kBoxQ, &map_dom, &bars.do_full, b * 32, i * kTQ, bh, kEvictLast);
        }
      };
      load_q(0);
      load_do(0);
      for (int il = 1; il < n; ++il) {
        if (QS == 2) { load_q(il); load_do(il); } else { load_do(il); load_q(il); }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, `_w` forms)
    // REDUX inside this branch: a shared-memory load is not provably warp-uniform for ptxas, and with ONE non-uniform operand
    // (every MMA takes tmem + const) it builds ALL operands of the issue loops in vector registers and wraps each UTCHMMA
    // into ELECT + four R2UR.BROADCAST (~16 instructions per MMA instead of ~3 on the uniform datapath)
    const uint32_t tmem = __reduce_max_sync(0xffffffffu, *tmem_slot);
    constexpr uint32_t idesc_sT = make_idesc(kFmtTF32, 128, kTQ, 0, 0);  // K-major x K-major
    constexpr uint32_t idesc_acc = make_idesc(kFmtTF32, 128, kD, 0, 1);  // TMEM x MN-major
    constexpr uint32_t idesc_dq = make_idesc(kFmtTF32, kTQ, kD, 1, 1);   // MN-major x MN-major, M = kTQ queries
    const uint32_t base = smem_u32(smem);
    // K-major [rows x kD fp32]: a k-step (8 tf32) is 32 B inside the 128-B row, 4 k-steps per 32-column box.
    // MN-major (32-byte atoms): a k-step is 8 "K" rows = 1 KB, 4-row groups 512 B apart, 32-element MN atoms one box apart.
    auto kmaj0 = [](uint32_t a) { return make_smem_desc(a, 16, 1024); };
    auto mnmaj0 = [](uint32_t a, uint32_t atom_stride) { return make_smem_desc(a, atom_stride, 512, kDescSw128Base32); };
    const uint64_t dKk = kmaj0(base + C::kOffKk), dVk = kmaj0(base + C::kOffVk), dKm = mnmaj0(base + C::kOffKm, kBoxKV),
                   dDSm = mnmaj0(base + C::kOffDS, kBoxKV);
    auto issue_sT = [&](uint64_t a0, uint64_t b0, uint32_t col) {  // S^T / dP^T: A = K / V tile, B = Q / dO tile, contraction over d
#pragma unroll
      for (int kk = 0; kk < kD / 8; ++kk)
        umma_ss_w<true>(tmem + col, desc_advance(a0, (kk >> 2) * kBoxKV + (kk & 3) * 32), desc_advance(b0, (kk >> 2) * C::kBoxQ + (kk & 3) * 32),
                        idesc_sT, kk > 0);
    };
    auto issue_ts = [&](uint32_t d_col, uint32_t a_col, uint64_t b0, bool acc) {  // dV / dK: A = P^T / dS^T in TMEM, contraction over the queries
#pragma unroll
      for (int kk = 0; kk < kTQ / 8; ++kk)
        umma_ts_w<true>(tmem + d_col, tmem + a_col + kk * 8, desc_advance(b0, kk * 1024), idesc_acc, (acc || kk > 0) ? 1u : 0u);
    };
    auto issue_dq = [&]() {  // dQ = dS K: contraction over the 128 keys
#pragma unroll
      for (int kk = 0; kk < kTile / 8; ++kk)
        umma_ss_w<true>(tmem + C::kColDQ, desc_advance(dDSm, kk * 1024), desc_advance(dKm, kk * 1024), idesc_dq, kk > 0);
    };
    auto qk = [&](int il) { return kmaj0(base + C::kOffQk + (il % QS) * C::kQBytes); };
    auto qm = [&](int il) { return mnmaj0(base + C::kOffQm + (il % QS) * C::kQBytes, C::kBoxQ); };
    const uint64_t dDOk = kmaj0(base + C::kOffDOk), dDOm = mnmaj0(base + C::kOffDOm, C::kBoxQ);

    mbar_wait_w(&bars.kv_full, 0);
    mbar_wait_w(&bars.q_full[0], 0);
    tc_fence_after();
    issue_sT(dKk, qk(0), C::kColS);
    umma_commit_w(&bars.s_full);
    mbar_wait_w(&bars.do_full, 0);
    tc_fence_after();
    issue_sT(dVk, dDOk, C::kColDP);
    umma_commit_w(&bars.dp_full);
    for (int il = 0; il < n; ++il) {
      mbar_wait_w(&bars.p_full, il & 1);
      tc_fence_after();
      issue_ts(C::kColDV, C::kColS, dDOm, il > 0);  // dV += P^T dO
      umma_commit_w(&bars.do_empty);                // dV(il) and dP(il) (issued earlier) are the last readers of dO(il)
      if (QS == 2 && il + 1 < n) {
        mbar_wait_w(&bars.q_full[(il + 1) % QS], ((il + 1) / QS) & 1);
        tc_fence_after();
        issue_sT(dKk, qk(il + 1), C::kColS);  // P(il) has been consumed by the MMAs just issued (in-order pipe)
        umma_commit_w(&bars.s_full);
      }
      mbar_wait_w(&bars.ds_full, il & 1);  // all of dP(il) has been read, dS^T(il) is in TMEM and in shared memory
      tc_fence_after();
      if (il + 1 < n) {
        mbar_wait_w(&bars.do_full, (il + 1) & 1);
        tc_fence_after();
        issue_sT(dVk, dDOk, C::kColDP);  // dP(il+1) first: it is what the compute warps wait for next
        umma_commit_w(&bars.dp_full);
      }
      issue_ts(C::kColDK, C::kColDS, qm(il), il > 0);  // dK += dS^T Q
      if (il > 0) mbar_wait_w(&bars.dq_drained, (il - 1) & 1);
      tc_fence_after();
      issue_dq();  // dQ(il) = dS K
      umma_commit_w(&bars.dq_full);
      umma_commit_w(&bars.ds_empty);
      umma_commit_w(&bars.q_empty[il % QS]);
      if (QS == 1 && il + 1 < n) {
        // single Q stage (d = 64): Q(il+1) can only be requested once dK(il) has released the stage
        mbar_wait_w(&bars.q_full[0], (il + 1) & 1);
        tc_fence_after();
        issue_sT(dKk, qk(il + 1), C::kColS);
        umma_commit_w(&bars.s_full);
      }
    }
    umma_commit_w(&bars.dkv_full);
  } else if (warp < 8) {
    // ------------------------------------------------------------------ compute warps
    reg_alloc<168>();
    constexpr int CT = C::kCT;             // this thread's query columns: [g * CT, g * CT + CT)
    const int g = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // key row within the tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + C::kColS + lane_base + g * CT;
    const uint32_t tDP = tmem + C::kColDP + lane_base + g * CT;
    const uint32_t tDS = tmem + C::kColDS + lane_base + g * CT;
    const float c2 = p.scale_log2;
    for (int il = 0; il < n; ++il) {
      const int i = i0 + il;
      const float* sL = sStats + (il & 1) * 2 * kTQ + g * CT;
      const float* sD = sL + kTQ;
      // ---- P^T = exp2(c S^T - L)
      mbar_wait(&bars.q_full[il % C::kQStages], (il / C::kQStages) & 1);  // L / delta of this Q tile have landed
      mbar_wait(&bars.s_full, il & 1);
      tc_fence_after();
      float pr[CT];
      {
        uint32_t* u = reinterpret_cast<uint32_t*>(pr);
#pragma unroll
        for (int h = 0; h < CT / 32; ++h) tmem_ld_x32(tS + h * 32, u + h * 32);
        tmem_wait_ld();
      }
#pragma unroll
      for (int c = 0; c < CT; c += 4) {
        const float4 L4 = *reinterpret_cast<const float4*>(sL + c);
        pr[c + 0] = ex2_approx(fmaf(pr[c + 0], c2, -L4.x));
        pr[c + 1] = ex2_approx(fmaf(pr[c + 1], c2, -L4.y));
        pr[c + 2] = ex2_approx(fmaf(pr[c + 2], c2, -L4.z));
        pr[c + 3] = ex2_approx(fmaf(pr[c + 3], c2, -L4.w));
      }
      if (p.causal && i * kTQ < j * kTile + kTile) {  // the tile reaches above the diagonal
        const int key = j * kTile + r, q0 = i * kTQ + g * CT;
#pragma unroll
        for (int c = 0; c < CT; ++c)
          if (q0 + c < key) pr[c] = 0.f;  // key index > query index
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < CT / 32; ++h) {
#pragma unroll
          for (int c = 0; c < 32; ++c) pk[c] = to_tf32(pr[h * 32 + c]);
          tmem_st_x32(tS + h * 32, pk);  // in place: this thread group has consumed these columns of S
        }
        // tau folded into the kept P values: dS^T = (tau P^T) o (dP^T - delta)
        const float2 tau2 = make_float2(p.tau, p.tau);
#pragma unroll
        for (int c = 0; c < CT; c += 2) {
          const float2 t2 = fmul2(make_float2(pr[c], pr[c + 1]), tau2);
          pr[c] = t2.x;
          pr[c + 1] = t2.y;
        }
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.p_full);
      // ---- dS^T = (tau P^T) o (dP^T - delta)  -> TMEM (A of dK) and shared memory (MN-major A of dQ)
      mbar_wait(&bars.dp_full, il & 1);
      tc_fence_after();
      if (il > 0) mbar_wait(&bars.ds_empty, (il - 1) & 1);  // dK(i-1) / dQ(i-1) have finished reading dS^T
#pragma unroll
      for (int hf = 0; hf < CT / 32; ++hf) {
        float dp[32];
        tmem_ld_x32(tDP + hf * 32, reinterpret_cast<uint32_t*>(dp));
        tmem_wait_ld();
        uint32_t dsb[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 D4 = *reinterpret_cast<const float4*>(sD + hf * 32 + c);
          const float2 a = fmul2(make_float2(pr[hf * 32 + c], pr[hf * 32 + c + 1]), fadd2(make_float2(dp[c], dp[c + 1]), make_float2(D4.x, D4.y)));
          const float2 b = fmul2(make_float2(pr[hf * 32 + c + 2], pr[hf * 32 + c + 3]), fadd2(make_float2(dp[c + 2], dp[c + 3]), make_float2(D4.z, D4.w)));
          dsb[c] = to_tf32(a.x);
          dsb[c + 1] = to_tf32(a.y);
          dsb[c + 2] = to_tf32(b.x);
          dsb[c + 3] = to_tf32(b.y);
        }
        tmem_st_x32(tDS + hf * 32, dsb);
        // box (g * CT) / 32 + hf = query columns [g CT + 32 hf, +32): row r = 128 B, 32-byte chunks XORed with (r & 3)
        uint8_t* row = sDS + ((g * CT) / 32 + hf) * kBoxKV + r * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<uint4*>(row + ((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4))) =
              make_uint4(dsb[c16 * 4], dsb[c16 * 4 + 1], dsb[c16 * 4 + 2], dsb[c16 * 4 + 3]);
      }
      tmem_wait_st();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ds_full);
    }
    // ---- epilogue: group 0 stores dV_j, group 1 stores dK_j (fp32, kD columns per thread)
    mbar_wait(&bars.dkv_full, 0);
    tc_fence_after();
    const int kv_row = j * kTile + r;
#pragma unroll
    for (int h = 0; h < kD / 32; ++h) {
      uint32_t rr[32];
      tmem_ld_x32(tmem + (g == 0 ? C::kColDV : C::kColDK) + lane_base + h * 32, rr);  // (warp-collective: every lane takes part)
      tmem_wait_ld();
      if (kv_row < T) {
        float* out = (g == 0 ? p.dv : p.dk) + ((size_t)bh * T + kv_row) * kD + h * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4) *reinterpret_cast<uint4*>(out + c) = make_uint4(rr[c], rr[c + 1], rr[c + 2], rr[c + 3]);
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ dQ drain warps
    reg_dealloc<88>();
    const int w = warp - 8;
    // kTQ = 128: lane = query row 32 w + lane.  kTQ = 64 (an M = 64 accumulator): rows 16 w + lane live in the first 16 lanes
    // of every 32-lane TMEM sub-partition, the other 16 lanes are not part of the tile.
    const bool owns = kTQ == 128 || lane < 16;
    const int r = kTQ == 128 ? w * 32 + lane : w * 16 + (lane & 15);
    const uint32_t tDQ = tmem + C::kColDQ + (uint32_t(w * 32) << 16);
    const bool leader = (warp == 8 && lane == 0);
    for (int il = 0; il < n; ++il) {
      const int i = i0 + il;
      mbar_wait(&bars.dq_full, il & 1);
      tc_fence_after();
      uint32_t rr[kD];
#pragma unroll
      for (int h = 0; h < kD / 32; ++h) tmem_ld_x32(tDQ + h * 32, rr + h * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.dq_drained);
      if (leader) tma_store_wait_read<0>();  // the staging slabs are no longer being read
      named_bar_sync(1, 128);
      if (owns) {
#pragma unroll
        for (int h = 0; h < kD / 32; ++h)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(sStage + h * C::kBoxQ + r * 128 + ((ch ^ (r & 7)) << 4)) =
                make_uint4(rr[h * 32 + ch * 4], rr[h * 32 + ch * 4 + 1], rr[h * 32 + ch * 4 + 2], rr[h * 32 + ch * 4 + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        for (int h = 0; h < kD / 32; ++h) tma_reduce_add_3d(&map_dq, sStage + h * C::kBoxQ, h * 32, i * kTQ, bh);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, 512);
}

// -delta = -rowsum(dO o O), lse -> log2 domain, both padded to a multiple of 128 rows per head; optionally zero-fills dQ
// (the main kernel reduce-adds into it).  The same pass writes copies of Q, K, V, dO ROUNDED to tf32 (cvt.rna): the tensor
// core truncates fp32 operands (one-sided, up to 2^-10 relative), which biases sums of same-sign terms -- dV = P^T dO with
// P >= 0 came out 5e-4 low relative, 1.5e-3 absolute at |dO| <= 1; rounding to nearest first halves the error and removes
// the bias, at the price of one more memory-bound pass over the four tensors.  kD / 4 lanes own one row.
__device__ __forceinline__ float4 round4_tf32(float4 x) {
  return make_float4(__uint_as_float(to_tf32(x.x)), __uint_as_float(to_tf32(x.y)), __uint_as_float(to_tf32(x.z)),
                     __uint_as_float(to_tf32(x.w)));
}
template <int kD>
__global__ void __launch_bounds__(256)
fa_bwd_tf32_prep_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                        const float* __restrict__ o, const float* __restrict__ dO, const float* __restrict__ lse,
                        float* __restrict__ qr, float* __restrict__ kr, float* __restrict__ vr, float* __restrict__ dor,
                        float* __restrict__ lse_pad, float* __restrict__ delta_pad, float* __restrict__ dq_zero, int T, int Tpad,
                        size_t total_rows_pad) {
  constexpr int kLanes = kD / 4, kRows = 32 / kLanes;  // lanes per row, rows per warp pass
  const int lane = threadIdx.x & 31;
  const int sub = lane / kLanes, li = lane % kLanes;
  const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t row = warp_global * kRows + sub; row < total_rows_pad; row += nwarps * kRows) {
    const size_t bh = row / Tpad;
    const int t = (int)(row - bh * Tpad);
    float acc = 0.f;
    if (t < T) {
      const size_t base = (bh * T + t) * kD + li * 4;
      const float4 a = *reinterpret_cast<const float4*>(o + base);
      const float4 b = *reinterpret_cast<const float4*>(dO + base);
      acc = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      *reinterpret_cast<float4*>(dor + base) = round4_tf32(b);
      *reinterpret_cast<float4*>(qr + base) = round4_tf32(*reinterpret_cast<const float4*>(q + base));
      *reinterpret_cast<float4*>(kr + base) = round4_tf32(*reinterpret_cast<const float4*>(k + base));
      *reinterpret_cast<float4*>(vr + base) = round4_tf32(*reinterpret_cast<const float4*>(v + base));
      if (dq_zero) *reinterpret_cast<float4*>(dq_zero + base) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int off = kLanes / 2; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (li == 0) {
      delta_pad[row] = t < T ? -acc : 0.f;
      lse_pad[row] = t < T ? lse[bh * T + t] * 1.4426950408889634f : INFINITY;
    }
  }
}

size_t pad_rows(int T) { return (size_t)((T + kTile - 1) / kTile) * kTile; }
size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <int kD, int kTQ>
cudaError_t launch(const float* q, const float* k, const float* v, const float* o, const float* dO, const float* lse, float* dq,
                   float* dk, float* dv, bool accumulate_dq, void* workspace, int BH, int T, bool causal, cudaStream_t st) {
  using C = Cfg<kD, kTQ>;
  const size_t Tpad = pad_rows(T);
  char* ws = static_cast<char*>(workspace);
  float* lse_pad = reinterpret_cast<float*>(ws);
  ws += align256((size_t)BH * Tpad * 4);
  float* delta_pad = reinterpret_cast<float*>(ws);
  ws += align256((size_t)BH * Tpad * 4);
  float* rounded[4];  // q, k, v, dO
  for (auto& r : rounded) {
    r = reinterpret_cast<float*>(ws);
    ws += align256((size_t)BH * T * kD * 4);
  }
  {
    const size_t rows = (size_t)BH * Tpad;
    constexpr int kRowsPerBlock = 8 * (32 / (kD / 4));
    size_t blocks = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
    if (blocks > 148 * 32) blocks = 148 * 32;
    fa_bwd_tf32_prep_kernel<kD><<<(unsigned)blocks, 256, 0, st>>>(q, k, v, o, dO, lse, rounded[0], rounded[1], rounded[2], rounded[3],
                                                                lse_pad, delta_pad, accumulate_dq ? nullptr : dq, T, (int)Tpad, rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  q = rounded[0];
  k = rounded[1];
  v = rounded[2];
  dO = rounded[3];
  CUtensorMap mq, mqm, mk, mkm, mv, mdo, mdom, mdq;
  // boxes of 32 columns (128 B); K-major copies: 128B swizzle (mode 1); MN-major copies: 128B swizzle with 32-byte atoms (mode 2)
  if (make_tmap_3d(&mq, q, 4, kD, T, BH, 32, kTQ, 1) || make_tmap_3d(&mqm, q, 4, kD, T, BH, 32, kTQ, 2) ||
      make_tmap_3d(&mk, k, 4, kD, T, BH, 32, kTile, 1) || make_tmap_3d(&mkm, k, 4, kD, T, BH, 32, kTile, 2) ||
      make_tmap_3d(&mv, v, 4, kD, T, BH, 32, kTile, 1) || make_tmap_3d(&mdo, dO, 4, kD, T, BH, 32, kTQ, 1) ||
      make_tmap_3d(&mdom, dO, 4, kD, T, BH, 32, kTQ, 2) || make_tmap_3d(&mdq, dq, 4, kD, T, BH, 32, kTQ, 1))
    return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(fa_bwd_tf32_kernel<kD, kTQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  if (e != cudaSuccess) return e;
  Params p;
  p.T = T;
  p.Tpad = (int)Tpad;
  p.causal = causal ? 1 : 0;
  p.tau = sqrtf(1.0f / (float)kD);
  p.scale_log2 = p.tau * 1.4426950408889634f;
  p.lse_pad = lse_pad;
  p.delta_pad = delta_pad;
  p.dk = dk;
  p.dv = dv;
  dim3 grid((unsigned)(Tpad / kTile), BH);
  fa_bwd_tf32_kernel<kD, kTQ><<<grid, kThreads, C::kSmemBytes, st>>>(mq, mqm, mk, mkm, mv, mdo, mdom, mdq, p);
  return cudaGetLastError();
}

}  // namespace

bool tc_bwd_tf32_supported(int d) { return d == 32 || d == 64; }

// padded L and -delta + tf32-rounded copies of Q, K, V, dO
size_t tc_bwd_tf32_workspace_bytes(int BH, int T, int d) {
  return 2 * align256((size_t)BH * pad_rows(T) * 4) + 4 * align256((size_t)BH * T * d * 4);
}

cudaError_t tc_bwd_tf32(const float* q, const float* k, const float* v, const float* o, const float* dO, const float* lse,
                        float* dq, float* dk, float* dv, bool accumulate_dq, void* workspace, int BH, int T, int d, bool causal,
                        cudaStream_t st) {
  if (BH > 65535 || !workspace) return cudaErrorInvalidValue;
  if (d == 32) return launch<32, 128>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, st);
  if (d == 64) return launch<64, 64>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, st);
  return cudaErrorInvalidValue;
}

}  // namespace fa
