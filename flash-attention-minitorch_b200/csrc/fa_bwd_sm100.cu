// fa_bwd_sm100.cu -- flash-attention backward for sm_100a: tcgen05.mma + TMEM + TMA.
//
// Replaces the reference kernels src/flash_attn2_bw.cu:21-263 (FA-2 backward: P = exp(S - L), atomicAdd
// into dQ) and src/flash_attn_bw.cu:21-261 (FA-1 backward; its (l, m) pair is folded into L up front).
//
// A work item is one 128-row K/V tile j of one (batch, head): the CTA streams the Q/dO tiles i that attend to it
// (i >= j when causal).  The kernel is persistent -- one CTA per SM draws items from a counter and keeps its pipelines
// running across them, see the comment above fa_bwd_kernel.  All five contractions of a tile pair run on the tensor pipe, with the
// score matrices held TRANSPOSED (lanes = key rows) so that P^T is directly the TMEM A operand of dV:
//     S^T  = K_j  Q_i^T          (smem x smem  -> TMEM)        P^T = exp2(c S^T - L_i)        [registers]
//     dP^T = V_j  dO_i^T         (smem x smem  -> TMEM)        dS^T = tau P^T o (dP^T - D_i)   [registers]
//     dV_j += P^T  dO_i          (TMEM x smem  -> TMEM, accumulates over i)
//     dK_j += dS^T Q_i           (smem x smem  -> TMEM, accumulates over i)
//     dQ_i  = dS   K_j           (smem(MN-major view of dS^T) x smem -> TMEM) -> fp32 TMA reduce-add to HBM
// The same shared-memory tile serves two roles through the descriptor's major-ness bit: Q_i / dO_i are
// K-major B operands of S^T / dP^T and MN-major B operands of dK / dV; dS^T is the K-major A of dK and the
// MN-major A of dQ; K_j is the K-major A of S^T and the MN-major B of dQ.  Nothing is transposed in memory.
//
// Warp roles (512 threads): warps 0-7 compute (thread = one key row x 64 query columns), warps 8-11 drain
// dQ (TMEM -> swizzled smem -> cp.reduce.async.bulk), warp 12 scheduler + TMA producer, warp 13 MMA issuer.
// TMEM (512 columns at d = 128): R0 | R1 | dV | dK.  The two 128-column regions R0 / R1 are time-shared by FOUR tenants so
// that no product ever waits for the drain of another (round 1 kept dQ in the dP columns, which put the chain
// dS(i) -> dQ(i) -> drain -> dP(i+1) -> dS(i+1) on the critical path, ~4100 cycles per tile pair):
//     R0:  S(i)  -> [compute warps pull it into registers at the start of the P phase] -> dP(i) -> [pulled into registers
//          at the start of the dS phase] -> S(i+1)
//     R1:  dQ(i-1) -> [drain warps pull it out] -> P(i) (bf16, stored at the END of the P phase) -> [dV(i) reads it] -> dQ(i)
// MMA issue order per step  dK(i-1) dQ(i-1) dP(i) dV(i) S(i+1)  (S(i+1) may be S(0) of the next work item):  every wait of
// the issuing warp is for an event that lies early in the compute warps' phase structure, so the tensor pipe runs ahead.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "fa_internal.h"
#include "fa_ptx.cuh"
#include "fa_tma_host.h"

// A/B switches for same-box experiments (tools/build_variant.sh + tools/ab.sh).
#ifndef FA_BWD_PREFETCH_KV
// L2 prefetch of the next item's K / V tiles (cp.async.bulk.prefetch.tensor an item ahead).  Worth 1 % when it was added (the K / V
// loads at an item boundary are HBM-cold); re-measured after the codegen fixes of the last session it COSTS 1 % everywhere
// (config 3: 5.53-5.56 against 5.57-5.62 ms, 5.953 against 6.024 ms power-capped; 64 heads 0.689 against 0.697 ms; config 4
// 23.78 against 23.87-24.1 ms): the extra TMA operations and L2 traffic are energy, and the step is power-capped.  Off.
#define FA_BWD_PREFETCH_KV 0
#endif
// k-step loops of the MMA issuer.  Round-2 history: while the issue code ran on vector registers (see the REDUX of the TMEM
// base in the issuer) the unroll factor was worth 10 % (1 = rolled: 7.1 ms = issue-bound, 2: 6.7, 4: 6.45-6.65, 8: 6.7-6.85:
// instruction cache) and the hash-heavy dropout variant wanted rolled loops (30.4 ms against 36.5 with 4).  On the uniform
// datapath an MMA costs ~3 issue slots and the factor is worth < 1 %: 1: 5.85-5.89 ms, 2: 5.81-5.87, 4: 5.85-5.86, 8: 5.79-5.82
// (config 4, dropout: 23.7-24.1 ms for every factor).
#ifndef FA_BWD_MMA_UNROLL_DROP
#define FA_BWD_MMA_UNROLL_DROP 8
#endif
#ifndef FA_BWD_MMA_UNROLL
#define FA_BWD_MMA_UNROLL 8
#endif

namespace fa {

namespace {

constexpr int kTile = 128;
#ifndef FA_BWD_HEAD_GROUP
#define FA_BWD_HEAD_GROUP 16  // same-box A/B at 64 | 128 | 512 heads (T = 4096): 8: 0.758 | 1.49 | 6.61 ms, 16: 0.738 | 1.475 | 6.60, 32: 0.750 | 1.505 | 6.70-6.95 (a group no longer fits the L2)
#endif
constexpr int kHeadGroup = FA_BWD_HEAD_GROUP;  // heads whose work items are interleaved longest-first (see `decode` in the kernel)
long long* g_bwd_trace = nullptr;  // developer timeline buffer (tc_bwd_set_trace)
int g_bwd_trace_t0 = 0;            // first tile pair of CTA 0 that the timeline records
int g_bwd_persistent = -1;         // -1: FA_B200_BWD_PERSISTENT (read once), default on

bool bwd_persistent() {
  if (g_bwd_persistent < 0) {
    const char* e = getenv("FA_B200_BWD_PERSISTENT");
    g_bwd_persistent = (e && e[0] == '0') ? 0 : 1;
  }
  return g_bwd_persistent != 0;
}
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
constexpr int kBwdThreads = 512;
constexpr int kBoxBytes = 128 * 128;  // 128 rows x 128 B

struct BwdParams {
  int T, Tpad, causal;
  int BH;
  int head_group;   // heads whose work items are interleaved longest-first (see `encode` in the kernel)
  int nkv;          // K/V tiles per head (= Tpad / 128)
  int total_items;  // BH * nkv work items (one item = one K/V tile of one head)
  int wide_stores;  // dk / dv are 32-byte aligned: 256-bit stores in the epilogue
  int persistent;   // 1: the CTAs draw items from *sched_counter until it runs dry; 0: CTA b works on item b only
  int* sched_counter;  // zeroed by the prep pass
  float tau, scale_log2;
  const float* lse_pad;    // [BH][Tpad]  -lse * log2(e); -inf on padding rows (=> P = 0 there)
  const float* delta_pad;  // [BH][Tpad]  -tau * rowsum(dO o O) (one packed FFMA2 gives tau (dP - delta)); 0 on padding rows
  void* dk;
  void* dv;
  float* dq_acc;  // fp32 dQ accumulator [BH][T][D] (workspace, or the caller's fp32 q_grad)
  Dropout drop;  // used by the kDrop instantiations only
  int dq_debug;      // timing experiments only (FA_B200_BWD_DEBUG_DQ): 1 = skip the reduce-adds, 2 = plain TMA stores
  int trace_t0;      // developer timeline: first tile pair (running count of CTA 0) that is recorded
  long long* trace;  // developer timeline (tools/fa_check trace): clock64 stamps of CTA 0, 16 slots per tile pair
};

template <int D>
struct BwdCfg {
  static constexpr int kBoxes = D / 64;
  static constexpr int kTileBytes = kBoxes * kBoxBytes;
  // setmaxnreg split of the 128 registers per thread the CTA is launched with (x 32 lanes: 8 compute, 4 drain, 4 other
  // warps).  The drain warps take a whole dQ tile (D registers) out of TMEM at once.  The dropout variant runs on the same
  // split since its dS phase re-reads the dropped P from TMEM instead of carrying mask words (it used 160 / 104 / 88 and
  // a two-part drain before: config 4 backward 28.1 -> 27.0 ms).
  static constexpr int kRegCompute = 136, kRegDrain = 152, kRegMisc = 88;       // 8*136 + 4*152 + 4*88 = 16*128
  static constexpr int kQStages = 2;
  // One dO stage at d = 128 (all that fits): the load of dO(t+1) can only start when dV(t) has finished, and under the
  // reduce-add stream a tile load takes ~2000-2400 cycles WHATEVER its size (measured with 16 KB half tiles in a ring of
  // two: same latency, and the N = 64 dP products cost more: 6.9 vs 6.5 ms; a ring of three halves needs 16 KB from the dQ
  // staging, and 8 KB staging slabs reduce-add at 16 instead of 22 B/clk: 7.8 ms).  load + dP + dV is the critical cycle.
  static constexpr int kDoStages = D >= 128 ? 1 : 2;
  static constexpr int kStageBytes = kBoxBytes;  // dQ staging slab: 128 rows x 32 fp32 (128-B swizzle)
  // shared-memory map (bytes).  The two K/V buffers swap roles from one work item to the next: item m keeps K in buffer
  // m & 1 and V in the other one, so that K of item m + 1 lands in the buffer V(m) leaves EARLY (after the last dP) and
  // the first S of the next item does not wait for the last dQ of this one.
  static constexpr int kOffKV = 0;
  static constexpr int kOffQ = kOffKV + 2 * kTileBytes;
  static constexpr int kOffDO = kOffQ + kQStages * kTileBytes;
  static constexpr int kOffDS = kOffDO + kDoStages * kTileBytes;  // dS^T: 128 x 128 bf16 = two boxes
  static constexpr int kOffStage = kOffDS + 2 * kBoxBytes;        // dQ staging: 2 slabs
  static constexpr int kOffStats = kOffStage + 2 * kStageBytes;   // [2 stages][L, delta][128] fp32
  static constexpr int kOffRowKeys = kOffStats + 2 * 2 * 128 * 4;  // dropout: [2 stages][64 query pairs] row keys of the mask hash
  static constexpr int kOffBars = kOffRowKeys + 2 * 64 * 4;
  static constexpr int kSmemBytes = kOffBars + 320;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
  // TMEM columns: R0 = S(i) then dP(i); R1 = dQ(i-1) then P(i)
  static constexpr uint32_t kColR0 = 0, kColR1 = 128, kColDV = 256, kColDK = 256 + D;
};

struct BwdBarriers {
  uint64_t k_full, v_full, k_free, v_free;                // per work item
  uint64_t q_full[2], q_empty[2], do_full[2], do_empty[2];  // per tile pair (rings)
  uint64_t s_full, s_loaded, p_full, dp_full, dp_loaded, ds_full, ds_empty, dq_full, dq_drained;  // per tile pair
  uint64_t dv_full, dk_full;                               // per work item
  uint64_t sched_full[2], sched_empty[2];                  // work-item ring
  int sched_item[2];  // (head << 16) | K/V tile, -1 = no more work
  uint32_t tmem_slot;
};
static_assert(sizeof(BwdBarriers) <= 320, "barrier block");

// 32 consecutive fp32 accumulator columns of one row -> global memory.  A thread owns a row, so a warp's store instruction
// touches 32 different lines: the epilogue is bound by tag lookups per instruction, and the 256-bit stores of sm_100
// (STG.256) halve their number; `wide` = the caller's pointers are 32-byte aligned.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <typename OutT>
__device__ __forceinline__ void store_chunk32(OutT* dst, const uint32_t* r, bool wide);
template <>
__device__ __forceinline__ void store_chunk32<float>(float* dst, const uint32_t* r, bool wide) {
  if (wide) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) st_global_v8(dst + i, r + i);
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      *reinterpret_cast<uint4*>(dst + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
  }
}
template <>
__device__ __forceinline__ void store_chunk32<__nv_bfloat16>(__nv_bfloat16* dst, const uint32_t* r, bool wide) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  if (wide) {
    st_global_v8(dst, v);
    st_global_v8(dst + 16, v + 8);
  } else {
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(dst + 2 * i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}

// The kernel is PERSISTENT: one CTA per SM draws work items (one K/V tile j of one head, heaviest first inside a head)
// from a global counter until it runs dry.  Every pipeline below is written against a RUNNING tile-pair counter `t`
// (mbarrier parities, ring stages) plus an item counter `m`, so the pipelines do not drain between items: while the
// tensor pipe finishes dQ / dK of the last tile pair of item m, K / Q of item m + 1 have already landed and S(0) of
// item m + 1 is issued ahead of that tail.  One-CTA-per-item launches paid ~10 k cycles of prologue + tail per item
// (15 % of the kernel at T = 4096: the short items near the diagonal are almost all overhead).
template <int D, typename OutT, bool kDrop>
__global__ void __launch_bounds__(kBwdThreads, 1)
fa_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
              const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
              const __grid_constant__ CUtensorMap map_dq, const BwdParams p) {
  using Cfg = BwdCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need 1024-B alignment
  uint8_t* sKV = smem + Cfg::kOffKV;
  uint8_t* sQ = smem + Cfg::kOffQ;
  uint8_t* sDO = smem + Cfg::kOffDO;
  uint8_t* sDS = smem + Cfg::kOffDS;
  uint8_t* sStage = smem + Cfg::kOffStage;
  float* sStats = reinterpret_cast<float*>(smem + Cfg::kOffStats);
  uint32_t* sRowKeys = reinterpret_cast<uint32_t*>(smem + Cfg::kOffRowKeys);
  BwdBarriers& bars = *reinterpret_cast<BwdBarriers*>(smem + Cfg::kOffBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T;
  const int nq = p.nkv;  // Q tiles per head == K/V tiles per head
#ifdef FA_B200_TRACE_HOOKS
  long long* const tr = blockIdx.x == 0 ? p.trace : nullptr;
  const int tr0 = p.trace_t0;
#endif
// developer timeline (tools/fa_check trace): compiled in only with -DFA_B200_TRACE_HOOKS (tools/build_variant.sh trace
// -DFA_B200_TRACE_HOOKS).  Each hook is ~8 instructions of predicate arithmetic even when the buffer is null -- about 100
// of the ~600 instructions a compute thread executes per tile pair, and in the power-capped steady state instructions are time.
#ifdef FA_B200_TRACE_HOOKS
#define FA_TR(t_, e_) do { if (tr && (t_) >= tr0 && (t_) < tr0 + 16 && lane == 0) tr[((t_) - tr0) * 16 + (e_)] = clock64(); } while (0)
#define FA_TR3(t_, e_) do { if (tr && (t_) >= tr0 && (t_) < tr0 + 16 && lane == 0) tr[256 + ((t_) - tr0) * 16 + (e_)] = clock64(); } while (0)
#define FA_TR2(t_, e_) do { if (tr && (t_) >= tr0 && (t_) < tr0 + 16 && lane == 0 && warp == 0) tr[256 + ((t_) - tr0) * 16 + (e_)] = clock64(); } while (0)
#else
#define FA_TR(t_, e_) do { } while (0)
#define FA_TR3(t_, e_) do { } while (0)
#define FA_TR2(t_, e_) do { } while (0)
#endif

  if (threadIdx.x == 0) {
    mbar_init(&bars.k_full, 1);
    mbar_init(&bars.v_full, 1);
    mbar_init(&bars.k_free, 1);
    mbar_init(&bars.v_free, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.q_full[s], 1);
      mbar_init(&bars.q_empty[s], 1);
      mbar_init(&bars.do_full[s], 1);
      mbar_init(&bars.do_empty[s], 1);
      mbar_init(&bars.sched_full[s], 1);
      mbar_init(&bars.sched_empty[s], 13);  // 8 compute + 4 drain + the MMA warp
    }
    mbar_init(&bars.s_full, 1);
    mbar_init(&bars.s_loaded, 8);
    mbar_init(&bars.p_full, 8);
    mbar_init(&bars.dp_full, 1);
    mbar_init(&bars.dp_loaded, 8);
    mbar_init(&bars.ds_full, 8);
    mbar_init(&bars.ds_empty, 1);
    mbar_init(&bars.dq_full, 1);
    mbar_init(&bars.dq_drained, 4);
    mbar_init(&bars.dv_full, 1);
    mbar_init(&bars.dk_full, 1);
    fence_mbar_init();
  }
  if (warp == 13) {
    tmem_alloc(&bars.tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_slot;

  // a consumer role picks up work item number m from the ring (all lanes read the same word)
  auto next_item = [&](int m) {
    mbar_wait(&bars.sched_full[m & 1], (m >> 1) & 1);
    const int item = *reinterpret_cast<volatile int*>(&bars.sched_item[m & 1]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars.sched_empty[m & 1]);
    return item;
  };

  // Item order: heads in groups of p.head_group (<= kHeadGroup, sized to the L2 by the launcher); inside a group all K/V tiles j = 0 (the longest items when causal) of the
  // group's heads come first, then all j = 1, ... -- the counter hands the longest items out first ACROSS heads, so the
  // last items a CTA can draw are the shortest ones (simulated schedule efficiency at 64 heads, the strong-scaling share
  // of one of 8 GPUs: 0.93 head by head, 0.99 in groups), while a group's tensors (16 heads x 8 MB) stay inside the L2.
  // The scheduler lane decodes an item number once and publishes (head << 16) | j (-1 = no more work) through the ring.
  auto encode = [&](int item) {
    if (item >= p.total_items) return -1;
    const int per_group = p.head_group * nq;
    const int grp = item / per_group, r = item - grp * per_group;
    const int gs = min(p.head_group, p.BH - grp * p.head_group);
    const int j = r / gs;
    return ((grp * p.head_group + (r - j * gs)) << 16) | j;
  };
  auto decode = [&](int code, int& bh, int& j) {
    bh = code >> 16;
    j = code & 0xffff;
  };

  if (warp >= 12) {
    reg_dealloc<Cfg::kRegMisc>();
  }
  if (warp == 12) {
    // ------------------------------------------------------------------ scheduler + TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      tma_prefetch_desc(&map_do);
      int item = p.persistent ? atomicAdd(p.sched_counter, 1) : (int)blockIdx.x;
      uint32_t t = 0;
      for (int m = 0;; ++m) {
        mbar_wait(&bars.sched_empty[m & 1], ((m >> 1) & 1) ^ 1);  // every consumer has read item m - 2
        const int code = encode(item);
        bars.sched_item[m & 1] = code;
        mbar_arrive(&bars.sched_full[m & 1]);
        if (code < 0) break;
        const int next = p.persistent ? atomicAdd(p.sched_counter, 1) : p.total_items;  // its latency hides under the loads
#if FA_BWD_PREFETCH_KV
        if (next < p.total_items) {
          // K / V of the next item come from HBM (no other CTA reads them): pull them into the L2 an item ahead, so that the
          // loads at the item boundary -- which sit on the critical path -- are L2 hits
          int nbh, nj;
          decode(encode(next), nbh, nj);
          for (int b = 0; b < Cfg::kBoxes; ++b) {
            tma_prefetch_l2_3d(&map_k, b * 64, nj * kTile, nbh);
            tma_prefetch_l2_3d(&map_v, b * 64, nj * kTile, nbh);
          }
        }
#endif
        int bh, j;
        decode(code, bh, j);
        const int i0 = p.causal ? j : 0, n = nq - i0;
        uint8_t* bufK = sKV + (m & 1) * Cfg::kTileBytes;
        uint8_t* bufV = sKV + ((m & 1) ^ 1) * Cfg::kTileBytes;
        // K first: its buffer held V of the previous item, which the last dP of that item released
        if (m > 0) mbar_wait(&bars.v_free, (m - 1) & 1);
        FA_TR3(t, 8);
        mbar_arrive_expect_tx(&bars.k_full, Cfg::kTileBytes);
        for (int b = 0; b < Cfg::kBoxes; ++b)
          tma_load_3d(bufK + b * kBoxBytes, &map_k, &bars.k_full, b * 64, j * kTile, bh, kEvictFirst);
        for (int il = 0; il < n; ++il, ++t) {
          const int i = i0 + il;
          const int s = t & 1;
          mbar_wait(&bars.q_empty[s], ((t >> 1) & 1) ^ 1);
          FA_TR3(t, 9);
          mbar_arrive_expect_tx(&bars.q_full[s], Cfg::kTileBytes + 1024);
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_3d(sQ + s * Cfg::kTileBytes + b * kBoxBytes, &map_q, &bars.q_full[s], b * 64, i * kTile, bh,
                        kEvictLast);
          const size_t soff = (size_t)bh * p.Tpad + (size_t)i * kTile;
          bulk_load_1d(sStats + s * 256, p.lse_pad + soff, 512, &bars.q_full[s]);
          bulk_load_1d(sStats + s * 256 + 128, p.delta_pad + soff, 512, &bars.q_full[s]);
          const int sd = t % Cfg::kDoStages;
          mbar_wait(&bars.do_empty[sd], ((t / Cfg::kDoStages) & 1) ^ 1);
          FA_TR3(t, 10);
          mbar_arrive_expect_tx(&bars.do_full[sd], Cfg::kTileBytes);
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_3d(sDO + sd * Cfg::kTileBytes + b * kBoxBytes, &map_do, &bars.do_full[sd], b * 64, i * kTile, bh,
                        kEvictLast);
          if (il == 0) {
            // V (first needed by dP(0), after S(0) and the start of the P phase) goes where K of the previous item was,
            // which the last dQ of that item releases
            if (m > 0) mbar_wait(&bars.k_free, (m - 1) & 1);
            FA_TR3(t, 11);
            mbar_arrive_expect_tx(&bars.v_full, Cfg::kTileBytes);
            for (int b = 0; b < Cfg::kBoxes; ++b)
              tma_load_3d(bufV + b * kBoxBytes, &map_v, &bars.v_full, b * 64, j * kTile, bh, kEvictFirst);
          }
        }
        item = next;
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp runs the issue code
    // REDUX inside this branch: a shared-memory load is not provably warp-uniform for ptxas, and with ONE non-uniform operand
    // (every MMA takes tmem + const) it builds ALL operands of the issue loops in vector registers and wraps each UTCHMMA
    // into ELECT + four R2UR.BROADCAST (~16 instructions per MMA instead of ~3 on the uniform datapath)
    const uint32_t tmem = __reduce_max_sync(0xffffffffu, bars.tmem_slot);
    // (uniform-register descriptors, see fa_ptx.cuh `_w` forms); one elected lane issues each MMA / commit.
    // ONE flat loop over the tile pairs of all items: an item boundary only changes descriptors and which commits fire,
    // so the boundary runs the same (instruction-cache-hot) code as the steady state -- a separate head / tail path was
    // measured at ~4400 cycles for its 24 MMAs (cold code: the SM's instruction cache is 32 KB).  The k-step loops are fully
    // unrolled: on the uniform datapath a k-step is ~3 instructions (see FA_BWD_MMA_UNROLL above).
    constexpr int kMmaUnroll = kDrop ? FA_BWD_MMA_UNROLL_DROP : FA_BWD_MMA_UNROLL;
    constexpr uint32_t idesc_sT = make_idesc(kFmtBF16, 128, 128, 0, 0);  // K-major x K-major
    constexpr uint32_t idesc_acc = make_idesc(kFmtBF16, 128, D, 0, 1);   // (TMEM | K-major) x MN-major
    constexpr uint32_t idesc_dq = make_idesc(kFmtBF16, 128, D, 1, 1);    // MN-major x MN-major
    const uint32_t aKV = smem_u32(sKV), aQ = smem_u32(sQ), aDO = smem_u32(sDO), aDS = smem_u32(sDS);
    // One base descriptor per (tile, major-ness); a k-step only adds to its low word.
    // K-major view of a [128 x D] tile: k-step kk of 16 elements = 32 B inside the 128-B swizzle row, boxes of 64 elements.
    // MN-major view: 16 "K" rows (2 KB) per step, MN atoms (64 elements) one box apart.
    auto kmaj0 = [](uint32_t base) { return make_smem_desc(base, 16, 1024); };
    auto mnmaj0 = [](uint32_t base) { return make_smem_desc(base, kBoxBytes, 1024); };
    auto koff = [](int kk) { return uint32_t((kk >> 2) * kBoxBytes + (kk & 3) * 32); };
    const uint64_t dDS_k = kmaj0(aDS), dDS_mn = mnmaj0(aDS);
    auto issue_sT = [&](uint64_t a0, uint64_t b0) {  // S^T or dP^T -> R0: K-major x K-major
#pragma unroll kMmaUnroll
      for (int kk = 0; kk < D / 16; ++kk)
        umma_ss_w(tmem + Cfg::kColR0, desc_advance(a0, koff(kk)), desc_advance(b0, koff(kk)), idesc_sT, kk > 0);
    };
    auto issue_dv = [&](uint64_t do_mn, bool acc) {  // A = P^T (bf16, two per column) in R1
#pragma unroll kMmaUnroll
      for (int kk = 0; kk < 8; ++kk)
        umma_ts_w(tmem + Cfg::kColDV, tmem + Cfg::kColR1 + (kk >> 2) * 64 + (kk & 3) * 8, desc_advance(do_mn, kk * 2048),
                  idesc_acc, (acc || kk > 0) ? 1u : 0u);
    };
    auto issue_dk = [&](uint64_t q_mn, bool acc) {
#pragma unroll kMmaUnroll
      for (int kk = 0; kk < 8; ++kk)
        umma_ss_w(tmem + Cfg::kColDK, desc_advance(dDS_k, koff(kk)), desc_advance(q_mn, kk * 2048), idesc_acc,
                  (acc || kk > 0) ? 1u : 0u);
    };
    auto issue_dq = [&](uint64_t k_mn) {
#pragma unroll kMmaUnroll
      for (int kk = 0; kk < 8; ++kk)
        umma_ss_w(tmem + Cfg::kColR1, desc_advance(dDS_mn, kk * 2048), desc_advance(k_mn, kk * 2048), idesc_dq, kk > 0);
    };
    auto q_tile = [&](uint32_t t) { return aQ + (t & 1) * Cfg::kTileBytes; };
    auto do_tile = [&](uint32_t t) { return aDO + (t % Cfg::kDoStages) * Cfg::kTileBytes; };
    // the ring word is read by all lanes; the max-reduction makes the value (and everything derived from it: loop
    // bounds, descriptors) provably warp-uniform for ptxas
    auto fetch = [&](int m) { return __reduce_max_sync(0xffffffffu, next_item(m)); };

    int m = 0;
    int item = fetch(0);
    if (item >= 0) {
      int n = nq - (p.causal ? (item & 0xffff) : 0);  // tile pairs of the current item
      int il = 0;                               // position of pair t inside its item
      uint32_t aK = aKV, aV = aKV + Cfg::kTileBytes;
      uint32_t prev_aK = 0;
      bool prev_first = false, prev_last = false, done = false;
      mbar_wait_w(&bars.k_full, 0);
      mbar_wait_w(&bars.q_full[0], 0);
      tc_fence_after();
      issue_sT(kmaj0(aK), kmaj0(q_tile(0)));
      umma_commit_w(&bars.s_full);
      for (uint32_t t = 0;; ++t) {
        if (t > 0) {
          // back half of pair t-1 (possibly the last pair of the previous item)
          mbar_wait_w(&bars.ds_full, (t - 1) & 1);
          tc_fence_after();
          FA_TR(t, 8);
          // dK(t-1) first: it releases the Q stage that Q(t+1) -- needed by S(t+1) below -- loads into; then dQ(t-1) -> R1
          // (P(t-1) was consumed by dV(t-1), issued before: the tensor pipe executes in order)
          issue_dk(mnmaj0(q_tile(t - 1)), !prev_first);
          umma_commit_w(&bars.q_empty[(t - 1) & 1]);
          if (prev_last) umma_commit_w(&bars.dk_full);
          FA_TR(t, 9);
          issue_dq(mnmaj0(prev_aK));
          umma_commit_w(&bars.dq_full);
          umma_commit_w(&bars.ds_empty);  // both readers of dS(t-1) are done
          if (prev_last) {
            umma_commit_w(&bars.k_free);  // last reader of that item's K
            FA_TR3(t, 15);
          }
        }
        if (done) break;
        const bool first = il == 0, last = il == n - 1;
        // dP(t) -> R0 (needs S(t) in the compute warps' registers)
        mbar_wait_w(&bars.s_loaded, t & 1);
        mbar_wait_w(&bars.do_full[t % Cfg::kDoStages], (t / Cfg::kDoStages) & 1);
        if (first) mbar_wait_w(&bars.v_full, m & 1);
        tc_fence_after();
        FA_TR(t, 10);
        issue_sT(kmaj0(aV), kmaj0(do_tile(t)));
        umma_commit_w(&bars.dp_full);
        if (last) umma_commit_w(&bars.v_free);  // last reader of V
        FA_TR(t, 11);
        // dV(t) <- P(t) in R1 (the compute warps stored it after the drain of dQ(t-1))
        mbar_wait_w(&bars.p_full, t & 1);
        tc_fence_after();
        FA_TR(t, 12);
        issue_dv(mnmaj0(do_tile(t)), !first);
        umma_commit_w(&bars.do_empty[t % Cfg::kDoStages]);
        if (last) umma_commit_w(&bars.dv_full);
        FA_TR(t, 13);
        prev_aK = aK;
        prev_first = first;
        prev_last = last;
        ++il;
        if (last) {
          ++m;
          item = fetch(m);
          if (item < 0) {
            done = true;
          } else {
            n = nq - (p.causal ? (item & 0xffff) : 0);
            il = 0;
            aK = aKV + (m & 1) * Cfg::kTileBytes;
            aV = aKV + ((m & 1) ^ 1) * Cfg::kTileBytes;
            FA_TR3(t + 1, 12);
            mbar_wait_w(&bars.k_full, m & 1);
            FA_TR3(t + 1, 13);
          }
        }
        if (!done) {
          // S(t+1) -> R0, once the compute warps hold dP(t) in registers.  At the end of an item that is S(0) of the NEXT
          // item: its K sits in the other K/V buffer, so it goes ahead of the dK / dQ of this item's last pair
          mbar_wait_w(&bars.dp_loaded, t & 1);
          mbar_wait_w(&bars.q_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
          tc_fence_after();
          issue_sT(kmaj0(aK), kmaj0(q_tile(t + 1)));
          umma_commit_w(&bars.s_full);
          FA_TR(t, 14);
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ compute warps
    reg_alloc<Cfg::kRegCompute>();
    const int g = warp >> 2;                 // which half of the 128 query columns
    const int r = (warp & 3) * 32 + lane;    // key row within the tile == TMEM lane
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tR0 = tmem + Cfg::kColR0 + lane_base + g * 64;  // S^T(i), then dP^T(i): this thread's 64 query columns
    const uint32_t tR1 = tmem + Cfg::kColR1 + lane_base + g * 64;  // P^T(i) (bf16, 32 columns) after dQ(i-1) has been drained
    const float2 c2v = make_float2(p.scale_log2, p.scale_log2);
    uint8_t* ds_row = sDS + g * kBoxBytes + r * 128;
    const uint32_t drop_thr = p.drop.thr;
    const float tau_keep = p.tau * p.drop.keep_scale;  // tau * Z for a kept element (dropout variant)
    uint32_t t = 0;
    for (int m = 0;; ++m) {
      const int item = next_item(m);
      if (item < 0) break;
      int bh, j;
      decode(item, bh, j);
      const int i0 = p.causal ? j : 0, n = nq - i0;
      // dropout: one 32-bit hash per 2 x 2 block of (query, key); this thread (one key row) uses the byte of its key
      // parity in the low half (even query) and in the high half (odd query) -- the same bits the forward drew
      const uint32_t kj = uint32_t(j * kTile + r);
      const uint32_t drop_head = drop_head_key(p.drop.seed_lo, p.drop.seed_hi, bh), drop_colterm = (kj >> 1) * kDropColMul;
      const uint32_t drop_b0 = kj & 1;  // byte b0: even query of the block, byte b0 + 2: odd query
      // mask selectors over the word pair (mine, theirs): queries c, c+1 take the even lane's word, c+2, c+3 the odd lane's
      const uint32_t drop_o0 = (lane & 1) * 4, drop_o1 = 4 - drop_o0;
      const uint32_t drop_sel2a = drop_sel_pair(drop_b0 + drop_o0, drop_b0 + 2 + drop_o0),
                     drop_sel2b = drop_sel_pair(drop_b0 + drop_o1, drop_b0 + 2 + drop_o1);
      for (int il = 0; il < n; ++il, ++t) {
        const int i = i0 + il;
        const float* sL = sStats + (t & 1) * 256 + g * 64;  // -L log2(e) of the 64 query columns (-inf on padding rows)
        const float* sD = sL + 128;                         // -tau * delta
        if constexpr (kDrop) {
          // The first round of the mask hash (row key of a query pair) does not depend on the key row: the 64 keys of this
          // Q tile are hashed once per CTA -- one each by the first 64 compute threads -- instead of once per thread and
          // word (half of the hash instructions of this loop).  Two buffers: after the barrier a warp is at most one tile
          // pair ahead of the slowest one.
          if (threadIdx.x < 64) sRowKeys[(t & 1) * 64 + threadIdx.x] = drop_row_key(drop_head, (uint32_t(i * kTile) >> 1) + threadIdx.x);
          named_bar_sync(2, 256);
        }
        // ---- P^T = exp2(c * S^T - L)
        mbar_wait(&bars.q_full[t & 1], (t >> 1) & 1);  // L / delta of this Q tile have landed
        mbar_wait(&bars.s_full, t & 1);
        tc_fence_after();
        if (warp == 0) FA_TR(t, 0);
        // P is kept as packed bf16 pairs (32 registers instead of 64 fp32): exactly the values the dV product sees, and
        // what leaves room to pull a whole dP row chunk (64 registers) out of TMEM in one go below
        uint32_t pk[32];
        {
          float pr[64];
          uint32_t* u = reinterpret_cast<uint32_t*>(pr);
          tmem_ld_x32(tR0, u);
          tmem_ld_x32(tR0 + 32, u + 32);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.s_loaded);  // R0 may take dP(i)
          FA_TR2(t, 0);
          const bool diag = p.causal && (i == j);
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const float4 L4 = *reinterpret_cast<const float4*>(sL + c);
            const float2 x0 = ffma2(make_float2(pr[c], pr[c + 1]), c2v, make_float2(L4.x, L4.y));
            const float2 x1 = ffma2(make_float2(pr[c + 2], pr[c + 3]), c2v, make_float2(L4.z, L4.w));
            pr[c + 0] = ex2_approx(x0.x);
            pr[c + 1] = ex2_approx(x0.y);
            pr[c + 2] = ex2_approx(x1.x);
            pr[c + 3] = ex2_approx(x1.y);
          }
          if (diag) {
            // a real branch: without the (empty) volatile asm the block is if-converted into 64 ISETP + 64 FSEL that
            // EVERY tile pair executes (~130 of ~800 instructions of the loop) although one pair in 16 needs them
            asm volatile("" ::: "memory");
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (g * 64 + c < r) pr[c] = 0.f;  // key index > query index
          }
          FA_TR2(t, 1);
#pragma unroll
          for (int c = 0; c < 64; c += 2) pk[c >> 1] = pack_bf16x2(pr[c], pr[c + 1]);
        }
        // R1 still holds dQ(t-1) until the drain warps have pulled it out
        if (t > 0) mbar_wait(&bars.dq_drained, (t - 1) & 1);
        FA_TR2(t, 2);
        if constexpr (kDrop) {  // P^T o keep feeds dV; pk[] itself stays undropped for dS
          uint32_t pd[32];
          const uint32_t* rk = sRowKeys + (t & 1) * 64 + g * 32 + (lane & 1);
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            // the two lanes of a key pair need the same block words: each hashes one, a shuffle gives the other
            const uint32_t mine = drop_keep_word(drop_mix(rk[c >> 1] ^ drop_colterm), drop_thr);
            const uint32_t theirs = __shfl_xor_sync(0xffffffffu, mine, 1);
            pd[c >> 1] = pk[c >> 1] & drop_prmt2(mine, theirs, drop_sel2a);
            pd[(c >> 1) + 1] = pk[(c >> 1) + 1] & drop_prmt2(mine, theirs, drop_sel2b);
          }
          tmem_st_x32(tR1, pd);
        } else {
          // P^T of query columns [64g, 64g+64) goes to R1 columns [64g, 64g+32)
          tmem_st_x32(tR1, pk);
        }
        tmem_wait_st();
        FA_TR2(t, 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full);
        if (warp == 0) FA_TR(t, 1);
        // ---- dS^T = P^T o (tau dP^T - tau delta)      (dropout: tau Z dP^T - tau delta)
        mbar_wait(&bars.dp_full, t & 1);
        tc_fence_after();
        if (warp == 0) FA_TR(t, 2);
        float dp[64];
        tmem_ld_x32(tR0, reinterpret_cast<uint32_t*>(dp));
        tmem_ld_x32(tR0 + 32, reinterpret_cast<uint32_t*>(dp) + 32);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.dp_loaded);  // R0 may take S(t+1)
        FA_TR2(t, 5);
        if (t > 0) mbar_wait(&bars.ds_empty, (t - 1) & 1);  // dK(t-1) / dQ(t-1) have finished reading dS
        FA_TR2(t, 6);
        const float2 tau2 = make_float2(p.tau, p.tau);
        const float2 tk2 = make_float2(tau_keep, tau_keep);
        auto unpack = [](uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };
        uint32_t pdr[16];  // dropout: 16 packed columns (32 queries) of the dropped P, re-read from R1
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {  // 16-byte chunk of this thread's 128-byte dS^T row: 8 query columns
          uint32_t out[4];
          // dropout: dS = tau Z (P o keep) o dP - P o tau delta.  The dropped P is what this thread stored into R1 for the dV
          // product -- it stays there until dQ(t) is issued, i.e. past this phase -- so it is read back (16 packed columns
          // at a time) instead of carrying 64 keep bits per thread through the P phase and turning them into per-element
          // selects here (-320 of ~1100 instructions per thread and tile pair)
          if constexpr (kDrop) {
            if ((ch & 3) == 0) {
              tmem_ld_x16(tR1 + (ch >> 2) * 16, pdr);
              tmem_wait_ld();
            }
          }
#pragma unroll
          for (int e = 0; e < 8; e += 4) {
            const float4 D4 = *reinterpret_cast<const float4*>(sD + ch * 8 + e);
            const int c = ch * 8 + e;
            const uint32_t w0 = pk[c >> 1], w1 = pk[(c >> 1) + 1];
            float2 a, b;
            if constexpr (kDrop) {
              const uint32_t d0 = pdr[(c >> 1) & 15], d1 = pdr[((c >> 1) + 1) & 15];
              const float2 u0 = fmul2(unpack(w0), make_float2(D4.x, D4.y)), u1 = fmul2(unpack(w1), make_float2(D4.z, D4.w));
              a = ffma2(fmul2(unpack(d0), make_float2(dp[c + 0], dp[c + 1])), tk2, u0);
              b = ffma2(fmul2(unpack(d1), make_float2(dp[c + 2], dp[c + 3])), tk2, u1);
            } else {
              const float2 t0 = ffma2(make_float2(dp[c + 0], dp[c + 1]), tau2, make_float2(D4.x, D4.y));
              const float2 t1 = ffma2(make_float2(dp[c + 2], dp[c + 3]), tau2, make_float2(D4.z, D4.w));
              a = fmul2(unpack(w0), t0);
              b = fmul2(unpack(w1), t1);
            }
            out[(e >> 1) + 0] = pack_bf16x2(a.x, a.y);
            out[(e >> 1) + 1] = pack_bf16x2(b.x, b.y);
          }
          *reinterpret_cast<uint4*>(ds_row + ((ch ^ (r & 7)) << 4)) = make_uint4(out[0], out[1], out[2], out[3]);
        }
        FA_TR2(t, 7);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.ds_full);
        if (warp == 0) FA_TR(t, 3);
        if (warp == 7) FA_TR(t, 4);
      }
      // ---- epilogue of the item: every warp stores its 32 key rows x D/2 columns of dV_j (complete since the last dV,
      // long ago), then of dK_j (the tail MMA of the item).  The MMA warp restarts the two accumulators (dV of the next
      // item after the next p_full, dK after the next ds_full), i.e. after this code.
      const int kv_row = j * kTile + r;
      const bool row_ok = kv_row < T;
      const size_t out_off = ((size_t)bh * T + (row_ok ? kv_row : 0)) * D + g * (D / 2);
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {  // rolled: this code runs once per item and is never instruction-cache-hot
        mbar_wait(which == 0 ? &bars.dv_full : &bars.dk_full, m & 1);
        tc_fence_after();
        OutT* out = reinterpret_cast<OutT*>(which == 0 ? p.dv : p.dk) + out_off;
        const uint32_t tAcc = tmem + (which == 0 ? Cfg::kColDV : Cfg::kColDK) + lane_base + g * (D / 2);
        const float mul = (kDrop && which == 0) ? p.drop.keep_scale : 1.f;  // dV accumulated (P o keep)^T dO: 1/(1-p) once here
#pragma unroll 1
        for (int c0 = 0; c0 < D / 2; c0 += 32) {
          uint32_t rr[32];
          tmem_ld_x32(tAcc + c0, rr);
          tmem_wait_ld();
          if constexpr (kDrop) {
#pragma unroll
            for (int x = 0; x < 32; ++x) rr[x] = __float_as_uint(__uint_as_float(rr[x]) * mul);
          }
          if (row_ok) store_chunk32<OutT>(out + c0, rr, p.wide_stores != 0);
        }
      }
      tc_fence_before();  // orders the TMEM reads above before the p_full / ds_full arrivals of the next item
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ dQ drain warps
    reg_alloc<Cfg::kRegDrain>();
    const int w = warp - 8;
    const int r = w * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t tDQ = tmem + Cfg::kColR1 + (uint32_t(w * 32) << 16);
    const bool leader = (warp == 8 && lane == 0);
    // staging addresses of this thread's eight 16-byte chunks (128-B swizzle): computed once -- per store they cost three ALU
    // instructions, 100 of the ~340 a drain thread executes per tile pair.  A pair fills an even number of slabs (D / 32), so the
    // slab a chunk goes to is a compile-time constant and folds into the store's immediate offset.
    static_assert((D / 32) % 2 == 0, "slab parity per pair");
    uint32_t st_addr[8];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) st_addr[ch] = smem_u32(sStage) + r * 128 + ((ch ^ (r & 7)) << 4);
    uint32_t t = 0;
    for (int m = 0;; ++m) {
      const int item = next_item(m);
      if (item < 0) break;
      int bh, j;
      decode(item, bh, j);
      const int i0 = p.causal ? j : 0, n = nq - i0;
      for (int il = 0; il < n; ++il, ++t) {
        const int i = i0 + il;
        mbar_wait(&bars.dq_full, t & 1);
        tc_fence_after();
        if (warp == 8) FA_TR(t, 5);
        // The whole dQ tile leaves TMEM at once (D registers per thread), so that P(t+1) -- which takes these columns --
        // is released after one TMEM read instead of after three staging rounds; the staging / reduce-adds below then
        // overlap the next tile pair.
        constexpr int kParts = 1, kColsPart = D / kParts;
#pragma unroll
        for (int part = 0; part < kParts; ++part) {
          uint32_t rr[kColsPart];
#pragma unroll
          for (int c = 0; c < kColsPart / 32; ++c) tmem_ld_x32(tDQ + part * kColsPart + c * 32, rr + c * 32);
          tmem_wait_ld();
          if (part == kParts - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.dq_drained);
            if (warp == 8) FA_TR(t, 6);
          }
#pragma unroll
          for (int c = 0; c < kColsPart / 32; ++c) {
            uint8_t* buf = sStage + (c & 1) * Cfg::kStageBytes;
            if (leader && (t > 0 || c >= 2)) tma_store_wait_read<1>();  // staging buffer no longer being read
            named_bar_sync(1, 128);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              st_shared_v4(st_addr[ch] + (c & 1) * Cfg::kStageBytes, rr[c * 32 + ch * 4], rr[c * 32 + ch * 4 + 1], rr[c * 32 + ch * 4 + 2],
                           rr[c * 32 + ch * 4 + 3]);
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (leader) {
#ifdef FA_B200_DEBUG_HOOKS  // timing experiments (wrong results): skip the reduce-adds / plain stores instead
              if (p.dq_debug == 0) tma_reduce_add_3d(&map_dq, buf, part * kColsPart + c * 32, i * kTile, bh);
              else if (p.dq_debug == 2) tma_store_3d(&map_dq, buf, part * kColsPart + c * 32, i * kTile, bh);
#else
              tma_reduce_add_3d(&map_dq, buf, part * kColsPart + c * 32, i * kTile, bh);
#endif
              tma_store_commit();
            }
          }
        }
        if (warp == 8) FA_TR(t, 7);
      }
    }
    if (leader) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, 512);
}

// delta = rowsum(dO o O), lse -> log2 domain, both padded to a multiple of 128 rows per head; the same pass
// zero-fills the fp32 dQ accumulator (one kernel instead of a 1 GB memset plus a second launch).
// D / 8 lanes own one row (16-byte loads), so a warp covers 256 / D rows... 32 / (D / 8) rows per pass.
template <int D>
__global__ void __launch_bounds__(256)
fa_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO,
                   const float* __restrict__ lse, float* __restrict__ lse_pad, float* __restrict__ delta_pad,
                   float* __restrict__ dq_zero, int T, int Tpad, size_t total_rows_pad, float tau,
                   int* __restrict__ sched_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *sched_counter = 0;  // work-item counter of the persistent main kernel
  constexpr int kLanesPerRow = D / 8;            // 16 (d = 128) or 8 (d = 64)
  constexpr int kRowsPerWarp = 32 / kLanesPerRow;
  const int lane = threadIdx.x & 31;
  const int sub = lane / kLanesPerRow, li = lane % kLanesPerRow;
  const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t row = warp_global * kRowsPerWarp + sub; row < total_rows_pad; row += nwarps * kRowsPerWarp) {
    const size_t bh = row / Tpad;
    const int t = (int)(row - bh * Tpad);
    float acc = 0.f;
    if (t < T) {
      const size_t base = (bh * T + t) * D + li * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + base);
      const uint4 b = *reinterpret_cast<const uint4*>(dO + base);
      const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* bh2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __bfloat1622float2(ah[i]), y = __bfloat1622float2(bh2[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
      if (dq_zero) {
        float4* z = reinterpret_cast<float4*>(dq_zero + base);
        z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int off = kLanesPerRow / 2; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (li == 0) {
      delta_pad[row] = t < T ? -tau * acc : 0.f;
      lse_pad[row] = t < T ? -lse[bh * T + t] * 1.4426950408889634f : -INFINITY;
    }
  }
}

// dq (bf16) = [dq +] acc
__global__ void fa_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, size_t n,
                                         int accumulate) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
    float4 a = reinterpret_cast<const float4*>(acc)[i];
    uint2* dst = reinterpret_cast<uint2*>(dq) + i;
    if (accumulate) {
      const uint2 old = *dst;
      const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&old);
      const float2 x = __bfloat1622float2(oh[0]), y = __bfloat1622float2(oh[1]);
      a.x += x.x; a.y += x.y; a.z += y.x; a.w += y.y;
    }
    *dst = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
  }
}

size_t pad_rows(int T) { return (size_t)((T + kTile - 1) / kTile) * kTile; }
size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <int D, typename OutT>
cudaError_t launch_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* o,
                       const __nv_bfloat16* dO, const float* lse, void* dq, void* dk, void* dv, bool accumulate_dq,
                       void* workspace, int BH, int T, bool causal, Dropout drop, cudaStream_t st) {
  using Cfg = BwdCfg<D>;
  constexpr bool kF32 = sizeof(OutT) == 4;
  const size_t Tpad = pad_rows(T);
  const size_t n = (size_t)BH * T * D;
  char* ws = static_cast<char*>(workspace);
  int* sched_counter = reinterpret_cast<int*>(ws);
  ws += 256;
  float* lse_pad = reinterpret_cast<float*>(ws);
  ws += align256((size_t)BH * Tpad * 4);
  float* delta_pad = reinterpret_cast<float*>(ws);
  ws += align256((size_t)BH * Tpad * 4);
  float* dq_acc = kF32 ? static_cast<float*>(dq) : reinterpret_cast<float*>(ws);
  cudaError_t e;
  {
    const size_t rows = (size_t)BH * Tpad;
    const bool zero = !kF32 || !accumulate_dq;  // fp32 gradients with accumulate: reduce straight into the caller's q_grad
    constexpr int kRowsPerBlock = 8 * (32 / (D / 8));
    size_t blocks = (rows + kRowsPerBlock - 1) / kRowsPerBlock;
    if (blocks > 148 * 32) blocks = 148 * 32;  // grid-stride beyond 32 blocks per SM
    fa_bwd_prep_kernel<D><<<(unsigned)blocks, 256, 0, st>>>(o, dO, lse, lse_pad, delta_pad, zero ? dq_acc : nullptr, T,
                                                          (int)Tpad, rows, sqrtf(1.0f / (float)D), sched_counter);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  CUtensorMap mq, mk, mv, mdo, mdq;
  if (make_tmap_3d(&mq, q, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mk, k, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mv, v, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mdo, dO, 2, D, T, BH, 64, kTile)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mdq, dq_acc, 4, D, T, BH, 32, kTile)) return cudaErrorInvalidValue;
  e = cudaFuncSetAttribute(fa_bwd_kernel<D, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(fa_bwd_kernel<D, OutT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (e != cudaSuccess) return e;
  BwdParams p;
  p.T = T;
  p.Tpad = (int)Tpad;
  p.causal = causal ? 1 : 0;
  p.tau = sqrtf(1.0f / (float)D);
  p.scale_log2 = p.tau * 1.4426950408889634f;
  p.lse_pad = lse_pad;
  p.delta_pad = delta_pad;
  p.dk = dk;
  p.dv = dv;
  p.wide_stores = ((reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 31) == 0 ? 1 : 0;
  p.dq_acc = dq_acc;
  p.drop = drop;
  p.trace = g_bwd_trace;
  p.trace_t0 = g_bwd_trace_t0;
  p.dq_debug = 0;
  p.BH = BH;
  {
    // a group's tensors (Q, K, V, dO, dK, dV in bf16 + the fp32 dQ accumulator: 16 bytes per element) should fit the L2
    // about twice over: 16 heads at T = 4096, d = 128 (same-box sweep above), 4 at T = 16384 (16: -1.7 % at config 4)
    const size_t per_head = (size_t)T * D * 16;
    p.head_group = (int)std::max<size_t>(1, std::min<size_t>(kHeadGroup, ((size_t)128 << 20) / per_head));
  }
  p.nkv = (int)(Tpad / kTile);
  p.total_items = BH * p.nkv;
  p.persistent = bwd_persistent() ? 1 : 0;
  p.sched_counter = sched_counter;
#ifdef FA_B200_DEBUG_HOOKS  // developer builds only (tools/build_variant.sh <name> -DFA_B200_DEBUG_HOOKS): results are wrong when set
  if (const char* e = getenv("FA_B200_BWD_DEBUG_DQ")) p.dq_debug = e[0] == 's' && e[1] == 'k' ? 1 : e[0] == 's' && e[1] == 't' ? 2 : 0;
#endif
  // persistent: one CTA per SM (1 CTA fits: 227 KB of shared memory, all 512 TMEM columns), items drawn dynamically
  const unsigned grid = p.persistent ? (unsigned)std::min(p.total_items, num_sms()) : (unsigned)p.total_items;
  if (drop.thr) fa_bwd_kernel<D, OutT, true><<<grid, kBwdThreads, Cfg::kSmemBytes, st>>>(mq, mk, mv, mdo, mdq, p);
  else fa_bwd_kernel<D, OutT, false><<<grid, kBwdThreads, Cfg::kSmemBytes, st>>>(mq, mk, mv, mdo, mdq, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (!kF32) {
    fa_bwd_dq_convert_kernel<<<148 * 8, 256, 0, st>>>(dq_acc, static_cast<__nv_bfloat16*>(dq), n, accumulate_dq ? 1 : 0);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace

bool tc_bwd_supported(int d) { return d == 64 || d == 128; }
// kernels one tc_bwd call launches: prep + main (+ the fp32 -> bf16 dQ conversion pass for bf16 gradients)
int tc_bwd_num_launches(bool grads_f32) {
  return grads_f32 ? 2 : 3;
}
void tc_bwd_set_trace(long long* buf) { g_bwd_trace = buf; }
void tc_bwd_set_trace_start(int t0) { g_bwd_trace_t0 = t0; }
void tc_bwd_set_persistent(int on) { g_bwd_persistent = on ? 1 : 0; }
int tc_bwd_get_persistent() { return bwd_persistent() ? 1 : 0; }

size_t tc_bwd_workspace_bytes(int BH, int T, int d) {
  // work-item counter + lse_pad + delta_pad + fp32 dQ accumulator (the latter only used for bf16 gradients)
  return 256 + 2 * align256((size_t)BH * pad_rows(T) * 4) + align256((size_t)BH * T * d * 4);
}

cudaError_t tc_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* o,
                   const __nv_bfloat16* dO, const float* lse, void* dq, void* dk, void* dv, bool grads_f32,
                   bool accumulate_dq, void* workspace, int BH, int T, int d, bool causal, cudaStream_t st,
                   Dropout drop) {
  // the work-item ring packs (head << 16) | K/V tile
  if (!workspace || BH >= 32768 || (T + kTile - 1) / kTile >= 65536) return cudaErrorInvalidValue;
  if (d == 128) {
    return grads_f32 ? launch_bwd<128, float>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, drop, st)
                     : launch_bwd<128, __nv_bfloat16>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, drop, st);
  }
  if (d == 64) {
    return grads_f32 ? launch_bwd<64, float>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, drop, st)
                     : launch_bwd<64, __nv_bfloat16>(q, k, v, o, dO, lse, dq, dk, dv, accumulate_dq, workspace, BH, T, causal, drop, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace fa
