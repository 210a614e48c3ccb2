// fa_fwd_sm100.cu -- flash-attention forward for sm_100a: tcgen05.mma + TMEM + TMA.
//
// Replaces the reference kernels src/flash_attn2_fw.cu:23-297 (FA-2 forward) and
// src/flash_attn_fw.cu:22-287 (FA-1 forward; same O, different saved statistics).
//
// A work item is 256 consecutive query rows of one (batch, head): two 128-row Q tiles that share
// every K/V tile streamed through shared memory.  The kernel is persistent: one CTA per SM draws items from a
// counter and keeps its K/V ring and its S / P / O hand-offs running across them (see the comment in the kernel).
// Warp roles:
//   warps 0-3   softmax group for Q tile 0   (thread r owns row r: TMEM lane r)
//   warps 4-7   softmax group for Q tile 1
//   warp  8     scheduler + TMA producer (one lane): per item Q, then the K/V ring
//   warp  9     MMA issuer   (one lane): S_t = Q_t K_j^T and O_t += P_t V_j
// (registers are moved from warps 8-11 to the softmax warps with setmaxnreg: 208 vs 80 per thread)
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [384,384+D).  P_t (bf16) overwrites the
// first 64 columns of S_t and is consumed as the TMEM A operand of the PV MMA, so probabilities
// never touch shared memory.  The MMA order  PV0(j) QK0(j+1) PV1(j) QK1(j+1)  keeps the tensor
// pipe busy with one tile while the other tile's softmax runs.
//
// Online softmax: running reference max m_ref per row; O is only rescaled in TMEM when the new
// row max exceeds m_ref by more than 2^8 in the exp2 domain (exact: the final 1/l normalisation
// and the saved logsumexp use the same m_ref), so the usual per-step O correction disappears
// after the first tiles.
#include <cuda.h>
#include <stdlib.h>
#include <math.h>

#include <algorithm>
#include <atomic>
#include <mutex>

#include "fa_internal.h"
#include "fa_ptx.cuh"
#include "fa_tma_host.h"

namespace fa {

long long* g_fwd_trace = nullptr;  // developer timeline buffer (tc_fwd_set_trace)
void tc_fwd_set_trace(long long* buf) { g_fwd_trace = buf; }

// Work-item counters of the persistent forward: {next item, finished CTAs} per slot.  The kernel leaves its slot zeroed
// (the last CTA resets it), so a launch needs no memset; consecutive launches rotate through the slots so that launches
// in flight on different streams never share one (up to kFwdSchedSlots concurrent launches per device).
constexpr int kFwdSchedSlots = 256;
__device__ int g_fwd_sched[kFwdSchedSlots][2];
int g_fwd_persistent = -1;  // -1: FA_B200_FWD_PERSISTENT (read once), default on
bool fwd_persistent() {
  if (g_fwd_persistent < 0) {
    const char* e = getenv("FA_B200_FWD_PERSISTENT");
    g_fwd_persistent = (e && e[0] == '0') ? 0 : 1;
  }
  return g_fwd_persistent != 0;
}
void tc_fwd_set_persistent(int on) { g_fwd_persistent = on ? 1 : 0; }
int tc_fwd_get_persistent() { return fwd_persistent() ? 1 : 0; }

namespace {

int* fwd_sched_slot() {
  static std::atomic<unsigned> next{0};
  static std::mutex mu;
  static int* base[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!base[dev]) {
      void* ptr = nullptr;
      if (cudaGetSymbolAddress(&ptr, g_fwd_sched) != cudaSuccess) return nullptr;
      base[dev] = static_cast<int*>(ptr);
    }
  }
  return base[dev] + 2 * (next.fetch_add(1) % kFwdSchedSlots);
}
int fwd_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

#ifndef FA_FWD_HEAD_GROUP
#define FA_FWD_HEAD_GROUP 32  // same-box A/B at 64 | 128 heads: 16: 0.243 | 0.468 ms, 32: 0.237 | 0.467, 64: 0.243 | 0.485
#endif
constexpr int kHeadGroup = FA_FWD_HEAD_GROUP;  // heads whose work items are interleaved longest-first (see `decode` in the kernel)
constexpr int kBlockM = 128;  // rows per Q tile (= TMEM lanes)
constexpr int kBlockN = 128;  // keys per K/V tile
constexpr int kFwdThreads = 384;  // warps 0-7 softmax, 8 TMA, 9 MMA, 10-11 idle (complete the 4-warp group for setmaxnreg)
#ifndef FA_FWD_PACKED
#define FA_FWD_PACKED 1  // FFMA2 / FADD2 in the softmax inner loop (measured: +2 % over scalar)
#endif
// Speculative exponent reference (FA_FWD_SPEC 1): the first 32 columns of a score tile are exponentiated against the
// reference max of the PREVIOUS step while the row max of the new tile is still being reduced (the lazy rescale keeps that
// reference valid unless the max grows by more than 2^8, in which case the chunk is recomputed).  It takes the 62-deep
// FMNMX3 tree off the S -> P -> PV -> QK chain, but measured (same-box A/B, config 3): 2.05-2.09 ms against 2.00-2.03 ms
// without it -- the interleaved max tree slows the first chunk by what it saves.  Off by default.
#ifndef FA_FWD_SPEC
#define FA_FWD_SPEC 0
#endif
// P is handed to the MMA warp in FA_FWD_PSIG pieces (1 or 2): with 2, the PV k-steps of the first 64 keys run under the
// exponentials of the last 64 (measured: 2.00-2.03 ms against 2.05-2.06 ms with one piece; four pieces cost more in
// tcgen05.wait::st + arrives than they hide, tools/experiments/fwd_quarter_pv_pingpong.patch).
#ifndef FA_FWD_PSIG
#define FA_FWD_PSIG 2
#endif
static_assert(FA_FWD_PSIG == 1 || FA_FWD_PSIG == 2, "the softmax warps signal P whole or in two halves (4 pieces: no gain in round 1, code removed)");

struct FwdParams {
  int T, BH, causal;
  int nblk;         // 256-row query blocks per head
  int total_items;  // BH * nblk work items
  int persistent;   // 1: CTAs draw items from sched[0] until it runs dry (sched[1] counts finished CTAs: the last one
                    // resets both words, so the slot is zero again for the next launch); 0: CTA b works on item b only
  int* sched;
  int wide_stores;  // o is 32-byte aligned: 256-bit stores in the epilogue
  int head_group;   // heads whose work items are interleaved longest-first (1 = head by head), see `decode`
  float tau;         // 1/sqrt(d)
  float scale_log2;  // tau * log2(e)
  void* o;
  float* lse;
  float* m_out;
  float* l_out;
  Dropout drop;  // used by the kDrop instantiations only
  long long* trace;  // developer timeline (tools/fa_check tracefwd): clock64 stamps of CTA (0, 0), 16 slots per K/V step
};

// kTF32 = false: bf16 storage, kind::f16 MMAs (UMMA_K = 16).  kTF32 = true: fp32 storage read directly by TMA,
// kind::tf32 MMAs (UMMA_K = 8); P is written back to TMEM as tf32 over all 128 columns of S and V is consumed
// MN-major through the 128B/32B-atom swizzle that UMMA requires for 32-bit MN-major operands.
template <int D, bool kTF32>
struct FwdCfg {
  static constexpr int kElem = kTF32 ? 4 : 2;
  static constexpr int kAtom = 128 / kElem;               // elements per 128-byte swizzle row (TMA box width)
  static constexpr int kBoxes = D / kAtom;                // boxes per tile row
  static constexpr int kBoxBytes = 128 * 128;             // 128 rows x 128 B
  static constexpr int kTileBytes = kBoxes * kBoxBytes;   // one 128 x D tile
  static constexpr int kStages = kTileBytes >= 32768 ? 2 : 4;
  static constexpr int kSmemBytes = (2 + 2 * kStages) * kTileBytes + 1024;
  static constexpr int kKStepsQK = D * kElem / 32;        // UMMA_K covers 32 bytes of the contraction dim
  static constexpr int kKStepsPV = kBlockN * kElem / 32;
  static_assert(kTileBytes <= 32768, "tile too large for the two-Q-tile layout");
};

struct FwdBarriers {
  uint64_t q_full, q_free;                  // per work item
  uint64_t sched_full[2], sched_empty[2];   // work-item ring
  int sched_item[2];  // (head << 16) | query block of the item, -1 = no more work (decoded once, by the scheduler lane)
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint64_t s_full[2], p_full[2][FA_FWD_PSIG], o_full[2];
};

// 32 consecutive output columns of one row -> global memory.  A thread owns a row, so a warp's store instruction touches
// 32 different lines; the 256-bit stores of sm_100 (STG.256, `wide`: the caller's pointer is 32-byte aligned) halve the
// number of instructions and tag lookups.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <typename OutT>
__device__ __forceinline__ void store_row_chunk32(OutT* dst, const float* x, bool wide);

template <>
__device__ __forceinline__ void store_row_chunk32<float>(float* dst, const float* x, bool wide) {
  if (wide) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) st_global_v8(dst + i, reinterpret_cast<const uint32_t*>(x) + i);
  } else {
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
  }
}
template <>
__device__ __forceinline__ void store_row_chunk32<__nv_bfloat16>(__nv_bfloat16* dst, const float* x, bool wide) {
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
  if (wide) {
    st_global_v8(dst, v);
    st_global_v8(dst + 16, v + 8);
  } else {
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(dst + 2 * i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}

template <int D, typename OutT, bool kTF32, bool kDrop>
__global__ void __launch_bounds__(kFwdThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
              const __grid_constant__ CUtensorMap map_v, const FwdParams p) {
  using Cfg = FwdCfg<D, kTF32>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * Cfg::kTileBytes;
  uint8_t* sV = sK + S * Cfg::kTileBytes;
  __shared__ FwdBarriers bars;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = p.T;
  const int nkv = (T + kBlockN - 1) / kBlockN;
  const int nblk = p.nblk;
  long long* const tr = blockIdx.x == 0 ? p.trace : nullptr;
// developer timeline (tools/fa_check tracefwd).  FA_B200_FWD_TRACE_HOOKS: bit 0 = the sites of the softmax warps, bits 1-3 =
// the sites of the MMA warp (slot 8: before PV of tile 0 | slots 9, 10: before PV of tile 1 | slot 11: end of the K/V step); a
// full timeline needs a -DFA_B200_FWD_TRACE_HOOKS=15 build (tools/build_variant.sh).  The shipped value is what measured
// fastest at config 3, interleaved same-box rounds (tools/batches/r02bs.sh, r02cd.sh):
//   all sites 1.891-1.907 ms | softmax sites only 1.933-1.971 | none 1.915-1.926 | MMA-warp sites only 1.876-1.884      (box A)
//   none 1.857-1.869 | slot 8 only 1.854-1.864 | 9, 10 only 1.850-1.885 | 11 only 1.857-1.872 | 8 + 9, 10: 1.816-1.819 |
//   8 + 11: 1.827-1.838 | 9, 10 + 11: 1.852-1.873 | all three 1.820-1.827                                                   (box B)
// The ~8 predicate instructions of a null hook cost the softmax warps 1 %; in the MMA warp, ahead of BOTH PV issues, they make the
// kernel 2 % FASTER.  The generated issue code is the same apart from the hooks (same UTCHMMA / R2UR / ELECT / branch counts),
// so it is their delay that acts; a suspend-time hint on the issuer's barrier waits (fewer polls instead of later ones) is
// 1 % slower, uniform-datapath issue (fewer instructions per MMA) 1 % slower too.  Unexplained; recorded in DESIGN.md 3.2e.
#ifndef FA_B200_FWD_TRACE_HOOKS
#define FA_B200_FWD_TRACE_HOOKS 6
#endif
#if FA_B200_FWD_TRACE_HOOKS  // bit 0: sites of the softmax warps (slots 0-7), bits 1-3: sites of the MMA warp (slot 8 | 9, 10 | 11)
#define FA_TRF(j_, e_) do { if ((FA_B200_FWD_TRACE_HOOKS & ((e_) == 8 ? 2 : (e_) == 9 || (e_) == 10 ? 4 : (e_) == 11 ? 8 : 1)) && tr && m == 0 && (j_) < 16 && lane == 0) tr[(j_) * 16 + (e_)] = clock64(); } while (0)
#else
#define FA_TRF(j_, e_) do { } while (0)
#endif

  // The kernel is PERSISTENT: one CTA per SM draws work items (one 256-row query block of one head, heaviest causal blocks
  // of a head first) from a global counter.  The K/V ring, the S / P / O hand-offs and their mbarrier parities run on
  // RUNNING counters, so nothing drains between items: Q of the next item is loaded as soon as the last QK of this one has
  // been issued (one K/V step before the item ends), its first scores are in TMEM before the softmax warps have stored O.
  // Item order: heads in groups of kHeadGroup; inside a group the last (longest, when causal) query blocks of all the
  // group's heads come first, then the next-to-last ones, ... -- the counter hands the longest items out first ACROSS
  // heads, so the last items a CTA can draw are the shortest (simulated schedule efficiency at 64 heads, the
  // strong-scaling share of one of 8 GPUs: 0.86 head by head, 0.98 in groups) while the group's Q / K / V / O
  // (32 heads x 4 MB) stay inside the L2.  Measured (fa_check time, T = 4096, d = 128, causal): 64 heads 0.237 against
  // 0.270 ms head by head; 512 heads (55 items per CTA: the tail is 2 % either way) 1.87 against 1.85 ms -- the group
  // order costs a little L2 locality -- so the launcher groups only when a CTA draws fewer than 32 items.
  auto encode = [&](int item) {  // scheduler lane only: item number -> (head << 16) | block
    if (item >= p.total_items) return -1;
    const int per_group = p.head_group * nblk;
    const int grp = item / per_group, r = item - grp * per_group;
    const int gs = min(p.head_group, p.BH - grp * p.head_group);
    const int x = r / gs;
    const int bh = grp * p.head_group + (r - x * gs);
    return (bh << 16) | (p.causal ? nblk - 1 - x : x);
  };
  auto decode = [&](int code, int& bh, int& blk, int& n0, int& n1) {
    bh = code >> 16;
    blk = code & 0xffff;
    n0 = p.causal ? min(2 * blk + 1, nkv) : nkv;
    n1 = p.causal ? min(2 * blk + 2, nkv) : nkv;
  };
  auto next_item = [&](int m) {
    mbar_wait(&bars.sched_full[m & 1], (m >> 1) & 1);
    const int item = *reinterpret_cast<volatile int*>(&bars.sched_item[m & 1]);
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars.sched_empty[m & 1]);
    return item;
  };

  if (threadIdx.x == 0) {
    mbar_init(&bars.q_full, 1);
    mbar_init(&bars.q_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.sched_full[i], 1);
      mbar_init(&bars.sched_empty[i], 9);  // 8 softmax warps + the MMA warp
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(&bars.k_full[i], 1);
      mbar_init(&bars.k_empty[i], 1);
      mbar_init(&bars.v_full[i], 1);
      mbar_init(&bars.v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars.s_full[t], 1);
      for (int h = 0; h < FA_FWD_PSIG; ++h) mbar_init(&bars.p_full[t][h], 4);  // one arrival per softmax warp
      mbar_init(&bars.o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;

  if (warp >= 8) {
    // All four warps of this group must execute the SAME setmaxnreg instruction (one code location),
    // and it must sit inside the role branch for ptxas to apply the budget to the right code.
    reg_dealloc<80>();
  if (warp == 8) {
    // ------------------------------------------------------------------ scheduler + TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      int item = p.persistent ? atomicAdd(&p.sched[0], 1) : (int)blockIdx.x;
      uint32_t jt = 0;  // running K/V step count: ring stage and parity
      for (int m = 0;; ++m) {
        mbar_wait(&bars.sched_empty[m & 1], ((m >> 1) & 1) ^ 1);  // every consumer has read item m - 2
        const int code = encode(item);
        bars.sched_item[m & 1] = code;
        mbar_arrive(&bars.sched_full[m & 1]);
        if (code < 0) break;
        const int next = p.persistent ? atomicAdd(&p.sched[0], 1) : p.total_items;
        int bh, blk, n0, n1;
        decode(code, bh, blk, n0, n1);
        if (m > 0) mbar_wait(&bars.q_free, (m - 1) & 1);  // the last QK of the previous item has read Q
        mbar_arrive_expect_tx(&bars.q_full, 2 * Cfg::kTileBytes);
        for (int t = 0; t < 2; ++t)
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_3d(sQ + t * Cfg::kTileBytes + b * Cfg::kBoxBytes, &map_q, &bars.q_full, b * Cfg::kAtom,
                        (2 * blk + t) * kBlockM, bh, kEvictFirst);
        for (int j = 0; j < n1; ++j, ++jt) {
          const int st = jt % S;
          const uint32_t ph = (jt / S) & 1;
          mbar_wait(&bars.k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&bars.k_full[st], Cfg::kTileBytes);
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_3d(sK + st * Cfg::kTileBytes + b * Cfg::kBoxBytes, &map_k, &bars.k_full[st], b * Cfg::kAtom, j * kBlockN,
                        bh, kEvictLast);
          mbar_wait(&bars.v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&bars.v_full[st], Cfg::kTileBytes);
          for (int b = 0; b < Cfg::kBoxes; ++b)
            tma_load_3d(sV + st * Cfg::kTileBytes + b * Cfg::kBoxBytes, &map_v, &bars.v_full[st], b * Cfg::kAtom, j * kBlockN,
                        bh, kEvictLast);
        }
        item = next;
      }
      if (p.persistent) {
        // every CTA draws exactly one item >= total; the last CTA to get here leaves the slot zeroed for the next launch
        __threadfence();
        if (atomicAdd(&p.sched[1], 1) == (int)gridDim.x - 1) {
          p.sched[0] = 0;
          p.sched[1] = 0;
          __threadfence();
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp runs the issue code
    // tf32 variants: TMEM base and running counters made provably warp-uniform (REDUX), which puts the issue loops on the
    // uniform datapath (fa_bwd_sm100.cu): config 2 0.273 -> 0.272 ms, the MT attention shape (BH = 88, T = 512,
    // d = 32) 0.023 -> 0.022 ms.  The bf16 variants measured 1 % SLOWER that way (1.83-1.84 -> 1.85 ms) and keep the old code.
    const uint32_t tmem_loaded = tmem;
    const uint32_t tmem = kTF32 ? __reduce_max_sync(0xffffffffu, tmem_base_smem) : tmem_loaded;
    // (warp-uniform descriptors, one elected lane issues each MMA / commit: the `_w` forms of fa_ptx.cuh)
    {
      constexpr uint32_t kFmt = kTF32 ? kFmtTF32 : kFmtBF16;
      constexpr uint32_t idesc_qk = make_idesc(kFmt, kBlockM, kBlockN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmt, kBlockM, D, 0, 1);
      const uint32_t sQa = smem_u32(sQ), sKa = smem_u32(sK), sVa = smem_u32(sV);
      // one base descriptor per tile, a k-step adds a compile-time constant to its low word (desc_advance)
      auto issue_qk = [&](int t, int st) {
        const uint64_t q0 = make_smem_desc(sQa + t * Cfg::kTileBytes, 16, 1024);
        const uint64_t k0 = make_smem_desc(sKa + st * Cfg::kTileBytes, 16, 1024);
#pragma unroll
        for (int kk = 0; kk < Cfg::kKStepsQK; ++kk) {
          const uint32_t off = (kk >> 2) * Cfg::kBoxBytes + (kk & 3) * 32;  // 4 k-steps of 32 B per swizzle row
          umma_ss_w<kTF32>(tmem + t * 128, desc_advance(q0, off), desc_advance(k0, off), idesc_qk, kk > 0);
        }
      };
      auto issue_pv = [&](int t, int st, bool acc, uint32_t par) {
        const uint32_t va = sVa + st * Cfg::kTileBytes;
        // bf16: 16 key rows (2 KB) per k-step, 64-column atoms one box apart, 128B swizzle.
        // tf32: P is one tf32 per TMEM column, 8 columns per k-step; V: 8 key rows (1 KB) per k-step, 4-row groups
        // 512 B apart, 32-column atoms one box apart, 128B swizzle with 32-B atomicity.
        const uint64_t v0 = kTF32 ? make_smem_desc(va, Cfg::kBoxBytes, 512, kDescSw128Base32) : make_smem_desc(va, Cfg::kBoxBytes, 1024);
        constexpr int kPerPiece = Cfg::kKStepsPV / FA_FWD_PSIG;
#pragma unroll
        for (int h = 0; h < FA_FWD_PSIG; ++h) {
          mbar_wait_w(&bars.p_full[t][h], par);  // piece h of P_t is in TMEM
          tc_fence_after();
#pragma unroll
          for (int k2 = 0; k2 < kPerPiece; ++k2) {
            const int kk = h * kPerPiece + k2;
            umma_ts_w<kTF32>(tmem + 256 + t * 128, tmem + t * 128 + kk * 8, desc_advance(v0, kk * (kTF32 ? 1024 : 2048)), idesc_pv,
                             (acc || kk > 0) ? 1u : 0u);
          }
        }
      };
      uint32_t jt = 0;          // running K/V step count (ring stage / parity)
      uint32_t c0 = 0, c1 = 0;  // running step counts of the two Q tiles (parities of their P / O hand-offs)
      for (int m = 0;; ++m) {
        // the max-reduction makes the item (and the loop bounds derived from it) provably warp-uniform for ptxas
        const int item = __reduce_max_sync(0xffffffffu, next_item(m));
        if (item < 0) break;
        int bh, blk, n0, n1;
        decode(item, bh, blk, n0, n1);
        if constexpr (kTF32) {  // loop-carried across items, where ptxas loses their uniformity
          jt = __reduce_max_sync(0xffffffffu, jt);
          c0 = __reduce_max_sync(0xffffffffu, c0);
          c1 = __reduce_max_sync(0xffffffffu, c1);
        }
        mbar_wait_w(&bars.q_full, m & 1);
        mbar_wait_w(&bars.k_full[jt % S], (jt / S) & 1);
        tc_fence_after();
        issue_qk(0, jt % S);
        umma_commit_w(&bars.s_full[0]);
        issue_qk(1, jt % S);
        umma_commit_w(&bars.s_full[1]);
        umma_commit_w(&bars.k_empty[jt % S]);
        if (n1 == 1) umma_commit_w(&bars.q_free);
        for (int j = 0; j < n1; ++j, ++jt) {
          const int st = jt % S;
          const uint32_t ph = (jt / S) & 1;
          const int jn = j + 1, stn = (jt + 1) % S;
          const uint32_t phn = ((jt + 1) / S) & 1;
          mbar_wait_w(&bars.v_full[st], ph);
          if (j < n0) {
            FA_TRF(j, 8);
            issue_pv(0, st, j > 0, c0 & 1);
            umma_commit_w(&bars.o_full[0]);
            ++c0;
          }
          if (jn < n0) {
            mbar_wait_w(&bars.k_full[stn], phn);
            tc_fence_after();
            issue_qk(0, stn);
            umma_commit_w(&bars.s_full[0]);
          }
          FA_TRF(j, 9);
          FA_TRF(j, 10);
          issue_pv(1, st, j > 0, c1 & 1);
          umma_commit_w(&bars.o_full[1]);
          umma_commit_w(&bars.v_empty[st]);
          ++c1;
          if (jn < n1) {
            mbar_wait_w(&bars.k_full[stn], phn);
            tc_fence_after();
            issue_qk(1, stn);
            umma_commit_w(&bars.s_full[1]);
            umma_commit_w(&bars.k_empty[stn]);
            if (jn == n1 - 1) umma_commit_w(&bars.q_free);  // the last QK of the item: Q of the next one may land
          }
          FA_TRF(j, 11);
        }
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ softmax groups
    reg_alloc<208>();  // 256*208 + 128*80 <= 384*168 (the CTA register pool fixed at launch)
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + t * 128 + lane_base;
    const uint32_t tO = tmem + 256 + t * 128 + lane_base;
    const float c2 = p.scale_log2;
    const uint32_t drop_thr = p.drop.thr;
    uint32_t cs = 0;  // running step count of this Q tile
    for (int m = 0;; ++m) {
      const int item = next_item(m);
      if (item < 0) break;
      int bh, blk, n0, n1;
      decode(item, bh, blk, n0, n1);
      const int qt = 2 * blk + t;
      const int nt = t == 0 ? n0 : n1;
      const int row_g = qt * kBlockM + r;
      float m_ref = -INFINITY;   // reference max (raw S units) used in the exponent
      float m_true = -INFINITY;  // true running row max (only needed for the FA-1 statistics)
      float l = 0.f;
      // dropout: one 32-bit hash per 2 x 2 block of (row, key); this thread uses the two bytes of its row parity
      const uint32_t drop_rowkey = drop_row_key(drop_head_key(p.drop.seed_lo, p.drop.seed_hi, bh), uint32_t(row_g) >> 1);
      const uint32_t drop_b0 = (row_g & 1) * 2;  // bytes b0, b0 + 1 of a block word belong to this row
      // mask selectors over the word pair (mine, theirs): columns cc, cc+1 take the even lane's word, cc+2, cc+3 the odd lane's
      const uint32_t drop_o0 = (lane & 1) * 4, drop_o1 = 4 - drop_o0;  // where the word of columns cc,cc+1 / cc+2,cc+3 sits
      const uint32_t drop_sel2a = drop_sel_pair(drop_b0 + drop_o0, drop_b0 + 1 + drop_o0),
                     drop_sel2b = drop_sel_pair(drop_b0 + drop_o1, drop_b0 + 1 + drop_o1);
      const uint32_t drop_selA0 = drop_sel_word(drop_b0 + drop_o0), drop_selB0 = drop_sel_word(drop_b0 + 1 + drop_o0),
                     drop_selA1 = drop_sel_word(drop_b0 + drop_o1), drop_selB1 = drop_sel_word(drop_b0 + 1 + drop_o1);

      for (int j = 0; j < nt; ++j, ++cs) {
        mbar_wait(&bars.s_full[t], cs & 1);
        tc_fence_after();
        if ((warp & 3) == 0) FA_TRF(j, 4 * t + 0);
        float s[128];
        {
          uint32_t* sr = reinterpret_cast<uint32_t*>(s);
          tmem_ld_x32(tS + 0, sr + 0);
          tmem_ld_x32(tS + 32, sr + 32);
          tmem_ld_x32(tS + 64, sr + 64);
          tmem_ld_x32(tS + 96, sr + 96);
          tmem_wait_ld();
        }
        if ((warp & 3) == 0) FA_TRF(j, 4 * t + 1);
        const bool diag = p.causal && (j == qt);
        const bool tail = (j == nkv - 1) && (T - j * kBlockN < kBlockN);
        if (diag || tail) {
          int limit = 127;
          if (diag) limit = r;
          if (tail) limit = min(limit, T - 1 - j * kBlockN);
  #pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c > limit) s[c] = -INFINITY;
        }
        // one chunk = 32 score columns: exponentials against the reference -nm / c2, row-sum partials, packed P words
        constexpr int kPk = kTF32 ? 32 : 16;
        const float2 c2v = make_float2(c2, c2);
        auto exp_chunk = [&](const int q4, const float nm_, uint32_t* pk, float2& la, float2& lb) {
          const float2 nmv = make_float2(nm_, nm_);
  #pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int cc = q4 * 32 + c;
            const float2 x0 = ffma2(make_float2(s[cc], s[cc + 1]), c2v, nmv);
            const float2 x1 = ffma2(make_float2(s[cc + 2], s[cc + 3]), c2v, nmv);
            float2 p0 = make_float2(ex2_mixed(x0.x, cc), ex2_mixed(x0.y, cc + 1));
            float2 p1;
            if (FA_EX2_POLY2_PERIOD > 0 && ((cc >> 2) % (FA_EX2_POLY2_PERIOD > 0 ? FA_EX2_POLY2_PERIOD : 1)) == 0) p1 = ex2_poly2(x1);
            else p1 = make_float2(ex2_mixed(x1.x, cc + 2), ex2_mixed(x1.y, cc + 3));
            if constexpr (kTF32) {  // round P to tf32 once, so that the row sum and the PV product see the same values
              p0 = make_float2(round_tf32(p0.x), round_tf32(p0.y));
              p1 = make_float2(round_tf32(p1.x), round_tf32(p1.y));
            }
            la = fadd2(la, p0);
            lb = fadd2(lb, p1);
            if constexpr (kTF32) {
              pk[c] = __float_as_uint(p0.x);
              pk[c + 1] = __float_as_uint(p0.y);
              pk[c + 2] = __float_as_uint(p1.x);
              pk[c + 3] = __float_as_uint(p1.y);
            } else {
              pk[c >> 1] = pack_bf16x2(p0.x, p0.y);
              pk[(c >> 1) + 1] = pack_bf16x2(p1.x, p1.y);
            }
            if constexpr (kDrop) {  // la / lb are the undropped row sum; only P fed to the PV product is thinned
              // the two lanes of a row pair need the same two block words: each hashes one, a shuffle gives the other
              const uint32_t j2 = (uint32_t(j * kBlockN + cc) >> 1) + (lane & 1);
              const uint32_t mine = drop_keep_word(drop_mix(drop_rowkey ^ (j2 * kDropColMul)), drop_thr);
              const uint32_t theirs = __shfl_xor_sync(0xffffffffu, mine, 1);
              if constexpr (kTF32) {
                pk[c] &= drop_prmt2(mine, theirs, drop_selA0);
                pk[c + 1] &= drop_prmt2(mine, theirs, drop_selB0);
                pk[c + 2] &= drop_prmt2(mine, theirs, drop_selA1);
                pk[c + 3] &= drop_prmt2(mine, theirs, drop_selB1);
              } else {
                pk[c >> 1] &= drop_prmt2(mine, theirs, drop_sel2a);
                pk[(c >> 1) + 1] &= drop_prmt2(mine, theirs, drop_sel2b);
              }
            }
          }
        };
        auto store_chunk = [&](const int q4, const uint32_t* pk) {
          if constexpr (kTF32) tmem_st_x32(tS + q4 * 32, pk);   // tf32: P fills all 128 columns of S
          else tmem_st_x16(tS + q4 * 16, pk);                    // bf16: two probabilities per column
        };
        auto signal_p = [&](const int h) {  // piece h of P_t is complete in TMEM
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.p_full[t][h]);
        };

        // Chunk 0 goes first, against the reference max of the previous step (speculation: independent of the max tree
        // below, so ptxas interleaves the two).  j == 0 has no reference yet (m_ref = -inf gives inf / NaN here): it always
        // takes the recompute branch, like any step whose row max outgrows the reference by more than 2^8.
        uint32_t pk0[kPk];
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
        float nm = -m_ref * c2;
        if (FA_FWD_SPEC) exp_chunk(0, nm, pk0, la, lb);
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
        const bool need = (j == 0) || (mx - m_ref) * c2 > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_ref, mx);
          if (j > 0) {
            // O_t must be quiescent: PV(t, j-1) retired, PV(t, j) not yet issued (it waits for our P).
            mbar_wait(&bars.o_full[t], (cs - 1) & 1);
            tc_fence_after();
            const float f = ex2_approx((m_ref - m_new) * c2);
            l *= f;
  #pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {
              uint32_t orr[32];
              tmem_ld_x32(tO + c0, orr);
              tmem_wait_ld();
  #pragma unroll
              for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
              tmem_st_x32(tO + c0, orr);
            }
          }
          m_ref = m_new;
          nm = -m_ref * c2;
          la = make_float2(0.f, 0.f);
          lb = make_float2(0.f, 0.f);
          if (FA_FWD_SPEC) exp_chunk(0, nm, pk0, la, lb);
        }
        if (!FA_FWD_SPEC) exp_chunk(0, nm, pk0, la, lb);
        if ((warp & 3) == 0) FA_TRF(j, 4 * t + 2);
        store_chunk(0, pk0);
  #pragma unroll
        for (int q4 = 1; q4 < 4; ++q4) {
          uint32_t pk[kPk];
          exp_chunk(q4, nm, pk, la, lb);
          store_chunk(q4, pk);
          if (FA_FWD_PSIG == 2 && q4 == 1) signal_p(0);
        }
        signal_p(FA_FWD_PSIG - 1);
        l += (la.x + la.y) + (lb.x + lb.y);
        if ((warp & 3) == 0) FA_TRF(j, 4 * t + 3);
      }


      // epilogue of the item: O_t / l -> global, statistics.  (The MMA warp already has the first scores of the next item
      // in S_t; it restarts O_t only after the next P hand-off of this group, i.e. after this code.)
      mbar_wait(&bars.o_full[t], (cs - 1) & 1);
      tc_fence_after();
      const float inv_l = (kDrop ? p.drop.keep_scale : 1.f) / l;
      const bool row_ok = row_g < T;
      OutT* orow = reinterpret_cast<OutT*>(p.o) + ((size_t)bh * T + (row_ok ? row_g : 0)) * D;
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t orr[32];
        tmem_ld_x32(tO + c0, orr);
        tmem_wait_ld();
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(orr[i]) * inv_l;
        if (row_ok) store_row_chunk32<OutT>(orow + c0, x, p.wide_stores != 0);
      }
      tc_fence_before();  // orders the TMEM reads above before the next p_full arrival
      if (row_ok) {
        const size_t idx = (size_t)bh * T + row_g;
        if (p.lse) p.lse[idx] = m_ref * p.tau + logf(l);
        if (p.m_out) p.m_out[idx] = m_true * p.tau;
        if (p.l_out) p.l_out[idx] = l * expf((m_ref - m_true) * p.tau);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

template <int D, typename OutT, bool kTF32, bool kDrop>
cudaError_t launch_fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, float* m_out, float* l_out,
                            int BH, int T, bool causal, Dropout drop, cudaStream_t st) {
  using Cfg = FwdCfg<D, kTF32>;
  CUtensorMap mq, mk, mv;
  if (make_tmap_3d(&mq, q, Cfg::kElem, D, T, BH, Cfg::kAtom, kBlockM)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mk, k, Cfg::kElem, D, T, BH, Cfg::kAtom, kBlockN)) return cudaErrorInvalidValue;
  if (make_tmap_3d(&mv, v, Cfg::kElem, D, T, BH, Cfg::kAtom, kBlockN, kTF32 ? 2 : 1)) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(fa_fwd_kernel<D, OutT, kTF32, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes);
  if (e != cudaSuccess) return e;
  FwdParams p;
  p.T = T;
  p.BH = BH;
  p.nblk = (T + 2 * kBlockM - 1) / (2 * kBlockM);
  p.total_items = BH * p.nblk;
  if (BH >= 32768 || p.nblk >= 65536) return cudaErrorInvalidValue;  // the work-item ring packs (head << 16) | block
  p.persistent = fwd_persistent() ? 1 : 0;
  p.sched = nullptr;
  if (p.persistent) {
    p.sched = fwd_sched_slot();
    if (!p.sched) return cudaErrorInvalidValue;
  }
  p.wide_stores = (reinterpret_cast<uintptr_t>(o) & 31) == 0 ? 1 : 0;
  p.causal = causal ? 1 : 0;
  p.tau = sqrtf(1.0f / (float)D);
  p.scale_log2 = p.tau * 1.4426950408889634f;
  p.o = o;
  p.lse = lse;
  p.m_out = m_out;
  p.l_out = l_out;
  p.drop = drop;
  p.trace = g_fwd_trace;
  // persistent: one CTA per SM (1 CTA fits: ~200 KB of shared memory, all 512 TMEM columns), items drawn dynamically
  const unsigned grid = p.persistent ? (unsigned)std::min(p.total_items, fwd_num_sms()) : (unsigned)p.total_items;
  // grouped order only where the tail matters (a CTA draws < 32 items); a group's Q / K / V / O should fit the L2
  const size_t per_head = (size_t)T * D * Cfg::kElem * 4;
  p.head_group = (p.persistent && p.total_items < 32 * (int)grid)
                     ? (int)std::max<size_t>(1, std::min<size_t>(kHeadGroup, ((size_t)128 << 20) / per_head)) : 1;
  fa_fwd_kernel<D, OutT, kTF32, kDrop><<<grid, kFwdThreads, Cfg::kSmemBytes, st>>>(mq, mk, mv, p);
  return cudaGetLastError();
}

template <int D, typename OutT, bool kTF32>
cudaError_t launch_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* m_out, float* l_out,
                       int BH, int T, bool causal, Dropout drop, cudaStream_t st) {
  return drop.thr ? launch_fwd_impl<D, OutT, kTF32, true>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st)
                  : launch_fwd_impl<D, OutT, kTF32, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
}

}  // namespace

bool tc_fwd_supported(int d) { return d == 64 || d == 128; }
// d = 128 can run on CTA pairs (fa_fwd_pair_sm100.cu: cta_group::2, P through shared memory).  Parity-green but measured
// SLOWER than the single-CTA kernel of this file (2.83 against 2.02 ms at config 3), so it is
// opt-in: FA_B200_FWD_PAIR=1 or fa_b200_set_fwd_pair(1).
int g_fwd_pair = -1;
bool fwd_pair_enabled() {
  if (g_fwd_pair < 0) {
    const char* e = getenv("FA_B200_FWD_PAIR");
    g_fwd_pair = (e && e[0] == '1') ? 1 : 0;
  }
  return g_fwd_pair == 1;
}
void set_fwd_pair(bool on) { g_fwd_pair = on ? 1 : 0; }
bool tc_fwd_tf32_supported(int d) { return d == 32 || d == 64; }

cudaError_t tc_fwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, void* o, bool o_is_f32,
                   float* lse, float* m_out, float* l_out, int BH, int T, int d, bool causal, cudaStream_t st,
                   Dropout drop) {
  if (d == 128 && fwd_pair_enabled()) return tc_fwd_pair(q, k, v, o, o_is_f32, lse, m_out, l_out, BH, T, causal, st, drop);
  if (d == 128) {
    return o_is_f32 ? launch_fwd<128, float, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st)
                    : launch_fwd<128, __nv_bfloat16, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
  }
  if (d == 64) {
    return o_is_f32 ? launch_fwd<64, float, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st)
                    : launch_fwd<64, __nv_bfloat16, false>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
  }
  return cudaErrorInvalidValue;
}

// fp32 storage in, fp32 out, tf32 tensor-core arithmetic
cudaError_t tc_fwd_tf32(const float* q, const float* k, const float* v, float* o, float* lse, float* m_out,
                        float* l_out, int BH, int T, int d, bool causal, cudaStream_t st, Dropout drop) {
  if (BH > 65535) return cudaErrorInvalidValue;
  if (d == 64) return launch_fwd<64, float, true>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
  if (d == 32) return launch_fwd<32, float, true>(q, k, v, o, lse, m_out, l_out, BH, T, causal, drop, st);
  return cudaErrorInvalidValue;
}

}  // namespace fa
