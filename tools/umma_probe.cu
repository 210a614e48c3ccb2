// umma_probe.cu -- hardware check of the hand-encoded UMMA / TMA / TMEM conventions in fa_ptx.cuh.
//
// One CTA computes D[128 x N] = A[128 x K] * B[N x K]^T with a single tile, for every operand
// form the attention kernels rely on:
//   A: smem K-major | smem MN-major | TMEM (written with tcgen05.st)
//   B: smem K-major | smem MN-major
//   element: bf16 (kind::f16, UMMA_K = 16) | fp32-as-tf32 (kind::tf32, UMMA_K = 8)
// and compares against a CPU fp64 product.  Descriptor LBO/SBO are runtime parameters so that
// alternative encodings can be swept in a single GPU session.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I flash-attention-minitorch_b200/csrc \
//        tools/umma_probe.cu -o tools/umma_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fa_ptx.cuh"
#include "fa_tma_host.h"

using namespace fa;

enum AForm { A_SMEM_K = 0, A_SMEM_MN = 1, A_TMEM = 2 };
enum BForm { B_SMEM_K = 0, B_SMEM_MN = 1 };

struct ProbeParams {
  int a_form, b_form;
  int M, N, K;          // M = 128
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep_bytes, b_kstep_bytes;  // start-address advance per UMMA_K step *within* a swizzle box
  uint32_t a_box_bytes, b_box_bytes;      // advance when a K step crosses into the next TMA box (K-major only)
  int a_ksteps_per_box, b_ksteps_per_box; // K-major: 4; MN-major: all steps in one "box walk"
  int a_nbox, b_nbox;                     // number of TMA boxes per operand
  int a_box_rows, b_box_rows;
  int a_box_c0_step, b_box_c0_step;       // element step of coordinate 0 between boxes
  uint64_t a_layout, b_layout;            // descriptor layout/version bits
};

template <typename T, bool kTF32>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const T* __restrict__ a_gl, float* __restrict__ d_gl, ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 65536;
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kElem = sizeof(T);
  constexpr int UMMA_K = 32 / kElem;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;          // D accumulator: columns [0, N)
  const uint32_t tmem_a = tmem + 256;    // A operand when A_TMEM

  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (p.a_form != A_TMEM) bytes += p.a_nbox * p.a_box_rows * 128;
    bytes += p.b_nbox * p.b_box_rows * 128;
    mbar_arrive_expect_tx(&bar_load, bytes);
    if (p.a_form != A_TMEM)
      for (int b = 0; b < p.a_nbox; ++b)
        tma_load_3d(sA + b * p.a_box_rows * 128, &map_a, &bar_load, b * p.a_box_c0_step, 0, 0, kEvictNormal);
    for (int b = 0; b < p.b_nbox; ++b)
      tma_load_3d(sB + b * p.b_box_rows * 128, &map_b, &bar_load, b * p.b_box_c0_step, 0, 0, kEvictNormal);
  }
  if (p.a_form == A_TMEM) {
    // thread t owns row t of A: write it to TMEM lanes, packed the way the MMA expects
    const int row = threadIdx.x;
    const uint32_t lane_addr = tmem_a + (uint32_t(warp * 32) << 16);
    if constexpr (kTF32) {
      for (int c0 = 0; c0 < p.K; c0 += 32) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(float(a_gl[row * p.K + c0 + i]));
        tmem_st_x32(lane_addr + c0, r);
      }
    } else {
      for (int c0 = 0; c0 < p.K; c0 += 64) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float lo = float(a_gl[row * p.K + c0 + 2 * i]);
          float hi = float(a_gl[row * p.K + c0 + 2 * i + 1]);
          r[i] = pack_bf16x2(lo, hi);
        }
        tmem_st_x32(lane_addr + c0 / 2, r);
      }
    }
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(kTF32 ? kFmtTF32 : kFmtBF16, 128, p.N, p.a_form == A_SMEM_MN ? 1 : 0,
                                      p.b_form == B_SMEM_MN ? 1 : 0);
    const int nk = p.K / UMMA_K;
    for (int kk = 0; kk < nk; ++kk) {
      uint32_t a_off = (kk / p.a_ksteps_per_box) * p.a_box_bytes + (kk % p.a_ksteps_per_box) * p.a_kstep_bytes;
      uint32_t b_off = (kk / p.b_ksteps_per_box) * p.b_box_bytes + (kk % p.b_ksteps_per_box) * p.b_kstep_bytes;
      uint64_t bdesc = make_smem_desc(smem_u32(sB) + b_off, p.b_lbo, p.b_sbo, p.b_layout);
      if (p.a_form == A_TMEM) {
        const uint32_t cols_per_step = kTF32 ? UMMA_K : UMMA_K / 2;
        umma_ts<kTF32>(tmem_d, tmem_a + kk * cols_per_step, bdesc, idesc, kk > 0);
      } else {
        uint64_t adesc = make_smem_desc(smem_u32(sA) + a_off, p.a_lbo, p.a_sbo, p.a_layout);
        umma_ss<kTF32>(tmem_d, adesc, bdesc, idesc, kk > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  // epilogue: thread t <- row t of D
  {
    const int row = threadIdx.x;
    const uint32_t lane_addr = tmem_d + (uint32_t(warp * 32) << 16);
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(lane_addr + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) d_gl[row * p.N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  (void)lane;
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

template <typename T, bool kTF32>
static int run_case(const char* name, int a_form, int b_form, int N, int K, int variant) {
  const int M = 128;
  constexpr int esize = sizeof(T);
  const int atom = 128 / esize;  // elements per 128-B swizzle row
  const int UMMA_K = 32 / esize;
  std::vector<float> A(M * K), B(N * K);
  srand(1234 + a_form * 7 + b_form * 3 + N + K);
  for (auto& x : A) x = kTF32 ? tf32_trunc((rand() % 2001 - 1000) / 1000.f) : bf16_round((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = kTF32 ? tf32_trunc((rand() % 2001 - 1000) / 1000.f) : bf16_round((rand() % 2001 - 1000) / 1000.f);

  // device storage in the layout the form implies
  std::vector<T> Ah(M * K), Bh(N * K);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      T v = T(A[m * K + k]);
      if (a_form == A_SMEM_MN) Ah[k * M + m] = v; else Ah[m * K + k] = v;
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      T v = T(B[n * K + k]);
      if (b_form == B_SMEM_MN) Bh[k * N + n] = v; else Bh[n * K + k] = v;
    }
  T *dA, *dB;
  float* dD;
  cudaMalloc(&dA, M * K * esize);
  cudaMalloc(&dB, N * K * esize);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, Ah.data(), M * K * esize, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), N * K * esize, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);

  ProbeParams p{};
  p.a_form = a_form; p.b_form = b_form; p.M = M; p.N = N; p.K = K;
  CUtensorMap map_a, map_b;
  p.a_layout = p.b_layout = kDescSw128;
  // variants >= 2: MN-major operands use the 128B/32B-atom swizzle (TMA mode 2, descriptor layout 1)
  const int mn_mode = variant >= 2 ? 2 : 1;
  memset(&map_a, 0, sizeof(map_a));
  // A
  // variant 5: K-major operands ALSO loaded with the 32B-atom swizzle and described with layout 1 -- would let one
  // fp32 tile serve both as a K-major and as an MN-major tf32 operand (needed by a tf32 backward)
  const int k_mode = variant == 5 ? 2 : 1;
  if (a_form == A_SMEM_K || a_form == A_TMEM) {
    // global [M][K], boxes [M rows][atom]
    if (make_tmap_3d(&map_a, dA, esize, K, M, 1, atom, M, k_mode)) return 1;
    p.a_nbox = K / atom; p.a_box_rows = M; p.a_box_c0_step = atom;
    p.a_lbo = 16; p.a_sbo = 1024; p.a_kstep_bytes = 32; p.a_ksteps_per_box = atom / UMMA_K; p.a_box_bytes = M * 128;
    if (variant == 5) p.a_layout = kDescSw128Base32;
  } else {
    // global [K][M], boxes [K rows][atom of M]
    if (make_tmap_3d(&map_a, dA, esize, M, K, 1, atom, K, mn_mode)) return 1;
    p.a_nbox = M / atom; p.a_box_rows = K; p.a_box_c0_step = atom;
    p.a_lbo = K * 128; p.a_sbo = 1024; p.a_kstep_bytes = UMMA_K * 128; p.a_ksteps_per_box = 1 << 20; p.a_box_bytes = 0;
    if (variant == 1) { p.a_lbo = 1024; p.a_sbo = K * 128; }
    if (variant >= 2) p.a_layout = kDescSw128Base32;
    if (variant == 2) { p.a_lbo = K * 128; p.a_sbo = 512; }
    if (variant == 3) { p.a_lbo = 512; p.a_sbo = K * 128; }
    if (variant == 4) { p.a_lbo = K * 128; p.a_sbo = 1024; }
  }
  if (b_form == B_SMEM_K) {
    if (make_tmap_3d(&map_b, dB, esize, K, N, 1, atom, N, k_mode)) return 1;
    p.b_nbox = K / atom; p.b_box_rows = N; p.b_box_c0_step = atom;
    p.b_lbo = 16; p.b_sbo = 1024; p.b_kstep_bytes = 32; p.b_ksteps_per_box = atom / UMMA_K; p.b_box_bytes = N * 128;
    if (variant == 5) p.b_layout = kDescSw128Base32;
  } else {
    if (make_tmap_3d(&map_b, dB, esize, N, K, 1, atom, K, mn_mode)) return 1;
    p.b_nbox = N / atom; p.b_box_rows = K; p.b_box_c0_step = atom;
    p.b_lbo = K * 128; p.b_sbo = 1024; p.b_kstep_bytes = UMMA_K * 128; p.b_ksteps_per_box = 1 << 20; p.b_box_bytes = 0;
    if (variant == 1) { p.b_lbo = 1024; p.b_sbo = K * 128; }
    if (variant >= 2) p.b_layout = kDescSw128Base32;
    if (variant == 2) { p.b_lbo = K * 128; p.b_sbo = 512; }
    if (variant == 3) { p.b_lbo = 512; p.b_sbo = K * 128; }
    if (variant == 4) { p.b_lbo = K * 128; p.b_sbo = 1024; }
  }
  const int smem_bytes = 65536 * 2 + 1024;
  cudaFuncSetAttribute(probe_kernel<T, kTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  probe_kernel<T, kTF32><<<1, 128, smem_bytes>>>(map_a, map_b, dA, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-34s v%d N=%3d K=%3d  CUDA ERROR %s\n", name, variant, N, K, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> D(M * N);
  cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += double(A[m * K + k]) * double(B[n * K + k]);
      double err = fabs(ref - D[m * N + n]);
      if (!(err <= maxerr)) maxerr = err;  // NaN-propagating
    }
  printf("%-34s v%d N=%3d K=%3d  max_abs_err=%.3e  %s\n", name, variant, N, K, maxerr, maxerr < 1e-2 ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-2 ? 0 : 3;
}


// ---------------------------------------------------------------------------------------------
// tcgen05.ld.16x256b register mapping (used by the backward's dQ drain: fragment layout gives each
// quad of lanes 32 contiguous bytes of one row, i.e. sector-complete red.global.add.v2.f32).
// Every TMEM cell (lane, col) is first written with lane * 256 + col through the 32x32b shape.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) ldshape_kernel(uint32_t* __restrict__ out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t mine = tmem + (uint32_t(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 256 + c0 + i;
    tmem_st_x32(mine + c0, r);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    tmem_ld_16x256b_x4(tmem + (uint32_t(warp * 32 + h * 16) << 16) + 8, r);  // columns [8, 40)
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) out[((warp * 2 + h) * 32 + lane) * 16 + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int run_ldshape() {
  uint32_t* d;
  cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
  ldshape_kernel<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("ldshape: CUDA ERROR %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<uint32_t> h(4 * 2 * 32 * 16);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int hf = 0; hf < 2; ++hf)
      for (int t = 0; t < 32; ++t)
        for (int i = 0; i < 16; ++i) {
          const uint32_t v = h[((w * 2 + hf) * 32 + t) * 16 + i];
          const int lane = v >> 8, col = v & 255;
          // hypothesis: repeat n = i / 4 covers columns [8n, 8n+8); regs 4n+{0,1}: row t/4, regs 4n+{2,3}: row t/4 + 8;
          // column pair 2 * (t % 4) + (i & 1)
          const int n = i / 4, exp_lane = w * 32 + hf * 16 + t / 4 + ((i & 2) ? 8 : 0), exp_col = 8 + 8 * n + 2 * (t % 4) + (i & 1);
          if (lane != exp_lane || col != exp_col) {
            if (bad < 24) printf("ldshape mismatch w=%d h=%d t=%2d reg=%2d: got (lane %3d, col %3d) expected (%3d, %3d)\n", w, hf, t, i, lane, col, exp_lane, exp_col);
            ++bad;
          }
        }
  printf("tcgen05.ld.16x256b.x4 fragment mapping: %s (%d mismatches)\n", bad ? "DIFFERENT FROM HYPOTHESIS" : "OK", bad);
  cudaFree(d);
  return bad != 0;
}

int main(int argc, char** argv) {
  int fails = 0;
  fails += run_ldshape();
  typedef __nv_bfloat16 bf;
  const bool full = argc > 1;
  if (full) {
    for (int N : {128, 64}) {
      fails += run_case<bf, false>("bf16 A=smemK  B=smemK", A_SMEM_K, B_SMEM_K, N, 128, 0) != 0;
      fails += run_case<bf, false>("bf16 A=smemK  B=smemK", A_SMEM_K, B_SMEM_K, N, 64, 0) != 0;
      fails += run_case<bf, false>("bf16 A=smemK  B=smemMN", A_SMEM_K, B_SMEM_MN, N, 128, 0) != 0;
      fails += run_case<bf, false>("bf16 A=tmem   B=smemMN", A_TMEM, B_SMEM_MN, N, 128, 0) != 0;
      fails += run_case<bf, false>("bf16 A=smemMN B=smemMN", A_SMEM_MN, B_SMEM_MN, N, 128, 0) != 0;
      fails += run_case<bf, false>("bf16 A=smemMN B=smemK", A_SMEM_MN, B_SMEM_K, N, 128, 0) != 0;
      fails += run_case<bf, false>("bf16 A=tmem   B=smemK", A_TMEM, B_SMEM_K, N, 128, 0) != 0;
      fails += run_case<float, true>("tf32 A=smemK  B=smemK", A_SMEM_K, B_SMEM_K, N, 64, 0) != 0;
      fails += run_case<float, true>("tf32 A=smemK  B=smemK", A_SMEM_K, B_SMEM_K, N, 128, 0) != 0;
    }
  }
  for (int N : {128, 64}) {
    fails += run_case<float, true>("tf32 A=tmem   B=smemK", A_TMEM, B_SMEM_K, N, 128, 0) != 0;
    for (int v : {2, 3, 4}) {
      fails += run_case<float, true>("tf32 A=smemK  B=smemMN(32B atom)", A_SMEM_K, B_SMEM_MN, N, 128, v) != 0;
      fails += run_case<float, true>("tf32 A=tmem   B=smemMN(32B atom)", A_TMEM, B_SMEM_MN, N, 128, v) != 0;
      fails += run_case<float, true>("tf32 A=smemMN B=smemK (32B atom)", A_SMEM_MN, B_SMEM_K, N, 64, v) != 0;
    }
  }
  // forms of the tf32 backward at d = 32: N = 32 accumulators (dV, dK, dQ), A from TMEM or MN-major shared memory,
  // B MN-major; and the score products with K = 32
  fails += run_case<float, true>("tf32 A=smemMN B=smemMN (32B atom)", A_SMEM_MN, B_SMEM_MN, 128, 128, 2) != 0;
  fails += run_case<float, true>("tf32 A=smemMN B=smemMN (32B atom)", A_SMEM_MN, B_SMEM_MN, 32, 128, 2) != 0;
  fails += run_case<float, true>("tf32 A=tmem   B=smemMN(32B atom)", A_TMEM, B_SMEM_MN, 32, 128, 2) != 0;
  fails += run_case<float, true>("tf32 A=smemK  B=smemK", A_SMEM_K, B_SMEM_K, 128, 32, 0) != 0;
  // K-major tf32 operands in the 32B-atom swizzle: measured on B200 -> "misaligned address" from the MMA (the 32-byte
  // k-step advance is not a legal start address in that layout), i.e. one fp32 tile cannot serve both as a K-major and
  // as an MN-major tf32 operand.  The launch faults the context, so it only runs on request: umma_probe full kmajor32
  if (argc > 2) run_case<float, true>("tf32 A=smemK  B=smemK  (both 32B atom)", A_SMEM_K, B_SMEM_K, 128, 64, 5);
  printf("umma_probe: %d failing cases\n", fails);
  return 0;
}
