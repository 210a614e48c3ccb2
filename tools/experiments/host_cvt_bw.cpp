// host_cvt_bw.cpp -- how fast can the box's host cores turn fp32 into bf16 (and back)?  Decides whether the host-pointer
// ABI should convert on the CPU (half the PCIe bytes) or on the GPU (what it does).  g++ -O3 -mavx2 -pthread
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void cvt_f32_bf16(const float* src, uint16_t* dst, size_t n) {
  const __m256i bias = _mm256_set1_epi32(0x7fff), one = _mm256_set1_epi32(1);
  size_t i = 0;
  for (; i + 16 <= n; i += 16) {
    __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 8));
    a = _mm256_srli_epi32(_mm256_add_epi32(a, _mm256_add_epi32(bias, _mm256_and_si256(_mm256_srli_epi32(a, 16), one))), 16);
    b = _mm256_srli_epi32(_mm256_add_epi32(b, _mm256_add_epi32(bias, _mm256_and_si256(_mm256_srli_epi32(b, 16), one))), 16);
    __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xd8);
    _mm256_stream_si256((__m256i*)(dst + i), p);
  }
  for (; i < n; ++i) { uint32_t u; memcpy(&u, src + i, 4); dst[i] = (uint16_t)((u + 0x7fff + ((u >> 16) & 1)) >> 16); }
}
static void cvt_bf16_f32(const uint16_t* src, float* dst, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    __m128i h = _mm_loadu_si128((const __m128i*)(src + i));
    _mm256_stream_si256((__m256i*)(dst + i), _mm256_slli_epi32(_mm256_cvtepu16_epi32(h), 16));
  }
  for (; i < n; ++i) { uint32_t u = (uint32_t)src[i] << 16; memcpy(dst + i, &u, 4); }
}
int main(int argc, char** argv) {
  const size_t n = (size_t)1 << 29;  // 2 GB of fp32
  float* a = (float*)aligned_alloc(64, n * 4);
  uint16_t* b = (uint16_t*)aligned_alloc(64, n * 2);
  for (size_t i = 0; i < n; ++i) a[i] = (float)(i & 1023) * 0.001f;
  memset(b, 0, n * 2);
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  for (int T : {1, 4, 8, 16, 32, 64}) {
    if ((unsigned)T > 2 * std::thread::hardware_concurrency()) break;
    for (int dir = 0; dir < 2; ++dir) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t)
        th.emplace_back([=] {
          const size_t lo = n / T * t, hi = t == T - 1 ? n : n / T * (t + 1);
          if (dir == 0) cvt_f32_bf16(a + lo, b + lo, hi - lo); else cvt_bf16_f32(b + lo, a + lo, hi - lo);
        });
      for (auto& x : th) x.join();
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      printf("threads %2d  %s  %.1f GB/s of fp32 (%.3f s)\n", T, dir == 0 ? "f32->bf16" : "bf16->f32", n * 4 / s * 1e-9, s);
    }
  }
  return 0;
}
