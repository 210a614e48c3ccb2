// fa_dropout.cuh -- counter-based dropout mask shared by every kernel (forward, backward, both layouts)
// and restated bit for bit by oracle/attention_oracle.py::dropout_keep_mask.
//
// The reference has no dropout on the flash path; the convention follows its Dropout module
// (minitorch/modules_basic.py:77-80): keep ~ Bernoulli(1 - p), kept probabilities scaled by 1/(1 - p), applied
// to P after the softmax normalisation; the saved logsumexp is that of the undropped scores.
//
// The mask is a pure function of (seed, head, query row i, key column j), so the forward (thread = query row)
// and the transposed backward (thread = key row) regenerate identical bits without storing anything:
//   head_key  = mix(seed_lo ^ mix(seed_hi + 0x9E3779B9 * (head + 1)))
//   bits      = mix(head_key ^ (i >> 1) * 0x9E3779B1 ^ (j >> 1) * 0x85EBCA77)         one 32-bit word per 2 x 2 block
//   keep(i,j) = byte[(i & 1) * 2 + (j & 1)] of bits  >=  thr,    thr = round(256 p)   (p_eff = thr / 256)
// mix() is the 32-bit multiply-xorshift finaliser "lowbias32" (a bijection with full avalanche), i.e. a keyed
// counter-mode generator: 10 integer ops per four mask bits, cheap enough to live next to the exp2 in the
// softmax warps.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD inline
#endif

namespace fa {

struct Dropout {
  uint32_t thr;      // 0 = dropout disabled; keep iff byte >= thr
  float keep_scale;  // 1 / (1 - thr / 256)
  uint32_t seed_lo, seed_hi;
};

constexpr uint32_t kDropRowMul = 0x9E3779B1u;
constexpr uint32_t kDropColMul = 0x85EBCA77u;

FA_HD uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
FA_HD uint32_t drop_head_key(uint32_t seed_lo, uint32_t seed_hi, uint32_t head) {
  return drop_mix(seed_lo ^ drop_mix(seed_hi + 0x9E3779B9u * (head + 1u)));
}
FA_HD uint32_t drop_block_bits(uint32_t head_key, uint32_t i, uint32_t j) {
  return drop_mix(head_key ^ ((i >> 1) * kDropRowMul) ^ ((j >> 1) * kDropColMul));
}
FA_HD bool drop_keep(uint32_t head_key, uint32_t i, uint32_t j, uint32_t thr) {
  const uint32_t bits = drop_block_bits(head_key, i, j);
  return ((bits >> (8u * ((i & 1u) * 2u + (j & 1u)))) & 0xffu) >= thr;
}

inline Dropout make_dropout(float p, unsigned long long seed) {
  Dropout d;
  int thr = (int)(p * 256.0f + 0.5f);
  if (!(p > 0.f)) thr = 0;
  if (thr > 255) thr = 255;
  d.thr = (uint32_t)thr;
  d.keep_scale = 256.0f / (256.0f - (float)thr);
  d.seed_lo = (uint32_t)(seed & 0xffffffffull);
  d.seed_hi = (uint32_t)(seed >> 32);
  return d;
}

}  // namespace fa
