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
//   row_key   = mix(head_key ^ (i >> 1) * 0x9E3779B1)                                 (two rounds: the row term is mixed
//   bits      = mix(row_key ^ (j >> 1) * 0x85EBCA77)     one 32-bit word per 2 x 2     before the column term comes in, so
//                                                         block                         the hash input is not an XOR of two
//                                                                                       linear terms)
//   keep(i,j) = (byte[(i & 1) * 2 + (j & 1)] of bits) & 0x7f  >=  thr,    thr = round(128 p)   (p_eff = thr / 128)
// mix() is the 32-bit multiply-xorshift finaliser "lowbias32" (a bijection with full avalanche), i.e. a keyed
// counter-mode generator.  Seven-bit uniforms make the four comparisons of a word one add:
//   w = (bits & 0x7f7f7f7f) + (0x80 - thr) * 0x01010101     ->  the top bit of byte e is keep(e)   (no carries)
// and PRMT's sign-replicate mode expands those top bits straight into AND-masks for packed bf16x2 / fp32 words.
// Two adjacent lanes (row pair in the forward, key pair in the backward) share every 2 x 2 block, so each
// computes half of the hashes and they swap with one shuffle: ~5 integer ops per mask bit in the softmax warps.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD inline
#endif

namespace fa {

struct Dropout {
  uint32_t thr;      // 0 = dropout disabled; keep iff (byte & 0x7f) >= thr
  float keep_scale;  // 1 / (1 - thr / 128)
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
// key of the row pair i >> 1 (the forward computes it once per thread, the transposed backward once per hashed word)
FA_HD uint32_t drop_row_key(uint32_t head_key, uint32_t i_pair) { return drop_mix(head_key ^ (i_pair * kDropRowMul)); }
FA_HD uint32_t drop_block_bits(uint32_t head_key, uint32_t i, uint32_t j) {
  return drop_mix(drop_row_key(head_key, i >> 1) ^ ((j >> 1) * kDropColMul));
}
// top bit of byte e (e = (i & 1) * 2 + (j & 1)) of the result = keep of that element of the 2 x 2 block
FA_HD uint32_t drop_keep_word(uint32_t bits, uint32_t thr) { return (bits & 0x7f7f7f7fu) + (0x80u - thr) * 0x01010101u; }
FA_HD bool drop_keep(uint32_t head_key, uint32_t i, uint32_t j, uint32_t thr) {
  const uint32_t w = drop_keep_word(drop_block_bits(head_key, i, j), thr);
  return ((w >> (8u * ((i & 1u) * 2u + (j & 1u)) + 7u)) & 1u) != 0u;
}
#if defined(__CUDACC__)
// PRMT with sign-replicate nibbles (8 | byte index): result byte = 0xff if the selected byte's top bit is set else 0x00
__device__ __forceinline__ uint32_t drop_prmt(uint32_t w, uint32_t selector) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "r"(selector));
  return r;
}
// The same over a PAIR of words: bytes 0-3 = a, 4-7 = b.  The kernels hash one block word per lane and get the partner
// lane's word by shuffle; which of the two belongs to which element pair depends on the lane parity, and folding that
// choice into the selector (byte index + 4) saves the two selects per four elements.
__device__ __forceinline__ uint32_t drop_prmt2(uint32_t a, uint32_t b, uint32_t selector) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(selector));
  return r;
}
// selector expanding the keep bits of bytes b0 and b1 into the low / high 16 bits of a mask word (packed bf16x2)
__device__ __forceinline__ uint32_t drop_sel_pair(uint32_t b0, uint32_t b1) {
  return (8u | b0) * 0x11u | ((8u | b1) * 0x11u << 8);
}
// selector expanding the keep bit of byte b into a full 32-bit mask (fp32 / tf32 words)
__device__ __forceinline__ uint32_t drop_sel_word(uint32_t b) { return (8u | b) * 0x1111u; }
#endif

inline Dropout make_dropout(float p, unsigned long long seed) {
  Dropout d;
  int thr = (int)(p * 128.0f + 0.5f);
  if (!(p > 0.f)) thr = 0;
  if (thr > 127) thr = 127;
  d.thr = (uint32_t)thr;
  d.keep_scale = 128.0f / (128.0f - (float)thr);
  d.seed_lo = (uint32_t)(seed & 0xffffffffull);
  d.seed_hi = (uint32_t)(seed >> 32);
  return d;
}

}  // namespace fa
