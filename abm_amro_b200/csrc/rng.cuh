// Counter-based RNG of the free-running mode (DESIGN.md "RNG spec") and the probability -> threshold map.
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), hand-written; key = 64-bit seed, counter =
// (row_lo, row_hi, sim_group, stream).  One call yields 128 bits = eight 16-bit chunks, one per
// simulation of a group of 8 consecutive simulations, for ONE patient-day record.  A Bernoulli(p) draw
// compares the 32-bit uniform (primary chunk << 16 | refinement chunk) with T(p) = #{r : r*2^-32 < p};
// the refinement chunk comes from stream+1 and is only evaluated when the primary chunk ties with the
// upper half of T (probability 2^-16), so the common case costs one Philox call per 8 updates.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace amro {

constexpr uint32_t kStreamStep = 0, kStreamInitCol = 2, kStreamInitDet = 4;  // +1 = refinement stream

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 16-bit chunk of simulation k (0..7) of the group
__device__ __forceinline__ uint32_t chunk16(const uint4& v, int k) {
  const uint32_t w = (k >> 1) == 0 ? v.x : (k >> 1) == 1 ? v.y : (k >> 1) == 2 ? v.z : v.w;
  return (k & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// T(p): number of 32-bit integers r with r * 2^-32 < p.  p <= 0 or NaN -> 0, p >= 1 -> 2^32.
__device__ __forceinline__ uint64_t threshold33(double p) {
  if (!(p > 0.0)) return 0ull;
  if (p >= 1.0) return 1ull << 32;
  return __double2ull_ru(p * 4294967296.0);  // exact product (power of two), rounded up = ceil
}

__device__ __forceinline__ double clamp01(double p) { return p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p); }

// Full 32-bit uniform of (row, global sim, stream): both chunks.  Used by the general kernels and the
// initial-state draws; the fast kernel evaluates the refinement lazily.
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t row, uint64_t sim, uint32_t stream) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint4 a = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 3), stream, k0, k1);
  const uint4 b = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 3), stream + 1u, k0, k1);
  const int k = (int)(sim & 7u);
  return (chunk16(a, k) << 16) | chunk16(b, k);
}

}  // namespace amro
