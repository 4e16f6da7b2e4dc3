// Counter-based RNG of the free-running mode (DESIGN.md "RNG spec") and the probability -> threshold map.
//
// Philox4x32 (Salmon, Moraes, Dror, Shaw, SC'11), hand-written; key = 64-bit seed.  The bulk stream -- the plane words of the
// daily draws, four calls per record and 32 simulations -- runs 7 rounds (Philox4x32-7: the smallest round count the authors
// report as passing BigCrush; 10 is their default with a safety margin); the dither and initial-state calls run 10.
//
// Daily step: the uniforms of a patient-day record are BIT-SLICED over groups of 32 consecutive simulations.  Plane word
// j (0..13) of (record, group g = global sim >> 5) is word (j & 3) of the call with counter (row_lo, row_hi, g, 8 + (j >> 2));
// bit (sim & 31) of plane j is bit j of the 14-bit draw u14(record, sim).  The bit-sliced kernel (kernels_bits.cu)
// compares the planes with threshold planes directly -- one LOP3 per plane and 32 simulations; the other kernels gather
// their simulations' bits.
//     Bernoulli(p) fires  <=>  u14 < B(p),      B(p) = (T30(p) + r16) >> 16  in [0, 2^14],
// with T30(p) = ceil(p * 2^30) (0 for p <= 0 or NaN, 2^30 for p >= 1) and r16 a DITHER: the 16-bit chunk of
// (stream 6, index, sim) of a call with counter (index_lo, index_hi, sim >> 3, 6), where index = the record's ward-day, or
// the all-ones index (one dither per simulation and run) when the probability does not depend on the ward (a record with
// W = C*w = 1 under a finite force of infection).  Rounding the 30-bit threshold to 14 bits with a uniform dither keeps
// the probability of every draw exact to 2^-30 (E[B(p)] = T30(p) / 2^16; "never" and "always" stay exact).  Draws that
// share a dither are coupled only through it: the count of new colonisations of a ward-day of n records is over-dispersed
// by at most n * 2^-14 relative to independent draws, and only where p < 2^-14.
// Initial state (streams 2..5, index = initial row): one call with counter (row_lo, row_hi, sim >> 3, stream) serves 8
// simulations; 32-bit uniform (chunk << 16 | chunk of stream+1) < T(p), T(p) = #{r : r * 2^-32 < p}.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace amro {

constexpr uint32_t kStreamInitCol = 2, kStreamInitDet = 4, kStreamDither = 6, kStreamPlanes = 8;  // init: +1 = low half; planes: 8..11
constexpr int kDrawBits = 14;                              // planes of the daily draw
constexpr uint64_t kRunDitherIndex = ~0ull;                // dither index of ward-independent probabilities

constexpr int kPlaneRounds = 7;                            // Philox rounds of the plane words (the dither / initial-state calls: 10)

template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
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

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  return philox4x32<10>(c0, c1, c2, c3, k0, k1);
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

// T30(p) and the dithered 14-bit threshold B(p) of the daily step
__device__ __forceinline__ uint32_t threshold30(double p) {
  if (!(p > 0.0)) return 0u;
  if (p >= 1.0) return 1u << 30;
  return (uint32_t)__double2ull_ru(p * 1073741824.0);  // exact product (power of two), rounded up = ceil
}
__device__ __forceinline__ uint32_t bernoulli14(double p, uint32_t r16) { return (threshold30(p) + r16) >> 16; }

__device__ __forceinline__ double clamp01(double p) { return p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p); }

// Full 32-bit uniform of (row, global sim, stream): both chunks (initial-state draws).
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint64_t row, uint64_t sim, uint32_t stream) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint4 a = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 3), stream, k0, k1);
  const uint4 b = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 3), stream + 1u, k0, k1);
  const int k = (int)(sim & 7u);
  return (chunk16(a, k) << 16) | chunk16(b, k);
}

// 16-bit chunk of (index, global sim, stream): the dither (stream 6)
__device__ __forceinline__ uint32_t philox_chunk(uint64_t seed, uint64_t index, uint64_t sim, uint32_t stream) {
  const uint4 a = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), (uint32_t)(sim >> 3), stream, (uint32_t)seed,
                                (uint32_t)(seed >> 32));
  return chunk16(a, (int)(sim & 7u));
}

// 14-bit daily draw of (row, global sim): its bit of each plane word (the gather the non-bit-sliced kernels do)
__device__ __forceinline__ uint32_t philox_u14(uint64_t seed, uint64_t row, uint64_t sim) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), b = (uint32_t)(sim & 31u);
  uint32_t u = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 a = philox4x32<kPlaneRounds>((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 5), kStreamPlanes + (uint32_t)c, k0, k1);
    u |= ((a.x >> b) & 1u) << (4 * c);
    u |= ((a.y >> b) & 1u) << (4 * c + 1);
    if (c < 3) { u |= ((a.z >> b) & 1u) << (4 * c + 2); u |= ((a.w >> b) & 1u) << (4 * c + 3); }
  }
  return u;
}

}  // namespace amro
