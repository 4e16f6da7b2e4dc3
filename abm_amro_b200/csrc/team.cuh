// Device helpers shared by the two fused kernels (kernels_fast.cu: 16-bit state codes; kernels_bits.cu: bit-sliced state):
// the team barrier, the split of a day's records over the warps of a team, the reference's transition probability, Philox
// with round keys from the kernel parameters, relaxed global REDs.
#pragma once
#include "common.hpp"
#include "kernels.hpp"
#include "rng.cuh"
#include "topology.hpp"

namespace amro {
namespace {

__device__ __forceinline__ void red_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// two adjacent packed words with one 64-bit RED (p is 8-byte aligned; no 16-bit field ever overflows, so nothing carries
// from the low word into the high one)
__device__ __forceinline__ void red_add2(uint32_t* p, uint32_t lo, uint32_t hi) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"((unsigned long long)lo | ((unsigned long long)hi << 32)) : "memory");
}

// Transition probability for unit weight (W = C), force of infection F -- the reference's expression order, one
// IEEE operation at a time (:95-112).  pc: 0 = uncolonised, 1 = colonised undetected, 2 = colonised detected.
__device__ __forceinline__ double class_probability(int pc, int h, double F, double alpha, double alpha2, double gamma, double w = 1.0) {
  const double W = pc ? w : 0.0;   // :95  W = C * w
  const double det = (pc == 2) ? 1.0 : 0.0;
  const double coef = __dadd_rn(__dmul_rn(__dsub_rn(1.0, det), __dsub_rn(1.0, alpha)),
                                __dmul_rn(det, __dsub_rn(1.0, alpha2)));
  double P = __dadd_rn(__dmul_rn(coef, W), __dmul_rn(__dsub_rn(1.0, W), F));
  P = __dadd_rn(P, __dmul_rn(h ? 1.0 : 0.0, gamma));
  return P;
}


// The plane words' Philox4x32 (kPlaneRounds rounds, rng.cuh) with the round keys taken from the kernel parameters (constant
// bank operands)
__device__ __forceinline__ uint4 philox_rk(const FastArgs& a, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < kPlaneRounds; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a.rk[2 * r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ a.rk[2 * r + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}


// ---- team barrier ------------------------------------------------------------------------------------------------
// The warps that share a team's slices (all warps of the team's CTAs) meet once per day: a warp may start day d+1 when
// every warp of the team has finished day d (its successors' states and tomorrow's pressure are then complete).  Arrival
// is a release increment of the team's counter by lane 0 after __syncwarp (which orders the other lanes' writes before
// it); waiting is a relaxed-load poll by lane 0 closed by one acquire load, then __syncwarp.  Nothing in the day loop synchronises a whole
// CTA, so the latency of one warp's barrier or table build is covered by the other warps of the SM.  A team of one CTA
// uses the hardware barrier instead.  The cooperative launch guarantees co-residency.
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Three mechanisms, same protocol (arrive after the day's last write, wait before the next day's first read):
//   kBarCta      a team of one CTA: __syncthreads
//   kBarCluster  the team is a thread-block cluster: barrier.cluster arrive.release / wait.acquire (hardware, split-phase)
//   kBarGlobal   any team size: a counter in global memory (release RED by lane 0 of each warp, relaxed poll + acquire load)
constexpr int kBarCta = 0, kBarGlobal = 1, kBarCluster = 2;
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
struct TeamBarrier {
  unsigned* ctr;
  unsigned target;
  unsigned n_warps;   // warps of the team
  int mode;
  __device__ __forceinline__ void arrive() {
    if (mode == kBarCta) return;
    __syncwarp();
    if (mode == kBarCluster) {
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      return;
    }
    target += n_warps;
    if ((threadIdx.x & 31) == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  }
  __device__ __forceinline__ void wait() {
    if (mode == kBarCta) { __syncthreads(); return; }
    if (mode == kBarCluster) {
      __syncwarp();
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      return;
    }
    if ((threadIdx.x & 31) == 0) {
#ifdef AMRO_POLL_ACQUIRE   // round 1 / first half of round 2: every poll an acquire load (LDG + CCTL.IVALL)
      while ((int)(ld_acquire(ctr) - target) < 0) {}
#else
      // relaxed polls (no L1 invalidate per poll: the SM's other warps keep their cached topology lines), then ONE acquire
      // load, which synchronises with the releases it observes (acquire load = load + L1 invalidate, without the MEMBAR of a fence)
      while ((int)(ld_relaxed(ctr) - target) < 0) {}
      (void)ld_acquire(ctr);
#endif
    }
    __syncwarp();
  }
};


// What a warp does on one day: its run [w0, w1) of the day's records (day-relative) and where they sit.  Static
// (topology only), so it is computed BEFORE the warp waits for the previous day to complete.
struct DayPlan {
  int64_t da;          // first position of the day
  int64_t seg0;        // first ward-day of the day
  uint32_t w0, w1;     // the run; w0 is a multiple of 32
  uint32_t sl_first, sl_last;   // ward-days (within the day) the run touches
  uint32_t n;          // records of the day
};
// chunks [floor(n_chunks * gw / n_gw), floor(n_chunks * (gw + 1) / n_gw)) without 64-bit division: with n_chunks = q n_gw + r,
// floor(n_chunks * g / n_gw) = q g + floor(r g / n_gw), and r g < n_gw^2 fits 32 bits for every team this library launches
__device__ __forceinline__ void split_chunks(uint32_t n_chunks, uint32_t gw, uint32_t n_gw, uint32_t& c0, uint32_t& c1) {
  if (n_gw < 65536u) {
    const uint32_t q = n_chunks / n_gw, r = n_chunks - q * n_gw;
    c0 = q * gw + (r * gw) / n_gw;
    c1 = q * (gw + 1u) + (r * (gw + 1u)) / n_gw;
  } else {
    c0 = (uint32_t)(((uint64_t)n_chunks * gw) / n_gw);
    c1 = (uint32_t)(((uint64_t)n_chunks * (gw + 1)) / n_gw);
  }
}
// `prev`: the warp's plan of the day before (or DayPlan{}): a day with as many records splits the same way
__device__ __forceinline__ DayPlan plan_day(const FastArgs& a, int64_t d, uint32_t gw, uint32_t n_gw, const DayPlan& prev) {
  DayPlan p{};
  if (d >= a.n_days) return p;
  p.da = __ldg(&a.day_start[d]);
  p.n = (uint32_t)(__ldg(&a.day_start[d + 1]) - p.da);
  p.seg0 = __ldg(&a.day_seg_start[d]);
  // the warp's run of 32-record chunks: consecutive trips stay in one ward for a while, so what the records add to
  // tomorrow's pressure is kept in registers across trips
  if (p.n == prev.n) { p.w0 = prev.w0; p.w1 = prev.w1; }
  else {
    uint32_t c0, c1;
    split_chunks((p.n + 31u) / 32u, gw, n_gw, c0, c1);
    p.w0 = min(p.n, 32u * c0);
    p.w1 = min(p.n, 32u * c1);
  }
  if (p.w0 < p.w1) {
    p.sl_first = __ldg(&a.meta[p.da + p.w0].x) & kInfoSegMask;
    p.sl_last = __ldg(&a.meta[p.da + p.w1 - 1].x) & kInfoSegMask;
  }
  return p;
}


}  // namespace
}  // namespace amro
