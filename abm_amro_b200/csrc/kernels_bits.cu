// BITS path: the ward-level colonisation step with the state BIT-SLICED over simulations -- summary runs (per-day counts,
// no trajectory) of hospitals with weights 1 and 1/2 and time_to_detect <= 6, free-running mode, in ONE persistent kernel.
//
// Replaces the same nest as kernels_fast.cu (src/amro/abm_wards.cpp:322-414 / :219-257 / :70-159 fused with the totals
// :435-509).  Where kernels_fast.cu keeps a 16-bit code per (record, simulation) and steps two simulations per 32-bit
// register, this kernel keeps ONE BIT PER PLANE: a 32-bit register holds one plane of 32 simulations of one record,
//   c    colonised                                  (reference: C = 1)
//   det  countdown at or below zero                 (D <= 0: detected; with c = 0 it is the initial-state quirk of :372,
//                                                     an uncolonised patient whose D reads 0)
//   q1   result due tomorrow  (D = 1)               (time_to_detect >= 1)
//   q2   result due in two days (D = 2)             (time_to_detect == 2)
// (time_to_detect 3 .. 6, the reference's default: planes det, q1, q2 hold a 3-bit code instead -- 0 nothing pending, 1 .. 6 days
// until the result, 7 detected; the instance compiled for TTD = 3 reads the actual value from the arguments)
// so a record x 32 simulations is one 16-byte word and every logical step of the update is one LOP3 for 32 simulations:
//   draws        the record's 14 plane words (rng.cuh) are compared with the THRESHOLD PLANES of its (ward-day, is_new):
//                lt = (t & ~u) | (~(t ^ u) & lt), least significant plane first -- one LOP3 per plane and probability
//                class, five classes (uncolonised / colonised / detected; tested positive: uncolonised / colonised)
//   transition   col = c ? (det ? fB : fA) : fU;  tested = col & hit & ~(det | q1 | q2);  det' = col & (det | q1);
//                q1' = col & q2;  q2' = tested      (:116-155: D -= 1 while a countdown runs, D = ttd on a positive test,
//                D = inf otherwise; c' = 0 clears everything)
//   counts       per-lane VERTICAL counters (bit-sliced ripple adders over the trips of a warp), turned into per-simulation
//                integers by a 32 x 32 bit transpose across the warp (five shuffle stages) + POPC when a ward-day / the day ends
// Data layout (HBM): the two-day state ring and the per-record meta words of kernels_fast.cu (same slots, same byte
// offsets: a slot holds kBW 16-byte words = kBW x 32 simulations), plus one bit per record "has a source" (src_mask).
// Ward pressure (the reference's sum(W), :99) is PULLED: the warp that builds a ward-day's tables counts the colonised
// plane of the ward-day's records in the state ring -- the states they start the day from, written by their predecessors
// yesterday -- with a vertical counter.  Nothing is added to a successor's ward when a record is stepped: no atomics, no
// transfer lists, no pressure buffers (kernels_fast.cu pushes; round 2 measured the push at a fifth of this kernel's
// instructions).  A ward-day that spans several runs is counted by each of them (config 2: 2.2 times).
// Work decomposition, team barrier and day plan: team.cuh, as in kernels_fast.cu; a team = 32 * kBW simulations.
#include <atomic>

#include "team.cuh"

namespace amro {
namespace {

constexpr int kBW = kSlicesPerTeam;                 // 16-byte words of a ring slot = groups of 32 simulations of a team
constexpr int kBitsSims = 32 * kBW;                 // simulations of a team
#ifndef AMRO_BITS_SEGS
#define AMRO_BITS_SEGS 4
#endif
constexpr int kBSegs = AMRO_BITS_SEGS;              // ring of ward-day tables per warp (power of two)
constexpr int kTabWords = 80;                       // a table row: 5 tables x 16 words (14 planes, the "always" mask, a zero)
constexpr int kTU = 0, kTA = 16, kTB = 32, kTUh = 48, kTAh = 64;   // word offsets of the tables in a row
// table rows of a warp: (ward-day slot, is_new) -- times the weight class (1 | 1/2) in the instances for hospitals with half
// weights -- plus one all-zero row: lanes without a record step against it
__host__ __device__ constexpr int zero_row(bool half) { return kBSegs * (half ? 4 : 2); }
constexpr int kCntPlanes = 5;                       // vertical counters count to 31
// (an instance for 6 CTAs per SM -- 80 registers, 100-900 B of spills -- was measured slower on every configuration and is gone)
constexpr int kFlushTrips = (1 << kCntPlanes) - 1;
static_assert(kDrawBits == 14, "table rows hold 14 planes + the always mask");

// 32 x 32 bit transpose across the warp: lane i holds row i; afterwards lane j holds column j (bit i = row i's bit j).
// Per stage the lanes of a pair exchange the blocks they do not keep (rotated into place before the shuffle).  K matrices
// at once: a stage's K shuffles are independent, so their latencies overlap (the per-day code is latency-bound: two
// warps per scheduler on config 2).
template <int K>
__device__ __forceinline__ void warp_transpose32_multi(uint32_t (&x)[K], int lane) {
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int k = 16 >> st;
    const uint32_t m = st == 0 ? 0x0000FFFFu : st == 1 ? 0x00FF00FFu : st == 2 ? 0x0F0F0F0Fu : st == 3 ? 0x33333333u : 0x55555555u;
    const bool upper = (lane & k) != 0;
    const int rot = upper ? k : 32 - k;
    const uint32_t keep = upper ? ~m : m;
    uint32_t y[K];
#pragma unroll
    for (int i = 0; i < K; ++i) y[i] = __shfl_xor_sync(0xFFFFFFFFu, __funnelshift_l(x[i], x[i], rot), k);
#pragma unroll
    for (int i = 0; i < K; ++i) x[i] = (x[i] & keep) | (y[i] & ~keep);
  }
}
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
  uint32_t t[1] = {x};
  warp_transpose32_multi<1>(t, lane);
  return t[0];
}

// Vertical counter: plane b holds bit b of a per-simulation count (32 simulations per register)
struct VCount {
  uint32_t v[kCntPlanes];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int b = 0; b < kCntPlanes; ++b) v[b] = 0u;
  }
  __device__ __forceinline__ void add(uint32_t x) {   // ripple-carry increment where x is set
#pragma unroll
    for (int b = 0; b < kCntPlanes; ++b) {
      const uint32_t t = v[b] & x;
      v[b] ^= x;
      x = t;
    }
  }
};
// lane s <- the warp's totals of simulation s of N counters after n_adds (warp-uniform) increments each; the counters are
// cleared.  All planes that can be non-zero are transposed together (one latency chain instead of one per plane).
template <int N, int NP>
__device__ __forceinline__ void vflush_planes(VCount* const (&vc)[N], uint32_t (&out)[N], int lane) {
  uint32_t t[N * NP];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int b = 0; b < NP; ++b) t[i * NP + b] = vc[i]->v[b];
  warp_transpose32_multi<N * NP>(t, lane);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    uint32_t c = 0u;
#pragma unroll
    for (int b = 0; b < NP; ++b) c += (uint32_t)__popc(t[i * NP + b]) << b;
    out[i] = c;
    vc[i]->clear();
  }
}
template <int N>
__device__ __forceinline__ void vflush(VCount* const (&vc)[N], uint32_t (&out)[N], int lane, uint32_t n_adds) {
  if (n_adds == 0u) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = 0u;
  } else if (n_adds < 16u) vflush_planes<N, kCntPlanes - 1>(vc, out, lane);
  else vflush_planes<N, kCntPlanes>(vc, out, lane);
}

// Shared memory (dynamic): per CTA the team's parameters and the run-constant tables, per warp a ring of table rows.
//   par   double [6][kBitsSims]                      alpha, beta, gamma, rho_hospital, alpha2, rho_imported (:79-84)
//   r16   uint32 [kBitsSims]                          the run's dither (rng.cuh: kRunDitherIndex)
//   rc    uint32 [is_new][word][48]                   tables A, B, Ah under a finite force of infection (they do not depend on the ward)
//   tab   uint32 [zero_row(HALF) + 1 rows][kTabWords] per warp: rows (ward-day slot, is_new[, weight class]) of the warp's word
//   prep  double [kBSegs][32] + uint32 [kBSegs][32]   per warp: beta / N and the dither of the next batch of ward-days (the part
//                                                     of a table that does not depend on the ward's pressure, computed while
//                                                     the warp would otherwise wait for the previous day to complete)
struct BitsShared {
  double* par;
  uint32_t* r16;
  uint32_t* rc;
  uint32_t* tab;
  double* prep_bn;
  uint32_t* prep_r16;
#ifdef AMRO_BITS_TMA
  uint4* tma;          // per warp: [2 stages][32 meta words | 32 state words]
  uint64_t* tma_bar;   // per warp: one mbarrier per stage
#endif
};
constexpr size_t kCtaSharedBytes = sizeof(double) * 6 * kBitsSims + sizeof(uint32_t) * kBitsSims + sizeof(uint32_t) * 2 * kBW * 48;
#ifdef AMRO_BITS_TMA   // the measured-and-rejected variant of the trip loop: inputs staged by TMA bulk copies
constexpr size_t kTmaBytes = 2 * 64 * sizeof(uint4) + 2 * sizeof(uint64_t);
#else
constexpr size_t kTmaBytes = 0;
#endif
__host__ __device__ constexpr size_t warp_shared_bytes(bool half) {
  return sizeof(uint32_t) * (zero_row(half) + 1) * kTabWords + (sizeof(double) + sizeof(uint32_t)) * kBSegs * 32 + kTmaBytes;
}
static_assert(kCtaSharedBytes % 16 == 0 && warp_shared_bytes(false) % 16 == 0 && warp_shared_bytes(true) % 16 == 0, "table rows are read as uint4");
__host__ __device__ inline size_t bits_smem_bytes(int threads, bool half) { return kCtaSharedBytes + (size_t)(threads / 32) * warp_shared_bytes(half); }
template <bool HALF>
__device__ __forceinline__ BitsShared carve_bits(unsigned char* raw, int warp) {
  BitsShared s;
  s.par = reinterpret_cast<double*>(raw);
  s.r16 = reinterpret_cast<uint32_t*>(raw + sizeof(double) * 6 * kBitsSims);
  s.rc = s.r16 + kBitsSims;
  unsigned char* mine = raw + kCtaSharedBytes + (size_t)warp * warp_shared_bytes(HALF);
  s.tab = reinterpret_cast<uint32_t*>(mine);
  s.prep_bn = reinterpret_cast<double*>(s.tab + (zero_row(HALF) + 1) * kTabWords);
  s.prep_r16 = reinterpret_cast<uint32_t*>(s.prep_bn + kBSegs * 32);
#ifdef AMRO_BITS_TMA
  s.tma = reinterpret_cast<uint4*>(s.prep_r16 + kBSegs * 32);
  s.tma_bar = reinterpret_cast<uint64_t*>(s.tma + 2 * 64);
#endif
  return s;
}

// What a warp needs about its team and its word: the warps of a team pair up -- a pair steps the same run of records,
// one warp per group of 32 simulations (word w of the 16-byte-word ring slots).
struct BitsCtx {
  uint4* pre;          // [2][ring_chunks][kBW][32], offset to the warp's word
  uint32_t* press;     // nullptr, or [n_days + 1][max_segs][kBW][32] offset to the warp's word: the pulled counts, kept for export
  uint32_t gsim0;      // global index of the first simulation of the warp's word (a multiple of 32)
  int64_t sim0;        // first simulation of the warp's word (run-relative)
  int w;               // the word
};

// element (uint4) index of slot q of a team's first word in a state-ring buffer laid out [chunk of 32 slots][word][32]
__device__ __forceinline__ int64_t ring_slot_b(uint32_t q) { return (int64_t)(q >> 5) * (kBW * 32) + (q & 31u); }

// 14-bit thresholds of two tables -> their plane words in a table row: after the transpose lane j holds plane j of the
// first table and lane 16 + j plane j of the second (bit 14 = "always": B = 2^14)
__device__ __forceinline__ void store_planes(uint32_t* row, int first, int second, uint32_t b_first, uint32_t b_second, int lane) {
  const uint32_t t = warp_transpose32(b_first | (b_second << 16), lane);
  if (lane < 16) row[first + lane] = t;
  else if (second >= 0) row[second + lane - 16] = t;
}

// ---- tables of the ward-days [b0, b1) (indices within day d) into the warp's ring; lane = simulation of the warp's word --
// Uncolonised patients feel the ward: P0 = F + gamma h with F = (beta / N) * sum(W) (:98-112), dithered per ward-day.  The
// colonised classes do not (W = 1: P = coef + 0 F + gamma h) unless F is not finite: their planes are copied from the
// run-constant tables, or built here for the simulations whose F is inf / NaN (N = 0).
// Two steps: prepare (beta / N and the dither: topology and parameters only -- done for a day's first batch BEFORE the warp
// waits for the previous day) and finish (the pressure is known).
__device__ __forceinline__ void prepare_tables_bits(const BitsShared& sh, const FastArgs& a, const BitsCtx& c, int64_t seg0, uint32_t b0, uint32_t b1) {
  const int lane = threadIdx.x & 31;
  const double beta = sh.par[kBitsSims + c.w * 32 + lane];
  const uint32_t gs = c.gsim0 + (uint32_t)lane;
  // two ward-days per pass (the division and the ten Philox rounds are two independent chains); a short tail is done twice
#pragma unroll 1
  for (uint32_t sl0 = b0; sl0 < b1; sl0 += 2) {
    double bn[2];
    uint32_t r16[2];
    int slot[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t sl = min(sl0 + (uint32_t)u, b1 - 1u);
      const int64_t seg = seg0 + sl;
      slot[u] = (int)(sl & (kBSegs - 1));
      bn[u] = __ddiv_rn(beta, __ldg(&a.seg_N[seg]));   // :98
      r16[u] = chunk16(philox4x32_10((uint32_t)seg, (uint32_t)((uint64_t)seg >> 32), gs >> 3, kStreamDither, (uint32_t)a.seed,
                                     (uint32_t)(a.seed >> 32)), (int)(gs & 7u));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      sh.prep_bn[slot[u] * 32 + lane] = bn[u];
      sh.prep_r16[slot[u] * 32 + lane] = r16[u];
    }
  }
}
// carry-save adder on bit planes: (hi, lo) = a + b + c per simulation
__device__ __forceinline__ void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t u = a ^ b;
  hi = (a & b) | (u & c);
  lo = u ^ c;
}
// four planes of weight one into a vertical counter: three carry-save adders and a three-plane ripple (11 LOP3 instead of
// the 40 of four increments); the counter must stay below 32
__device__ __forceinline__ void vadd4(VCount& vc, const uint32_t (&x)[4]) {
  static_assert(kCntPlanes == 5, "planes 0..1 take the carry-save sums, planes 2..4 the ripple");
  uint32_t t0, t1, f;
  csa(t0, vc.v[0], vc.v[0], x[0], x[1]);
  csa(t1, vc.v[0], vc.v[0], x[2], x[3]);
  csa(f, vc.v[1], vc.v[1], t0, t1);
  const uint32_t c2 = vc.v[2] & f;
  vc.v[2] ^= f;
  const uint32_t c3 = vc.v[3] & c2;
  vc.v[3] ^= c2;
  vc.v[4] ^= c3;
}

// sum(W) of ward-days (:95-99) for the 32 simulations of the warp's word: the colonised plane of the states the records
// [rs, re) (positions within the day) start the day from, counted with vertical counters -- lane = record while counting,
// lane = simulation in the result.  Records without a source (admitted today) have no state yet: their ring slot holds
// somebody else's old state and is masked out (src_mask: one bit per record, static).  Four 32-record chunks per step and
// ward-day through the carry-save adders.
constexpr int kPullGroup = 4;
__device__ __forceinline__ void pull_load(uint32_t (&x)[kPullGroup], const uint32_t* pre_x, const uint32_t* mask_day, uint32_t k, uint32_t k1,
                                          uint32_t rs, uint32_t re, int lane) {
  // no branches: chunks past the ward-day's end load its last chunk again and are masked out (pos >= re), so that all the
  // loads of a step issue back to back and the selects wait once
  uint32_t m[kPullGroup], v[kPullGroup];
#pragma unroll
  for (int j = 0; j < kPullGroup; ++j) {
    const uint32_t kk = min(k + (uint32_t)j, k1 - 1u);
    m[j] = __ldg(mask_day + kk);
    v[j] = __ldcg(pre_x + (int64_t)kk * (kBW * 32 * 4));
  }
#pragma unroll
  for (int j = 0; j < kPullGroup; ++j) {
    const uint32_t pos = (k + (uint32_t)j) * 32u + (uint32_t)lane;
    x[j] = pos >= rs && pos < re && ((m[j] >> lane) & 1u) ? v[j] : 0u;
  }
}
// A step of D groups for each of N ward-days: all the loads first (one round trip to L2), then the adders
template <int N, int D>
__device__ __forceinline__ void pull_step(VCount (&vc)[N], const uint32_t* pre_x, const uint32_t* mask_day, uint32_t (&k)[N], const uint32_t (&k1)[N],
                                          const uint32_t (&rows)[N + 1], int lane) {
  uint32_t x[D][N][kPullGroup];
#pragma unroll
  for (int q = 0; q < D; ++q)
#pragma unroll
    for (int u = 0; u < N; ++u) pull_load(x[q][u], pre_x, mask_day, k[u] + (uint32_t)(q * kPullGroup), k1[u], rows[u], rows[u + 1], lane);
#pragma unroll
  for (int q = 0; q < D; ++q)
#pragma unroll
    for (int u = 0; u < N; ++u) vadd4(vc[u], x[q][u]);
#pragma unroll
  for (int u = 0; u < N; ++u) k[u] += (uint32_t)(D * kPullGroup);
}
// N ward-days at once: their counters are flushed together (one chain of transposes).  Steps of up to kPullDepth groups per
// ward-day (12 chunks: a whole config-2 ward in one round trip), shorter ones for what is left.
#ifndef AMRO_BITS_PULL_DEPTH
#define AMRO_BITS_PULL_DEPTH 3
#endif
constexpr int kPullDepth = AMRO_BITS_PULL_DEPTH;
template <int N>
__device__ __forceinline__ void pull_pressure(const uint32_t* pre_x, const uint32_t* mask_day, const uint32_t (&rows)[N + 1], uint32_t (&cnt)[N], int lane) {
  VCount vc[N];
  uint32_t k[N], k1[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    vc[u].clear();
    cnt[u] = 0u;
    k[u] = rows[u] >> 5; k1[u] = max((rows[u + 1] + 31u) >> 5, k[u] + 1u);
  }
  VCount* vp[N];
#pragma unroll
  for (int u = 0; u < N; ++u) vp[u] = &vc[u];
  VCount* const (&vr)[N] = vp;
  uint32_t n_since = 0u;   // increments since the counters were flushed
  for (;;) {
    uint32_t rem = 0u;
#pragma unroll
    for (int u = 0; u < N; ++u) rem = max(rem, k1[u] > k[u] ? k1[u] - k[u] : 0u);
    if (rem == 0u) break;
    // (a lane's counter grows by one per chunk of its ward-day actually present: at most min(step, rem))
#if AMRO_BITS_PULL_DEPTH >= 4
    if (rem > 3u * kPullGroup) { pull_step<N, 4>(vc, pre_x, mask_day, k, k1, rows, lane); n_since += min(rem, 4u * kPullGroup); }
    else
#endif
    if (rem > 2u * kPullGroup) { pull_step<N, 3>(vc, pre_x, mask_day, k, k1, rows, lane); n_since += min(rem, 3u * kPullGroup); }
    else if (rem > (uint32_t)kPullGroup) { pull_step<N, 2>(vc, pre_x, mask_day, k, k1, rows, lane); n_since += rem; }
    else { pull_step<N, 1>(vc, pre_x, mask_day, k, k1, rows, lane); n_since += rem; }
    if (n_since + (uint32_t)(kPullDepth * kPullGroup) > (uint32_t)kFlushTrips) {
      uint32_t o[N];
      vflush<N>(vr, o, lane, n_since);
#pragma unroll
      for (int u = 0; u < N; ++u) cnt[u] += o[u];
      n_since = 0u;
    }
  }
  uint32_t o[N];
  vflush<N>(vr, o, lane, n_since);
#pragma unroll
  for (int u = 0; u < N; ++u) cnt[u] += o[u];
}

// The same count for launches that keep many warps per SM busy with short wards: there the SM has other warps to run while a
// load is on its way, and what counts is the number of instructions -- chunks past the ward-day's end are skipped instead of
// loaded and masked, one group per step (measured: profiles/sweeps.txt).  The host picks the variant per launch (pull_deep).
__device__ __forceinline__ void pull_load_lean(uint32_t (&x)[kPullGroup], const uint32_t* pre_x, const uint32_t* mask_day, uint32_t k, uint32_t k1,
                                               uint32_t rs, uint32_t re, int lane) {
#pragma unroll
  for (int j = 0; j < kPullGroup; ++j) {
    const uint32_t kk = k + (uint32_t)j, pos = kk * 32u + (uint32_t)lane;
    x[j] = 0u;
    if (kk < k1) {   // warp-uniform
      const uint32_t m = __ldg(mask_day + kk);
      const uint32_t v = __ldcg(pre_x + (int64_t)kk * (kBW * 32 * 4));   // loaded by every lane of the chunk (no divergent section), then masked
      x[j] = pos >= rs && pos < re && ((m >> lane) & 1u) ? v : 0u;
    }
  }
}
template <int N>
__device__ __forceinline__ void pull_pressure_lean(const uint32_t* pre_x, const uint32_t* mask_day, const uint32_t (&rows)[N + 1], uint32_t (&cnt)[N], int lane) {
  constexpr uint32_t kGroupsPerFlush = 7;
  VCount vc[N];
  uint32_t k[N], k1[N], groups = 0u;
#pragma unroll
  for (int u = 0; u < N; ++u) {
    vc[u].clear();
    cnt[u] = 0u;
    k[u] = rows[u] >> 5; k1[u] = (rows[u + 1] + 31u) >> 5;
    groups = max(groups, (k1[u] - k[u] + (uint32_t)kPullGroup - 1u) / (uint32_t)kPullGroup);
  }
  VCount* vp[N];
#pragma unroll
  for (int u = 0; u < N; ++u) vp[u] = &vc[u];
  VCount* const (&vr)[N] = vp;
  while (groups > 0u) {
    const uint32_t g = min(groups, kGroupsPerFlush);
    for (uint32_t i = 0; i < g; ++i) {
      uint32_t x[N][kPullGroup];
#pragma unroll
      for (int u = 0; u < N; ++u) pull_load_lean(x[u], pre_x, mask_day, k[u], k1[u], rows[u], rows[u + 1], lane);
#pragma unroll
      for (int u = 0; u < N; ++u) {
        vadd4(vc[u], x[u]);
        k[u] += (uint32_t)kPullGroup;
      }
    }
    groups -= g;
    uint32_t o[N];
    vflush<N>(vr, o, lane, g * (uint32_t)kPullGroup);
#pragma unroll
    for (int u = 0; u < N; ++u) cnt[u] += o[u];
  }
}

// What finish_tables_bits needs about the day
struct DayIn {
  const uint32_t* pre_x;      // colonised plane of the day's state-ring buffer: word .x of slot 0 of the warp's word, + 4 * lane
  const uint32_t* mask_day;   // src_mask of the day's first chunk (hospitals with half weights: the full-weight records only)
  const uint32_t* half_day;   // nullptr, or the same for the records that weigh 1/2
  const int64_t* seg_rows;    // seg_row_start of the day's first ward-day
  int64_t da;                 // first position of the day
  uint32_t* press_out;        // nullptr, or where the day's counts go ([ward-day of the day][kBW][32], + lane)
  bool test_h, test_a;
};
// deep: the launch keeps few warps per SM busy, or its wards are long (latency-bound counting: deep, branch-free pull steps)
template <bool HALF>
__device__ __forceinline__ void finish_tables_bits(const BitsShared& sh, const BitsCtx& c, const DayIn& in, uint32_t b0, uint32_t b1, bool deep) {
  const int lane = threadIdx.x & 31;
  const int sim = c.w * 32 + lane;
  const bool test_h = in.test_h, test_a = in.test_a;
  const double alpha = sh.par[sim], gamma = sh.par[2 * kBitsSims + sim], alpha2 = sh.par[4 * kBitsSims + sim];
  const double rho_h = sh.par[3 * kBitsSims + sim], rho_i = sh.par[5 * kBitsSims + sim];
  // two ward-days per pass: their dependent chains (fp64, transposes) overlap; a short tail is built twice
#pragma unroll 1
  for (uint32_t sl0 = b0; sl0 < b1; sl0 += 2) {
    uint32_t cnt[2], r16[2];
    double bn[2];
    int slot[2];
    uint32_t rows[3];
    if (lane < 3) rows[0] = (uint32_t)(__ldg(in.seg_rows + min(sl0 + (uint32_t)lane, b1)) - in.da);
    rows[2] = __shfl_sync(0xFFFFFFFFu, rows[0], 2); rows[1] = __shfl_sync(0xFFFFFFFFu, rows[0], 1); rows[0] = __shfl_sync(0xFFFFFFFFu, rows[0], 0);
    uint32_t cnth[2] = {0u, 0u};   // colonised records that weigh 1/2 (hospitals with patients split over two wards)
    auto pull = [&](const uint32_t* mask, uint32_t (&out)[2]) {
      if (sl0 + 1u < b1) {
        if (deep) pull_pressure<2>(in.pre_x, mask, rows, out, lane);
        else pull_pressure_lean<2>(in.pre_x, mask, rows, out, lane);
      } else {   // a single ward-day: its tables are built twice
        const uint32_t r1[2] = {rows[0], rows[1]};
        uint32_t c1[1];
        if (deep) pull_pressure<1>(in.pre_x, mask, r1, c1, lane);
        else pull_pressure_lean<1>(in.pre_x, mask, r1, c1, lane);
        out[0] = out[1] = c1[0];
      }
    };
    pull(in.mask_day, cnt);
    if (HALF) pull(in.half_day, cnth);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t sl = min(sl0 + (uint32_t)u, b1 - 1u);
      slot[u] = (int)(sl & (kBSegs - 1));
      // every run that touches the ward-day writes the same count (in halves when the hospital has half weights, like k_fast)
      if (in.press_out) in.press_out[(int64_t)sl * (kBW * 32)] = HALF ? 2u * cnt[u] + cnth[u] : cnt[u];
      bn[u] = sh.prep_bn[slot[u] * 32 + lane];
      r16[u] = sh.prep_r16[slot[u] * 32 + lane];
    }
    double F[2];
    bool finite[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      // :98-99, sum(W) = colonised full-weight records + half of the colonised half-weight ones (exact in any order)
      F[u] = __dmul_rn(bn[u], HALF ? __dadd_rn((double)cnt[u], __dmul_rn(0.5, (double)cnth[u])) : (double)cnt[u]);
      finite[u] = fabs(F[u]) <= 1.7976931348623157e308;
    }
    const bool all_finite = __all_sync(0xFFFFFFFFu, finite[0] && finite[1]);
    // the uncolonised classes of both ward-days and both is_new classes: four pairs of 14-bit thresholds, transposed together
    uint32_t tw[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool flag = h ? test_a : test_h;
      const double rho = h ? rho_i : rho_h;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const double P0 = class_probability(0, h, F[u], alpha, alpha2, gamma);
        const uint32_t bU = (threshold30(P0) + r16[u]) >> 16;
        const uint32_t bUh = flag ? (threshold30(__dmul_rn(clamp01(P0), rho)) + r16[u]) >> 16 : 0u;   // nobody is tested: never
        tw[h * 2 + u] = bU | (bUh << 16);
      }
    }
    warp_transpose32_multi<4>(tw, lane);   // lane j: plane j of U, lane 16 + j: plane j of Uh (bit 14 = "always": B = 2^14)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool flag = h ? test_a : test_h;
      const double rho = h ? rho_i : rho_h;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t* row = sh.tab + ((slot[u] * 2 + h) * (HALF ? 2 : 1)) * kTabWords;
        row[(lane < 16 ? kTU : kTUh - 16) + lane] = tw[h * 2 + u];
        if (all_finite) {
          const uint4* rc = reinterpret_cast<const uint4*>(sh.rc + (h * kBW + c.w) * 48);
          if (lane < 8) reinterpret_cast<uint4*>(row + kTA)[lane] = rc[lane];
          else if (lane < 12) reinterpret_cast<uint4*>(row + kTAh)[lane - 8] = flag ? rc[lane] : make_uint4(0u, 0u, 0u, 0u);
        } else {
          const uint32_t rd = finite[u] ? sh.r16[sim] : r16[u];   // rng.cuh: ward-free probabilities take the run's dither
          const double P1 = class_probability(1, h, F[u], alpha, alpha2, gamma), P2 = class_probability(2, h, F[u], alpha, alpha2, gamma);
          const uint32_t bA = (threshold30(P1) + rd) >> 16, bB = (threshold30(P2) + rd) >> 16;
          const uint32_t bAh = flag ? (threshold30(__dmul_rn(clamp01(P1), rho)) + rd) >> 16 : 0u;
          store_planes(row, kTA, kTB, bA, bB, lane);
          store_planes(row, kTAh, -1, bAh, 0u, lane);
        }
      }
    }
    if (HALF) {
      // rows of the records that weigh 1/2: uncolonised ones feel the ward like everybody (W = 0: the planes above), colonised
      // ones with W = 1/2 (P = coef / 2 + F / 2 + gamma h, :108-112): ward-dependent, dithered per ward-day
      uint32_t ta[4], tah[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool flag = h ? test_a : test_h;
        const double rho = h ? rho_i : rho_h;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const double P1 = class_probability(1, h, F[u], alpha, alpha2, gamma, 0.5), P2 = class_probability(2, h, F[u], alpha, alpha2, gamma, 0.5);
          const uint32_t bA = (threshold30(P1) + r16[u]) >> 16, bB = (threshold30(P2) + r16[u]) >> 16;
          ta[h * 2 + u] = bA | (bB << 16);
          tah[h * 2 + u] = flag ? (threshold30(__dmul_rn(clamp01(P1), rho)) + r16[u]) >> 16 : 0u;
        }
      }
      warp_transpose32_multi<4>(ta, lane);
      warp_transpose32_multi<4>(tah, lane);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t* row = sh.tab + ((slot[u] * 2 + h) * 2 + 1) * kTabWords;
          row[(lane < 16 ? kTU : kTUh - 16) + lane] = tw[h * 2 + u];
          row[(lane < 16 ? kTA : kTB - 16) + lane] = ta[h * 2 + u];
          if (lane < 16) row[kTAh + lane] = tah[h * 2 + u];
        }
    }
  }
}

#ifdef AMRO_BITS_TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// per-simulation day totals (:445-459 fused): lane s adds the count of simulation s of the warp's word
__device__ __forceinline__ void add_day_counts(const FastArgs& a, int64_t d, int64_t sim0, uint32_t cc, uint32_t cd) {
  const int lane = threadIdx.x & 31;
  if (d < a.counts_ld) {
    const int64_t at = d + a.counts_ld * (sim0 + lane);
    if (cc && a.counts_col) atomicAdd(&a.counts_col[at], (int)cc);
    if (cd && a.counts_det) atomicAdd(&a.counts_det[at], (int)cd);
  }
}

// What a warp keeps across the days of a run: d % tsh, d % tsa (:391-392; the schedules are 64-bit the way the reference's
// unsigned % reads them), kept without the division
struct DayState {
#ifdef AMRO_BITS_TMA
  uint32_t tma_stage, tma_phase;   // next stage of the warp's two-stage ring, expected parity of each stage's mbarrier
#endif
  uint64_t th, ta;
  __device__ __forceinline__ void next(const FastArgs& a) {
    th = th + 1u == a.tsh ? 0u : th + 1u;
    ta = ta + 1u == a.tsa ? 0u : ta + 1u;
  }
};
// ---- one day of the warp's word: the run of the day's records that belongs to this warp's pair (:384-407) ---------------
template <int TTD, bool HALF>
__device__ __forceinline__ void process_day_bits(const BitsShared& sh, const FastArgs& a, const BitsCtx& c, int d, const DayPlan& pl,
                                                 int64_t seg0_next, TeamBarrier& bar, DayState& ds) {
  const int lane = threadIdx.x & 31;
#ifdef AMRO_BITS_TMA
  uint32_t tma_stage = ds.tma_stage, tma_phase = ds.tma_phase;
  asm volatile("fence.proxy.async;" ::: "memory");   // yesterday's states (generic-proxy writes, acquired at the barrier) before the async-proxy reads
#endif
  VCount vcol, vdet;
  vcol.clear(); vdet.clear();
  uint32_t since = 0;   // trips since the day counters were flushed

  if (pl.w0 < pl.w1 && c.sim0 < a.S) {   // a word beyond the last simulation (a 32-member run: half a team) has nothing to step
    const int64_t ring = a.ring_chunks * (kBW * 32), da = pl.da, seg0 = pl.seg0;
#ifndef AMRO_BITS_NO_PREFETCH
    // Topology lines this warp will ask for one after the other when the day is over (tomorrow's plan: the ward-days at the ends
    // of its run, their N and row ranges -- tomorrow's run is today's when the census is steady): requested now, into L1,
    // without registers; they arrive while the trips run.  Wrong guesses cost nothing but the request.
    if (lane == 0 && d + 1 < (int)a.n_days) {
      const int64_t da_next = da + pl.n, sn = min(seg0_next + pl.sl_first, a.n_seg - 1);
      prefetch_l1(&a.meta[min(da_next + pl.w0, a.R - 1)]);
      prefetch_l1(&a.meta[min(da_next + pl.w1 - 1, a.R - 1)]);
      prefetch_l1(&a.seg_N[sn]);
      prefetch_l1(&a.seg_row_start[sn]);
    }
#endif
    const uint4* pday = c.pre + (d & 1) * ring;
    const uint4* pc = pday + (int64_t)(pl.w0 >> 5) * (kBW * 32) + lane;
    char* nxt = reinterpret_cast<char*>(c.pre + ((d + 1) & 1) * ring);
    const uint4* pm = a.meta + da + pl.w0 + lane;
    const uint32_t g32 = c.gsim0 >> 5;
    const uint32_t ttd_bit[3] = {(a.ttd & 1) ? ~0u : 0u, (a.ttd & 2) ? ~0u : 0u, (a.ttd & 4) ? ~0u : 0u};   // TTD > 2: the code a positive test starts
    DayIn in;
    in.pre_x = reinterpret_cast<const uint32_t*>(pday) + 4 * lane;
    in.mask_day = a.src_mask + __ldg(&a.day_chunk_start[d]);
    in.half_day = HALF ? a.half_mask + __ldg(&a.day_chunk_start[d]) : nullptr;
    in.seg_rows = a.seg_row_start + seg0;
    in.da = da;
    in.press_out = c.press ? c.press + (int64_t)d * a.max_segs_per_day * (kBW * 32) + lane : nullptr;
    in.test_h = ds.th == 0u; in.test_a = ds.ta == 0u;

    const uint32_t n_run = pl.w1 - pl.w0;
    for (uint32_t sl = pl.sl_first; sl <= pl.sl_last; sl += kBSegs) {
      const uint32_t n_tab = min((uint32_t)kBSegs, pl.sl_last + 1u - sl);
      __syncwarp();   // the previous batch's rows are no longer read
      if (sl != pl.sl_first) { prepare_tables_bits(sh, a, c, seg0, sl, sl + n_tab); __syncwarp(); }   // the first batch was prepared before the wait
#ifdef AMRO_EXP_NOTABLES   // timing experiment (wrong results): tables built on the first two days only
      if (d < 2)
#endif
      finish_tables_bits<HALF>(sh, c, in, sl, sl + n_tab, a.pull_deep != 0);
      __syncwarp();
      const uint32_t rb = max(pl.w0, (uint32_t)(__ldg(in.seg_rows + sl) - da)) - pl.w0;
      const uint32_t re = min(pl.w1, (uint32_t)(__ldg(in.seg_rows + sl + n_tab) - da)) - pl.w0;
      uint32_t tb = rb >> 5;
      const uint32_t t_end = (re + 31u) >> 5;
      while (tb < t_end) {
        const uint32_t te = min(t_end, tb + ((uint32_t)kFlushTrips - since));
        const uint4* pmt = pm + (int64_t)tb * 32;
        const uint4* pct = pc + (int64_t)tb * (kBW * 32);
#ifdef AMRO_BITS_TMA
        // rejected variant: the 512 B of record words and the 512 B of state words of a trip are staged in shared memory by
        // two TMA bulk copies, one trip ahead, completion on an mbarrier per stage (lane 0 issues, every lane waits)
        auto tma_issue = [&](uint32_t trip, uint32_t stg) {
          const uint32_t n_rec = min(32u, n_run - trip * 32u);
          mbar_expect(&sh.tma_bar[stg], n_rec * 16u + 512u);
          bulk_g2s(sh.tma + stg * 64, pm - lane + (int64_t)trip * 32, n_rec * 16u, &sh.tma_bar[stg]);
          bulk_g2s(sh.tma + stg * 64 + 32, pc - lane + (int64_t)trip * (kBW * 32), 512u, &sh.tma_bar[stg]);
        };
        __syncwarp();
        if (lane == 0 && tb * 32u < n_run) tma_issue(tb, tma_stage);
        for (uint32_t i = tb * 32u + lane; i < te * 32u + (uint32_t)lane; i += 32u) {   // warp-uniform trip count
          const uint32_t trip = i >> 5;
          __syncwarp();   // every lane has read the other stage (the previous trip's)
          if (lane == 0 && trip + 1u < te && (trip + 1u) * 32u < n_run) tma_issue(trip + 1u, tma_stage ^ 1u);
          uint4 meta = make_uint4(0u, kNone, kNone, 0u), pre = make_uint4(0u, 0u, 0u, 0u);
          if (trip * 32u < n_run) {
            mbar_wait(&sh.tma_bar[tma_stage], (tma_phase >> tma_stage) & 1u);
            tma_phase ^= 1u << tma_stage;
            if (i < n_run) meta = sh.tma[tma_stage * 64 + lane];
            pre = sh.tma[tma_stage * 64 + 32 + lane];
          }
          tma_stage ^= 1u;
#else
        // software pipeline: the loads of the next trip are in flight while this one computes
        uint4 n_meta = make_uint4(0u, kNone, kNone, 0u), n_prev = make_uint4(0u, 0u, 0u, 0u);
        if (tb * 32u + (uint32_t)lane < n_run) { n_meta = __ldg(pmt); n_prev = __ldcg(pct); }
        for (uint32_t i = tb * 32u + lane; i < te * 32u + (uint32_t)lane; i += 32u) {   // warp-uniform trip count
          const uint4 meta = n_meta;
          uint4 pre = n_prev;
          pmt += 32; pct += kBW * 32;
          n_meta = make_uint4(0u, kNone, kNone, 0u);
          if (i + 32u < min(n_run, te * 32u)) { n_meta = __ldg(pmt); n_prev = __ldcg(pct); }
#endif
          const uint32_t info = meta.x, segl = info & kInfoSegMask, src = (info >> kInfoSrcShift) & 3u;
          const int h = (int)(info >> kInfoHShift);
          const int64_t orig = a.identity_order ? da + pl.w0 + i : (int64_t)meta.w;
          const bool on = i < n_run && segl - sl < n_tab;   // this batch's records
          int row_i = (int)(segl & (kBSegs - 1)) * 2 + h;
          if (HALF) row_i = row_i * 2 + (int)((info >> kInfoHalfShift) & 1u);   // the record weighs 1/2
          const uint4* tr = reinterpret_cast<const uint4*>(sh.tab + (on ? row_i : zero_row(HALF)) * kTabWords);
          const uint32_t next = on ? meta.y : kNone;
          if (!on || src == kSrcNone) pre = make_uint4(0u, 0u, 0u, 0u);   // (C = 0, D = inf) x 32, :359-360

          // ---- draws: u < t per probability class, least significant plane first; four planes per Philox call (rng.cuh) ----
          uint32_t lt[5] = {0u, 0u, 0u, 0u, 0u}, always[5];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint4 u = philox_rk(a, (uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), g32, kStreamPlanes + (uint32_t)q4);
            const uint32_t up[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 5; ++t) {
              const uint4 th = tr[t * 4 + q4];
              const uint32_t tp[4] = {th.x, th.y, th.z, th.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (q4 * 4 + j < kDrawBits) lt[t] = (tp[j] & ~up[j]) | (~(tp[j] ^ up[j]) & lt[t]);
                else if (q4 * 4 + j == kDrawBits) always[t] = tp[j];
              }
            }
          }
          const uint32_t fU = lt[0] | always[0], fA = lt[1] | always[1], fB = lt[2] | always[2];
          const uint32_t fUh = lt[3] | always[3], fAh = lt[4] | always[4];
          // ---- transition (:116-155) ----
          uint4 post;
          uint32_t col, det_post;
          if (TTD <= 2) {
            const uint32_t cc = pre.x, det = pre.y, q1 = pre.z, q2 = pre.w;
            col = (cc & ((det & fB) | (~det & fA))) | (~cc & fU);                         // :116-120
            const uint32_t hit = (cc & fAh) | (~cc & fUh);                                // tested positive today (:134 / :146)
            const uint32_t tested = col & hit & ~(det | q1 | q2);                         // no countdown running: D = ttd
            post.x = col;
            post.y = col & (det | q1);                                                    // D -= 1 reaches or stays below 0 (:131-132)
            post.z = 0u; post.w = 0u;
            if (TTD == 0) post.y |= tested;
            if (TTD == 1) post.z = tested;
            if (TTD == 2) { post.z = col & q2; post.w = tested; }
            det_post = post.y;
          } else {
            // time_to_detect 3 .. 6: planes y, z, w are a 3-bit code n per simulation -- 0 no result pending, 1 .. 6 days until
            // the result (the reference's countdown D), 7 detected (D <= 0; with c = 0 the initial-state quirk)
            const uint32_t cc = pre.x, n0 = pre.y, n1 = pre.z, n2 = pre.w;
            const uint32_t det = n0 & n1 & n2, any = n0 | n1 | n2;
            col = (cc & ((det & fB) | (~det & fA))) | (~cc & fU);
            const uint32_t hit = (cc & fAh) | (~cc & fUh);
            const uint32_t tested = col & hit & ~any;                                     // no countdown running: D = ttd (:134 / :146)
            const uint32_t one = n0 & ~n1 & ~n2;                                          // D = 1: the result arrives (D -= 1 reaches 0)
            const uint32_t to7 = col & (det | one);                                       // (one implies a running countdown)
            const uint32_t dec = col & any & ~det & ~one;                                 // D -= 1 (:131-132), still positive
            const uint32_t m1 = ~(n1 ^ n0), m2 = n2 ^ (~n0 & ~n1);                        // n - 1 (bit 0: ~n0)
            post.x = col;
            post.y = to7 | (dec & ~n0) | (tested & ttd_bit[0]);
            post.z = to7 | (dec & m1) | (tested & ttd_bit[1]);
            post.w = to7 | (dec & m2) | (tested & ttd_bit[2]);
            det_post = post.y & post.z & post.w;
          }
          // :402-407 the successor starts from this state -- and its ward counts it tomorrow (pull_pressure)
          if (next != kNone) __stcg(reinterpret_cast<uint4*>(nxt + next), post);
          vcol.add(col);
          vdet.add(det_post);
        }
        since += te - tb;
        tb = te;
        if (since == (uint32_t)kFlushTrips) {   // the counters are full
          VCount* const vc[2] = {&vcol, &vdet};
          uint32_t v[2];
          vflush_planes<2, kCntPlanes>(vc, v, lane);
          add_day_counts(a, d, c.sim0, v[0], v[1]);
          since = 0;
        }
      }
    }
  }
  // The transposes of the day totals (registers and shuffles only) run while the last state stores are on their way to L2 --
  // the release below waits for those anyway.  The totals' atomics are nobody's input: after the arrive.
  VCount* const vc[2] = {&vcol, &vdet};
  uint32_t v[2];
  vflush<2>(vc, v, lane, since);
  bar.arrive();   // tomorrow's states are out
  add_day_counts(a, d, c.sim0, v[0], v[1]);
#ifdef AMRO_BITS_TMA
  ds.tma_stage = tma_stage; ds.tma_phase = tma_phase;
#endif
}

template <int TTD, int THREADS, int MINB, bool HALF>
__global__ void __launch_bounds__(THREADS, MINB) k_bits(const __grid_constant__ FastArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kWarps = THREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BitsShared sh = carve_bits<HALF>(smem_raw, warp);
  const int big = a.teams_plus_one * (a.ctas_per_team + 1);
  const bool in_big = (int)blockIdx.x < big;
  const int G = a.ctas_per_team + (in_big ? 1 : 0);
  const int64_t team = in_big ? blockIdx.x / G : a.teams_plus_one + ((int)blockIdx.x - big) / G;
  const int rank = in_big ? blockIdx.x % G : ((int)blockIdx.x - big) % G;
  // the warps of a team pair up: pair `gw` of `n_gw` steps a run of the day's records, warp `w` of the pair its word.  A team
  // whose second word holds no simulation (the last team of a run, or the only one: 32 members per GPU) does not pair: every
  // warp takes its own run of word 0 (twice as many, half as long).
  static_assert(kWarps % kBW == 0 && kBW == 2, "whole pairs per CTA");
  const uint32_t words = (a.S - team * kBitsSims > 32) ? (uint32_t)kBW : 1u;
  const uint32_t gw_all = (uint32_t)(rank * kWarps + warp), gw = gw_all / words, n_gw = (uint32_t)(G * kWarps) / words;
  TeamBarrier bar{a.team_barrier + team, 0u, (unsigned)(G * kWarps), G == 1 ? kBarCta : (a.cluster_barrier ? kBarCluster : kBarGlobal)};
  const int64_t ring = a.ring_chunks * (kBW * 32);

  uint4* team_pre = a.pre + team * 2 * ring;
  // the pulled counts of every ward-day are kept only when the caller asked for them (pressure_out): [team][n_days][max_segs][kBW][32]
  uint32_t* team_press = a.pressure ? a.pressure + team * a.press_days * a.max_segs_per_day * (kBW * 32) : nullptr;
  const uint32_t team_gsim0 = (uint32_t)(a.sim_offset + (uint64_t)team * kBitsSims);
  const int64_t team_sim0 = team * kBitsSims;
  BitsCtx c;
  c.w = (int)(gw_all % words);
  c.pre = team_pre + c.w * 32;
  c.press = team_press ? team_press + c.w * 32 : nullptr;
  c.gsim0 = team_gsim0 + (uint32_t)c.w * 32u;
  c.sim0 = team_sim0 + c.w * 32;

  // ---- team setup: parameters, run-constant tables and the all-zero rows to shared memory --------------------------------
  for (int q = tid; q < 6 * kBitsSims; q += THREADS) sh.par[q] = a.params[(team_sim0 + q % kBitsSims) + a.p_ld * (q / kBitsSims)];
  for (int q = lane; q < kTabWords; q += 32) sh.tab[zero_row(HALF) * kTabWords + q] = 0u;
  __syncthreads();
  for (int it = warp; it < 2 * kBW; it += kWarps) {   // (is_new, word): lane = simulation
    const int h = it / kBW, w = it % kBW, sim = w * 32 + lane;
    const double alpha = sh.par[sim], gamma = sh.par[2 * kBitsSims + sim], alpha2 = sh.par[4 * kBitsSims + sim];
    const double rho = sh.par[(h ? 5 : 3) * kBitsSims + sim];
    const uint32_t gs = team_gsim0 + (uint32_t)sim;
    const uint32_t rd = chunk16(philox4x32_10((uint32_t)kRunDitherIndex, (uint32_t)(kRunDitherIndex >> 32), gs >> 3, kStreamDither,
                                              (uint32_t)a.seed, (uint32_t)(a.seed >> 32)), (int)(gs & 7u));
    if (h == 0) sh.r16[sim] = rd;
    const double P1 = class_probability(1, h, 0.0, alpha, alpha2, gamma), P2 = class_probability(2, h, 0.0, alpha, alpha2, gamma);
    uint32_t* rc = sh.rc + (h * kBW + w) * 48;
    store_planes(rc, 0, 16, (threshold30(P1) + rd) >> 16, (threshold30(P2) + rd) >> 16, lane);
    store_planes(rc, 32, -1, (threshold30(__dmul_rn(clamp01(P1), rho)) + rd) >> 16, 0u, lane);
  }
  __syncthreads();

  // ---- initial state (:363-372) into the day-0 buffer: a lane draws 8 simulations of a row, four lanes make a word -----
  // (the ward pressure of day 0 is pulled from these states like any other day's)
  {
    const int64_t n = a.n_init, per = ((n + G - 1) / G + 7) / 8 * 8;
    const int64_t lo = rank * per, hi = min(n, lo + per);
    for (int64_t r0 = lo + warp * 8; r0 < hi; r0 += kWarps * 8) {
      const int64_t r = r0 + (lane >> 2);
      const int g = lane & 3;
      const bool live = r < hi;
      const uint32_t slot = live ? __ldg(&a.init_slot[r]) : kNone;   // kNone: overwritten by a predecessor before it is ever read
#pragma unroll 1
      for (int w = 0; w < kBW; ++w) {
        uint32_t c0 = 0u, d0 = 0u, valid = 0u;
        if (slot != kNone) {
          const uint32_t r_lo = (uint32_t)r, r_hi = (uint32_t)((uint64_t)r >> 32), gsl = ((team_gsim0 + (uint32_t)w * 32u) >> 3) + (uint32_t)g;
          const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
          const uint4 rc_hi = philox4x32_10(r_lo, r_hi, gsl, kStreamInitCol, k0, k1), rc_lo = philox4x32_10(r_lo, r_hi, gsl, kStreamInitCol + 1u, k0, k1);
          const uint4 rd_hi = philox4x32_10(r_lo, r_hi, gsl, kStreamInitDet, k0, k1), rd_lo = philox4x32_10(r_lo, r_hi, gsl, kStreamInitDet + 1u, k0, k1);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int64_t sim = team_sim0 + w * 32 + g * 8 + k;
            if (sim < a.S) {
              const double pcol = a.icp[r * a.icp_rs + sim * a.icp_cs], pdet = a.idp[r * a.idp_rs + sim * a.idp_cs];
              const bool cb = (uint64_t)((chunk16(rc_hi, k) << 16) | chunk16(rc_lo, k)) < threshold33(pcol);
              const bool db = (uint64_t)((chunk16(rd_hi, k) << 16) | chunk16(rd_lo, k)) < threshold33(pdet);
              c0 |= (cb ? 1u : 0u) << (g * 8 + k);
              d0 |= (db ? 1u : 0u) << (g * 8 + k);
              valid |= 1u << (g * 8 + k);
            }
          }
        }
        c0 |= __shfl_xor_sync(0xFFFFFFFFu, c0, 1); c0 |= __shfl_xor_sync(0xFFFFFFFFu, c0, 2);
        d0 |= __shfl_xor_sync(0xFFFFFFFFu, d0, 1); d0 |= __shfl_xor_sync(0xFFFFFFFFu, d0, 2);
        valid |= __shfl_xor_sync(0xFFFFFFFFu, valid, 1); valid |= __shfl_xor_sync(0xFFFFFFFFu, valid, 2);
        if (slot != kNone) {
          // :372  D0 = d0*c0 is a 0/1 FLAG that the dynamics read as a countdown on day 0: uncolonised -> D = 0 (det without
          // c); colonised, flag 0 -> D = 0: detected; flag 1 -> D = 1: due tomorrow if ttd >= 1, else neither pending nor
          // detected (it is re-tested like D = inf)
          const uint32_t flag1 = c0 & d0;
          if (g == 0) {
            if (TTD <= 2) __stcg(&team_pre[ring_slot_b(slot) + w * 32], make_uint4(c0, valid & ~flag1, TTD >= 1 ? flag1 : 0u, 0u));
            else __stcg(&team_pre[ring_slot_b(slot) + w * 32], make_uint4(c0, valid, valid & ~flag1, valid & ~flag1));   // code 7, or 1 where D0 = 1
          }
        }
      }
    }
  }
  bar.arrive();

  // ---- day loop (:381): one team barrier per day ------------------------------------------------------------------
  DayPlan plan = plan_day(a, 0, gw, n_gw, DayPlan{});
#ifdef AMRO_BITS_TMA
  DayState ds{0u, 0u, 0u, 0u};
  if (lane == 0) { mbar_init(&sh.tma_bar[0]); mbar_init(&sh.tma_bar[1]); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
#else
  DayState ds{0u, 0u};
#endif
  const int n_days = (int)a.n_days;
  for (int d = 0; d < n_days; ++d) {
    const int64_t seg0_next = __ldg(&a.day_seg_start[d + 1]);
    // what does not depend on the previous day's results: the first batch's beta / N and dithers
    if (plan.w0 < plan.w1 && c.sim0 < a.S) prepare_tables_bits(sh, a, c, plan.seg0, plan.sl_first, min(plan.sl_last + 1u, plan.sl_first + (uint32_t)kBSegs));
#ifdef AMRO_EXP_NOWAIT   // timing experiment (wrong results)
    if (d < 2)
#endif
    bar.wait();
    process_day_bits<TTD, HALF>(sh, a, c, d, plan, seg0_next, bar, ds);
    ds.next(a);
    plan = plan_day(a, d + 1, gw, n_gw, plan);   // static: ready before the next wait ends
  }
  if (bar.mode == kBarCluster) bar.wait();   // every arrive of a cluster barrier is paired with a wait
}

// src_mask: bit j of word day_chunk_start[d] + k <=> record 32 k + j of day d has a source (a predecessor or an initial row).
// Hospitals with half weights get two masks: src_mask = the sourced records that weigh 1, half_mask = those that weigh 1/2.
// One warp per chunk (grid.y = day).
__global__ void k_bits_src_mask(const uint4* __restrict__ meta, const int64_t* __restrict__ day_start, const int64_t* __restrict__ day_chunk_start,
                                int64_t n_days, uint32_t* __restrict__ src_mask, uint32_t* __restrict__ half_mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_x = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = blockIdx.y; d < n_days; d += gridDim.y) {
    const int64_t da = day_start[d], n = day_start[d + 1] - da, n_chunks = (n + 31) / 32;
    for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < n_chunks; k += warps_x) {
      const int64_t i = k * 32 + lane;
      const uint32_t info = i < n ? __ldg(&meta[da + i].x) : 0u;
      const bool has = i < n && ((info >> kInfoSrcShift) & 3u) != kSrcNone;
      const bool half = has && half_mask != nullptr && ((info >> kInfoHalfShift) & 1u);
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, has && !half), mh = __ballot_sync(0xFFFFFFFFu, half);
      if (lane == 0) {
        src_mask[day_chunk_start[d] + k] = m;
        if (half_mask) half_mask[day_chunk_start[d] + k] = mh;
      }
    }
  }
}

// ward pressure of a finished run -> out[ward-day + ld * simulation] (press_days = n_days + 1)
__global__ void k_bits_export_pressure(const uint32_t* __restrict__ pressure, const int64_t* __restrict__ day_seg_start, int64_t n_days,
                                       int64_t max_segs, int64_t press_days, int64_t S, int32_t* __restrict__ out, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_days * max_segs * S) return;
  const int64_t s = i % S, sl = (i / S) % max_segs, d = i / (S * max_segs);
  const int64_t seg = day_seg_start[d] + sl;
  if (seg >= day_seg_start[d + 1]) return;
  const int64_t team = s / kBitsSims, w = (s % kBitsSims) / 32, lane = s % 32;
  out[seg + ld * s] = (int32_t)pressure[((team * press_days + d) * max_segs + sl) * (kBW * 32) + w * 32 + lane];
}

template <int TTD, bool HALF>
const void* bits_kernel_of(int shape) {
  return shape == kShapeRoomy ? (const void*)k_bits<TTD, kTeamThreads, kTeamMinBlocks - 1, HALF>
                              : (const void*)k_bits<TTD, kTeamThreads, kTeamMinBlocks, HALF>;
}
template <bool HALF>
const void* bits_kernel_half(int ttd, int shape) {
  return ttd == 0 ? bits_kernel_of<0, HALF>(shape) : ttd == 1 ? bits_kernel_of<1, HALF>(shape) : ttd == 2 ? bits_kernel_of<2, HALF>(shape)
                                                                                                    : bits_kernel_of<3, HALF>(shape);
}
const void* bits_kernel_ptr(int ttd, int shape, bool half) { return half ? bits_kernel_half<true>(ttd, shape) : bits_kernel_half<false>(ttd, shape); }

}  // namespace

int bits_sims_per_team() { return kBitsSims; }
int bits_words_per_team() { return kBW; }
size_t bits_shared_bytes(bool half) { return bits_smem_bytes(kTeamThreads, half); }

// Per (device, instance): the dynamic shared memory opt-in and the occupancy of the instance are asked for once, not on every
// run (a parameter sweep calls amro_simulate thousands of times).
namespace {
constexpr int kMaxCachedDevices = 16;
std::atomic<int> g_bits_cap[kMaxCachedDevices][4][4][2];        // co-resident CTAs + 1 (0 = not asked yet)
inline int ttd_index(int ttd) { return ttd < 0 ? 0 : (ttd > 3 ? 3 : ttd); }
}  // namespace
int bits_max_coresident_ctas(int device, int ttd, int shape, bool half) {
  std::atomic<int>* slot = (device >= 0 && device < kMaxCachedDevices && shape >= 0 && shape < 4) ? &g_bits_cap[device][ttd_index(ttd)][shape][half ? 1 : 0] : nullptr;
  if (slot) { const int v = slot->load(std::memory_order_acquire); if (v > 0) return v - 1; }
  int per_sm = 0, sms = 0;
  const void* k = bits_kernel_ptr(ttd, shape, half);
  const size_t smem = bits_smem_bytes(kTeamThreads, half);
  AMRO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AMRO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kTeamThreads, smem));
  if (slot) slot->store(per_sm * sms + 1, std::memory_order_release);
  return per_sm * sms;
}

int bits_max_coresident_clusters(int ttd, int shape, int cluster_size, bool half) {
  const void* k = bits_kernel_ptr(ttd, shape, half);
  const size_t smem = bits_smem_bytes(kTeamThreads, half);
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cluster_size > 8 && cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)cluster_size * 1024u);
  cfg.blockDim = dim3((unsigned)kTeamThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void bits_launch(cudaStream_t st, const FastArgs& a, int shape, int grid) {
  const void* k = bits_kernel_ptr(a.ttd, shape, a.half_mask != nullptr);
  const size_t smem = bits_smem_bytes(kTeamThreads, a.half_mask != nullptr);
  {
    int device = -1;
    AMRO_CUDA(cudaGetDevice(&device));
    const bool known = device >= 0 && device < kMaxCachedDevices && shape >= 0 && shape < 4 &&
                       g_bits_cap[device][ttd_index(a.ttd)][shape][a.half_mask != nullptr ? 1 : 0].load(std::memory_order_acquire) > 0;
    if (!known) AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // else: set when the occupancy was asked for
  }
  void* params[] = {(void*)&a};
  if (a.cluster_barrier) {
    if (a.ctas_per_team > 8) AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)kTeamThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.ctas_per_team; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AMRO_CUDA(cudaLaunchKernelExC(&cfg, k, params));
  } else if (a.ctas_per_team > 1 || a.teams_plus_one > 0) {
    AMRO_CUDA(cudaLaunchCooperativeKernel(k, dim3((unsigned)grid), dim3((unsigned)kTeamThreads), params, smem, st));
  } else {
    AMRO_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3((unsigned)kTeamThreads), params, smem, st));
  }
}

void bits_build_src_mask(cudaStream_t st, const uint4* meta, const int64_t* day_start, const int64_t* day_chunk_start, int64_t n_days,
                         int64_t max_rows_per_day, uint32_t* src_mask, uint32_t* half_mask) {
  if (n_days <= 0) return;
  const int64_t chunks = (max_rows_per_day + 31) / 32;
  const unsigned gx = (unsigned)std::min<int64_t>(std::max<int64_t>((chunks + 7) / 8, 1), 4096);
  const unsigned gy = (unsigned)std::min<int64_t>(n_days, 65535);
  k_bits_src_mask<<<dim3(gx, gy), 256, 0, st>>>(meta, day_start, day_chunk_start, n_days, src_mask, half_mask);
  AMRO_CUDA(cudaGetLastError());
}

void bits_export_pressure(cudaStream_t st, const uint32_t* pressure, const int64_t* day_seg_start, int64_t n_days, int64_t max_segs,
                          int64_t press_days, int64_t S, int32_t* out, int64_t ld) {
  const int64_t n = n_days * max_segs * S;
  if (n <= 0) return;
  k_bits_export_pressure<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pressure, day_seg_start, n_days, max_segs, press_days, S, out, ld);
}

}  // namespace amro
