// FAST path: the ward-level colonisation step for a whole run in ONE persistent cooperative kernel.
//
// Replaces the nest  simulate_discrete_model_internal_one -> progress_patients_1_timestep ->
// progress_patients_probability_ward_1_timestep  (src/amro/abm_wards.cpp:322-414 / :219-257 / :70-159) fused
// with total_positive_colonized/_detected (:435-509), for hospitals with weights in {1, 1/2}, is_new in {0,1},
// time_to_detect >= 0 and next_day links that connect consecutive days.
//
// State of one (record, simulation): a 16-bit code.
//   0x0000        U   uncolonised                              (reference: C=0, D=inf)
//   0x0800        U0  uncolonised with a zero countdown        (C=0, D=0: the initial-state quirk of :372)
//   0x1000        A   colonised, no test result coming         (C=1, D=inf)
//   0x4000 - D        colonised with a finite countdown D: pending (0x20C0..0x3FFF) while 0 < D <= ttd, detected
//                     (0x4000..0x5F40) once D <= 0.  The reference's daily `D -= 1` (:131-132) is `code + 1`.
//   Every code is the bit pattern of a finite non-negative fp16 number and integer order = fp16 order, so class tests
//   (code >= boundary) and the Bernoulli compares run as packed half-precision compares (HSET2: one instruction per TWO
//   simulations, on the FMA pipe, result a half-word mask or 1.0 per half-word), and the per-lane counters are packed
//   half adds.  Two simulations share a 32-bit word and are stepped together (SIMD within a register); U and A have
//   one code each (U -> A | 0x4000-ttd, A -> A | 0x4000-ttd), a running countdown is `code + 1`, U0 (first day only)
//   steps like D = 0 with the probability of an uncolonised patient.
//
// Data layout (HBM):
//   ring      [team][2][chunk of 32 records][slice][32] one 16-byte word = 8 simulations of one record
//             (fast_ring_slot); buffer d&1 holds the state the records of day d start from.  A record's step
//             WRITES its result into its successor's slot of buffer (d+1)&1 (next_day is a scatter, :402-407),
//             so reading the pre-step state needs no dependent load.  Records without a predecessor ignore the
//             slot and start uncolonised (:359-360).
//   pressure  [team][ward-day][slice][4] packed 16-bit counts (in halves when the hospital has weight-1/2 records):
//             colonised weight a ward-day STARTS with = the reference's sum(W) (:99), accumulated by the previous
//             day's pass (every record adds its new C to the ward-day of its successor): one pass over a day's
//             records instead of two.
//   meta      one uint4 per record shared by all simulations: (is_new | source kind | weight bits | ward-day index
//             inside its day), successor's byte offset in the ring, successor's ward-day, original row.
// Work decomposition: a TEAM of ctas_per_team CTAs owns kSlicesPerTeam slices (8 simulations each) for every day; a
// lane holds one record x 16 simulations, lanes map to consecutive records, so state loads/stores are 128-bit and
// coalesced.  Teams never talk to each other; inside a team the warps split each day's records into contiguous runs
// and meet at one barrier per day (__syncthreads, the cluster barrier, or a release/acquire counter in global memory).
// Per ward-day, per simulation, the transition probabilities a record can have (class x is_new x weight) are evaluated
// in fp64 exactly as the reference evaluates them and turned into dithered 14-bit thresholds (rng.cuh) that every warp
// builds into its own shared-memory table ring.
#include "common.hpp"
#include "kernels.hpp"
#include "rng.cuh"
#include "topology.hpp"
#include "team.cuh"

#define AMRO_PRAGMA_STR(x) _Pragma(#x)
#define AMRO_PRAGMA_UNROLL(n) AMRO_PRAGMA_STR(unroll n)
namespace amro {
#ifdef AMRO_PROFILE_PHASES
__device__ unsigned long long g_phase_cycles[8];   // development only: summed cycles of thread 0 of every CTA per phase
#define PHASE_T0() long long _pt = clock64()
#define PHASE_DECL long long& _pt,
#define PHASE_ARG _pt,
#define PHASE_ADD(i) do { if (threadIdx.x == 0) { long long _n = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(_n - _pt)); _pt = _n; } } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_DECL
#define PHASE_ARG
#define PHASE_ADD(i) do {} while (0)
#endif
namespace {

constexpr uint32_t kClsU = 0, kClsA = 1, kClsU0 = 2, kClsPn = 3, kClsB = 4;
constexpr uint32_t kCodeU0 = 0x0800u, kCodeA = 0x1000u, kCodeCnt = 0x2000u, kCodeZero = 0x4000u;  // class boundaries
static_assert(kCodeZero - kMaxCountdown > kCodeCnt && kCodeZero + kMaxCountdown < 0x7C00u, "codes must stay finite fp16 patterns");
__host__ __device__ inline uint32_t cls_of(uint32_t code) {
  return code >= kCodeZero ? kClsB : code >= kCodeCnt ? kClsPn : code >= kCodeA ? kClsA : code >= kCodeU0 ? kClsU0 : kClsU;
}
constexpr int kTabRow = 5;                      // 16-byte entries per (ward-day, is_new): see SharedView
constexpr int kFlushEvery = 1024;              // row iterations between flushes of the 16-bit packed per-lane counters
constexpr uint32_t kHalves = 0x00010001u;      // one in each half-word
constexpr uint32_t kCodeA2 = kCodeA * kHalves, kCodeZero2 = kCodeZero * kHalves, kCodeCnt2 = kCodeCnt * kHalves, kCodeU02 = kCodeU0 * kHalves;
constexpr uint32_t kOne2 = 0x3C003C00u;        // half2 (1.0, 1.0)
constexpr uint32_t kTwo2 = 0x40004000u;        // half2 (2.0, 2.0)

// packed half-precision compares / adds on 32-bit registers holding two fp16 patterns
__device__ __forceinline__ uint32_t h2_ge_mask(uint32_t x, uint32_t y) {  // 0xFFFF per half-word where x >= y
  uint32_t m;
  asm("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(y));
  return m;
}
__device__ __forceinline__ uint32_t h2_lt_mask(uint32_t x, uint32_t y) {
  uint32_t m;
  asm("set.lt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(y));
  return m;
}
__device__ __forceinline__ uint32_t h2_ge_one(uint32_t x, uint32_t y) {   // 1.0 per half-word where x >= y
  uint32_t m;
  asm("set.ge.f16x2.f16x2 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(y));
  return m;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t x, uint32_t y) {
  uint32_t r;
  asm("add.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
  return r;
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t x, uint32_t y, uint32_t z) {
  uint32_t r;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z));
  return r;
}
// two small non-negative integers held as halves -> packed 16-bit integers
__device__ __forceinline__ uint32_t h2_to_u16x2(uint32_t x) {
  uint32_t lo, hi;
  asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.rni.u32.f16 %0, l; cvt.rni.u32.f16 %1, h; }" : "=r"(lo), "=r"(hi) : "r"(x));
  return lo | (hi << 16);
}

struct SliceParams {  // the 8 simulations' parameter rows (abm_wards.cpp:79-84)
  double alpha[8], beta[8], gamma[8], rho_h[8], alpha2[8], rho_i[8];
  // 30-bit thresholds that do not depend on the ward-day (finite force of infection), [kind][is_new][simulation]:
  // kind 0: stays colonised, undetected (1 - alpha + gamma h); 1: stays colonised, detected (1 - alpha2 + gamma h);
  // 2: kind 0 and tested positive (clamp(P) * rho)
  uint32_t t30[3][2][8];
  uint32_t r16_run[8];   // the run's dither of those thresholds (rng.cuh: kRunDitherIndex)
};

// probability class of a state class: U, U0 -> 0; A, pending -> 1; detected -> 2
__device__ __forceinline__ int prob_class(uint32_t cls) { return cls == kClsB ? 2 : (cls == kClsA || cls == kClsPn) ? 1 : 0; }

// code delta of a colonised outcome: (class, tested positive today) -> new code - old code  (mod 2^16)   [replay mode]
__device__ __forceinline__ uint32_t code_delta(uint32_t cls, bool hit, int ttd) {
  const uint32_t newdet = kCodeZero - (uint32_t)ttd;  // D = ttd after the step (:135 / :147)
  if (cls == kClsU) return hit ? newdet : kCodeA;      // D = inf (:138 / :150)
  if (cls == kClsA) return (hit ? newdet : kCodeA) - kCodeA;
  if (cls == kClsU0) return (kCodeZero + 1u) - kCodeU0;  // D = 0 <= ttd: D -= 1 (:131-132 / :143-144)
  return 1u;                                            // pending / detected: D -= 1
}

// Shared memory (dynamic): the slice parameters, then for EACH WARP the tables of up to kWarpTableSegs ward-days.
//   free-running: tab  uint4 at [((ward-day*4 + is_new*2 + half weight)*kNS + slice)*kTabRow + e]; word j of an entry belongs
//                      to simulations 2j (low half) and 2j+1 (high half); 14-bit Bernoulli thresholds b (rng.cuh):
//                        e=0  bU               uncolonised:        colonised today  <=>  v14 < bU
//                        e=1  bU ^ bA          A / pending
//                        e=2  bA ^ bB          detected
//                        e=3  bUh              uncolonised, tested positive today (0 when nobody is tested)
//                        e=4  bUh ^ bAh        A
//   replay:       P    double at [(((ward-day*4 + is_new*2 + half weight)*kNS + slice)*3 + class)*8 + sim]: the transition probabilities themselves
constexpr int kWarpTableSegs = 8;              // ring of ward-day tables per warp (power of two)
constexpr int kRowsPerSeg = 4;                 // (is_new, weight class): weight 1 / weight 1/2
constexpr int kZeroRow = kWarpTableSegs * kRowsPerSeg;   // + one all-zero row: lanes without a record step against it
constexpr int kNS = kSlicesPerTeam;   // slices (of 8 simulations) a team steps together: a thread carries one record x 16 simulations
// element (uint4) index of slot q of the team's first slice in a state-ring buffer laid out [chunk of 32 slots][slice][32]
__host__ __device__ inline int64_t ring_slot(uint32_t q) { return (int64_t)(q >> 5) * (kNS * 32) + (q & 31u); }
struct WarpShared {
  uint4* tab;
  double* P;
};
__host__ __device__ inline size_t warp_table_bytes(bool replay) {
  return (size_t)(kWarpTableSegs * kRowsPerSeg + 1) * kNS * (replay ? 24 * sizeof(double) : kTabRow * sizeof(uint4));
}
__host__ __device__ inline size_t fast_smem_bytes(bool replay, int threads) {
  return kNS * sizeof(SliceParams) + (size_t)(threads / 32) * warp_table_bytes(replay);
}
template <bool REPLAY>
__device__ __forceinline__ WarpShared carve_shared(unsigned char* raw, int warp) {
  WarpShared v{};
  unsigned char* mine = raw + (size_t)warp * warp_table_bytes(REPLAY);
  if constexpr (REPLAY) v.P = reinterpret_cast<double*>(mine);
  else v.tab = reinterpret_cast<uint4*>(mine);
  return v;
}

// The 14-bit daily draws of one record's 16 simulations (a team = one half of a group of 32 simulations): the group's
// plane words (rng.cuh: 4 calls, 14 planes), this half of each as the rows of a 16 x 16 bit matrix held two rows per
// register, transposed in registers (three block-swap stages between registers, one inside each).  v[i] = the draws of
// simulations 2i (low half-word) and 2i + 1 (high) of the team.
__device__ __forceinline__ void draws_u14x16(const FastArgs& a, int64_t orig, uint32_t g32, uint32_t half, uint32_t (&v)[8]) {
  uint4 q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = philox_rk(a, (uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), g32, kStreamPlanes + (uint32_t)c);
  const uint32_t sel = half ? 0x7632u : 0x5410u;
  v[0] = __byte_perm(q[0].x, q[0].y, sel); v[1] = __byte_perm(q[0].z, q[0].w, sel);
  v[2] = __byte_perm(q[1].x, q[1].y, sel); v[3] = __byte_perm(q[1].z, q[1].w, sel);
  v[4] = __byte_perm(q[2].x, q[2].y, sel); v[5] = __byte_perm(q[2].z, q[2].w, sel);
  v[6] = __byte_perm(q[3].x, q[3].y, sel); v[7] = 0u;   // planes 14 and 15 do not exist (kDrawBits = 14)
  static_assert(kDrawBits == 14, "the transpose below drops planes 14 and 15");
#pragma unroll
  for (int st = 0; st < 3; ++st) {
    const int k = 8 >> st, kk = k >> 1;
    const uint32_t m = (st == 0 ? 0x00FFu : st == 1 ? 0x0F0Fu : 0x3333u) * kHalves;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i & kk) continue;
      const uint32_t t = ((v[i] >> k) ^ v[i + kk]) & m;
      v[i + kk] ^= t;
      v[i] ^= t << k;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t t = ((v[i] >> 1) ^ (v[i] >> 16)) & 0x5555u;
    v[i] ^= (t << 16) ^ (t << 1);
  }
}

// One record, eight simulations, free-running mode; two simulations per 32-bit word, stepped together.
//   class masks   detected / running countdown / colonised: packed half compares of the code with the class boundaries
//   thresholds    b1 = bU ^ (colonised & (bU^bA)) ^ (detected & (bA^bB)),  b2 = bUh ^ (colonised & (bUh^bAh))
//   draws         colonised today = v14 < b1, tested positive = v14 < b2 (one uniform for both, so
//                 P(colonised and positive) = P * rho exactly)
//   transition    countdown: code + 1;  U / A: A, or 0x4000 - ttd when tested positive;  not colonised: 0 (:153-155)
// FIRST: the record's state may hold U0 (uncolonised, D = 0; day of the initial state only): probability of an uncolonised
// patient, transition of a countdown at 0.
template <bool FIRST>
__device__ __forceinline__ uint4 step_record(const uint4* __restrict__ tab, const uint4 pre, const uint4 rnd, uint32_t x2) {
  const uint4 tU = tab[0], tXA = tab[1], tXB = tab[2], tUh = tab[3], tXAh = tab[4];
  const uint32_t prew[4] = {pre.x, pre.y, pre.z, pre.w}, rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
  const uint32_t bU[4] = {tU.x, tU.y, tU.z, tU.w}, xA[4] = {tXA.x, tXA.y, tXA.z, tXA.w}, xB[4] = {tXB.x, tXB.y, tXB.z, tXB.w};
  const uint32_t bH[4] = {tUh.x, tUh.y, tUh.z, tUh.w}, xH[4] = {tXAh.x, tXAh.y, tXAh.z, tXAh.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w = prew[j];
    const uint32_t v = rw[j];   // two 14-bit draws
    const uint32_t mdet = h2_ge_mask(w, kCodeZero2);
    uint32_t mcnt = h2_ge_mask(w, kCodeCnt2);
    const uint32_t mcol = h2_ge_mask(w, kCodeA2);
    if constexpr (FIRST) {
      const uint32_t mu0 = h2_ge_mask(w, kCodeU02) & ~mcol;
      w = (w & ~mu0) | (kCodeZero2 & mu0);
      mcnt |= mu0;
    }
    const uint32_t b1 = bU[j] ^ (mcol & xA[j]) ^ (mdet & xB[j]);
    const uint32_t b2 = bH[j] ^ (mcol & xH[j]);
    const uint32_t m1 = h2_lt_mask(v, b1), m2 = h2_lt_mask(v, b2);
    const uint32_t t = m2 & x2 & ~mcnt;                                 // U / A tested positive: A -> 0x4000 - ttd
    const uint32_t q = (mcnt & (w + kHalves)) | (~mcnt & kCodeA2);      // D -= 1 (:131-132)  |  A
    o[j] = m1 & (q ^ t);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// One record, eight simulations, replay mode: the caller's uniforms, compared in fp64 as the reference does.
__device__ __forceinline__ uint4 step_record_replay(const WarpShared& sh, const SliceParams& par, const FastArgs& a, const uint4 pre, int row, int h,
                                                    int64_t orig, int64_t sim0, bool flag, bool on) {
  const uint32_t prew[4] = {pre.x, pre.y, pre.z, pre.w};
  uint32_t postw[4] = {0, 0, 0, 0};
  const double* rho_tab = h ? par.rho_i : par.rho_h;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t v = (prew[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    const uint32_t cls = cls_of(v);
    const int64_t sim = sim0 + k;
    const double P = sh.P[(row * 3 + prob_class(cls)) * 8 + k];
    bool col = false, hit = false;
    if (on && sim < a.S) {
      if (a.p_out) a.p_out[orig + a.R * sim] = P;
      col = a.u_col[orig + a.u_ld * sim] < P;                                                  // :116
      // a finite countdown is always <= ttd here (ttd >= 0), so only U and A reach the test (:131 / :134)
      if (col && cls <= kClsA && flag) hit = a.u_det[orig + a.u_ld * sim] <= rho_tab[k];       // :134 / :146
    }
    const uint32_t o = col ? ((v + code_delta(cls, hit, a.ttd)) & 0xFFFFu) : 0u;               // :153-155
    postw[k >> 1] |= o << (16 * (k & 1));
  }
  return make_uint4(postw[0], postw[1], postw[2], postw[3]);
}

}  // namespace

// Everything a warp needs about its team (kNS adjacent slices of 8 simulations; a lane carries one record x 8*kNS sims).
struct TeamCtx {
  uint4* pre;          // [2][ring_chunks][kNS][32]
  uint4* traj;         // [R][kNS] or nullptr
  uint32_t* press;     // [n_seg][kNS][4]
  uint32_t gslice0;    // global index of the team's first slice (counter word 2 of the dither / initial-state calls)
  uint32_t g32, half;  // the team's 16 simulations: half `half` of the group of 32 simulations g32 (daily draws, rng.cuh)
  int64_t sim0;        // first simulation of the team (run-relative)
};

// ---- tables of the ward-days [b0, b1) (indices within day d), built by ONE warp into its ring of table rows ----------
// lane -> (slice, ward-day, simulation): two ward-days per pass, both slices and both is_new classes at once.

// The probabilities of colonised patients (:103-112 with W = 1) do not depend on the ward's pressure unless the force of
// infection is not finite (N = 0): their 30-bit thresholds are evaluated once per run (SliceParams::t30) and only the
// ward-day's dither is added here; a non-finite force takes the general expression.
template <bool REPLAY, bool LEAN>
__device__ __forceinline__ void build_tables(WarpShared sh, const SliceParams* par, const FastArgs& a, const uint32_t* press, uint32_t gslice0,
                                          int64_t seg0, uint32_t b0, uint32_t b1, bool test_h, bool test_a) {
  const int lane = threadIdx.x & 31;
  // lane -> (slice, ward-day of the pass, simulation); each lane builds the rows of both is_new classes of its ward-day.
  // kNS == 2: two ward-days per pass; kNS == 1: four
  const int s = kNS == 2 ? (lane >> 4) : 0, sub = kNS == 2 ? ((lane >> 3) & 1) : (lane >> 3), k = lane & 7;
  constexpr uint32_t kPerPass = kNS == 2 ? 2u : 4u;
  const SliceParams& p = par[s];
  const double beta = p.beta[k], alpha = p.alpha[k], alpha2 = p.alpha2[k], gamma = p.gamma[k];
  // the loads of the next pass are in flight while this one computes
  auto lane_seg = [&](uint32_t sl0) { return min(sl0 + (uint32_t)sub, b1 - 1u); };   // a short tail is simply built more than once
  uint32_t pk_next = 0u;
  double n_next = 1.0;
  if (b0 < b1) {
    const int64_t seg = seg0 + lane_seg(b0);
    pk_next = __ldcg(&press[(seg * kNS + s) * 4 + (k >> 1)]);
    n_next = __ldg(&a.seg_N[seg]);
  }
#pragma unroll 1
  for (uint32_t sl0 = b0; sl0 < b1; sl0 += kPerPass) {
    const uint32_t sl = lane_seg(sl0);
    const int64_t seg = seg0 + sl;
    const int slot = (int)(sl & (kWarpTableSegs - 1));
    const uint32_t pk = pk_next;
    const double n_seg = n_next;
    if (sl0 + kPerPass < b1) {
      const int64_t sn = seg0 + lane_seg(sl0 + kPerPass);
      pk_next = __ldcg(&press[(sn * kNS + s) * 4 + (k >> 1)]);
      n_next = __ldg(&a.seg_N[sn]);
    }
    const uint32_t cnt = (k & 1) ? (pk >> 16) : (pk & 0xFFFFu);
    // :98-99  F = (beta / N) * sum(W).  sum(W) is exact: the count of colonised records, or, when some records weigh 1/2,
    // the count in halves times 0.5 (sums of halves are exact in float64 in any order, Armadillo's two accumulators included)
    const double sumW = (!LEAN && a.has_half) ? __dmul_rn((double)cnt, 0.5) : (double)cnt;
    const double F = __dmul_rn(__ddiv_rn(beta, n_seg), sumW);
    const bool finite = fabs(F) <= 1.7976931348623157e308;
    const uint32_t r16 = REPLAY ? 0u : chunk16(philox4x32_10((uint32_t)seg, (uint32_t)((uint64_t)seg >> 32), gslice0 + (uint32_t)s, kStreamDither,
                                                            (uint32_t)a.seed, (uint32_t)(a.seed >> 32)), k);
#pragma unroll 1
    for (int wc = 0; wc <= ((!LEAN && a.has_half) ? 1 : 0); ++wc) {   // weight class of the record: 1, then 1/2
      const double w = wc ? 0.5 : 1.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {   // hospitalised, then arrivals: two independent chains
        const bool flag = h ? test_a : test_h;
        const double rho = h ? p.rho_i[k] : p.rho_h[k];
        const int trow = ((slot * kRowsPerSeg + h * 2 + wc) * kNS + s);
        if constexpr (REPLAY) {
#pragma unroll
          for (int pc = 0; pc < 3; ++pc) sh.P[(trow * 3 + pc) * 8 + k] = class_probability(pc, h, F, alpha, alpha2, gamma, w);
        } else {
          // dithered 14-bit thresholds (rng.cuh); U and A are tested today if it is a testing day, a running countdown is
          // never re-tested (:131 / :134)
          const double P0 = class_probability(0, h, F, alpha, alpha2, gamma, w);
          uint32_t t30[5];
          t30[0] = threshold30(P0);
          t30[3] = threshold30(__dmul_rn(clamp01(P0), rho));
          if (finite && wc == 0) {
            t30[1] = p.t30[0][h][k]; t30[2] = p.t30[1][h][k]; t30[4] = p.t30[2][h][k];
          } else {   // a colonised half-weight record feels the ward's pressure too: P = coef/2 + F/2 (:108-109)
            const double P1 = class_probability(1, h, F, alpha, alpha2, gamma, w);
            t30[1] = threshold30(P1);
            t30[2] = threshold30(class_probability(2, h, F, alpha, alpha2, gamma, w));
            t30[4] = threshold30(__dmul_rn(clamp01(P1), rho));
          }
          // ward-free thresholds take the run's dither, the others the ward-day's (rng.cuh)
          const uint32_t r16c = (finite && wc == 0) ? p.r16_run[k] : r16;
          uint32_t e[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) e[q] = (t30[q] + ((q == 0 || q == 3) ? r16 : r16c)) >> 16;
          if (!flag) e[3] = e[4] = 0u;
          // pair up with the neighbouring simulation (lane ^ 1): word j = sims (2j, 2j+1)
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, e[q], 1);
            e[q] = (k & 1) ? (other | (e[q] << 16)) : (e[q] | (other << 16));
          }
          if (!(k & 1)) {
            uint32_t* row = reinterpret_cast<uint32_t*>(sh.tab + trow * kTabRow) + (k >> 1);
            row[0] = e[0]; row[4] = e[0] ^ e[1]; row[8] = e[1] ^ e[2]; row[12] = e[3]; row[16] = e[3] ^ e[4];
          }
        }
      }
    }
  }
}

// Per-day totals (abm_wards.cpp:445-459 fused), out of line (once per warp and day): per-lane halves -> packed 16-bit
// integers -> transposing butterfly over the warp (each step halves the registers and doubles the lanes summed) -> one RED
// per lane = (slice, metric, simulation).  cnt[q] / cnt[kNS*4 + q]: colonised / detected counts of state word q.
__device__ __forceinline__ void flush_day_counts(const FastArgs& a, int64_t d, int64_t sim0, const uint32_t (&cnt)[kNS * 8]) {
  const int lane = threadIdx.x & 31;
  constexpr int kRegs = kNS * 8;   // 8 or 16 registers of two packed counts
  static_assert(kRegs == 8 || kRegs == 16, "butterfly sizes");
  uint32_t r[kRegs];
#pragma unroll
  for (int q = 0; q < kNS * 4; ++q) {     // per-lane counts are at most kFlushEvery: exact as halves, no carry between halves
    r[(q >> 2) * 8 + (q & 3)] = h2_to_u16x2(cnt[q]);
    r[(q >> 2) * 8 + 4 + (q & 3)] = h2_to_u16x2(cnt[kNS * 4 + q]);
  }
#pragma unroll
  for (int w = kRegs / 2, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int q = 0; q < w; ++q) {
      const uint32_t send = up ? r[q] : r[q + w], keep = up ? r[q + w] : r[q];
      r[q] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, bit);
    }
  }
#pragma unroll
  for (int bit = 16 / kRegs; bit >= 1; bit >>= 1) r[0] += __shfl_xor_sync(0xFFFFFFFFu, r[0], bit);
  // the register a lane ends up with: t = slice(0/1 bit) | metric(1) | word(2); its low / high half = simulations 2*word, 2*word+1
  const int t = lane >> (kRegs == 16 ? 1 : 2);
  const int sl = t >> 3, metric = (t >> 2) & 1, k = (t & 3) * 2 + (lane & 1);
  const uint32_t v = (lane & 1) ? (r[0] >> 16) : (r[0] & 0xFFFFu);
  int32_t* dst = metric ? a.counts_det : a.counts_col;
  const bool emit = kRegs == 16 || !(lane & 2);
  if (emit && v != 0 && dst && d < a.counts_ld) atomicAdd(&dst[d + a.counts_ld * (sim0 + sl * 8 + k)], (int)v);
}

// What the warp's records add to the ward-day most of them feed (warp-accumulated halves -> one RED per word), out of line.
__device__ __forceinline__ void push_pressure(uint32_t* dst, const uint32_t (&acc)[kNS * 4]) {
  uint32_t v[kNS * 4];
#pragma unroll
  for (int q = 0; q < kNS * 4; ++q) v[q] = __reduce_add_sync(0xFFFFFFFFu, h2_to_u16x2(acc[q]));
  if ((threadIdx.x & 31) == 0 && dst) {
#pragma unroll
    for (int q = 0; q < kNS * 4; q += 2) red_add2(dst + q, v[q], v[q + 1]);
  }
}

// ---- one day of the team's slices: the run of the day's records that belongs to this warp (abm_wards.cpp:384-407) ----
// FIRST: the day of the initial state (its records may hold U0); a separate instance keeps that variant of the step out of
// the loop every other day runs.
// LEAN: a summary run (no trajectory) of a hospital with unit weights whose rows are already in (day, ward) order -- the
// common case, compiled without the code (and the registers) of the other cases.
// RED64: the transfers' pressure words are added two at a time (64-bit REDs).  Faster where the register budget is roomy
// (config 2: 5.99 -> 5.68 ms with the lean instance), slower in the 128-register lean instance (particle sweep: 36.5 -> 38.7 ms).
template <bool REPLAY, bool FIRST, bool LEAN, bool RED64>
__device__ __forceinline__ void process_day(PHASE_DECL const WarpShared& sh, const SliceParams* par, const FastArgs& a,
                                            const TeamCtx& c, int64_t d, const DayPlan& pl, TeamBarrier& bar) {
  const int lane = threadIdx.x & 31;
  constexpr int kNW = kNS * 4;              // state words of a record: word s*4 + j = simulations 8s + 2j, 8s + 2j + 1
  uint32_t ccol[kNW], cdet[kNW];            // per-lane day totals, packed halves
#pragma unroll
  for (int q = 0; q < kNW; ++q) ccol[q] = cdet[q] = 0;

  auto flush_counts = [&]() {
    uint32_t cnt[kNS * 8];
#pragma unroll
    for (int q = 0; q < kNW; ++q) { cnt[q] = ccol[q]; cnt[kNW + q] = cdet[q]; ccol[q] = cdet[q] = 0; }
    flush_day_counts(a, d, c.sim0, cnt);
  };

  if (pl.w0 < pl.w1) {
    const int64_t ring = a.ring_chunks * (kNS * 32), da = pl.da, seg0 = pl.seg0;
    const bool test_h = ((uint64_t)d % a.tsh) == 0, test_a = ((uint64_t)d % a.tsa) == 0;  // :391-392
    const uint4* pc = c.pre + (d & 1) * ring + (int64_t)(pl.w0 >> 5) * (kNS * 32) + lane;   // this lane's slot of the run's first chunk
    char* nxt = reinterpret_cast<char*>(c.pre + ((d + 1) & 1) * ring);
    const uint4* pm = a.meta + da + pl.w0 + lane;
    uint4* pt = (!LEAN && c.traj) ? c.traj + (da + pl.w0 + lane) * kNS : nullptr;
    const uint32_t x2 = kCodeA2 ^ ((kCodeZero - (uint32_t)a.ttd) * kHalves);   // A <-> "D = ttd" (:135 / :147)

    uint32_t pacc[kNW], wmain = kNone, wsegl = 0xFFFFFFFFu;  // wmain / wsegl are warp-uniform
#pragma unroll
    for (int q = 0; q < kNW; ++q) pacc[q] = 0;
    auto flush_warp = [&]() {
      uint32_t acc[kNW];   // a copy: the accumulators themselves must stay in registers
#pragma unroll
      for (int q = 0; q < kNW; ++q) { acc[q] = pacc[q]; pacc[q] = 0; }
      push_pressure(wmain != kNone ? c.press + (int64_t)wmain * kNW : nullptr, acc);
    };

    // The run is stepped in batches of at most kWarpTableSegs ward-days: build their tables, then run the trips that hold
    // their records.  A trip that straddles two batches is run once per batch, each time with the other batch's lanes
    // switched off (they step an uncolonised state against the all-zero table row: nothing happens, nothing is stored).
    const uint32_t n_run = pl.w1 - pl.w0;
    uint32_t since = 0;   // trips since the per-lane half counters were flushed (they stay exact up to kFlushEvery)
    for (uint32_t sl = pl.sl_first; sl <= pl.sl_last; sl += kWarpTableSegs) {
      const uint32_t n_tab = min((uint32_t)kWarpTableSegs, pl.sl_last + 1u - sl);
      __syncwarp();   // the previous batch's rows are no longer read
      build_tables<REPLAY, LEAN>(sh, par, a, c.press, c.gslice0, seg0, sl, sl + n_tab, test_h, test_a);
      __syncwarp();
      PHASE_ADD(0);  // table
      // trips (32-record chunks of the run) that hold records of these ward-days
      const int64_t sb = seg0 + sl;
      const uint32_t rb = max(pl.w0, (uint32_t)(__ldg(&a.seg_row_start[sb]) - da)) - pl.w0;
      const uint32_t re = min(pl.w1, (uint32_t)(__ldg(&a.seg_row_start[sb + n_tab]) - da)) - pl.w0;
      uint32_t tb = rb >> 5;
      const uint32_t t_end = (re + 31u) >> 5;
      while (tb < t_end) {
        const uint32_t te = min(t_end, tb + (kFlushEvery - since));
        const uint4* pmt = pm + (int64_t)tb * 32;
        const uint4* pct = pc + (int64_t)tb * (kNS * 32);
        uint4* ptt = pt ? pt + (int64_t)tb * (kNS * 32) : nullptr;
        // software pipeline: the loads of the next trip are in flight while this one computes
        uint4 n_meta = make_uint4(0u, kNone, kNone, 0u), n_prev[kNS];
#pragma unroll
        for (int s = 0; s < kNS; ++s) n_prev[s] = make_uint4(0u, 0u, 0u, 0u);
        if (tb * 32u + (uint32_t)lane < n_run) {
          n_meta = __ldg(pmt);
#pragma unroll
          for (int s = 0; s < kNS; ++s) n_prev[s] = __ldcg(pct + s * 32);
        }
#ifdef AMRO_TRIP_UNROLL
        AMRO_PRAGMA_UNROLL(AMRO_TRIP_UNROLL)
#endif
        for (uint32_t i = tb * 32u + lane; i < te * 32u + (uint32_t)lane; i += 32u) {   // i: run-relative; warp-uniform trip count
          const uint4 meta = n_meta;
          uint4 prev[kNS];
#pragma unroll
          for (int s = 0; s < kNS; ++s) prev[s] = n_prev[s];
          pmt += 32; pct += kNS * 32;
          n_meta = make_uint4(0u, kNone, kNone, 0u);
          if (i + 32u < min(n_run, te * 32u)) {
            n_meta = __ldg(pmt);
#pragma unroll
            for (int s = 0; s < kNS; ++s) n_prev[s] = __ldcg(pct + s * 32);
          }
          const uint32_t info = meta.x, segl = info & kInfoSegMask, src = (info >> kInfoSrcShift) & 3u;
          const int h = (int)(info >> kInfoHShift);
          const int64_t orig = (LEAN || a.identity_order) ? da + pl.w0 + i : (int64_t)meta.w;
          const bool on = i < n_run && segl - sl < n_tab;   // this batch's records (unsigned compare: segl >= sl too)
          const uint32_t wc = (info >> kInfoHalfShift) & 1u;   // the record weighs 1/2
          const int row = on ? ((int)(segl & (kWarpTableSegs - 1)) * kRowsPerSeg + h * 2 + (int)wc) * kNS : kZeroRow * kNS;
          const uint32_t next = on ? meta.y : kNone, nseg = on ? meta.z : kNone;
          uint32_t colw[kNW], detw[kNW];
          uint32_t u14[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if constexpr (!REPLAY) draws_u14x16(a, orig, c.g32, c.half, u14);
#pragma unroll
          for (int s = 0; s < kNS; ++s) {
            uint4 pre = prev[s];
            if (!on || src == kSrcNone) pre = make_uint4(0u, 0u, 0u, 0u);  // (C=0, D=inf) x 8, :359-360
            uint4 post;
            if constexpr (REPLAY) {
              post = step_record_replay(sh, par[s], a, pre, row + s, h, orig, c.sim0 + s * 8, h ? test_a : test_h, on);
            } else {
              const uint4 rnd = make_uint4(u14[(s & 1) * 4], u14[(s & 1) * 4 + 1], u14[(s & 1) * 4 + 2], u14[(s & 1) * 4 + 3]);
              post = step_record<FIRST>(sh.tab + (row + s) * kTabRow, pre, rnd, x2);
            }
            if (next != kNone) __stcg(reinterpret_cast<uint4*>(nxt + next) + s * 32, post);     // :402-407 the successor starts from this state
            if (ptt && on) __stcg(ptt + s, post);
            const uint32_t pw[4] = {post.x, post.y, post.z, post.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {                      // 1.0 per half-word: colonised (:477-485), D <= 0 (:500-509)
              colw[s * 4 + j] = h2_ge_one(pw[j], kCodeA2);
              detw[s * 4 + j] = h2_ge_one(pw[j], kCodeZero2);
            }
          }
#pragma unroll
          for (int q = 0; q < kNW; ++q) { ccol[q] = h2_add(ccol[q], colw[q]); cdet[q] = h2_add(cdet[q], detw[q]); }

          // ---- feed the successor's ward-day (its sum(W) of tomorrow, :95-99) --------------------------------
          // Most records of a warp's run feed the same ward-day (same ward, next day): lanes keep a running packed sum
          // for it, reduced across the warp (REDUX) and added with one RED per word only when that ward-day changes;
          // records that go elsewhere (transfers) add on their own.
          const uint32_t seg_lo = max(sl, __shfl_sync(0xFFFFFFFFu, segl, 0));   // the ward of the trip's first record of this batch (about)
          if (__builtin_expect(seg_lo != wsegl, 0)) {
            wsegl = seg_lo;
            const uint32_t m = __ldg(&a.seg_main_next[seg0 + seg_lo]);
            if (m != wmain) { flush_warp(); wmain = m; }
          }
          const bool has = nseg != kNone, mine = nseg == wmain;
          // what the record adds to its successor's ward-day: 1, or -- when the hospital has half-weight records and the
          // pressure is therefore counted in halves -- 2 if the successor is a whole record and 1 if it is a half one
          const bool two = (!LEAN && a.has_half) && !((info >> kInfoNextHalfShift) & 1u);   // the weight is the SUCCESSOR's (W = C * w of the row it becomes)
          const uint32_t mine2 = mine ? (two ? kTwo2 : kOne2) : 0u;
#pragma unroll
          for (int q = 0; q < kNW; ++q) pacc[q] = h2_fma(colw[q], mine2, pacc[q]);
          if (has && !mine) {
            uint32_t* dst = c.press + (int64_t)nseg * kNW;
            // half flags (0x3C00 per half-word) -> integer 1 (or 2) per half-word with one multiply-high on the FMA pipe:
            // floor(0x3C00 * 279621 / 2^32) = 1, floor(0x3C000000 * 279621 / 2^32) = 65536, and their sum for both
            const uint32_t mul = two ? 559242u : 279621u;
#pragma unroll
            for (int q = 0; q < kNW; q += 2) {
              if constexpr (!RED64) { red_add(dst + q, __umulhi(colw[q], mul)); red_add(dst + q + 1, __umulhi(colw[q + 1], mul)); }
              else red_add2(dst + q, __umulhi(colw[q], mul), __umulhi(colw[q + 1], mul));
            }
          }
          if (ptt) ptt += 32 * kNS;
        }
        since += te - tb;
        tb = te;
        if (since == kFlushEvery) { flush_counts(); since = 0; }
      }
    }
    PHASE_ADD(1);  // records (thread 0's view)
    flush_warp();
  }
  bar.arrive();   // tomorrow's pressure and states are out; the day totals below are nobody's input
  PHASE_ADD(4);
  flush_counts();
  PHASE_ADD(2);  // totals
}

template <bool REPLAY, bool LEAN, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_fast(const __grid_constant__ FastArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kWarps = THREADS / 32;
  constexpr bool kRed64 = !LEAN || (THREADS == kTeamThreads && MINB < kTeamMinBlocks);   // the roomy instance
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  SliceParams* par = reinterpret_cast<SliceParams*>(smem_raw);   // [kNS]
  const WarpShared sh = carve_shared<REPLAY>(smem_raw + kNS * sizeof(SliceParams), warp);
  // one team per kNS slices of 8 simulations; the first teams_plus_one teams have ctas_per_team + 1 CTAs (the grid is sized
  // to give every SM the same number of CTAs: an SM with one CTA more than its neighbours would make its teams wait)
  const int big = a.teams_plus_one * (a.ctas_per_team + 1);
  const bool in_big = (int)blockIdx.x < big;
  const int G = a.ctas_per_team + (in_big ? 1 : 0);
  const int64_t team = in_big ? blockIdx.x / G : a.teams_plus_one + ((int)blockIdx.x - big) / G;
  const int rank = in_big ? blockIdx.x % G : ((int)blockIdx.x - big) % G;
  const uint32_t gw = (uint32_t)(rank * kWarps + warp), n_gw = (uint32_t)(G * kWarps);   // this warp within its team
  TeamBarrier bar{a.team_barrier + team, 0u, n_gw, G == 1 ? kBarCta : (a.cluster_barrier ? kBarCluster : kBarGlobal)};
  const int64_t ring = a.ring_chunks * (kNS * 32);

  TeamCtx c;
  c.pre = a.pre + team * 2 * ring;
  c.traj = a.traj ? a.traj + team * a.R * kNS : nullptr;
  c.press = a.pressure + team * a.n_seg * (kNS * 4);
  c.gslice0 = (uint32_t)((a.sim_offset >> 3) + (uint64_t)team * kNS);
  static_assert(REPLAY || kNS == 2, "the free-running draws of a team are one half of a 32-simulation group");
  c.g32 = c.gslice0 >> 2; c.half = (c.gslice0 >> 1) & 1u;   // sim_offset is a multiple of 16 (api.cu)
  c.sim0 = team * (kNS * 8);

  // ---- team setup: parameters and the all-zero table row to shared memory, pressure zeroed -----------------------
  if (tid < 48 * kNS) {
    const int s = tid / 48, col = (tid % 48) >> 3, k = tid & 7;
    const double v = a.params[(c.sim0 + s * 8 + k) + a.p_ld * col];
    SliceParams& p = par[s];
    double* dst = col == 0 ? p.alpha : col == 1 ? p.beta : col == 2 ? p.gamma : col == 3 ? p.rho_h : col == 4 ? p.alpha2 : p.rho_i;
    dst[k] = v;
  }
  if constexpr (REPLAY) {
    for (int q = lane; q < kNS * 24; q += 32) sh.P[kZeroRow * kNS * 24 + q] = 0.0;
  } else {
    for (int q = lane; q < kNS * kTabRow; q += 32) sh.tab[kZeroRow * kNS * kTabRow + q] = make_uint4(0u, 0u, 0u, 0u);
  }
  {
    const int64_t n = a.n_seg * (kNS * 4), per = (n + G - 1) / G;
    const int64_t lo = rank * per, hi = min(n, lo + per);
    for (int64_t i = lo + tid; i < hi; i += THREADS) __stcg(&c.press[i], 0u);
  }
  __syncthreads();
  if (tid < 16 * kNS) {   // thresholds of colonised patients under a finite force of infection (0 * F = 0)
    const int s = tid >> 4, h = (tid >> 3) & 1, k = tid & 7;
    SliceParams& p = par[s];
    const double P1 = class_probability(1, h, 0.0, p.alpha[k], p.alpha2[k], p.gamma[k]);
    const double P2 = class_probability(2, h, 0.0, p.alpha[k], p.alpha2[k], p.gamma[k]);
    p.t30[0][h][k] = threshold30(P1);
    p.t30[1][h][k] = threshold30(P2);
    p.t30[2][h][k] = threshold30(__dmul_rn(clamp01(P1), h ? p.rho_i[k] : p.rho_h[k]));
    if (!REPLAY && h == 0)
      p.r16_run[k] = chunk16(philox4x32_10((uint32_t)kRunDitherIndex, (uint32_t)(kRunDitherIndex >> 32), c.gslice0 + (uint32_t)s, kStreamDither,
                                           (uint32_t)a.seed, (uint32_t)(a.seed >> 32)), k);
  }
  __syncthreads();   // parameters visible to every warp of the CTA
  bar.arrive();
  bar.wait();

  // ---- initial state (abm_wards.cpp:363-372) into the day-0 buffer -----------------------------------------
  {
    const int64_t n = a.n_init, per = ((n + G - 1) / G + 31) / 32 * 32;
    const int64_t lo = rank * per, hi = min(n, lo + per);
    for (int64_t r = lo + tid; r < hi; r += THREADS) {
      const uint32_t slot = a.init_slot[r];
      if (slot == kNone) continue;  // overwritten by a predecessor before it is ever read
#pragma unroll 1
      for (int s = 0; s < kNS; ++s) {
        uint32_t w[4] = {0, 0, 0, 0}, cw[4] = {0, 0, 0, 0};
        uint4 rc_hi = make_uint4(0, 0, 0, 0), rc_lo = rc_hi, rd_hi = rc_hi, rd_lo = rc_hi;
        if constexpr (!REPLAY) {   // one call per stream serves the slice's eight simulations (rng.cuh: philox_u32)
          const uint32_t r0 = (uint32_t)r, r1 = (uint32_t)((uint64_t)r >> 32), g = c.gslice0 + (uint32_t)s;
          const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
          rc_hi = philox4x32_10(r0, r1, g, kStreamInitCol, k0, k1); rc_lo = philox4x32_10(r0, r1, g, kStreamInitCol + 1u, k0, k1);
          rd_hi = philox4x32_10(r0, r1, g, kStreamInitDet, k0, k1); rd_lo = philox4x32_10(r0, r1, g, kStreamInitDet + 1u, k0, k1);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int64_t sim = c.sim0 + s * 8 + k;
          bool c0 = false, d0 = false;
          if (sim < a.S) {
            const double pc = a.icp[r * a.icp_rs + sim * a.icp_cs], pd = a.idp[r * a.idp_rs + sim * a.idp_cs];
            if (REPLAY) {
              c0 = a.u_icol[r + a.ui_ld * sim] < pc;
              d0 = a.u_idet[r + a.ui_ld * sim] < pd;
            } else {
              c0 = (uint64_t)((chunk16(rc_hi, k) << 16) | chunk16(rc_lo, k)) < threshold33(pc);
              d0 = (uint64_t)((chunk16(rd_hi, k) << 16) | chunk16(rd_lo, k)) < threshold33(pd);
            }
          }
          // :372  D0 = d0*c0 is a 0/1 FLAG that the dynamics read as a countdown on day 0:
          //   uncolonised -> D = 0 (U0);  colonised, flag 0 -> D = 0: detected;  flag 1 -> D = 1: pending if
          //   ttd >= 1, else (1 > ttd) it is neither pending nor detected and gets re-tested like D = inf
          const uint32_t code = !c0 ? kCodeU0 : (!d0 ? kCodeZero : (a.ttd >= 1 ? kCodeZero - 1u : kCodeA));
          w[k >> 1] |= code << (16 * (k & 1));
          cw[k >> 1] |= (c0 ? 1u : 0u) << (16 * (k & 1));
        }
        __stcg(&c.pre[ring_slot(slot) + s * 32], make_uint4(w[0], w[1], w[2], w[3]));
        const uint32_t iseg = a.init_seg[r];
        // initial rows sit on day 0 (fast-path condition): position = slot; a whole record counts 2 when pressure is in halves
        const uint32_t unit = (a.has_half && !((__ldg(&a.meta[a.day_start[0] + slot].x) >> kInfoHalfShift) & 1u)) ? 2u : 1u;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (cw[j]) red_add(&c.press[((int64_t)iseg * kNS + s) * 4 + j], cw[j] * unit);
      }
    }
  }
  bar.arrive();

  // ---- day loop (abm_wards.cpp:381): one team barrier per day ----------------------------------------------
  PHASE_T0();
  DayPlan plan = plan_day(a, 0, gw, n_gw, DayPlan{});
  for (int64_t d = 0; d < a.n_days; ++d) {
    bar.wait();
    PHASE_ADD(3);  // team barrier wait
    if (d == 0) process_day<REPLAY, true, LEAN, kRed64>(PHASE_ARG sh, par, a, c, d, plan, bar);   // initial rows sit on day 0 (fast-path condition)
    else process_day<REPLAY, false, LEAN, kRed64>(PHASE_ARG sh, par, a, c, d, plan, bar);
    plan = plan_day(a, d + 1, gw, n_gw, plan);   // static: ready before the next wait ends
  }
  if (bar.mode == kBarCluster) bar.wait();   // every arrive of a cluster barrier is paired with a wait
}

// ---- packed trajectory -> the reference's float64 columns ------------------------------------------------------
namespace {
// rows [r0, r0 + nr) x simulations [s0, s0 + ns) of the run -> Cout / Dout [nr x ns], element (r - r0) + ld * (s - s0)
__global__ void k_fast_expand(const uint4* __restrict__ traj, const int64_t* __restrict__ pos_of, int64_t R, int64_t S,
                              int64_t r0, int64_t nr, int64_t s0, int64_t ns, double* __restrict__ Cout, double* __restrict__ Dout, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_sl = (ns + 7) / 8;  // s0 is a multiple of 8
  if (i >= nr * n_sl) return;
  const int64_t rr = i % nr, sl = i / nr, r = r0 + rr;
  const int64_t slice = s0 / 8 + sl, team = slice / kNS, in_team = slice % kNS;
  const int64_t p = pos_of ? pos_of[r] : r;
  const uint4 v = __ldg(&traj[(team * R + p) * kNS + in_team]);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t s = sl * 8 + k;
    if (s >= ns || s0 + s >= S) break;
    const uint32_t x = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    Cout[rr + ld * s] = (x >= kCodeA) ? 1.0 : 0.0;
    Dout[rr + ld * s] = (x >= kCodeCnt) ? (double)((int)kCodeZero - (int)x) : __longlong_as_double(0x7FF0000000000000ll);
  }
}

__global__ void k_fast_export_pressure(const uint32_t* __restrict__ pressure, int64_t n_seg, int64_t S,
                                       int32_t* __restrict__ out, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg * S) return;
  const int64_t seg = i % n_seg, s = i / n_seg;
  const int k = (int)(s & 7);
  const int64_t slice = s >> 3, team = slice / kNS, in_team = slice % kNS;
  const uint32_t w = pressure[((team * n_seg + seg) * kNS + in_team) * 4 + (k >> 1)];
  out[seg + ld * s] = (int32_t)((w >> (16 * (k & 1))) & 0xFFFFu);
}
}  // namespace

#ifdef AMRO_PROFILE_PHASES
extern "C" void amro_debug_phase_cycles(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles));
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); }
}
#endif

// Two launch shapes of the same kernel:
//   WIDE  one big CTA per SM owns a slice by itself (team of 1: only __syncthreads between days); slices beyond the SM
//         count run in further waves.  For ensembles with at least about one slice per SM.
//   TEAM  small CTAs, several per SM, cooperate on a slice with one team barrier per day (cooperative launch).  For
//         few simulations and many records per day.
namespace {
template <bool REPLAY, bool LEAN>
const void* kernel_ptr_of(int shape) {
  return shape == kShapeWide    ? (const void*)k_fast<REPLAY, LEAN, kWideThreads, 1>
         : shape == kShapeRoomy ? (const void*)k_fast<REPLAY, LEAN, kTeamThreads, kTeamMinBlocks - 1>
                                : (const void*)k_fast<REPLAY, LEAN, kTeamThreads, kTeamMinBlocks>;
}
// the lean instance exists for free-running runs only (replay is a correctness mode)
const void* kernel_ptr(bool replay, bool lean, int shape) {
  return replay ? kernel_ptr_of<true, false>(shape) : lean ? kernel_ptr_of<false, true>(shape) : kernel_ptr_of<false, false>(shape);
}
}  // namespace

int fast_threads(int shape) { return shape == kShapeWide ? kWideThreads : kTeamThreads; }
size_t fast_shared_bytes(bool replay, int shape) { return fast_smem_bytes(replay, fast_threads(shape)); }

int fast_max_coresident_ctas(int device, bool replay, bool lean, int shape) {
  int per_sm = 0, sms = 0;
  const void* k = kernel_ptr(replay, lean, shape);
  const size_t smem = fast_smem_bytes(replay, fast_threads(shape));
  AMRO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AMRO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, fast_threads(shape), smem));
  return per_sm * sms;
}

// CTAs of the cluster launch that can be co-resident, in clusters (0 if the shape cannot be launched as clusters)
int fast_max_coresident_clusters(int device, bool replay, bool lean, int shape, int cluster_size) {
  (void)device;
  const void* k = kernel_ptr(replay, lean, shape);
  const size_t smem = fast_smem_bytes(replay, fast_threads(shape));
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cluster_size > 8 && cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)cluster_size * 1024u);
  cfg.blockDim = dim3((unsigned)fast_threads(shape));
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void fast_launch(cudaStream_t st, const FastArgs& a, bool replay, bool lean, int shape, int grid) {
  const void* k = kernel_ptr(replay, lean, shape);
  const size_t smem = fast_smem_bytes(replay, fast_threads(shape));
  AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {(void*)&a};
  if (a.cluster_barrier) {   // every team is one thread-block cluster (the hardware co-schedules its CTAs)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)fast_threads(shape));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.ctas_per_team; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AMRO_CUDA(cudaLaunchKernelExC(&cfg, k, params));
  } else if (a.ctas_per_team > 1 || a.teams_plus_one > 0) {
    AMRO_CUDA(cudaLaunchCooperativeKernel(k, dim3((unsigned)grid), dim3((unsigned)fast_threads(shape)), params, smem, st));
  } else {
    AMRO_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3((unsigned)fast_threads(shape)), params, smem, st));
  }
}

void fast_expand(cudaStream_t st, const uint4* traj, const int64_t* pos_of, int64_t R, int64_t S, int64_t r0, int64_t nr, int64_t s0,
                 int64_t ns, double* Cout, double* Dout, int64_t ld) {
  const int64_t n = nr * ((ns + 7) / 8);
  if (n <= 0) return;
  k_fast_expand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(traj, pos_of, R, S, r0, nr, s0, ns, Cout, Dout, ld);
}

void fast_export_pressure(cudaStream_t st, const uint32_t* pressure, int64_t n_seg, int64_t S, int32_t* out, int64_t ld) {
  const int64_t n = n_seg * S;
  if (n <= 0) return;
  k_fast_export_pressure<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pressure, n_seg, S, out, ld);
}

}  // namespace amro
