// FAST path: the ward-level colonisation step for a whole run in ONE persistent cooperative kernel.
//
// Replaces the nest  simulate_discrete_model_internal_one -> progress_patients_1_timestep ->
// progress_patients_probability_ward_1_timestep  (src/amro/abm_wards.cpp:322-414 / :219-257 / :70-159) fused
// with total_positive_colonized/_detected (:435-509), for hospitals with unit weights, is_new in {0,1},
// time_to_detect >= 0 and next_day links that connect consecutive days.
//
// State of one (record, simulation): a 16-bit code whose top 3 bits are its class.
//   0x0000        U   uncolonised                              (reference: C=0, D=inf)
//   0x2000        A   colonised, no test result coming         (C=1, D=inf)
//   0x4000        U0  uncolonised with a zero countdown        (C=0, D=0: the initial-state quirk of :372)
//   0x8000 - D        colonised with a finite countdown D: class 3 (0x6000..0x7FFF) while the result is pending
//                     (0 < D <= ttd), class 4 (0x8000..) once detected (D <= 0).  The reference's daily
//                     `D -= 1` (:131-132) is `code += 1`, and "pending -> detected" is the carry into bit 15.
//   Two simulations share a 32-bit word and are stepped together (SIMD within a register): class membership, the
//   Bernoulli compares and the transition are all "add a constant, replicate bit 15 / bit 31 into a half-word mask
//   (one PRMT), select with LOP3".  U and A have one code each (U -> A | 0x8000-ttd, A -> A | 0x8000-ttd), a running
//   countdown is `code + 1`, U0 (first day only) steps like D = 0 with the probability of an uncolonised patient.
//
// Data layout (HBM):
//   pre       [set][2][max records per day]  one 16-byte word = 8 simulations per record; buffer d&1 holds the
//             state the records of day d start from.  A record's step WRITES its result into its successor's
//             slot of buffer (d+1)&1 (next_day is a scatter, :402-407), so reading the pre-step state needs no
//             dependent load.  Records without a predecessor ignore the slot and start uncolonised (:359-360).
//   pressure  [set][ward-day][4] packed 16-bit counts: colonised patients a ward-day STARTS with = the
//             reference's sum(W) (:99), accumulated by the previous day's pass (every record adds its new C to
//             the ward-day of its successor): one pass over a day's records instead of two.
//   meta      three 32-bit words per record shared by all simulations: (is_new, source, ward-day), successor
//             slot, successor ward-day.
// Work decomposition: a TEAM of ctas_per_team CTAs owns one slice (8 simulations) for every day; lanes map to
// consecutive records, so state loads/stores are 128-bit and coalesced.  Teams never talk to each other; inside
// a team the CTAs split each day's records and meet at one barrier per day.  Per ward-day, per simulation, the
// transition probabilities a record can have (class x is_new) are evaluated in fp64 exactly as the reference
// evaluates them and turned into integer thresholds kept in shared memory.
#include "common.hpp"
#include "kernels.hpp"
#include "rng.cuh"
#include "topology.hpp"

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

constexpr uint32_t kClsU = 0, kClsA = 1, kClsU0 = 2, kClsPn = 3, kClsB = 4, kNumCls = 5;
constexpr uint32_t kClsShift = 13;
constexpr uint32_t kCodeA = kClsA << kClsShift, kCodeU0 = kClsU0 << kClsShift, kCodeZero = kClsB << kClsShift;
constexpr int kTabRow = 5;                      // 16-byte entries per (ward-day, is_new): see SharedView
constexpr int kFlushEvery = 1024;              // row iterations between flushes of the 16-bit packed per-lane counters
constexpr uint32_t kHalves = 0x00010001u;      // one in each half-word
constexpr uint32_t kCodeA2 = kCodeA * kHalves, kCodeZero2 = kCodeZero * kHalves;

__device__ __forceinline__ void red_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 0xFFFF in every half-word whose top bit (bit 15 / bit 31) is set: PRMT with sign replication of bytes 1 and 3
__device__ __forceinline__ uint32_t half_mask(uint32_t x) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(x));
  return m;
}

struct SliceParams {  // the 8 simulations' parameter rows (abm_wards.cpp:79-84)
  double alpha[8], beta[8], gamma[8], rho_h[8], alpha2[8], rho_i[8];
};

// Transition probability for unit weight (W = C), force of infection F -- the reference's expression order, one
// IEEE operation at a time (:95-112).  pc: 0 = uncolonised, 1 = colonised undetected, 2 = colonised detected.
__device__ __forceinline__ double class_probability(int pc, int h, double F, double alpha, double alpha2, double gamma) {
  const double W = pc ? 1.0 : 0.0;
  const double det = (pc == 2) ? 1.0 : 0.0;
  const double coef = __dadd_rn(__dmul_rn(__dsub_rn(1.0, det), __dsub_rn(1.0, alpha)),
                                __dmul_rn(det, __dsub_rn(1.0, alpha2)));
  double P = __dadd_rn(__dmul_rn(coef, W), __dmul_rn(__dsub_rn(1.0, W), F));
  P = __dadd_rn(P, __dmul_rn(h ? 1.0 : 0.0, gamma));
  return P;
}

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

// Shared memory (dynamic): the tables of up to a.table_segs ward-days, then the slice parameters.
//   free-running: tab  uint4[table_segs*2*kTabRow] at [(ward-day*2 + is_new)*kTabRow + e]; word j of an entry belongs
//                      to simulations 2j (low half) and 2j+1 (high half); 15-bit Bernoulli thresholds b (rng.cuh):
//                        e=0  bU               uncolonised:        colonised today  <=>  v15 + bU  >= 0x8000
//                        e=1  bU ^ bA          A / pending
//                        e=2  bA ^ bB          detected
//                        e=3  bUh              uncolonised, tested positive today (0 when nobody is tested)
//                        e=4  bUh ^ bAh        A
//   replay:       P    double[table_segs*2][3][8]         the transition probabilities themselves
struct SharedView {
  uint4* tab;
  double* P;
  SliceParams* par;
};
__host__ __device__ inline size_t fast_smem_bytes(bool replay, int table_segs) {
  const size_t rows = (size_t)table_segs * 2;
  return (replay ? rows * 24 * sizeof(double) : rows * kTabRow * sizeof(uint4)) + sizeof(SliceParams) + 16;
}
template <bool REPLAY>
__device__ __forceinline__ SharedView carve_shared(unsigned char* raw, int table_segs) {
  SharedView v{};
  const size_t rows = (size_t)table_segs * 2;
  if constexpr (REPLAY) {
    v.P = reinterpret_cast<double*>(raw);
    v.par = reinterpret_cast<SliceParams*>(raw + rows * 24 * sizeof(double));
  } else {
    v.tab = reinterpret_cast<uint4*>(raw);
    v.par = reinterpret_cast<SliceParams*>(raw + rows * kTabRow * sizeof(uint4));
  }
  return v;
}

// Philox4x32-10 with the round keys taken from the kernel parameters (constant bank operands)
__device__ __forceinline__ uint4 philox_rk(const FastArgs& a, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t kx) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a.rk[2 * r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ a.rk[2 * r + 1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  (void)kx;
  return make_uint4(c0, c1, c2, c3);
}

// One record, eight simulations, free-running mode; two simulations per 32-bit word, stepped together.
//   class masks   detected: bit 15 of the code; running countdown: code >= 0x6000; colonised: code >= 0x2000 -- an add
//                 that moves the boundary to bit 15, then half_mask()
//   thresholds    b1 = bU ^ (colonised & (bU^bA)) ^ (detected & (bA^bB)),  b2 = bUh ^ (colonised & (bUh^bAh))
//   draws         colonised today = bit 15 of v15 + b1, tested positive = bit 15 of v15 + b2 (one uniform for both, so
//                 P(colonised and positive) = P * rho exactly)
//   transition    countdown: code + 1;  U / A: A, or 0x8000 - ttd when tested positive;  not colonised: 0 (:153-155)
// FIRST: the record's state may hold U0 (uncolonised, D = 0; day of the initial state only): probability of an uncolonised
// patient, transition of a countdown at 0.
// colf[j] gets one bit per half-word: colonised after the step.
template <bool FIRST>
__device__ __forceinline__ uint4 step_record(const uint4* __restrict__ tab, const uint4 pre, const uint4 rnd, uint32_t x2,
                                             uint32_t (&colf)[4]) {
  const uint4 tU = tab[0], tXA = tab[1], tXB = tab[2], tUh = tab[3], tXAh = tab[4];
  const uint32_t prew[4] = {pre.x, pre.y, pre.z, pre.w}, rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
  const uint32_t bU[4] = {tU.x, tU.y, tU.z, tU.w}, xA[4] = {tXA.x, tXA.y, tXA.z, tXA.w}, xB[4] = {tXB.x, tXB.y, tXB.z, tXB.w};
  const uint32_t bH[4] = {tUh.x, tUh.y, tUh.z, tUh.w}, xH[4] = {tXAh.x, tXAh.y, tXAh.z, tXAh.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w = prew[j];
    const uint32_t av = rw[j] & 0x7FFF7FFFu;
    const uint32_t mdet = half_mask(w);
    uint32_t mcnt = half_mask(w + 0x20002000u);      // code >= 0x6000 (codes stay below 0xA000: no carry between halves)
    uint32_t mcol = half_mask(w + 0x60006000u);      // code >= 0x2000
    if constexpr (FIRST) {
      const uint32_t mge4 = half_mask(w + 0x40004000u);
      const uint32_t mu0 = mge4 & ~mcnt;              // U0 = 0x4000
      mcol &= ~mu0;
      w = (w & ~mu0) | (kCodeZero2 & mu0);
      mcnt = mge4;
    }
    const uint32_t b1 = bU[j] ^ (mcol & xA[j]) ^ (mdet & xB[j]);
    const uint32_t b2 = bH[j] ^ (mcol & xH[j]);
    const uint32_t m1 = half_mask(av + b1), m2 = half_mask(av + b2);
    const uint32_t t = m2 & x2 & ~mcnt;                                 // U / A tested positive: A -> 0x8000 - ttd
    const uint32_t q = (mcnt & (w + kHalves)) | (~mcnt & kCodeA2);      // D -= 1 (:131-132)  |  A
    o[j] = m1 & (q ^ t);
    colf[j] = m1 & kHalves;
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// One record, eight simulations, replay mode: the caller's uniforms, compared in fp64 as the reference does.
__device__ __forceinline__ uint4 step_record_replay(const SharedView& sh, const SliceParams& par, const FastArgs& a, const uint4 pre, int row, int h,
                                                    int64_t orig, int64_t sim0, bool flag) {
  const uint32_t prew[4] = {pre.x, pre.y, pre.z, pre.w};
  uint32_t postw[4] = {0, 0, 0, 0};
  const double* rho_tab = h ? par.rho_i : par.rho_h;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t v = (prew[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    const uint32_t cls = v >> kClsShift;
    const int64_t sim = sim0 + k;
    const double P = sh.P[(row * 3 + prob_class(cls)) * 8 + k];
    bool col = false, hit = false;
    if (sim < a.S) {
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

// Everything a CTA needs to step its slice (8 simulations).
struct SliceCtx {
  uint4* pre;
  uint4* traj;
  uint32_t* press;
  uint32_t gslice;
  int64_t sim0;
};

// ---- one day of one slice: the part of the day's records that belongs to this CTA (abm_wards.cpp:384-407) ----------
template <bool REPLAY, int THREADS>
__device__ __forceinline__ void process_day(PHASE_DECL const SharedView& sh, const FastArgs& a, const SliceCtx& c, int64_t d, int rank, int G) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t ring = a.max_rows_per_day;
  const int64_t da = a.day_start[d], db = a.day_start[d + 1];
  const uint32_t n = (uint32_t)(db - da), per = ((n + G - 1) / G + 31) / 32 * 32;
  const uint32_t lo = min(n, (uint32_t)rank * per), hi = min(n, lo + per);   // this CTA's records, day-relative
  const bool test_h = ((uint64_t)d % a.tsh) == 0, test_a = ((uint64_t)d % a.tsa) == 0;  // :391-392
  const int64_t seg0 = a.day_seg_start[d];
  const uint4* cur = c.pre + (d & 1) * ring;
  uint4* nxt = c.pre + ((d + 1) & 1) * ring;
  uint32_t* press = c.press;
  const uint32_t* m_info = a.meta_info + da;
  const uint32_t* m_next = a.meta_next + da;
  const uint32_t* m_nseg = a.meta_nseg + da;
  const SliceParams& par = *sh.par;
  const int tsegs = a.table_segs;
  uint32_t ccol[4] = {0, 0, 0, 0}, cdet[4] = {0, 0, 0, 0};
  int since_flush = 0;
  const uint32_t x2 = kCodeA2 ^ ((kCodeZero - (uint32_t)a.ttd) * kHalves);   // A <-> "D = ttd" (:135 / :147)

  // per-day totals (abm_wards.cpp:445-459 fused): warp-reduce the packed per-lane counters, then one RED per
  // (simulation, metric) and warp
  auto flush_counts = [&]() {
    uint32_t tot[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tot[j] = __reduce_add_sync(0xFFFFFFFFu, ccol[j]);
      tot[4 + j] = __reduce_add_sync(0xFFFFFFFFu, cdet[j]);
      ccol[j] = 0;
      cdet[j] = 0;
    }
    if (lane < 16 && d < a.counts_ld) {
      uint32_t w = tot[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) w = ((lane >> 1) == j) ? tot[j] : w;
      const uint32_t v = (lane & 1) ? (w >> 16) : (w & 0xFFFFu);
      int32_t* dst = (lane < 8) ? a.counts_col : a.counts_det;
      if (v != 0 && dst) atomicAdd(&dst[d + a.counts_ld * (c.sim0 + (lane & 7))], (int)v);
    }
    since_flush = 0;
  };

  PHASE_ADD(5);  // day set-up
  if (lo < hi) {
    const int64_t seg_first = seg0 + (m_info[lo] & kInfoSegMask);
    const int64_t seg_last = seg0 + (m_info[hi - 1] & kInfoSegMask);
    for (int64_t sb = seg_first; sb <= seg_last; sb += tsegs) {
      const int64_t se = min(seg_last + 1, sb + tsegs);
      // ---- threshold / delta tables of ward-days [sb, se) ------------------------------------------------
      __syncthreads();
      for (int t0 = 0; t0 < (int)(se - sb) * 16; t0 += THREADS) {   // uniform trip count: the pairing below shuffles
        const int t = t0 + tid;
        const bool valid = t < (int)(se - sb) * 16;
        const int si = valid ? (t >> 4) : 0, h = (t >> 3) & 1, k = t & 7;
        const int64_t seg = sb + si;
        const uint32_t pk = __ldcg(&press[seg * 4 + (k >> 1)]);
        const uint32_t cnt = (k & 1) ? (pk >> 16) : (pk & 0xFFFFu);
        // :98-99  F = (beta / N) * sum(W); with unit weights sum(W) is the exact integer count
        const double F = __dmul_rn(__ddiv_rn(par.beta[k], a.seg_N[seg]), (double)cnt);
        const bool flag = h ? test_a : test_h;
        const double rho = h ? par.rho_i[k] : par.rho_h[k];
        double P[3];
#pragma unroll
        for (int pc = 0; pc < 3; ++pc) P[pc] = class_probability(pc, h, F, par.alpha[k], par.alpha2[k], par.gamma[k]);
        if constexpr (REPLAY) {
          if (valid) {
#pragma unroll
            for (int pc = 0; pc < 3; ++pc) sh.P[((si * 2 + h) * 3 + pc) * 8 + k] = P[pc];
          }
        } else {
          // dithered 15-bit thresholds (rng.cuh); U and A are tested today if it is a testing day, a running countdown is
          // never re-tested (:131 / :134)
          const uint32_t r16 = chunk16(philox4x32_10((uint32_t)seg, (uint32_t)((uint64_t)seg >> 32), c.gslice, kStreamDither,
                                                     (uint32_t)a.seed, (uint32_t)(a.seed >> 32)), k);
          uint32_t e[5];
          e[0] = bernoulli15(P[0], r16);
          e[1] = bernoulli15(P[1], r16);
          e[2] = bernoulli15(P[2], r16);
          e[3] = flag ? bernoulli15(__dmul_rn(clamp01(P[0]), rho), r16) : 0u;
          e[4] = flag ? bernoulli15(__dmul_rn(clamp01(P[1]), rho), r16) : 0u;
          // pair up with the neighbouring simulation (lane ^ 1): word j = sims (2j, 2j+1)
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, e[q], 1);
            e[q] = (k & 1) ? (other | (e[q] << 16)) : (e[q] | (other << 16));
          }
          if (valid && !(k & 1)) {
            uint32_t* row = reinterpret_cast<uint32_t*>(sh.tab + (si * 2 + h) * kTabRow) + (k >> 1);
            row[0] = e[0]; row[4] = e[0] ^ e[1]; row[8] = e[1] ^ e[2]; row[12] = e[3]; row[16] = e[3] ^ e[4];
          }
        }
      }
      __syncthreads();
      PHASE_ADD(0);  // table
      // ---- the records of those ward-days that belong to this CTA ----------------------------------------
      const uint32_t rb = max(lo, (uint32_t)(a.seg_row_start[sb] - da)), re = min(hi, (uint32_t)(a.seg_row_start[se] - da));
      // Each warp owns a contiguous run of 32-record chunks (dealt out evenly): consecutive trips of a warp stay in the
      // same ward, so what its records add to tomorrow's pressure can be kept in registers across trips.
      constexpr int kWarps = THREADS / 32;
      const uint32_t n_chunks = (re - rb + 31) / 32;
      const uint32_t w0 = min(re, rb + 32u * (uint32_t)(((uint64_t)n_chunks * (tid >> 5)) / kWarps));
      const uint32_t w1 = min(re, rb + 32u * (uint32_t)(((uint64_t)n_chunks * ((tid >> 5) + 1)) / kWarps));
      const uint32_t first = w0 + lane, step = 32u, lim = w1, trips_end = w1 + (uint32_t)lane;
      // software pipeline: the loads of the next trip are in flight while this one computes
      uint32_t n_info = 0, n_next = kNone, n_nseg = kNone;
      uint4 n_prev = make_uint4(0u, 0u, 0u, 0u);
      if (first < lim) {
        n_info = __ldg(&m_info[first]); n_next = __ldg(&m_next[first]); n_nseg = __ldg(&m_nseg[first]);
        n_prev = __ldcg(&cur[first]);
      }
      uint32_t pacc[4] = {0, 0, 0, 0}, wmain = kNone, wsegl = 0xFFFFFFFFu;  // wmain / wsegl are warp-uniform
      auto flush_warp = [&]() {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = __reduce_add_sync(0xFFFFFFFFu, pacc[j]); pacc[j] = 0; }
        if (lane == 0 && wmain != kNone) {
          uint32_t* dst = press + (int64_t)wmain * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) red_add(dst + j, v[j]);
        }
      };

      for (uint32_t i = first; i < trips_end; i += step) {  // warp-uniform trip count (warp collectives below)
        const uint32_t info = n_info, next = n_next;
        uint4 prev = n_prev;
        uint32_t colw[4] = {0, 0, 0, 0}, detw[4] = {0, 0, 0, 0};
        uint32_t nseg = i < lim ? n_nseg : kNone;
        if (i + step < lim) {
          n_info = __ldg(&m_info[i + step]); n_next = __ldg(&m_next[i + step]);
          n_nseg = __ldg(&m_nseg[i + step]); n_prev = __ldcg(&cur[i + step]);
        }
        if (i < lim) {
          const int64_t p = da + i;
          const int64_t orig = a.orig_row ? a.orig_row[p] : p;
          const int h = (int)(info >> kInfoHShift);
          const int row = ((int)((int64_t)(info & kInfoSegMask) + seg0 - sb)) * 2 + h;
          const uint32_t src = (info >> kInfoSrcShift) & 3u;
          if (src == kSrcNone) prev = make_uint4(0u, 0u, 0u, 0u);  // (C=0, D=inf) x 8, :359-360
          uint4 post;
          if constexpr (REPLAY) {
            post = step_record_replay(sh, par, a, prev, row, h, orig, c.sim0, h ? test_a : test_h);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t pw = j == 0 ? post.x : j == 1 ? post.y : j == 2 ? post.z : post.w;
              colw[j] = ((pw >> 13) | (pw >> 15)) & kHalves;             // classes 1, 3 (bit 13) and 4 (bit 15)
            }
          } else {
            // U0 only exists in states that come straight from the initial draw
            const uint4 rnd = philox_rk(a, (uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), c.gslice, kStreamStep, 0);
            if (src == kSrcInit) post = step_record<true>(sh.tab + row * kTabRow, prev, rnd, x2, colw);
            else post = step_record<false>(sh.tab + row * kTabRow, prev, rnd, x2, colw);
          }
          if (next != kNone) __stcg(&nxt[next], post);     // :402-407 the successor starts from this state
          if (c.traj) __stcg(&c.traj[p], post);
          const uint32_t pw[4] = {post.x, post.y, post.z, post.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) detw[j] = (pw[j] >> 15) & kHalves;   // class 4: D <= 0
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { ccol[j] += colw[j]; cdet[j] += detw[j]; }

        // ---- feed the successor's ward-day (its sum(W) of tomorrow, :95-99) --------------------------------
        // Most records of a warp's run feed the same ward-day (same ward, next day): lanes keep a running packed sum
        // for it, reduced across the warp (REDUX) and added with one RED per word only when that ward-day changes;
        // records that go elsewhere (transfers) add on their own.
        {
          // the ward-day most records of this chunk feed: same ward as the chunk's first record, next day (warp-uniform)
          const uint32_t segl0 = __shfl_sync(0xFFFFFFFFu, info & kInfoSegMask, 0);
          if (segl0 != wsegl) {
            wsegl = segl0;
            const uint32_t m = __ldg(&a.seg_main_next[seg0 + segl0]);
            if (m != wmain) { flush_warp(); wmain = m; }
          }
          const bool has = nseg != kNone, mine = nseg == wmain;
#pragma unroll
          for (int j = 0; j < 4; ++j) pacc[j] += mine ? colw[j] : 0u;
          if (has && !mine) {
            uint32_t* dst = press + (int64_t)nseg * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) red_add(dst + j, colw[j]);
          }
        }
        if (++since_flush == kFlushEvery) flush_counts();
      }
      flush_warp();
      PHASE_ADD(1);  // records (thread 0's view)
    }
  }
  flush_counts();
  PHASE_ADD(2);  // totals
}

// Split-phase barrier among the CTAs of a team (the pattern of a cooperative-groups grid sync on a per-team counter;
// the cooperative launch guarantees co-residency).  A team of one CTA only needs __syncthreads.
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
struct TeamBarrier {
  unsigned* ctr;
  unsigned target;
  int G;
  __device__ __forceinline__ void arrive() {
    __syncthreads();
    if (G > 1) {
      target += (unsigned)G;
      // release: the CTA's state / pressure writes are ordered before thread 0's increment by the bar.sync above
      if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    }
  }
  __device__ __forceinline__ void wait() {
    if (G > 1) {
      // acquire: poll relaxed, fence once; what another CTA wrote is read afterwards with ld.cg / atomics (L2)
      if (threadIdx.x == 0) {
        while ((int)(ld_relaxed(ctr) - target) < 0) {}
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      __syncthreads();
    }
  }
};

template <bool REPLAY, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_fast(const __grid_constant__ FastArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SharedView sh = carve_shared<REPLAY>(smem_raw, a.table_segs);
  const int tid = threadIdx.x;
  const int G = a.ctas_per_team;
  const int64_t slice = blockIdx.x / G;                // one team per slice of 8 simulations
  const int rank = blockIdx.x % G;
  TeamBarrier bar{a.team_barrier + slice, 0u, G};
  const int64_t ring = a.max_rows_per_day;

  SliceCtx c;
  c.pre = a.pre + slice * 2 * ring;
  c.traj = a.traj ? a.traj + slice * a.R : nullptr;
  c.press = a.pressure + slice * a.n_seg * 4;
  c.gslice = (uint32_t)((a.sim_offset >> 3) + (uint64_t)slice);
  c.sim0 = slice * 8;

  // ---- slice setup: parameters to shared memory, pressure zeroed ---------------------------------------------
  if (tid < 48) {
    const int col = tid >> 3, k = tid & 7;
    const double v = a.params[(c.sim0 + k) + a.p_ld * col];
    SliceParams& par = *sh.par;
    double* dst = col == 0 ? par.alpha : col == 1 ? par.beta : col == 2 ? par.gamma : col == 3 ? par.rho_h
                 : col == 4 ? par.alpha2 : par.rho_i;
    dst[k] = v;
  }
  {
    const int64_t n = a.n_seg * 4, per = (n + G - 1) / G;
    const int64_t lo = rank * per, hi = min(n, lo + per);
    for (int64_t i = lo + tid; i < hi; i += THREADS) __stcg(&c.press[i], 0u);
  }
  bar.arrive();
  bar.wait();

  // ---- initial state (abm_wards.cpp:363-372) into the day-0 buffer -----------------------------------------
  {
    const int64_t n = a.n_init, per = ((n + G - 1) / G + 31) / 32 * 32;
    const int64_t lo = rank * per, hi = min(n, lo + per);
    for (int64_t r = lo + tid; r < hi; r += THREADS) {
      const uint32_t slot = a.init_slot[r];
      if (slot == kNone) continue;  // overwritten by a predecessor before it is ever read
      uint32_t w[4] = {0, 0, 0, 0}, cw[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int64_t sim = c.sim0 + k;
        bool c0 = false, d0 = false;
        if (sim < a.S) {
          const double pc = a.icp[r * a.icp_rs + sim * a.icp_cs], pd = a.idp[r * a.idp_rs + sim * a.idp_cs];
          if (REPLAY) {
            c0 = a.u_icol[r + a.ui_ld * sim] < pc;
            d0 = a.u_idet[r + a.ui_ld * sim] < pd;
          } else {
            c0 = (uint64_t)philox_u32(a.seed, (uint64_t)r, a.sim_offset + sim, kStreamInitCol) < threshold33(pc);
            d0 = (uint64_t)philox_u32(a.seed, (uint64_t)r, a.sim_offset + sim, kStreamInitDet) < threshold33(pd);
          }
        }
        // :372  D0 = d0*c0 is a 0/1 FLAG that the dynamics read as a countdown on day 0:
        //   uncolonised -> D = 0 (U0);  colonised, flag 0 -> D = 0: detected;  flag 1 -> D = 1: pending if
        //   ttd >= 1, else (1 > ttd) it is neither pending nor detected and gets re-tested like D = inf
        const uint32_t code = !c0 ? kCodeU0 : (!d0 ? kCodeZero : (a.ttd >= 1 ? kCodeZero - 1u : kCodeA));
        w[k >> 1] |= code << (16 * (k & 1));
        cw[k >> 1] |= (c0 ? 1u : 0u) << (16 * (k & 1));
      }
      __stcg(&c.pre[slot], make_uint4(w[0], w[1], w[2], w[3]));
      const uint32_t iseg = a.init_seg[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) if (cw[j]) red_add(&c.press[(int64_t)iseg * 4 + j], cw[j]);
    }
  }
  bar.arrive();

  // ---- day loop (abm_wards.cpp:381): one team barrier per day ----------------------------------------------
  PHASE_T0();
  for (int64_t d = 0; d < a.n_days; ++d) {
    bar.wait();
    PHASE_ADD(3);  // team barrier wait
    process_day<REPLAY, THREADS>(PHASE_ARG sh, a, c, d, rank, G);
    bar.arrive();
    PHASE_ADD(4);  // arrive
  }
  bar.wait();
}

// ---- packed trajectory -> the reference's float64 columns ------------------------------------------------------
namespace {
__global__ void k_fast_expand(const uint4* __restrict__ traj, const int64_t* __restrict__ pos_of, int64_t R, int64_t S,
                              int64_t s0, int64_t ns, double* __restrict__ Cout, double* __restrict__ Dout, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_sl = (ns + 7) / 8;  // s0 is a multiple of 8
  if (i >= R * n_sl) return;
  const int64_t r = i % R, sl = i / R;
  const int64_t slice = s0 / 8 + sl;
  const int64_t p = pos_of ? pos_of[r] : r;
  const uint4 v = __ldg(&traj[slice * R + p]);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t s = sl * 8 + k;
    if (s >= ns || s0 + s >= S) break;
    const uint32_t x = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    const uint32_t cls = x >> kClsShift;
    Cout[r + ld * s] = (cls == kClsA || cls >= kClsPn) ? 1.0 : 0.0;
    Dout[r + ld * s] = (cls >= kClsPn) ? (double)((int)kCodeZero - (int)x) : __longlong_as_double(0x7FF0000000000000ll);
  }
}

__global__ void k_fast_export_pressure(const uint32_t* __restrict__ pressure, int64_t n_seg, int64_t S,
                                       int32_t* __restrict__ out, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg * S) return;
  const int64_t seg = i % n_seg, s = i / n_seg;
  const int k = (int)(s & 7);
  const uint32_t w = pressure[((s >> 3) * n_seg + seg) * 4 + (k >> 1)];
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
template <bool REPLAY>
const void* kernel_ptr(int shape) {
  return shape == kShapeWide ? (const void*)k_fast<REPLAY, kWideThreads, 1> : (const void*)k_fast<REPLAY, kTeamThreads, kTeamMinBlocks>;
}
}  // namespace

size_t fast_shared_bytes(bool replay, int table_segs) { return fast_smem_bytes(replay, table_segs); }
int fast_threads(int shape) { return shape == kShapeWide ? kWideThreads : kTeamThreads; }

int fast_max_coresident_ctas(int device, bool replay, int shape, int table_segs) {
  int per_sm = 0, sms = 0;
  const void* k = replay ? kernel_ptr<true>(shape) : kernel_ptr<false>(shape);
  const size_t smem = fast_smem_bytes(replay, table_segs);
  AMRO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AMRO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, fast_threads(shape), smem));
  return per_sm * sms;
}

void fast_launch(cudaStream_t st, const FastArgs& a, bool replay, int shape, int grid) {
  const void* k = replay ? kernel_ptr<true>(shape) : kernel_ptr<false>(shape);
  const size_t smem = fast_smem_bytes(replay, a.table_segs);
  AMRO_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {(void*)&a};
  if (a.ctas_per_team > 1)
    AMRO_CUDA(cudaLaunchCooperativeKernel(k, dim3((unsigned)grid), dim3((unsigned)fast_threads(shape)), params, smem, st));
  else
    AMRO_CUDA(cudaLaunchKernel(k, dim3((unsigned)grid), dim3((unsigned)fast_threads(shape)), params, smem, st));
}

void fast_expand(cudaStream_t st, const uint4* traj, const int64_t* pos_of, const int64_t* day_start, int64_t n_days,
                 int64_t R, int64_t S, int64_t s0, int64_t ns, double* Cout, double* Dout, int64_t ld) {
  (void)day_start; (void)n_days;
  const int64_t n = R * ((ns + 7) / 8);
  if (n <= 0) return;
  k_fast_expand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(traj, pos_of, R, S, s0, ns, Cout, Dout, ld);
}

void fast_export_pressure(cudaStream_t st, const uint32_t* pressure, int64_t n_seg, int64_t S, int32_t* out, int64_t ld) {
  const int64_t n = n_seg * S;
  if (n <= 0) return;
  k_fast_export_pressure<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pressure, n_seg, S, out, ld);
}

}  // namespace amro
