// FAST path: the ward-level colonisation step for a whole run in ONE persistent cooperative kernel.
//
// Replaces the nest  simulate_discrete_model_internal_one -> progress_patients_1_timestep ->
// progress_patients_probability_ward_1_timestep  (src/amro/abm_wards.cpp:322-414 / :219-257 / :70-159) fused
// with total_positive_colonized/_detected (:435-509), for hospitals with unit weights, is_new in {0,1},
// time_to_detect >= 0 and next_day links that connect consecutive days.
//
// State of one (record, simulation): 16 bits = class << 13 | absolute detection day.
//   class 0 U   uncolonised                              (reference: C=0, D=inf)
//         1 A   colonised, no test result coming         (C=1, D=inf)
//         2 U0  uncolonised with a zero countdown        (C=0, D=0: the initial-state quirk of :372)
//         3 Pn  colonised, result pending                (C=1, 0 < D <= ttd)
//         5 B   colonised, detected                      (C=1, D <= 0)
//   The reference decrements D once per day; here a pending/detected record stores Dabs = D + day, which
//   does not change from day to day, and the class bit "detected" is refreshed when the state is written.
//   D of the record of day d after its step is Dabs - d - 1 (fast_expand turns the codes back into C / D).
//   Bit 13 of a code = colonised, bit 15 = detected: the per-day totals are two shifts and a mask.
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
namespace {

constexpr uint32_t kClsU = 0, kClsA = 1, kClsU0 = 2, kClsPn = 3, kClsB = 5;
constexpr uint32_t kClsShift = 13;
constexpr uint32_t kCodeA = kClsA << kClsShift, kCodeU0 = kClsU0 << kClsShift, kCodePn = kClsPn << kClsShift,
                   kCodeB = kClsB << kClsShift;
constexpr int kSlots = 6, kSlotStride = 9, kRowStride = kSlots * kSlotStride;  // padded: conflict-free lookups
constexpr int kFlushEvery = 1024;  // row iterations between flushes of the 16-bit packed per-lane counters

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Barrier among the CTAs of one team (the pattern of a cooperative-groups grid sync, on a per-team counter).
// The cooperative launch guarantees the CTAs are co-resident.
__device__ __forceinline__ void team_barrier(unsigned* ctr, unsigned& target, int ctas) {
  __syncthreads();
  if (ctas > 1) {
    target += (unsigned)ctas;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      while ((int)(ld_acquire(ctr) - target) < 0) {}
      __threadfence();
    }
    __syncthreads();
  }
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

template <bool REPLAY>
struct Shared;
template <>
struct Shared<false> {
  uint2 hi[kTableSegs * 2 * kRowStride];    // {T(P) >> 16, T(P*rho) >> 16} at [(seg*2+h)*54 + class*9 + sim]
  ushort2 lo[kTableSegs * 2 * kRowStride];  // low 16 bits, read only on a tie
  SliceParams par;
  int cnt[16];
};
template <>
struct Shared<true> {
  double P[kTableSegs * 2][3][8];
  SliceParams par;
  int cnt[16];
};

struct DayConst {        // uniform per day
  uint32_t newdet;       // code a newly tested-positive record gets: D = ttd after the step
  uint32_t quirk;        // code of a U0 record that gets colonised: D = -1 after the step
  uint32_t bump_lim;     // a pending record with Dabs <= bump_lim is detected from tomorrow on
  bool test_h, test_a;
};

__device__ __forceinline__ uint32_t next_code(uint32_t v, uint32_t cls, bool col, bool dh, const DayConst& dc, bool has_u0) {
  const bool keep = v >= kCodePn;                                  // pending / detected: countdown runs on (:131-132)
  const bool bump = (v - kCodePn) <= dc.bump_lim;                  // pending whose D reaches 0 with this step
  const uint32_t v2 = v + (bump ? (kCodeB - kCodePn) : 0u);
  uint32_t r1 = dh ? dc.newdet : kCodeA;                           // :134-138 / :146-150
  if (has_u0) r1 = (cls == kClsU0) ? dc.quirk : r1;
  const uint32_t r2 = keep ? v2 : r1;
  return col ? r2 : 0u;                                            // :153-155
}

// One record, eight simulations.  Returns the packed post-step codes.
template <bool REPLAY, bool HAS_U0>
__device__ __forceinline__ uint4 step_record(const Shared<REPLAY>& sh, const FastArgs& a, const uint4 pre, int ti, int h,
                                             int64_t orig, uint32_t gslice, int64_t sim0, uint32_t key0, uint32_t key1,
                                             const DayConst& dc) {
  const uint32_t prew[4] = {pre.x, pre.y, pre.z, pre.w};
  uint32_t postw[4];
  if constexpr (!REPLAY) {
    const uint2* tab = sh.hi + (ti * 2 + h) * kRowStride;
    const uint4 rnd = philox4x32_10((uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), gslice, kStreamStep, key0, key1);
    const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
    bool tie = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t v = (k & 1) ? (prew[k >> 1] >> 16) : (prew[k >> 1] & 0xFFFFu);
      const uint32_t c = (k & 1) ? (rw[k >> 1] >> 16) : (rw[k >> 1] & 0xFFFFu);
      const uint32_t cls = v >> kClsShift;
      const uint2 T = tab[cls * kSlotStride + k];
      tie |= (c == T.x) | (c == T.y);
      const uint32_t o = next_code(v, cls, c < T.x, c < T.y, dc, HAS_U0);
      if (k & 1) postw[k >> 1] |= o << 16; else postw[k >> 1] = o;
    }
    if (tie) {  // probability 2^-15 per update: redo the record with the refinement chunks (full 32-bit uniforms)
      const ushort2* tlo = sh.lo + (ti * 2 + h) * kRowStride;
      const uint4 rf = philox4x32_10((uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), gslice, kStreamStep + 1u, key0, key1);
      const uint32_t fw[4] = {rf.x, rf.y, rf.z, rf.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t v = (prew[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
        const uint32_t c = (rw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
        const uint32_t f = (fw[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
        const uint64_t u = ((uint64_t)c << 16) | f;
        const uint32_t cls = v >> kClsShift;
        const uint2 Th = tab[cls * kSlotStride + k];
        const ushort2 Tl = tlo[cls * kSlotStride + k];
        const bool col = u < (((uint64_t)Th.x << 16) | Tl.x), dh = u < (((uint64_t)Th.y << 16) | Tl.y);
        const uint32_t o = next_code(v, cls, col, dh, dc, true);
        if (k & 1) postw[k >> 1] |= o << 16; else postw[k >> 1] = o;
      }
    }
  } else {
    const double* rho_tab = h ? sh.par.rho_i : sh.par.rho_h;
    const bool flag = h ? dc.test_a : dc.test_h;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t v = (k & 1) ? (prew[k >> 1] >> 16) : (prew[k >> 1] & 0xFFFFu);
      const uint32_t cls = v >> kClsShift;
      const int pc = (cls == kClsB) ? 2 : (int)(cls & 1u);
      const int64_t sim = sim0 + k;
      const double P = sh.P[ti * 2 + h][pc][k];
      bool col = false, dh = false;
      if (sim < a.S) {
        if (a.p_out) a.p_out[orig + a.R * sim] = P;
        col = a.u_col[orig + a.u_ld * sim] < P;                                              // :116
        if (col && cls < kClsU0 && flag) dh = a.u_det[orig + a.u_ld * sim] <= rho_tab[k];    // :134 / :146
      }
      const uint32_t o = next_code(v, cls, col, dh, dc, true);
      if (k & 1) postw[k >> 1] |= o << 16; else postw[k >> 1] = o;
    }
  }
  return make_uint4(postw[0], postw[1], postw[2], postw[3]);
}

}  // namespace

template <bool REPLAY>
__global__ void __launch_bounds__(kFastThreads, kFastMinBlocks) k_fast(const FastArgs a) {
  __shared__ Shared<REPLAY> sh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int team = blockIdx.x / a.ctas_per_team, rank = blockIdx.x % a.ctas_per_team;
  if (team >= a.teams) return;
  unsigned bar_target = 0;
  unsigned* bar = a.team_barrier + team;
  const int G = a.ctas_per_team;
  const uint32_t key0 = (uint32_t)a.seed, key1 = (uint32_t)(a.seed >> 32);
  const int64_t ring = a.max_rows_per_day;

  if (tid < 16) sh.cnt[tid] = 0;

  for (int64_t slice = team; slice < a.n_slices; slice += a.teams) {
    const int64_t set = a.sets_by_team ? (int64_t)team : slice;
    uint4* pre = a.pre + set * 2 * ring;
    uint4* traj = a.traj ? a.traj + slice * a.R : nullptr;
    uint32_t* press = a.pressure + set * a.n_seg * 4;
    const uint32_t gslice = (uint32_t)((a.sim_offset >> 3) + (uint64_t)slice);
    const int64_t sim0 = slice * 8;

    // ---- slice setup: parameters to shared memory, pressure zeroed -----------------------------------------
    __syncthreads();
    if (tid < 48) {
      const int c = tid >> 3, k = tid & 7;
      const double v = a.params[(sim0 + k) + a.p_ld * c];
      double* dst = c == 0 ? sh.par.alpha : c == 1 ? sh.par.beta : c == 2 ? sh.par.gamma : c == 3 ? sh.par.rho_h
                   : c == 4 ? sh.par.alpha2 : sh.par.rho_i;
      dst[k] = v;
    }
    {
      const int64_t n = a.n_seg * 4, per = (n + G - 1) / G;
      const int64_t lo = rank * per, hi = min(n, lo + per);
      for (int64_t i = lo + tid; i < hi; i += kFastThreads) __stcg(&press[i], 0u);
    }
    team_barrier(bar, bar_target, G);

    // ---- initial state (abm_wards.cpp:363-372) into the day-0 buffer -----------------------------------------
    {
      const int64_t n = a.n_init, per = ((n + G - 1) / G + 31) / 32 * 32;
      const int64_t lo = rank * per, hi = min(n, lo + per);
      for (int64_t r = lo + tid; r < hi; r += kFastThreads) {
        const uint32_t slot = a.init_slot[r];
        if (slot == kNone) continue;  // overwritten by a predecessor before it is ever read
        uint32_t w[4] = {0, 0, 0, 0}, cw[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int64_t sim = sim0 + k;
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
          const uint32_t code = !c0 ? kCodeU0 : (!d0 ? (kCodeB | 0u) : (a.ttd >= 1 ? (kCodePn | 1u) : kCodeA));
          w[k >> 1] |= code << (16 * (k & 1));
          cw[k >> 1] |= (c0 ? 1u : 0u) << (16 * (k & 1));
        }
        __stcg(&pre[slot], make_uint4(w[0], w[1], w[2], w[3]));
        const uint32_t iseg = a.init_seg[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (cw[j]) red_add(&press[(int64_t)iseg * 4 + j], cw[j]);
      }
    }
    team_barrier(bar, bar_target, G);

    // ---- day loop (abm_wards.cpp:381) -------------------------------------------------------------------
    for (int64_t d = 0; d < a.n_days; ++d) {
      const int64_t da = a.day_start[d], db = a.day_start[d + 1];
      const int64_t n = db - da, per = ((n + G - 1) / G + 31) / 32 * 32;
      const int64_t lo = da + rank * per, hi = min(db, lo + per);
      DayConst dc;
      dc.test_h = ((uint64_t)d % a.tsh) == 0;  // :391-392
      dc.test_a = ((uint64_t)d % a.tsa) == 0;
      dc.newdet = (a.ttd == 0 ? kCodeB : kCodePn) | (uint32_t)(d + 1 + a.ttd);
      dc.quirk = kCodeB | (uint32_t)d;
      dc.bump_lim = (uint32_t)(d + 1);
      const int64_t seg0 = a.day_seg_start[d];
      const uint4* cur = pre + (d & 1) * ring;
      uint4* nxt = pre + ((d + 1) & 1) * ring;
      uint32_t ccol[4] = {0, 0, 0, 0}, cdet[4] = {0, 0, 0, 0};
      int since_flush = 0;

      auto flush_counts = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, ccol[j]);
          const uint32_t e = __reduce_add_sync(0xFFFFFFFFu, cdet[j]);
          if (lane == 0) {
            if (c & 0xFFFFu) atomicAdd(&sh.cnt[2 * j], (int)(c & 0xFFFFu));
            if (c >> 16) atomicAdd(&sh.cnt[2 * j + 1], (int)(c >> 16));
            if (e & 0xFFFFu) atomicAdd(&sh.cnt[8 + 2 * j], (int)(e & 0xFFFFu));
            if (e >> 16) atomicAdd(&sh.cnt[8 + 2 * j + 1], (int)(e >> 16));
          }
          ccol[j] = 0;
          cdet[j] = 0;
        }
        since_flush = 0;
      };

      if (lo < hi) {
        const int64_t seg_first = seg0 + (a.meta_info[lo] & kInfoSegMask);
        const int64_t seg_last = seg0 + (a.meta_info[hi - 1] & kInfoSegMask);
        for (int64_t sb = seg_first; sb <= seg_last; sb += kTableSegs) {
          const int64_t se = min(seg_last + 1, sb + kTableSegs);
          // ---- threshold tables of ward-days [sb, se) ------------------------------------------------------
          __syncthreads();
          for (int t = tid; t < (int)(se - sb) * 16; t += kFastThreads) {
            const int si = t >> 4, h = (t >> 3) & 1, k = t & 7;
            const int64_t seg = sb + si;
            const uint32_t pk = __ldcg(&press[seg * 4 + (k >> 1)]);
            const uint32_t cnt = (k & 1) ? (pk >> 16) : (pk & 0xFFFFu);
            // :98-99  F = (beta / N) * sum(W); with unit weights sum(W) is the exact integer count
            const double F = __dmul_rn(__ddiv_rn(sh.par.beta[k], a.seg_N[seg]), (double)cnt);
            const bool flag = h ? dc.test_a : dc.test_h;
            const double rho = h ? sh.par.rho_i[k] : sh.par.rho_h[k];
#pragma unroll
            for (int pc = 0; pc < 3; ++pc) {
              const double P = class_probability(pc, h, F, sh.par.alpha[k], sh.par.alpha2[k], sh.par.gamma[k]);
              if constexpr (REPLAY) {
                sh.P[si * 2 + h][pc][k] = P;
              } else {
                const uint64_t T = threshold33(P);
                const uint64_t TD = flag ? threshold33(__dmul_rn(clamp01(P), rho)) : 0ull;
                const uint2 e_test = make_uint2((uint32_t)(T >> 16), (uint32_t)(TD >> 16));
                const uint2 e_none = make_uint2((uint32_t)(T >> 16), 0u);
                const ushort2 l_test = make_ushort2((unsigned short)(T & 0xFFFFu), (unsigned short)(TD & 0xFFFFu));
                const ushort2 l_none = make_ushort2((unsigned short)(T & 0xFFFFu), 0);
                const int row = (si * 2 + h) * kRowStride + k;
                // U and A are tested today; U0, Pn and B keep their countdown (no new test)
                if (pc == 0) { sh.hi[row + kClsU * kSlotStride] = e_test; sh.lo[row + kClsU * kSlotStride] = l_test;
                               sh.hi[row + kClsU0 * kSlotStride] = e_none; sh.lo[row + kClsU0 * kSlotStride] = l_none; }
                if (pc == 1) { sh.hi[row + kClsA * kSlotStride] = e_test; sh.lo[row + kClsA * kSlotStride] = l_test;
                               sh.hi[row + kClsPn * kSlotStride] = e_none; sh.lo[row + kClsPn * kSlotStride] = l_none; }
                if (pc == 2) { sh.hi[row + kClsB * kSlotStride] = e_none; sh.lo[row + kClsB * kSlotStride] = l_none; }
              }
            }
          }
          __syncthreads();
          // ---- the records of those ward-days that belong to this CTA ----------------------------------------
          const int64_t rb = max(lo, a.seg_row_start[sb]), re = min(hi, a.seg_row_start[se]);
          for (int64_t base = rb; base < re; base += kFastThreads) {  // CTA-uniform trip count (warp collectives below)
            const int64_t p = base + tid;
            const bool valid = p < re;
            uint32_t colw[4] = {0, 0, 0, 0}, detw[4] = {0, 0, 0, 0};
            uint32_t nseg = kNone;
            if (valid) {
              const uint32_t info = __ldg(&a.meta_info[p]);
              const uint32_t next = __ldg(&a.meta_next[p]);
              nseg = __ldg(&a.meta_nseg[p]);
              uint4 prev = __ldcg(&cur[p - da]);
              const int64_t orig = a.orig_row ? a.orig_row[p] : p;
              const int h = (int)(info >> kInfoHShift);
              const int ti = (int)((int64_t)(info & kInfoSegMask) + seg0 - sb);
              if (((info >> kInfoSrcShift) & 3u) == kSrcNone) prev = make_uint4(0u, 0u, 0u, 0u);  // (C=0, D=inf) x 8, :359-360
              uint4 post;
              if (d == 0) post = step_record<REPLAY, true>(sh, a, prev, ti, h, orig, gslice, sim0, key0, key1, dc);
              else post = step_record<REPLAY, false>(sh, a, prev, ti, h, orig, gslice, sim0, key0, key1, dc);
              if (next != kNone) __stcg(&nxt[next], post);     // :402-407 the successor starts from this state
              if (traj) __stcg(&traj[p], post);
              const uint32_t pw[4] = {post.x, post.y, post.z, post.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                colw[j] = (pw[j] >> 13) & 0x00010001u;
                detw[j] = (pw[j] >> 15) & 0x00010001u;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { ccol[j] += colw[j]; cdet[j] += detw[j]; }

            // ---- feed the successor's ward-day (its sum(W) of tomorrow, :95-99) --------------------------------
            // Most lanes of a warp feed the same ward-day: they are summed with a warp reduction and added by one
            // lane; the rest (transfers) add on their own.
            const bool has = nseg != kNone;
            const unsigned hm = __ballot_sync(0xFFFFFFFFu, has);
            if (hm) {
              const int leader = __ffs(hm) - 1;
              const uint32_t lseg = __shfl_sync(0xFFFFFFFFu, nseg, leader);
              const bool mine = nseg == lseg;
              uint32_t v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = __reduce_add_sync(0xFFFFFFFFu, mine ? colw[j] : 0u);
              if (has && (lane == leader || !mine)) {
                uint32_t* dst = press + (int64_t)nseg * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) red_add(dst + j, mine ? v[j] : colw[j]);
              }
            }
            if (++since_flush == kFlushEvery) flush_counts();
          }
        }
      }
      // ---- per-day totals of this CTA (abm_wards.cpp:445-459 fused) -------------------------------------------
      flush_counts();
      __syncthreads();
      if (tid < 16) {
        const int v = sh.cnt[tid];
        if (v != 0 && d < a.counts_ld) {
          int32_t* dst = (tid < 8) ? a.counts_col : a.counts_det;
          if (dst) atomicAdd(&dst[d + a.counts_ld * (sim0 + (tid & 7))], v);
        }
        sh.cnt[tid] = 0;
      }
      team_barrier(bar, bar_target, G);
    }
  }
}

// ---- packed trajectory -> the reference's float64 columns ------------------------------------------------------
namespace {
__global__ void k_fast_expand(const uint4* __restrict__ traj, const int64_t* __restrict__ pos_of,
                              const int64_t* __restrict__ day_start, int64_t n_days, int64_t R, int64_t S, int64_t s0,
                              int64_t ns, double* __restrict__ Cout, double* __restrict__ Dout, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_sl = (ns + 7) / 8;  // s0 is a multiple of 8
  if (i >= R * n_sl) return;
  const int64_t r = i % R, sl = i / R;
  const int64_t slice = s0 / 8 + sl;
  const int64_t p = pos_of ? pos_of[r] : r;
  // day of the record: last d with day_start[d] <= p
  int64_t lo = 0, hi = n_days;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (day_start[mid] <= p) lo = mid; else hi = mid;
  }
  const int64_t day = lo;
  const uint4 v = __ldg(&traj[slice * R + p]);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t s = sl * 8 + k;
    if (s >= ns || s0 + s >= S) break;
    const uint32_t x = (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    const uint32_t cls = x >> kClsShift;
    const bool finite = cls >= kClsPn;
    Cout[r + ld * s] = (double)((x >> 13) & 1u);
    Dout[r + ld * s] = finite ? (double)((int64_t)(x & 0x1FFFu) - day - 1) : __longlong_as_double(0x7FF0000000000000ll);
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

int fast_max_coresident_ctas(int device, bool replay) {
  int per_sm = 0, sms = 0;
  AMRO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (replay) AMRO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fast<true>, kFastThreads, 0));
  else AMRO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fast<false>, kFastThreads, 0));
  return per_sm * sms;
}

void fast_launch(cudaStream_t st, const FastArgs& a, bool replay, int grid) {
  void* params[] = {(void*)&a};
  if (replay) AMRO_CUDA(cudaLaunchCooperativeKernel((void*)k_fast<true>, dim3(grid), dim3(kFastThreads), params, 0, st));
  else AMRO_CUDA(cudaLaunchCooperativeKernel((void*)k_fast<false>, dim3(grid), dim3(kFastThreads), params, 0, st));
}

void fast_expand(cudaStream_t st, const uint4* traj, const int64_t* pos_of, const int64_t* day_start, int64_t n_days,
                 int64_t R, int64_t S, int64_t s0, int64_t ns, double* Cout, double* Dout, int64_t ld) {
  const int64_t n = R * ((ns + 7) / 8);
  if (n <= 0) return;
  k_fast_expand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(traj, pos_of, day_start, n_days, R, S, s0, ns, Cout, Dout, ld);
}

void fast_export_pressure(cudaStream_t st, const uint32_t* pressure, int64_t n_seg, int64_t S, int32_t* out, int64_t ld) {
  const int64_t n = n_seg * S;
  if (n <= 0) return;
  k_fast_export_pressure<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pressure, n_seg, S, out, ld);
}

}  // namespace amro
