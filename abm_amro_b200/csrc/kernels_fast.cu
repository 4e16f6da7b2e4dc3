// FAST path: the ward-level colonisation step for a whole run in ONE persistent cooperative kernel.
//
// Replaces the nest  simulate_discrete_model_internal_one -> progress_patients_1_timestep ->
// progress_patients_probability_ward_1_timestep  (src/amro/abm_wards.cpp:322-414 / :219-257 / :70-159) fused
// with total_positive_colonized/_detected (:435-509) for hospitals with unit weights and is_new in {0,1}.
//
// Data layout (HBM):
//   state       [slice][slot]  one 16-byte word = 8 simulations x int16 (2*D + C; D = 16383 is +inf) per
//               patient-day record; slot = record position (trajectory) or a two-day ring
//   pressure    [slice][segment][8] int32: colonised patients a ward-day STARTS with = the reference's
//               sum(W) (:99).  It is accumulated by the previous day's pass (each record adds its new C to
//               the ward-day of its next_day successor), so a day needs ONE pass over its records.
//   meta_*      three 32-bit words per record: predecessor link, (is_new, ward-day index), successor's
//               ward-day.  Shared by every simulation.
// Work decomposition: a TEAM of ctas_per_team CTAs owns one slice (8 simulations) for every day; lanes map
// to consecutive records, so state / uniform loads are 128-bit and coalesced across the warp.  Teams never
// talk to each other; inside a team the CTAs split each day's records and meet at one barrier per day
// (records of day d+1 read the state day d wrote).  Per ward-day, per simulation, the six transition
// probabilities a record can have (C in {0,1} x detected x is_new) are evaluated in fp64 exactly as the
// reference evaluates them and turned into integer thresholds kept in shared memory.
#include <cooperative_groups.h>

#include "common.hpp"
#include "kernels.hpp"
#include "rng.cuh"
#include "topology.hpp"

namespace amro {
namespace {

constexpr int kInf0 = 2 * kInfCode;      // packed (C=0, D=inf)
constexpr int kInf1 = 2 * kInfCode + 1;  // packed (C=1, D=inf)
constexpr int kFlushEvery = 512;         // row iterations between flushes of the 16-bit packed counters

// ---- memory helpers: state crosses CTAs of a team through L2, never through a (non-coherent) L1 ------------
__device__ __forceinline__ uint4 ld_cg(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(uint4* p, const uint4& v) { __stcg(p, v); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Barrier among the CTAs of one team (same pattern as a cooperative-groups grid sync, on a per-team counter).
// Cooperative launch guarantees the CTAs are co-resident.
__device__ __forceinline__ void team_barrier(unsigned* ctr, unsigned& target, int ctas) {
  __syncthreads();
  if (ctas > 1) {
    target += (unsigned)ctas;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(ctr, 1u);
      while ((int)(ld_acquire(ctr) - target) < 0) { __nanosleep(32); }
      __threadfence();
    }
    __syncthreads();
  }
}

struct SliceParams {  // the 8 simulations' parameter rows (abm_wards.cpp:79-84)
  double alpha[8], beta[8], gamma[8], rho_h[8], alpha2[8], rho_i[8];
};

// The transition probability of class cls (0: C=0, 1: C=1 undetected, 2: C=1 detected) for is_new = h in a
// ward-day with force of infection F, for unit weight -- the reference's expression order, one IEEE
// operation at a time (:95-112).
__device__ __forceinline__ double class_probability(int cls, int h, double F, double alpha, double alpha2, double gamma) {
  const double W = cls ? 1.0 : 0.0;
  const double det = (cls == 2) ? 1.0 : 0.0;
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
  uint2 hi[kTableSegs][2][3][8];     // {T(P) >> 16, T(P*rho) >> 16} per (ward-day, is_new, class, sim)
  ushort2 lo[kTableSegs][2][3][8];   // low 16 bits, read only on a tie
  SliceParams par;
  int cnt[16];
};
template <>
struct Shared<true> {
  double P[kTableSegs][2][3][8];
  SliceParams par;
  int cnt[16];
};


}  // namespace

template <bool REPLAY>
__global__ void __launch_bounds__(kFastThreads) k_fast(const FastArgs a) {
  __shared__ Shared<REPLAY> sh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int team = blockIdx.x / a.ctas_per_team, rank = blockIdx.x % a.ctas_per_team;
  if (team >= a.teams) return;
  unsigned bar_target = 0;
  unsigned* bar = a.team_barrier + team;
  const int G = a.ctas_per_team;
  const uint32_t key0 = (uint32_t)a.seed, key1 = (uint32_t)(a.seed >> 32);
  const int ttd2 = 2 * a.ttd + 1;
  const int64_t scratch_stride_state = a.n_slots, scratch_stride_init = a.n_init, scratch_stride_press = a.n_seg * 8;

  if (tid < 16) sh.cnt[tid] = 0;

  for (int64_t slice = team; slice < a.n_slices; slice += a.teams) {
    const int64_t sidx = a.ring ? (int64_t)team : slice;  // scratch is per team when nothing is kept per slice
    uint4* st = a.state + sidx * scratch_stride_state;
    uint4* init = a.init_state + sidx * scratch_stride_init;
    int32_t* press = a.pressure + sidx * scratch_stride_press;
    const uint32_t gslice = (uint32_t)((a.sim_offset >> 3) + (uint64_t)slice);
    const int64_t sim0 = slice * 8;

    // ---- slice setup: parameters to shared memory, pressure zeroed -------------------------------------
    __syncthreads();
    if (tid < 48) {
      const int c = tid >> 3, k = tid & 7;
      const double v = a.params[(sim0 + k) + a.p_ld * c];
      double* dst = c == 0 ? sh.par.alpha : c == 1 ? sh.par.beta : c == 2 ? sh.par.gamma : c == 3 ? sh.par.rho_h
                   : c == 4 ? sh.par.alpha2 : sh.par.rho_i;
      dst[k] = v;
    }
    {
      const int64_t n = a.n_seg * 8, per = (n + G - 1) / G;
      const int64_t lo = rank * per, hi = min(n, lo + per);
      for (int64_t i = lo + tid; i < hi; i += kFastThreads) __stcg(&press[i], 0);
    }
    team_barrier(bar, bar_target, G);

    // ---- initial state (abm_wards.cpp:363-372) ---------------------------------------------------------
    {
      const int64_t n = a.n_init, per = ((n + G - 1) / G + 31) / 32 * 32;
      const int64_t lo = rank * per, hi = min(n, lo + per);
      for (int64_t r = lo + tid; r < hi; r += kFastThreads) {
        uint32_t w[4] = {0, 0, 0, 0};
        const uint32_t iseg = a.init_seg[r];
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
          // :372  D0 = d0*c0 is a 0/1 FLAG that the dynamics later read as a countdown: packed 2*D + C
          const uint32_t s = c0 ? (d0 ? 3u : 1u) : 0u;
          w[k >> 1] |= s << (16 * (k & 1));
          if (c0 && iseg != kNone) atomicAdd(&press[(int64_t)iseg * 8 + k], 1);
        }
        st_cg(&init[r], make_uint4(w[0], w[1], w[2], w[3]));
      }
    }
    team_barrier(bar, bar_target, G);

    // ---- day loop (abm_wards.cpp:381) -------------------------------------------------------------------
    for (int64_t d = 0; d < a.n_days; ++d) {
      const int64_t da = a.day_start[d], db = a.day_start[d + 1];
      const int64_t n = db - da, per = ((n + G - 1) / G + 31) / 32 * 32;
      const int64_t lo = da + rank * per, hi = min(db, lo + per);
      const bool test_h = ((uint64_t)d % a.tsh) == 0, test_a = ((uint64_t)d % a.tsa) == 0;  // :391-392
      const int64_t seg0 = a.day_seg_start[d];
      const int64_t ring_cur = a.ring ? (d & 1) * a.max_rows_per_day - da : 0;
      const int64_t ring_prev = (a.ring && d > 0) ? ((d - 1) & 1) * a.max_rows_per_day - a.day_start[d - 1] : 0;
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
        const int64_t seg_first = seg0 + (a.meta_info[lo] & 0x7FFFFFFFu);
        const int64_t seg_last = seg0 + (a.meta_info[hi - 1] & 0x7FFFFFFFu);
        for (int64_t sb = seg_first; sb <= seg_last; sb += kTableSegs) {
          const int64_t se = min(seg_last + 1, sb + kTableSegs);
          // ---- threshold tables of ward-days [sb, se) ------------------------------------------------------
          __syncthreads();
          for (int t = tid; t < (int)(se - sb) * 16; t += kFastThreads) {
            const int si = t >> 4, h = (t >> 3) & 1, k = t & 7;
            const int64_t seg = sb + si;
            const int cnt = __ldcg(&press[seg * 8 + k]);
            // :98-99  F = (beta / N) * sum(W); with unit weights sum(W) is the exact integer count
            const double F = __dmul_rn(__ddiv_rn(sh.par.beta[k], a.seg_N[seg]), (double)cnt);
            const bool flag = h ? test_a : test_h;
            const double rho = h ? sh.par.rho_i[k] : sh.par.rho_h[k];
#pragma unroll
            for (int cls = 0; cls < 3; ++cls) {
              const double P = class_probability(cls, h, F, sh.par.alpha[k], sh.par.alpha2[k], sh.par.gamma[k]);
              if constexpr (REPLAY) {
                sh.P[si][h][cls][k] = P;
              } else {
                const uint64_t T = threshold33(P);
                const uint64_t TD = flag ? threshold33(__dmul_rn(clamp01(P), rho)) : 0ull;
                sh.hi[si][h][cls][k] = make_uint2((uint32_t)(T >> 16), (uint32_t)(TD >> 16));
                sh.lo[si][h][cls][k] = make_ushort2((unsigned short)(T & 0xFFFFu), (unsigned short)(TD & 0xFFFFu));
              }
            }
          }
          __syncthreads();
          // ---- the records of those ward-days that belong to this CTA ----------------------------------------
          const int64_t rb = max(lo, a.seg_row_start[sb]), re = min(hi, a.seg_row_start[se]);
          for (int64_t base = rb; base < re; base += kFastThreads) {  // CTA-uniform trip count: warp collectives below
            const int64_t p = base + tid;
            const bool valid = p < re;
            uint32_t colw[4] = {0, 0, 0, 0}, detw[4] = {0, 0, 0, 0};
            uint32_t nseg = kNone;
            if (valid) {
            const uint32_t info = a.meta_info[p], src = a.meta_src[p];
            nseg = a.meta_nseg[p];
            const int h = (int)(info >> 31);
            const int ti = (int)((int64_t)(info & 0x7FFFFFFFu) + seg0 - sb);
            const uint32_t kind = src >> kSrcShift, val = src & kSrcMask;
            uint4 pre;
            if (kind == kSrcPrev) pre = ld_cg(&st[(p - val) + ring_prev]);
            else if (kind == kSrcInit) pre = ld_cg(&init[val]);
            else pre = make_uint4(0x7FFE7FFEu, 0x7FFE7FFEu, 0x7FFE7FFEu, 0x7FFE7FFEu);  // (C=0, D=inf) x 8  (:359-360)
            const int64_t orig = a.orig_row ? a.orig_row[p] : p;
            const uint32_t prew[4] = {pre.x, pre.y, pre.z, pre.w};
            uint32_t postw[4];

            if constexpr (!REPLAY) {
              const uint4 rnd = philox4x32_10((uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), gslice, kStreamStep, key0, key1);
              const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
              uint32_t ties = 0;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int s = (k & 1) ? ((int)prew[k >> 1] >> 16) : (int)(short)(prew[k >> 1] & 0xFFFFu);
                const uint32_t c = (k & 1) ? (rw[k >> 1] >> 16) : (rw[k >> 1] & 0xFFFFu);
                const int cls = (s & 1) ? ((s <= 1) ? 2 : 1) : 0;
                const uint2 T = sh.hi[ti][h][cls][k];
                const bool col = c < T.x, dh = c < T.y;
                ties |= (uint32_t)((c == T.x) | (c == T.y)) << k;
                const bool pend = s <= ttd2;
                const int sp = !col ? kInf0 : (pend ? ((s | 1) - 2) : (dh ? ttd2 : kInf1));
                const uint32_t half = (uint32_t)sp & 0xFFFFu;
                const uint32_t cb = col ? 1u : 0u, db_ = (col && sp <= 1) ? 1u : 0u;
                if (k & 1) { postw[k >> 1] |= half << 16; colw[k >> 1] |= cb << 16; detw[k >> 1] |= db_ << 16; }
                else { postw[k >> 1] = half; colw[k >> 1] = cb; detw[k >> 1] = db_; }
              }
              if (ties) {  // probability 2^-15 per update: resolve with the refinement chunks, all 8 sims again
                const uint4 rf = philox4x32_10((uint32_t)orig, (uint32_t)((uint64_t)orig >> 32), gslice, kStreamStep + 1u, key0, key1);
                const uint32_t fw[4] = {rf.x, rf.y, rf.z, rf.w};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const int s = (k & 1) ? ((int)prew[k >> 1] >> 16) : (int)(short)(prew[k >> 1] & 0xFFFFu);
                  const uint32_t c = (k & 1) ? (rw[k >> 1] >> 16) : (rw[k >> 1] & 0xFFFFu);
                  const uint32_t f = (k & 1) ? (fw[k >> 1] >> 16) : (fw[k >> 1] & 0xFFFFu);
                  const uint64_t u = ((uint64_t)c << 16) | f;
                  const int cls = (s & 1) ? ((s <= 1) ? 2 : 1) : 0;
                  const uint2 Th = sh.hi[ti][h][cls][k];
                  const ushort2 Tl = sh.lo[ti][h][cls][k];
                  const bool col = u < (((uint64_t)Th.x << 16) | Tl.x), dh = u < (((uint64_t)Th.y << 16) | Tl.y);
                  const bool pend = s <= ttd2;
                  const int sp = !col ? kInf0 : (pend ? ((s | 1) - 2) : (dh ? ttd2 : kInf1));
                  const uint32_t half = (uint32_t)sp & 0xFFFFu;
                  const uint32_t cb = col ? 1u : 0u, db_ = (col && sp <= 1) ? 1u : 0u;
                  if (k & 1) { postw[k >> 1] |= half << 16; colw[k >> 1] |= cb << 16; detw[k >> 1] |= db_ << 16; }
                  else { postw[k >> 1] = half; colw[k >> 1] = cb; detw[k >> 1] = db_; }
                }
              }
            } else {
              const double* rho_tab = h ? sh.par.rho_i : sh.par.rho_h;
              const bool flag = h ? test_a : test_h;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int s = (k & 1) ? ((int)prew[k >> 1] >> 16) : (int)(short)(prew[k >> 1] & 0xFFFFu);
                const int64_t sim = sim0 + k;
                const int cls = (s & 1) ? ((s <= 1) ? 2 : 1) : 0;
                const double P = sh.P[ti][h][cls][k];
                bool col = false, dh = false;
                const bool pend = s <= ttd2;
                if (sim < a.S) {
                  if (a.p_out) a.p_out[orig + a.R * sim] = P;
                  col = a.u_col[orig + a.u_ld * sim] < P;                                  // :116
                  if (col && !pend && flag) dh = a.u_det[orig + a.u_ld * sim] <= rho_tab[k];  // :134 / :146
                }
                const int sp = !col ? kInf0 : (pend ? ((s | 1) - 2) : (dh ? ttd2 : kInf1));
                const uint32_t half = (uint32_t)sp & 0xFFFFu;
                const uint32_t cb = col ? 1u : 0u, db_ = (col && sp <= 1) ? 1u : 0u;
                if (k & 1) { postw[k >> 1] |= half << 16; colw[k >> 1] |= cb << 16; detw[k >> 1] |= db_ << 16; }
                else { postw[k >> 1] = half; colw[k >> 1] = cb; detw[k >> 1] = db_; }
              }
            }

            st_cg(&st[p + ring_cur], make_uint4(postw[0], postw[1], postw[2], postw[3]));
            }  // valid
#pragma unroll
            for (int j = 0; j < 4; ++j) { ccol[j] += colw[j]; cdet[j] += detw[j]; }

            // ---- feed the successor's ward-day (its sum(W) of tomorrow, :95-99) --------------------------------
            // Most lanes of a warp feed the same ward-day: they are summed with a warp reduction and added by
            // one lane; the rest (transfers) add on their own.
            constexpr unsigned act = 0xFFFFFFFFu;
            const bool has = nseg != kNone;
            const unsigned hm = __ballot_sync(act, has);
            if (hm) {
              const int leader = __ffs(hm) - 1;
              const uint32_t lseg = __shfl_sync(act, nseg, leader);
              const bool mine = has && nseg == lseg;
              uint32_t sum[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) sum[j] = __reduce_add_sync(act, mine ? colw[j] : 0u);
              if (lane == leader) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (sum[j] & 0xFFFFu) atomicAdd(&press[(int64_t)lseg * 8 + 2 * j], (int)(sum[j] & 0xFFFFu));
                  if (sum[j] >> 16) atomicAdd(&press[(int64_t)lseg * 8 + 2 * j + 1], (int)(sum[j] >> 16));
                }
              } else if (has && !mine) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (colw[j] & 0xFFFFu) atomicAdd(&press[(int64_t)nseg * 8 + 2 * j], 1);
                  if (colw[j] >> 16) atomicAdd(&press[(int64_t)nseg * 8 + 2 * j + 1], 1);
                }
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
__global__ void k_fast_expand(const uint4* __restrict__ state, int64_t n_slots, const int64_t* __restrict__ pos_of,
                              int64_t R, int64_t S, int64_t s0, int64_t ns, double* __restrict__ Cout,
                              double* __restrict__ Dout, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_sl = (ns + 7) / 8;  // s0 is a multiple of 8
  if (i >= R * n_sl) return;
  const int64_t r = i % R, sl = i / R;
  const int64_t slice = s0 / 8 + sl;
  const int64_t p = pos_of ? pos_of[r] : r;
  const uint4 v = __ldg(&state[slice * n_slots + p]);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t s = sl * 8 + k;
    if (s >= ns || s0 + s >= S) break;
    const int x = (k & 1) ? ((int)w[k >> 1] >> 16) : (int)(short)(w[k >> 1] & 0xFFFFu);
    const int D = x >> 1;
    Cout[r + ld * s] = (double)(x & 1);
    Dout[r + ld * s] = (D == kInfCode) ? __longlong_as_double(0x7FF0000000000000ll) : (double)D;
  }
}

__global__ void k_fast_export_pressure(const int32_t* __restrict__ pressure, int64_t n_seg, int64_t S,
                                       int32_t* __restrict__ out, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg * S) return;
  const int64_t seg = i % n_seg, s = i / n_seg;
  out[seg + ld * s] = pressure[((s >> 3) * n_seg + seg) * 8 + (s & 7)];
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

void fast_expand(cudaStream_t st, const uint4* state, int64_t n_slots, const int64_t* pos_of, int64_t R,
                 int64_t S, int64_t s0, int64_t ns, double* Cout, double* Dout, int64_t ld) {
  const int64_t n = R * ((ns + 7) / 8);
  if (n <= 0) return;
  k_fast_expand<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(state, n_slots, pos_of, R, S, s0, ns, Cout, Dout, ld);
}

void fast_export_pressure(cudaStream_t st, const int32_t* pressure, int64_t n_seg, int64_t S, int32_t* out, int64_t ld) {
  const int64_t n = n_seg * S;
  if (n <= 0) return;
  k_fast_export_pressure<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pressure, n_seg, S, out, ld);
}

}  // namespace amro
