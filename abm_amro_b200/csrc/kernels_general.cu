// GENERAL path kernels: the reference's semantics on float64 state planes, for any input the reference
// accepts (fractional weights, arbitrary is_new / countdown values, same-day or backward next_day links).
// They also serve the single-day entry points, whose state arrives as arbitrary float64 columns.
//
// Reference: src/amro/abm_wards.cpp:70-159 (ward step), :359-372 (initial state), :402-407 (propagate),
// :435-464 (totals).  Floating point: every product/sum is a separate IEEE operation (__dmul_rn /
// __dadd_rn, never contracted into FMA) in the operation order of the reference's expressions, so the
// transition probabilities are bit-identical to the CPU's.
#include "kernels.hpp"
#include "rng.cuh"

namespace amro {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7FF0000000000000ll); }

// ---- initial state (abm_wards.cpp:359-372) --------------------------------------------------------------
__global__ void k_gen_init(double* __restrict__ C, double* __restrict__ D, int64_t ld, int64_t R, int64_t S,
                           int64_t n_init, const double* __restrict__ icp, int64_t icp_rs, int64_t icp_cs,
                           const double* __restrict__ idp, int64_t idp_rs, int64_t idp_cs, int replay,
                           const double* __restrict__ u_icol, const double* __restrict__ u_idet, int64_t ui_ld,
                           uint64_t seed, uint64_t sim_offset) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const int64_t r = i % R, s = i / R;
  double c = 0.0, d = dinf();
  if (r < n_init) {
    const double pc = icp[r * icp_rs + s * icp_cs], pd = idp[r * idp_rs + s * idp_cs];
    bool c0, d0;
    if (replay) {
      c0 = u_icol[r + ui_ld * s] < pc;
      d0 = u_idet[r + ui_ld * s] < pd;
    } else {
      c0 = (uint64_t)philox_u32(seed, (uint64_t)r, sim_offset + s, kStreamInitCol) < threshold33(pc);
      d0 = (uint64_t)philox_u32(seed, (uint64_t)r, sim_offset + s, kStreamInitDet) < threshold33(pd);
    }
    c = c0 ? 1.0 : 0.0;
    d = (c0 && d0) ? 1.0 : 0.0;  // :372 a 0/1 flag stored where the dynamics read a countdown
  }
  C[r + ld * s] = c;
  D[r + ld * s] = d;
}

// ---- force of infection per (ward-day, sim) (abm_wards.cpp:95-99) ------------------------------------------
// One warp per (ward-day, group of 32 simulations).  The sum over the ward's rows keeps the reference's order -- the two
// accumulators of arrayops::accumulate, rows in the reference's row order -- so fractional weights sum to the same bits: a
// lane owns one simulation and adds serially.  The loads are transposed through shared memory: a tile of 32 rows x 32
// simulations is read with the lanes along the ROWS (consecutive addresses of a column of the state plane), then walked with
// the lanes along the simulations.
constexpr int kPressWarps = 4;
__global__ void __launch_bounds__(kPressWarps * 32) k_gen_pressure(GenArgs a, int64_t seg_begin, int64_t seg_end) {
  __shared__ double tile[kPressWarps][32][33];
  __shared__ double wrow[kPressWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t groups = (a.S + 31) / 32;
  const int64_t item = (int64_t)blockIdx.x * kPressWarps + warp;
  if (item >= (seg_end - seg_begin) * groups) return;
  const int64_t sl = item / groups, s0 = (item % groups) * 32, seg = seg_begin + sl;
  const int64_t p0 = a.seg_row_start[seg], p1 = a.seg_row_start[seg + 1];
  const int n_s = (int)min((int64_t)32, a.S - s0);
  const int64_t s = s0 + lane;
  double acc1 = 0.0, acc2 = 0.0;   // rows at even / odd offsets within the ward (a ward's odd last row goes to acc1: same rule)
  for (int64_t t0 = p0; t0 < p1; t0 += 32) {
    const int n_r = (int)min((int64_t)32, p1 - t0);
    int64_t r = 0;
    if (lane < n_r) { r = a.order[t0 + lane]; wrow[warp][lane] = a.w[r * a.hw_stride]; }
#pragma unroll 8
    for (int j = 0; j < n_s; ++j)
      if (lane < n_r) tile[warp][j][lane] = a.C[r + a.ld * (s0 + j)];
    __syncwarp();
    if (lane < n_s) {
      // t0 - p0 is a multiple of 32: the parity of a row's offset within the ward is the parity of its index in the tile
      int k = 0;
      for (; k + 1 < n_r; k += 2) {
        acc1 = __dadd_rn(acc1, __dmul_rn(tile[warp][lane][k], wrow[warp][k]));
        acc2 = __dadd_rn(acc2, __dmul_rn(tile[warp][lane][k + 1], wrow[warp][k + 1]));
      }
      if (k < n_r) acc1 = __dadd_rn(acc1, __dmul_rn(tile[warp][lane][k], wrow[warp][k]));
    }
    __syncwarp();
  }
  if (lane < n_s) {
    const double beta = a.params[s + a.p_ld * 1];
    a.F[sl * a.S + s] = __dmul_rn(__ddiv_rn(beta, a.seg_N[seg]), __dadd_rn(acc1, acc2));
    // the ward-day's dither of this simulation (rng.cuh), once per ward-day instead of once per record
    if (a.dither) a.dither[sl * a.S + s] = philox_chunk(a.seed, (uint64_t)seg, a.sim_offset + s, kStreamDither);
  }
}
// dither of ward-free probabilities: one per simulation and run, kept behind the ward-days' rows of `dither`
__global__ void k_gen_run_dither(GenArgs a, int64_t row) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < a.S) a.dither[row * a.S + s] = philox_chunk(a.seed, kRunDitherIndex, a.sim_offset + s, kStreamDither);
}

// ---- per (row, sim) update (abm_wards.cpp:102-156) ---------------------------------------------------------
__global__ void k_gen_update(GenArgs a, int64_t row_begin, int64_t row_end, int64_t seg_begin, int test_h, int test_a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = row_end - row_begin;
  if (i >= n * a.S) return;
  const int64_t p = row_begin + i % n, s = i / n;
  const int64_t r = a.order[p];
  const int64_t sl = (int64_t)a.seg_of_pos[p] - seg_begin;
  const double alpha = a.params[s], gamma = a.params[s + a.p_ld * 2];
  const double rho_h = a.params[s + a.p_ld * 3], alpha2 = a.params[s + a.p_ld * 4], rho_i = a.params[s + a.p_ld * 5];
  const double force = a.F[sl * a.S + s];
  const double h = a.h[r * a.hw_stride], w = a.w[r * a.hw_stride];
  const double c = a.C[r + a.ld * s];
  double d = a.D[r + a.ld * s];

  const double W = __dmul_rn(c, w);                                                    // :95
  const double det = (d <= 0.0) ? 1.0 : 0.0;                                           // :102
  const double coef = __dadd_rn(__dmul_rn(__dsub_rn(1.0, det), __dsub_rn(1.0, alpha)),
                                __dmul_rn(det, __dsub_rn(1.0, alpha2)));               // :103-105
  double P = __dadd_rn(__dmul_rn(coef, W), __dmul_rn(__dsub_rn(1.0, W), force));       // :108-109
  P = __dadd_rn(P, __dmul_rn(h, gamma));                                               // :112
  if (a.p_out) a.p_out[r + a.p_out_ld * s] = P;

  bool col;
  uint32_t v14 = 0, r16 = 0;
  if (a.replay) {
    col = a.u_col[r + a.u_ld * s] < P;                                                 // :116
  } else {  // rng.cuh: 14-bit draw against the dithered threshold (dither of the ward-day, or of the run when P is ward-free)
    const bool ward_free = (W == 1.0) && (fabs(force) <= 1.7976931348623157e308);
    v14 = philox_u14(a.seed, (uint64_t)r, a.sim_offset + s);
    r16 = philox_chunk(a.seed, ward_free ? kRunDitherIndex : (uint64_t)a.seg_of_pos[p], a.sim_offset + s, kStreamDither);
    col = v14 < bernoulli14(P, r16);
  }
  const bool h0 = (h == 0.0), h1 = (h == 1.0);
  if (col && (h0 || h1)) {                                                             // :129 / :141
    if (d <= (double)a.ttd) {
      d = __dsub_rn(d, 1.0);                                                           // :131-132
    } else {
      const bool flag = h0 ? (test_h != 0) : (test_a != 0);
      const double rho = h0 ? rho_h : rho_i;
      bool hit = false;
      if (flag) {
        if (a.replay) hit = a.u_det[r + a.u_ld * s] <= rho;                            // :134 / :146
        else hit = v14 < bernoulli14(__dmul_rn(clamp01(P), rho), r16);
      }
      d = hit ? (double)a.ttd : dinf();                                                // :135,138
    }
  } else {
    d = dinf();                                                                        // :153-155
  }
  a.C[r + a.ld * s] = col ? 1.0 : 0.0;
  a.D[r + a.ld * s] = d;
}

// The same update in free-running mode for runs whose simulations start on a 32-boundary of the global index: one thread per
// (row, group of 32 simulations).  The plane words of the daily draws belong to (row, group) (rng.cuh): the thread computes the
// four Philox calls ONCE and takes each simulation's bit of every plane, where the per-(row, sim) kernel above repeats the
// calls for each of the 32 simulations; the dithers come from k_gen_pressure / k_gen_run_dither.  Consecutive lanes are
// consecutive rows, so the loads of a simulation's column are coalesced as above.
__global__ void k_gen_update_groups(GenArgs a, int64_t row_begin, int64_t row_end, int64_t seg_begin, int test_h, int test_a,
                                    int64_t run_dither_row) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = row_end - row_begin, groups = (a.S + 31) / 32;
  if (i >= n * groups) return;
  const int64_t p = row_begin + i % n, gi = i / n;
  const int64_t r = a.order[p];
  const int64_t sl = (int64_t)a.seg_of_pos[p] - seg_begin;
  const double h = a.h[r * a.hw_stride], w = a.w[r * a.hw_stride];
  const bool h0 = (h == 0.0), h1 = (h == 1.0);
  uint32_t pw[16];
  {
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32), g = (uint32_t)((a.sim_offset >> 5) + (uint64_t)gi);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 v = philox4x32<kPlaneRounds>((uint32_t)r, (uint32_t)((uint64_t)r >> 32), g, kStreamPlanes + (uint32_t)c, k0, k1);
      pw[4 * c] = v.x; pw[4 * c + 1] = v.y; pw[4 * c + 2] = v.z; pw[4 * c + 3] = v.w;
    }
  }
  const int64_t s_end = min(a.S, gi * 32 + 32);
  for (int64_t s = gi * 32; s < s_end; ++s) {
    const int b = (int)(s & 31);
    uint32_t v14 = 0;
#pragma unroll
    for (int j = 0; j < kDrawBits; ++j) v14 |= ((pw[j] >> b) & 1u) << j;
    const double alpha = a.params[s], gamma = a.params[s + a.p_ld * 2];
    const double rho_h = a.params[s + a.p_ld * 3], alpha2 = a.params[s + a.p_ld * 4], rho_i = a.params[s + a.p_ld * 5];
    const double force = a.F[sl * a.S + s];
    const double c = a.C[r + a.ld * s];
    double d = a.D[r + a.ld * s];
    const double W = __dmul_rn(c, w);                                                    // :95
    const double det = (d <= 0.0) ? 1.0 : 0.0;                                           // :102
    const double coef = __dadd_rn(__dmul_rn(__dsub_rn(1.0, det), __dsub_rn(1.0, alpha)),
                                  __dmul_rn(det, __dsub_rn(1.0, alpha2)));               // :103-105
    double P = __dadd_rn(__dmul_rn(coef, W), __dmul_rn(__dsub_rn(1.0, W), force));       // :108-109
    P = __dadd_rn(P, __dmul_rn(h, gamma));                                               // :112
    if (a.p_out) a.p_out[r + a.p_out_ld * s] = P;
    const bool ward_free = (W == 1.0) && (fabs(force) <= 1.7976931348623157e308);
    const uint32_t r16 = a.dither[(ward_free ? run_dither_row : sl) * a.S + s];
    const bool col = v14 < bernoulli14(P, r16);
    if (col && (h0 || h1)) {                                                             // :129 / :141
      if (d <= (double)a.ttd) {
        d = __dsub_rn(d, 1.0);                                                           // :131-132
      } else {
        const bool flag = h0 ? (test_h != 0) : (test_a != 0);
        const double rho = h0 ? rho_h : rho_i;
        const bool hit = flag && v14 < bernoulli14(__dmul_rn(clamp01(P), rho), r16);     // :134 / :146
        d = hit ? (double)a.ttd : dinf();                                                // :135,138
      }
    } else {
      d = dinf();                                                                        // :153-155
    }
    a.C[r + a.ld * s] = col ? 1.0 : 0.0;
    a.D[r + a.ld * s] = d;
  }
}

// ---- next_day propagation (abm_wards.cpp:402-407): gather the sources' post-step state, then scatter ------
__global__ void k_gen_gather(const double* __restrict__ C, const double* __restrict__ D, int64_t ld, int64_t S,
                             const int64_t* __restrict__ src, int64_t n, double* __restrict__ tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * S) return;
  const int64_t j = i % n, s = i / n;
  tmp[i] = C[src[j] + ld * s];
  tmp[n * S + i] = D[src[j] + ld * s];
}
__global__ void k_gen_scatter(double* __restrict__ C, double* __restrict__ D, int64_t ld, int64_t S,
                              const int64_t* __restrict__ dst, int64_t n, const double* __restrict__ tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * S) return;
  const int64_t j = i % n, s = i / n;
  C[dst[j] + ld * s] = tmp[i];
  D[dst[j] + ld * s] = tmp[n * S + i];
}

// ---- per-day totals (abm_wards.cpp:445-459) -----------------------------------------------------------------
// A warp walks kCountSpan consecutive 32-row chunks of one simulation column (lanes = consecutive rows).  While whole chunks
// belong to one day the hits are kept in a register; a chunk that touches several days is merged with match_any (one atomic
// per distinct day).  Rows of a day are consecutive in every hospital whose rows are in day order: one atomic per span then.
constexpr int kCountSpan = 16;
__global__ void k_gen_count(const double* __restrict__ plane, int64_t ld, int64_t R, int64_t S,
                            const int32_t* __restrict__ count_day, int mode, int32_t* __restrict__ counts, int64_t ld_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_col = (R + 32 * kCountSpan - 1) / (32 * kCountSpan);
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= warps_per_col * S) return;
  const int64_t s = wid / warps_per_col, r0 = (wid % warps_per_col) * (32 * kCountSpan);
  int cur_day = -1, acc = 0;   // warp-uniform
  for (int k = 0; k < kCountSpan; ++k) {
    const int64_t r = r0 + 32 * k + lane;
    int day = -1;
    bool hit = false;
    if (r < R) {
      day = count_day[r];
      const double v = plane[r + ld * s];
      hit = (day >= 0) && (mode == 1 ? (v == 1.0) : (v <= 0.0));
    }
    const unsigned same = __match_any_sync(0xFFFFFFFFu, day);
    const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
    const int day0 = __shfl_sync(0xFFFFFFFFu, day, 0);
    if (same == 0xFFFFFFFFu && day0 == cur_day) { acc += __popc(hits); continue; }
    if (acc && lane == 0) atomicAdd(&counts[cur_day + ld_out * s], acc);
    acc = 0; cur_day = -1;
    if (same == 0xFFFFFFFFu) { cur_day = day0; acc = day0 >= 0 ? __popc(hits) : 0; continue; }
    if (day >= 0 && lane == __ffs(same) - 1) {
      const int n = __popc(same & hits);
      if (n) atomicAdd(&counts[day + ld_out * s], n);
    }
  }
  if (acc && lane == 0) atomicAdd(&counts[cur_day + ld_out * s], acc);
}

__global__ void k_i32_to_f64(const int32_t* __restrict__ in, double* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}

// per-day moments over simulations: one warp per day (counts are exact integers in float64 up to 2^53)
__global__ void k_day_moments(const int32_t* __restrict__ col, const int32_t* __restrict__ det, int64_t n_days, int64_t S,
                              int64_t ld_c, double* __restrict__ out, int64_t ld) {
  const int64_t day = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (day >= n_days) return;
  double a = 0, b = 0, c = 0, d = 0;
  for (int64_t s = lane; s < S; s += 32) {
    const double x = (double)col[day + ld_c * s], y = (double)det[day + ld_c * s];
    a += x; b += x * x; c += y; d += y * y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    c += __shfl_xor_sync(0xFFFFFFFFu, c, o); d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
  }
  if (lane == 0) { out[day] = a; out[day + ld] = b; out[day + 2 * ld] = c; out[day + 3 * ld] = d; }
}

// The same moments delivered to every GPU of the box in the same pass: lane 0 of a day's warp stores the four sums into the
// slot of this rank in EVERY rank's buffer (peer memory mapped over NVLink / NVSwitch; the own buffer is one of them).  No
// rank waits for another inside a run: readers add the slots up after the barrier that ends a batch of runs.
struct PeerDst { double* p[kMaxPeers]; };
__global__ void k_day_moments_peers(const int32_t* __restrict__ col, const int32_t* __restrict__ det, int64_t n_days, int64_t S,
                                    int64_t ld_c, PeerDst dst, int n_dst, int64_t ld) {
  const int64_t day = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (day >= n_days) return;
  double a = 0, b = 0, c = 0, d = 0;
  for (int64_t s = lane; s < S; s += 32) {
    const double x = (double)col[day + ld_c * s], y = (double)det[day + ld_c * s];
    a += x; b += x * x; c += y; d += y * y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    c += __shfl_xor_sync(0xFFFFFFFFu, c, o); d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
  }
  // lanes 0 .. n_dst-1 each serve one destination (the stores of a day go out together)
  if (lane < n_dst) {
    double* out = dst.p[lane];
    out[day] = a; out[day + ld] = b; out[day + 2 * ld] = c; out[day + 3 * ld] = d;
  }
}

// per-simulation squared distance between the per-day counts and observed series: one warp per simulation (its days are
// contiguous: coalesced).  Sums of squares of integers minus data, accumulated in float64 in a fixed lane/shuffle order.
__global__ void k_distance(const int32_t* __restrict__ col, const int32_t* __restrict__ det, int64_t n_days, int64_t S, int64_t ld_c,
                           const double* __restrict__ obs_col, const double* __restrict__ obs_det, double* __restrict__ out) {
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  double acc = 0.0;
  for (int64_t d = lane; d < n_days; d += 32) {
    if (obs_col) { const double o = obs_col[d]; if (o == o) { const double e = (double)col[d + ld_c * s] - o; acc += e * e; } }
    if (obs_det) { const double o = obs_det[d]; if (o == o) { const double e = (double)det[d + ld_c * s] - o; acc += e * e; } }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0) out[s] = acc;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

void gen_init(cudaStream_t st, double* C, double* D, int64_t ld, int64_t R, int64_t S, int64_t n_init,
              const double* icp, int64_t icp_rs, int64_t icp_cs, const double* idp, int64_t idp_rs, int64_t idp_cs,
              int replay, const double* u_icol, const double* u_idet, int64_t ui_ld, uint64_t seed, uint64_t sim_offset) {
  if (R * S == 0) return;
  k_gen_init<<<blocks_for(R * S), kThreads, 0, st>>>(C, D, ld, R, S, n_init, icp, icp_rs, icp_cs, idp, idp_rs, idp_cs,
                                                     replay, u_icol, u_idet, ui_ld, seed, sim_offset);
}
void gen_pressure(cudaStream_t st, const GenArgs& a, int64_t seg_begin, int64_t seg_end) {
  const int64_t n = (seg_end - seg_begin) * ((a.S + 31) / 32);   // warps
  if (n <= 0) return;
  k_gen_pressure<<<(unsigned)((n + kPressWarps - 1) / kPressWarps), kPressWarps * 32, 0, st>>>(a, seg_begin, seg_end);
}
void gen_run_dither(cudaStream_t st, const GenArgs& a) {
  if (!a.dither || a.S <= 0) return;
  k_gen_run_dither<<<blocks_for(a.S), kThreads, 0, st>>>(a, a.run_dither_row);
}
void gen_update(cudaStream_t st, const GenArgs& a, int64_t row_begin, int64_t row_end, int64_t seg_begin, int test_h, int test_a) {
  const int64_t n = (row_end - row_begin) * a.S;
  if (n <= 0) return;
  if (!a.replay && a.dither && (a.sim_offset & 31u) == 0u) {
    const int64_t m = (row_end - row_begin) * ((a.S + 31) / 32);
    k_gen_update_groups<<<blocks_for(m), kThreads, 0, st>>>(a, row_begin, row_end, seg_begin, test_h, test_a, a.run_dither_row);
    return;
  }
  k_gen_update<<<blocks_for(n), kThreads, 0, st>>>(a, row_begin, row_end, seg_begin, test_h, test_a);
}
void gen_propagate(cudaStream_t st, double* C, double* D, int64_t ld, int64_t S, const int64_t* push_src,
                   const int64_t* push_dst, int64_t n_push, double* tmp) {
  const int64_t n = n_push * S;
  if (n <= 0) return;
  k_gen_gather<<<blocks_for(n), kThreads, 0, st>>>(C, D, ld, S, push_src, n_push, tmp);
  k_gen_scatter<<<blocks_for(n), kThreads, 0, st>>>(C, D, ld, S, push_dst, n_push, tmp);
}
void gen_count(cudaStream_t st, const double* plane, int64_t ld, int64_t R, int64_t S, const int32_t* count_day,
               int mode, int32_t* counts, int64_t ld_out) {
  if (R * S == 0) return;
  const int64_t warps = (R + 32 * kCountSpan - 1) / (32 * kCountSpan) * S;
  k_gen_count<<<blocks_for(warps * 32), kThreads, 0, st>>>(plane, ld, R, S, count_day, mode, counts, ld_out);
}
void day_moments(cudaStream_t st, const int32_t* col, const int32_t* det, int64_t n_days, int64_t S, int64_t ld_c, double* out, int64_t ld) {
  if (n_days <= 0) return;
  k_day_moments<<<blocks_for(n_days * 32), kThreads, 0, st>>>(col, det, n_days, S, ld_c, out, ld);
}
void day_moments_peers(cudaStream_t st, const int32_t* col, const int32_t* det, int64_t n_days, int64_t S, int64_t ld_c,
                       double* const* dst, int n_dst, int64_t ld) {
  if (n_days <= 0 || n_dst <= 0) return;
  PeerDst d{};
  for (int i = 0; i < n_dst; ++i) d.p[i] = dst[i];
  k_day_moments_peers<<<blocks_for(n_days * 32), kThreads, 0, st>>>(col, det, n_days, S, ld_c, d, n_dst, ld);
}
void distance_to_data(cudaStream_t st, const int32_t* col, const int32_t* det, int64_t n_days, int64_t S, int64_t ld_c,
                      const double* obs_col, const double* obs_det, double* out) {
  if (S <= 0) return;
  k_distance<<<blocks_for(S * 32), kThreads, 0, st>>>(col, det, n_days, S, ld_c, obs_col, obs_det, out);
}
void counts_to_double(cudaStream_t st, const int32_t* counts, double* out, int64_t n) {
  if (n <= 0) return;
  k_i32_to_f64<<<blocks_for(n), kThreads, 0, st>>>(counts, out, n);
}

}  // namespace amro
