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
// One thread per (segment, sim) walks the ward's rows in the reference's row order with the two
// accumulators of arrayops::accumulate, so fractional weights sum to the same bits.
__global__ void k_gen_pressure(GenArgs a, int64_t seg_begin, int64_t seg_end) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_seg = seg_end - seg_begin;
  if (i >= n_seg * a.S) return;
  const int64_t s = i % a.S, sl = i / a.S, seg = seg_begin + sl;
  const int64_t p0 = a.seg_row_start[seg], p1 = a.seg_row_start[seg + 1];
  double acc1 = 0.0, acc2 = 0.0;
  int64_t p = p0;
  for (; p + 1 < p1; p += 2) {
    const int64_t r0 = a.order[p], r1 = a.order[p + 1];
    acc1 = __dadd_rn(acc1, __dmul_rn(a.C[r0 + a.ld * s], a.w[r0 * a.hw_stride]));
    acc2 = __dadd_rn(acc2, __dmul_rn(a.C[r1 + a.ld * s], a.w[r1 * a.hw_stride]));
  }
  if (p < p1) {
    const int64_t r0 = a.order[p];
    acc1 = __dadd_rn(acc1, __dmul_rn(a.C[r0 + a.ld * s], a.w[r0 * a.hw_stride]));
  }
  const double beta = a.params[s + a.p_ld * 1];
  a.F[sl * a.S + s] = __dmul_rn(__ddiv_rn(beta, a.seg_N[seg]), __dadd_rn(acc1, acc2));
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
// Lanes = consecutive rows of one simulation column; lanes of the same day are merged with match_any so a
// warp issues one atomic per distinct day it touches.
__global__ void k_gen_count(const double* __restrict__ plane, int64_t ld, int64_t R, int64_t S,
                            const int32_t* __restrict__ count_day, int mode, int32_t* __restrict__ counts, int64_t ld_out) {
  const int64_t rows_per_block = blockDim.x;
  const int64_t n_row_blocks = (R + rows_per_block - 1) / rows_per_block;
  const int64_t s = blockIdx.x / n_row_blocks;
  const int64_t r = (blockIdx.x % n_row_blocks) * rows_per_block + threadIdx.x;
  int day = -1;
  bool hit = false;
  if (r < R && s < S) {
    day = count_day[r];
    const double v = plane[r + ld * s];
    hit = (day >= 0) && (mode == 1 ? (v == 1.0) : (v <= 0.0));
  }
  const unsigned same = __match_any_sync(0xFFFFFFFFu, day);
  const unsigned hits = __ballot_sync(0xFFFFFFFFu, hit);
  const int lane = threadIdx.x & 31;
  if (day >= 0 && lane == __ffs(same) - 1) {
    const int n = __popc(same & hits);
    if (n) atomicAdd(&counts[day + ld_out * s], n);
  }
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
  const int64_t n = (seg_end - seg_begin) * a.S;
  if (n <= 0) return;
  k_gen_pressure<<<blocks_for(n), kThreads, 0, st>>>(a, seg_begin, seg_end);
}
void gen_update(cudaStream_t st, const GenArgs& a, int64_t row_begin, int64_t row_end, int64_t seg_begin, int test_h, int test_a) {
  const int64_t n = (row_end - row_begin) * a.S;
  if (n <= 0) return;
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
  const int64_t n_row_blocks = (R + kThreads - 1) / kThreads;
  k_gen_count<<<(unsigned)(n_row_blocks * S), kThreads, 0, st>>>(plane, ld, R, S, count_day, mode, counts, ld_out);
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
