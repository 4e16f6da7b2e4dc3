// Launcher declarations shared by api.cu and the two kernel translation units.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace amro {

// ----------------------------------------------------------------------------------------------------
// GENERAL path: float64 state planes laid out like the reference's matrix columns (element (row, sim) at
// plane[row + ld*sim]), any weights / is_new / countdown values.  One launch per phase per day.
// ----------------------------------------------------------------------------------------------------
struct GenArgs {
  double* C;                 // colonised plane  (reference columns 6 .. 6+S-1)
  double* D;                 // countdown plane  (reference columns 6+S .. 6+2S-1)
  int64_t ld;
  const double* h;           // per ORIGINAL row
  const double* w;
  int64_t hw_stride;         // element stride of h / w (1)
  const int64_t* order;      // position -> original row
  const uint32_t* seg_of_pos;  // position -> absolute segment
  const int64_t* seg_row_start;
  const double* seg_N;
  const double* params;      // [>=S x 6]
  int64_t p_ld;
  int64_t S;
  double* F;                 // scratch [n_segments_of_day x S]: force of infection
  int ttd;
  int replay;
  const double* u_col;       // [rows x S], element (orig row, sim) at u[row + u_ld*sim]
  const double* u_det;
  int64_t u_ld;
  uint64_t seed;
  uint64_t sim_offset;
  double* p_out;             // optional [rows x S]
  int64_t p_out_ld;
  uint32_t* dither;          // free-running mode: [(ward-days of a day) + 1 x S] dithers of the day's ward-days (k_gen_pressure)
  int64_t run_dither_row;    // ... and, in row run_dither_row, the run's dither of ward-free probabilities
};

void gen_init(cudaStream_t st, double* C, double* D, int64_t ld, int64_t R, int64_t S, int64_t n_init,
              const double* icp, int64_t icp_rs, int64_t icp_cs, const double* idp, int64_t idp_rs, int64_t idp_cs,
              int replay, const double* u_icol, const double* u_idet, int64_t ui_ld, uint64_t seed, uint64_t sim_offset);
void gen_pressure(cudaStream_t st, const GenArgs& a, int64_t seg_begin, int64_t seg_end);
void gen_run_dither(cudaStream_t st, const GenArgs& a);   // once per run, before the first day
void gen_update(cudaStream_t st, const GenArgs& a, int64_t row_begin, int64_t row_end, int64_t seg_begin,
                int test_h, int test_a);
void gen_propagate(cudaStream_t st, double* C, double* D, int64_t ld, int64_t S, const int64_t* push_src,
                   const int64_t* push_dst, int64_t n_push, double* tmp);
// counts[day + ld_out*sim] += #rows of that day with (mode 1: v == 1) or (mode 0: v <= 0)
void gen_count(cudaStream_t st, const double* plane, int64_t ld, int64_t R, int64_t S, const int32_t* count_day,
               int mode, int32_t* counts, int64_t ld_out);
void counts_to_double(cudaStream_t st, const int32_t* counts, double* out, int64_t n);
// out[day + ld*{0,1,2,3}] = sum / sum of squares over simulations of col[day + ld_c*s], then of det[...]
// out[s] = sum over days with data (non-NaN) of (col - obs_col)^2 + (det - obs_det)^2; either series may be nullptr
void distance_to_data(cudaStream_t st, const int32_t* col, const int32_t* det, int64_t n_days, int64_t S, int64_t ld_c,
                      const double* obs_col, const double* obs_det, double* out);
// summary.cu: X [n_days x S] (device, column-major, ld) -> out [n_days x (3 + nq)] (device, ld_out): day, mean, population sd, quantiles
// (arma::quantile, Hyndman & Fan 5) over the simulations; the same operation order as the host routine (bit-identical)
void device_summary(cudaStream_t st, const double* X, int64_t n_days, int64_t S, int64_t ld, const double* q_dev, int64_t nq, double* out,
                    int64_t ld_out);
// the same, stored into n_dst <= kMaxPeers buffers (this GPU's and its peers' over NVLink): the run's collective as peer stores
constexpr int kMaxPeers = 16;
void day_moments_peers(cudaStream_t st, const int32_t* col, const int32_t* det, int64_t n_days, int64_t S, int64_t ld_c,
                       double* const* dst, int n_dst, int64_t ld);
void day_moments(cudaStream_t st, const int32_t* col, const int32_t* det, int64_t n_days, int64_t S, int64_t ld_c, double* out, int64_t ld);

// ----------------------------------------------------------------------------------------------------
// FAST path: packed 16-bit state, 8 simulations (one 16-byte word) per patient-day record and slice; one
// persistent cooperative kernel runs the initial state and every day.  See kernels_fast.cu.
// ----------------------------------------------------------------------------------------------------
constexpr int kSimsPerSlice = 8;
#ifndef AMRO_SLICES_PER_TEAM
#define AMRO_SLICES_PER_TEAM 2
#endif
constexpr int kSlicesPerTeam = AMRO_SLICES_PER_TEAM;           // a team (the CTAs that share one barrier) steps this many slices together
constexpr int kMaxCountdown = 8000;        // |countdown| representable in a 13-bit field: days and time_to_detect limits
#ifndef AMRO_WIDE_THREADS
#define AMRO_WIDE_THREADS 512
#endif
#ifndef AMRO_TEAM_THREADS
#define AMRO_TEAM_THREADS 128
#endif
#ifndef AMRO_TEAM_MINBLOCKS
#define AMRO_TEAM_MINBLOCKS 4
#endif
constexpr int kShapeWide = 0, kShapeTeam = 1, kShapeRoomy = 2;   // Roomy: Team with the register budget of 3 CTAs per SM
constexpr int kWideThreads = AMRO_WIDE_THREADS;     // one CTA per SM
constexpr int kTeamThreads = AMRO_TEAM_THREADS;
constexpr int kTeamMinBlocks = AMRO_TEAM_MINBLOCKS; // CTAs per SM the team shape's register budget is sized for

struct FastArgs {
  // topology (device)
  const uint4* meta;             // per record: (is_new | source | ward-day within the day, byte offset of the successor's slot in a
                                 // state-ring buffer or kNone, successor's ward-day or kNone, original row (low 32 bits))
  const uint32_t* seg_main_next; // per ward-day: the ward-day most of its records feed (same ward, next day)
  const uint32_t* init_seg;
  const uint32_t* init_slot;
  int identity_order;            // records are stepped in their original row order (row = position)
  int has_half;                  // some records weigh 1/2: ward pressure is counted in halves, tables exist per weight
  const int64_t* day_start;
  const int64_t* day_seg_start;
  const int64_t* seg_row_start;
  const double* seg_N;
  int64_t R, n_init, n_days, n_seg, ring_chunks;   // ring_chunks: 32-record chunks of the fullest day
  // run
  const double* params; int64_t p_ld;   // [S_pad x 6] device copy, padded rows are zeros
  const double* icp; int64_t icp_rs, icp_cs;
  const double* idp; int64_t idp_rs, idp_cs;
  int64_t S;                     // real simulations
  int64_t n_slices;              // ceil(S / 8)
  uint64_t seed, sim_offset;
  uint32_t rk[20];               // Philox round keys (key + r*Weyl), consumed straight from the constant bank
  int ttd;
  uint64_t tsh, tsa;             // schedules converted the way the reference's unsigned % does
  // replay
  const double* u_col; const double* u_det; int64_t u_ld;
  const double* u_icol; const double* u_idet; int64_t ui_ld;
  double* p_out;                 // optional [R x S] (ld = R)
  // state
  uint4* pre;                    // [n_teams][2][ring_chunks][slices per team][32]: pre-step state of the current / next day
  uint4* traj;                   // nullptr, or [n_teams][R][slices per team]: post-step state of every record
  uint32_t* pressure;            // [n_teams][n_seg][slices per team][4]: 8 x 16-bit colonised counts per ward-day and slice
  int32_t* counts_col;           // [n_days_out x S_pad8] ld = counts_ld  (zeroed by the caller)
  int32_t* counts_det;
  int64_t counts_ld;
  unsigned int* team_barrier;    // [n_teams] zeroed by the caller
  int ctas_per_team;             // CTAs that share a team's slices (1 = no inter-CTA barrier) ...
  int teams_plus_one;            // ... the first teams_plus_one teams have one more, so that the grid fills the SMs evenly
  int cluster_barrier;           // 1: every team is a thread-block cluster and meets at the hardware cluster barrier
  // bit-sliced kernel (kernels_bits.cu): ward pressure is pulled from the state ring; `pressure` is nullptr or the export buffer
  // [n_teams][press_days = n_days + 1][max_segs_per_day][words per team][32] of 32-bit counts
  int64_t max_segs_per_day;
  int press_days;
  const uint32_t* src_mask;        // one bit per record: it has a source (a predecessor or an initial row) and weighs 1; word = 32-record chunk of a day
  const uint32_t* half_mask;       // nullptr, or the same for the records that weigh 1/2
  const int64_t* day_chunk_start;  // [n_days + 1] first chunk of every day in src_mask
  int pull_deep;                   // which variant counts the ward pressure (kernels_bits.cu: pull_pressure / pull_pressure_lean)
};

// element (uint4) index of slot q of a team's first slice in a state-ring buffer laid out [chunk of 32 slots][slice][32]
__host__ __device__ inline int64_t fast_ring_slot(uint32_t q) { return (int64_t)(q >> 5) * (kSlicesPerTeam * 32) + (q & 31u); }
size_t fast_shared_bytes(bool replay, int shape);
int fast_threads(int shape);
// CTAs of the given launch shape that can be co-resident (bound of a cooperative launch)
int fast_max_coresident_ctas(int device, bool replay, bool lean, int shape);
int fast_max_coresident_clusters(int device, bool replay, bool lean, int shape, int cluster_size);
// lean: no trajectory, unit weights, rows already in (day, ward) order: the instance compiled without the other cases
void fast_launch(cudaStream_t st, const FastArgs& a, bool replay, bool lean, int shape, int grid);

// rows [r0, r0 + nr) x sims [s0, s0 + ns) of the run (s0 a multiple of 8): Cout[(row - r0) + ld*(s - s0)] = C, Dout[...] = D
void fast_expand(cudaStream_t st, const uint4* traj, const int64_t* pos_of, int64_t R, int64_t S, int64_t r0, int64_t nr, int64_t s0,
                 int64_t ns, double* Cout, double* Dout, int64_t ld);
// out[seg + ld*sim] = colonised count the ward-day started with
void fast_export_pressure(cudaStream_t st, const uint32_t* pressure, int64_t n_seg, int64_t S, int32_t* out, int64_t ld);


// ----------------------------------------------------------------------------------------------------
// BITS path: the state bit-sliced over simulations (one bit per plane, 32 simulations per register): summary runs of
// unit-weight hospitals with time_to_detect <= 2 in free-running mode.  See kernels_bits.cu.  Same ring / meta / launch
// shapes (kShapeTeam, kShapeRoomy) as the FAST path; a team steps bits_sims_per_team() simulations.
// ----------------------------------------------------------------------------------------------------
constexpr int kBitsMaxTtd = 6;   // 0 .. 2: shift-register planes; 3 .. 6: a 3-bit countdown code per simulation
int bits_sims_per_team();
int bits_words_per_team();   // groups of 32 simulations of a team: the warps of a team pair up, one warp per word
size_t bits_shared_bytes(bool half);   // half: the instances for hospitals with weights 1/2 (more table rows per warp)
int bits_max_coresident_ctas(int device, int ttd, int shape, bool half);
int bits_max_coresident_clusters(int ttd, int shape, int cluster_size, bool half);
void bits_launch(cudaStream_t st, const FastArgs& a, int shape, int grid);
// out[ward-day + ld*sim] = colonised count the ward-day started with (press_days = n_days + 1 runs only)
// src_mask / day_chunk_start of a compiled hospital (device arrays; day_chunk_start from the host copy of day_start)
void bits_build_src_mask(cudaStream_t st, const uint4* meta, const int64_t* day_start, const int64_t* day_chunk_start, int64_t n_days,
                         int64_t max_rows_per_day, uint32_t* src_mask, uint32_t* half_mask /* nullptr: no half weights */);
void bits_export_pressure(cudaStream_t st, const uint32_t* pressure, const int64_t* day_seg_start, int64_t n_days, int64_t max_segs,
                          int64_t press_days, int64_t S, int32_t* out, int64_t ld);

}  // namespace amro
