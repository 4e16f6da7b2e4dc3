/* amro_b200 -- C ABI of the B200-native ward-level colonisation path of ABM_AMRO.
 *
 * This header is the drop-in boundary.  The reference exposes the path as seven pybind11 functions of
 * the top-level module `amro` (/root/reference/src/amro/amro.cpp:18-318, prototypes in
 * /root/reference/src/amro/abm_wards.hpp:5-54); there is no C ABI in the reference, so each entry point
 * below states which of those functions it replaces.  Signatures use plain pointers and sizes only (no
 * torch / numpy / armadillo types).  The reference-side binding a maintainer would add is shown in
 * INTEGRATION.md; abm_amro_b200/csrc/amro_module.cpp is that binding, built as the module `amro`.
 *
 * Matrix convention (the reference's, abm_wards.cpp:87-90,291-295,335-338): float64, column-major,
 * element (r,c) at ptr[r + ld*c].
 *   ward_matrix [R x 6]          day | ward | MRN | is_new h | weight w | next_day
 *   state matrix [R x (6+2S)]    the six columns above, then C_0..C_{S-1}, then D_0..D_{S-1}
 *   total_patients_per_ward      day | ward | N
 *   parameters [>=S x 6]         alpha | beta | gamma | rho_hospital | alpha2 | rho_imported
 *
 * Every function returns AMRO_OK or an error class; amro_last_error() gives the message of the calling
 * thread's last failure.  The error classes map onto what the reference raises through pybind11
 * (SURVEY.md 8b): AMRO_EINDEX -> IndexError (armadillo bounds errors), AMRO_ERUNTIME -> RuntimeError
 * (armadillo size mismatches), AMRO_EINVAL -> ValueError (inputs the reference would crash on or that
 * this implementation refuses), AMRO_ECUDA -> RuntimeError (no device / CUDA failure: there is no CPU
 * fallback).
 */
#ifndef AMRO_B200_H
#define AMRO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMRO_B200_VERSION "0.2.2"  /* amro.__version__ of the reference (setup.py:6,18; amro.cpp:320-324) */

enum amro_status { AMRO_OK = 0, AMRO_EINDEX = 1, AMRO_ERUNTIME = 2, AMRO_EINVAL = 3, AMRO_ECUDA = 4 };

/* Uniform source.  PHILOX is the free-running production mode (counter-based, DESIGN.md "RNG spec");
 * REPLAY consumes caller-supplied uniforms, one per (row, simulation) and draw slot, and is bit-exact
 * against the reference fed the same numbers. */
enum amro_rng_mode { AMRO_RNG_PHILOX = 0, AMRO_RNG_REPLAY = 1 };

/* Which device implementation runs.  AUTO picks FAST when the hospital qualifies (unit weights, is_new
 * in {0,1}, every record stepped exactly once and read by at most its next_day successor), else GENERAL,
 * the float64 kernels that follow the reference's semantics for any input. */
enum amro_path { AMRO_PATH_AUTO = 0, AMRO_PATH_FAST = 1, AMRO_PATH_GENERAL = 2 };

const char* amro_last_error(void);
const char* amro_version(void);
int amro_device_count(void);                 /* 0 when no CUDA device is usable */

/* ------------------------------------------------------------------------------------------------
 * Hospital topology: ward_matrix + total_patients_per_ward compiled once into the device-resident form
 * the kernels read (rows grouped by (day, ward), predecessor links = inverse of next_day, per ward-day
 * N).  Replaces the per-day find/unique/gather/scatter of abm_wards.cpp:235-253,384-407.
 * ---------------------------------------------------------------------------------------------- */
typedef struct amro_topology amro_topology;

typedef struct amro_topology_info {
  int64_t n_rows;            /* R                                                              */
  int64_t n_rows_stepped;    /* rows whose day is an integer in [0, total_days]                */
  int64_t n_init;            /* rows of the initial-state block                                */
  int64_t total_days;        /* max(total_patients_per_ward[:,0]) (abm_wards.cpp:356)          */
  int64_t n_days_out;        /* max(ward_matrix[:,0]) + 1: rows of the per-day totals (:440)   */
  int64_t n_segments;        /* (day, ward) groups                                             */
  int64_t max_rows_per_day;
  int32_t fast_eligible;     /* 1 if AMRO_PATH_FAST can run this hospital                      */
  int32_t ring_eligible;     /* same as fast_eligible (every predecessor link spans exactly one day); kept for layout */
  int32_t identity_order;    /* 1 if the input rows are already grouped by (day, ward)         */
  int32_t device;
  double  compile_ms;        /* host time spent compiling + uploading                          */
  char    why_not_fast[128]; /* first disqualifying property, "" if eligible                   */
} amro_topology_info;

int  amro_topology_create(const double* ward_matrix, int64_t n_rows, int64_t n_cols, int64_t ld,
                          const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                          int64_t n_init, int device, amro_topology** out);
void amro_topology_destroy(amro_topology* topo);
int  amro_topology_get_info(const amro_topology* topo, amro_topology_info* info);

/* ------------------------------------------------------------------------------------------------
 * Synthetic hospitals generated on the device, straight into the compiled form (no float64 tables: configs 4 and 5 of
 * BASELINE.json have 3.65e8 and 7.3e9 patient-day records).  "Bed model": `census` beds, always occupied; after every day the
 * patient of a bed is discharged with probability p_discharge and replaced by a new patient (is_new = 1) in a uniformly
 * drawn ward, otherwise moves to a uniformly drawn ward with probability p_transfer.  Records of a day are ordered by
 * (ward, bed); total_patients_per_ward is the bed count of every occupied ward; the initial block is day 0.
 * A generated handle runs the fused kernels (AMRO_PATH_FAST) only.
 * ---------------------------------------------------------------------------------------------- */
typedef struct amro_hospital_spec {
  int64_t n_wards, census, n_days;
  int64_t seed;
  double  p_discharge, p_transfer;
} amro_hospital_spec;
int amro_topology_generate(const amro_hospital_spec* spec, int device, amro_topology** out);
/* The same hospital as the reference's float64 tables, on the host (small sizes; what the parity tests compile and hand to
 * the oracle).  First call with ward_matrix = tppw = NULL for the sizes; then with ward_matrix [n_rows x 6] and tppw
 * [t_rows x 3] (column-major, ld = rows, *t_rows as returned). */
int amro_hospital_tables(const amro_hospital_spec* spec, int64_t* n_rows, int64_t* t_rows, double* ward_matrix, double* tppw);

/* ------------------------------------------------------------------------------------------------
 * Whole-run simulation on a compiled topology.  Replaces simulate_discrete_model_internal_one
 * (abm_wards.cpp:322-414) fused with total_positive_colonized/_detected (:435-509): state stays in HBM
 * across all days, only the per-(day, simulation) counts -- and, if asked, the trajectory -- leave.
 * ---------------------------------------------------------------------------------------------- */
enum amro_kernel_kind {
  AMRO_KERNEL_NONE = 0,
  AMRO_KERNEL_CODES = 1,   /* kernels_fast.cu: 16-bit state codes (trajectories, half weights, replay)   */
  AMRO_KERNEL_BITS = 2,    /* kernels_bits.cu: state bit-sliced over simulations (summary runs)          */
  AMRO_KERNEL_FLOAT64 = 3  /* kernels_general.cu: the reference's float64 planes                         */
};

typedef struct amro_sim_args {
  int64_t n_sims;              /* S: simulations run by THIS call (this GPU's shard)                      */
  int64_t sim_offset;          /* global index of simulation 0 (Philox counters; keeps results independent
                                  of how an ensemble is sharded over GPUs)                               */
  int64_t seed;                /* the 64-bit Philox key (the reference's `int seed`, sign-extended)       */
  int32_t time_to_detect;
  int32_t testing_schedule_hospitalized;
  int32_t testing_schedule_arrivals;
  int32_t rng_mode;            /* enum amro_rng_mode                                                      */
  int32_t path;                /* enum amro_path                                                          */
  int32_t keep_trajectory;     /* 0: two-day ring of state (summary runs); 1: keep every record's state   */
  int32_t pointers_on_device;  /* 0: every pointer below is host memory; 1: device memory (no copies)     */
  int32_t async;               /* with pointers_on_device: 1 = enqueue on `stream` and return without waiting
                                  (kernel_ms / total_ms are then 0); the caller synchronises the stream      */

  /* parameters [>= S x 6], column stride p_ld */
  const double* parameters;  int64_t p_rows;  int64_t p_ld;
  /* initial probabilities, element (r, s) at ptr[r*rs + s*cs]; rs = cs = 0 broadcasts a scalar
     (simulate_and_collapse, abm_wards.cpp:593-597) */
  const double* icp;  int64_t icp_rs;  int64_t icp_cs;
  const double* idp;  int64_t idp_rs;  int64_t idp_cs;

  /* REPLAY only: uniforms per (row, sim): u_col/u_det [R x S] ld = u_ld; u_icol/u_idet [n_init x S] */
  const double* u_col;  const double* u_det;  int64_t u_ld;
  const double* u_icol; const double* u_idet; int64_t ui_ld;

  /* outputs (any may be NULL) */
  int32_t* counts_colonized;   /* [n_days_out x S] column-major, ld = n_days_out: #rows with C == 1      */
  int32_t* counts_detected;    /* [n_days_out x S]: #rows with D <= 0                                    */
  double*  state_out;          /* [R x 2S] ld = so_ld: C columns then D columns of the reference matrix   */
  int64_t  so_ld;              /*   (requires keep_trajectory)                                            */
  double*  p_out;              /* [R x S] ld = R: transition probability used for every update            */
  double*  day_moments;        /* [n_days_out x 4] ld = n_days_out: per day, over this call's simulations: sum and sum
                                  of squares of the colonised counts, then of the detected counts -- what one
                                  all-reduce over the GPUs of an ensemble needs for simulate_and_collapse           */
  int32_t* pressure_out;       /* FAST path: [n_segments x S] ld = n_segments: colonised count per
                                  ward-day before its step (the reference's sum(W), :99)                  */
  /* distance to data (what a parameter-inference loop needs back from a 64k-particle sweep is one number per
     particle, not the [days x S] counts): distance[s] = sum over days of (colonised(day, s) - observed_colonized[day])^2
     + (detected(day, s) - observed_detected[day])^2; either series may be NULL, NaN entries are days without data */
  const double* observed_colonized;   /* [n_days_out] or NULL                                             */
  const double* observed_detected;    /* [n_days_out] or NULL                                             */
  double*  distance;           /* [S] or NULL                                                              */
  /* summaries on the device (summary_of_total_positive, abm_wards.cpp:526-540, of this run's counts): per day,
     over the simulations of this call: day, mean, population sd and the given quantiles (arma::quantile) */
  const double* quantiles;     /* [n_quantiles] (host memory, also with pointers_on_device) or NULL        */
  int64_t  n_quantiles;
  double*  summary_colonized;  /* [n_days_out x (3 + n_quantiles)] column-major, ld = n_days_out, or NULL  */
  double*  summary_detected;
  void*    stream;             /* cudaStream_t when pointers_on_device, else ignored                      */

  /* filled on return */
  int32_t  path_used;          /* AMRO_PATH_FAST or AMRO_PATH_GENERAL                                     */
  int32_t  launches;           /* kernels launched by this call                                           */
  float    kernel_ms;          /* device time of the simulation kernels (CUDA events)                     */
  float    total_ms;           /* device time incl. copies                                                */
  int32_t  kernel_kind;        /* enum amro_kernel_kind: which kernel stepped the records                 */
  int32_t  grid;               /* CTAs of the fused kernel's launch (0 on the float64 path)               */
  int32_t  ctas_per_team;      /* CTAs that share a group of simulations and meet once per day            */
  int32_t  team_barrier;       /* 0: __syncthreads (one CTA per team), 1: counter in global memory, 2: cluster barrier */
} amro_sim_args;

/* Thread safety: calls on ONE handle are serialised by a mutex inside the handle (they share its device workspaces);
 * different handles are independent.  An asynchronous call (pointers_on_device + async) returns with its work pending on
 * `stream`: later calls on the same handle must be enqueued on the same stream, or ordered after it by the caller.
 * FAST needs sim_offset to be a multiple of 32 (a group of 32 simulations shares the plane words of the daily draws). */
int amro_simulate(amro_topology* topo, amro_sim_args* args);
/* Device time (CUDA events) of the fused kernel of the last n runs on this handle, oldest first (n <= 64): how an
 * asynchronous caller times its runs after synchronising.  Waits for those runs to finish. */
int amro_topology_kernel_times(amro_topology* topo, int n, float* ms);

/* ------------------------------------------------------------------------------------------------
 * One-call equivalents of the reference's Python entry points (host pointers).
 * ---------------------------------------------------------------------------------------------- */

/* The one-call entry points below keep the last few compiled hospitals on the device, keyed by a hash of the CONTENT of
 * ward_matrix / total_patients_per_ward, so that optimisation loops over the parameters (the use the reference's
 * tutorial recommends simulate_and_collapse for) do not recompile the hospital every call.  This frees them. */
void amro_topology_cache_clear(void);
/* The cache keeps at most four hospitals and at most AMRO_CACHE_BYTES (environment, default 8 GiB) of host + device memory,
 * evicting the least recently used; a hospital is never freed while a call is running on it.  Lookups, compiled hospitals
 * held and their footprint (any pointer may be NULL). */
void amro_topology_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries, int64_t* bytes);

/* amro.simulate_discrete_model (amro.cpp:18-91 -> abm_wards.cpp:322).  out is [R x (6+2S)], ld = R. */
int amro_simulate_discrete_model(const double* icp, int64_t i_rows, int64_t i_cols, int64_t i_ld,
                                 const double* idp, int64_t d_rows, int64_t d_cols, int64_t d_ld,
                                 const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                                 const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                                 const double* parameters, int64_t p_rows, int64_t p_cols, int64_t p_ld,
                                 int time_to_detect, int tsh, int tsa, int seed, int device, double* out);

/* amro.simulate_and_collapse (amro.cpp:273-318 -> abm_wards.cpp:566).  out is [n_days_out x 5]; call with
 * out = NULL to get n_days_out in *n_days.  bug_compatible != 0 reproduces the reference's column layout
 * (day, sd colonised, sd detected, 0, 0; abm_wards.cpp:625-628), 0 gives the documented one
 * (day, mean colonised, mean detected, sd colonised, sd detected). */
int amro_simulate_and_collapse(const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                               const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                               double icp, double idp, int64_t initial_patients,
                               double alpha, double beta, double gamma, double rho_hospital,
                               double alpha2, double rho_imported, int64_t n_sims,
                               int64_t time_to_detect, int64_t tsh, int64_t tsa, int seed,
                               int device, int bug_compatible, int64_t* n_days, double* out);
/* The same with the result buffer supplied by a callback once the number of days is known (no shape query, no second pass
 * over ward_matrix): alloc(n_days_out, ctx) returns room for [n_days_out x 5] doubles, or NULL to abort. */
typedef double* (*amro_alloc_fn)(int64_t n_days_out, void* ctx);
int amro_simulate_and_collapse_alloc(const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                                     const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                                     double icp, double idp, int64_t initial_patients,
                                     double alpha, double beta, double gamma, double rho_hospital,
                                     double alpha2, double rho_imported, int64_t n_sims,
                                     int64_t time_to_detect, int64_t tsh, int64_t tsa, int seed,
                                     int device, int bug_compatible, amro_alloc_fn alloc, void* ctx);

/* amro.progress_patients_1_timestep (amro.cpp:157-222 -> abm_wards.cpp:219) and
 * amro.progress_patients_probability_ward_1_timestep (amro.cpp:93-155 -> abm_wards.cpp:70) when
 * tppw == NULL (then total_patients is the ward's N).  ward_matrix [n x (6+2S)] is updated IN PLACE, as
 * the reference does to a borrowed array.  rng_mode REPLAY reads u_col/u_det [n x S] (ld = n);
 * PHILOX uses (seed, call_index) -- the reference instead continues the process-global rand() stream. */
int amro_progress_patients(double* ward_matrix, int64_t n_rows, int64_t n_cols, int64_t ld,
                           const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                           double total_patients,
                           const double* parameters, int64_t p_rows, int64_t p_cols, int64_t p_ld,
                           int64_t n_sims, int time_to_detect, int detect_hosp, int detect_arr,
                           int rng_mode, int64_t seed, int64_t call_index,
                           const double* u_col, const double* u_det,
                           int device, double* p_out);

/* amro.total_positive_colonized / _detected (amro.cpp:224-252 -> abm_wards.cpp:435-509).
 * mode 1: count == 1 over columns 6..6+n_sims-1; mode 0: count <= 0 over columns 6+n_sims..n_cols-1.
 * out is [n_days x n_out_cols]; call with out = NULL to get the shape. */
int amro_total_positive(const double* model, int64_t n_rows, int64_t n_cols, int64_t ld,
                        int64_t n_sims, int mode, int device,
                        int64_t* n_days, int64_t* n_out_cols, double* out);

/* The run's collective as peer stores (the reference has no counterpart: its OpenMP loop over simulations,
 * abm_wards.cpp:92, shares one address space).  Per-day sum and sum of squares over this GPU's simulations of the colonised and
 * detected counts (device arrays [n_days x n_sims], ld_counts, as amro_simulate leaves them with pointers_on_device), stored
 * as [n_days x 4] (ld = n_days) into EVERY destination: dst[i] are device pointers -- this GPU's buffer and its peers'
 * buffers mapped over NVLink (CUDA IPC / torch symmetric memory), n_dst <= 16.  One small kernel on `stream`, no host
 * synchronisation, no rank waits for another; the reader adds the ranks' slots after its own barrier. */
int amro_day_moments_peers(const int32_t* counts_colonized, const int32_t* counts_detected, int64_t n_days, int64_t n_sims,
                           int64_t ld_counts, double* const* dst, int n_dst, int device, void* stream);

/* The same on the device (per-day segmented sort across the simulations; bit-identical to the host routine): worth it from a
 * few hundred thousand values on. */
int amro_summary_of_total_positive_device(const double* positive, int64_t n_days, int64_t n_sims, int64_t ld,
                                          const double* quantiles, int64_t n_q, int device, double* out);

/* amro.summary_of_total_positive (amro.cpp:254-271 -> abm_wards.cpp:526-540): day, mean, population sd and
 * Armadillo-definition quantiles across simulations.  Host arithmetic (tiny post-processing of
 * [days x S] counts); out is [n_days x (3+n_q)]. */
int amro_summary_of_total_positive(const double* positive, int64_t n_days, int64_t n_sims, int64_t ld,
                                   const double* quantiles, int64_t n_q, double* out);

#ifdef __cplusplus
}
#endif
#endif /* AMRO_B200_H */
