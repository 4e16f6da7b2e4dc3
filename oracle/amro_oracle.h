/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the ABM_AMRO ward-level colonisation path.
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The shipped path (abm_amro_b200/) never links,
 * imports or calls anything in oracle/ and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement bit-for-bit
 * against the unmodified reference compiled from /root/reference (oracle/Makefile `ref` target) and,
 * where /root/reference is absent (GPU box), against the golden fixtures under tests/golden/ that
 * tools/make_golden.py generated from that compiled reference.
 *
 * All matrices are float64, column-major (Armadillo / numpy order='F'), element (r,c) at r + ld*c.
 * State matrix columns (reference src/amro/abm_wards.cpp:87-90,335-338):
 *   0 day | 1 ward | 2 MRN | 3 is_new h | 4 weight w | 5 next_day | 6..6+S-1 C_s | 6+S..6+2S-1 D_s
 */
#ifndef AMRO_ORACLE_H
#define AMRO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  AMRO_ORACLE_RNG_GLIBC  = 0, /* srand/rand + mt19937_64 fill: the reference's native stream          */
  AMRO_ORACLE_RNG_STREAM = 1, /* flat buffer popped in the reference's call order (ARMA_RNG_ALT build) */
  AMRO_ORACLE_RNG_TENSOR = 2, /* per-(row,sim) uniforms: what the GPU replay mode consumes             */
  AMRO_ORACLE_RNG_PHILOX = 3  /* the GPU's free-running counter-based stream (DESIGN.md "RNG spec")    */
};

enum { AMRO_ORACLE_OK = 0, AMRO_ORACLE_EINDEX = -1, AMRO_ORACLE_ERUNTIME = -2 };

typedef struct amro_oracle_rng {
  int32_t  kind;
  int32_t  _pad;
  /* STREAM */
  const double* stream;
  uint64_t stream_len;
  uint64_t stream_pos;
  /* TENSOR: u_col/u_det are [R x S] (ld = u_ld), u_icol/u_idet are [n_init x S] (ld = ui_ld) */
  const double* u_col;
  const double* u_det;
  uint64_t u_ld;
  const double* u_icol;
  const double* u_idet;
  uint64_t ui_ld;
  /* PHILOX */
  uint64_t seed;
  uint64_t sim_offset;
  uint64_t ward_day;   /* index of the ward-day (ward step) being processed: counts ward steps in the reference's visiting
                        * order (day ascending, ward ascending); reset by simulate, advanced by the day step */
  /* statistics */
  uint64_t n_draws;
  /* optional recording of the uniform consumed at each slot (GLIBC / STREAM kinds); untouched slots
   * keep whatever the caller put there (tests pre-fill with NaN) */
  double*  rec_col;
  double*  rec_det;
  uint64_t rec_ld;
  double*  rec_icol;
  double*  rec_idet;
  uint64_t reci_ld;
} amro_oracle_rng;

const char* amro_oracle_last_error(void);

/* Philox4x32-10 (Salmon et al. 2011) and the probability -> 33-bit threshold map of the RNG spec. */
void     amro_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void     amro_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
uint64_t amro_oracle_threshold(double p);
/* the 32-bit uniform integer the spec assigns to (row, global sim, draw kind 0=step 1=init-col 2=init-det) */
uint32_t amro_oracle_philox_u32(uint64_t seed, uint64_t row, uint64_t sim, uint32_t kind);
/* daily step: 16-bit chunk of (index, sim, stream), 30-bit threshold and its dithered 14-bit rounding */
uint32_t amro_oracle_philox_chunk(uint64_t seed, uint64_t index, uint64_t sim, uint32_t stream);
/* bit-sliced daily uniforms: plane word j (0..13) of (row, group of 32 simulations), and the 14-bit draw of (row, sim) */
uint32_t amro_oracle_philox_plane(uint64_t seed, uint64_t row, uint64_t group32, unsigned j);
uint32_t amro_oracle_philox_u14(uint64_t seed, uint64_t row, uint64_t sim);
uint32_t amro_oracle_threshold30(double p);
uint32_t amro_oracle_bernoulli14(double p, uint32_t r16);

/* reference progress_patients_probability_ward_1_timestep, abm_wards.cpp:70-159.
 * row_ids[i] = global row index of local row i (NULL = identity); p_out (optional) receives the
 * transition probability P of every (row, sim) at p_out[row_id + p_out_ld*sim]. */
int amro_oracle_ward_step(double* wm, uint64_t n_rows, uint64_t n_cols, uint64_t ld,
                          double total_patients,
                          const double* params, uint64_t p_rows, uint64_t p_cols, uint64_t p_ld,
                          uint64_t n_sims, int time_to_detect, int detect_hosp, int detect_arr,
                          amro_oracle_rng* rng, const uint64_t* row_ids,
                          double* p_out, uint64_t p_out_ld);

/* reference progress_patients_1_timestep, abm_wards.cpp:219-257 */
int amro_oracle_day_step(double* wm, uint64_t n_rows, uint64_t n_cols, uint64_t ld,
                         const double* tppw, uint64_t t_rows, uint64_t t_cols, uint64_t t_ld,
                         const double* params, uint64_t p_rows, uint64_t p_cols, uint64_t p_ld,
                         uint64_t n_sims, int time_to_detect, int detect_hosp, int detect_arr,
                         amro_oracle_rng* rng, const uint64_t* row_ids,
                         double* p_out, uint64_t p_out_ld);

/* reference simulate_discrete_model_internal_one, abm_wards.cpp:322-414.
 * out is [R x (6+2S)] with ld = R, allocated by the caller. */
int amro_oracle_simulate(const double* icp, uint64_t i_rows, uint64_t i_cols, uint64_t i_ld,
                         const double* idp, uint64_t d_rows, uint64_t d_cols, uint64_t d_ld,
                         const double* ward_matrix, uint64_t R, uint64_t w_cols, uint64_t w_ld,
                         const double* tppw, uint64_t t_rows, uint64_t t_cols, uint64_t t_ld,
                         const double* params, uint64_t p_rows, uint64_t p_cols, uint64_t p_ld,
                         int time_to_detect, int tsh, int tsa, int seed,
                         amro_oracle_rng* rng, double* out, double* p_out, uint64_t p_out_ld);

/* reference total_positive (abm_wards.cpp:435-464): mode 1 counts ==1, mode 0 counts <=0, over columns
 * col_init..col_end.  Returns the number of days (max(col0)+1) or a negative error; out (may be NULL to
 * query the size) is [days x (col_end-col_init+1)] with ld = days. */
int64_t amro_oracle_total_positive(const double* m, uint64_t n_rows, uint64_t n_cols, uint64_t ld,
                                   uint64_t col_init, uint64_t col_end, int mode, double* out);

/* reference summary_of_total_positive (abm_wards.cpp:526-540): out is [n_days x (3+n_q)], ld = n_days */
int amro_oracle_summary(const double* pos, uint64_t n_days, uint64_t n_sims, uint64_t ld,
                        const double* q, uint64_t n_q, double* out);

/* row statistics with Armadillo's exact operation order (mean(X,1), stddev(X,1,1)) */
void amro_oracle_row_mean_sd(const double* pos, uint64_t n_days, uint64_t n_sims, uint64_t ld,
                             double* mean_out, double* sd_out);

/* reference simulate_and_collapse (abm_wards.cpp:566-631) incl. its column bug: out is [n_days x 5];
 * returns n_days (>=0) or a negative error; pass out = NULL first to query n_days. */
int64_t amro_oracle_simulate_and_collapse(const double* ward_matrix, uint64_t R, uint64_t w_cols, uint64_t w_ld,
                                          const double* tppw, uint64_t t_rows, uint64_t t_cols, uint64_t t_ld,
                                          double icp, double idp, uint64_t initial_patients,
                                          double alpha, double beta, double gamma, double rho_hospital,
                                          double alpha2, double rho_imported, uint64_t n_sims,
                                          uint64_t time_to_detect, uint64_t tsh, uint64_t tsa, int seed,
                                          amro_oracle_rng* rng, double* out);

#ifdef __cplusplus
}
#endif
#endif
