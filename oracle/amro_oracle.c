/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the ABM_AMRO ward-level colonisation path.
 * See amro_oracle.h for the role of this file and its parity status (PINNED against the compiled
 * reference and the golden fixtures).  Nothing under abm_amro_b200/ may depend on it.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.  The code is a
 * from-scratch restatement of the algorithm (explicit loops over a column-major array), not a copy of
 * the reference's Armadillo expressions.
 *
 * Floating point: compiled with -ffp-contract=off; every product and sum below is rounded separately,
 * in the operation order Armadillo's expression templates produce on x86-64 without FMA.
 */
#include "amro_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static _Thread_local char g_err[256];
const char* amro_oracle_last_error(void) { return g_err; }
static int fail(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }

#define AT(m, ld, r, c) ((m)[(uint64_t)(r) + (uint64_t)(ld) * (uint64_t)(c)])

/* ---------------------------------------------------------------------------------------------
 * Uniform sources
 * ------------------------------------------------------------------------------------------- */

/* std::mt19937_64 (libstdc++ <random>), used by Armadillo's fill::randu for N > 1
 * (armadillo_bits/arma_rng.hpp:335-345). */
typedef struct { uint64_t x[312]; int idx; } mt64;
static void mt64_seed(mt64* g, uint64_t s) {
  g->x[0] = s;
  for (int i = 1; i < 312; ++i) g->x[i] = 6364136223846793005ULL * (g->x[i - 1] ^ (g->x[i - 1] >> 62)) + (uint64_t)i;
  g->idx = 312;
}
static uint64_t mt64_next(mt64* g) {
  if (g->idx >= 312) {
    const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
    for (int i = 0; i < 312; ++i) {
      uint64_t y = (g->x[i] & UM) | (g->x[(i + 1) % 312] & LM);
      g->x[i] = g->x[(i + 156) % 312] ^ (y >> 1) ^ ((y & 1ULL) ? 0xB5026F5AA96619E9ULL : 0ULL);
    }
    g->idx = 0;
  }
  uint64_t y = g->x[g->idx++];
  y ^= (y >> 29) & 0x5555555555555555ULL;
  y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
  y ^= (y << 37) & 0xFFF7EEE000000000ULL;
  y ^= (y >> 43);
  return y;
}
/* std::uniform_real_distribution<double>(0,1) over a 64-bit engine: generate_canonical<double,53>
 * takes ONE draw, divides by 2^64 and clamps below 1 (libstdc++ bits/random.tcc). */
static double mt64_uniform(mt64* g) {
  double r = (double)mt64_next(g) / 18446744073709551616.0;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

/* scalar arma::randu(): double(rand()) * (1/RAND_MAX)  (armadillo_bits/arma_rng_cxx03.hpp:81-86) */
static double glibc_randu(void) { return (double)rand() * (1.0 / 2147483647.0); }

/* Philox4x32-R, Salmon/Moraes/Dror/Shaw SC'11 (the published algorithm; constants as in Random123). */
void amro_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < rounds; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void amro_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { amro_oracle_philox4x32(ctr, key, 10, out); }
#define AMRO_PLANE_ROUNDS 7 /* RNG spec: the plane words of the daily draws run Philox4x32-7, everything else Philox4x32-10 */

/* RNG spec: number of 32-bit integers r with r * 2^-32 < p, i.e. the Bernoulli(p) threshold on a
 * uniform u = r * 2^-32 in [0,1).  p <= 0 or NaN -> 0 (never), p >= 1 -> 2^32 (always). */
uint64_t amro_oracle_threshold(double p) {
  if (!(p > 0.0)) return 0;
  if (p >= 1.0) return 4294967296ULL;
  return (uint64_t)ceil(p * 4294967296.0);
}

/* RNG spec, daily step (abm_amro_b200/csrc/rng.cuh).  The uniforms of a record are BIT-SLICED over groups of 32
 * consecutive simulations: plane word j (0..13) of (row, group g = sim >> 5) is word (j & 3) of the Philox call with
 * counter (row_lo, row_hi, g, 8 + (j >> 2)) and 7 rounds; bit (sim & 31) of plane j is bit j of the 14-bit draw u14(row, sim).
 * Bernoulli(p) fires iff u14 < B(p) with B(p) = (T30(p) + r16) >> 16, T30(p) = ceil(p * 2^30) clamped to [0, 2^30], and
 * r16 a 16-bit DITHER: the 30-bit threshold is rounded to 14 bits so that E[B(p)] = T30(p) / 2^16 exactly.  The dither is
 * the 16-bit chunk of (stream 6, index, sim) with index = the record's ward-day -- or the all-ones index (one dither per
 * simulation and run) when the probability does not depend on the ward: a record with W = C*w = 1 under a finite force of
 * infection, whose P = coef + 0*F + gamma*h (:108-112). */
uint32_t amro_oracle_philox_chunk(uint64_t seed, uint64_t index, uint64_t sim, uint32_t stream) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)(sim >> 3), stream};
  uint32_t a[4];
  amro_oracle_philox4x32_10(ctr, key, a);
  unsigned k = (unsigned)(sim & 7u);
  return (a[k >> 1] >> (16u * (k & 1u))) & 0xFFFFu;
}
uint32_t amro_oracle_philox_plane(uint64_t seed, uint64_t row, uint64_t group32, unsigned j) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {(uint32_t)row, (uint32_t)(row >> 32), (uint32_t)group32, 8u + (j >> 2)};
  uint32_t a[4];
  amro_oracle_philox4x32(ctr, key, AMRO_PLANE_ROUNDS, a);
  return a[j & 3u];
}
uint32_t amro_oracle_philox_u14(uint64_t seed, uint64_t row, uint64_t sim) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t u = 0;
  for (unsigned c = 0; c < 4; ++c) {
    uint32_t ctr[4] = {(uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 5), 8u + c};
    uint32_t a[4];
    amro_oracle_philox4x32(ctr, key, AMRO_PLANE_ROUNDS, a);
    for (unsigned q = 0; q < 4 && 4 * c + q < 14; ++q) u |= ((a[q] >> (sim & 31u)) & 1u) << (4 * c + q);
  }
  return u;
}
uint32_t amro_oracle_threshold30(double p) {
  if (!(p > 0.0)) return 0;
  if (p >= 1.0) return 1u << 30;
  return (uint32_t)ceil(p * 1073741824.0);
}
uint32_t amro_oracle_bernoulli14(double p, uint32_t r16) { return (amro_oracle_threshold30(p) + r16) >> 16; }

/* RNG spec, initial state: the 32-bit uniform of (row, sim, kind) is (chunk of stream 2*kind) << 16 | chunk of stream
 * 2*kind + 1, compared with T(p) = #{r : r * 2^-32 < p}. */
uint32_t amro_oracle_philox_u32(uint64_t seed, uint64_t row, uint64_t sim, uint32_t kind) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t ctr[4] = {(uint32_t)row, (uint32_t)(row >> 32), (uint32_t)(sim >> 3), 2u * kind};
  uint32_t a[4], b[4];
  amro_oracle_philox4x32_10(ctr, key, a);
  ctr[3] = 2u * kind + 1u;
  amro_oracle_philox4x32_10(ctr, key, b);
  unsigned k = (unsigned)(sim & 7u);
  uint32_t hi = (a[k >> 1] >> (16u * (k & 1u))) & 0xFFFFu;
  uint32_t lo = (b[k >> 1] >> (16u * (k & 1u))) & 0xFFFFu;
  return (hi << 16) | lo;
}

static int stream_pop(amro_oracle_rng* g, double* u) {
  if (g->stream_pos >= g->stream_len) return fail(AMRO_ORACLE_ERUNTIME, "replay buffer exhausted");
  *u = g->stream[g->stream_pos++];
  return 0;
}

static double clamp01(double p) { return p < 0.0 ? 0.0 : (p > 1.0 ? 1.0 : p); }

/* ---------------------------------------------------------------------------------------------
 * a1  progress_patients_probability_ward_1_timestep   (src/amro/abm_wards.cpp:70-159)
 * ------------------------------------------------------------------------------------------- */
int amro_oracle_ward_step(double* wm, uint64_t n_rows, uint64_t n_cols, uint64_t ld,
                          double total_patients,
                          const double* params, uint64_t p_rows, uint64_t p_cols, uint64_t p_ld,
                          uint64_t n_sims, int ttd, int detect_hosp, int detect_arr,
                          amro_oracle_rng* rng, const uint64_t* row_ids,
                          double* p_out, uint64_t p_out_ld) {
  const uint64_t H = 3, Wc = 4, C0 = 6; /* :87-90 */
  double* prob = (double*)malloc(sizeof(double) * (n_rows ? n_rows : 1));
  if (!prob) return fail(AMRO_ORACLE_ERUNTIME, "out of memory");
  /* PHILOX: the plane words of every (row, group of 32 simulations) of this call, drawn once (speed only) */
  uint32_t* planes = NULL;
  const uint64_t g_lo = rng->sim_offset >> 5, n_grp = n_sims ? ((rng->sim_offset + n_sims - 1) >> 5) - g_lo + 1 : 0;
  if (rng->kind == AMRO_ORACLE_RNG_PHILOX && n_rows && n_sims) {
    planes = (uint32_t*)malloc(sizeof(uint32_t) * n_rows * n_grp * 14);
    if (!planes) { free(prob); return fail(AMRO_ORACLE_ERUNTIME, "out of memory"); }
    for (uint64_t i = 0; i < n_rows; ++i)
      for (uint64_t g = 0; g < n_grp; ++g)
        for (unsigned j = 0; j < 14; ++j)
          planes[(i * n_grp + g) * 14 + j] = amro_oracle_philox_plane(rng->seed, row_ids ? row_ids[i] : i, g_lo + g, j);
  }
  for (uint64_t s = 0; s < n_sims; ++s) { /* :92 */
    const uint64_t cc = C0 + s, dc = C0 + s + n_sims;
    /* Armadillo bounds checks: .col(cc) first (:95), then parameters(s, beta) (:98), .col(dc) (:102) */
    if (cc >= n_cols || Wc >= n_cols) { free(planes); free(prob); return fail(AMRO_ORACLE_EINDEX, "Mat::col(): index out of bounds"); }
    if (s >= p_rows || 1 >= p_cols) { free(planes); free(prob); return fail(AMRO_ORACLE_EINDEX, "Mat::operator(): index out of bounds"); }
    if (dc >= n_cols) { free(planes); free(prob); return fail(AMRO_ORACLE_EINDEX, "Mat::col(): index out of bounds"); }
    if (5 >= p_cols) { free(planes); free(prob); return fail(AMRO_ORACLE_EINDEX, "Mat::operator(): index out of bounds"); }
    const double alpha = AT(params, p_ld, s, 0), beta = AT(params, p_ld, s, 1), gamma = AT(params, p_ld, s, 2);
    const double rho_h = AT(params, p_ld, s, 3), alpha2 = AT(params, p_ld, s, 4), rho_i = AT(params, p_ld, s, 5);

    /* :95  W = C .* w, then :99 sum(W) in arrayops::accumulate's two-accumulator order
     * (armadillo_bits/arrayops_meat.hpp:917-936) */
    double acc1 = 0.0, acc2 = 0.0;
    uint64_t i;
    for (i = 0; i < n_rows; ++i) AT(wm, ld, i, cc) = AT(wm, ld, i, cc) * AT(wm, ld, i, Wc);
    for (i = 0; i + 1 < n_rows; i += 2) { acc1 += AT(wm, ld, i, cc); acc2 += AT(wm, ld, i + 1, cc); }
    if (i < n_rows) acc1 += AT(wm, ld, i, cc);
    const double force = (beta / total_patients) * (acc1 + acc2); /* :98-99 */

    const double om_a = 1.0 - alpha, om_a2 = 1.0 - alpha2;
    for (i = 0; i < n_rows; ++i) {
      const double W = AT(wm, ld, i, cc);
      const double det = (AT(wm, ld, i, dc) <= 0.0) ? 1.0 : 0.0;          /* :102 */
      const double coef = ((1.0 - det) * om_a) + (det * om_a2);            /* :103-105 */
      double p = (coef * W) + ((1.0 - W) * force);                         /* :108-109 */
      p = p + (AT(wm, ld, i, H) * gamma);                                  /* :112 */
      prob[i] = p;
    }

    for (i = 0; i < n_rows; ++i) { /* :115 */
      const uint64_t gid = row_ids ? row_ids[i] : i;
      const double p = prob[i], h = AT(wm, ld, i, H);
      double* D = &AT(wm, ld, i, dc);
      int col, tie_kind = rng->kind;
      uint32_t v14 = 0, r16 = 0;
      if (p_out) p_out[gid + p_out_ld * s] = p;
      /* :116-120 colonisation draw, always taken */
      if (tie_kind == AMRO_ORACLE_RNG_PHILOX) {
        /* W (the column still holds C*w of this row) = 1 under a finite force: P does not depend on the ward */
        const int ward_free = (AT(wm, ld, i, cc) == 1.0) && isfinite(force);
        {
          const uint64_t sim = rng->sim_offset + s;
          const uint32_t* pl = planes + (i * n_grp + ((sim >> 5) - g_lo)) * 14;
          for (unsigned j = 0; j < 14; ++j) v14 |= ((pl[j] >> (sim & 31u)) & 1u) << j;   /* = amro_oracle_philox_u14 */
        }
        r16 = amro_oracle_philox_chunk(rng->seed, ward_free ? UINT64_MAX : rng->ward_day, rng->sim_offset + s, 6);
        col = v14 < amro_oracle_bernoulli14(p, r16);
        rng->n_draws++;
      } else {
        double u;
        if (tie_kind == AMRO_ORACLE_RNG_GLIBC) u = glibc_randu();
        else if (tie_kind == AMRO_ORACLE_RNG_STREAM) { if (stream_pop(rng, &u)) { free(planes); free(prob); return AMRO_ORACLE_ERUNTIME; } }
        else u = rng->u_col[gid + rng->u_ld * s];
        if (rng->rec_col) rng->rec_col[gid + rng->rec_ld * s] = u;
        rng->n_draws++;
        col = (u < p);
      }
      AT(wm, ld, i, cc) = col ? 1.0 : 0.0;

      /* :129-155 detection bookkeeping */
      int is_h0 = (h == 0.0), is_h1 = (h == 1.0);
      if (col && (is_h0 || is_h1)) {
        const int flag = is_h0 ? detect_hosp : detect_arr;
        const double rho = is_h0 ? rho_h : rho_i;
        if (*D <= (double)ttd) {
          *D -= 1.0;                                                       /* :131-132 / :143-144 */
        } else {
          int hit = 0;
          if (flag) {                                                      /* short-circuit: draw only if testing today */
            if (tie_kind == AMRO_ORACLE_RNG_PHILOX) {
              hit = v14 < amro_oracle_bernoulli14(clamp01(p) * rho, r16); /* same uniform: P(col and hit) = p*rho */
            } else {
              double u2;
              if (tie_kind == AMRO_ORACLE_RNG_GLIBC) u2 = glibc_randu();
              else if (tie_kind == AMRO_ORACLE_RNG_STREAM) { if (stream_pop(rng, &u2)) { free(planes); free(prob); return AMRO_ORACLE_ERUNTIME; } }
              else u2 = rng->u_det[gid + rng->u_ld * s];
              if (rng->rec_det) rng->rec_det[gid + rng->rec_ld * s] = u2;
              rng->n_draws++;
              hit = (u2 <= rho);                                           /* :134 / :146 */
            }
          }
          *D = hit ? (double)ttd : INFINITY;                               /* :135,138 / :147,150 */
        }
      } else {
        *D = INFINITY;                                                     /* :153-155 */
      }
    }
  }
  free(planes);
  free(prob);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * a2  progress_patients_1_timestep   (src/amro/abm_wards.cpp:219-257)
 * ------------------------------------------------------------------------------------------- */
static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

int amro_oracle_day_step(double* wm, uint64_t n_rows, uint64_t n_cols, uint64_t ld,
                         const double* tppw, uint64_t t_rows, uint64_t t_cols, uint64_t t_ld,
                         const double* params, uint64_t p_rows, uint64_t p_cols, uint64_t p_ld,
                         uint64_t n_sims, int ttd, int detect_hosp, int detect_arr,
                         amro_oracle_rng* rng, const uint64_t* row_ids,
                         double* p_out, uint64_t p_out_ld) {
  if (n_rows == 0) return 0; /* unique() of an empty column: no wards */
  if (1 >= n_cols) return fail(AMRO_ORACLE_EINDEX, "Mat::col(): index out of bounds");
  if (1 >= t_cols && t_rows) return fail(AMRO_ORACLE_EINDEX, "Mat::col(): index out of bounds");
  /* :235 active wards = sorted unique values of column 1 */
  double* wards = (double*)malloc(sizeof(double) * n_rows);
  uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * n_rows);
  uint64_t* gids = (uint64_t*)malloc(sizeof(uint64_t) * n_rows);
  if (!wards || !idx || !gids) { free(wards); free(idx); free(gids); return fail(AMRO_ORACLE_ERUNTIME, "out of memory"); }
  for (uint64_t i = 0; i < n_rows; ++i) wards[i] = AT(wm, ld, i, 1);
  qsort(wards, n_rows, sizeof(double), cmp_double);
  uint64_t n_w = 0;
  for (uint64_t i = 0; i < n_rows; ++i) if (n_w == 0 || wards[i] != wards[n_w - 1]) wards[n_w++] = wards[i];

  int rc = 0;
  for (uint64_t w = 0; w < n_w && rc == 0; ++w) { /* :238 */
    uint64_t n = 0;
    for (uint64_t i = 0; i < n_rows; ++i) if (AT(wm, ld, i, 1) == wards[w]) idx[n++] = i; /* :241 */
    /* :242,:249  N = column 2 of the FIRST tppw row whose ward matches */
    uint64_t hit = t_rows;
    for (uint64_t k = 0; k < t_rows; ++k) if (AT(tppw, t_ld, k, 1) == wards[w]) { hit = k; break; }
    if (hit == t_rows || 2 >= t_cols) { rc = fail(AMRO_ORACLE_EINDEX, "Mat::operator(): index out of bounds"); break; }
    const double N = AT(tppw, t_ld, hit, 2);
    /* :246 gather, :249 step, :253 scatter */
    double* tmp = (double*)malloc(sizeof(double) * n * n_cols);
    if (!tmp) { rc = fail(AMRO_ORACLE_ERUNTIME, "out of memory"); break; }
    for (uint64_t c = 0; c < n_cols; ++c) for (uint64_t i = 0; i < n; ++i) tmp[i + n * c] = AT(wm, ld, idx[i], c);
    for (uint64_t i = 0; i < n; ++i) gids[i] = row_ids ? row_ids[idx[i]] : idx[i];
    rc = amro_oracle_ward_step(tmp, n, n_cols, n, N, params, p_rows, p_cols, p_ld, n_sims, ttd,
                               detect_hosp, detect_arr, rng, gids, p_out, p_out_ld);
    if (rc == 0) for (uint64_t c = 0; c < n_cols; ++c) for (uint64_t i = 0; i < n; ++i) AT(wm, ld, idx[i], c) = tmp[i + n * c];
    rng->ward_day++;
    free(tmp);
  }
  free(wards); free(idx); free(gids);
  return rc;
}

/* ---------------------------------------------------------------------------------------------
 * a3  simulate_discrete_model_internal_one   (src/amro/abm_wards.cpp:322-414)
 * ------------------------------------------------------------------------------------------- */
static int fill_uniform(amro_oracle_rng* rng, double* mem, uint64_t n_rows, uint64_t n_sims, int which /*0 col, 1 det*/) {
  const uint64_t N = n_rows * n_sims;
  if (rng->kind == AMRO_ORACLE_RNG_GLIBC) {
    /* armadillo_bits/arma_rng.hpp:335-345: one element -> scalar draw; otherwise a local mt19937_64
     * seeded with one rand(), filled in memory (column-major) order */
    if (N == 1) { mem[0] = glibc_randu(); }
    else { mt64 g; mt64_seed(&g, (uint64_t)rand()); for (uint64_t i = 0; i < N; ++i) mem[i] = mt64_uniform(&g); }
  } else if (rng->kind == AMRO_ORACLE_RNG_STREAM) {
    for (uint64_t i = 0; i < N; ++i) if (stream_pop(rng, &mem[i])) return AMRO_ORACLE_ERUNTIME;
  } else if (rng->kind == AMRO_ORACLE_RNG_TENSOR) {
    const double* src = which ? rng->u_idet : rng->u_icol;
    for (uint64_t s = 0; s < n_sims; ++s) for (uint64_t r = 0; r < n_rows; ++r) mem[r + n_rows * s] = src[r + rng->ui_ld * s];
  }
  double* rec = which ? rng->rec_idet : rng->rec_icol;
  if (rec && rng->kind != AMRO_ORACLE_RNG_PHILOX)
    for (uint64_t s = 0; s < n_sims; ++s) for (uint64_t r = 0; r < n_rows; ++r) rec[r + rng->reci_ld * s] = mem[r + n_rows * s];
  rng->n_draws += N;
  return 0;
}

int amro_oracle_simulate(const double* icp, uint64_t i_rows, uint64_t i_cols, uint64_t i_ld,
                         const double* idp, uint64_t d_rows, uint64_t d_cols, uint64_t d_ld,
                         const double* ward_matrix, uint64_t R, uint64_t w_cols, uint64_t w_ld,
                         const double* tppw, uint64_t t_rows, uint64_t t_cols, uint64_t t_ld,
                         const double* params, uint64_t p_rows, uint64_t p_cols, uint64_t p_ld,
                         int ttd, int tsh, int tsa, int seed,
                         amro_oracle_rng* rng, double* out, double* p_out, uint64_t p_out_ld) {
  const uint64_t S = i_cols, ncol = w_cols + 2 * S, C0 = 6;
  if (rng->kind == AMRO_ORACLE_RNG_GLIBC) srand((unsigned)seed); /* :333 -> arma_rng_cxx03.hpp:48-52 */
  if (rng->kind == AMRO_ORACLE_RNG_PHILOX) rng->seed = (uint64_t)(int64_t)seed;
  rng->ward_day = 0;

  if (t_rows == 0 || t_cols == 0) return fail(AMRO_ORACLE_ERUNTIME, "max(): object has no elements");
  double dmax = AT(tppw, t_ld, 0, 0); /* :356 */
  for (uint64_t k = 1; k < t_rows; ++k) if (AT(tppw, t_ld, k, 0) > dmax) dmax = AT(tppw, t_ld, k, 0);
  if (!(dmax >= 0.0)) return fail(AMRO_ORACLE_ERUNTIME, "negative or NaN day in total_patients_per_ward");
  const uint64_t total_days = (uint64_t)dmax;

  /* :359-360 zero matrix, detection columns = +inf */
  for (uint64_t c = 0; c < ncol; ++c) for (uint64_t r = 0; r < R; ++r) AT(out, R, r, c) = 0.0;
  for (uint64_t s = 0; s < S; ++s) {
    if (C0 + S + s >= ncol) return fail(AMRO_ORACLE_EINDEX, "Mat::cols(): indices out of bounds or incorrectly used");
    for (uint64_t r = 0; r < R; ++r) AT(out, R, r, C0 + S + s) = INFINITY;
  }

  /* :363-372 initial state */
  double* u1 = (double*)malloc(sizeof(double) * (i_rows * S + 1));
  double* u2 = (double*)malloc(sizeof(double) * (d_rows * S + 1));
  if (!u1 || !u2) { free(u1); free(u2); return fail(AMRO_ORACLE_ERUNTIME, "out of memory"); }
  int rc = fill_uniform(rng, u1, i_rows, S, 0);
  if (rc == 0 && (i_rows == 0 || S == 0 || i_rows > R || C0 + S > ncol))
    rc = fail(AMRO_ORACLE_EINDEX, "Mat::submat(): indices out of bounds or incorrectly used");
  if (rc == 0) rc = fill_uniform(rng, u2, d_rows, S, 1);
  if (rc == 0 && d_cols != S) rc = fail(AMRO_ORACLE_ERUNTIME, "relational operator: incompatible matrix dimensions");
  if (rc == 0 && d_rows != i_rows) rc = fail(AMRO_ORACLE_ERUNTIME, "element-wise multiplication: incompatible matrix dimensions");
  if (rc == 0 && C0 + 2 * S != ncol) rc = fail(AMRO_ORACLE_ERUNTIME, "copy into submatrix: incompatible matrix dimensions");
  if (rc) { free(u1); free(u2); return rc; }
  for (uint64_t s = 0; s < S; ++s) for (uint64_t r = 0; r < i_rows; ++r) {
    int c0, d0;
    if (rng->kind == AMRO_ORACLE_RNG_PHILOX) {
      c0 = (uint64_t)amro_oracle_philox_u32(rng->seed, r, rng->sim_offset + s, 1) < amro_oracle_threshold(AT(icp, i_ld, r, s));
      d0 = (uint64_t)amro_oracle_philox_u32(rng->seed, r, rng->sim_offset + s, 2) < amro_oracle_threshold(AT(idp, d_ld, r, s));
    } else {
      c0 = u1[r + i_rows * s] < AT(icp, i_ld, r, s);
      d0 = u2[r + d_rows * s] < AT(idp, d_ld, r, s);
    }
    AT(out, R, r, C0 + s) = c0 ? 1.0 : 0.0;
    AT(out, R, r, C0 + S + s) = (d0 && c0) ? 1.0 : 0.0; /* :372 a 0/1 FLAG stored where a countdown is read later */
  }
  free(u1); free(u2);

  /* :375 topology columns (requires exactly 6) */
  if (w_cols != C0) return fail(AMRO_ORACLE_ERUNTIME, "copy into submatrix: incompatible matrix dimensions");
  for (uint64_t c = 0; c < C0; ++c) for (uint64_t r = 0; r < R; ++r) AT(out, R, r, c) = AT(ward_matrix, w_ld, r, c);

  if (tsh == 0 || tsa == 0) return fail(AMRO_ORACLE_ERUNTIME, "integer modulo by zero (testing schedule == 0)");
  uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (R ? R : 1));
  if (!idx) return fail(AMRO_ORACLE_ERUNTIME, "out of memory");
  rc = 0;
  for (uint64_t day = 0; day <= total_days && rc == 0; ++day) { /* :381 */
    /* :384 the day's rows of total_patients_per_ward, in order */
    uint64_t nt = 0;
    for (uint64_t k = 0; k < t_rows; ++k) if (AT(tppw, t_ld, k, 0) == (double)day) nt++;
    double* tday = (double*)malloc(sizeof(double) * (nt * t_cols + 1));
    uint64_t j = 0;
    for (uint64_t k = 0; k < t_rows; ++k) if (AT(tppw, t_ld, k, 0) == (double)day) {
      for (uint64_t c = 0; c < t_cols; ++c) tday[j + nt * c] = AT(tppw, t_ld, k, c);
      j++;
    }
    /* :387-388 the day's rows of the state matrix */
    uint64_t n = 0;
    for (uint64_t r = 0; r < R; ++r) if (AT(out, R, r, 0) == (double)day) idx[n++] = r;
    double* wd = (double*)malloc(sizeof(double) * (n * ncol + 1));
    for (uint64_t c = 0; c < ncol; ++c) for (uint64_t i = 0; i < n; ++i) wd[i + n * c] = AT(out, R, idx[i], c);
    /* :391-392 day is unsigned 64-bit, the schedules are int -> converted to unsigned 64-bit */
    const int th = (day % (uint64_t)(int64_t)tsh) == 0, ta = (day % (uint64_t)(int64_t)tsa) == 0;
    rc = amro_oracle_day_step(wd, n, ncol, n, tday, nt, t_cols, nt, params, p_rows, p_cols, p_ld, S, ttd, th, ta,
                              rng, idx, p_out, p_out_ld); /* :395 */
    if (rc == 0) {
      for (uint64_t c = 0; c < ncol; ++c) for (uint64_t i = 0; i < n; ++i) AT(out, R, idx[i], c) = wd[i + n * c]; /* :399 */
      /* :402-407 propagate post-step state along next_day > 0, sources in ascending row order (last wins),
       * values taken from the day's copy (so a same-day target does not feed forward) */
      for (uint64_t i = 0; i < n && rc == 0; ++i) {
        const double nd = wd[i + n * 5];
        if (nd > 0.0) {
          const uint64_t t = (nd >= 18446744073709551616.0) ? UINT64_MAX : (uint64_t)nd;
          if (t >= R) { rc = fail(AMRO_ORACLE_EINDEX, "Mat::elem(): index out of bounds"); break; }
        }
      }
      if (rc == 0) for (uint64_t i = 0; i < n; ++i) {
        const double nd = wd[i + n * 5];
        if (nd > 0.0) {
          const uint64_t t = (uint64_t)nd;
          for (uint64_t s = 0; s < 2 * S; ++s) AT(out, R, t, C0 + s) = wd[i + n * (C0 + s)];
        }
      }
    }
    free(tday); free(wd);
  }
  free(idx);
  return rc;
}

/* ---------------------------------------------------------------------------------------------
 * a4  total_positive / _colonized / _detected   (src/amro/abm_wards.cpp:435-509)
 * ------------------------------------------------------------------------------------------- */
int64_t amro_oracle_total_positive(const double* m, uint64_t n_rows, uint64_t n_cols, uint64_t ld,
                                   uint64_t col_init, uint64_t col_end, int mode, double* out) {
  if (n_rows == 0 || n_cols == 0) return fail(AMRO_ORACLE_ERUNTIME, "max(): object has no elements");
  double dmax = AT(m, ld, 0, 0); /* :440 */
  for (uint64_t r = 1; r < n_rows; ++r) if (AT(m, ld, r, 0) > dmax) dmax = AT(m, ld, r, 0);
  if (!(dmax >= 0.0)) return fail(AMRO_ORACLE_ERUNTIME, "negative or NaN day");
  const uint64_t days = (uint64_t)dmax + 1;
  if (col_end >= n_cols || col_init > col_end) return fail(AMRO_ORACLE_EINDEX, "Mat::cols(): indices out of bounds or incorrectly used");
  if (!out) return (int64_t)days;
  const uint64_t nc = col_end - col_init + 1;
  for (uint64_t d = 0; d < days; ++d) for (uint64_t c = 0; c < nc; ++c) {
    double cnt = 0.0;
    for (uint64_t r = 0; r < n_rows; ++r) if (AT(m, ld, r, 0) == (double)d) {
      const double v = AT(m, ld, r, col_init + c);
      if (mode == 0 ? (v <= 0.0) : (mode == 1 ? (v == 1.0) : 0)) cnt += 1.0; /* :451-455 */
      else if (mode != 0 && mode != 1) cnt += v;                               /* :450 plain conversion */
    }
    AT(out, days, d, c) = cnt;
  }
  return (int64_t)days;
}

/* mean(X,1) and stddev(X,1,1) with Armadillo's operation order
 * (armadillo_bits/op_mean_meat.hpp:105-133, op_var_meat.hpp:85-106,175-219, arrayops_meat.hpp:917-936) */
void amro_oracle_row_mean_sd(const double* pos, uint64_t n_days, uint64_t n_sims, uint64_t ld,
                             double* mean_out, double* sd_out) {
  for (uint64_t d = 0; d < n_days; ++d) {
    if (mean_out) {
      double acc = 0.0;
      for (uint64_t s = 0; s < n_sims; ++s) acc += AT(pos, ld, d, s);
      mean_out[d] = acc / (double)n_sims;
    }
    if (sd_out) {
      double var = 0.0;
      if (n_sims >= 2) {
        double a1 = 0.0, a2 = 0.0;
        uint64_t i;
        for (i = 0; i + 1 < n_sims; i += 2) { a1 += AT(pos, ld, d, i); a2 += AT(pos, ld, d, i + 1); }
        if (i < n_sims) a1 += AT(pos, ld, d, i);
        const double m = (a1 + a2) / (double)n_sims;
        double q = 0.0, l = 0.0;
        for (i = 0; i + 1 < n_sims; i += 2) {
          const double ti = m - AT(pos, ld, d, i), tj = m - AT(pos, ld, d, i + 1);
          q += (ti * ti) + (tj * tj);
          l += ti + tj;
        }
        if (i < n_sims) { const double ti = m - AT(pos, ld, d, i); q += ti * ti; l += ti; }
        var = (q - (l * l) / (double)n_sims) / (double)n_sims;
      }
      sd_out[d] = sqrt(var);
    }
  }
}

/* a6  summary_of_total_positive (abm_wards.cpp:526-540); quantiles follow Armadillo's
 * glue_quantile::worker (Hyndman & Fan definition 5, armadillo_bits/glue_quantile_meat.hpp:22-80) */
int amro_oracle_summary(const double* pos, uint64_t n_days, uint64_t n_sims, uint64_t ld,
                        const double* q, uint64_t n_q, double* out) {
  double* mean = (double*)malloc(sizeof(double) * (n_days + 1));
  double* sd = (double*)malloc(sizeof(double) * (n_days + 1));
  double* y = (double*)malloc(sizeof(double) * (n_sims + 1));
  if (!mean || !sd || !y) { free(mean); free(sd); free(y); return fail(AMRO_ORACLE_ERUNTIME, "out of memory"); }
  amro_oracle_row_mean_sd(pos, n_days, n_sims, ld, mean, sd);
  const double N = (double)n_sims, pmin = (1.0 - 0.5) / N, pmax = (N - 0.5) / N;
  for (uint64_t d = 0; d < n_days; ++d) {
    AT(out, n_days, d, 0) = (double)d;
    AT(out, n_days, d, 1) = mean[d];
    AT(out, n_days, d, 2) = sd[d];
    for (uint64_t s = 0; s < n_sims; ++s) y[s] = AT(pos, ld, d, s);
    qsort(y, n_sims, sizeof(double), cmp_double);
    for (uint64_t i = 0; i < n_q; ++i) {
      const double p = q[i];
      double v;
      if (p < pmin) v = (p < 0.0) ? -INFINITY : y[0];
      else if (p > pmax) v = (p > 1.0) ? INFINITY : y[n_sims - 1];
      else {
        const uint64_t k = (uint64_t)floor(N * p + 0.5);
        const double pk = ((double)k - 0.5) / N;
        const double w = (p - pk) * N;
        v = ((1.0 - w) * y[k - 1]) + (w * y[k]);
      }
      AT(out, n_days, d, 3 + i) = v;
    }
  }
  free(mean); free(sd); free(y);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * a5  simulate_and_collapse   (src/amro/abm_wards.cpp:566-631)
 * ------------------------------------------------------------------------------------------- */
int64_t amro_oracle_simulate_and_collapse(const double* ward_matrix, uint64_t R, uint64_t w_cols, uint64_t w_ld,
                                          const double* tppw, uint64_t t_rows, uint64_t t_cols, uint64_t t_ld,
                                          double icp, double idp, uint64_t initial_patients,
                                          double alpha, double beta, double gamma, double rho_hospital,
                                          double alpha2, double rho_imported, uint64_t n_sims,
                                          uint64_t ttd, uint64_t tsh, uint64_t tsa, int seed,
                                          amro_oracle_rng* rng, double* out) {
  const uint64_t S = n_sims;
  if (R == 0 || w_cols == 0) return fail(AMRO_ORACLE_ERUNTIME, "max(): object has no elements");
  double dmax = AT(ward_matrix, w_ld, 0, 0);
  for (uint64_t r = 1; r < R; ++r) if (AT(ward_matrix, w_ld, r, 0) > dmax) dmax = AT(ward_matrix, w_ld, r, 0);
  if (!(dmax >= 0.0)) return fail(AMRO_ORACLE_ERUNTIME, "negative or NaN day");
  const uint64_t days = (uint64_t)dmax + 1;
  if (!out) return (int64_t)days;

  double* par = (double*)malloc(sizeof(double) * (S * 6 + 1));          /* :584-590 */
  double* ic = (double*)malloc(sizeof(double) * (initial_patients * S + 1)); /* :593-597 */
  double* id = (double*)malloc(sizeof(double) * (initial_patients * S + 1));
  double* full = (double*)malloc(sizeof(double) * (R * (w_cols + 2 * S) + 1));
  double* colz = (double*)malloc(sizeof(double) * (days * S + 1));
  double* detz = (double*)malloc(sizeof(double) * (days * S + 1));
  double* sdc = (double*)malloc(sizeof(double) * (days + 1));
  double* sdd = (double*)malloc(sizeof(double) * (days + 1));
  int64_t rc = 0;
  if (!par || !ic || !id || !full || !colz || !detz || !sdc || !sdd) rc = fail(AMRO_ORACLE_ERUNTIME, "out of memory");
  if (rc == 0) {
    const double v[6] = {alpha, beta, gamma, rho_hospital, alpha2, rho_imported};
    for (uint64_t c = 0; c < 6; ++c) for (uint64_t s = 0; s < S; ++s) par[s + S * c] = v[c];
    for (uint64_t i = 0; i < initial_patients * S; ++i) { ic[i] = icp; id[i] = idp; }
    rc = amro_oracle_simulate(ic, initial_patients, S, initial_patients, id, initial_patients, S, initial_patients,
                              ward_matrix, R, w_cols, w_ld, tppw, t_rows, t_cols, t_ld, par, S, 6, S,
                              (int)ttd, (int)tsh, (int)tsa, seed, rng, full, NULL, 0); /* :600-610 */
  }
  if (rc == 0) {
    const uint64_t nc = w_cols + 2 * S;
    int64_t d1 = amro_oracle_total_positive(full, R, nc, R, 6, 6 + S - 1, 1, colz);       /* :613 */
    int64_t d2 = amro_oracle_total_positive(full, R, nc, R, 6 + S, nc - 1, 0, detz);      /* :614 */
    if (d1 < 0 || d2 < 0) rc = d1 < 0 ? d1 : d2;
  }
  if (rc == 0) {
    amro_oracle_row_mean_sd(colz, days, S, days, NULL, sdc); /* :617-620 */
    amro_oracle_row_mean_sd(detz, days, S, days, NULL, sdd);
    for (uint64_t d = 0; d < days; ++d) {
      AT(out, days, d, 0) = (double)d;   /* :624 */
      AT(out, days, d, 1) = sdc[d];      /* :625 then overwritten by :627 -> sd, not mean */
      AT(out, days, d, 2) = sdd[d];      /* :626 then overwritten by :628 */
      AT(out, days, d, 3) = 0.0;         /* never written (zero-initialised arma::mat) */
      AT(out, days, d, 4) = 0.0;
    }
    rc = (int64_t)days;
  }
  free(par); free(ic); free(id); free(full); free(colz); free(detz); free(sdc); free(sdd);
  return rc;
}
