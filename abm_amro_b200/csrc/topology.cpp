// Host-side topology compiler (see topology.hpp).
#include "topology.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <atomic>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <unordered_map>

#include "common.hpp"

namespace amro {
namespace {

inline double at(const double* m, int64_t ld, int64_t r, int64_t c) { return m[r + ld * c]; }

// doubles compare equal <=> same key (folds -0.0 onto +0.0; NaN never reaches here)
inline uint64_t dkey(double x) {
  x += 0.0;
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return u;
}
struct PairHash {
  size_t operator()(const std::pair<uint64_t, uint64_t>& k) const {
    uint64_t h = k.first * 0x9E3779B97F4A7C15ull;
    h ^= (k.second + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};

constexpr int64_t kMaxDays = 10'000'000;

// Hospitals of millions of rows are compiled by a few threads: every phase below that is a loop over rows or positions is cut
// into contiguous blocks (results do not depend on the number of threads).  AMRO_COMPILE_THREADS caps the threads (1 = serial),
// AMRO_COMPILE_PAR_MIN_ROWS sets the size below which one thread does everything (tests force the threaded path with it).
int compile_threads(int64_t n) {
  const char* e_min = std::getenv("AMRO_COMPILE_PAR_MIN_ROWS");
  const char* e_cap = std::getenv("AMRO_COMPILE_THREADS");
  const int64_t min_rows = e_min ? std::atoll(e_min) : (int64_t)200'000;
  const int cap = e_cap ? std::max(1, std::atoi(e_cap)) : 16;
  if (n < std::max<int64_t>(min_rows, 2)) return 1;
  const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
  return (int)std::max<int64_t>(1, std::min<int64_t>(std::min(hw, cap), n / 2));
}
// scratch array of n elements set to `v` by n_thr threads (a std::vector would be filled -- and its pages touched -- by one)
template <class T>
std::unique_ptr<T[]> filled(int64_t n, T v, int n_thr);
// f(lo, hi, block) on n_thr contiguous blocks of [0, n); exceptions of a block are rethrown (the first block's first)
template <class F>
void parallel_blocks(int64_t n, int n_thr, F f) {
  if (n_thr <= 1) { f((int64_t)0, n, 0); return; }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> err((size_t)n_thr);
  for (int t = 0; t < n_thr; ++t)
    pool.emplace_back([&, t] {
      try { f(n * t / n_thr, n * (t + 1) / n_thr, t); } catch (...) { err[(size_t)t] = std::current_exception(); }
    });
  for (auto& th : pool) th.join();
  for (auto& e : err) if (e) std::rethrow_exception(e);
}

// integer-valued, in [0, limit] -> the integer; else -1   (the reference compares `col == day` with an
// unsigned loop counter converted to double, abm_wards.cpp:384,387,447)
inline int64_t as_day(double v, int64_t limit) {
  if (!(v >= 0.0) || v > (double)limit) return -1;
  const int64_t i = (int64_t)v;
  return ((double)i == v) ? i : -1;
}

template <class T>
std::unique_ptr<T[]> filled(int64_t n, T v, int n_thr) {
  std::unique_ptr<T[]> a(new T[(size_t)std::max<int64_t>(n, 1)]);
  T* p = a.get();
  parallel_blocks(n, n_thr, [&](int64_t lo, int64_t hi, int) { std::fill(p + lo, p + hi, v); });
  return a;
}

}  // namespace

void compile_topology(const double* wm, int64_t R, int64_t n_cols, int64_t ld,
                      const double* tp, int64_t K, int64_t t_cols, int64_t t_ld,
                      int64_t n_init, HostTopology& T) {
  if (R < 0 || K < 0 || n_init < 0) fail(AMRO_EINVAL, "negative size");
  // abm_wards.cpp:356  max() of an empty column throws in Armadillo
  if (K == 0 || t_cols == 0) fail(AMRO_ERUNTIME, "max(): object has no elements");
  double dmax = at(tp, t_ld, 0, 0);
  for (int64_t k = 1; k < K; ++k) dmax = std::max(dmax, at(tp, t_ld, k, 0));
  if (!(dmax >= 0.0) || dmax > (double)kMaxDays)
    fail(AMRO_EINVAL, "total_patients_per_ward: largest day must be in [0, %lld]", (long long)kMaxDays);
  // abm_wards.cpp:365  submat(0, 6, n_init-1, ...) needs 1 <= n_init <= R
  if (n_init < 1 || n_init > R) fail(AMRO_EINDEX, "Mat::submat(): indices out of bounds or incorrectly used");
  // abm_wards.cpp:375  copying ward_matrix into the first 6 columns needs exactly 6 columns
  if (n_cols != 6) fail(AMRO_ERUNTIME, "copy into submatrix: incompatible matrix dimensions: %lldx6 and %lldx%lld",
                        (long long)R, (long long)R, (long long)n_cols);

  T = HostTopology();
  T.R = R;
  T.n_init = n_init;
  T.total_days = (int64_t)dmax;
  const int64_t D = T.total_days + 1;

  // ---- per original row: day, totals day, h, w --------------------------------------------------
  const std::unique_ptr<int64_t[]> row_day_a(new int64_t[(size_t)std::max<int64_t>(R, 1)]);   // every element is written in the row pass
  int64_t* const row_day = row_day_a.get();
  T.h.resize(R);
  T.w.resize(R);
  T.count_day.resize(R);   // every element is written below
  const int n_thr = compile_threads(R);
  double cmax = R ? at(wm, ld, 0, 0) : 0.0;
  {
    std::vector<double> part((size_t)n_thr, cmax);
    parallel_blocks(R, n_thr, [&](int64_t lo, int64_t hi, int t) {
      double m = part[(size_t)t];
      for (int64_t r = lo; r < hi; ++r) m = std::max(m, at(wm, ld, r, 0));
      part[(size_t)t] = m;
    });
    for (double m : part) cmax = std::max(cmax, m);
  }
  if (!(cmax >= 0.0) || cmax > (double)kMaxDays)
    fail(AMRO_EINVAL, "ward_matrix: largest day must be in [0, %lld]", (long long)kMaxDays);
  T.n_days_out = (int64_t)cmax + 1;
  // Sortedness: the stepped rows, in row order, must not go back in (day, ward), and no stepped row may follow a row that is
  // never stepped.  Every block reports its own verdict and its first / last stepped row; the blocks are stitched afterwards.
  struct Block {
    int64_t n_stepped = 0, first_unstepped = -1, last_stepped = -1, nan_row = -1;
    bool ordered = true, unit = true, dyadic = true, half = false, binary_h = true;
    int64_t first_day = -1, last_day = -1;
    double first_ward = 0.0, last_ward = 0.0;
  };
  std::vector<Block> blocks((size_t)n_thr);
  parallel_blocks(R, n_thr, [&](int64_t lo, int64_t hi, int t) {
    Block b;
    for (int64_t r = lo; r < hi; ++r) {
      const double dv = at(wm, ld, r, 0), wv = at(wm, ld, r, 1);
      const int64_t d = as_day(dv, T.total_days);
      row_day[r] = d;
      T.count_day[r] = (int32_t)as_day(dv, T.n_days_out - 1);
      T.h[r] = at(wm, ld, r, 3);
      T.w[r] = at(wm, ld, r, 4);
      if (d >= 0) {
        if (wv != wv) { if (b.nan_row < 0) b.nan_row = r; continue; }
        if (b.last_stepped >= 0) {
          if (d < b.last_day || (d == b.last_day && wv < b.last_ward)) b.ordered = false;
        } else { b.first_day = d; b.first_ward = wv; }
        b.last_stepped = r; b.last_day = d; b.last_ward = wv;
        b.n_stepped++;
        if (T.w[r] != 1.0) b.unit = false;
        if (T.w[r] == 0.5) b.half = true;
        else if (T.w[r] != 1.0) b.dyadic = false;
        if (!(T.h[r] == 0.0 || T.h[r] == 1.0)) b.binary_h = false;
      } else if (b.first_unstepped < 0) {
        b.first_unstepped = r;
      }
    }
    blocks[(size_t)t] = b;
  });
  bool sorted = true;
  T.unit_weights = true;
  T.dyadic_weights = true;
  T.binary_h = true;
  {
    int64_t first_unstepped = -1, last_stepped = -1, prev_day = -1;
    double prev_ward = 0.0;
    for (const Block& b : blocks) {
      if (b.nan_row >= 0) fail(AMRO_EINVAL, "ward_matrix: NaN ward identifier in row %lld", (long long)b.nan_row);
      T.Rp += b.n_stepped;
      if (!b.unit) T.unit_weights = false;
      if (!b.dyadic) T.dyadic_weights = false;
      if (b.half) T.has_half = true;
      if (!b.binary_h) T.binary_h = false;
      if (!b.ordered) sorted = false;
      if (b.last_stepped >= 0) {
        if (last_stepped >= 0 && (b.first_day < prev_day || (b.first_day == prev_day && b.first_ward < prev_ward))) sorted = false;
        last_stepped = b.last_stepped; prev_day = b.last_day; prev_ward = b.last_ward;
      }
      if (first_unstepped < 0 && b.first_unstepped >= 0) first_unstepped = b.first_unstepped;
    }
    if (first_unstepped >= 0 && last_stepped > first_unstepped) sorted = false;
  }
  T.identity_order = sorted;

  // ---- visiting order: (day, ward, original index), never-stepped rows last ------------------------
  T.order.resize(R);
  if (sorted) {
    parallel_blocks(R, n_thr, [&](int64_t lo, int64_t hi, int) { for (int64_t r = lo; r < hi; ++r) T.order[r] = r; });
  } else {
    int64_t a = 0, b = T.Rp;
    for (int64_t r = 0; r < R; ++r) {
      if (row_day[r] >= 0) T.order[a++] = r; else T.order[b++] = r;
    }
    std::stable_sort(T.order.begin(), T.order.begin() + T.Rp, [&](int64_t x, int64_t y) {
      if (row_day[x] != row_day[y]) return row_day[x] < row_day[y];
      return at(wm, ld, x, 1) < at(wm, ld, y, 1);
    });
  }
  T.pos_of.resize(R);
  parallel_blocks(R, n_thr, [&](int64_t lo, int64_t hi, int) { for (int64_t p = lo; p < hi; ++p) T.pos_of[T.order[p]] = p; });

  // ---- first tppw row of every (day, ward)   (abm_wards.cpp:384 then :242,249) ---------------------
  std::unordered_map<std::pair<uint64_t, uint64_t>, int64_t, PairHash> first_row;
  first_row.reserve((size_t)K * 2);
  if (t_cols >= 2) {
    for (int64_t k = 0; k < K; ++k) {
      const int64_t d = as_day(at(tp, t_ld, k, 0), T.total_days);
      const double wv = at(tp, t_ld, k, 1);
      if (d < 0 || wv != wv) continue;
      first_row.emplace(std::make_pair((uint64_t)d, dkey(wv)), k);
    }
  }

  // ---- days and (day, ward) segments ----------------------------------------------------------------
  T.day_start.assign(D + 1, 0);
  T.day_seg_start.assign(D + 1, 0);
  std::vector<uint32_t> seg_of_pos(T.Rp);
  {
    int64_t p = 0;
    for (int64_t d = 0; d < D; ++d) {
      T.day_start[d] = p;
      T.day_seg_start[d] = (int64_t)T.seg_N.size();
      while (p < T.Rp && row_day[T.order[p]] == d) {
        const double wv = at(wm, ld, T.order[p], 1);
        const int64_t s0 = p;
        while (p < T.Rp && row_day[T.order[p]] == d && at(wm, ld, T.order[p], 1) == wv) ++p;
        auto it = first_row.find(std::make_pair((uint64_t)d, dkey(wv)));
        if (it == first_row.end() || t_cols < 3)
          fail(AMRO_EINDEX, "Mat::operator(): index out of bounds (day %lld: ward %g has no row in total_patients_per_ward)",
               (long long)d, wv);
        if (T.seg_N.size() >= (size_t)0x7FFFFFF0) fail(AMRO_EINVAL, "too many (day, ward) segments");
        const uint32_t sid = (uint32_t)T.seg_N.size();
        for (int64_t q = s0; q < p; ++q) seg_of_pos[q] = sid;
        T.seg_row_start.push_back(s0);
        T.seg_N.push_back(at(tp, t_ld, it->second, 2));
        T.seg_ward.push_back(wv);
        T.max_seg_rows = std::max(T.max_seg_rows, p - s0);
      }
      T.max_rows_per_day = std::max(T.max_rows_per_day, p - T.day_start[d]);
      T.max_segs_per_day = std::max(T.max_segs_per_day, (int64_t)T.seg_N.size() - T.day_seg_start[d]);
    }
    T.day_start[D] = p;
    T.day_seg_start[D] = (int64_t)T.seg_N.size();
    T.seg_row_start.push_back(p);
    T.n_seg = (int64_t)T.seg_N.size();
  }

  // ---- next_day links: per day the surviving copies; per stepped row its effective predecessor -------
  // abm_wards.cpp:402-407: after day d, rows of d with next_day > 0 (ascending row order) overwrite the
  // C/D columns of row trunc(next_day); a later source wins.  A stepped row therefore starts its own day
  // from the last write made on an EARLIER day (or its initial state), and a write made on its own or a
  // later day only changes what the returned matrix shows for it.
  const std::unique_ptr<int64_t[]> pre_src_a = filled<int64_t>(T.Rp, -1, n_thr);   // original row of the effective predecessor
  int64_t* const pre_src = pre_src_a.get();
  T.day_push_start.assign(D + 1, 0);
  T.simple_final = (T.Rp == R);
  bool ring = true;
  // Threaded pass for the ordinary hospital -- no row is the target of two links: then every link is its day's winner, a
  // stepped target's predecessor is its one source if that source lives on an earlier day, and the push list of a day is its
  // links in ascending source order.  Anything else (several sources for one row) takes the sequential pass below, which
  // replays the reference's order of writes.
  bool links_done = false;
  if (n_thr > 1) {
    const std::unique_ptr<int64_t[]> target_a = filled<int64_t>(T.Rp, -1, n_thr);   // per position: the row its record writes to, or -1
    const std::unique_ptr<int64_t[]> src_of_a = filled<int64_t>(R, -1, n_thr);      // per row: its source (meaningful where n_src == 1)
    int64_t* const target = target_a.get();
    int64_t* const src_of = src_of_a.get();
    std::unique_ptr<std::atomic<uint32_t>[]> n_src(new std::atomic<uint32_t>[(size_t)std::max<int64_t>(R, 1)]);   // per row: links that point to it
    parallel_blocks(R, n_thr, [&](int64_t lo, int64_t hi, int) { for (int64_t r = lo; r < hi; ++r) n_src[(size_t)r].store(0u, std::memory_order_relaxed); });
    std::atomic<int64_t> bad_pos{INT64_MAX};
    std::atomic<bool> multi{false};
    parallel_blocks(T.Rp, n_thr, [&](int64_t lo, int64_t hi, int) {
      for (int64_t p = lo; p < hi; ++p) {
        const int64_t r = T.order[p];
        const double nd = at(wm, ld, r, 5);
        if (!(nd > 0.0)) continue;
        if (!(nd < 18446744073709551616.0) || (uint64_t)nd >= (uint64_t)R) {
          int64_t cur = bad_pos.load(std::memory_order_relaxed);
          while (p < cur && !bad_pos.compare_exchange_weak(cur, p, std::memory_order_relaxed)) {}
          continue;
        }
        const int64_t t = (int64_t)(uint64_t)nd;
        target[p] = t;
        if (n_src[(size_t)t].fetch_add(1u, std::memory_order_relaxed) != 0u) multi.store(true, std::memory_order_relaxed);
        src_of[(size_t)t] = r;
      }
    });
    // (an out-of-range link: the sequential pass reports the first one in the reference's order of writes)
    if (!multi.load() && bad_pos.load() == INT64_MAX) {
      std::vector<int64_t> n_push((size_t)D, 0);
      std::atomic<bool> not_final{false}, not_ring{false};
      parallel_blocks(D, std::min<int64_t>(n_thr, std::max<int64_t>(D, 1)), [&](int64_t dlo, int64_t dhi, int) {
        for (int64_t d = dlo; d < dhi; ++d) {
          int64_t n = 0;
          for (int64_t p = T.day_start[d]; p < T.day_start[d + 1]; ++p) {
            const int64_t src = src_of[(size_t)T.order[p]];
            if (src >= 0 && row_day[src] >= 0 && row_day[src] < d) {   // written on an earlier day: this is what the row starts from
              pre_src[p] = src;
              if (row_day[src] != d - 1) not_ring.store(true, std::memory_order_relaxed);
            }
            const int64_t t = target[p];
            if (t >= 0) {
              ++n;
              if (row_day[t] < 0 || row_day[t] <= d) not_final.store(true, std::memory_order_relaxed);
            }
          }
          n_push[(size_t)d] = n;
        }
      });
      for (int64_t d = 0; d < D; ++d) T.day_push_start[d + 1] = T.day_push_start[d] + n_push[(size_t)d];
      T.push_src.resize((size_t)T.day_push_start[D]);
      T.push_dst.resize((size_t)T.day_push_start[D]);
      parallel_blocks(D, std::min<int64_t>(n_thr, std::max<int64_t>(D, 1)), [&](int64_t dlo, int64_t dhi, int) {
        std::vector<std::pair<int64_t, int64_t>> ev;
        for (int64_t d = dlo; d < dhi; ++d) {
          int64_t q = T.day_push_start[d];
          if (sorted) {
            for (int64_t p = T.day_start[d]; p < T.day_start[d + 1]; ++p)
              if (target[p] >= 0) { T.push_src[(size_t)q] = T.order[p]; T.push_dst[(size_t)q] = target[p]; ++q; }
          } else {   // positions of a day are in (ward, row) order; the reference writes in row order
            ev.clear();
            for (int64_t p = T.day_start[d]; p < T.day_start[d + 1]; ++p) if (target[p] >= 0) ev.emplace_back(T.order[p], target[p]);
            std::sort(ev.begin(), ev.end());
            for (const auto& e : ev) { T.push_src[(size_t)q] = e.first; T.push_dst[(size_t)q] = e.second; ++q; }
          }
        }
      });
      if (not_final.load()) T.simple_final = false;
      if (not_ring.load()) ring = false;
      links_done = true;
    }
  }
  if (!links_done) {
  struct Written { int64_t writer, day; };   // last source row and the day it wrote on: one cache line per target
  std::vector<Written> written(R, Written{-1, -1});
  T.push_src.reserve((size_t)T.Rp);
  T.push_dst.reserve((size_t)T.Rp);
  std::vector<int64_t> ev;
  for (int64_t d = 0; d < D; ++d) {
    const int64_t a = T.day_start[d], b = T.day_start[d + 1];
    for (int64_t p = a; p < b; ++p) {
      const int64_t src = written[T.order[p]].writer;
      pre_src[p] = src;
      if (src >= 0 && row_day[src] != d - 1) ring = false;
    }
    T.day_push_start[d] = (int64_t)T.push_src.size();
    ev.clear();
    for (int64_t p = a; p < b; ++p) {
      const int64_t r = T.order[p];
      if (at(wm, ld, r, 5) > 0.0) ev.push_back(r);
    }
    if (!sorted) std::sort(ev.begin(), ev.end());
    const size_t base = T.push_src.size();
    for (size_t i = ev.size(); i-- > 0;) {  // reverse: the last source of a target is its winner
      const int64_t r = ev[i];
      const double nd = at(wm, ld, r, 5);
      if (!(nd < 18446744073709551616.0) || (uint64_t)nd >= (uint64_t)R)
        fail(AMRO_EINDEX, "Mat::elem(): index out of bounds (row %lld: next_day %g)", (long long)r, nd);
      const int64_t t = (int64_t)(uint64_t)nd;
      if (written[t].day == d) continue;
      written[t] = Written{r, d};
      T.push_src.push_back(r);
      T.push_dst.push_back(t);
      if (row_day[t] < 0 || row_day[t] <= d) T.simple_final = false;
    }
    std::reverse(T.push_src.begin() + base, T.push_src.end());
    std::reverse(T.push_dst.begin() + base, T.push_dst.end());
  }
  T.day_push_start[D] = (int64_t)T.push_src.size();
  }

  // ---- FAST path encodings ------------------------------------------------------------------------------
  T.why_not_fast.clear();
  if (!T.dyadic_weights) T.why_not_fast = "weights other than 1 or 1/2";
  else if (!T.binary_h) T.why_not_fast = "is_new values other than 0/1";
  else if (T.Rp != R) T.why_not_fast = "rows whose day is never stepped";
  else if (!T.simple_final) T.why_not_fast = "next_day links that point to the same or an earlier day";
  else if (!ring) T.why_not_fast = "next_day links that skip a day";
  else if (T.max_seg_rows > (T.has_half ? 32767 : 65535)) T.why_not_fast = "a ward-day with more than 65535 records (32767 with half weights)";
  else if (T.max_rows_per_day >= (int64_t)0xFFFFFFF0u) T.why_not_fast = "too many records in one day";
  else if (T.max_segs_per_day >= (int64_t)kInfoSegMask) T.why_not_fast = "too many wards in one day";
  if (T.why_not_fast.empty()) {
    T.meta_info.resize(T.Rp);   // written for every position below
    T.meta_next.resize(T.Rp);
    T.meta_nseg.resize(T.Rp);
    parallel_blocks(T.Rp, n_thr, [&](int64_t lo, int64_t hi, int) {
      std::fill(T.meta_next.begin() + lo, T.meta_next.begin() + hi, kNone);
      std::fill(T.meta_nseg.begin() + lo, T.meta_nseg.begin() + hi, kNone);
    });
    T.init_seg.assign(n_init, kNone);
    T.init_slot.assign(n_init, kNone);
    T.day_has_init.assign(D, 0);
    T.seg_main_next.assign(T.n_seg, kNone);
    {
      // same ward on the following day, where it exists
      std::unordered_map<std::pair<uint64_t, uint64_t>, uint32_t, PairHash> seg_of;
      seg_of.reserve((size_t)T.n_seg * 2);
      for (int64_t d = 0; d < D; ++d)
        for (int64_t sg = T.day_seg_start[d]; sg < T.day_seg_start[d + 1]; ++sg)
          seg_of.emplace(std::make_pair((uint64_t)d, dkey(T.seg_ward[sg])), (uint32_t)sg);
      for (int64_t d = 0; d + 1 < D; ++d)
        for (int64_t sg = T.day_seg_start[d]; sg < T.day_seg_start[d + 1]; ++sg) {
          auto it = seg_of.find(std::make_pair((uint64_t)(d + 1), dkey(T.seg_ward[sg])));
          if (it != seg_of.end()) T.seg_main_next[sg] = it->second;
        }
    }
    // two passes over the positions (blocks of days): what a record says about itself, then what its successor adds to it --
    // a record has one successor, so no two positions write the same predecessor
    std::atomic<bool> init_late{false};
    const int n_day_thr = (int)std::min<int64_t>(n_thr, std::max<int64_t>(D, 1));
    parallel_blocks(D, n_day_thr, [&](int64_t dlo, int64_t dhi, int) {
      for (int64_t d = dlo; d < dhi; ++d)
        for (int64_t p = T.day_start[d]; p < T.day_start[d + 1]; ++p) {
          const int64_t r = T.order[p];
          const uint32_t seg_local = seg_of_pos[p] - (uint32_t)T.day_seg_start[d];
          uint32_t kind = kSrcNone;
          if (pre_src[p] >= 0) kind = kSrcPrev;
          else if (r < n_init) {
            kind = kSrcInit;
            if (d != 0) init_late.store(true, std::memory_order_relaxed);
            T.init_seg[r] = seg_of_pos[p];
            T.init_slot[r] = (uint32_t)(p - T.day_start[d]);
            T.day_has_init[d] = 1;
          }
          T.meta_info[p] = (T.h[r] == 1.0 ? (1u << kInfoHShift) : 0u) | (kind << kInfoSrcShift) |
                           (T.w[r] == 0.5 ? (1u << kInfoHalfShift) : 0u) | seg_local;
        }
    });
    parallel_blocks(D, n_day_thr, [&](int64_t dlo, int64_t dhi, int) {
      for (int64_t d = dlo; d < dhi; ++d)
        for (int64_t p = T.day_start[d]; p < T.day_start[d + 1]; ++p) {
          if (pre_src[p] < 0) continue;
          const int64_t pp = T.pos_of[pre_src[p]];  // in day d-1 (ring property)
          T.meta_next[pp] = (uint32_t)(p - T.day_start[d]);
          T.meta_nseg[pp] = seg_of_pos[p];
          if (T.w[T.order[p]] == 0.5) T.meta_info[pp] |= 1u << kInfoNextHalfShift;   // what pp's state adds to this ward-day's sum(W) is w of THIS row
        }
    });
    if (init_late.load()) T.why_not_fast = "initial-state rows outside day 0";
  }
  T.fast_eligible = T.why_not_fast.empty();
  if (!T.fast_eligible) {
    T.meta_info.clear(); T.meta_next.clear(); T.meta_nseg.clear(); T.init_seg.clear(); T.init_slot.clear(); T.seg_main_next.clear();
  }
}

void compile_day(const double* wm, int64_t n, int64_t n_cols, int64_t ld,
                 const double* tp, int64_t K, int64_t t_cols, int64_t t_ld, double total_patients,
                 std::vector<int64_t>& order, std::vector<int64_t>& seg_row_start, std::vector<double>& seg_N) {
  order.resize(n);
  std::iota(order.begin(), order.end(), (int64_t)0);
  seg_row_start.clear();
  seg_N.clear();
  if (tp == nullptr) {  // ward step: all rows are one ward with the caller's N
    if (n > 0) { seg_row_start.push_back(0); seg_N.push_back(total_patients); }
    seg_row_start.push_back(n);
    return;
  }
  if (n == 0) { seg_row_start.push_back(0); return; }  // unique() of an empty column: nothing to do
  if (n_cols < 2) fail(AMRO_EINDEX, "Mat::col(): index out of bounds");
  if (t_cols < 2) fail(AMRO_EINDEX, "Mat::col(): index out of bounds");
  for (int64_t r = 0; r < n; ++r)
    if (at(wm, ld, r, 1) != at(wm, ld, r, 1)) fail(AMRO_EINVAL, "ward_matrix: NaN ward identifier in row %lld", (long long)r);
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return at(wm, ld, x, 1) < at(wm, ld, y, 1); });
  int64_t p = 0;
  while (p < n) {
    const double wv = at(wm, ld, order[p], 1);
    seg_row_start.push_back(p);
    while (p < n && at(wm, ld, order[p], 1) == wv) ++p;
    int64_t hit = -1;
    for (int64_t k = 0; k < K; ++k) if (at(tp, t_ld, k, 1) == wv) { hit = k; break; }  // :242, first match
    if (hit < 0 || t_cols < 3) fail(AMRO_EINDEX, "Mat::operator(): index out of bounds (ward %g has no row in total_patients_per_ward)", wv);
    seg_N.push_back(at(tp, t_ld, hit, 2));
  }
  seg_row_start.push_back(n);
}

}  // namespace amro
