// Host-side topology compiler: turns the reference's ward_matrix / total_patients_per_ward tables into
// the grouped, link-inverted form the kernels read.  Pure C++ (no CUDA) so it is testable on a CPU box.
//
// It replaces the work the reference redoes every day of every run (abm_wards.cpp:235-253 unique/find per
// ward, :384-388 find per day, :402-407 next_day scatter) by doing it once per hospital.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace amro {

constexpr uint32_t kNone = 0xFFFFFFFFu;          // "no segment" / "no successor"
// meta_info word of a record: bit 31 = is_new, bits 30:29 = where its pre-step state comes from, bit 28 = weight 1/2 (a
// patient split over two wards), bit 27 = its successor weighs 1/2, bits 26:0 = index of its ward-day within the day
constexpr uint32_t kSrcNone = 0u, kSrcPrev = 1u, kSrcInit = 2u;
constexpr uint32_t kInfoHShift = 31, kInfoSrcShift = 29, kInfoHalfShift = 28, kInfoNextHalfShift = 27,
                   kInfoSegMask = (1u << kInfoNextHalfShift) - 1u;

// std::vector whose resize() leaves trivially-constructible elements uninitialised: the per-row arrays of a hospital of millions of
// rows are then first written -- and their pages first touched -- by the compiler's threads instead of one zero-filling loop
template <class T>
struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAlloc<U>; };
  NoInitAlloc() = default;
  template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new ((void*)p) U;
    else ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
template <class T> using RowVec = std::vector<T, NoInitAlloc<T>>;

struct HostTopology {
  int64_t R = 0, n_init = 0;
  int64_t total_days = 0;    // max(tppw[:,0])                      (abm_wards.cpp:356)
  int64_t n_days_out = 0;    // max(ward_matrix[:,0]) + 1           (abm_wards.cpp:440)
  int64_t Rp = 0;            // stepped rows: day integer in [0, total_days]
  int64_t n_seg = 0, max_rows_per_day = 0, max_segs_per_day = 0, max_seg_rows = 0;
  bool identity_order = false, unit_weights = false, binary_h = false, simple_final = false;
  bool dyadic_weights = false;   // every stepped row weighs 1 or 1/2 (the FAST path counts ward pressure in halves then)
  bool has_half = false;         // ... and some weigh 1/2
  bool fast_eligible = false;
  std::string why_not_fast;

  // position p in [0,R): stepped rows first, sorted by (day, ward, original index) -- the order the
  // reference visits them (day loop :381, unique() ascending wards :235, find() ascending rows :241);
  // rows that are never stepped follow in original order.
  RowVec<int64_t> order;               // [R]  position -> original row
  RowVec<int64_t> pos_of;              // [R]  original row -> position
  std::vector<int64_t> day_start;      // [total_days+2] first position of each day
  std::vector<int64_t> day_seg_start;  // [total_days+2] first segment of each day
  std::vector<int64_t> seg_row_start;  // [n_seg+1] first position of each (day, ward) segment
  std::vector<double>  seg_N;          // [n_seg]  N of the first matching tppw row (:242,249)
  std::vector<double>  seg_ward;       // [n_seg]
  RowVec<double>  h, w;                // [R] per ORIGINAL row: is_new, weight
  RowVec<int32_t> count_day;           // [R] per ORIGINAL row: day index for the totals, -1 if none

  // GENERAL path: per day, the surviving (last-writer-wins) next_day copies   (:402-407)
  std::vector<int64_t> day_push_start; // [total_days+2]
  RowVec<int64_t> push_src, push_dst;  // original row indices

  // FAST path per position (valid when fast_eligible)
  RowVec<uint32_t> meta_info;          // [Rp] h<<31 | source kind<<29 | segment index within the day
  RowVec<uint32_t> meta_next;          // [Rp] position of the successor record within the NEXT day, or kNone
  RowVec<uint32_t> meta_nseg;          // [Rp] absolute segment of that successor, or kNone
  std::vector<uint32_t> init_seg;      // [n_init] absolute segment of init-sourced stepped rows, or kNone
  std::vector<uint32_t> init_slot;     // [n_init] position of that row within its day, or kNone
  std::vector<uint8_t>  day_has_init;  // [total_days+1] 1 if the day holds init-sourced rows
  std::vector<uint32_t> seg_main_next; // [n_seg] the ward-day most of this ward-day's records feed (same ward, next day), or kNone
};

// Throws amro::Error (AMRO_EINDEX / AMRO_ERUNTIME / AMRO_EINVAL) mirroring what the reference raises.
void compile_topology(const double* wm, int64_t R, int64_t n_cols, int64_t ld,
                      const double* tp, int64_t K, int64_t t_cols, int64_t t_ld,
                      int64_t n_init, HostTopology& T);

// Grouping for the single-day entry points (progress_patients_1_timestep: wards only, no day column
// match; abm_wards.cpp:235-249).  tp == nullptr: one segment with N = total_patients (ward step, :70).
void compile_day(const double* wm, int64_t n, int64_t n_cols, int64_t ld,
                 const double* tp, int64_t K, int64_t t_cols, int64_t t_ld, double total_patients,
                 std::vector<int64_t>& order, std::vector<int64_t>& seg_row_start, std::vector<double>& seg_N);

}  // namespace amro
