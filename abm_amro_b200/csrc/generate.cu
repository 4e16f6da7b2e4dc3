// Synthetic hospitals generated ON THE DEVICE, straight into the compiled form the fused kernels read:
// configs 4 and 5 of BASELINE.json (3.65e8 and 7.3e9 patient-day records) are far beyond what float64 tables through the
// numpy boundary can carry (17.5 GB and 350 GB), so their topology never exists on the host.
//
// The hospital ("bed model", the law of the survey's generator): `census` beds, each always occupied.  Day 0:
// every bed holds a new patient (is_new = 1) in a uniformly drawn ward.  After each day the patient of a bed is discharged
// with probability p_discharge and replaced by a new patient in a uniformly drawn ward (is_new = 1, no predecessor);
// otherwise the patient moves to a uniformly drawn ward with probability p_transfer, else stays (is_new = 0, its record is
// the successor of today's).  The records of a day are ordered by (ward, bed); row = day * census + position, so rows are
// already in (day, ward) order and the original row of a record is its position.  total_patients_per_ward = the bed count of
// every occupied ward.  All draws are Philox4x32-10 calls with counter (bed_lo, bed_hi, day, stream) and key = seed.
//
// amro_hospital_tables builds the SAME hospital as the reference's float64 tables on the host (small sizes: the parity
// tests compile those tables with topology.cpp and compare arrays and results with the device-generated handle).
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "handle.cuh"

namespace amro {
namespace {

constexpr uint32_t kGenStreamDay0 = 0x47454E00u, kGenStreamMove = 0x47454E01u;

struct Draw4 { uint32_t x, y, z, w; };
__host__ __device__ inline Draw4 gen_philox(uint64_t bed, uint32_t day, uint32_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)bed, c1 = (uint32_t)(bed >> 32), c2 = day, c3 = stream, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Draw4{c0, c1, c2, c3};
}
__host__ __device__ inline uint32_t gen_ward(uint32_t u, uint32_t n_wards) { return 1u + (uint32_t)(((uint64_t)u * n_wards) >> 32); }
__host__ __device__ inline uint64_t gen_threshold(double p) {   // u < T  <=>  u * 2^-32 < p
  if (!(p > 0.0)) return 0ull;
  if (p >= 1.0) return 1ull << 32;
  return (uint64_t)(p * 4294967296.0);
}
// what happens to the patient of `bed` after day `day`: bit 0 = discharged (a new patient tomorrow), bits 31:1 = tomorrow's ward
__host__ __device__ inline uint32_t gen_move(uint64_t bed, uint32_t day, uint32_t ward, uint32_t n_wards, uint64_t t_dis, uint64_t t_tra, uint64_t seed) {
  const Draw4 u = gen_philox(bed, day, kGenStreamMove, seed);
  if ((uint64_t)u.x < t_dis) return (gen_ward(u.y, n_wards) << 1) | 1u;
  if ((uint64_t)u.z < t_tra) return gen_ward(u.w, n_wards) << 1;
  return ward << 1;
}

void check_spec(const amro_hospital_spec* s) {
  if (!s) fail(AMRO_EINVAL, "spec is NULL");
  if (s->n_wards < 1 || s->n_wards > 60000) fail(AMRO_EINVAL, "n_wards must be in [1, 60000]");
  if (s->census < 1 || s->census >= ((int64_t)1 << 31)) fail(AMRO_EINVAL, "census must be in [1, 2^31)");
  if (s->n_days < 1 || s->n_days > kMaxCountdown) fail(AMRO_EINVAL, "n_days must be in [1, %d]", kMaxCountdown);
}

// ---- device kernels --------------------------------------------------------------------------------------------------
__global__ void k_gen_day0(uint32_t* ward, uint32_t* flags, int64_t B, uint32_t n_wards, uint64_t seed) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  ward[b] = gen_ward(gen_philox((uint64_t)b, 0u, kGenStreamDay0, seed).x, n_wards);
  flags[b] = 1u;   // bit 0: is_new
}
// flags_cur bit 1 <- discharged after today; ward_nxt / flags_nxt bit 0 (is_new) of tomorrow
__global__ void k_gen_move(const uint32_t* ward_cur, uint32_t* flags_cur, uint32_t* ward_nxt, uint32_t* flags_nxt, int64_t B, uint32_t day,
                           uint32_t n_wards, uint64_t t_dis, uint64_t t_tra, uint64_t seed) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t m = gen_move((uint64_t)b, day, ward_cur[b], n_wards, t_dis, t_tra, seed);
  flags_cur[b] = (flags_cur[b] & 1u) | ((m & 1u) << 1);
  ward_nxt[b] = m >> 1;
  flags_nxt[b] = m & 1u;
}
__global__ void k_gen_iota(uint32_t* v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}
__global__ void k_gen_hist(const uint32_t* ward, int64_t B, uint32_t* count) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) atomicAdd(&count[ward[b]], 1u);
}
__global__ void k_gen_pos(const uint32_t* sorted_beds, int64_t B, uint32_t* pos) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < B) pos[sorted_beds[p]] = (uint32_t)p;
}
// One thread: the ward-days of day d (occupied wards in ascending order) and, for yesterday's ward-days, the same ward today.
// scal[0] = segments so far (in: this day's first; out: the next day's first), scal[1] = largest ward-day, scal[2] = most ward-days of a day
__global__ void k_gen_segments(const uint32_t* count, uint32_t n_wards, int64_t day, int64_t B, uint32_t* segl /*[n_wards+1]*/,
                               int64_t* seg_row_start, double* seg_N, uint32_t* seg_ward, uint32_t* seg_main_next, int64_t prev_first,
                               int64_t* scal, int64_t* day_seg_start) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t first = scal[0];
  int64_t j = first, at = day * B;
  uint32_t biggest = (uint32_t)scal[1];
  for (uint32_t w = 1; w <= n_wards; ++w) {
    const uint32_t c = count[w];
    if (c == 0) { segl[w] = kNone; continue; }
    segl[w] = (uint32_t)(j - first);
    seg_row_start[j] = at;
    seg_N[j] = (double)c;
    seg_ward[j] = w;
    seg_main_next[j] = kNone;
    at += c;
    biggest = max(biggest, c);
    ++j;
  }
  seg_row_start[j] = at;
  for (int64_t q = prev_first; q < first; ++q) {   // yesterday's ward-days feed the same ward today, where it is occupied
    const uint32_t l = segl[seg_ward[q]];
    seg_main_next[q] = l == kNone ? kNone : (uint32_t)(first + l);
  }
  day_seg_start[day] = first;
  day_seg_start[day + 1] = j;
  scal[0] = j;
  scal[1] = biggest;
  scal[2] = max(scal[2], j - first);
}
// the packed per-record words of day e (kernels.hpp FastArgs::meta), one thread per bed
__global__ void k_gen_meta(uint4* meta, uint32_t* init_seg, uint32_t* init_slot, int64_t B, int64_t e, int has_next_day,
                           const uint32_t* ward_e, const uint32_t* flags_e, const uint32_t* pos_e, const uint32_t* segl_e,
                           const uint32_t* ward_n, const uint32_t* pos_n, const uint32_t* segl_n, int64_t seg_first_e, int64_t seg_first_n) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t p = pos_e[b], f = flags_e[b], sl = segl_e[ward_e[b]];
  const uint32_t kind = e == 0 ? kSrcInit : ((f & 1u) ? kSrcNone : kSrcPrev);
  uint32_t next = kNone, nseg = kNone;
  if (has_next_day && !(f & 2u)) {
    next = (uint32_t)(fast_ring_slot(pos_n[b]) * (int64_t)sizeof(uint4));
    nseg = (uint32_t)(seg_first_n + segl_n[ward_n[b]]);
  }
  const int64_t row = e * B + p;
  meta[row] = make_uint4(((f & 1u) << kInfoHShift) | (kind << kInfoSrcShift) | sl, next, nseg, (uint32_t)row);
  if (e == 0) { init_slot[p] = p; init_seg[p] = (uint32_t)(seg_first_e + sl); }
}

}  // namespace
}  // namespace amro

using namespace amro;

extern "C" int amro_hospital_tables(const amro_hospital_spec* spec, int64_t* n_rows, int64_t* t_rows, double* wm, double* tp) {
  return guarded([&] {
    check_spec(spec);
    const int64_t B = spec->census, D = spec->n_days, R = B * D;
    const uint32_t W = (uint32_t)spec->n_wards;
    const uint64_t seed = (uint64_t)spec->seed, t_dis = gen_threshold(spec->p_discharge), t_tra = gen_threshold(spec->p_transfer);
    if (R > ((int64_t)1 << 28)) fail(AMRO_EINVAL, "amro_hospital_tables is for small hospitals (tests); this one has %lld rows", (long long)R);
    std::vector<uint32_t> ward(B), flags(B), ward_n(B), flags_n(B), beds(B), pos(B), pos_n(B);
    std::vector<int64_t> mrn(B);
    for (int64_t b = 0; b < B; ++b) { ward[b] = gen_ward(gen_philox((uint64_t)b, 0u, kGenStreamDay0, seed).x, W); flags[b] = 1u; mrn[b] = b + 1; }
    int64_t next_mrn = B + 1, K = 0;
    auto order_day = [&](const std::vector<uint32_t>& wd, std::vector<uint32_t>& ps) {   // records of a day in (ward, bed) order
      for (int64_t b = 0; b < B; ++b) beds[b] = (uint32_t)b;
      std::stable_sort(beds.begin(), beds.end(), [&](uint32_t x, uint32_t y) { return wd[x] < wd[y]; });
      for (int64_t p = 0; p < B; ++p) ps[beds[p]] = (uint32_t)p;
    };
    order_day(ward, pos);
    for (int64_t d = 0; d < D; ++d) {
      std::vector<int64_t> count(W + 1, 0);
      for (int64_t b = 0; b < B; ++b) count[ward[b]]++;
      for (uint32_t w = 1; w <= W; ++w)
        if (count[w]) {
          if (tp) { tp[K] = (double)d; tp[K + *t_rows] = (double)w; tp[K + 2 * *t_rows] = (double)count[w]; }
          ++K;
        }
      for (int64_t b = 0; b < B; ++b) {
        const uint32_t m = gen_move((uint64_t)b, (uint32_t)d, ward[b], W, t_dis, t_tra, seed);
        flags[b] = (flags[b] & 1u) | ((m & 1u) << 1);
        ward_n[b] = m >> 1;
        flags_n[b] = m & 1u;
      }
      if (d + 1 < D) order_day(ward_n, pos_n);
      if (wm)
        for (int64_t b = 0; b < B; ++b) {
          const int64_t r = d * B + pos[b];
          wm[r] = (double)d; wm[r + R] = (double)ward[b]; wm[r + 2 * R] = (double)mrn[b];
          wm[r + 3 * R] = (double)(flags[b] & 1u); wm[r + 4 * R] = 1.0;
          wm[r + 5 * R] = (d + 1 < D && !(flags[b] & 2u)) ? (double)((d + 1) * B + pos_n[b]) : -1.0;
        }
      for (int64_t b = 0; b < B; ++b) if (flags[b] & 2u) mrn[b] = next_mrn++;
      ward.swap(ward_n); flags.swap(flags_n); pos.swap(pos_n);
    }
    if (n_rows) *n_rows = R;
    if (t_rows && !tp) *t_rows = K;
  });
}

extern "C" int amro_topology_generate(const amro_hospital_spec* spec, int device, amro_topology** out) {
  return guarded([&] {
    if (!out) fail(AMRO_EINVAL, "out is NULL");
    *out = nullptr;
    check_spec(spec);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); fail(AMRO_ECUDA, "no usable CUDA device; amro_b200 has no CPU fallback"); }
    if (device < 0 || device >= n_dev) fail(AMRO_EINVAL, "device %d out of range (have %d)", device, n_dev);
    AMRO_CUDA(cudaSetDevice(device));
    const auto t0 = std::chrono::steady_clock::now();
    const int64_t B = spec->census, D = spec->n_days, R = B * D;
    const uint32_t W = (uint32_t)spec->n_wards;
    const uint64_t seed = (uint64_t)spec->seed, t_dis = gen_threshold(spec->p_discharge), t_tra = gen_threshold(spec->p_transfer);
    if (((B + 31) / 32) * 32 * kSlicesPerTeam * (int64_t)sizeof(uint4) >= ((int64_t)1 << 32)) fail(AMRO_EINVAL, "census too large for 32-bit slot offsets");

    std::unique_ptr<amro_topology> t(new amro_topology);
    t->device = device;
    t->generated = true;
    const int64_t max_seg = std::min<int64_t>((int64_t)W, B) * D;
    t->meta.alloc((size_t)R);
    t->day_start.alloc((size_t)(D + 1));
    t->day_seg_start.alloc((size_t)(D + 1));
    t->seg_row_start.alloc((size_t)(max_seg + 1));
    t->seg_N.alloc((size_t)max_seg);
    t->seg_main_next.alloc((size_t)max_seg);
    t->init_seg.alloc((size_t)B);
    t->init_slot.alloc((size_t)B);
    DevBuf<uint32_t> seg_ward, ward[2], flags[2], pos[2], segl[2], count, iota, sorted_keys, sorted_beds;
    DevBuf<int64_t> scal;
    DevBuf<unsigned char> tmp;
    seg_ward.alloc((size_t)max_seg);
    for (int i = 0; i < 2; ++i) { ward[i].alloc((size_t)B); flags[i].alloc((size_t)B); pos[i].alloc((size_t)B); segl[i].alloc((size_t)W + 1); }
    count.alloc((size_t)W + 1); iota.alloc((size_t)B); sorted_keys.alloc((size_t)B); sorted_beds.alloc((size_t)B);
    scal.alloc(3);
    AMRO_CUDA(cudaMemsetAsync(scal.p, 0, 3 * sizeof(int64_t), 0));
    {
      std::vector<int64_t> ds((size_t)(D + 1));
      for (int64_t d = 0; d <= D; ++d) ds[d] = d * B;
      AMRO_CUDA(cudaMemcpyAsync(t->day_start.p, ds.data(), sizeof(int64_t) * (D + 1), cudaMemcpyHostToDevice, 0));
      AMRO_CUDA(cudaStreamSynchronize(0));
    }
    int key_bits = 1;
    while ((1u << key_bits) <= W) ++key_bits;
    size_t tmp_bytes = 0;
    AMRO_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ward[0].p, sorted_keys.p, iota.p, sorted_beds.p, (int)B, 0, key_bits, 0));
    tmp.alloc(tmp_bytes);
    const unsigned grid = (unsigned)((B + 255) / 256);
    k_gen_iota<<<grid, 256>>>(iota.p, B);
    k_gen_day0<<<grid, 256>>>(ward[0].p, flags[0].p, B, W, seed);

    std::vector<int64_t> firsts((size_t)(D + 1), 0);   // first ward-day of every day (host copy of day_seg_start)
    for (int64_t d = 0; d < D; ++d) {
      const int c = (int)(d & 1), n = c ^ 1;
      // the day's ward-days, positions of its records
      AMRO_CUDA(cudaMemsetAsync(count.p, 0, sizeof(uint32_t) * (W + 1), 0));
      k_gen_hist<<<grid, 256>>>(ward[c].p, B, count.p);
      k_gen_segments<<<1, 32>>>(count.p, W, d, B, segl[c].p, t->seg_row_start.p, t->seg_N.p, seg_ward.p, t->seg_main_next.p,
                               d > 0 ? firsts[d - 1] : 0, scal.p, t->day_seg_start.p);
      AMRO_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, ward[c].p, sorted_keys.p, iota.p, sorted_beds.p, (int)B, 0, key_bits, 0));
      k_gen_pos<<<grid, 256>>>(sorted_beds.p, B, pos[c].p);
      // the number of ward-days so far: where this day's begin (one small copy per day)
      int64_t s3[3];
      AMRO_CUDA(cudaMemcpy(s3, scal.p, sizeof s3, cudaMemcpyDeviceToHost));
      firsts[d + 1] = s3[0];
      // yesterday's records: their successors' positions and ward-days are known now
      if (d > 0)
        k_gen_meta<<<grid, 256>>>(t->meta.p, t->init_seg.p, t->init_slot.p, B, d - 1, 1, ward[n].p, flags[n].p, pos[n].p, segl[n].p, ward[c].p,
                                  pos[c].p, segl[c].p, firsts[d - 1], firsts[d]);
      // who leaves after today, and tomorrow's wards
      if (d + 1 < D) k_gen_move<<<grid, 256>>>(ward[c].p, flags[c].p, ward[n].p, flags[n].p, B, (uint32_t)d, W, t_dis, t_tra, seed);
      if (d + 1 == D)
        k_gen_meta<<<grid, 256>>>(t->meta.p, t->init_seg.p, t->init_slot.p, B, d, 0, ward[c].p, flags[c].p, pos[c].p, segl[c].p, ward[c].p, pos[c].p,
                                  segl[c].p, firsts[d], firsts[d]);
    }
    AMRO_CUDA(cudaGetLastError());
    int64_t s3[3];
    AMRO_CUDA(cudaMemcpy(s3, scal.p, sizeof s3, cudaMemcpyDeviceToHost));

    HostTopology& H = t->H;
    H.R = H.Rp = R; H.n_init = B; H.total_days = D - 1; H.n_days_out = D;
    H.n_seg = s3[0]; H.max_seg_rows = s3[1]; H.max_segs_per_day = s3[2]; H.max_rows_per_day = B;
    H.identity_order = H.unit_weights = H.binary_h = H.simple_final = H.dyadic_weights = true;
    H.has_half = false;
    H.fast_eligible = true;
    t->compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = t.release();
  });
}

// device arrays of a handle, for the parity tests of the generator against topology.cpp (not part of the public header)
extern "C" int amro_topology_debug_device_arrays(amro_topology* topo, void* meta /*[Rp] uint4*/, int64_t* seg_row_start /*[n_seg+1]*/,
                                                 double* seg_N, uint32_t* seg_main_next, int64_t* day_seg_start /*[D+1]*/,
                                                 uint32_t* init_seg, uint32_t* init_slot) {
  return guarded([&] {
    if (!topo || topo->device < 0) fail(AMRO_EINVAL, "no device handle");
    AMRO_CUDA(cudaSetDevice(topo->device));
    const HostTopology& H = topo->H;
    const int64_t D = H.total_days + 1;
    if (meta) AMRO_CUDA(cudaMemcpy(meta, topo->meta.p, sizeof(uint4) * H.Rp, cudaMemcpyDeviceToHost));
    if (seg_row_start) AMRO_CUDA(cudaMemcpy(seg_row_start, topo->seg_row_start.p, sizeof(int64_t) * (H.n_seg + 1), cudaMemcpyDeviceToHost));
    if (seg_N) AMRO_CUDA(cudaMemcpy(seg_N, topo->seg_N.p, sizeof(double) * H.n_seg, cudaMemcpyDeviceToHost));
    if (seg_main_next) AMRO_CUDA(cudaMemcpy(seg_main_next, topo->seg_main_next.p, sizeof(uint32_t) * H.n_seg, cudaMemcpyDeviceToHost));
    if (day_seg_start) AMRO_CUDA(cudaMemcpy(day_seg_start, topo->day_seg_start.p, sizeof(int64_t) * (D + 1), cudaMemcpyDeviceToHost));
    if (init_seg) AMRO_CUDA(cudaMemcpy(init_seg, topo->init_seg.p, sizeof(uint32_t) * H.n_init, cudaMemcpyDeviceToHost));
    if (init_slot) AMRO_CUDA(cudaMemcpy(init_slot, topo->init_slot.p, sizeof(uint32_t) * H.n_init, cudaMemcpyDeviceToHost));
  });
}
