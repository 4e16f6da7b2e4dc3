// C ABI of libamro_b200.so (include/amro_b200.h): host orchestration around the CUDA kernels.
// There is no CPU fallback: every compute entry point needs a CUDA device and fails with AMRO_ECUDA otherwise.
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.hpp"
#include "kernels.hpp"
#include "topology.hpp"
#include "handle.cuh"

namespace amro {

static thread_local char g_last_error[512] = "";
void set_last_error(const char* msg) { snprintf(g_last_error, sizeof g_last_error, "%s", msg); }

namespace {

struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

void use_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    fail(AMRO_ECUDA, "no usable CUDA device (%s); amro_b200 has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) fail(AMRO_EINVAL, "device %d out of range (have %d)", device, n);
  AMRO_CUDA(cudaSetDevice(device));
}

// strided host matrix -> dense device matrix [rows x cols], ld = rows
void upload_matrix(DevBuf<double>& dst, const double* src, int64_t rows, int64_t cols, int64_t ld, cudaStream_t st) {
  dst.alloc((size_t)std::max<int64_t>(rows * cols, 1));
  if (rows == 0 || cols == 0) return;
  AMRO_CUDA(cudaMemcpy2DAsync(dst.p, rows * sizeof(double), src, ld * sizeof(double), rows * sizeof(double), cols,
                              cudaMemcpyHostToDevice, st));
}

}  // namespace
}  // namespace amro

using namespace amro;

template <class T>
static void ensure_size(DevBuf<T>& b, size_t n) {
  if (b.n < n || b.p == nullptr) b.alloc(n);
}

extern "C" {

const char* amro_last_error(void) { return g_last_error; }
const char* amro_version(void) { return AMRO_B200_VERSION; }

int amro_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int amro_topology_create(const double* ward_matrix, int64_t n_rows, int64_t n_cols, int64_t ld,
                         const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                         int64_t n_init, int device, amro_topology** out) {
  return guarded([&] {
    if (!out) fail(AMRO_EINVAL, "out is NULL");
    *out = nullptr;
    Timer tm;
    std::unique_ptr<amro_topology> t(new amro_topology);
    compile_topology(ward_matrix, n_rows, n_cols, ld, tppw, t_rows, t_cols, t_ld, n_init, t->H);
    t->device = device;
    if (device >= 0) {  // device < 0: host-only handle (inspection / CPU tests); amro_simulate refuses it
      use_device(device);
      const HostTopology& H = t->H;
      t->day_start.upload(H.day_start.data(), H.day_start.size());
      t->day_seg_start.upload(H.day_seg_start.data(), H.day_seg_start.size());
      t->seg_row_start.upload(H.seg_row_start.data(), H.seg_row_start.size());
      t->seg_N.upload(H.seg_N.data(), H.seg_N.size());
      if (!H.identity_order) t->pos_of.upload(H.pos_of.data(), H.pos_of.size());
      if (H.fast_eligible) {
        // one 16-byte word per record: what a lane needs about its record comes with ONE load, and the successor's slot
        // is already the byte offset inside a state-ring buffer laid out [chunk of 32][slice of the team][32]
        std::unique_ptr<uint4[]> meta(new uint4[(size_t)std::max<int64_t>(H.Rp, 1)]);   // packed by a few threads (58 MB for config 2)
        {
          const int n_thr = H.Rp < 200000 ? 1 : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
          std::vector<std::thread> pool;
          auto pack = [&](int64_t lo, int64_t hi) {
            for (int64_t p = lo; p < hi; ++p) {
              const uint32_t nx = H.meta_next[p];
              meta[(size_t)p] = make_uint4(H.meta_info[p], nx == kNone ? kNone : (uint32_t)(fast_ring_slot(nx) * sizeof(uint4)),
                                           H.meta_nseg[p], (uint32_t)H.order[p]);
            }
          };
          for (int k = 1; k < n_thr; ++k) pool.emplace_back(pack, H.Rp * k / n_thr, H.Rp * (k + 1) / n_thr);
          pack(0, H.Rp / n_thr);
          for (auto& th : pool) th.join();
        }
        t->meta.upload(meta.get(), (size_t)H.Rp);
        t->init_seg.upload(H.init_seg.data(), H.init_seg.size());
        t->init_slot.upload(H.init_slot.data(), H.init_slot.size());
        t->seg_main_next.upload(H.seg_main_next.data(), H.seg_main_next.size());
      }
      AMRO_CUDA(cudaStreamSynchronize(0));
    }
    t->compile_ms = tm.ms();
    *out = t.release();
  });
}

void amro_topology_destroy(amro_topology* topo) {
  if (!topo) return;
  if (topo->device >= 0) cudaSetDevice(topo->device);
  delete topo;
}

int amro_topology_get_info(const amro_topology* topo, amro_topology_info* info) {
  return guarded([&] {
    if (!topo || !info) fail(AMRO_EINVAL, "NULL argument");
    const HostTopology& H = topo->H;
    std::memset(info, 0, sizeof *info);
    info->n_rows = H.R; info->n_rows_stepped = H.Rp; info->n_init = H.n_init; info->total_days = H.total_days;
    info->n_days_out = H.n_days_out; info->n_segments = H.n_seg; info->max_rows_per_day = H.max_rows_per_day;
    info->fast_eligible = H.fast_eligible; info->ring_eligible = H.fast_eligible; info->identity_order = H.identity_order;
    info->device = topo->device; info->compile_ms = topo->compile_ms;
    snprintf(info->why_not_fast, sizeof info->why_not_fast, "%s", H.why_not_fast.c_str());
  });
}

}  // extern "C"

extern "C" int amro_day_moments_peers(const int32_t* counts_colonized, const int32_t* counts_detected, int64_t n_days, int64_t n_sims,
                                      int64_t ld_counts, double* const* dst, int n_dst, int device, void* stream) {
  return guarded([&] {
    if (!counts_colonized || !counts_detected || !dst) fail(AMRO_EINVAL, "NULL argument");
    if (n_days < 0 || n_sims < 1 || ld_counts < n_days) fail(AMRO_EINVAL, "bad shape");
    if (n_dst < 1 || n_dst > kMaxPeers) fail(AMRO_EINVAL, "between 1 and %d destinations", kMaxPeers);
    for (int i = 0; i < n_dst; ++i) if (!dst[i]) fail(AMRO_EINVAL, "destination %d is NULL", i);
    use_device(device);
    day_moments_peers((cudaStream_t)stream, counts_colonized, counts_detected, n_days, n_sims, ld_counts, dst, n_dst, n_days);
    AMRO_CUDA(cudaGetLastError());
  });
}

extern "C" int amro_topology_kernel_times(amro_topology* topo, int n, float* ms) {
  return guarded([&] {
    if (!topo || !ms || n < 0) fail(AMRO_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lock(topo->mu);
    if ((uint64_t)n > topo->n_fused_runs || n > amro_topology::kKernelEvents) fail(AMRO_EINVAL, "only %llu fused runs (at most %d) are on record",
                                                                                 (unsigned long long)topo->n_fused_runs, amro_topology::kKernelEvents);
    if (topo->device >= 0) AMRO_CUDA(cudaSetDevice(topo->device));
    for (int i = 0; i < n; ++i) {   // ms[0] = the oldest of the last n
      EventPair& e = topo->kernel_events[(topo->n_fused_runs - (uint64_t)n + (uint64_t)i) % amro_topology::kKernelEvents];
      AMRO_CUDA(cudaEventSynchronize(e.b));
      ms[i] = e.ms();
    }
  });
}

// Host arrays of a compiled topology, for the CPU tests of the compiler (not part of the public header).
extern "C" int amro_topology_debug_arrays(const amro_topology* topo, const int64_t** order, const int64_t** day_start,
                                          const int64_t** day_seg_start, const int64_t** seg_row_start,
                                          const double** seg_N, const uint32_t** meta_next, const uint32_t** meta_info,
                                          const uint32_t** meta_nseg, const uint32_t** init_seg,
                                          const int64_t** push_src, const int64_t** push_dst, int64_t* n_push) {
  return guarded([&] {
    if (!topo) fail(AMRO_EINVAL, "NULL topology");
    const HostTopology& H = topo->H;
    *order = H.order.data(); *day_start = H.day_start.data(); *day_seg_start = H.day_seg_start.data();
    *seg_row_start = H.seg_row_start.data(); *seg_N = H.seg_N.data();
    *meta_next = H.meta_next.empty() ? nullptr : H.meta_next.data();
    *meta_info = H.meta_info.empty() ? nullptr : H.meta_info.data();
    *meta_nseg = H.meta_nseg.empty() ? nullptr : H.meta_nseg.data();
    *init_seg = H.init_seg.empty() ? nullptr : H.init_seg.data();
    *push_src = H.push_src.data(); *push_dst = H.push_dst.data(); *n_push = (int64_t)H.push_src.size();
  });
}

// --------------------------------------------------------------------------------------------------------
// amro_simulate
// --------------------------------------------------------------------------------------------------------
namespace {

struct DeviceInputs {  // device views of the run's inputs (uploaded copies or the caller's device pointers)
  const double* params; int64_t p_ld;
  const double* icp; int64_t icp_rs, icp_cs;
  const double* idp; int64_t idp_rs, idp_cs;
  const double* u_col; const double* u_det; int64_t u_ld;
  const double* u_icol; const double* u_idet; int64_t ui_ld;
};

constexpr bool kClusterBarrierByDefault = true;
constexpr int kMaxClusterCtas = 16;   // sm_100 allows clusters of up to 16 CTAs (more than 8 is "non-portable": opted in per kernel)

// amro_sim_args::distance: per-simulation squared distance of the per-day counts to the observed series (device kernel;
// with host pointers the two short series go up and S doubles come back)
void emit_distance(amro_topology* T, amro_sim_args* A, const int32_t* col, const int32_t* det, int64_t S, bool on_device,
                   cudaStream_t st, int& launches) {
  if (!A->distance) return;
  if (!A->observed_colonized && !A->observed_detected) fail(AMRO_EINVAL, "distance needs observed_colonized and/or observed_detected");
  const int64_t Dd = T->H.n_days_out;
  if (on_device) {
    distance_to_data(st, col, det, Dd, S, Dd, A->observed_colonized, A->observed_detected, A->distance);
  } else {
    DevBuf<double> oc, od, out;
    if (A->observed_colonized) oc.upload(A->observed_colonized, (size_t)Dd, st);
    if (A->observed_detected) od.upload(A->observed_detected, (size_t)Dd, st);
    out.alloc((size_t)S);
    distance_to_data(st, col, det, Dd, S, Dd, oc.p, od.p, out.p);
    AMRO_CUDA(cudaMemcpyAsync(A->distance, out.p, sizeof(double) * (size_t)S, cudaMemcpyDeviceToHost, st));
    AMRO_CUDA(cudaStreamSynchronize(st));
  }
  launches++;
}

// amro_sim_args::summary_*: summary_of_total_positive (abm_wards.cpp:526-540) of this run's counts, on the device (summary.cu)
void emit_summaries(amro_topology* T, amro_sim_args* A, const int32_t* col, const int32_t* det, int64_t S, bool on_device, cudaStream_t st,
                    int& launches) {
  if (!A->summary_colonized && !A->summary_detected) return;
  if (A->n_quantiles < 0 || (A->n_quantiles > 0 && !A->quantiles)) fail(AMRO_EINVAL, "summary: quantiles is NULL");
  const int64_t Dd = T->H.n_days_out, nq = A->n_quantiles, cols = 3 + nq;
  DevBuf<double> q, x, out;
  if (nq) q.upload(A->quantiles, (size_t)nq, st);
  x.alloc((size_t)(Dd * S));
  if (!on_device) out.alloc((size_t)(Dd * cols));
  const int32_t* src[2] = {col, det};
  double* dst[2] = {A->summary_colonized, A->summary_detected};
  for (int k = 0; k < 2; ++k) {
    if (!dst[k]) continue;
    counts_to_double(st, src[k], x.p, Dd * S);
    device_summary(st, x.p, Dd, S, Dd, q.p, nq, on_device ? dst[k] : out.p, Dd);
    launches += 2 + (nq ? 4 : 0);
    if (!on_device) {
      AMRO_CUDA(cudaMemcpyAsync(dst[k], out.p, sizeof(double) * Dd * cols, cudaMemcpyDeviceToHost, st));
      AMRO_CUDA(cudaStreamSynchronize(st));
    }
  }
  AMRO_CUDA(cudaStreamSynchronize(st));   // the scratch buffers die here
}

uint64_t schedule_u64(int v) { return (uint64_t)(int64_t)v; }  // int -> unsigned 64-bit, as `uword % int` does (:391-392)

void run_fast(amro_topology* T, amro_sim_args* A, const DeviceInputs& in, cudaStream_t st, bool on_device) {
  const HostTopology& H = T->H;
  const bool replay = A->rng_mode == AMRO_RNG_REPLAY;
  constexpr int64_t kTeamSims = kSimsPerSlice * kSlicesPerTeam;
  const int64_t S = A->n_sims, S_pad = (S + kTeamSims - 1) / kTeamSims * kTeamSims, n_slices = S_pad / kSimsPerSlice;
  const int64_t n_teams = n_slices / kSlicesPerTeam;   // padded simulations have zero parameters and are never read back
  const bool want_traj = A->keep_trajectory || A->state_out;
  bool lean = !replay && !want_traj && !H.has_half && H.identity_order;   // the kernel instance without those cases
  if (std::getenv("AMRO_FAST_NOLEAN")) lean = false;
  // Launch shape (kernels_fast.cu).  TEAM (128-thread CTAs, four per SM, G of them per pair of slices) measured equal or better
  // than WIDE (one 768-thread CTA per slice) on every configuration tried (profiles/sweeps.txt), so it is the default;
  // AMRO_FAST_SHAPE=w forces WIDE (kept for the tests and for comparison); AMRO_FAST_NOLEAN=1 runs the full instance where the
  // lean one would do and AMRO_FAST_VERBOSE prints the launch shape chosen (development).
  int sms = 0;
  AMRO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, T->device));
  int shape = kShapeTeam;
  if (const char* e = std::getenv("AMRO_FAST_SHAPE")) shape = (e[0] == 'w') ? kShapeWide : kShapeTeam;
  if (shape == kShapeWide && fast_shared_bytes(replay, shape) > (size_t)227 * 1024) shape = kShapeTeam;   // replay tables of 16 warps do not fit
  // Team size: all resident CTA slots are dealt out to the teams, the first `extra` teams taking one more, so that every SM
  // hosts the same number of CTAs (measured: a grid that leaves some SMs one CTA short is ~5-10 % slower, because the
  // fuller SMs make their teams' barriers late).  AMRO_FAST_G fixes the team size, AMRO_FAST_CTAS_PER_SM caps the slots.
  int G = 1, extra = 0;
  if (shape == kShapeTeam) {
    int cap = fast_max_coresident_ctas(T->device, replay, lean, shape);
    if (cap <= 0) fail(AMRO_ECUDA, "fast kernel cannot be made resident on this device");
    if (const char* e = std::getenv("AMRO_FAST_CTAS_PER_SM")) cap = std::min(cap, std::max(1, std::atoi(e)) * sms);
    // every warp pays a fixed price per day (barrier, tables, totals): it wants about a dozen 32-record trips to spread it
    // over -- measured on config 2: 7 CTAs per team (11 trips per warp and day) beat both 6 and 8-10 (profiles/sweeps.txt)
    const int64_t useful = std::max<int64_t>(1, H.max_rows_per_day / (11 * kTeamThreads));
    if (n_teams <= cap) {
      G = (int)(cap / n_teams); extra = (int)(cap % n_teams);
      if (G >= useful) { G = (int)useful; extra = 0; }
    }
  }
  if (const char* e = std::getenv("AMRO_FAST_G")) { G = std::max(1, std::atoi(e)); extra = 0; }
  // Teams of 2..8 equal CTAs run as thread-block clusters when all of them can be resident at once: the hardware cluster
  // barrier replaces the counter in global memory.  AMRO_FAST_BARRIER=global / cluster forces one.
  // A grid that leaves a quarter of the CTA slots empty anyway runs the instance compiled for three CTAs per SM (168
  // registers per thread instead of 128: the scheduler has room to overlap the eight pairs of a record): +3 % on config 2
  if (shape == kShapeTeam && !std::getenv("AMRO_FAST_SHAPE") && ((int64_t)n_teams * G + extra <= (int64_t)(kTeamMinBlocks - 1) * sms || std::getenv("AMRO_FAST_ROOMY")))
    shape = kShapeRoomy;   // the lean instance stays lean (with 64-bit REDs there: kernels_fast.cu)
  int use_cluster = 0;
  if (shape != kShapeWide && G > 1 && G <= kMaxClusterCtas && extra == 0) {   // larger teams (few simulations, big days) keep their size
    const char* e = std::getenv("AMRO_FAST_BARRIER");
    const bool want = e ? (e[0] == 'c') : kClusterBarrierByDefault;
    const int Gc = G;
    if (want && fast_max_coresident_clusters(T->device, replay, lean, shape, Gc) >= n_teams) { use_cluster = 1; G = Gc; }
  }
  const int64_t ring_chunks = std::max<int64_t>((H.max_rows_per_day + 31) / 32, 1);
  if (!H.identity_order && H.R >= ((int64_t)1 << 32)) fail(AMRO_EINVAL, "fast path: more than 2^32 rows need the original row order");
  if (ring_chunks * 32 * kSlicesPerTeam * (int64_t)sizeof(uint4) >= ((int64_t)1 << 32)) fail(AMRO_EINVAL, "fast path: day too large for 32-bit slot offsets");

  ensure_size(T->ws_pre, (size_t)(n_teams * 2 * ring_chunks * 32 * kSlicesPerTeam));
  if (want_traj) ensure_size(T->ws_traj, (size_t)(n_slices * std::max<int64_t>(H.R, 1)));
  ensure_size(T->ws_pressure, (size_t)(n_slices * std::max<int64_t>(H.n_seg, 1) * 4));
  ensure_size(T->ws_counts, (size_t)(2 * H.n_days_out * S_pad));
  ensure_size(T->ws_params, (size_t)(S_pad * 6));
  AMRO_CUDA(cudaMemsetAsync(T->ws_counts.p, 0, sizeof(int32_t) * 2 * H.n_days_out * S_pad, st));
  AMRO_CUDA(cudaMemsetAsync(T->ws_params.p, 0, sizeof(double) * S_pad * 6, st));
  AMRO_CUDA(cudaMemcpy2DAsync(T->ws_params.p, S_pad * sizeof(double), in.params, in.p_ld * sizeof(double),
                              S * sizeof(double), 6, cudaMemcpyDeviceToDevice, st));

  FastArgs f{};
  f.meta = T->meta.p;
  f.init_seg = T->init_seg.p; f.init_slot = T->init_slot.p; f.seg_main_next = T->seg_main_next.p;
  f.identity_order = H.identity_order ? 1 : 0;
  f.has_half = H.has_half ? 1 : 0;
  f.day_start = T->day_start.p; f.day_seg_start = T->day_seg_start.p; f.seg_row_start = T->seg_row_start.p;
  f.seg_N = T->seg_N.p;
  f.R = H.R; f.n_init = H.n_init; f.n_days = H.total_days + 1; f.n_seg = H.n_seg; f.ring_chunks = ring_chunks;
  f.params = T->ws_params.p; f.p_ld = S_pad;
  f.icp = in.icp; f.icp_rs = in.icp_rs; f.icp_cs = in.icp_cs;
  f.idp = in.idp; f.idp_rs = in.idp_rs; f.idp_cs = in.idp_cs;
  f.S = S; f.n_slices = n_slices;
  f.seed = (uint64_t)A->seed; f.sim_offset = (uint64_t)A->sim_offset;
  f.ttd = A->time_to_detect;
  for (int r = 0; r < 10; ++r) {
    f.rk[2 * r] = (uint32_t)f.seed + (uint32_t)r * 0x9E3779B9u;
    f.rk[2 * r + 1] = (uint32_t)(f.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  f.tsh = schedule_u64(A->testing_schedule_hospitalized); f.tsa = schedule_u64(A->testing_schedule_arrivals);
  f.u_col = in.u_col; f.u_det = in.u_det; f.u_ld = in.u_ld;
  f.u_icol = in.u_icol; f.u_idet = in.u_idet; f.ui_ld = in.ui_ld;
  DevBuf<double> d_p;
  if (A->p_out) {
    if (!replay) fail(AMRO_EINVAL, "p_out on the fast path needs rng_mode = AMRO_RNG_REPLAY");
    if (on_device) f.p_out = A->p_out;
    else { d_p.alloc((size_t)(H.R * S)); AMRO_CUDA(cudaMemsetAsync(d_p.p, 0xFF, sizeof(double) * H.R * S, st)); f.p_out = d_p.p; }
  }
  f.pre = T->ws_pre.p; f.traj = want_traj ? T->ws_traj.p : nullptr;
  f.pressure = T->ws_pressure.p;
  f.counts_col = T->ws_counts.p; f.counts_det = T->ws_counts.p + H.n_days_out * S_pad; f.counts_ld = H.n_days_out;
  ensure_size(T->ws_barrier, (size_t)n_teams);
  AMRO_CUDA(cudaMemsetAsync(T->ws_barrier.p, 0, sizeof(unsigned) * n_teams, st));
  f.team_barrier = T->ws_barrier.p; f.ctas_per_team = G; f.teams_plus_one = extra; f.cluster_barrier = use_cluster;

  if (std::getenv("AMRO_FAST_VERBOSE"))   // development: the launch shape chosen
    std::fprintf(stderr, "[amro] fast launch: shape %d lean %d teams %lld G %d extra %d cluster %d grid %lld threads %d smem %zu\n", shape, (int)lean,
                 (long long)n_teams, G, extra, use_cluster, (long long)(n_teams * G + extra), fast_threads(shape), fast_shared_bytes(replay, shape));
  EventPair& ev = T->next_kernel_events();
  AMRO_CUDA(cudaEventRecord(ev.a, st));
  fast_launch(st, f, replay, lean, shape, (int)(n_teams * G + extra));
  AMRO_CUDA(cudaEventRecord(ev.b, st));
  A->launches = 1;
  A->kernel_kind = AMRO_KERNEL_CODES; A->grid = (int)(n_teams * G + extra); A->ctas_per_team = G;
  A->team_barrier = (G == 1 && extra == 0) ? 0 : (use_cluster ? 2 : 1);

  // ---- outputs ---------------------------------------------------------------------------------------------
  const size_t cbytes = sizeof(int32_t) * (size_t)(H.n_days_out * S);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  // host destinations are pageable as a rule: stage through the handle's pinned buffer (full PCIe speed), copy out after the sync
  int32_t* staged = (!on_device && (A->counts_colonized || A->counts_detected)) ? T->pinned(2 * (size_t)(H.n_days_out * S)) : nullptr;
  if (A->counts_colonized) AMRO_CUDA(cudaMemcpyAsync(staged ? staged : A->counts_colonized, f.counts_col, cbytes, kind, st));
  if (A->counts_detected) AMRO_CUDA(cudaMemcpyAsync(staged ? staged + H.n_days_out * S : A->counts_detected, f.counts_det, cbytes, kind, st));
  if (A->day_moments) {
    if (on_device) day_moments(st, f.counts_col, f.counts_det, H.n_days_out, S, H.n_days_out, A->day_moments, H.n_days_out);
    else {
      DevBuf<double> dm;
      dm.alloc((size_t)(4 * H.n_days_out));
      day_moments(st, f.counts_col, f.counts_det, H.n_days_out, S, H.n_days_out, dm.p, H.n_days_out);
      AMRO_CUDA(cudaMemcpyAsync(A->day_moments, dm.p, sizeof(double) * 4 * H.n_days_out, cudaMemcpyDeviceToHost, st));
      AMRO_CUDA(cudaStreamSynchronize(st));
    }
    A->launches++;
  }
  emit_distance(T, A, f.counts_col, f.counts_det, S, on_device, st, A->launches);
  emit_summaries(T, A, f.counts_col, f.counts_det, S, on_device, st, A->launches);
  const int64_t* pos_of = H.identity_order ? nullptr : T->pos_of.p;
  if (A->state_out) {
    if (on_device) {
      fast_expand(st, f.traj, pos_of, H.R, S, 0, H.R, 0, S, A->state_out, A->state_out + A->so_ld * S, A->so_ld);
      A->launches++;
    } else {
      // The reference's float64 columns leave in tiles of (a block of rows) x (8 simulations): a tile is expanded on the device,
      // copied into one of two PINNED staging buffers on a second stream, and copied from there into the caller's (pageable)
      // matrix by the host while the next tile is in flight -- expansion, PCIe transfer and the host copy overlap.
      AMRO_CUDA(cudaStreamSynchronize(st));   // the trajectory is complete
      const int64_t tile_rows = std::max<int64_t>(1, std::min<int64_t>(H.R, ((int64_t)1 << 20)));
      const size_t tile_elems = (size_t)tile_rows * 16;   // 8 C columns + 8 D columns
      DevBuf<double> dbuf[2];
      double* hbuf[2] = {nullptr, nullptr};
      cudaStream_t s_copy = nullptr;
      cudaEvent_t expanded[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
      struct Tile { int64_t r0, nr, s0, ns; };
      auto cleanup = [&] {
        for (int b = 0; b < 2; ++b) { if (hbuf[b]) cudaFreeHost(hbuf[b]); if (expanded[b]) cudaEventDestroy(expanded[b]); if (copied[b]) cudaEventDestroy(copied[b]); }
        if (s_copy) cudaStreamDestroy(s_copy);
      };
      try {
        AMRO_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
          dbuf[b].alloc(tile_elems);
          AMRO_CUDA(cudaMallocHost(&hbuf[b], tile_elems * sizeof(double)));
          AMRO_CUDA(cudaEventCreateWithFlags(&expanded[b], cudaEventDisableTiming));
          AMRO_CUDA(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
        }
        std::vector<Tile> tiles;
        for (int64_t s0 = 0; s0 < S; s0 += 8)
          for (int64_t r0 = 0; r0 < H.R; r0 += tile_rows) tiles.push_back(Tile{r0, std::min(tile_rows, H.R - r0), s0, std::min<int64_t>(8, S - s0)});
        // staging [16][nr] -> the caller's C and D columns, one host thread per simulation of the tile (a freshly allocated
        // result matrix is first touched here: the page faults spread over the threads)
        auto to_caller = [&](const Tile& t, int b) {
          auto one = [&, b](int64_t k) {
            std::memcpy(A->state_out + A->so_ld * (t.s0 + k) + t.r0, hbuf[b] + k * t.nr, sizeof(double) * (size_t)t.nr);
            std::memcpy(A->state_out + A->so_ld * (S + t.s0 + k) + t.r0, hbuf[b] + (8 + k) * t.nr, sizeof(double) * (size_t)t.nr);
          };
          if (t.nr * t.ns < (1 << 16)) { for (int64_t k = 0; k < t.ns; ++k) one(k); return; }
          std::vector<std::thread> pool;
          for (int64_t k = 1; k < t.ns; ++k) pool.emplace_back(one, k);
          one(0);
          for (auto& th : pool) th.join();
        };
        for (size_t k = 0; k < tiles.size(); ++k) {
          const int b = (int)(k & 1);
          const Tile& t = tiles[k];
          if (k >= 2) AMRO_CUDA(cudaStreamWaitEvent(st, copied[b], 0));      // the device buffer has been read
          fast_expand(st, f.traj, pos_of, H.R, S, t.r0, t.nr, t.s0, t.ns, dbuf[b].p, dbuf[b].p + 8 * t.nr, t.nr);
          A->launches++;
          AMRO_CUDA(cudaEventRecord(expanded[b], st));
          AMRO_CUDA(cudaStreamWaitEvent(s_copy, expanded[b], 0));   // hbuf[b] was handed to the caller one iteration ago
          AMRO_CUDA(cudaMemcpyAsync(hbuf[b], dbuf[b].p, sizeof(double) * 16 * (size_t)t.nr, cudaMemcpyDeviceToHost, s_copy));
          AMRO_CUDA(cudaEventRecord(copied[b], s_copy));
          if (k >= 1) {   // the previous tile: wait for its transfer, hand it to the caller while this tile is in flight
            AMRO_CUDA(cudaEventSynchronize(copied[b ^ 1]));
            to_caller(tiles[k - 1], b ^ 1);
          }
        }
        if (!tiles.empty()) {
          const int b = (int)((tiles.size() - 1) & 1);
          AMRO_CUDA(cudaEventSynchronize(copied[b]));
          to_caller(tiles.back(), b);
        }
      } catch (...) { cleanup(); throw; }
      cleanup();
    }
  }
  if (A->pressure_out) {
    if (on_device) { fast_export_pressure(st, f.pressure, H.n_seg, S, A->pressure_out, H.n_seg); }
    else {
      DevBuf<int32_t> pb;
      pb.alloc((size_t)(H.n_seg * S));
      fast_export_pressure(st, f.pressure, H.n_seg, S, pb.p, H.n_seg);
      AMRO_CUDA(cudaMemcpyAsync(A->pressure_out, pb.p, sizeof(int32_t) * H.n_seg * S, cudaMemcpyDeviceToHost, st));
      AMRO_CUDA(cudaStreamSynchronize(st));
    }
    A->launches++;
  }
  if (A->p_out && !on_device) AMRO_CUDA(cudaMemcpyAsync(A->p_out, d_p.p, sizeof(double) * H.R * S, cudaMemcpyDeviceToHost, st));
  A->path_used = AMRO_PATH_FAST;
  if (on_device && A->async) { A->kernel_ms = 0.f; return; }   // everything is enqueued on the caller's stream
  AMRO_CUDA(cudaStreamSynchronize(st));
  if (staged && A->counts_colonized) std::memcpy(A->counts_colonized, staged, cbytes);
  if (staged && A->counts_detected) std::memcpy(A->counts_detected, staged + H.n_days_out * S, cbytes);
  A->kernel_ms = ev.ms();
}

// BITS path (kernels_bits.cu): summary runs of unit-weight hospitals in free-running mode with time_to_detect <= 2 -- the
// state bit-sliced over simulations, a team = 64 simulations.  Same topology arrays, ring slots and outputs as run_fast.
bool bits_eligible(const amro_topology* T, const amro_sim_args* A) {
  const HostTopology& H = T->H;
  if (std::getenv("AMRO_FAST_NOBITS")) return false;
  if (!std::getenv("AMRO_FAST_BITS")) {
    // The bit-sliced kernel counts a ward-day's pressure once per run that touches it.  Hospitals whose wards are many times
    // longer than a run (one ward of 10 000 beds: every warp would count the whole ward to step 280 records) go to the kernel
    // that pushes pressure instead (measured: 1.16 against 1.89 ms, profiles/sweeps.txt).  The run length is estimated the way
    // run_bits sizes its teams: all CTA slots dealt out to the teams, at least 8 trips of 32 records per warp.
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, T->device) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
    const int64_t n_teams = (A->n_sims + bits_sims_per_team() - 1) / bits_sims_per_team();
    const int64_t pairs = std::max<int64_t>(1, (int64_t)sms * kTeamMinBlocks / std::max<int64_t>(n_teams, 1)) * (kTeamThreads / 32) / bits_words_per_team();
    const int64_t run = std::max<int64_t>(256, H.max_rows_per_day / pairs);
    if (H.max_seg_rows > 8 * run) return false;
  }
  return A->rng_mode != AMRO_RNG_REPLAY && !A->keep_trajectory && !A->state_out && !A->p_out &&
         A->time_to_detect >= 0 && A->time_to_detect <= kBitsMaxTtd;
}

void run_bits(amro_topology* T, amro_sim_args* A, const DeviceInputs& in, cudaStream_t st, bool on_device) {
  const HostTopology& H = T->H;
  const int64_t kTeamSims = bits_sims_per_team();
  const int64_t S = A->n_sims, S_pad = (S + kTeamSims - 1) / kTeamSims * kTeamSims, n_teams = S_pad / kTeamSims;
  const int ttd = A->time_to_detect;
  int sms = 0;
  AMRO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, T->device));
  // Team size: as in run_fast -- all resident CTA slots dealt out to the teams, capped where a warp would get fewer than
  // about `trips` 32-record trips a day (every warp pays a fixed price per day: barrier, tables, count flushes).
  // AMRO_BITS_G fixes the team size, AMRO_BITS_CTAS_PER_SM caps the slots, AMRO_BITS_TRIPS sets the target.
  int shape = kShapeTeam;
  if (const char* e = std::getenv("AMRO_BITS_SHAPE")) shape = e[0] == 'r' ? kShapeRoomy : kShapeTeam;
  const bool shape_forced = std::getenv("AMRO_BITS_SHAPE") != nullptr;
  const bool half = H.has_half;
  int cap = bits_max_coresident_ctas(T->device, ttd, shape, half);
  if (cap <= 0) fail(AMRO_ECUDA, "bit-sliced kernel cannot be made resident on this device");
  if (const char* e = std::getenv("AMRO_BITS_CTAS_PER_SM")) cap = std::min(cap, std::max(1, std::atoi(e)) * sms);
  // A pair of warps steps a run of records (one warp per word).  Team size: the largest number of whole CTAs-per-SM layers
  // (the grid then loads every SM alike) that still leaves a warp about `trips` 32-record trips a day -- measured on config 2:
  // 18 CTAs per team (2 per SM, 8.7 trips) beat 12, 16, 24 and 36 (profiles/sweeps.txt).
  // (A ward-day's pressure is counted by every run that touches it, but counting a chunk costs a thirteenth of stepping it and a
  // warp counts its wards whole however short its run is: more, shorter runs still win -- measured on a one-ward hospital.)
  int trips = 8;
  if (const char* e = std::getenv("AMRO_BITS_TRIPS")) trips = std::max(1, std::atoi(e));
  int G = 1, extra = 0;
  const int64_t useful = std::max<int64_t>(1, H.max_rows_per_day * bits_words_per_team() / ((int64_t)trips * kTeamThreads));
  if (n_teams <= cap) {
    G = (int)(cap / n_teams); extra = (int)(cap % n_teams);
    if (G >= useful) {
      extra = 0;
      G = 1;
      bool found = false;
      for (int layers = cap / sms; layers >= 1; --layers) {
        const int g = (int)((int64_t)layers * sms / n_teams);
        if (g >= 1 && g <= useful + useful / 8) { G = g; found = true; break; }
      }
      if (!found) G = (int)useful;   // few teams: even one CTA per SM would be more than the records feed -- part of the SMs then
    }
  }
  if (const char* e = std::getenv("AMRO_BITS_G")) { G = std::max(1, std::atoi(e)); extra = 0; }
  if (!shape_forced && (int64_t)n_teams * G + extra <= (int64_t)(kTeamMinBlocks - 1) * sms) shape = kShapeRoomy;
  int use_cluster = 0;
  if (G > 1 && G <= kMaxClusterCtas && extra == 0) {
    const char* e = std::getenv("AMRO_FAST_BARRIER");
    const bool want = e ? (e[0] == 'c') : kClusterBarrierByDefault;
    if (want && bits_max_coresident_clusters(ttd, shape, G, half) >= n_teams) use_cluster = 1;
  }
  if ((G > 1 || extra > 0) && !use_cluster && (int64_t)n_teams * G + extra > bits_max_coresident_ctas(T->device, ttd, shape, half))
    fail(AMRO_ECUDA, "bit-sliced kernel: the teams cannot be co-resident");
  const int64_t ring_chunks = std::max<int64_t>((H.max_rows_per_day + 31) / 32, 1);
  if (!H.identity_order && H.R >= ((int64_t)1 << 32)) fail(AMRO_EINVAL, "fast path: more than 2^32 rows need the original row order");
  if (ring_chunks * 32 * kSlicesPerTeam * (int64_t)sizeof(uint4) >= ((int64_t)1 << 32)) fail(AMRO_EINVAL, "fast path: day too large for 32-bit slot offsets");
  const int64_t n_days = H.total_days + 1;
  const int64_t max_segs = std::max<int64_t>(H.max_segs_per_day, 1);
  const int64_t press_days = n_days + 1;   // the export buffer's days (only when pressure_out is asked for)

  T->ensure_pull();
  ensure_size(T->ws_pre, (size_t)(n_teams * 2 * ring_chunks * 32 * kSlicesPerTeam));
  if (A->pressure_out) ensure_size(T->ws_pressure, (size_t)(n_teams * press_days * max_segs * kTeamSims));
  ensure_size(T->ws_counts, (size_t)(2 * H.n_days_out * S_pad));
  ensure_size(T->ws_params, (size_t)(S_pad * 6));
  AMRO_CUDA(cudaMemsetAsync(T->ws_counts.p, 0, sizeof(int32_t) * 2 * H.n_days_out * S_pad, st));
  AMRO_CUDA(cudaMemsetAsync(T->ws_params.p, 0, sizeof(double) * S_pad * 6, st));
  AMRO_CUDA(cudaMemcpy2DAsync(T->ws_params.p, S_pad * sizeof(double), in.params, in.p_ld * sizeof(double),
                              S * sizeof(double), 6, cudaMemcpyDeviceToDevice, st));

  FastArgs f{};
  f.meta = T->meta.p;
  f.init_seg = T->init_seg.p; f.init_slot = T->init_slot.p; f.seg_main_next = T->seg_main_next.p;
  f.identity_order = H.identity_order ? 1 : 0;
  f.has_half = 0;
  f.day_start = T->day_start.p; f.day_seg_start = T->day_seg_start.p; f.seg_row_start = T->seg_row_start.p;
  f.seg_N = T->seg_N.p;
  f.R = H.R; f.n_init = H.n_init; f.n_days = n_days; f.n_seg = H.n_seg; f.ring_chunks = ring_chunks;
  f.params = T->ws_params.p; f.p_ld = S_pad;
  f.icp = in.icp; f.icp_rs = in.icp_rs; f.icp_cs = in.icp_cs;
  f.idp = in.idp; f.idp_rs = in.idp_rs; f.idp_cs = in.idp_cs;
  f.S = S; f.n_slices = S_pad / kSimsPerSlice;
  f.seed = (uint64_t)A->seed; f.sim_offset = (uint64_t)A->sim_offset;
  f.ttd = ttd;
  for (int r = 0; r < 10; ++r) {
    f.rk[2 * r] = (uint32_t)f.seed + (uint32_t)r * 0x9E3779B9u;
    f.rk[2 * r + 1] = (uint32_t)(f.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
  }
  f.tsh = schedule_u64(A->testing_schedule_hospitalized); f.tsa = schedule_u64(A->testing_schedule_arrivals);
  f.pre = T->ws_pre.p; f.traj = nullptr;
  f.pressure = A->pressure_out ? T->ws_pressure.p : nullptr;
  f.src_mask = T->src_mask.p; f.half_mask = H.has_half ? T->half_mask.p : nullptr; f.day_chunk_start = T->day_chunk_start.p;
  {
    // Counting a ward's pressure: deep branch-free steps (12 chunks per ward-day in one round trip) where the launch is bound by
    // latency -- few warps per SM (config 2: 8) or long wards (configs 4 and 5) -- and lean groups where
    // many warps share an SM and the wards are short (the 64k-particle sweep).  AMRO_BITS_PULL=deep|lean overrides.
    const double warps_per_sm = (double)(n_teams * G + extra) * (kTeamThreads / 32) / sms;
    f.pull_deep = (warps_per_sm <= 8.5 || H.max_seg_rows > 1024) ? 1 : 0;
    if (const char* e = std::getenv("AMRO_BITS_PULL")) f.pull_deep = e[0] == 'd';
  }
  f.counts_col = T->ws_counts.p; f.counts_det = T->ws_counts.p + H.n_days_out * S_pad; f.counts_ld = H.n_days_out;
  ensure_size(T->ws_barrier, (size_t)n_teams);
  AMRO_CUDA(cudaMemsetAsync(T->ws_barrier.p, 0, sizeof(unsigned) * n_teams, st));
  f.team_barrier = T->ws_barrier.p; f.ctas_per_team = G; f.teams_plus_one = extra; f.cluster_barrier = use_cluster;
  f.max_segs_per_day = max_segs; f.press_days = (int)press_days;

  if (std::getenv("AMRO_FAST_VERBOSE"))
    std::fprintf(stderr, "[amro] bits launch: shape %d ttd %d teams %lld G %d extra %d cluster %d grid %lld threads %d smem %zu\n", shape, ttd,
                 (long long)n_teams, G, extra, use_cluster, (long long)(n_teams * G + extra), kTeamThreads, bits_shared_bytes(half));
  EventPair& ev = T->next_kernel_events();
  AMRO_CUDA(cudaEventRecord(ev.a, st));
  bits_launch(st, f, shape, (int)(n_teams * G + extra));
  AMRO_CUDA(cudaEventRecord(ev.b, st));
  A->launches = 1;
  A->kernel_kind = AMRO_KERNEL_BITS; A->grid = (int)(n_teams * G + extra); A->ctas_per_team = G;
  A->team_barrier = (G == 1 && extra == 0) ? 0 : (use_cluster ? 2 : 1);

  const size_t cbytes = sizeof(int32_t) * (size_t)(H.n_days_out * S);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  // host destinations are pageable as a rule: stage through the handle's pinned buffer (full PCIe speed), copy out after the sync
  int32_t* staged = (!on_device && (A->counts_colonized || A->counts_detected)) ? T->pinned(2 * (size_t)(H.n_days_out * S)) : nullptr;
  if (A->counts_colonized) AMRO_CUDA(cudaMemcpyAsync(staged ? staged : A->counts_colonized, f.counts_col, cbytes, kind, st));
  if (A->counts_detected) AMRO_CUDA(cudaMemcpyAsync(staged ? staged + H.n_days_out * S : A->counts_detected, f.counts_det, cbytes, kind, st));
  if (A->day_moments) {
    if (on_device) day_moments(st, f.counts_col, f.counts_det, H.n_days_out, S, H.n_days_out, A->day_moments, H.n_days_out);
    else {
      DevBuf<double> dm;
      dm.alloc((size_t)(4 * H.n_days_out));
      day_moments(st, f.counts_col, f.counts_det, H.n_days_out, S, H.n_days_out, dm.p, H.n_days_out);
      AMRO_CUDA(cudaMemcpyAsync(A->day_moments, dm.p, sizeof(double) * 4 * H.n_days_out, cudaMemcpyDeviceToHost, st));
      AMRO_CUDA(cudaStreamSynchronize(st));
    }
    A->launches++;
  }
  emit_distance(T, A, f.counts_col, f.counts_det, S, on_device, st, A->launches);
  emit_summaries(T, A, f.counts_col, f.counts_det, S, on_device, st, A->launches);
  if (A->pressure_out) {
    if (on_device) bits_export_pressure(st, f.pressure, f.day_seg_start, n_days, max_segs, press_days, S, A->pressure_out, H.n_seg);
    else {
      DevBuf<int32_t> pb;
      pb.alloc((size_t)(H.n_seg * S));
      bits_export_pressure(st, f.pressure, f.day_seg_start, n_days, max_segs, press_days, S, pb.p, H.n_seg);
      AMRO_CUDA(cudaMemcpyAsync(A->pressure_out, pb.p, sizeof(int32_t) * H.n_seg * S, cudaMemcpyDeviceToHost, st));
      AMRO_CUDA(cudaStreamSynchronize(st));
    }
    A->launches++;
  }
  A->path_used = AMRO_PATH_FAST;
  if (on_device && A->async) { A->kernel_ms = 0.f; return; }   // everything is enqueued on the caller's stream
  AMRO_CUDA(cudaStreamSynchronize(st));
  if (staged && A->counts_colonized) std::memcpy(A->counts_colonized, staged, cbytes);
  if (staged && A->counts_detected) std::memcpy(A->counts_detected, staged + H.n_days_out * S, cbytes);
  A->kernel_ms = ev.ms();
}

void run_general(amro_topology* T, amro_sim_args* A, const DeviceInputs& in, cudaStream_t st, bool on_device) {
  const HostTopology& H = T->H;
  if (T->generated) fail(AMRO_EINVAL, "a hospital generated on the device has no host tables: it runs the fused kernels only");
  T->ensure_general();
  const bool replay = A->rng_mode == AMRO_RNG_REPLAY;
  const int64_t S = A->n_sims, R = H.R, D = H.total_days + 1;
  DevBuf<double> planes, F, tmp, d_p;
  DevBuf<int32_t> counts;
  double *C, *Dp;
  int64_t ld;
  if (on_device && A->state_out) { C = A->state_out; Dp = A->state_out + A->so_ld * S; ld = A->so_ld; }
  else { planes.alloc((size_t)(2 * R * S)); C = planes.p; Dp = planes.p + R * S; ld = R; }
  F.alloc((size_t)(std::max<int64_t>(H.max_segs_per_day, 1) * S));
  DevBuf<uint32_t> dither;
  if (!replay) dither.alloc((size_t)((std::max<int64_t>(H.max_segs_per_day, 1) + 1) * S));
  int64_t max_push = 0;
  for (int64_t d = 0; d < D; ++d) max_push = std::max(max_push, H.day_push_start[d + 1] - H.day_push_start[d]);
  tmp.alloc((size_t)(std::max<int64_t>(max_push, 1) * 2 * S));
  counts.alloc((size_t)(2 * H.n_days_out * S));
  AMRO_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(int32_t) * 2 * H.n_days_out * S, st));

  GenArgs g{};
  g.C = C; g.D = Dp; g.ld = ld;
  g.h = T->h.p; g.w = T->w.p; g.hw_stride = 1;
  g.order = T->order.p; g.seg_of_pos = T->seg_of_pos.p; g.seg_row_start = T->seg_row_start.p; g.seg_N = T->seg_N.p;
  g.params = in.params; g.p_ld = in.p_ld; g.S = S; g.F = F.p; g.ttd = A->time_to_detect;
  g.replay = replay; g.u_col = in.u_col; g.u_det = in.u_det; g.u_ld = in.u_ld;
  g.seed = (uint64_t)A->seed; g.sim_offset = (uint64_t)A->sim_offset;
  g.dither = replay ? nullptr : dither.p; g.run_dither_row = std::max<int64_t>(H.max_segs_per_day, 1);
  if (A->p_out) {
    if (on_device) g.p_out = A->p_out;
    else { d_p.alloc((size_t)(R * S)); AMRO_CUDA(cudaMemsetAsync(d_p.p, 0xFF, sizeof(double) * R * S, st)); g.p_out = d_p.p; }
    g.p_out_ld = R;
  }
  const uint64_t tsh = schedule_u64(A->testing_schedule_hospitalized), tsa = schedule_u64(A->testing_schedule_arrivals);

  EventPair ev;
  AMRO_CUDA(cudaEventRecord(ev.a, st));
  int launches = 1;
  gen_init(st, C, Dp, ld, R, S, H.n_init, in.icp, in.icp_rs, in.icp_cs, in.idp, in.idp_rs, in.idp_cs, replay,
           in.u_icol, in.u_idet, in.ui_ld, g.seed, g.sim_offset);
  if (!replay) { gen_run_dither(st, g); launches++; }
  for (int64_t d = 0; d < D; ++d) {
    const int64_t sa = H.day_seg_start[d], sb = H.day_seg_start[d + 1];
    const int64_t ra = H.day_start[d], rb = H.day_start[d + 1];
    if (rb > ra) {
      gen_pressure(st, g, sa, sb);
      gen_update(st, g, ra, rb, sa, ((uint64_t)d % tsh) == 0, ((uint64_t)d % tsa) == 0);
      launches += 2;
    }
    const int64_t pa = H.day_push_start[d], pb = H.day_push_start[d + 1];
    if (pb > pa) { gen_propagate(st, C, Dp, ld, S, T->push_src.p + pa, T->push_dst.p + pa, pb - pa, tmp.p); launches += 2; }
  }
  const bool want_sum = A->summary_colonized || A->summary_detected;
  if (A->counts_colonized || A->day_moments || A->distance || want_sum) { gen_count(st, C, ld, R, S, T->count_day.p, 1, counts.p, H.n_days_out); launches++; }
  if (A->counts_detected || A->day_moments || A->distance || want_sum) { gen_count(st, Dp, ld, R, S, T->count_day.p, 0, counts.p + H.n_days_out * S, H.n_days_out); launches++; }
  DevBuf<double> dm;
  if (A->day_moments) {
    double* dst = A->day_moments;
    if (!on_device) { dm.alloc((size_t)(4 * H.n_days_out)); dst = dm.p; }
    day_moments(st, counts.p, counts.p + H.n_days_out * S, H.n_days_out, S, H.n_days_out, dst, H.n_days_out);
    launches++;
  }
  emit_distance(T, A, counts.p, counts.p + H.n_days_out * S, S, on_device, st, launches);
  emit_summaries(T, A, counts.p, counts.p + H.n_days_out * S, S, on_device, st, launches);
  AMRO_CUDA(cudaEventRecord(ev.b, st));
  AMRO_CUDA(cudaGetLastError());

  const size_t cbytes = sizeof(int32_t) * (size_t)(H.n_days_out * S);
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (A->counts_colonized) AMRO_CUDA(cudaMemcpyAsync(A->counts_colonized, counts.p, cbytes, kind, st));
  if (A->counts_detected) AMRO_CUDA(cudaMemcpyAsync(A->counts_detected, counts.p + H.n_days_out * S, cbytes, kind, st));
  if (A->state_out && !on_device)
    AMRO_CUDA(cudaMemcpy2DAsync(A->state_out, A->so_ld * sizeof(double), planes.p, R * sizeof(double), R * sizeof(double),
                                2 * S, cudaMemcpyDeviceToHost, st));
  if (A->p_out && !on_device) AMRO_CUDA(cudaMemcpyAsync(A->p_out, d_p.p, sizeof(double) * R * S, cudaMemcpyDeviceToHost, st));
  if (A->day_moments && !on_device) AMRO_CUDA(cudaMemcpyAsync(A->day_moments, dm.p, sizeof(double) * 4 * H.n_days_out, cudaMemcpyDeviceToHost, st));
  AMRO_CUDA(cudaStreamSynchronize(st));
  A->kernel_ms = ev.ms();
  A->launches = launches;
  A->path_used = AMRO_PATH_GENERAL;
  A->kernel_kind = AMRO_KERNEL_FLOAT64; A->grid = 0; A->ctas_per_team = 0; A->team_barrier = 0;
}

}  // namespace

extern "C" int amro_simulate(amro_topology* T, amro_sim_args* A) {
  return guarded([&] {
    if (!T || !A) fail(AMRO_EINVAL, "NULL argument");
    if (T->device < 0) fail(AMRO_ECUDA, "topology was compiled host-only (device < 0); amro_b200 has no CPU fallback");
    std::lock_guard<std::mutex> lock(T->mu);
    use_device(T->device);
    const HostTopology& H = T->H;
    const int64_t S = A->n_sims;
    if (S < 1) fail(AMRO_EINDEX, "Mat::submat(): indices out of bounds or incorrectly used");  // :365 with zero simulations
    if (!A->parameters || !A->icp || !A->idp) fail(AMRO_EINVAL, "parameters / icp / idp must not be NULL");
    if (A->p_rows < S) fail(AMRO_EINDEX, "Mat::operator(): index out of bounds");                // :98 parameters(s, .)
    if (A->p_ld < S) fail(AMRO_EINVAL, "p_ld (%lld) is smaller than n_sims (%lld)", (long long)A->p_ld, (long long)S);
    if (A->state_out && A->so_ld < H.R) fail(AMRO_EINVAL, "so_ld (%lld) is smaller than the number of rows (%lld)", (long long)A->so_ld, (long long)H.R);
    if (A->async && !A->pointers_on_device) fail(AMRO_EINVAL, "async needs pointers_on_device");
    if (A->async && (A->summary_colonized || A->summary_detected)) fail(AMRO_EINVAL, "summaries are not available in async runs (their scratch is per call)");
    if (A->async && A->path == AMRO_PATH_GENERAL) fail(AMRO_EINVAL, "async runs need the fused kernels (path AUTO or FAST)");
    if (A->testing_schedule_hospitalized == 0 || A->testing_schedule_arrivals == 0)
      fail(AMRO_EINVAL, "testing schedule of 0: the reference divides by zero here (abm_wards.cpp:391-392)");
    const bool replay = A->rng_mode == AMRO_RNG_REPLAY;
    if (replay && (!A->u_col || !A->u_det || !A->u_icol || !A->u_idet))
      fail(AMRO_EINVAL, "replay mode needs u_col, u_det, u_icol and u_idet");
    if (replay && (A->u_ld < H.R || A->ui_ld < H.n_init))
      fail(AMRO_EINVAL, "u_ld / ui_ld (%lld / %lld) are smaller than the rows they stride over (%lld / %lld)", (long long)A->u_ld,
           (long long)A->ui_ld, (long long)H.R, (long long)H.n_init);
    if ((A->state_out || A->p_out) && (double)H.R * (double)S * 8.0 > 64e9)
      fail(AMRO_EINVAL, "state_out / p_out of %lld x %lld float64 is too large; run a slice of the simulations", (long long)H.R, (long long)S);

    int path = A->path;
    const bool fast_ok = H.fast_eligible && A->time_to_detect >= 0 && A->time_to_detect <= kMaxCountdown &&
                         H.total_days <= kMaxCountdown && (A->sim_offset % 32 == 0);   // a team steps an aligned half of a group of 32 simulations (rng.cuh)
    if (path == AMRO_PATH_AUTO) path = fast_ok ? AMRO_PATH_FAST : AMRO_PATH_GENERAL;
    if (path == AMRO_PATH_FAST && !fast_ok)
      fail(AMRO_EINVAL, "fast path not available: %s", H.fast_eligible ? "time_to_detect / number of days / sim_offset out of range" : H.why_not_fast.c_str());
    if (A->async && path != AMRO_PATH_FAST) fail(AMRO_EINVAL, "async runs need the fused kernels; this run takes the float64 path");
    if (path == AMRO_PATH_GENERAL && (double)H.R * (double)S * 16.0 > 120e9)
      fail(AMRO_EINVAL, "general path needs %lld x %lld x 16 bytes of state; too large", (long long)H.R, (long long)S);

    const bool on_device = A->pointers_on_device != 0;
    cudaStream_t st = on_device ? (cudaStream_t)A->stream : (cudaStream_t)0;
    EventPair ev;
    AMRO_CUDA(cudaEventRecord(ev.a, st));

    DeviceInputs in{};
    DevBuf<double> d_par, d_icp, d_idp, d_uc, d_ud, d_uic, d_uid;
    if (on_device) {
      in.params = A->parameters; in.p_ld = A->p_ld;
      in.icp = A->icp; in.icp_rs = A->icp_rs; in.icp_cs = A->icp_cs;
      in.idp = A->idp; in.idp_rs = A->idp_rs; in.idp_cs = A->idp_cs;
      in.u_col = A->u_col; in.u_det = A->u_det; in.u_ld = A->u_ld;
      in.u_icol = A->u_icol; in.u_idet = A->u_idet; in.ui_ld = A->ui_ld;
    } else {
      upload_matrix(d_par, A->parameters, S, 6, A->p_ld, st);
      in.params = d_par.p; in.p_ld = S;
      auto up_prob = [&](DevBuf<double>& dst, const double* src, int64_t rs, int64_t cs, const double*& dp, int64_t& drs, int64_t& dcs) {
        if (rs == 0 && cs == 0) { dst.upload(src, 1, st); dp = dst.p; drs = 0; dcs = 0; return; }
        std::vector<double> dense((size_t)(H.n_init * S));
        for (int64_t s = 0; s < S; ++s) for (int64_t r = 0; r < H.n_init; ++r) dense[r + H.n_init * s] = src[r * rs + s * cs];
        dst.upload(dense.data(), dense.size(), st);
        AMRO_CUDA(cudaStreamSynchronize(st));  // `dense` dies at the end of this lambda
        dp = dst.p; drs = 1; dcs = H.n_init;
      };
      up_prob(d_icp, A->icp, A->icp_rs, A->icp_cs, in.icp, in.icp_rs, in.icp_cs);
      up_prob(d_idp, A->idp, A->idp_rs, A->idp_cs, in.idp, in.idp_rs, in.idp_cs);
      if (replay) {
        upload_matrix(d_uc, A->u_col, H.R, S, A->u_ld, st); upload_matrix(d_ud, A->u_det, H.R, S, A->u_ld, st);
        upload_matrix(d_uic, A->u_icol, H.n_init, S, A->ui_ld, st); upload_matrix(d_uid, A->u_idet, H.n_init, S, A->ui_ld, st);
        in.u_col = d_uc.p; in.u_det = d_ud.p; in.u_ld = H.R; in.u_icol = d_uic.p; in.u_idet = d_uid.p; in.ui_ld = H.n_init;
      }
    }
    if (path == AMRO_PATH_FAST && bits_eligible(T, A)) run_bits(T, A, in, st, on_device);
    else if (path == AMRO_PATH_FAST) run_fast(T, A, in, st, on_device);
    else run_general(T, A, in, st, on_device);
    if (on_device && A->async) { A->total_ms = 0.f; return; }
    AMRO_CUDA(cudaEventRecord(ev.b, st));
    AMRO_CUDA(cudaEventSynchronize(ev.b));
    A->total_ms = ev.ms();
  });
}

// --------------------------------------------------------------------------------------------------------
// One-call equivalents of the reference's Python entry points
// --------------------------------------------------------------------------------------------------------
namespace {

// ---- topology cache of the one-call entry points ----------------------------------------------------
// Optimisation loops call amro.simulate_and_collapse thousands of times on ONE hospital with different parameters
// (tutorials/Test_2024.ipynb:927).  Compiling + uploading the hospital costs far more than a run, so the one-call entry
// points keep the last few compiled hospitals, keyed by a 128-bit hash of the CONTENT of ward_matrix and
// total_patients_per_ward (a caller that edits its arrays in place gets a new topology), n_init and the device.
struct Hash128 { uint64_t a = 0x9E3779B97F4A7C15ull, b = 0xC2B2AE3D27D4EB4Full; };
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
void hash_bytes(Hash128& h, const void* data, size_t n_bytes) {
  const uint64_t* p = static_cast<const uint64_t*>(data);  // inputs are arrays of doubles
  const size_t n = n_bytes / 8;
  uint64_t a0 = h.a, a1 = h.b, a2 = h.a ^ 0x165667B19E3779F9ull, a3 = h.b ^ 0x27D4EB2F165667C5ull;
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {   // four independent multiply-rotate lanes: memory-bound on one core
    a0 = rotl64(a0 ^ p[i], 27) * 0x9FB21C651E98DF25ull;
    a1 = rotl64(a1 ^ p[i + 1], 31) * 0xD6E8FEB86659FD93ull;
    a2 = rotl64(a2 ^ p[i + 2], 33) * 0xA0761D6478BD642Full;
    a3 = rotl64(a3 ^ p[i + 3], 29) * 0xE7037ED1A0B428DBull;
  }
  for (; i < n; ++i) a0 = rotl64(a0 ^ p[i], 27) * 0x9FB21C651E98DF25ull;
  h.a = rotl64(a0 ^ a2, 17) * 0x9E3779B97F4A7C15ull ^ (a1 + n);
  h.b = rotl64(a1 ^ a3, 23) * 0xC2B2AE3D27D4EB4Full ^ (a0 + n);
}
// A large dense matrix is hashed in slabs by a few threads (the hash of a 175 MB ward_matrix on one core costs more than
// the run it saves compiling for); the slab hashes are then chained in order, so the key does not depend on the thread count.
void hash_dense(Hash128& h, const double* m, size_t n) {
  constexpr size_t kSlab = (size_t)1 << 18;   // doubles per slab (2 MB): one core reads ~6 GB/s, so the slabs go to up to 16 threads
  const size_t n_slabs = (n + kSlab - 1) / kSlab;
  if (n_slabs <= 2) { hash_bytes(h, m, sizeof(double) * n); return; }
  std::vector<Hash128> part(n_slabs);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t n_thr = std::min<size_t>(std::min<size_t>(hw, 16), n_slabs);
  std::vector<std::thread> pool;
  for (size_t t = 0; t < n_thr; ++t)
    pool.emplace_back([&, t] {
      for (size_t sl = t; sl < n_slabs; sl += n_thr) {
        const size_t lo = sl * kSlab, hi = std::min(n, lo + kSlab);
        hash_bytes(part[sl], m + lo, sizeof(double) * (hi - lo));
      }
    });
  for (auto& th : pool) th.join();
  hash_bytes(h, part.data(), sizeof(Hash128) * n_slabs);
}
void hash_matrix(Hash128& h, const double* m, int64_t rows, int64_t cols, int64_t ld) {
  const int64_t dims[3] = {rows, cols, 0};
  hash_bytes(h, dims, sizeof dims);
  if (ld == rows) hash_dense(h, m, (size_t)(rows * cols));
  else for (int64_t c = 0; c < cols; ++c) hash_dense(h, m + c * ld, (size_t)rows);
}

// What a call can tell about its tables without reading them: where they are, their shapes, and a sample of their contents
// (64 bytes out of every ~1/256th).  Two calls with the same quick key almost always pass the same tables -- "almost": the
// content hash decides, the quick key only picks the hospital a run may START on while the tables are being hashed.
struct QuickKey {
  const void* wm = nullptr; const void* tp = nullptr;
  int64_t dims[6] = {0, 0, 0, 0, 0, 0};
  Hash128 sample;
  bool operator==(const QuickKey& o) const {
    return wm == o.wm && tp == o.tp && std::memcmp(dims, o.dims, sizeof dims) == 0 && sample.a == o.sample.a && sample.b == o.sample.b;
  }
};
void sample_matrix(Hash128& h, const double* m, int64_t rows, int64_t cols, int64_t ld) {
  const int64_t n = rows * cols;
  if (n <= 0) return;
  if (ld != rows || n <= 256 * 8) { for (int64_t c = 0; c < cols; ++c) hash_bytes(h, m + c * ld, sizeof(double) * (size_t)std::min<int64_t>(rows, 64)); return; }
  const int64_t step = n / 256;
  for (int64_t i = 0; i + 8 <= n; i += step) hash_bytes(h, m + i, 64);
  hash_bytes(h, m + n - 8, 64);
}
QuickKey quick_key(const double* wm, int64_t R, int64_t w_cols, int64_t w_ld, const double* tp, int64_t K, int64_t t_cols, int64_t t_ld) {
  QuickKey q;
  q.wm = wm; q.tp = tp;
  const int64_t d[6] = {R, w_cols, w_ld, K, t_cols, t_ld};
  std::memcpy(q.dims, d, sizeof d);
  sample_matrix(q.sample, wm, R, w_cols, w_ld);
  sample_matrix(q.sample, tp, K, t_cols, t_ld);
  return q;
}

struct CachedTopology {
  Hash128 key;
  QuickKey quick;                         // of the call that last used the entry
  int64_t n_init = 0;
  int device = -1;
  uint64_t stamp = 0;
  size_t bytes = 0;                       // host + device footprint of the compiled hospital (estimate)
  std::shared_ptr<amro_topology> t;       // a run in progress holds a reference: eviction never frees a hospital in use
};
constexpr int kTopoCacheSlots = 4;
std::mutex g_cache_mutex;                 // guards the table below, NOT the runs (those lock their own handle)
CachedTopology g_cache[kTopoCacheSlots];
uint64_t g_cache_clock = 0;
int64_t g_cache_hits = 0, g_cache_misses = 0;

size_t cache_budget_bytes() {             // AMRO_CACHE_BYTES overrides the default of 8 GiB
  if (const char* e = std::getenv("AMRO_CACHE_BYTES")) return (size_t)std::strtoull(e, nullptr, 10);
  return (size_t)8 << 30;
}
size_t topology_bytes(const amro_topology& t) {
  const HostTopology& H = t.H;
  // host: order, pos_of (8 B each), h, w (8 B each), count_day, the four meta arrays (4 B each), push lists; device: 16 B of
  // meta per record + the general path's arrays when uploaded + workspaces that grow with the runs
  size_t b = (size_t)H.R * (8 + 8 + 8 + 8 + 4) + (size_t)H.Rp * 16 + (H.push_src.size() + H.push_dst.size()) * 8;
  b += (size_t)H.Rp * 16 + (size_t)H.R * 8;
  b += t.ws_pre.n * 16 + t.ws_traj.n * 16 + t.ws_pressure.n * 4 + t.ws_counts.n * 4 + (t.src_mask.n + t.half_mask.n) * 4;
  return b;
}
void destroy_topology(amro_topology* t) { amro_topology_destroy(t); }

// Evict least-recently-used hospitals (never the one being inserted / returned) until the footprint fits the budget.
void enforce_cache_budget(const CachedTopology* keep) {
  const size_t budget = cache_budget_bytes();
  for (;;) {
    size_t total = 0;
    CachedTopology* victim = nullptr;
    for (auto& c : g_cache) {
      if (!c.t) continue;
      c.bytes = topology_bytes(*c.t);
      total += c.bytes;
      if (&c != keep && (!victim || c.stamp < victim->stamp)) victim = &c;
    }
    if (total <= budget || !victim) return;
    *victim = CachedTopology{};
  }
}

// Returns a reference to a compiled hospital held by the cache.  The lookup key is the CONTENT of the two tables (a caller
// that edits its arrays in place gets a new topology), n_init and the device.
std::shared_ptr<amro_topology> lookup_or_compile(const Hash128& key, const QuickKey& quick, const double* wm, int64_t R, int64_t w_cols, int64_t w_ld,
                                                 const double* tp, int64_t K, int64_t t_cols, int64_t t_ld, int64_t n_init, int device) {
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (auto& c : g_cache)
      if (c.t && c.key.a == key.a && c.key.b == key.b && c.n_init == n_init && c.device == device) {
        c.stamp = ++g_cache_clock;
        c.quick = quick;
        ++g_cache_hits;
        return c.t;
      }
    ++g_cache_misses;
  }
  amro_topology* raw = nullptr;
  int rc = amro_topology_create(wm, R, w_cols, w_ld, tp, K, t_cols, t_ld, n_init, device, &raw);
  if (rc == AMRO_ECUDA || (rc == AMRO_ERUNTIME && std::strstr(amro_last_error(), "memory"))) {
    // out of device / host memory: let go of every cached hospital nobody is running and try once more
    amro_topology_cache_clear();
    rc = amro_topology_create(wm, R, w_cols, w_ld, tp, K, t_cols, t_ld, n_init, device, &raw);
  }
  if (rc) throw Error(rc, amro_last_error());
  std::shared_ptr<amro_topology> t(raw, destroy_topology);
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  CachedTopology* slot = &g_cache[0];
  for (auto& c : g_cache) {
    if (!c.t) { slot = &c; break; }
    if (c.stamp < slot->stamp) slot = &c;
  }
  *slot = CachedTopology{key, quick, n_init, device, ++g_cache_clock, 0, t};
  enforce_cache_budget(slot);
  return t;
}
Hash128 content_key(const double* wm, int64_t R, int64_t w_cols, int64_t w_ld, const double* tp, int64_t K, int64_t t_cols, int64_t t_ld) {
  Hash128 key;
  Timer th;
  hash_matrix(key, wm, R, w_cols, w_ld);
  hash_matrix(key, tp, K, t_cols, t_ld);
  if (std::getenv("AMRO_FAST_VERBOSE")) std::fprintf(stderr, "[amro] cache key of %lld + %lld rows hashed in %.3f ms\n", (long long)R, (long long)K, th.ms());
  return key;
}

// run(topology) on the cached hospital of these tables.  Hashing 175 MB of tables takes as long as the run it guards, so a call
// whose QUICK key (pointers, shapes, a sample of the contents) matches a cached hospital starts the run on that hospital at once
// and hashes the tables on other threads meanwhile; the content hash has the last word: if it differs (the tables were edited
// in place where the sample did not look) the speculative results are thrown away and the run is repeated on the right hospital.
template <class Run>
void run_on_cached(const double* wm, int64_t R, int64_t w_cols, int64_t w_ld, const double* tp, int64_t K, int64_t t_cols, int64_t t_ld,
                   int64_t n_init, int device, Run run) {
  const QuickKey quick = quick_key(wm, R, w_cols, w_ld, tp, K, t_cols, t_ld);
  std::shared_ptr<amro_topology> cand;
  Hash128 cand_key;
  if (!std::getenv("AMRO_CACHE_NO_SPECULATION")) {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (auto& c : g_cache)
      if (c.t && c.quick == quick && c.n_init == n_init && c.device == device) { cand = c.t; cand_key = c.key; break; }
  }
  if (cand) {
    Hash128 full;
    std::thread hasher([&] { full = content_key(wm, R, w_cols, w_ld, tp, K, t_cols, t_ld); });
    std::exception_ptr err;
    try { run(cand); } catch (...) { err = std::current_exception(); }
    hasher.join();
    if (full.a == cand_key.a && full.b == cand_key.b) {
      {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        for (auto& c : g_cache) if (c.t == cand) c.stamp = ++g_cache_clock;
        ++g_cache_hits;
      }
      if (err) std::rethrow_exception(err);
      return;
    }
    cand.reset();   // the tables changed under the same pointers: whatever the speculative run wrote is overwritten below
    run(lookup_or_compile(full, quick, wm, R, w_cols, w_ld, tp, K, t_cols, t_ld, n_init, device));
    return;
  }
  const Hash128 full = content_key(wm, R, w_cols, w_ld, tp, K, t_cols, t_ld);
  run(lookup_or_compile(full, quick, wm, R, w_cols, w_ld, tp, K, t_cols, t_ld, n_init, device));
}

// amro_simulate on a cached hospital; on a CUDA failure (typically out of memory: the workspaces of other cached hospitals
// stay allocated between calls) the other hospitals are dropped and the run is tried once more.
void simulate_cached(const std::shared_ptr<amro_topology>& topo, amro_sim_args& a) {
  int rc = amro_simulate(topo.get(), &a);
  if (rc == AMRO_ECUDA) {
    cudaGetLastError();
    {
      std::lock_guard<std::mutex> lock(g_cache_mutex);
      for (auto& c : g_cache) if (c.t && c.t != topo) c = CachedTopology{};
    }
    rc = amro_simulate(topo.get(), &a);
  }
  if (rc) throw Error(rc, amro_last_error());
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  for (auto& c : g_cache) if (c.t == topo) { enforce_cache_budget(&c); break; }
}

}  // namespace

extern "C" void amro_topology_cache_clear(void) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  for (auto& c : g_cache) c = CachedTopology{};
}

extern "C" void amro_topology_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries, int64_t* bytes) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  int64_t n = 0, b = 0;
  for (auto& c : g_cache) if (c.t) { ++n; b += (int64_t)topology_bytes(*c.t); }
  if (hits) *hits = g_cache_hits;
  if (misses) *misses = g_cache_misses;
  if (entries) *entries = n;
  if (bytes) *bytes = b;
}

namespace {

// mean(X,1) / stddev(X,1,1) over simulations with Armadillo's operation order
// (armadillo_bits/op_mean_meat.hpp:105-133, op_var_meat.hpp:85-106,175-219, arrayops_meat.hpp:917-936)
// X: the counts as they come back from the device (int32) or a float64 matrix; every value is converted to double first
template <class X>
void row_mean_sd(const X* xs, int64_t n_days, int64_t S, int64_t ld, double* mean, double* sd) {
  // Simulation-major passes over the column-major counts (contiguous in memory), one accumulator set per day: every day
  // sees exactly the operation order of the per-day loops of Armadillo (mean: one accumulator; variance: accumulate's two
  // accumulators for the mean, then op_var's pairwise (q, l) sums).
  auto x = [xs, ld](int64_t d, int64_t s) { return (double)xs[d + ld * s]; };
  const size_t n = (size_t)n_days;
  if (mean) {
    std::vector<double> acc(n, 0.0);
    for (int64_t s = 0; s < S; ++s) for (int64_t d = 0; d < n_days; ++d) acc[d] += x(d, s);
    for (int64_t d = 0; d < n_days; ++d) mean[d] = acc[d] / (double)S;
  }
  if (sd) {
    if (S < 2) { for (int64_t d = 0; d < n_days; ++d) sd[d] = 0.0; return; }
    std::vector<double> a1(n, 0.0), a2(n, 0.0), m(n), q(n, 0.0), l(n, 0.0);
    int64_t i;
    for (i = 0; i + 1 < S; i += 2) for (int64_t d = 0; d < n_days; ++d) { a1[d] += x(d, i); a2[d] += x(d, i + 1); }
    if (i < S) for (int64_t d = 0; d < n_days; ++d) a1[d] += x(d, i);
    for (int64_t d = 0; d < n_days; ++d) m[d] = (a1[d] + a2[d]) / (double)S;
    for (i = 0; i + 1 < S; i += 2)
      for (int64_t d = 0; d < n_days; ++d) {
        const double ti = m[d] - x(d, i), tj = m[d] - x(d, i + 1);
        q[d] += (ti * ti) + (tj * tj);
        l[d] += ti + tj;
      }
    if (i < S) for (int64_t d = 0; d < n_days; ++d) { const double ti = m[d] - x(d, i); q[d] += ti * ti; l[d] += ti; }
    for (int64_t d = 0; d < n_days; ++d) sd[d] = std::sqrt((q[d] - (l[d] * l[d]) / (double)S) / (double)S);
  }
}

}  // namespace

extern "C" int amro_simulate_discrete_model(const double* icp, int64_t i_rows, int64_t i_cols, int64_t i_ld,
                                            const double* idp, int64_t d_rows, int64_t d_cols, int64_t d_ld,
                                            const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                                            const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                                            const double* parameters, int64_t p_rows, int64_t p_cols, int64_t p_ld,
                                            int time_to_detect, int tsh, int tsa, int seed, int device, double* out) {
  return guarded([&] {
    const int64_t S = i_cols, R = n_rows;
    if (!out) fail(AMRO_EINVAL, "out is NULL");
    // argument checks in the order the reference trips over them (abm_wards.cpp:356-375)
    if (t_rows == 0 || t_cols == 0) fail(AMRO_ERUNTIME, "max(): object has no elements");
    if (S < 1 || i_rows < 1 || i_rows > R) fail(AMRO_EINDEX, "Mat::submat(): indices out of bounds or incorrectly used");
    if (d_cols != S) fail(AMRO_ERUNTIME, "relational operator: incompatible matrix dimensions: %lldx%lld and %lldx%lld",
                          (long long)d_rows, (long long)S, (long long)d_rows, (long long)d_cols);
    if (d_rows != i_rows) fail(AMRO_ERUNTIME, "element-wise multiplication: incompatible matrix dimensions: %lldx%lld and %lldx%lld",
                               (long long)d_rows, (long long)S, (long long)i_rows, (long long)S);
    if (p_cols < 6) fail(AMRO_EINDEX, "Mat::operator(): index out of bounds");
    run_on_cached(ward_matrix, R, w_cols, w_ld, tppw, t_rows, t_cols, t_ld, i_rows, device, [&](const std::shared_ptr<amro_topology>& topo) {
      amro_sim_args a{};
      a.n_sims = S; a.sim_offset = 0; a.time_to_detect = time_to_detect;
      a.testing_schedule_hospitalized = tsh; a.testing_schedule_arrivals = tsa; a.seed = seed;
      a.rng_mode = AMRO_RNG_PHILOX; a.path = AMRO_PATH_AUTO; a.keep_trajectory = 1;
      a.parameters = parameters; a.p_rows = p_rows; a.p_ld = p_ld;
      a.icp = icp; a.icp_rs = 1; a.icp_cs = i_ld;
      a.idp = idp; a.idp_rs = 1; a.idp_cs = d_ld;
      a.state_out = out + 6 * R; a.so_ld = R;
      simulate_cached(topo, a);
      // the packed trajectory (2 bytes per record and simulation) is of no use to the next call: give it back
      std::lock_guard<std::mutex> lock(topo->mu);
      topo->ws_traj.release();
    });
    for (int64_t c = 0; c < 6; ++c) std::memcpy(out + c * R, ward_matrix + c * w_ld, sizeof(double) * (size_t)R);  // :375
  });
}

namespace {
// amro.simulate_and_collapse (abm_wards.cpp:566-631); `alloc(n_days_out, ctx)` supplies the [n_days_out x 5] result buffer once
// the number of days is known (from the compiled hospital: no separate pass over ward_matrix)
void collapse_impl(const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                   const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                   double icp, double idp, int64_t initial_patients,
                   double alpha, double beta, double gamma, double rho_hospital,
                   double alpha2, double rho_imported, int64_t n_sims,
                   int64_t time_to_detect, int64_t tsh, int64_t tsa, int seed,
                   int device, int bug_compatible, int64_t* n_days, amro_alloc_fn alloc, void* ctx) {
  if (n_rows < 1 || w_cols < 1) fail(AMRO_ERUNTIME, "max(): object has no elements");
  const int64_t S = n_sims;
  if (S < 1) fail(AMRO_EINDEX, "Mat::submat(): indices out of bounds or incorrectly used");
  int64_t Dd = 0;
  std::vector<double> par((size_t)(S * 6));
  const double v[6] = {alpha, beta, gamma, rho_hospital, alpha2, rho_imported};
  for (int c = 0; c < 6; ++c) for (int64_t s = 0; s < S; ++s) par[s + S * c] = v[c];     // :584-590
  std::vector<int32_t> cc, cd;
  run_on_cached(ward_matrix, n_rows, w_cols, w_ld, tppw, t_rows, t_cols, t_ld, initial_patients, device, [&](const std::shared_ptr<amro_topology>& topo) {
    Dd = topo->H.n_days_out;
    cc.assign((size_t)(Dd * S), 0); cd.assign((size_t)(Dd * S), 0);
    amro_sim_args a{};
    a.n_sims = S; a.time_to_detect = (int)time_to_detect;                                    // uword -> const int& (:606-608)
    a.testing_schedule_hospitalized = (int)tsh; a.testing_schedule_arrivals = (int)tsa; a.seed = seed;
    a.rng_mode = AMRO_RNG_PHILOX; a.path = AMRO_PATH_AUTO; a.keep_trajectory = 0;
    a.parameters = par.data(); a.p_rows = S; a.p_ld = S;
    a.icp = &icp; a.idp = &idp;                                                              // scalar broadcast (:593-597)
    a.counts_colonized = cc.data(); a.counts_detected = cd.data();
    Timer ts;
    simulate_cached(topo, a);
    if (std::getenv("AMRO_FAST_VERBOSE")) std::fprintf(stderr, "[amro] simulate_and_collapse: run %.3f ms (kernel %.3f)\n", ts.ms(), a.kernel_ms);
  });
  double* out = alloc(Dd, ctx);
  if (!out) fail(AMRO_ERUNTIME, "simulate_and_collapse: no result buffer");
  std::vector<double> mc(Dd), md(Dd), sc(Dd), sd(Dd);
  row_mean_sd(cc.data(), Dd, S, Dd, mc.data(), sc.data());                                 // :617-620
  row_mean_sd(cd.data(), Dd, S, Dd, md.data(), sd.data());
  for (int64_t d = 0; d < Dd; ++d) {
    out[d] = (double)d;
    if (bug_compatible) {  // :625-628 write sd over the mean columns and leave 3-4 zero
      out[d + Dd * 1] = sc[d]; out[d + Dd * 2] = sd[d]; out[d + Dd * 3] = 0.0; out[d + Dd * 4] = 0.0;
    } else {
      out[d + Dd * 1] = mc[d]; out[d + Dd * 2] = md[d]; out[d + Dd * 3] = sc[d]; out[d + Dd * 4] = sd[d];
    }
  }
  if (n_days) *n_days = Dd;
}
double* alloc_given(int64_t, void* ctx) { return static_cast<double*>(ctx); }
}  // namespace

extern "C" int amro_simulate_and_collapse(const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                                          const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                                          double icp, double idp, int64_t initial_patients,
                                          double alpha, double beta, double gamma, double rho_hospital,
                                          double alpha2, double rho_imported, int64_t n_sims,
                                          int64_t time_to_detect, int64_t tsh, int64_t tsa, int seed,
                                          int device, int bug_compatible, int64_t* n_days, double* out) {
  return guarded([&] {
    if (n_rows < 1 || w_cols < 1) fail(AMRO_ERUNTIME, "max(): object has no elements");
    if (!out) {  // shape query: max(ward_matrix[:,0]) + 1 (abm_wards.cpp:440)
      double m = ward_matrix[0];
      for (int64_t r = 1; r < n_rows; ++r) m = std::max(m, ward_matrix[r]);
      if (!(m >= 0.0) || m > 1e7) fail(AMRO_EINVAL, "ward_matrix: largest day out of range");
      if (n_days) *n_days = (int64_t)m + 1;
      return;
    }
    collapse_impl(ward_matrix, n_rows, w_cols, w_ld, tppw, t_rows, t_cols, t_ld, icp, idp, initial_patients, alpha, beta, gamma,
                  rho_hospital, alpha2, rho_imported, n_sims, time_to_detect, tsh, tsa, seed, device, bug_compatible, n_days, alloc_given, out);
  });
}

extern "C" int amro_simulate_and_collapse_alloc(const double* ward_matrix, int64_t n_rows, int64_t w_cols, int64_t w_ld,
                                                const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                                                double icp, double idp, int64_t initial_patients,
                                                double alpha, double beta, double gamma, double rho_hospital,
                                                double alpha2, double rho_imported, int64_t n_sims,
                                                int64_t time_to_detect, int64_t tsh, int64_t tsa, int seed,
                                                int device, int bug_compatible, amro_alloc_fn alloc, void* ctx) {
  return guarded([&] {
    if (!alloc) fail(AMRO_EINVAL, "alloc is NULL");
    collapse_impl(ward_matrix, n_rows, w_cols, w_ld, tppw, t_rows, t_cols, t_ld, icp, idp, initial_patients, alpha, beta, gamma,
                  rho_hospital, alpha2, rho_imported, n_sims, time_to_detect, tsh, tsa, seed, device, bug_compatible, nullptr, alloc, ctx);
  });
}

extern "C" int amro_progress_patients(double* ward_matrix, int64_t n_rows, int64_t n_cols, int64_t ld,
                                      const double* tppw, int64_t t_rows, int64_t t_cols, int64_t t_ld,
                                      double total_patients,
                                      const double* parameters, int64_t p_rows, int64_t p_cols, int64_t p_ld,
                                      int64_t n_sims, int time_to_detect, int detect_hosp, int detect_arr,
                                      int rng_mode, int64_t seed, int64_t call_index,
                                      const double* u_col, const double* u_det, int device, double* p_out) {
  return guarded([&] {
    const int64_t S = n_sims, n = n_rows;
    std::vector<int64_t> order, seg_row_start;
    std::vector<double> seg_N;
    compile_day(ward_matrix, n, n_cols, ld, tppw, t_rows, t_cols, t_ld, total_patients, order, seg_row_start, seg_N);
    const int64_t n_seg = (int64_t)seg_N.size();
    if (S == 0 || n_seg == 0) return;
    // bounds the reference's element accessors enforce (abm_wards.cpp:92-112)
    if (6 + 2 * S > n_cols) fail(AMRO_EINDEX, "Mat::col(): index out of bounds");
    if (p_rows < S || p_cols < 6) fail(AMRO_EINDEX, "Mat::operator(): index out of bounds");
    const bool replay = rng_mode == AMRO_RNG_REPLAY;
    if (replay && (!u_col || !u_det)) fail(AMRO_EINVAL, "replay mode needs u_col and u_det");
    use_device(device);
    cudaStream_t st = 0;
    DevBuf<double> M, par, F, uc, ud, dp;
    DevBuf<int64_t> d_order, d_srs;
    DevBuf<uint32_t> d_sop;
    DevBuf<double> d_N;
    upload_matrix(M, ward_matrix, n, 6 + 2 * S, ld, st);
    upload_matrix(par, parameters, S, 6, p_ld, st);
    std::vector<uint32_t> sop((size_t)n);
    for (int64_t s = 0; s < n_seg; ++s) for (int64_t p = seg_row_start[s]; p < seg_row_start[s + 1]; ++p) sop[p] = (uint32_t)s;
    d_order.upload(order.data(), order.size(), st); d_srs.upload(seg_row_start.data(), seg_row_start.size(), st);
    d_sop.upload(sop.data(), sop.size(), st); d_N.upload(seg_N.data(), seg_N.size(), st);
    F.alloc((size_t)(n_seg * S));
    if (replay) { uc.upload(u_col, (size_t)(n * S), st); ud.upload(u_det, (size_t)(n * S), st); }
    if (p_out) { dp.alloc((size_t)(n * S)); AMRO_CUDA(cudaMemsetAsync(dp.p, 0xFF, sizeof(double) * n * S, st)); }
    GenArgs g{};
    g.C = M.p + n * 6; g.D = M.p + n * (6 + S); g.ld = n;
    g.h = M.p + n * 3; g.w = M.p + n * 4; g.hw_stride = 1;
    g.order = d_order.p; g.seg_of_pos = d_sop.p; g.seg_row_start = d_srs.p; g.seg_N = d_N.p;
    g.params = par.p; g.p_ld = S; g.S = S; g.F = F.p; g.ttd = time_to_detect;
    g.replay = replay; g.u_col = uc.p; g.u_det = ud.p; g.u_ld = n;
    g.seed = (uint64_t)seed + 0x9E3779B97F4A7C15ull * (uint64_t)call_index; g.sim_offset = 0;
    g.p_out = p_out ? dp.p : nullptr; g.p_out_ld = n;
    gen_pressure(st, g, 0, n_seg);
    gen_update(st, g, 0, n, 0, detect_hosp, detect_arr);
    AMRO_CUDA(cudaGetLastError());
    // the reference mutates the caller's matrix in place (C and D columns)
    AMRO_CUDA(cudaMemcpy2DAsync(ward_matrix + 6 * ld, ld * sizeof(double), M.p + n * 6, n * sizeof(double), n * sizeof(double),
                                2 * S, cudaMemcpyDeviceToHost, st));
    if (p_out) AMRO_CUDA(cudaMemcpyAsync(p_out, dp.p, sizeof(double) * n * S, cudaMemcpyDeviceToHost, st));
    AMRO_CUDA(cudaStreamSynchronize(st));
  });
}

extern "C" int amro_total_positive(const double* model, int64_t n_rows, int64_t n_cols, int64_t ld,
                                   int64_t n_sims, int mode, int device,
                                   int64_t* n_days, int64_t* n_out_cols, double* out) {
  return guarded([&] {
    if (n_rows < 1 || n_cols < 1) fail(AMRO_ERUNTIME, "max(): object has no elements");                 // :440
    double m = model[0];
    for (int64_t r = 1; r < n_rows; ++r) m = std::max(m, model[r]);
    if (!(m >= 0.0) || m > 1e7) fail(AMRO_EINVAL, "model_colonized: largest day out of range");
    const int64_t Dd = (int64_t)m + 1;
    // :480-481 (mode 1: columns 6 .. 6+n_sims-1)   :503-504 (mode 0: columns 6+n_sims .. n_cols-1)
    const int64_t c0 = mode == 1 ? 6 : 6 + n_sims, c1 = mode == 1 ? 6 + n_sims - 1 : n_cols - 1;
    if (c1 >= n_cols || c0 > c1 + 0 || c0 < 0) fail(AMRO_EINDEX, "Mat::cols(): indices out of bounds or incorrectly used");
    const int64_t nc = c1 - c0 + 1;
    if (n_days) *n_days = Dd;
    if (n_out_cols) *n_out_cols = nc;
    if (!out) return;
    use_device(device);
    cudaStream_t st = 0;
    std::vector<int32_t> cday((size_t)n_rows);
    for (int64_t r = 0; r < n_rows; ++r) {
      const double v = model[r];
      const int64_t i = (v >= 0.0 && v <= (double)(Dd - 1)) ? (int64_t)v : -1;
      cday[r] = (i >= 0 && (double)i == v) ? (int32_t)i : -1;
    }
    DevBuf<double> plane, dout;
    DevBuf<int32_t> d_day, cnt;
    upload_matrix(plane, model + c0 * ld, n_rows, nc, ld, st);
    d_day.upload(cday.data(), cday.size(), st);
    cnt.alloc((size_t)(Dd * nc));
    dout.alloc((size_t)(Dd * nc));
    AMRO_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * Dd * nc, st));
    gen_count(st, plane.p, n_rows, n_rows, nc, d_day.p, mode == 1 ? 1 : 0, cnt.p, Dd);
    counts_to_double(st, cnt.p, dout.p, Dd * nc);
    AMRO_CUDA(cudaGetLastError());
    AMRO_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * Dd * nc, cudaMemcpyDeviceToHost, st));
    AMRO_CUDA(cudaStreamSynchronize(st));
  });
}

extern "C" int amro_summary_of_total_positive_device(const double* positive, int64_t n_days, int64_t n_sims, int64_t ld,
                                                     const double* quantiles, int64_t n_q, int device, double* out) {
  return guarded([&] {
    if (n_days < 0 || n_sims < 0 || n_q < 0) fail(AMRO_EINVAL, "negative size");
    if (n_days == 0) return;
    if (n_sims == 0) fail(AMRO_ERUNTIME, "copy into submatrix: incompatible matrix dimensions");
    use_device(device);
    cudaStream_t st = 0;
    DevBuf<double> x, q, o;
    upload_matrix(x, positive, n_days, n_sims, ld, st);
    if (n_q) q.upload(quantiles, (size_t)n_q, st);
    o.alloc((size_t)(n_days * (3 + n_q)));
    device_summary(st, x.p, n_days, n_sims, n_days, q.p, n_q, o.p, n_days);
    AMRO_CUDA(cudaMemcpyAsync(out, o.p, sizeof(double) * n_days * (3 + n_q), cudaMemcpyDeviceToHost, st));
    AMRO_CUDA(cudaStreamSynchronize(st));
  });
}

extern "C" int amro_summary_of_total_positive(const double* positive, int64_t n_days, int64_t n_sims, int64_t ld,
                                              const double* quantiles, int64_t n_q, double* out) {
  return guarded([&] {
    if (n_days < 0 || n_sims < 0 || n_q < 0) fail(AMRO_EINVAL, "negative size");
    if (n_days == 0) return;
    if (n_sims == 0) fail(AMRO_ERUNTIME, "copy into submatrix: incompatible matrix dimensions");
    std::vector<double> mean(n_days), sd(n_days), y(n_sims);
    row_mean_sd(positive, n_days, n_sims, ld, mean.data(), sd.data());                       // :534-535
    // :536-537 arma::quantile = Hyndman & Fan definition 5 (armadillo_bits/glue_quantile_meat.hpp:22-80)
    const double N = (double)n_sims, pmin = (1.0 - 0.5) / N, pmax = (N - 0.5) / N;
    for (int64_t d = 0; d < n_days; ++d) {
      out[d] = (double)d;
      out[d + n_days] = mean[d];
      out[d + 2 * n_days] = sd[d];
      for (int64_t s = 0; s < n_sims; ++s) y[s] = positive[d + ld * s];
      std::sort(y.begin(), y.end());
      for (int64_t i = 0; i < n_q; ++i) {
        const double p = quantiles[i];
        double v;
        if (p < pmin) v = (p < 0.0) ? -INFINITY : y.front();
        else if (p > pmax) v = (p > 1.0) ? INFINITY : y.back();
        else {
          const int64_t k = (int64_t)std::floor(N * p + 0.5);
          const double pk = ((double)k - 0.5) / N;
          const double wgt = (p - pk) * N;
          v = ((1.0 - wgt) * y[k - 1]) + (wgt * y[std::min<int64_t>(k, n_sims - 1)]);   // k == n_sims only with wgt == 0
        }
        out[d + (3 + i) * n_days] = v;
      }
    }
  });
}
