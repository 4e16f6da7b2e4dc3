// The topology handle of the C ABI (include/amro_b200.h: amro_topology): a compiled hospital on one GPU, its device arrays
// and the workspaces its runs reuse.  Shared by api.cu (compile from the caller's tables) and generate.cu (synthetic
// hospitals generated on the device).
#pragma once
#include <memory>
#include <mutex>
#include <vector>

#include "common.hpp"
#include "kernels.hpp"
#include "topology.hpp"

namespace amro {
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  float ms() const { float t = 0.f; cudaEventElapsedTime(&t, a, b); return t; }
};
}  // namespace amro

struct amro_topology {
  amro::HostTopology H;
  bool generated = false;   // made by amro_topology_generate: device arrays only (no host tables, fused kernels only)
  std::mutex mu;   // amro_simulate calls on one handle are serialised: they share the handle's workspaces
  // CUDA events around the fused kernel of the last kKernelEvents runs (asynchronous runs are timed after the fact:
  // amro_topology_kernel_times)
  static constexpr int kKernelEvents = 64;
  std::unique_ptr<amro::EventPair[]> kernel_events;
  uint64_t n_fused_runs = 0;
  amro::EventPair& next_kernel_events() {
    if (!kernel_events) kernel_events.reset(new amro::EventPair[kKernelEvents]);
    return kernel_events[n_fused_runs++ % kKernelEvents];
  }
  int device = -1;
  double compile_ms = 0.0;
  int coresident[2] = {0, 0};  // cooperative-launch capacity of k_fast<false/true>
  // shared / fast path
  amro::DevBuf<int64_t> day_start, day_seg_start, seg_row_start, pos_of;
  amro::DevBuf<double> seg_N;
  amro::DevBuf<uint4> meta;   // packed per-record words of the fast path (kernels.hpp FastArgs::meta)
  amro::DevBuf<uint32_t> init_seg, init_slot, seg_main_next;
  // bit-sliced kernel (built on first use from the device arrays above: the same for compiled and generated hospitals)
  bool pull_ready = false;
  amro::DevBuf<uint32_t> src_mask, half_mask;
  amro::DevBuf<int64_t> day_chunk_start;
  void ensure_pull() {
    if (pull_ready) return;
    const int64_t n_days = H.total_days + 1;
    std::vector<int64_t> ds((size_t)(n_days + 1)), dc((size_t)(n_days + 1), 0);
    AMRO_CUDA(cudaMemcpy(ds.data(), day_start.p, sizeof(int64_t) * (n_days + 1), cudaMemcpyDeviceToHost));
    for (int64_t d = 0; d < n_days; ++d) dc[d + 1] = dc[d] + (ds[d + 1] - ds[d] + 31) / 32;
    day_chunk_start.upload(dc.data(), dc.size());
    src_mask.alloc((size_t)std::max<int64_t>(dc[n_days], 1));
    if (H.has_half) half_mask.alloc((size_t)std::max<int64_t>(dc[n_days], 1));
    amro::bits_build_src_mask(0, meta.p, day_start.p, day_chunk_start.p, n_days, H.max_rows_per_day, src_mask.p, H.has_half ? half_mask.p : nullptr);
    AMRO_CUDA(cudaStreamSynchronize(0));
    pull_ready = true;
  }
  // general path (uploaded on first use)
  bool general_ready = false;
  amro::DevBuf<int64_t> order, push_src, push_dst;
  amro::DevBuf<uint32_t> seg_of_pos;
  amro::DevBuf<double> h, w;
  amro::DevBuf<int32_t> count_day;
  // workspace reused across runs (parameter-inference loops call amro_simulate thousands of times)
  amro::DevBuf<uint4> ws_pre, ws_traj;
  amro::DevBuf<uint32_t> ws_pressure;
  amro::DevBuf<int32_t> ws_counts;
  amro::DevBuf<double> ws_params;
  amro::DevBuf<unsigned int> ws_barrier;
  // pinned host staging of the per-day counts for host-pointer callers (pageable destinations copy at a fraction of PCIe speed)
  int32_t* pinned_counts = nullptr;
  size_t pinned_counts_n = 0;
  int32_t* pinned(size_t n) {
    if (n > pinned_counts_n) {
      if (pinned_counts) cudaFreeHost(pinned_counts);
      pinned_counts = nullptr; pinned_counts_n = 0;
      if (cudaMallocHost(&pinned_counts, n * sizeof(int32_t)) != cudaSuccess) { cudaGetLastError(); pinned_counts = nullptr; return nullptr; }
      pinned_counts_n = n;
    }
    return pinned_counts;
  }
  ~amro_topology() { if (pinned_counts) cudaFreeHost(pinned_counts); }

  void ensure_general() {
    if (general_ready) return;
    order.upload(H.order.data(), H.order.size());
    push_src.upload(H.push_src.data(), H.push_src.size());
    push_dst.upload(H.push_dst.data(), H.push_dst.size());
    std::vector<uint32_t> sop((size_t)H.Rp);
    for (int64_t s = 0; s < H.n_seg; ++s)
      for (int64_t p = H.seg_row_start[s]; p < H.seg_row_start[s + 1]; ++p) sop[p] = (uint32_t)s;
    seg_of_pos.upload(sop.data(), sop.size());
    h.upload(H.h.data(), H.h.size());
    w.upload(H.w.data(), H.w.size());
    count_day.upload(H.count_day.data(), H.count_day.size());
    AMRO_CUDA(cudaStreamSynchronize(0));
    general_ready = true;
  }
};

