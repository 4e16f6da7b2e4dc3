// Shared host-side helpers: status + thread-local error message, CUDA error checks, RAII device buffers.
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/amro_b200.h"

namespace amro {

// Exception carrying one of the amro_status classes; thrown by host code, caught at the C ABI.
struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

[[noreturn]] inline void fail(int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(status, buf);
}

void set_last_error(const char* msg);

// Runs f(), translating exceptions into a status + thread-local message.
template <class F>
int guarded(F&& f) {
  try {
    f();
    return AMRO_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.status;
  } catch (const std::bad_alloc&) {
    set_last_error("out of host memory");
    return AMRO_ERUNTIME;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return AMRO_ERUNTIME;
  }
}

}  // namespace amro

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define AMRO_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      ::amro::fail(AMRO_ECUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                   cudaGetErrorString(_e));                                                          \
  } while (0)

namespace amro {
// Device allocation owned by a handle or a call; freed on destruction.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess)
      fail(AMRO_ECUDA, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    n = count;
  }
  void upload(const T* src, size_t count, cudaStream_t st = 0) {
    alloc(count);
    if (count) AMRO_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
  }
};
}  // namespace amro
#endif
