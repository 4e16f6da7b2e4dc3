// summary_of_total_positive on the device: per day, over the simulations -- mean, population standard
// deviation and arma::quantile (src/amro/abm_wards.cpp:526-540).  What a 64k-particle sweep needs back is [days x (3 + nq)]
// numbers, not the 94 MB of per-(day, particle) counts; and the host routine sorts 65,536 doubles per day on one core.
//
//   mean / sd    one thread per day walks the day's values in the order Armadillo does (op_mean: one accumulator; op_var:
//                accumulate's two accumulators for the mean, then pairwise (q, l) sums) -- the same IEEE operations in the
//                same order as the host routine, so the results are bit-identical for any input, not just integer counts;
//                a day's values are a strided row of the column-major matrix, adjacent threads read adjacent days: coalesced
//   quantiles    the matrix is transposed into day-major segments, every segment sorted with CUB's segmented radix sort
//                (double keys), and one thread per (day, quantile) interpolates: Hyndman & Fan definition 5
//                (armadillo_bits/glue_quantile_meat.hpp:22-80)
#include <cub/device/device_segmented_radix_sort.cuh>

#include "common.hpp"
#include "kernels.hpp"

namespace amro {
namespace {

__global__ void k_summary_moments(const double* __restrict__ X, int64_t n_days, int64_t S, int64_t ld, double* __restrict__ out, int64_t ld_out) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_days) return;
  const double* x = X + d;
  const double n = (double)S;
  double acc = 0.0;
  for (int64_t s = 0; s < S; ++s) acc = __dadd_rn(acc, x[ld * s]);
  out[d] = (double)d;
  out[d + ld_out] = __ddiv_rn(acc, n);                                            // :534 mean(X, 1)
  double var = 0.0;
  if (S >= 2) {                                                                    // :535 stddev(X, 1, 1): population variance
    double a1 = 0.0, a2 = 0.0;
    int64_t i;
    for (i = 0; i + 1 < S; i += 2) { a1 = __dadd_rn(a1, x[ld * i]); a2 = __dadd_rn(a2, x[ld * (i + 1)]); }
    if (i < S) a1 = __dadd_rn(a1, x[ld * i]);
    const double m = __ddiv_rn(__dadd_rn(a1, a2), n);
    double q = 0.0, l = 0.0;
    for (i = 0; i + 1 < S; i += 2) {
      const double ti = __dsub_rn(m, x[ld * i]), tj = __dsub_rn(m, x[ld * (i + 1)]);
      q = __dadd_rn(q, __dadd_rn(__dmul_rn(ti, ti), __dmul_rn(tj, tj)));
      l = __dadd_rn(l, __dadd_rn(ti, tj));
    }
    if (i < S) { const double ti = __dsub_rn(m, x[ld * i]); q = __dadd_rn(q, __dmul_rn(ti, ti)); l = __dadd_rn(l, ti); }
    var = __ddiv_rn(__dsub_rn(q, __ddiv_rn(__dmul_rn(l, l), n)), n);
  }
  out[d + 2 * ld_out] = sqrt(var);
}

// [n_days x S] column-major -> day-major segments, through a shared-memory tile (both sides coalesced)
__global__ void k_summary_transpose(const double* __restrict__ X, int64_t n_days, int64_t S, int64_t ld, double* __restrict__ Y) {
  __shared__ double tile[32][33];
  const int64_t d0 = (int64_t)blockIdx.x * 32, s0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t d = d0 + threadIdx.x, s = s0 + j;
    if (d < n_days && s < S) tile[j][threadIdx.x] = X[d + ld * s];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t d = d0 + j, s = s0 + threadIdx.x;
    if (d < n_days && s < S) Y[d * S + s] = tile[threadIdx.x][j];
  }
}

__global__ void k_summary_offsets(int64_t n_days, int64_t S, int64_t* off) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d <= n_days) off[d] = d * S;
}

__global__ void k_summary_quantiles(const double* __restrict__ sorted, int64_t n_days, int64_t S, const double* __restrict__ qs, int64_t nq,
                                    double* __restrict__ out, int64_t ld_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_days * nq) return;
  const int64_t d = i % n_days, j = i / n_days;
  const double* y = sorted + d * S;
  const double N = (double)S, p = qs[j];
  const double pmin = __ddiv_rn(0.5, N), pmax = __ddiv_rn(__dsub_rn(N, 0.5), N);
  double v;
  if (p < pmin) v = (p < 0.0) ? -INFINITY : y[0];
  else if (p > pmax) v = (p > 1.0) ? INFINITY : y[S - 1];
  else {
    const int64_t k = (int64_t)floor(__dadd_rn(__dmul_rn(N, p), 0.5));
    const double pk = __ddiv_rn(__dsub_rn((double)k, 0.5), N);
    const double w = __dmul_rn(__dsub_rn(p, pk), N);
    v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, w), y[k - 1]), __dmul_rn(w, y[k < S ? k : S - 1]));   // k == S only with w == 0
  }
  out[d + (3 + j) * ld_out] = v;
}

}  // namespace

// X [n_days x S] (device, column-major, ld) -> out [n_days x (3 + nq)] (device, column-major, ld_out): day, mean, sd, quantiles
void device_summary(cudaStream_t st, const double* X, int64_t n_days, int64_t S, int64_t ld, const double* q_dev, int64_t nq, double* out,
                    int64_t ld_out) {
  if (n_days <= 0) return;
  k_summary_moments<<<(unsigned)((n_days + 63) / 64), 64, 0, st>>>(X, n_days, S, ld, out, ld_out);
  if (nq > 0) {
    DevBuf<double> Y, Z;
    DevBuf<int64_t> off;
    DevBuf<unsigned char> tmp;
    Y.alloc((size_t)(n_days * S)); Z.alloc((size_t)(n_days * S)); off.alloc((size_t)(n_days + 1));
    k_summary_transpose<<<dim3((unsigned)((n_days + 31) / 32), (unsigned)((S + 31) / 32)), dim3(32, 8), 0, st>>>(X, n_days, S, ld, Y.p);
    k_summary_offsets<<<(unsigned)((n_days + 256) / 256), 256, 0, st>>>(n_days, S, off.p);
    if (n_days * S >= ((int64_t)1 << 31)) fail(AMRO_EINVAL, "summary: more than 2^31 values");
    size_t bytes = 0;
    AMRO_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, bytes, Y.p, Z.p, (int)(n_days * S), (int)n_days, off.p, off.p + 1, 0, 64, st));
    tmp.alloc(bytes);
    AMRO_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(tmp.p, bytes, Y.p, Z.p, (int)(n_days * S), (int)n_days, off.p, off.p + 1, 0, 64, st));
    k_summary_quantiles<<<(unsigned)((n_days * nq + 255) / 256), 256, 0, st>>>(Z.p, n_days, S, q_dev, nq, out, ld_out);
    AMRO_CUDA(cudaGetLastError());
    AMRO_CUDA(cudaStreamSynchronize(st));   // the scratch buffers die here
  }
  AMRO_CUDA(cudaGetLastError());
}

}  // namespace amro
