// The drop-in Python module `amro`: same seven functions, argument names and defaults as the reference's
// pybind11 module (/root/reference/src/amro/amro.cpp:18-318), bound to the C ABI of libamro_b200.so
// (include/amro_b200.h) instead of the Armadillo implementation.  No carma: arguments are taken as plain
// numpy arrays (any dtype/layout is converted to float64 Fortran order), results are new float64
// Fortran-ordered arrays that own their memory.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/amro_b200.h"

namespace py = pybind11;
using namespace pybind11::literals;
using farray = py::array_t<double, py::array::f_style | py::array::forcecast>;

namespace {

int g_device = -1;
std::atomic<long long> g_step_calls{0};
long long g_seed = 1;  // the reference's single-step functions continue the global rand() stream (default seed 1)

int device() {
  if (g_device < 0) {
    const char* e = std::getenv("AMRO_DEVICE");
    g_device = e ? std::atoi(e) : 0;
  }
  return g_device;
}

// error classes of the C ABI -> what the reference raises through pybind11
void check(int rc) {
  if (rc == AMRO_OK) return;
  const std::string msg = amro_last_error();
  if (rc == AMRO_EINDEX) throw py::index_error(msg);
  if (rc == AMRO_EINVAL) throw py::value_error(msg);
  throw std::runtime_error(msg);
}

struct Mat {  // 1-D -> column, like carma (numpytoarma.h:74-76); anything else than 1-D/2-D is refused
  farray a;
  int64_t rows, cols, ld;
  explicit Mat(farray arr) : a(std::move(arr)) {
    if (a.ndim() == 1) { rows = a.shape(0); cols = 1; }
    else if (a.ndim() == 2) { rows = a.shape(0); cols = a.shape(1); }
    else throw std::runtime_error("CARMA: Number of dimensions must be 1 <= ndim <= 2");
    ld = rows > 0 ? rows : 1;
  }
  const double* data() const { return a.data(); }
  double* mutable_data() { return a.mutable_data(); }
};

py::array_t<double> fmat(int64_t rows, int64_t cols) {
  return py::array_t<double>({(py::ssize_t)rows, (py::ssize_t)cols},
                             {(py::ssize_t)sizeof(double), (py::ssize_t)(sizeof(double) * (rows > 0 ? rows : 1))});
}

py::array_t<double> simulate_discrete_model(farray icp_, farray idp_, farray wm_, farray tp_, farray par_,
                                            int time_to_detect, int tsh, int tsa, int seed) {
  Mat icp(icp_), idp(idp_), wm(wm_), tp(tp_), par(par_);
  auto out = fmat(wm.rows, wm.cols + 2 * icp.cols);
  g_seed = seed;
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = amro_simulate_discrete_model(icp.data(), icp.rows, icp.cols, icp.ld, idp.data(), idp.rows, idp.cols, idp.ld,
                                      wm.data(), wm.rows, wm.cols, wm.ld, tp.data(), tp.rows, tp.cols, tp.ld,
                                      par.data(), par.rows, par.cols, par.ld, time_to_detect, tsh, tsa, seed, device(),
                                      out.mutable_data());
  }
  check(rc);
  return out;
}

py::array_t<double> progress(farray wm_, const double* tp, int64_t t_rows, int64_t t_cols, int64_t t_ld, double N,
                             farray par_, int64_t n_sims, int ttd, bool dh, bool da) {
  Mat wm(wm_), par(par_);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = amro_progress_patients(wm.mutable_data(), wm.rows, wm.cols, wm.ld, tp, t_rows, t_cols, t_ld, N, par.data(),
                                par.rows, par.cols, par.ld, n_sims, ttd, dh, da, AMRO_RNG_PHILOX, g_seed,
                                g_step_calls.fetch_add(1), nullptr, nullptr, device(), nullptr);
  }
  check(rc);
  // the reference returns a copy (abm_wards.cpp:158,256) besides mutating a borrowed input in place
  auto out = fmat(wm.rows, wm.cols);
  std::memcpy(out.mutable_data(), wm.data(), sizeof(double) * (size_t)(wm.rows * wm.cols));
  return out;
}

py::array_t<double> progress_ward(farray wm, double total_patients, farray par, int64_t n_sims, int ttd, bool dh, bool da) {
  return progress(std::move(wm), nullptr, 0, 0, 0, total_patients, std::move(par), n_sims, ttd, dh, da);
}

py::array_t<double> progress_day(farray wm, farray tp_, farray par, int64_t n_sims, int ttd, bool dh, bool da) {
  Mat tp(tp_);
  return progress(std::move(wm), tp.data(), tp.rows, tp.cols, tp.ld, 0.0, std::move(par), n_sims, ttd, dh, da);
}

py::array_t<double> total_positive(farray m_, int64_t n_sims, int mode) {
  Mat m(m_);
  int64_t days = 0, nc = 0;
  check(amro_total_positive(m.data(), m.rows, m.cols, m.ld, n_sims, mode, device(), &days, &nc, nullptr));
  auto out = fmat(days, nc);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = amro_total_positive(m.data(), m.rows, m.cols, m.ld, n_sims, mode, device(), &days, &nc, out.mutable_data());
  }
  check(rc);
  return out;
}

py::array_t<double> summary_of_total_positive(farray pos_, farray q_) {
  Mat pos(pos_), q(q_);
  const int64_t nq = q.rows * q.cols;
  auto out = fmat(pos.rows, 3 + nq);
  // from a few hundred thousand values on, the per-day sort across the simulations runs on the device (bit-identical results)
  if (pos.rows * pos.cols >= (int64_t)1 << 18 && amro_device_count() > 0)
    check(amro_summary_of_total_positive_device(pos.data(), pos.rows, pos.cols, pos.ld, q.data(), nq, device(), out.mutable_data()));
  else
    check(amro_summary_of_total_positive(pos.data(), pos.rows, pos.cols, pos.ld, q.data(), nq, out.mutable_data()));
  return out;
}

py::array_t<double> simulate_and_collapse(farray wm_, farray tp_, double icp, double idp, int64_t initial_patients,
                                          double alpha, double beta, double gamma, double rho_hospital, double alpha2,
                                          double rho_imported, int64_t n_sims, int64_t ttd, int64_t tsh, int64_t tsa,
                                          int seed, bool bug_compatible) {
  Mat wm(wm_), tp(tp_);
  g_seed = seed;
  py::array_t<double> out;
  // the result array is created (with the GIL re-acquired) once the library knows the number of days
  auto alloc = [](int64_t days, void* ctx) -> double* {
    py::gil_scoped_acquire gil;
    auto* o = static_cast<py::array_t<double>*>(ctx);
    *o = fmat(days, 5);
    return o->mutable_data();
  };
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = amro_simulate_and_collapse_alloc(wm.data(), wm.rows, wm.cols, wm.ld, tp.data(), tp.rows, tp.cols, tp.ld, icp, idp,
                                          initial_patients, alpha, beta, gamma, rho_hospital, alpha2, rho_imported, n_sims, ttd,
                                          tsh, tsa, seed, device(), bug_compatible, alloc, &out);
  }
  check(rc);
  return out;
}

}  // namespace

PYBIND11_MODULE(amro, m) {
  m.doc() = "Stochastic model of antimicrobial resistant organisms in a healthcare facility -- B200-native "
            "implementation of the ABM_AMRO ward-level colonisation path (drop-in for the reference module `amro`).";

  // Defaults are the reference's literal ones (amro.cpp:83-91,149-155,216-222,236-237,251-252,270-271,314,318):
  // its tests rely on time_to_detect.. seed = 6,7,8,9 and n_sims = 2.
  m.def("simulate_discrete_model", &simulate_discrete_model,
        "Simulate every patient-day record of every simulation; returns the [R x (6+2S)] state matrix.",
        "initial_colonized_probability"_a, "initial_detected_probability"_a, "ward_matrix"_a,
        "total_patients_per_ward"_a, "parameters"_a, "time_to_detect"_a = 6, "testing_schedule_hospitalized"_a = 7,
        "testing_schedule_arrivals"_a = 8, "seed"_a = 9);
  m.def("progress_patients_probability_ward_1_timestep", &progress_ward,
        "Progress the patients of ONE ward one day (updates a float64 Fortran-ordered ward_matrix in place and returns a copy).",
        "ward_matrix"_a, "total_patients"_a = 2.0, "parameters"_a, "n_sims"_a = 4, "time_to_detect"_a = 5,
        "detect_today_hospitalized"_a = true, "detect_today_arrival"_a = true);
  m.def("progress_patients_1_timestep", &progress_day,
        "Progress the patients of all wards one day (updates a float64 Fortran-ordered ward_matrix in place and returns a copy).",
        "ward_matrix"_a, "total_patients_per_ward"_a, "parameters"_a, "n_sims"_a = 4, "time_to_detect"_a = 5,
        "detect_today_hospitalized"_a = true, "detect_today_arrival"_a = true);
  m.def("total_positive_colonized", [](farray m_, int64_t n_sims) { return total_positive(std::move(m_), n_sims, 1); },
        "Daily number of colonised records per simulation.", "model_colonized"_a, "n_sims"_a = 2);
  m.def("total_positive_detected", [](farray m_, int64_t n_sims) { return total_positive(std::move(m_), n_sims, 0); },
        "Daily number of detected records per simulation.", "model_colonized"_a, "n_sims"_a = 2);
  m.def("summary_of_total_positive", &summary_of_total_positive,
        "Day, mean, population sd and quantiles across simulations.", "model_colonized"_a, "quantiles"_a);
  m.def("simulate_and_collapse", &simulate_and_collapse,
        "Simulate n_sims identical-parameter runs and collapse them to per-day statistics.  bug_compatible=True (default) "
        "returns the reference's actual columns (day, sd colonised, sd detected, 0, 0); False returns the documented "
        "(day, mean colonised, mean detected, sd colonised, sd detected).",
        "ward_matrix"_a, "total_patients_per_ward"_a, "initial_colonized_probability"_a,
        "initial_detected_probability"_a, "initial_patients"_a, "alpha"_a, "beta"_a, "gamma"_a, "rho_hospital"_a,
        "alpha2"_a, "rho_imported"_a, "n_sims"_a = 100, "time_to_detect"_a, "testing_schedule_hospitalized"_a,
        "testing_schedule_arrivals"_a, "seed"_a = 2384235, py::kw_only(), "bug_compatible"_a = true);

  m.def("set_device", [](int d) { g_device = d; }, "CUDA device used by this module (default: $AMRO_DEVICE or 0).");
  m.def("device_count", &amro_device_count);
  m.def("clear_cache", &amro_topology_cache_clear, "Free the compiled hospitals the one-call functions keep on the device.");
  m.attr("__version__") = amro_version();
}
