// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// Replacement RNG for Armadillo's official hook  -DARMA_RNG_ALT=<this header>
// (/root/reference/armadillo_12_6_1/armadillo:143-145; the members required are the ones
// used in armadillo_bits/arma_rng.hpp:78-92,109-111,217-219,300-302,323-325).
//
// Every uniform the UNMODIFIED reference sources ask for -- the scalar draws of
// src/amro/abm_wards.cpp:116,134,146 and the fill::randu matrices of :363,:369 -- is popped,
// in call order, from a flat buffer the test harness installs through the two C entry
// points in replay_hooks.cpp.  Nothing here is copied from the reference.
#pragma once
#include <cstddef>
#include <cmath>
#include <stdexcept>

extern "C" {
extern const double* amro_replay_buf;
extern std::size_t   amro_replay_len;
extern std::size_t   amro_replay_cur;
}

class arma_rng_alt
  {
  public:
  typedef unsigned int seed_type;

  inline static void set_seed(const seed_type) {}

  inline static double randu_val()
    {
    if(amro_replay_cur >= amro_replay_len) { throw std::runtime_error("replay buffer exhausted"); }
    return amro_replay_buf[amro_replay_cur++];
    }

  inline static int    randi_val()     { return int(randu_val() * 2147483647.0); }
  inline static int    randi_max_val() { return 2147483647; }
  inline static double randn_val()     { return 0.0; }

  template<typename eT>
  inline static void randn_dual_val(eT& out1, eT& out2) { out1 = eT(0); out2 = eT(0); }

  template<typename eT>
  inline static void randi_fill(eT* mem, const unsigned long long N, const int a, const int b)
    {
    for(unsigned long long i = 0; i < N; ++i) { mem[i] = eT(a + int(randu_val() * double(b - a + 1))); }
    }
  };
