// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
// Storage + ctypes-callable entry points for replay_rng_alt.hpp.
#include <cstddef>
extern "C" {
const double* amro_replay_buf = nullptr;
std::size_t   amro_replay_len = 0;
std::size_t   amro_replay_cur = 0;

void amro_replay_set(const double* buf, std::size_t n) { amro_replay_buf = buf; amro_replay_len = n; amro_replay_cur = 0; }
std::size_t amro_replay_pos(void) { return amro_replay_cur; }
}
