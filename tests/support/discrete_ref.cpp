// discrete_ref.cpp -- TEST INFRASTRUCTURE: the standard library's own std::discrete_distribution, constructed anew for every draw the way
// the reference does (matchBase.hpp:120-140), for tests/test_s4pcs_plan.py to compare the planner's table-free replica with.
#include <cstdint>
#include <random>
#include <vector>

extern "C" void hop_ref_draw_discrete(const float *w, int n, uint32_t seed, int draws, int32_t *out) {
  std::mt19937 engine(seed);
  for (int k = 0; k < draws; ++k) {
    std::discrete_distribution<> d(w, w + n);
    out[k] = d(engine);
  }
}
