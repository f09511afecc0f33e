// Host-side check of dl::round_div (d-liom_b200/csrc/dl_math.cuh, host + device code) against lround(x / res), the reference's
// cell index (voxel_filter.cc:126-131, hybrid_grid.h:430-435): random coordinates plus values within +-4 ulp of every k + 0.5
// rounding boundary, for the resolutions the pipeline uses. Prints "n=<cases> bad=<mismatches>"; exit status 1 on a mismatch.
#include <cstdio>
#include <random>

#include "../../d-liom_b200/csrc/dl_math.cuh"

int main() {
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> ux(-200.f, 200.f);
  const float ress[] = {0.075f, 0.15f, 0.1f, 0.05f, 0.45f, 2.f, 1.f, 0.6875f, 3.3f, 1e-3f, 7.5f};
  long bad = 0, n = 0;
  for (float res : ress) {
    const dl::CellDivider d = dl::make_divider(res);
    for (int i = 0; i < 200000; ++i) {
      float x = ux(rng);
      if (i % 3 == 0) {  // adversarial: near a boundary (k + 0.5) * res
        const int k = (int)(rng() % 4000) - 2000;
        x = (k + 0.5f) * res;
        const int ulps = (int)(rng() % 9) - 4;
        for (int u = 0; u < (ulps < 0 ? -ulps : ulps); ++u) x = nextafterf(x, ulps > 0 ? 1e9f : -1e9f);
      }
      bad += dl::round_to_int(x / res) != dl::round_div(x, d);
      ++n;
    }
  }
  const dl::CellDivider d = dl::make_divider(0.15f);
  for (float x : {1e7f, -1e7f, 3e6f, 629145.6f, 629145.7f, 1e-30f, -0.f, 0.f, 0.075f, -0.075f, 0.22500001f}) {
    bad += dl::round_to_int(x / 0.15f) != dl::round_div(x, d);
    ++n;
  }
  printf("n=%ld bad=%ld\n", n, bad);
  return bad != 0;
}
