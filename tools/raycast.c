/* Workload generator helper (NOT product, NOT oracle): exact ray casting of the analytic test scene
 * (planes, axis-aligned boxes, spheres) for tools/synth.py. Plain C, OpenMP if available. */
#include <math.h>
#include <stdint.h>

void synth_raycast(int64_t n, const double* origins, const double* dirs, double ground_z, double facade_y,
                   int n_box, const double* box_lo, const double* box_hi, int n_sph, const double* sph_c,
                   double sph_r, double max_range, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double* o = origins + 3 * i;
    const double* d = dirs + 3 * i;
    double best = INFINITY;
    double t = (ground_z - o[2]) / d[2];
    if (t > 0 && t < best) best = t;
    for (int s = -1; s <= 1; s += 2) {
      t = (s * facade_y - o[1]) / d[1];
      if (t > 0 && t < best) best = t;
    }
    for (int b = 0; b < n_box; ++b) {
      double tn = -INFINITY, tf = INFINITY;
      for (int a = 0; a < 3; ++a) {
        const double inv = 1.0 / d[a];
        double t1 = (box_lo[3 * b + a] - o[a]) * inv, t2 = (box_hi[3 * b + a] - o[a]) * inv;
        if (t1 > t2) { const double tmp = t1; t1 = t2; t2 = tmp; }
        if (t1 > tn) tn = t1;
        if (t2 < tf) tf = t2;
      }
      if (tn <= tf && tn > 0 && tn < best) best = tn;
    }
    for (int s = 0; s < n_sph; ++s) {
      const double oc[3] = {o[0] - sph_c[3 * s], o[1] - sph_c[3 * s + 1], o[2] - sph_c[3 * s + 2]};
      const double b = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
      const double disc = b * b - (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sph_r * sph_r);
      if (disc > 0) {
        t = -b - sqrt(disc);
        if (t > 0 && t < best) best = t;
      }
    }
    out[i] = best > max_range ? INFINITY : best;
  }
}
