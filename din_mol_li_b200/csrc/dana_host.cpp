// dana_host.cpp — host-side set-up helpers (include/dml_host.h).  Product code, independent of oracle/.
#include "../../include/dml_host.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" {

void dmlh_rng_init(dmlh_rng *r, int32_t idum) { r->idum = idum; r->ix = -1; r->iy = -1; r->stored = 0; r->g = 0.0; r->calls = 0; }

// Park–Miller (Schrage) combined with a 13/17/5 xorshift; the mantissa scale is the float below 1 over 2^31-1.
double dmlh_ran(dmlh_rng *r) {
  static const double am = (double)nextafterf(1.0f, -1.0f) / 2147483647.0;
  r->calls++;
  if (r->idum <= 0 || r->iy < 0) {
    int32_t a = std::abs(r->idum);
    r->iy = (888889999 ^ a) | 1;
    r->ix = 777755555 ^ a;
    r->idum = a + 1;
  }
  uint32_t x = (uint32_t)r->ix;
  x ^= x << 13; x ^= x >> 17; x ^= x << 5;
  r->ix = (int32_t)x;
  int32_t k = r->iy / 127773;
  r->iy = 16807 * (r->iy - k * 127773) - 2836 * k;
  if (r->iy < 0) r->iy += 2147483647;
  return am * (double)((2147483647 & (r->ix ^ r->iy)) | 1);
}

// Marsaglia polar method; the squared radius and the scale factor live in single precision like the reference's rsq.
double dmlh_gasdev(dmlh_rng *r) {
  if (r->stored) { r->stored = 0; return r->g; }
  double a, b; float s;
  do {
    a = 2.0 * dmlh_ran(r) - 1.0;
    b = 2.0 * dmlh_ran(r) - 1.0;
    s = (float)(a * a + b * b);
  } while (!((double)s > 0.0 && (double)s < 1.0));
  float f = (float)std::sqrt(-2.0 * (double)logf(s) / (double)s);
  r->g = b * (double)f; r->stored = 1;
  return a * (double)f;
}

int32_t dmlh_pos_inic(dmlh_rng *r, double xi, double yi, double alto, double *xyz, int32_t cap) {
  const double rmin = 3.2;
  int32_t n = (int32_t)(1.0 * xi * yi * alto * (double)6.022e-4f);
  if (n > cap) return -n;
  // bucket grid with cells >= rmin; x,y periodic, z open
  int gx = std::max(1, (int)(xi / rmin)), gy = std::max(1, (int)(yi / rmin)), gz = std::max(1, (int)(alto / rmin));
  bool grid = gx >= 3 && gy >= 3;
  double sx = xi / gx, sy = yi / gy, sz = alto / gz;
  std::vector<int> head(grid ? (size_t)gx * gy * gz : 0, -1), next(n, -1);
  auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v >= hi ? hi - 1 : v); };
  const double obx = 1.0 / xi, oby = 1.0 / yi;
  for (int32_t i = 0; i < n; ++i) {
    int tries = 0;
    for (;;) {
      if (++tries > 10000) return -1;
      double p[3];
      p[0] = dmlh_ran(r) * xi; p[1] = dmlh_ran(r) * yi; p[2] = dmlh_ran(r) * alto + 0.0;
      bool clash = false;
      auto close = [&](int j) {
        double dx = p[0] - xyz[3 * j], dy = p[1] - xyz[3 * j + 1], dz = p[2] - xyz[3 * j + 2];
        dx = dx - xi * std::round(dx * obx); dy = dy - yi * std::round(dy * oby);
        return (dx * dx + dy * dy) + dz * dz < rmin * rmin;
      };
      if (!grid) { for (int j = 0; j < i && !clash; ++j) clash = close(j); }
      else {
        int cx = clampi((int)(p[0] / sx), gx), cy = clampi((int)(p[1] / sy), gy), cz = clampi((int)(p[2] / sz), gz);
        for (int dz = -1; dz <= 1 && !clash; ++dz) {
          int z = cz + dz; if (z < 0 || z >= gz) continue;
          for (int dy = -1; dy <= 1 && !clash; ++dy) for (int dx = -1; dx <= 1 && !clash; ++dx) {
            int c = (cx + dx + gx) % gx + gx * ((cy + dy + gy) % gy + gy * z);
            for (int j = head[c]; j >= 0; j = next[j]) if (close(j)) { clash = true; break; }
          }
        }
      }
      if (clash) continue;
      xyz[3 * i] = p[0]; xyz[3 * i + 1] = p[1]; xyz[3 * i + 2] = p[2];
      if (grid) {
        int c = clampi((int)(p[0] / sx), gx) + gx * (clampi((int)(p[1] / sy), gy) + gy * clampi((int)(p[2] / sz), gz));
        next[i] = head[c]; head[c] = i;
      }
      break;
    }
  }
  char buf[64];
  for (int64_t k = 0; k < (int64_t)n * 3; ++k) { snprintf(buf, sizeof buf, "%.12f", xyz[k]); xyz[k] = strtod(buf, nullptr); }
  return n;
}

} // extern "C"
