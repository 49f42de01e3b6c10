// Host-side unit test of the radix butterflies of saclaymocks_b200/csrc/smk_fft.cuh against a naive DFT in double
// (compiled with nvcc, runs on the CPU: tests/test_fft_butterflies_cpu.py).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../saclaymocks_b200/csrc/smk_fft.cuh"

template <int R, bool INV>
static double check() {
  float2 v[R];
  double xr[R], xi[R];
  srand(R * 2 + INV);
  for (int t = 0; t < R; ++t) {
    xr[t] = rand() / (double)RAND_MAX - 0.5;
    xi[t] = rand() / (double)RAND_MAX - 0.5;
    v[t] = make_float2((float)xr[t], (float)xi[t]);
    xr[t] = v[t].x; xi[t] = v[t].y;
  }
  smk::Butterfly<R, INV>::run(v);
  double err = 0, nrm = 0;
  for (int q = 0; q < R; ++q) {
    double yr = 0, yi = 0;
    for (int t = 0; t < R; ++t) {
      double a = (INV ? 2.0 : -2.0) * M_PI * t * q / R;
      yr += xr[t] * cos(a) - xi[t] * sin(a);
      yi += xr[t] * sin(a) + xi[t] * cos(a);
    }
    err += (v[q].x - yr) * (v[q].x - yr) + (v[q].y - yi) * (v[q].y - yi);
    nrm += yr * yr + yi * yi;
  }
  return sqrt(err / nrm);
}

int main() {
  double worst = 0;
#define T(R) { double a = check<R, false>(), b = check<R, true>(); printf("radix %d fwd %.3e inv %.3e\n", R, a, b); \
               worst = fmax(worst, fmax(a, b)); }
  T(2) T(3) T(4) T(5) T(8) T(16) T(24) T(32)
  printf("worst %.3e\n", worst);
  return worst < 5e-7 ? 0 : 1;
}
