// Host-side unit test of saclaymocks_b200/csrc/smk_fft.cuh against a naive DFT in double (compiled with nvcc, runs on
// the CPU: tests/test_fft_butterflies_cpu.py): every radix butterfly, and every plan (radix sequence, sub-transform
// sizes, inter-stage twiddles W_L[q * o * N/M], digit-reversed output order nat()) through a host restatement of the
// in-place decimation-in-frequency stage dif_stage() executes on the device.
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

// One DIF stage exactly as dif_stage() runs it (same index arithmetic, same twiddle table look-up), all butterflies
// of one line in sequence.
template <class P, int STAGE, bool INV>
static void host_stage(float2* x, const float2* tw, int twmul) {
  constexpr int R = P::radix(STAGE), M = P::sub(STAGE), MQ = M / R, NB = P::N / R;
  constexpr bool LAST = (STAGE == P::S - 1);
  for (int j = 0; j < NB; ++j) {
    const int b = j / MQ, o = j - b * MQ;
    float2 v[R];
    for (int t = 0; t < R; ++t) v[t] = x[b * M + o + t * MQ];
    smk::Butterfly<R, INV>::run(v);
    if (!LAST) {
      const int oc = o * ((P::N / M) * twmul);
      for (int q = 1; q < R; ++q) {
        float2 w = tw[q * oc];
        if (INV) w.y = -w.y;
        v[q] = smk::cmul(v[q], w);
      }
    }
    for (int q = 0; q < R; ++q) x[b * M + o + q * MQ] = v[q];
  }
}
template <class P, int S0, bool INV>
static void host_stages(float2* x, const float2* tw, int twmul) {
  if constexpr (S0 < P::S) {
    host_stage<P, S0, INV>(x, tw, twmul);
    host_stages<P, S0 + 1, INV>(x, tw, twmul);
  }
}

template <int N, bool INV>
static double check_plan(int twmul) {
  using P = typename smk::PlanFor<N>::type;
  static_assert(P::N == N, "plan size");
  const int L = N * twmul;
  float2* tw = new float2[L];
  for (int k = 0; k < L; ++k) tw[k] = make_float2((float)cos(-2.0 * M_PI * k / L), (float)sin(-2.0 * M_PI * k / L));
  float2* x = new float2[N];
  double *xr = new double[N], *xi = new double[N];
  srand(N + INV);
  for (int t = 0; t < N; ++t) {
    x[t] = make_float2((float)(rand() / (double)RAND_MAX - 0.5), (float)(rand() / (double)RAND_MAX - 0.5));
    xr[t] = x[t].x; xi[t] = x[t].y;
  }
  host_stages<P, 0, INV>(x, tw, twmul);
  // naive DFT at a subset of output indices (all of them for N <= 512)
  const int step = N <= 512 ? 1 : 7;
  double err = 0, nrm = 0;
  bool perm_ok = true;
  for (int p = 0; p < N; ++p) perm_ok = perm_ok && P::pos(P::nat(p)) == p;
  if (P::S == 2)   // closed form the z kernels use to read a two-stage plan's outputs without a re-sort (ZTraits::nat)
    for (int k = 0; k < N; ++k) perm_ok = perm_ok && P::pos(k) == (k % P::radix(0)) * P::radix(1) + k / P::radix(0);
  for (int p = 0; p < N; p += step) {
    const int k = P::nat(p);
    double yr = 0, yi = 0;
    for (int t = 0; t < N; ++t) {
      const double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long long)t * k % N) / N;
      yr += xr[t] * cos(a) - xi[t] * sin(a);
      yi += xr[t] * sin(a) + xi[t] * cos(a);
    }
    err += (x[p].x - yr) * (x[p].x - yr) + (x[p].y - yi) * (x[p].y - yi);
    nrm += yr * yr + yi * yi;
  }
  delete[] tw; delete[] x; delete[] xr; delete[] xi;
  return perm_ok ? sqrt(err / nrm) : 1.0;
}

int main() {
  double worst = 0;
#define T(R) { double a = check<R, false>(), b = check<R, true>(); printf("radix %d fwd %.3e inv %.3e\n", R, a, b); \
               worst = fmax(worst, fmax(a, b)); }
  T(2) T(3) T(4) T(5) T(8) T(16) T(24) T(32)
  printf("worst butterfly %.3e\n", worst);
  double worst_plan = 0;
  // twmul = 1: strided passes (table W_N); twmul = 2: z passes (table W_NZ with NZ = 2 M)
#define PL(N) { double a = check_plan<N, false>(1), b = check_plan<N, true>(1), c = check_plan<N, true>(2);           \
                printf("plan %d (%d stages, first radix %d) fwd %.3e inv %.3e inv/twmul2 %.3e\n", N,                  \
                       smk::PlanFor<N>::type::S, smk::PlanFor<N>::type::radix(0), a, b, c);                            \
                worst_plan = fmax(worst_plan, fmax(a, fmax(b, c))); }
  PL(4) PL(8) PL(12) PL(16) PL(24) PL(32) PL(48) PL(64) PL(96) PL(128) PL(256) PL(384) PL(512) PL(768) PL(1024)
  PL(2048) PL(2560) PL(4096)
  printf("worst plan %.3e\n", worst_plan);
  return (worst < 5e-7 && worst_plan < 2e-6) ? 0 : 1;
}
