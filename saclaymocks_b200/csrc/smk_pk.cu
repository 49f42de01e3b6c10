// Spectral weight tables on the GPU: W = float32(sqrt(float32(max(spline(|k|), 0)) / Vcell)) for every mode of this
// rank's k-slab -- the arithmetic of bin/interpolate_pk.py:17-26, 63-77 (16 CPU processes and a 20-minute SLURM
// budget in the reference, submit_mocks.py:1142-1146).  The cubic spline is evaluated in float64 from its
// piecewise-polynomial form (breaks + 4 coefficients per interval), |k| in the reference's float32 rounding order.
#include <math.h>

#include "smk_internal.h"

namespace smk {

__global__ void pk_weights_kernel(PkParams p) {
  const size_t n = (size_t)p.nx * p.nyl * p.nzh;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int iz = (int)(idx % p.nzh);
    const size_t t = idx / p.nzh;
    const int iy = (int)(t % p.nyl), ix = (int)(t / p.nyl);
    const float kx = __ldg(p.kx + ix), ky = __ldg(p.ky + p.y0 + iy), kz = __ldg(p.kz + iz);
    const float k = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz)));
    const double x = (double)k;
    // interval i with breaks[i] <= x < breaks[i+1]; the end intervals extrapolate, like FITPACK's splev
    int lo = 0, hi = p.nint;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(p.breaks + mid) <= x) lo = mid; else hi = mid;
    }
    const double dx = x - __ldg(p.breaks + lo);
    const double c0 = __ldg(p.coefs + lo), c1 = __ldg(p.coefs + p.nint + lo), c2 = __ldg(p.coefs + 2 * (size_t)p.nint + lo),
                 c3 = __ldg(p.coefs + 3 * (size_t)p.nint + lo);
    const double v = fma(fma(fma(c0, dx, c1), dx, c2), dx, c3);
    const float pf = (float)fmax(v, 0.0);
    p.out[idx] = __fsqrt_rn(__fdiv_rn(pf, p.vcell));
  }
}

// ---- 3-D power spectrum estimator: bins |boxk|^2 of this rank's k-slab in |k| with the
// Hermitian multiplicity of the half-complex layout (planes kz = 0 and kz = Nyquist count once, the others twice).
// Per bin: sum of mult * |boxk|^2, sum of mult, sum of mult * |k| (double).  Block-local histograms in shared memory,
// one atomicAdd per bin and block at the end.
struct PkEstParams {
  const float2* boxk;
  const float *kx, *ky, *kz;
  int nx, nyl, nzh, pitch, y0, nz_even;
  int nbins;
  float kmin, inv_dk;
  double* out;       // [3][nbins]
};

__global__ void __launch_bounds__(256) pk_estimate_kernel(PkEstParams p) {
  extern __shared__ double sh[];                 // [3][nbins]
  for (int i = threadIdx.x; i < 3 * p.nbins; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const size_t n = (size_t)p.nx * p.nyl * p.nzh;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int iz = (int)(idx % p.nzh);
    const size_t t = idx / p.nzh;
    const int iy = (int)(t % p.nyl), ix = (int)(t / p.nyl);
    const float kx = __ldg(p.kx + ix), ky = __ldg(p.ky + p.y0 + iy), kz = __ldg(p.kz + iz);
    const float k = sqrtf(kx * kx + ky * ky + kz * kz);
    const int b = (int)floorf((k - p.kmin) * p.inv_dk);
    if (b < 0 || b >= p.nbins) continue;
    const float2 v = p.boxk[((size_t)ix * p.nyl + iy) * p.pitch + iz];
    const double m = (iz == 0 || (p.nz_even && iz == p.nzh - 1)) ? 1.0 : 2.0;
    atomicAdd(sh + b, m * ((double)v.x * v.x + (double)v.y * v.y));
    atomicAdd(sh + p.nbins + b, m);
    atomicAdd(sh + 2 * p.nbins + b, m * (double)k);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * p.nbins; i += blockDim.x)
    if (sh[i] != 0.0) atomicAdd(p.out + i, sh[i]);
}

int launch_pk_estimate(const float2* boxk, const float* kx, const float* ky, const float* kz, int nx, int nyl, int nzh,
                       int pitch, int y0, int nz, int nbins, double kmin, double kmax, double* out, cudaStream_t st) {
  if (nbins < 1 || nbins > 2048 || !(kmax > kmin)) { set_error("smk_pk_estimate: need 1..2048 bins and kmax > kmin"); return SMK_ERR_ARG; }
  PkEstParams p{boxk, kx, ky, kz, nx, nyl, nzh, pitch, y0, nz % 2 == 0, nbins, (float)kmin,
                (float)(nbins / (kmax - kmin)), out};
  pk_estimate_kernel<<<148 * 8, 256, 3 * nbins * sizeof(double), st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

int launch_pk_weights(const PkParams& p, cudaStream_t st) {
  pk_weights_kernel<<<148 * 16, 256, 0, st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

}  // namespace smk

// ---- float64 1-D complex FFT on the GPU for the P(k) <-> xi(r) pair of LogNormalP (py/SaclayMocks/powerspectrum.py:
// 145-200: two np.fft.fft of 2^20 and 2^19 points).  Stockham autosort radix 2: log2(n) passes over global memory,
// ping-ponging between two buffers, twiddles from sincospi in float64 (error ~1e-16 like pocketfft's).  The arrays
// are a few MB and the passes a few microseconds each: nothing here needs shared memory.
namespace smk {

// one pass: sub-transform length len (n / stride), stride = number of interleaved sequences
__global__ void __launch_bounds__(256) fft1d_pass_kernel(const double2* __restrict__ x, double2* __restrict__ y, int len,
                                                         int stride, int half_n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= half_n) return;
  const int m = len >> 1;
  const int p = t / stride, q = t - p * stride;            // butterfly p of the current length, sequence q
  double s, c;
  sincospi(-2.0 * (double)p / (double)len, &s, &c);        // w = exp(-2 pi i p / len)
  const double2 a = x[q + stride * p], b = x[q + stride * (p + m)];
  const double2 d = make_double2(a.x - b.x, a.y - b.y);
  y[q + stride * (2 * p)] = make_double2(a.x + b.x, a.y + b.y);
  y[q + stride * (2 * p + 1)] = make_double2(d.x * c - d.y * s, d.x * s + d.y * c);
}

}  // namespace smk

extern "C" int smk_fft1d_f64(smk_ctx* ctx, int n, const double* in, double* out, double* work) {
  using namespace smk;
  if (!in || !out || !work || n < 2 || (n & (n - 1))) { set_error("smk_fft1d_f64: n must be a power of two >= 2"); return SMK_ERR_ARG; }
  if (in == out || in == work || out == work) { set_error("smk_fft1d_f64: in, out and work must be distinct"); return SMK_ERR_ARG; }
  cudaStream_t st = smk_ctx_stream(ctx);
  int passes = 0;
  for (int v = n; v > 1; v >>= 1) ++passes;
  // the last pass must land in `out`: the first pass writes to out when the number of passes is odd, else to work
  const double2* src = reinterpret_cast<const double2*>(in);
  double2* a = reinterpret_cast<double2*>((passes & 1) ? out : work);
  double2* b = reinterpret_cast<double2*>((passes & 1) ? work : out);
  const int half = n / 2;
  for (int len = n, stride = 1; len > 1; len >>= 1, stride <<= 1) {
    fft1d_pass_kernel<<<(half + 255) / 256, 256, 0, st>>>(src, a, len, stride, half);
    SMK_CUDA_OK(cudaGetLastError());
    src = a;
    double2* tmp = a; a = b; b = tmp;
  }
  return SMK_OK;
}
