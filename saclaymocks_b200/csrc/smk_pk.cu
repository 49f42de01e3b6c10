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

int launch_pk_weights(const PkParams& p, cudaStream_t st) {
  pk_weights_kernel<<<148 * 16, 256, 0, st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

}  // namespace smk
