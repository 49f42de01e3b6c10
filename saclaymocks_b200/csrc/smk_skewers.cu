// Skewer gather: batched ReadSpec (bin/make_spectra.py:90-139 with ComputeWeight :40-62 and
// computeRho :66-68) for every quasar of one x-slab in one launch.
//
// For each line-of-sight pixel: the cell that contains it (fp64 index arithmetic, exactly the
// reference's expressions), a (2*dmax+1)^3 Gaussian-weighted average of up to 10 fields, the
// eta_par = x_i x_j eta_ij / r^2 and v_par = v.r/|r| contractions.  The 343 reference weights
// exp(-|dr|^2 / 2 DX^2) factorise as wx*wy*wz, so 21 exponentials are evaluated per pixel instead
// of 343; sums run in float32 (|error| ~1e-6 of the field rms, far inside the 1e-5 tolerance on F).
#include <math.h>

#include "smk_internal.h"

namespace smk {

struct SkewerParams {
  const float* f[10];
  int nx, ny, nz;          // full box
  double dx, dy, dz, r0;
  int ix0, nxs;            // slab: global planes [ix0, ix0+nxs)
  double xmin, xmax;       // ownership: xmin < X <= xmax
  int rsd, dla;
  int nqso, npix;
  const double* qso;       // [nqso][4] X,Y,Z,R
  const int* npix_forest;  // [nqso]
  const double* rvec;      // [npix]
  float* delta_l;
  float* eta_par;
  float* vpar;
};

template <int DMAX, int NF>
__device__ __forceinline__ void gather(const SkewerParams& p, int ixl, int iy, int iz, const float (&wx)[2 * DMAX + 1],
                                       const float (&wy)[2 * DMAX + 1], const float (&wz)[2 * DMAX + 1],
                                       float (&acc)[NF]) {
  constexpr int W = 2 * DMAX + 1;
#pragma unroll
  for (int f = 0; f < NF; ++f) acc[f] = 0.f;
  int lz[W];
#pragma unroll
  for (int c = 0; c < W; ++c) lz[c] = min(max(iz + c - DMAX, 0), p.nz - 1);
  for (int a = 0; a < W; ++a) {
    int la = min(max(ixl + a - DMAX, 0), p.nxs - 1);
    for (int b = 0; b < W; ++b) {
      int lb = min(max(iy + b - DMAX, 0), p.ny - 1);
      float wab = wx[a] * wy[b];
      long long row = ((long long)la * p.ny + lb) * p.nz;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float* __restrict__ src = p.f[f] + row;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < W; ++c) s = fmaf(wz[c], __ldg(src + lz[c]), s);
        acc[f] = fmaf(wab, s, acc[f]);
      }
    }
  }
}

template <int DMAX>
__global__ void __launch_bounds__(128) skewers_kernel(SkewerParams p, int nchunk) {
  constexpr int W = 2 * DMAX + 1;
  const int q = blockIdx.x / nchunk;
  const int i = (blockIdx.x - q * nchunk) * blockDim.x + threadIdx.x;
  if (i >= p.npix) return;
  const double X = p.qso[4 * q], Y = p.qso[4 * q + 1], Z = p.qso[4 * q + 2], R = p.qso[4 * q + 3];
  const double r = p.rvec[i];
  const double xv = r * X / R;                       // make_spectra.py:443-450 (same operation order)
  if (!(xv > p.xmin) || !(xv <= p.xmax)) return;     // pixel belongs to another slab
  const size_t o = (size_t)q * p.npix + i;
  if (i >= p.npix_forest[q]) {                       // make_spectra.py:99-101
    p.delta_l[o] = -1000000.f;
    if (p.eta_par) p.eta_par[o] = 0.f;
    if (p.vpar) p.vpar[o] = 0.f;
    return;
  }
  const double yv = r * Y / R, zv = r * Z / R;
  const double LX = p.dx * p.nx, LY = p.dy * p.ny, LZ = p.dz * p.nz;
  // make_spectra.py:47-49,54-55 (int() truncates toward zero)
  const int ix = (int)((xv + LX / 2) / p.dx), iy = (int)((yv + LY / 2) / p.dy), iz = (int)((zv + LZ / 2 - p.r0) / p.dz);
  const float ox = (float)((ix + 0.5) * p.dx - LX / 2 - xv);
  const float oy = (float)((iy + 0.5) * p.dy - LY / 2 - yv);
  const float oz = (float)((iz + 0.5) * p.dz - LZ / 2 + p.r0 - zv);
  const float inv_sig2 = (float)(1.0 / (2.0 * p.dx * p.dx));
  const float fdx = (float)p.dx, fdy = (float)p.dy, fdz = (float)p.dz;
  float wx[W], wy[W], wz[W];
  float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
  for (int a = 0; a < W; ++a) {
    float tx = (a - DMAX) * fdx + ox, ty = (a - DMAX) * fdy + oy, tz = (a - DMAX) * fdz + oz;
    wx[a] = expf(-tx * tx * inv_sig2);
    wy[a] = expf(-ty * ty * inv_sig2);
    wz[a] = expf(-tz * tz * inv_sig2);
    sx += wx[a];
    sy += wy[a];
    sz += wz[a];
  }
  const float inv_sw = 1.0f / (sx * sy * sz);
  const int ixl = ix - p.ix0;
  if (p.rsd && p.dla) {
    float acc[10];
    gather<DMAX, 10>(p, ixl, iy, iz, wx, wy, wz, acc);
    const double RR = xv * xv + yv * yv + zv * zv;
    double e = (xv * (double)(acc[1] * inv_sw) * xv + yv * (double)(acc[2] * inv_sw) * yv +
                zv * (double)(acc[3] * inv_sw) * zv + 2 * xv * (double)(acc[4] * inv_sw) * yv +
                2 * xv * (double)(acc[5] * inv_sw) * zv + 2 * yv * (double)(acc[6] * inv_sw) * zv) / RR;
    double v = ((double)(acc[7] * inv_sw) * xv + (double)(acc[8] * inv_sw) * yv + (double)(acc[9] * inv_sw) * zv) /
               sqrt(RR);
    p.delta_l[o] = acc[0] * inv_sw;
    p.eta_par[o] = (float)e;
    p.vpar[o] = (float)v;
  } else if (p.rsd) {
    float acc[7];
    gather<DMAX, 7>(p, ixl, iy, iz, wx, wy, wz, acc);
    const double RR = xv * xv + yv * yv + zv * zv;
    double e = (xv * (double)(acc[1] * inv_sw) * xv + yv * (double)(acc[2] * inv_sw) * yv +
                zv * (double)(acc[3] * inv_sw) * zv + 2 * xv * (double)(acc[4] * inv_sw) * yv +
                2 * xv * (double)(acc[5] * inv_sw) * zv + 2 * yv * (double)(acc[6] * inv_sw) * zv) / RR;
    p.delta_l[o] = acc[0] * inv_sw;
    p.eta_par[o] = (float)e;
    if (p.vpar) p.vpar[o] = 0.f;
  } else {
    float acc[1];
    gather<DMAX, 1>(p, ixl, iy, iz, wx, wy, wz, acc);
    p.delta_l[o] = acc[0] * inv_sw;
    if (p.eta_par) p.eta_par[o] = 0.f;
    if (p.vpar) p.vpar[o] = 0.f;
  }
}

int launch_skewers(const SkewerParams& p, int dmax, cudaStream_t st) {
  if (p.nqso == 0 || p.npix == 0) return SMK_OK;
  const int NT = 128;
  int nchunk = (p.npix + NT - 1) / NT;
  long long nblocks = (long long)nchunk * p.nqso;
  if (nblocks > 2147483647LL) { set_error("smk_skewers: too many quasars for one launch"); return SMK_ERR_ARG; }
  switch (dmax) {
    case 1: skewers_kernel<1><<<(unsigned)nblocks, NT, 0, st>>>(p, nchunk); break;
    case 2: skewers_kernel<2><<<(unsigned)nblocks, NT, 0, st>>>(p, nchunk); break;
    case 3: skewers_kernel<3><<<(unsigned)nblocks, NT, 0, st>>>(p, nchunk); break;
    case 4: skewers_kernel<4><<<(unsigned)nblocks, NT, 0, st>>>(p, nchunk); break;
    default: set_error("smk_skewers: dmax must be in 1..4"); return SMK_ERR_UNSUPPORTED;
  }
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

}  // namespace smk

extern "C" int smk_skewers(smk_ctx* ctx, const smk_geom* g, const float* const fields[10], int ix0, int nxs,
                           double xmin, double xmax, int rsd, int dla, int nqso, const double* qso_xyzr,
                           const int* npix_forest, const double* rvec, int npix, float* delta_l, float* eta_par,
                           float* vpar) {
  using namespace smk;
  if (!g || !fields || !fields[0] || !delta_l || (nqso > 0 && (!qso_xyzr || !npix_forest || !rvec))) {
    set_error("smk_skewers: null argument");
    return SMK_ERR_ARG;
  }
  if (rsd && (!eta_par || !fields[1] || !fields[2] || !fields[3] || !fields[4] || !fields[5] || !fields[6])) {
    set_error("smk_skewers: rsd needs the six eta fields and eta_par");
    return SMK_ERR_ARG;
  }
  if (rsd && dla && (!vpar || !fields[7] || !fields[8] || !fields[9])) {
    set_error("smk_skewers: dla needs the three velocity fields and vpar");
    return SMK_ERR_ARG;
  }
  SkewerParams p{};
  for (int i = 0; i < 10; ++i) p.f[i] = fields[i];
  p.nx = g->nx; p.ny = g->ny; p.nz = g->nz;
  p.dx = g->dx; p.dy = g->dy; p.dz = g->dz; p.r0 = g->r0;
  p.ix0 = ix0; p.nxs = nxs; p.xmin = xmin; p.xmax = xmax;
  p.rsd = rsd; p.dla = (rsd && dla);
  p.nqso = nqso; p.npix = npix;
  p.qso = qso_xyzr; p.npix_forest = npix_forest; p.rvec = rvec;
  p.delta_l = delta_l; p.eta_par = eta_par; p.vpar = vpar;
  return launch_skewers(p, g->dmax, smk_ctx_stream(ctx));
}
