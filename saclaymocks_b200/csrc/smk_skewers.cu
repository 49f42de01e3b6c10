// Skewer gather: batched ReadSpec (bin/make_spectra.py:90-139 with ComputeWeight :40-62 and
// computeRho :66-68) for every quasar of one x-slab in one launch.
//
// For each line-of-sight pixel: the cell that contains it (fp64 index arithmetic, exactly the
// reference's expressions), a (2*dmax+1)^3 Gaussian-weighted average of up to 10 fields, the
// eta_par = x_i x_j eta_ij / r^2 and v_par = v.r/|r| contractions.  The 343 reference weights
// exp(-|dr|^2 / 2 DX^2) factorise as wx*wy*wz, so 21 exponentials are evaluated per pixel instead
// of 343; sums run in float32 (|error| ~1e-6 of the field rms, far inside the 1e-5 tolerance on F).
#include <math.h>
#include <stdlib.h>

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
  // fused FGPA epilogue (smk_skewers_fgpa): flux = exp(-a exp(b G (delta_l + delta_s + c eta_par))), util.py:421-433
  const float* delta_s;    // [nqso][npix] or null
  const float *fg_G, *fg_a, *fg_b, *fg_c;   // [npix]
  float* flux;             // [nqso][npix] or null (no epilogue)
  int pfd;                 // distance (rows of the window) of the L1 row prefetch
  int pf;                  // tuning bits: 1 = L2 prefetch of the next x slab of the window, 2 = L1 prefetch of the next row, 4 = L1 prefetch of the next x slab
};

// Blackwell packed FP32: one FFMA2 / FMUL2 issues two fused multiply-adds (fma.rn.f32x2, sm_100+)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// FGPA of one pixel, the float32 arithmetic of fgpa_kernel (smk_spectra1d.cu) operation for operation
__device__ __forceinline__ void store_flux(const SkewerParams& p, size_t o, int i, float delta_l, float eta) {
  float d = delta_l;
  if (p.delta_s) d += p.delta_s[o];
  if (p.eta_par && p.fg_c) d += __ldg(p.fg_c + i) * eta;
  const float tau_over_a = expf(__ldg(p.fg_b + i) * __ldg(p.fg_G + i) * d);   // util.py:427
  p.flux[o] = expf(-__ldg(p.fg_a + i) * tau_over_a);                          // util.py:428
}

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

// ---- register-blocked variant: each thread owns P consecutive pixels of one sightline and walks the UNION of
// their (2*DMAX+1)^3 windows once, so that every field value loaded is used by P pixels.  P consecutive pixels
// span (P-1)*pixel < one cell, hence the union is at most one cell wider per axis; a pixel's weight is zero outside
// its own window, which keeps the result identical to the reference's truncated Gaussian sum.
template <int DMAX, int P, int NF, bool INTERIOR>
__device__ __forceinline__ void gather_multi(const SkewerParams& p, const float* const (&fp)[NF], int bx, int by,
                                             int bz, int nxu, int nyu, const int (&dix)[P], const float (&ox)[P],
                                             const float* wy_s /* [b][q] of this thread, stride 128 */,
                                             const float (&wz)[P][2 * DMAX + 2], float inv_sig2, float (&acc)[NF][P],
                                             float (&sx)[P]) {
  constexpr int WU = 2 * DMAX + 2;
  const float fdx = (float)p.dx;
#pragma unroll
  for (int f = 0; f < NF; ++f)
#pragma unroll
    for (int q = 0; q < P; ++q) acc[f][q] = 0.f;
  int lz[WU];
#pragma unroll
  for (int c = 0; c < WU; ++c) lz[c] = INTERIOR ? c : min(max(bz - DMAX + c, 0), p.nz - 1);
  const int z0 = INTERIOR ? bz - DMAX : 0;      // INTERIOR: the z window [bz-DMAX, bz+DMAX+1] needs no clamping
#pragma unroll
  for (int q = 0; q < P; ++q) sx[q] = 0.f;
  const unsigned plane = (unsigned)p.ny * (unsigned)p.nz;
  for (int a = 0; a < nxu; ++a) {
    const int la = INTERIOR ? (bx - DMAX + a - p.ix0) : min(max(bx - DMAX + a - p.ix0, 0), p.nxs - 1);
    float wxa[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      const int m = a - DMAX - dix[q];                      // cell offset from pixel q's own cell
      const float t = m * fdx + ox[q];
      wxa[q] = (m >= -DMAX && m <= DMAX) ? __expf(-t * t * inv_sig2) : 0.f;
      sx[q] += wxa[q];
    }
    for (int b = 0; b < nyu; ++b) {
      const int lb = INTERIOR ? (by - DMAX + b) : min(max(by - DMAX + b, 0), p.ny - 1);
      float wab[P];
#pragma unroll
      for (int q = 0; q < P; ++q) wab[q] = wxa[q] * wy_s[(b * P + q) * 128];
      const size_t row = (size_t)la * plane + (unsigned)lb * (unsigned)p.nz + z0;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float* __restrict__ src = fp[f] + row;
        float r[WU];
#pragma unroll
        for (int c = 0; c < WU; ++c) r[c] = __ldg(src + lz[c]);
#pragma unroll
        for (int q = 0; q < P; ++q) {
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < WU; ++c) s = fmaf(wz[q][c], r[c], s);
          acc[f][q] = fmaf(wab[q], s, acc[f][q]);
        }
      }
    }
  }
}

// Same walk with packed arithmetic: the z contraction runs on pairs of cells (4 FMUL2/FFMA2 + one add per pixel and
// row instead of 8 FFMA) and the row accumulation on pairs of pixels (P/2 FFMA2 instead of P FFMA): 30 instead of 44
// issue slots per (row, field) at P = 4.  wz2[q][j] = (wz[q][2j], wz[q][2j+1]); acc2[f][h] = pixels (2h, 2h+1).
template <int DMAX, int P, int NF, bool INTERIOR>
__device__ __forceinline__ void gather_multi2(const SkewerParams& p, const float* const (&fp)[NF], int bx, int by,
                                              int bz, int nxu, int nyu, const int (&dix)[P], const float (&ox)[P],
                                              const float* wy_s, const float2 (&wz2)[P][DMAX + 1], float inv_sig2,
                                              float2 (&acc2)[NF][P / 2], float (&sx)[P]) {
  constexpr int WU = 2 * DMAX + 2;
  static_assert(P % 2 == 0, "packed gather works on pixel pairs");
  const float fdx = (float)p.dx;
#pragma unroll
  for (int f = 0; f < NF; ++f)
#pragma unroll
    for (int h = 0; h < P / 2; ++h) acc2[f][h] = make_float2(0.f, 0.f);
  int lz[WU];
#pragma unroll
  for (int c = 0; c < WU; ++c) lz[c] = INTERIOR ? c : min(max(bz - DMAX + c, 0), p.nz - 1);
  const int z0 = INTERIOR ? bz - DMAX : 0;
#pragma unroll
  for (int q = 0; q < P; ++q) sx[q] = 0.f;
  const unsigned plane = (unsigned)p.ny * (unsigned)p.nz;
  for (int a = 0; a < nxu; ++a) {
    const int la = INTERIOR ? (bx - DMAX + a - p.ix0) : min(max(bx - DMAX + a - p.ix0, 0), p.nxs - 1);
    if (INTERIOR && (p.pf & 1) && a + 1 < nxu) {
      // the rows of the next x slab of the window are known now and needed ~8 row iterations from now: pull the
      // sector at the centre of each into L2 (the loads of a row otherwise expose one DRAM latency per iteration)
      const size_t nrow = (size_t)(la + 1) * plane + (unsigned)(by - DMAX) * (unsigned)p.nz + z0 + DMAX;
      for (int b = 0; b < nyu; ++b)
#pragma unroll
        for (int f = 0; f < NF; ++f) asm volatile("prefetch.global.L2 [%0];" ::"l"(fp[f] + nrow + (size_t)b * p.nz));
    }
    if (INTERIOR && (p.pf & 4) && a + 1 < nxu) {       // same, into L1 (both sectors the window can straddle)
      const size_t nrow = (size_t)(la + 1) * plane + (unsigned)(by - DMAX) * (unsigned)p.nz + z0;
      for (int b = 0; b < nyu; ++b)
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(fp[f] + nrow + (size_t)b * p.nz));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(fp[f] + nrow + (size_t)b * p.nz + WU - 1));
        }
    }
    float wxa[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      const int m = a - DMAX - dix[q];
      const float t = m * fdx + ox[q];
      wxa[q] = (m >= -DMAX && m <= DMAX) ? __expf(-t * t * inv_sig2) : 0.f;
      sx[q] += wxa[q];
    }
    for (int b = 0; b < nyu; ++b) {
      const int lb = INTERIOR ? (by - DMAX + b) : min(max(by - DMAX + b, 0), p.ny - 1);
      if (INTERIOR && (p.pf & 2)) {
        // L1 prefetch of the window row p.pfd iterations ahead (wrapping into the next x slab)
        int nb = b + p.pfd, na = a;
        if (nb >= nyu) { nb -= nyu; ++na; }
        if (na < nxu) {
          const size_t nrow = (size_t)(la + (na - a)) * plane + (unsigned)(by - DMAX + nb) * (unsigned)p.nz + z0;
#pragma unroll
          for (int f = 0; f < NF; ++f) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(fp[f] + nrow));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(fp[f] + nrow + WU - 1));
          }
        }
      }
      float2 wab2[P / 2];
#pragma unroll
      for (int h = 0; h < P / 2; ++h)
        wab2[h] = make_float2(wxa[2 * h] * wy_s[(b * P + 2 * h) * 128], wxa[2 * h + 1] * wy_s[(b * P + 2 * h + 1) * 128]);
      const size_t row = (size_t)la * plane + (unsigned)lb * (unsigned)p.nz + z0;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float* __restrict__ src = fp[f] + row;
        float2 r2[WU / 2];
#pragma unroll
        for (int j = 0; j < WU / 2; ++j) r2[j] = make_float2(__ldg(src + lz[2 * j]), __ldg(src + lz[2 * j + 1]));
#pragma unroll
        for (int h = 0; h < P / 2; ++h) {
          float2 s0 = fmul2(wz2[2 * h][0], r2[0]), s1 = fmul2(wz2[2 * h + 1][0], r2[0]);
#pragma unroll
          for (int j = 1; j < WU / 2; ++j) {
            s0 = ffma2(wz2[2 * h][j], r2[j], s0);
            s1 = ffma2(wz2[2 * h + 1][j], r2[j], s1);
          }
          acc2[f][h] = ffma2(wab2[h], make_float2(s0.x + s0.y, s1.x + s1.y), acc2[f][h]);
        }
      }
    }
  }
}

template <int DMAX, int P>
__device__ __forceinline__ void pixel_xyz(const SkewerParams& p, int q, int i, double& xv, double& yv, double& zv) {
  const double R = p.qso[4 * q + 3], r = p.rvec[i];
  xv = r * p.qso[4 * q] / R;                 // make_spectra.py:443-452 (same operation order)
  yv = r * p.qso[4 * q + 1] / R;
  zv = r * p.qso[4 * q + 2] / R;
}

template <int DMAX, int P, int MINB, bool F2>
__global__ void __launch_bounds__(128, MINB) skewers_multi_kernel(const __grid_constant__ SkewerParams p, int nchunk) {
  constexpr int WU = 2 * DMAX + 2;
  const int q = blockIdx.x / nchunk;
  const int i0 = ((blockIdx.x - q * nchunk) * blockDim.x + threadIdx.x) * P;
  if (i0 >= p.npix) return;
  const int nfor = p.npix_forest[q];
  const double LX = p.dx * p.nx, LY = p.dy * p.ny, LZ = p.dz * p.nz;
  const float inv_sig2 = (float)(1.0 / (2.0 * p.dx * p.dx));
  const float fdz = (float)p.dz;
  unsigned actmask = 0;
  int ix[P], iy[P], iz[P];
  float ox[P], oy[P], oz[P];
  int bx = 1 << 30, by = 1 << 30, bz = 1 << 30, tx = -(1 << 30), ty = -(1 << 30), tz = -(1 << 30);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int i = i0 + k;
    ix[k] = iy[k] = iz[k] = 0;
    ox[k] = oy[k] = oz[k] = 0.f;
    if (i >= p.npix) continue;
    double xv, yv, zv;
    pixel_xyz<DMAX, P>(p, q, i, xv, yv, zv);
    if (!(xv > p.xmin) || !(xv <= p.xmax)) continue;            // owned by another slab
    if (i >= nfor) {                                            // make_spectra.py:99-101
      const size_t o = (size_t)q * p.npix + i;
      p.delta_l[o] = -1000000.f;
      if (p.eta_par) p.eta_par[o] = 0.f;
      if (p.vpar) p.vpar[o] = 0.f;
      if (p.flux) store_flux(p, o, i, -1000000.f, 0.f);
      continue;
    }
    ix[k] = (int)((xv + LX / 2) / p.dx);                        // make_spectra.py:47-49
    iy[k] = (int)((yv + LY / 2) / p.dy);
    iz[k] = (int)((zv + LZ / 2 - p.r0) / p.dz);
    ox[k] = (float)((ix[k] + 0.5) * p.dx - LX / 2 - xv);        // cell centre - pixel, make_spectra.py:54-56
    oy[k] = (float)((iy[k] + 0.5) * p.dy - LY / 2 - yv);
    oz[k] = (float)((iz[k] + 0.5) * p.dz - LZ / 2 + p.r0 - zv);
    actmask |= 1u << k;
    bx = min(bx, ix[k]); by = min(by, iy[k]); bz = min(bz, iz[k]);
    tx = max(tx, ix[k]); ty = max(ty, iy[k]); tz = max(tz, iz[k]);
  }
  if (!actmask) return;
  // the host guarantees (P-1)*pixel < cell size, so tx-bx, ty-by, tz-bz are 0 or 1
  const int nxu = 2 * DMAX + 1 + (tx - bx), nyu = 2 * DMAX + 1 + (ty - by);
  int dix[P], diy[P];
  float wz[P][WU], sz[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const bool act = (actmask >> k) & 1;
    dix[k] = act ? ix[k] - bx : 100;      // an inactive pixel gets zero weight everywhere
    diy[k] = iy[k] - by;
    sz[k] = 0.f;
#pragma unroll
    for (int c = 0; c < WU; ++c) {
      const int m = c - DMAX - (iz[k] - bz);
      const float t = m * fdz + oz[k];
      wz[k][c] = (act && m >= -DMAX && m <= DMAX) ? __expf(-t * t * inv_sig2) : 0.f;
      sz[k] += wz[k][c];
    }
  }
  float2 wz2[P][DMAX + 1];      // packed copy for the FFMA2 path (dead code otherwise)
#pragma unroll
  for (int k = 0; k < P; ++k)
#pragma unroll
    for (int j = 0; j <= DMAX; ++j) wz2[k][j] = make_float2(wz[k][2 * j], wz[k][2 * j + 1]);
  // y weights of the union window, once per thread, shared by the two field-group passes: [b][q][thread]
  __shared__ float s_wy[(2 * DMAX + 2) * P * 128];
  float* wy_s = s_wy + threadIdx.x;
  float sy[P];
  {
    const float fdy = (float)p.dy;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      sy[k] = 0.f;
#pragma unroll
      for (int b = 0; b < WU; ++b) {
        const int m = b - DMAX - diy[k];
        const float t = m * fdy + oy[k];
        const float w = (m >= -DMAX && m <= DMAX) ? __expf(-t * t * inv_sig2) : 0.f;
        wy_s[(b * P + k) * 128] = w;
        sy[k] += w;
      }
    }
  }
  // whole union window inside the slab: no index clamping, z offsets become immediates
  const bool interior = bx - DMAX - p.ix0 >= 0 && bx - DMAX - p.ix0 + nxu <= p.nxs && by - DMAX >= 0 &&
                        by - DMAX + nyu <= p.ny && bz - DMAX >= 0 && bz + DMAX + 1 < p.nz;
  float sx[P];
  const int NFI = p.rsd ? (p.dla ? 10 : 7) : 1;
  float d0[P], inv_sw[P];
  double eta[P], vel[P];
#pragma unroll
  for (int k = 0; k < P; ++k) { eta[k] = 0.0; vel[k] = 0.0; }
#define SMK_GATHER(NF_, ACC, SX, SY_UNUSED)                                                                              \
  if constexpr (F2) {                                                                                                    \
    float2 acc2_[NF_][P / 2];                                                                                            \
    if (interior) gather_multi2<DMAX, P, NF_, true>(p, fp, bx, by, bz, nxu, nyu, dix, ox, wy_s, wz2, inv_sig2, acc2_, SX); \
    else gather_multi2<DMAX, P, NF_, false>(p, fp, bx, by, bz, nxu, nyu, dix, ox, wy_s, wz2, inv_sig2, acc2_, SX);       \
    _Pragma("unroll") for (int f_ = 0; f_ < NF_; ++f_)                                                                   \
      _Pragma("unroll") for (int h_ = 0; h_ < P / 2; ++h_) {                                                             \
        ACC[f_][2 * h_] = acc2_[f_][h_].x;                                                                               \
        ACC[f_][2 * h_ + 1] = acc2_[f_][h_].y;                                                                           \
      }                                                                                                                  \
  } else {                                                                                                               \
    if (interior) gather_multi<DMAX, P, NF_, true>(p, fp, bx, by, bz, nxu, nyu, dix, ox, wy_s, wz, inv_sig2, ACC, SX);  \
    else gather_multi<DMAX, P, NF_, false>(p, fp, bx, by, bz, nxu, nyu, dix, ox, wy_s, wz, inv_sig2, ACC, SX);          \
  }
  if (NFI == 1) {
    float acc[1][P];
    const float* const fp[1] = {p.f[0]};
    SMK_GATHER(1, acc, sx, 0)
#pragma unroll
    for (int k = 0; k < P; ++k) {
      inv_sw[k] = ((actmask >> k) & 1) ? 1.0f / (sx[k] * sy[k] * sz[k]) : 0.f;
      d0[k] = acc[0][k];
    }
  } else {
    {   // group A: delta, eta_xx, eta_yy, eta_zz, eta_xy
      float acc[5][P];
      const float* const fp[5] = {p.f[0], p.f[1], p.f[2], p.f[3], p.f[4]};
      SMK_GATHER(5, acc, sx, 0)
#pragma unroll
      for (int k = 0; k < P; ++k) {
        inv_sw[k] = ((actmask >> k) & 1) ? 1.0f / (sx[k] * sy[k] * sz[k]) : 0.f;
        d0[k] = acc[0][k];
        double xv, yv, zv;
        pixel_xyz<DMAX, P>(p, q, min(i0 + k, p.npix - 1), xv, yv, zv);
        eta[k] = xv * (double)(acc[1][k] * inv_sw[k]) * xv + yv * (double)(acc[2][k] * inv_sw[k]) * yv +
                 zv * (double)(acc[3][k] * inv_sw[k]) * zv + 2 * xv * (double)(acc[4][k] * inv_sw[k]) * yv;
      }
    }
    float sx2[P];
    if (NFI == 10) {   // group B: eta_xz, eta_yz, vx, vy, vz
      float acc[5][P];
      const float* const fp[5] = {p.f[5], p.f[6], p.f[7], p.f[8], p.f[9]};
      SMK_GATHER(5, acc, sx2, 0)
#pragma unroll
      for (int k = 0; k < P; ++k) {
        double xv, yv, zv;
        pixel_xyz<DMAX, P>(p, q, min(i0 + k, p.npix - 1), xv, yv, zv);
        eta[k] += 2 * xv * (double)(acc[0][k] * inv_sw[k]) * zv + 2 * yv * (double)(acc[1][k] * inv_sw[k]) * zv;
        vel[k] = (double)(acc[2][k] * inv_sw[k]) * xv + (double)(acc[3][k] * inv_sw[k]) * yv +
                 (double)(acc[4][k] * inv_sw[k]) * zv;
      }
    } else {           // group B': eta_xz, eta_yz
      float acc[2][P];
      const float* const fp[2] = {p.f[5], p.f[6]};
      SMK_GATHER(2, acc, sx2, 0)
#pragma unroll
      for (int k = 0; k < P; ++k) {
        double xv, yv, zv;
        pixel_xyz<DMAX, P>(p, q, min(i0 + k, p.npix - 1), xv, yv, zv);
        eta[k] += 2 * xv * (double)(acc[0][k] * inv_sw[k]) * zv + 2 * yv * (double)(acc[1][k] * inv_sw[k]) * zv;
      }
    }
  }
#undef SMK_GATHER
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if (!((actmask >> k) & 1)) continue;
    const size_t o = (size_t)q * p.npix + i0 + k;
    double xv, yv, zv;
    pixel_xyz<DMAX, P>(p, q, i0 + k, xv, yv, zv);
    const double RR = xv * xv + yv * yv + zv * zv;
    const float dl = d0[k] * inv_sw[k];
    const float ep = NFI >= 7 ? (float)(eta[k] / RR) : 0.f;
    p.delta_l[o] = dl;
    if (p.eta_par) p.eta_par[o] = ep;
    if (p.vpar) p.vpar[o] = NFI == 10 ? (float)(vel[k] / sqrt(RR)) : 0.f;
    if (p.flux) store_flux(p, o, i0 + k, dl, ep);
  }
}

template <int P, int MINB, bool F2 = false>
static int launch_multi(const SkewerParams& p, cudaStream_t st) {
  const int NT = 128;
  int nchunk = (p.npix + NT * P - 1) / (NT * P);
  long long nblocks = (long long)nchunk * p.nqso;
  if (nblocks > 2147483647LL) { set_error("smk_skewers: too many quasars for one launch"); return SMK_ERR_ARG; }
  skewers_multi_kernel<3, P, MINB, F2><<<(unsigned)nblocks, NT, 0, st>>>(p, nchunk);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

// *fused (if given) tells whether the kernel that ran has the FGPA epilogue (the register-blocked one has)
int launch_skewers(const SkewerParams& p, int dmax, double pixel_step, cudaStream_t st, bool* fused = nullptr) {
  if (fused) *fused = false;
  if (p.nqso == 0 || p.npix == 0) return SMK_OK;
  const int NT = 128;
  // register-blocked kernel: valid while P consecutive pixels cannot cross two cell boundaries on any axis
  const double cell = fmin(p.dx, fmin(p.dy, p.dz));
  // default: 4 pixels per thread, packed FFMA2 arithmetic, 3 CTAs per SM (measured on B200 at 512 x 512 x 1536, whole
  // skewer stage: plain FFMA 18.4 ms, FFMA2 17.6-17.8 ms, FFMA2 + L1 prefetch of the next row 16.6 ms)
  int variant = 42;
  const char* env = getenv("SMK_SKEW_VARIANT");      // tuning knob: pixels per thread (2, 3, 4), 44 (4 px, 4 CTAs/SM),
                                                     // 42 / 442 (4 px, FFMA2 arithmetic, 3 / 4 CTAs/SM)
  if (env) variant = atoi(env);
  const int PB = (variant == 44 || variant == 42 || variant == 442) ? 4 : variant;
  const bool multi = (dmax == 3) && pixel_step > 0 && (PB - 1) * pixel_step < cell;
  if (multi) {
    if (fused) *fused = true;
    switch (variant) {
      case 2: return launch_multi<2, 5>(p, st);
      case 3: return launch_multi<3, 4>(p, st);
      case 44: return launch_multi<4, 4>(p, st);
      case 42: return launch_multi<4, 3, true>(p, st);     // packed FFMA2 arithmetic
      case 442: return launch_multi<4, 4, true>(p, st);
      default: return launch_multi<4, 3>(p, st);
    }
  }
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

static int skewers_impl(smk_ctx* ctx, const smk_geom* g, const float* const fields[10], int ix0, int nxs,
                        double xmin, double xmax, int rsd, int dla, int nqso, const double* qso_xyzr,
                        const int* npix_forest, const double* rvec, int npix, float* delta_l, float* eta_par,
                        float* vpar, const float* delta_s, const float* growthf, const float* fa, const float* fb,
                        const float* fc, float* flux) {
  using namespace smk;
  if (nqso == 0 || npix == 0) return SMK_OK;     // empty catalogue: nothing to do (empty tensors have null pointers)
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
  p.delta_s = delta_s; p.fg_G = growthf; p.fg_a = fa; p.fg_b = fb; p.fg_c = fc; p.flux = flux;
  { const char* e = getenv("SMK_SKEW_PF"); p.pf = e ? atoi(e) : 2; }
  { const char* e = getenv("SMK_SKEW_PFD"); p.pfd = e ? atoi(e) : 1; if (p.pfd < 1 || p.pfd > 7) p.pfd = 1; }
  // largest step between consecutive pixels of the grid (uniform 0.2 Mpc/h in the reference); decides whether the
  // register-blocked kernel may be used.  SMK_SKEWERS_SIMPLE=1 forces the one-pixel-per-thread kernel.
  double step = g->pixel_step;
  const char* env = getenv("SMK_SKEWERS_SIMPLE");
  if (env && env[0] == '1') step = 0.0;
  bool fused = false;
  int rc = launch_skewers(p, g->dmax, step, smk_ctx_stream(ctx), &fused);
  if (rc != SMK_OK || !flux || fused) return rc;
  // the one-pixel-per-thread kernel has no epilogue: FGPA as a separate pass over the rows
  return smk_fgpa(ctx, nqso, npix, delta_l, delta_s, eta_par, growthf, fa, fb, fc, flux);
}

extern "C" int smk_skewers(smk_ctx* ctx, const smk_geom* g, const float* const fields[10], int ix0, int nxs,
                           double xmin, double xmax, int rsd, int dla, int nqso, const double* qso_xyzr,
                           const int* npix_forest, const double* rvec, int npix, float* delta_l, float* eta_par,
                           float* vpar) {
  return skewers_impl(ctx, g, fields, ix0, nxs, xmin, xmax, rsd, dla, nqso, qso_xyzr, npix_forest, rvec, npix, delta_l,
                      eta_par, vpar, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int smk_skewers_fgpa(smk_ctx* ctx, const smk_geom* g, const float* const fields[10], int ix0, int nxs,
                                double xmin, double xmax, int rsd, int dla, int nqso, const double* qso_xyzr,
                                const int* npix_forest, const double* rvec, int npix, float* delta_l, float* eta_par,
                                float* vpar, const float* delta_s, const float* growthf, const float* a, const float* b,
                                const float* c, float* flux) {
  if (!flux || !growthf || !a || !b || (rsd && !c)) {
    smk::set_error("smk_skewers_fgpa: flux and the FGPA parameter vectors are required");
    return SMK_ERR_ARG;
  }
  return skewers_impl(ctx, g, fields, ix0, nxs, xmin, xmax, rsd, dla, nqso, qso_xyzr, npix_forest, rvec, npix, delta_l,
                      eta_par, vpar, delta_s, growthf, a, b, c, flux);
}
