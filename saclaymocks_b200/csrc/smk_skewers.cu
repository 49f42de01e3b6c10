// Skewer gather: batched ReadSpec (bin/make_spectra.py:90-139 with ComputeWeight :40-62 and
// computeRho :66-68) for every quasar of one x-slab in one launch.
//
// For each line-of-sight pixel: the cell that contains it (fp64 index arithmetic, exactly the
// reference's expressions), a (2*dmax+1)^3 Gaussian-weighted average of up to 10 fields, the
// eta_par = x_i x_j eta_ij / r^2 and v_par = v.r/|r| contractions.  The 343 reference weights
// exp(-|dr|^2 / 2 DX^2) factorise as wx*wy*wz, so 21 exponentials are evaluated per pixel instead
// of 343; sums run in float32 (|error| ~1e-6 of the field rms, far inside the 1e-5 tolerance on F).
//
// Two gather kernels share the per-pixel set-up and the epilogue:
//   * skewers_tma_kernel (default): every warp owns 32*P consecutive pixels of one sightline and brings the box of
//     cells their windows can touch into shared memory with one 3-D TMA load per field (cp.async.bulk.tensor, one
//     mbarrier per warp), two fields per stage; all window reads are then shared-memory loads and the z contraction runs
//     on packed FFMA2 with the horizontal add deferred to the end of the field.
//   * skewers_multi_kernel: the same arithmetic on global memory through L1 with index clamping -- the path of the
//     segments the TMA kernel hands back (windows that reach over a box edge, segments too oblique for the staged box).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

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
  int nseg;                // segments of 32*P pixels per sightline (list mode of skewers_multi_kernel, TMA kernel)
  const int* list;         // skewers_multi_kernel: null = every segment, else [0] = count, [1..] = q * nseg + segment
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

// exp(x) for the Gaussian window weights: ex2.approx(x log2 e), what __expf evaluates for arguments that cannot
// underflow (here x >= -6.2), without its range check and the branches that come with it
__device__ __forceinline__ float exp_weight(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
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

// ---- register-blocked gathers: each thread owns P consecutive pixels of one sightline and walks the UNION of their
// (2*DMAX+1)^3 windows once, so that every field value loaded is used by P pixels.  P consecutive pixels span
// (P-1)*pixel < one cell, hence the union is at most one cell wider per axis; a pixel's weight is zero outside its
// own window, which keeps the result identical to the reference's truncated Gaussian sum.
constexpr int DMAX = 3;
constexpr int WU = 2 * DMAX + 2;          // union window width (cells)

template <int P>
__device__ __forceinline__ void pixel_xyz(const SkewerParams& p, int q, int i, double& xv, double& yv, double& zv) {
  const double R = p.qso[4 * q + 3], r = p.rvec[i];
  xv = r * p.qso[4 * q] / R;                 // make_spectra.py:443-452 (same operation order)
  yv = r * p.qso[4 * q + 1] / R;
  zv = r * p.qso[4 * q + 2] / R;
}

// Per-thread state of P consecutive pixels: cells, in-cell offsets, the union window and the separable weights along
// y and z (x weights are re-evaluated inside the walk, one exponential per pixel and window plane).  Pixels are packed
// in pairs (2h, 2h+1) wherever a weight multiplies both with one FMUL2.
template <int P>
struct PixelGroup {
  static_assert(P % 2 == 0 && P <= 4, "pixels are processed in pairs; xlo packs one byte per pixel");
  unsigned actmask;            // pixels of this thread that are owned by the slab and inside the forest
  int bx, by, bz;              // lowest cell of the union of the windows' centres
  int nxu, nyu;                // union window extent along x / y (7 or 8)
  int tz;                      // highest centre cell along z
  int dix[P];                  // centre cell of pixel k minus bx (100 for an inactive pixel: zero weight everywhere)
  float ox[P];                 // cell centre - pixel along x (set-up only; the walk uses tx0 / xlo)
  float tx0[P];                // x offset of window plane a = 0 from pixel k: plane a sits at tx0 + a * dx
  unsigned xlo;                // first window plane of pixel k in bits [8k, 8k+8) (its window is planes xlo .. xlo + 6)
  float2 wy2[P / 2][WU];       // y weights over the union window of the pixel pair (2h, 2h+1)
  float2 wz2[P][WU / 2];       // z weights over the union window, packed in pairs of cells
  float syz[P];                // (sum of y weights) * (sum of z weights)
};

// Set-up of make_spectra.py:47-62 for the P pixels starting at i0 of sightline q.  Writes the sentinels of
// make_spectra.py:99-101 for owned pixels beyond the forest.  Returns false when no pixel is left to gather.
template <int P>
__device__ __forceinline__ bool pixel_group_setup(const SkewerParams& p, int q, int i0, PixelGroup<P>& g) {
  const int nfor = p.npix_forest[q];
  const double LX = p.dx * p.nx, LY = p.dy * p.ny, LZ = p.dz * p.nz;
  const double inv_dx = 1.0 / p.dx, inv_dy = 1.0 / p.dy, inv_dz = 1.0 / p.dz;
  const double R = p.qso[4 * q + 3];
  const double ux = p.qso[4 * q] / R, uy = p.qso[4 * q + 1] / R, uz = p.qso[4 * q + 2] / R;
  const float inv_sig2 = (float)(1.0 / (2.0 * p.dx * p.dx));
  const float fdy = (float)p.dy, fdz = (float)p.dz;
  g.actmask = 0;
  int ix[P], iy[P], iz[P];
  float oy[P], oz[P];
  int bx = 1 << 30, by = 1 << 30, bz = 1 << 30, tx = -(1 << 30), ty = -(1 << 30), tz = -(1 << 30);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int i = i0 + k;
    ix[k] = iy[k] = iz[k] = 0;
    g.ox[k] = oy[k] = oz[k] = 0.f;
    if (i >= p.npix) continue;
    // Pixel position and cell: the reference's expressions are r * X / R and int((x + L/2) / D) (make_spectra.py:443-452,
    // :47-49).  They are first evaluated with reciprocals (r * (X / R), (x + L/2) * (1 / D): no float64 division per
    // pixel); the results can differ from the reference's by a few ulp, which changes the cell index or the slab that
    // owns the pixel only within ~1e-13 of a cell / slab boundary -- anything within 1e-9 of one is re-evaluated with
    // the reference's own operations, so the integer work stays exactly the reference's.
    const double r = p.rvec[i];
    double xv = r * ux, yv = r * uy, zv = r * uz;
    double cx = (xv + LX / 2) * inv_dx, cy = (yv + LY / 2) * inv_dy, cz = (zv + LZ / 2 - p.r0) * inv_dz;
    const double fx = cx - floor(cx), fy = cy - floor(cy), fz = cz - floor(cz);
    const double EPS = 1e-9;
    if (fx < EPS || fx > 1 - EPS || fy < EPS || fy > 1 - EPS || fz < EPS || fz > 1 - EPS ||
        fabs(xv - p.xmin) < EPS * LX || fabs(xv - p.xmax) < EPS * LX) {
      pixel_xyz<P>(p, q, i, xv, yv, zv);
      cx = (xv + LX / 2) / p.dx; cy = (yv + LY / 2) / p.dy; cz = (zv + LZ / 2 - p.r0) / p.dz;
    }
    if (!(xv > p.xmin) || !(xv <= p.xmax)) continue;            // owned by another slab
    if (i >= nfor) {                                            // make_spectra.py:99-101
      const size_t o = (size_t)q * p.npix + i;
      p.delta_l[o] = -1000000.f;
      if (p.eta_par) p.eta_par[o] = 0.f;
      if (p.vpar) p.vpar[o] = 0.f;
      if (p.flux) store_flux(p, o, i, -1000000.f, 0.f);
      continue;
    }
    ix[k] = (int)cx;                                            // int() truncates toward zero
    iy[k] = (int)cy;
    iz[k] = (int)cz;
    g.ox[k] = (float)((ix[k] + 0.5) * p.dx - LX / 2 - xv);      // cell centre - pixel, make_spectra.py:54-56
    oy[k] = (float)((iy[k] + 0.5) * p.dy - LY / 2 - yv);
    oz[k] = (float)((iz[k] + 0.5) * p.dz - LZ / 2 + p.r0 - zv);
    g.actmask |= 1u << k;
    bx = min(bx, ix[k]); by = min(by, iy[k]); bz = min(bz, iz[k]);
    tx = max(tx, ix[k]); ty = max(ty, iy[k]); tz = max(tz, iz[k]);
  }
  if (!g.actmask) return false;
  // the host guarantees (P-1)*pixel < cell size, so tx-bx, ty-by, tz-bz are 0 or 1
  g.bx = bx; g.by = by; g.bz = bz; g.tz = tz;
  g.nxu = 2 * DMAX + 1 + (tx - bx);
  g.nyu = 2 * DMAX + 1 + (ty - by);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const bool act = (g.actmask >> k) & 1;
    g.dix[k] = act ? ix[k] - bx : 100;      // an inactive pixel gets zero weight everywhere
    g.tx0[k] = (float)(-DMAX - g.dix[k]) * (float)p.dx + g.ox[k];
    if (k == 0) g.xlo = 0;
    g.xlo |= (unsigned)g.dix[k] << (8 * k);
    const int diy = iy[k] - by, diz = iz[k] - bz;
    float sy = 0.f, sz = 0.f;
    float wz[WU];
#pragma unroll
    for (int c = 0; c < WU; ++c) {
      const int mz = c - DMAX - diz, my = c - DMAX - diy;
      const float tzz = mz * fdz + oz[k], tyy = my * fdy + oy[k];
      wz[c] = (act && mz >= -DMAX && mz <= DMAX) ? exp_weight(-tzz * tzz * inv_sig2) : 0.f;
      const float wyc = (act && my >= -DMAX && my <= DMAX) ? exp_weight(-tyy * tyy * inv_sig2) : 0.f;
      if (k & 1) g.wy2[k / 2][c].y = wyc; else g.wy2[k / 2][c].x = wyc;
      sz += wz[c];
      sy += wyc;
    }
#pragma unroll
    for (int j = 0; j < WU / 2; ++j) g.wz2[k][j] = make_float2(wz[2 * j], wz[2 * j + 1]);
    g.syz[k] = sy * sz;
  }
  return true;
}

// a lane without pixels in a warp that still has some: zero weights everywhere
template <int P>
__device__ __forceinline__ void pixel_group_idle(PixelGroup<P>& g) {
  g.actmask = 0;
  g.bx = g.by = g.bz = g.tz = 0;
  g.nxu = g.nyu = 2 * DMAX + 1;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    g.dix[k] = 100; g.ox[k] = 0.f; g.syz[k] = 1.f; g.tx0[k] = 0.f;
    if (k == 0) g.xlo = 0;
    g.xlo |= 100u << (8 * k);
#pragma unroll
    for (int j = 0; j < WU / 2; ++j) g.wz2[k][j] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int h = 0; h < P / 2; ++h)
#pragma unroll
    for (int c = 0; c < WU; ++c) g.wy2[h][c] = make_float2(0.f, 0.f);
}

// whole union window inside the slab (no index clamping needed)
template <int P>
__device__ __forceinline__ bool group_interior(const SkewerParams& p, const PixelGroup<P>& g) {
  return g.bx - DMAX - p.ix0 >= 0 && g.bx - DMAX - p.ix0 + g.nxu <= p.nxs && g.by - DMAX >= 0 &&
         g.by - DMAX + g.nyu <= p.ny && g.bz - DMAX >= 0 && g.tz + DMAX < p.nz;
}

// Contraction of the gathered fields along the line of sight (make_spectra.py:116-127).  Every pixel of a sightline
// has the same direction u = (X, Y, Z) / R_QSO, and the reference's x_i eta_ij x_j / r^2 and v_i x_i / r are
// homogeneous of degree zero in the pixel position, so they are evaluated with u once per sightline:
// eta_par = u_i eta_ij u_j / |u|^2, v_par = v_i u_i / |u| (float64 like the reference's products).
// The six / three coefficients are evaluated in float64 from the catalogue's float64 (X, Y, Z, R) and rounded to
// float32 once; the sums of six (three) float32 terms then carry ~1e-7 of the largest eta_ij (v_i), against the
// 5e-6 the rows are compared at.
struct Direction {
  float ux, uy, uz, inv_u2, inv_u;
};
__device__ __forceinline__ Direction sightline_direction(const SkewerParams& p, int q) {
  const double R = p.qso[4 * q + 3];
  const double ux = p.qso[4 * q] / R, uy = p.qso[4 * q + 1] / R, uz = p.qso[4 * q + 2] / R;
  const double u2 = ux * ux + uy * uy + uz * uz;
  Direction d;
  d.ux = (float)ux; d.uy = (float)uy; d.uz = (float)uz;
  d.inv_u2 = (float)(1.0 / u2);
  d.inv_u = (float)(1.0 / sqrt(u2));
  return d;
}
// coefficient of field f (delta, eta_xx, eta_yy, eta_zz, eta_xy, eta_xz, eta_yz, vx, vy, vz) in eta_par / v_par
__device__ __forceinline__ float field_coefficient(const Direction& d, int f) {
  switch (f) {
    case 1: return d.ux * d.ux * d.inv_u2;
    case 2: return d.uy * d.uy * d.inv_u2;
    case 3: return d.uz * d.uz * d.inv_u2;
    case 4: return 2.f * d.ux * d.uy * d.inv_u2;
    case 5: return 2.f * d.ux * d.uz * d.inv_u2;
    case 6: return 2.f * d.uy * d.uz * d.inv_u2;
    case 7: return d.ux * d.inv_u;
    case 8: return d.uy * d.inv_u;
    case 9: return d.uz * d.inv_u;
    default: return 1.f;
  }
}

template <int P>
struct PixelResult {
  float d0[P], eta[P], vel[P];
};

// fold field f of one walk (two partial sums per pixel) into the running results
template <int P>
__device__ __forceinline__ void fold_field(int f, float coef, const float2 (&acc2)[P], const float (&inv_sw)[P],
                                           PixelResult<P>& r) {
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const float val = (acc2[k].x + acc2[k].y) * inv_sw[k];
    if (f == 0) r.d0[k] = val;
    else if (f <= 6) r.eta[k] = fmaf(coef, val, r.eta[k]);
    else r.vel[k] = fmaf(coef, val, r.vel[k]);
  }
}

template <int P>
__device__ __forceinline__ void result_store(const SkewerParams& p, int q, int i0, unsigned actmask, int nf,
                                             const PixelResult<P>& r) {
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if (!((actmask >> k) & 1)) continue;
    const size_t o = (size_t)q * p.npix + i0 + k;
    const float dl = r.d0[k];
    const float ep = nf >= 7 ? r.eta[k] : 0.f;
    p.delta_l[o] = dl;
    if (p.eta_par) p.eta_par[o] = ep;
    if (p.vpar) p.vpar[o] = nf == 10 ? r.vel[k] : 0.f;
    if (p.flux) store_flux(p, o, i0 + k, dl, ep);
  }
}

// sum of the x weights of every pixel over its window (the third factor of the weight normalisation)
template <int P>
__device__ __forceinline__ void sum_x_weights(const SkewerParams& p, const PixelGroup<P>& g, float (&sx)[P]) {
  const float fdx = (float)p.dx;
  const float inv_sig2 = (float)(1.0 / (2.0 * p.dx * p.dx));
#pragma unroll
  for (int k = 0; k < P; ++k) sx[k] = 0.f;
#pragma unroll
  for (int a = 0; a < WU; ++a) {
    const float ta = (float)a * fdx;          // the same expressions as walk_window: the weights summed are the weights used
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const unsigned m = (unsigned)a - ((g.xlo >> (8 * k)) & 0xffu);
      const float t = ta + g.tx0[k];
      sx[k] += (a < g.nxu && m <= 2u * DMAX) ? exp_weight(-t * t * inv_sig2) : 0.f;
    }
  }
}

// One walk over the union window for NF fields.  load(f, a, b, r2) fetches the 8 consecutive z cells of window row
// (a, b) of field f as four pairs.  Rows 0..6 of every window plane are walked unconditionally (one straight line of
// code: the loads of a row are issued under the arithmetic of the one before), row 7 only when some pixel of the warp
// has it in its window (nyu == 8).  z contraction on pairs of cells (4 FMUL2/FFMA2 per pixel and row); rows are
// accumulated per plane with the y weight as the broadcast operand of one FFMA2, the plane then enters the running sum
// with its x weight -- no per-row weight product.  The two halves of every accumulator are added at the end of the walk.
template <int P, int NF, class Load>
__device__ __forceinline__ void walk_window(const SkewerParams& p, const PixelGroup<P>& g, int nxu, int nyu, Load load,
                                            float2 (&acc2)[NF][P]) {
  const float fdx = (float)p.dx;
  const float inv_sig2 = (float)(1.0 / (2.0 * p.dx * p.dx));
#pragma unroll
  for (int f = 0; f < NF; ++f)
#pragma unroll
    for (int k = 0; k < P; ++k) acc2[f][k] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int a = 0; a < nxu; ++a) {
    float wxa[P];
    const float ta = (float)a * fdx;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const unsigned m = (unsigned)a - ((g.xlo >> (8 * k)) & 0xffu);      // plane a is inside pixel k's window iff m <= 6
      const float t = ta + g.tx0[k];
      wxa[k] = m <= 2u * DMAX ? exp_weight(-t * t * inv_sig2) : 0.f;
    }
    float2 pl2[NF][P];                         // this plane: sum over its rows of wy * (z contraction)
    auto row = [&](int b, bool first) {
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        float2 r2[WU / 2];
        load(f, a, b, r2);
#pragma unroll
        for (int k = 0; k < P; ++k) {
          float2 s2 = fmul2(g.wz2[k][0], r2[0]);
#pragma unroll
          for (int j = 1; j < WU / 2; ++j) s2 = ffma2(g.wz2[k][j], r2[j], s2);
          const float w = (k & 1) ? g.wy2[k / 2][b].y : g.wy2[k / 2][b].x;
          pl2[f][k] = first ? fmul2(make_float2(w, w), s2) : ffma2(make_float2(w, w), s2, pl2[f][k]);
        }
      }
    };
#pragma unroll
    for (int b = 0; b < WU - 1; ++b) row(b, b == 0);
    if (nyu == WU) row(WU - 1, false);
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
      for (int k = 0; k < P; ++k) acc2[f][k] = ffma2(make_float2(wxa[k], wxa[k]), pl2[f][k], acc2[f][k]);
  }
}

// ---------------------------------------------------------------- global-memory walk (clamped indices)
// All segments (list == null: one CTA per 4 segments of a sightline) or the segments the staged kernel handed back
// (list mode: a fixed grid walks list[1 .. list[0]]).
template <int P>
__device__ __forceinline__ void global_segment(const SkewerParams& p, int q, int seg) {
  const int lane = threadIdx.x & 31;
  const int i0 = (seg * 32 + lane) * P;
  if (i0 >= p.npix || seg * 32 * P >= p.npix_forest[q]) return;      // beyond the forest: skewers_sentinel_kernel's pixels
  PixelGroup<P> g;
  if (!pixel_group_setup<P>(p, q, i0, g)) return;
  const int nf = p.rsd ? (p.dla ? 10 : 7) : 1;
  const Direction dir = sightline_direction(p, q);
  const unsigned plane = (unsigned)p.ny * (unsigned)p.nz;
  PixelResult<P> r;
  float inv_sw[P], sx[P];
  sum_x_weights<P>(p, g, sx);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    r.d0[k] = 0.f; r.eta[k] = 0.f; r.vel[k] = 0.f;
    inv_sw[k] = ((g.actmask >> k) & 1) ? 1.0f / (sx[k] * g.syz[k]) : 0.f;
  }
  // window indices clamped to the slab (documented deviation from the reference's unchecked gather, include/smk.h;
  // a no-op for a window inside the slab), rows fetched through L1
  int lz[WU];
#pragma unroll
  for (int c = 0; c < WU; ++c) lz[c] = min(max(g.bz - DMAX + c, 0), p.nz - 1);
#pragma unroll 1
  for (int f = 0; f < nf; ++f) {
    const float* __restrict__ field = p.f[f];
    float2 acc2[1][P];
    walk_window<P, 1>(p, g, g.nxu, g.nyu,
                      [&](int, int a, int b, float2 (&r2)[WU / 2]) {
                        const int la = min(max(g.bx - DMAX + a - p.ix0, 0), p.nxs - 1);
                        const int lb = min(max(g.by - DMAX + b, 0), p.ny - 1);
                        const float* __restrict__ row = field + ((size_t)la * plane + (unsigned)lb * (unsigned)p.nz);
#pragma unroll
                        for (int j = 0; j < WU / 2; ++j) r2[j] = make_float2(__ldg(row + lz[2 * j]), __ldg(row + lz[2 * j + 1]));
                      },
                      acc2);
    fold_field<P>(f, field_coefficient(dir, f), acc2[0], inv_sw, r);
  }
  result_store<P>(p, q, i0, g.actmask, nf, r);
}

template <int P>
__global__ void __launch_bounds__(128, 2) skewers_multi_kernel(const __grid_constant__ SkewerParams p) {
  const int warp = threadIdx.x >> 5;
  if (p.list == nullptr) {
    const long long s = (long long)blockIdx.x * 4 + warp;
    if (s >= (long long)p.nqso * p.nseg) return;
    const int q = (int)(s / p.nseg);
    global_segment<P>(p, q, (int)(s - (long long)q * p.nseg));
  } else {
    const int n = p.list[0];
#pragma unroll 1
    for (int e = blockIdx.x * 4 + warp; e < n; e += gridDim.x * 4) {
      const int s = p.list[1 + e];
      const int q = s / p.nseg;
      global_segment<P>(p, q, s - q * p.nseg);
    }
  }
}

// ---------------------------------------------------------------- staged walk: TMA box loads into shared memory
struct alignas(64) SkewerTmaParams {
  CUtensorMap map[10];     // one 3-D tiled map per field: dims (z, y, x-planes of the slab), box (zl, yw, xw)
  SkewerParams p;
  int xw, yw, zl;          // box extent in cells
  int box_elems;           // floats per staged box (xw * yw * zl rounded up to 128 B)
  int* handback;           // [0] = count, [1..] = q * nseg + segment of the segments left to skewers_multi_kernel
  int* work;               // [0] = entries, [1] = fetch cursor, [2..] = q * nseg + segment (skewers_worklist_kernel)
};

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// 3-D box load: coordinates (z, y, x) in elements, z a multiple of 4 (16 bytes); out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// 8 consecutive floats of shared memory at a 32-bit shared address (4-byte aligned): explicit ld.shared with
// immediate offsets, so that the address arithmetic stays one add per row
__device__ __forceinline__ void lds_row8(uint32_t addr, float2 (&r2)[WU / 2]) {
  static_assert(WU == 8, "eight cells per window row");
  asm volatile("ld.shared.f32 %0, [%8];\n\tld.shared.f32 %1, [%8+4];\n\tld.shared.f32 %2, [%8+8];\n\t"
               "ld.shared.f32 %3, [%8+12];\n\tld.shared.f32 %4, [%8+16];\n\tld.shared.f32 %5, [%8+20];\n\t"
               "ld.shared.f32 %6, [%8+24];\n\tld.shared.f32 %7, [%8+28];"
               : "=f"(r2[0].x), "=f"(r2[0].y), "=f"(r2[1].x), "=f"(r2[1].y), "=f"(r2[2].x), "=f"(r2[2].y), "=f"(r2[3].x),
                 "=f"(r2[3].y)
               : "r"(addr));
}

// Owned pixels of segments that lie entirely beyond the forest (make_spectra.py:99-101: delta_l = -1e6, eta_par =
// v_par = 0, and F of that): the gather kernels leave such segments at once, this kernel writes their sentinels.
// SEG = pixels per segment of the gather that follows (the segment that holds the end of the forest is the gather's).
__global__ void __launch_bounds__(256) skewers_sentinel_kernel(const __grid_constant__ SkewerParams p, int seg_pixels,
                                                               int chunks) {
  const int q = blockIdx.x / chunks;
  const int i = (blockIdx.x - q * chunks) * 256 + threadIdx.x;
  const int nfor = p.npix_forest[q];
  const int start = (nfor + seg_pixels - 1) / seg_pixels * seg_pixels;
  if (i < start || i >= p.npix) return;
  const double xv = p.rvec[i] * p.qso[4 * q] / p.qso[4 * q + 3];       // make_spectra.py:443-448
  if (!(xv > p.xmin) || !(xv <= p.xmax)) return;
  const size_t o = (size_t)q * p.npix + i;
  p.delta_l[o] = -1000000.f;
  if (p.eta_par) p.eta_par[o] = 0.f;
  if (p.vpar) p.vpar[o] = 0.f;
  if (p.flux) store_flux(p, o, i, -1000000.f, 0.f);
}

// Work list of the staged gather: the segments (q * nseg + segment) that hold forest pixels this slab can own, the
// segments of a sightline next to each other.  work[0] = number of entries, work[1] = fetch cursor of the gather's
// persistent warps, entries from work[2].  One thread per sightline.
__global__ void __launch_bounds__(256) skewers_worklist_kernel(const __grid_constant__ SkewerParams p, int seg_pixels,
                                                               int* __restrict__ work) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= p.nqso) return;
  const int nfor = min(p.npix_forest[q], p.npix);
  const double X = p.qso[4 * q], R = p.qso[4 * q + 3];
  // a segment can own pixels here iff the x range of its forest pixels meets (xmin, xmax] (x is monotonic along the
  // sightline); the margin keeps a segment whose end pixel sits on the slab boundary to the last bit
  const double margin = 1e-9 * (fabs(p.xmin) + fabs(p.xmax) + 1.0);
  auto live = [&](int s) {
    const int ia = s * seg_pixels, ib = min(ia + seg_pixels, nfor) - 1;
    const double xa = p.rvec[ia] * X / R, xb = p.rvec[ib] * X / R;
    return fmax(xa, xb) > p.xmin - margin && fmin(xa, xb) <= p.xmax + margin;
  };
  const int ns = (nfor + seg_pixels - 1) / seg_pixels;
  int n = 0;
  for (int s = 0; s < ns; ++s) n += live(s) ? 1 : 0;
  if (n == 0) return;
  int at = 2 + atomicAdd(work, n);
  for (int s = 0; s < ns; ++s)
    if (live(s)) work[at++] = q * p.nseg + s;
}

// Persistent gather: a grid of single-warp CTAs, as many as fit the SMs, each fetching segments of 32*P pixels from the
// work list until it is empty.  Every warp has its own boxes and mbarriers; NFG fields are staged and walked together,
// and the boxes of the next stage are in flight while the current ones are walked (two buffers, one mbarrier each).
template <int P, int NFG, int MINB>
__global__ void __launch_bounds__(32, MINB) skewers_tma_kernel(const __grid_constant__ SkewerTmaParams t) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const SkewerParams& p = t.p;
  const int lane = threadIdx.x;
  const uint32_t stage_bytes = (uint32_t)NFG * t.box_elems * 4u;
  const uint32_t box = smem_u32(smraw);
  const uint32_t bar = box + 2u * stage_bytes;
  if (lane == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;              // bit b: phase of buffer b's mbarrier (carried from segment to segment)
  const int nwork = t.work[0];
  const int nf = p.rsd ? (p.dla ? 10 : 7) : 1;
  const uint32_t row_bytes = (uint32_t)t.zl * 4u, plane_bytes = (uint32_t)(t.yw * t.zl) * 4u;
  const uint32_t field_bytes = (uint32_t)t.box_elems * 4u;
  const uint32_t box_bytes = (uint32_t)(t.xw * t.yw * t.zl) * (uint32_t)sizeof(float);
#pragma unroll 1
  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(t.work + 1, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= nwork) break;
    const int s = t.work[2 + w];
    const int q = s / p.nseg, seg = s - q * p.nseg;
    const int i0 = (seg * 32 + lane) * P;
    PixelGroup<P> g;
    const bool act = (i0 < p.npix) && pixel_group_setup<P>(p, q, i0, g);
    if (!__any_sync(0xffffffffu, act)) continue;
    if (!act) pixel_group_idle<P>(g);
    // the box of the warp: lowest window corner of any live lane; every live lane's 8 x 8 x 8 read must fit into it
    // and every window must lie inside the slab (TMA fills what is outside the tensor with zeros, which only ever meet
    // the zero weights of the 8th row / plane / cell)
    const int BIG = 1 << 30;
    const int x0 = __reduce_min_sync(0xffffffffu, act ? g.bx : BIG), x1 = __reduce_max_sync(0xffffffffu, act ? g.bx : -BIG);
    const int y0 = __reduce_min_sync(0xffffffffu, act ? g.by : BIG), y1 = __reduce_max_sync(0xffffffffu, act ? g.by : -BIG);
    const int z0 = __reduce_min_sync(0xffffffffu, act ? g.bz : BIG), z1 = __reduce_max_sync(0xffffffffu, act ? g.bz : -BIG);
    // TMA wants the box to start on a 16-byte boundary along the contiguous axis (measured: an innermost coordinate
    // that is not a multiple of 4 floats raises "illegal instruction", tools/micro/tma_box_test.cu): round z down
    const int zb = (z0 - DMAX) & ~3;
    const bool fits = (x1 - x0 + WU <= t.xw) && (y1 - y0 + WU <= t.yw) && (z1 - DMAX + WU - zb <= t.zl);
    const bool inside = __all_sync(0xffffffffu, !act || group_interior<P>(p, g));
    if (!fits || !inside) {         // hand the segment back to the global-memory kernel
      if (lane == 0) t.handback[1 + atomicAdd(t.handback, 1)] = s;
      continue;
    }
    const int nxu = __reduce_max_sync(0xffffffffu, g.nxu), nyu = __reduce_max_sync(0xffffffffu, g.nyu);
    const uint32_t mine = act ? (uint32_t)(((g.bx - x0) * t.yw + (g.by - y0)) * t.zl + (g.bz - DMAX - zb)) * 4u : 0u;
    // a stage = the boxes of fields [f0, f0 + NFG) (a last stage that is not full loads its last field again)
    auto issue = [&](int f0, int buf) {
      if (lane == 0) {
        mbar_expect_tx(bar + 8u * buf, NFG * box_bytes);
#pragma unroll
        for (int f = 0; f < NFG; ++f)
          tma_load_3d(box + buf * stage_bytes + f * field_bytes, &t.map[min(f0 + f, nf - 1)], zb, y0 - DMAX,
                      x0 - DMAX - p.ix0, bar + 8u * buf);
      }
    };
    issue(0, 0);                    // in flight under the rest of the set-up
    const Direction dir = sightline_direction(p, q);
    PixelResult<P> r;
    float inv_sw[P], sx[P];
    sum_x_weights<P>(p, g, sx);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      r.d0[k] = 0.f; r.eta[k] = 0.f; r.vel[k] = 0.f;
      inv_sw[k] = ((g.actmask >> k) & 1) ? 1.0f / (sx[k] * g.syz[k]) : 0.f;
    }
    // ONE copy of the walk in the instruction stream: the warps of an SM are at different stages, and separate copies
    // per stage thrashed the instruction cache (profiles/README.md)
#pragma unroll 1
    for (int f0 = 0, it = 0; f0 < nf; f0 += NFG, ++it) {
      const int buf = it & 1;
      // the other buffer was last read by the walk before this one (every lane passed its __syncwarp): refill it now
      if (f0 + NFG < nf) issue(f0 + NFG, buf ^ 1);
      mbar_wait(bar + 8u * buf, (parity >> buf) & 1u);
      parity ^= 1u << buf;
      const uint32_t base = box + buf * stage_bytes + mine;
      float2 acc2[NFG][P];
      walk_window<P, NFG>(p, g, nxu, nyu,
                          [&](int f, int a, int b, float2 (&r2)[WU / 2]) {
                            lds_row8(base + f * field_bytes + a * plane_bytes + b * row_bytes, r2);
                          },
                          acc2);
      __syncwarp();                 // every lane is done with this buffer before the stage after next overwrites it
#pragma unroll
      for (int f = 0; f < NFG; ++f)
        if (f == 0 || f0 + f < nf) fold_field<P>(f0 + f, field_coefficient(dir, f0 + f), acc2[f], inv_sw, r);
    }
    if (act) result_store<P>(p, q, i0, g.actmask, nf, r);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point table (libsmk.so does not link libcuda)
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// bookkeeping of the calling thread's last gather (smk_skewers_stats)
struct LastGather { const int* handback = nullptr; long long nsegs = 0; int xw = 0, yw = 0, zl = 0; cudaStream_t st = nullptr; };
static thread_local LastGather g_last;

constexpr int SKEW_BOX_MAX = 16;          // largest staged box extent along x / y (cells)

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static int launch_sentinels(const SkewerParams& p, int seg_pixels, cudaStream_t st) {
  const int chunks = (p.npix + 255) / 256;
  const long long nblocks = (long long)p.nqso * chunks;
  if (nblocks > 2147483647LL) { set_error("smk_skewers: too many quasars for one launch"); return SMK_ERR_ARG; }
  skewers_sentinel_kernel<<<(unsigned)nblocks, 256, 0, st>>>(p, seg_pixels, chunks);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

// Staged launch: one warp per CTA, every warp stages its own boxes, so the number of warps an SM holds is set by
// shared memory (and registers) alone.  Returns SMK_ERR_UNSUPPORTED (without touching the error string) when the
// inputs do not meet TMA's alignment rules, in which case the caller uses the global-memory kernel for everything.
template <int P, int NFG, int MINB>
static int launch_staged(smk_ctx* ctx, const smk_geom* g, SkewerParams p, cudaStream_t st) {
  EncodeTiledFn encode = encode_tiled_fn();
  const int nf = p.rsd ? (p.dla ? 10 : 7) : 1;
  if (!encode || p.nz % 4 != 0) return SMK_ERR_UNSUPPORTED;
  for (int f = 0; f < nf; ++f)
    if ((uintptr_t)p.f[f] & 15) return SMK_ERR_UNSUPPORTED;
  p.nseg = (p.npix + 32 * P - 1) / (32 * P);
  p.list = nullptr;
  SkewerTmaParams t{};
  t.p = p;
  // extent of the box a segment of 32*P pixels can need: cells crossed along the axis + the 8-cell union window
  const double lseg = (32 * P - 1) * g->pixel_step;
  auto extent = [&](double dir, double d) { return (int)floor(lseg * fmin(fabs(dir), 1.0) / d) + 1 + WU; };
  const double dirx = g->dir_x_max > 0 ? g->dir_x_max : 0.25, diry = g->dir_y_max > 0 ? g->dir_y_max : 0.25;
  t.xw = extent(dirx, g->dx) < SKEW_BOX_MAX ? extent(dirx, g->dx) : SKEW_BOX_MAX;
  t.yw = extent(diry, g->dy) < SKEW_BOX_MAX ? extent(diry, g->dy) : SKEW_BOX_MAX;
  t.zl = round_up(extent(1.0, g->dz) + 3, 4);     // rows of whole 16-byte units, origin rounded down to one (+3)
  if (t.zl > 256) return SMK_ERR_UNSUPPORTED;
  t.box_elems = round_up(t.xw * t.yw * t.zl, 32);
  for (int f = 0; f < nf; ++f) {
    const cuuint64_t dims[3] = {(cuuint64_t)p.nz, (cuuint64_t)p.ny, (cuuint64_t)p.nxs};
    const cuuint64_t strides[2] = {(cuuint64_t)p.nz * sizeof(float), (cuuint64_t)p.nz * p.ny * sizeof(float)};
    const cuuint32_t boxd[3] = {(cuuint32_t)t.zl, (cuuint32_t)t.yw, (cuuint32_t)t.xw};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = encode(&t.map[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p.f[f], dims, strides, boxd, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return SMK_ERR_UNSUPPORTED;
  }
  const long long nsegs = (long long)p.nqso * p.nseg;
  if (nsegs > 2147483647LL / 2) { set_error("smk_skewers: too many quasars for one launch"); return SMK_ERR_ARG; }
  // scratch: hand-back list [1 + nsegs] and work list [2 + nsegs]
  t.handback = (int*)smk_ctx_scratch(ctx, (size_t)(2 * nsegs + 4) * sizeof(int));
  if (!t.handback) return SMK_ERR_CUDA;
  t.work = t.handback + (nsegs + 2);
  SMK_CUDA_OK(cudaMemsetAsync(t.handback, 0, sizeof(int), st));
  SMK_CUDA_OK(cudaMemsetAsync(t.work, 0, 2 * sizeof(int), st));
  { const int rc = launch_sentinels(p, 32 * P, st); if (rc) return rc; }
  skewers_worklist_kernel<<<(p.nqso + 255) / 256, 256, 0, st>>>(p, 32 * P, t.work);
  SMK_CUDA_OK(cudaGetLastError());
  auto kern = skewers_tma_kernel<P, NFG, MINB>;
  const size_t smem = (size_t)2 * NFG * t.box_elems * sizeof(float) + 16;      // two stages of NFG boxes + two mbarriers
  if (smem > 227 * 1024) return SMK_ERR_UNSUPPORTED;
  SMK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, nsm = 0, per_sm = 0;
  SMK_CUDA_OK(cudaGetDevice(&dev));
  SMK_CUDA_OK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  SMK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
  if (per_sm < 1) return SMK_ERR_UNSUPPORTED;
  long long grid = (long long)nsm * per_sm;
  if (grid > nsegs) grid = nsegs;
  kern<<<(unsigned)grid, 32, smem, st>>>(t);
  SMK_CUDA_OK(cudaGetLastError());
  g_last.handback = t.handback; g_last.nsegs = nsegs; g_last.xw = t.xw; g_last.yw = t.yw; g_last.zl = t.zl; g_last.st = st;
  // the segments handed back (box edges, oblique segments): fixed grid over the list
  SkewerParams pl = p;
  pl.list = t.handback;
  skewers_multi_kernel<P><<<148 * 2, 128, 0, st>>>(pl);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

// *fused (if given) tells whether the kernel that ran has the FGPA epilogue (the register-blocked ones have)
int launch_skewers(smk_ctx* ctx, const smk_geom* g, SkewerParams& p, int dmax, double pixel_step, bool staged,
                   cudaStream_t st, bool* fused = nullptr) {
  if (fused) *fused = false;
  g_last = LastGather();
  if (p.nqso == 0 || p.npix == 0) return SMK_OK;
  const int NT = 128;
  // register-blocked kernels: valid while P consecutive pixels cannot cross two cell boundaries on any axis
  const double cell = fmin(p.dx, fmin(p.dy, p.dz));
  auto blocked_ok = [&](int P) { return (dmax == DMAX) && pixel_step > 0 && (P - 1) * pixel_step < cell; };
  if (blocked_ok(4)) {
    if (fused) *fused = true;
    if (staged && ctx) {
      // 4 pixels per thread, one field per stage, up to 12 single-warp CTAs per SM: the measured best of the variants
      // tried on B200 (two fields per stage; 8 / 12 / 16 warps per SM; 6 and 8 pixels per thread: profiles/README.md)
      const int rc = launch_staged<4, 1, 12>(ctx, g, p, st);
      if (rc != SMK_ERR_UNSUPPORTED) return rc;
    }
    p.nseg = (p.npix + 32 * 4 - 1) / (32 * 4);
    p.list = nullptr;
    const long long nblocks = ((long long)p.nqso * p.nseg + 3) / 4;
    if (nblocks > 2147483647LL) { set_error("smk_skewers: too many quasars for one launch"); return SMK_ERR_ARG; }
    { const int rc = launch_sentinels(p, 32 * 4, st); if (rc) return rc; }
    skewers_multi_kernel<4><<<(unsigned)nblocks, NT, 0, st>>>(p);
    SMK_CUDA_OK(cudaGetLastError());
    return SMK_OK;
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
  // largest step between consecutive pixels of the grid (uniform 0.2 Mpc/h in the reference); decides whether the
  // register-blocked kernels may be used.  Parity-test switch (smk_set_option "skewers_kernel"): 1 = the global-memory
  // walk for every segment, 2 = the one-pixel-per-thread kernel.
  double step = g->pixel_step;
  const int which = smk_option("skewers_kernel");
  if (which == 2) step = 0.0;
  const bool staged = which == 0;
  bool fused = false;
  int rc = launch_skewers(ctx, g, p, g->dmax, step, staged, smk_ctx_stream(ctx), &fused);
  if (rc != SMK_OK || !flux || fused) return rc;
  // the one-pixel-per-thread kernel has no epilogue: FGPA as a separate pass over the rows
  return smk_fgpa(ctx, nqso, npix, delta_l, delta_s, eta_par, growthf, fa, fb, fc, flux);
}

extern "C" int smk_skewers_stats(smk_ctx* ctx, long long* segments, long long* handed_back, int box[3]) {
  using namespace smk;
  (void)ctx;
  if (segments) *segments = g_last.nsegs;
  if (handed_back) *handed_back = 0;
  if (box) { box[0] = g_last.xw; box[1] = g_last.yw; box[2] = g_last.zl; }
  if (g_last.handback && handed_back) {
    int n = 0;
    SMK_CUDA_OK(cudaMemcpyAsync(&n, g_last.handback, sizeof(int), cudaMemcpyDeviceToHost, g_last.st));
    SMK_CUDA_OK(cudaStreamSynchronize(g_last.st));
    *handed_back = n;
  }
  return SMK_OK;
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
