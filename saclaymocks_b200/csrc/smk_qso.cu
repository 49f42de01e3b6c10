// Quasar drawing on device-resident boxes: the cell loop of bin/draw_qso.py:228-251 (ptot from the three
// lognormal boxes) and :394-480 (cond1 & cond2 & cond3, random position in the cell, redshift-space shift of the
// quasar redshift) for one x-slab, one thread per cell.
//
// The reference evaluates everything for every cell of every z plane; here a thread first evaluates cond1
// (rnd1 < norm * ptot, true for ~4e-4 of the cells) and only the survivors go through the n(z) rejection, the
// (ra, dec) rotation and the RSD shift.  Arithmetic is float64 in the reference's operation order (this file is
// compiled with -fmad=false so that nvcc does not contract a*b+c), the two float32 roundings of ptot
// (np.exp in place on the float32 box, `p **= a(z)` cast back to float32) are kept, and all 1-D tables are
// interpolated with scipy.interpolate.interp1d's slope*(x-x_lo)+y_lo.  Uniform variates come either from the
// caller (the reference's legacy NumPy stream: bit-identical selection given identical boxes) or from
// Philox4x32-10 keyed by (seed, global cell index).  Selected quasars are appended to a record buffer with their
// (plane, ix, iy) key; the host orders them like the reference's np.where.
#include <math.h>
#include <stdlib.h>

#include "smk_internal.h"
#include "smk_philox.cuh"

namespace smk {

// searchsorted(x, v, side='left') clipped to [1, n-1], then slope * (v - x_lo) + y_lo  (scipy interp1d, linear)
__device__ __forceinline__ double interp1d(const double* __restrict__ x, const double* __restrict__ y, int n, double v) {
  int lo = 0, hi = n;                       // first index with x[idx] >= v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(x + mid) < v) lo = mid + 1; else hi = mid;
  }
  int idx = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
  const double xl = __ldg(x + idx - 1), xh = __ldg(x + idx), yl = __ldg(y + idx - 1), yh = __ldg(y + idx);
  const double slope = (yh - yl) / (xh - xl);
  return slope * (v - xl) + yl;
}

__device__ __forceinline__ double bias_qso(double z) { return 3.7 * pow((1 + z) / (1 + 2.33), 1.7); }      // util.py:513
// util.py:517 with bias_qso(z) evaluated once per redshift (the reference evaluates the same expression per call)
__device__ __forceinline__ double a_of_z(double z, double bias_z, double zb, double bias_zb) {
  return bias_z * (1 + zb) / (bias_zb * (1 + z));
}

__device__ __forceinline__ double diffmod(double a, double b, double c) {       // util.py:116-120 (python %: sign of c)
  double d = fmod(a - b, c);
  if (d != 0.0 && (d < 0.0) != (c < 0.0)) d += c;
  return fmin(d, c - d);
}

// 53-bit uniform in [0,1) from Philox: same construction as NumPy's random_sample ((a >> 5) * 2^26 + (b >> 6)) / 2^53
// Counter word 3 carries a domain tag: the white noise of the boxes (smk_philox.cuh) uses tag 0, the cond1 variates
// 0x51 and these streams 0x52, so that no two consumers of the same seed ever share a Philox block.
__device__ __forceinline__ void philox_uniform2(uint64_t seed, uint64_t ctr, uint32_t stream, double& u0, double& u1) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), stream, 0x52u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  u0 = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0;
  u1 = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) / 9007199254740992.0;
}

// Everything after cond1 for one cell (draw_qso.py:403-480): n(z) rejection, position in the cell, (ra, dec, R), RSD shift
// of the redshift, cond3, record.  Float64 in the reference's operation order.
__device__ __noinline__ void finish_cell(const smk_qso_params& p, size_t idx, int ix, int iy, int mz, double u2, double ux, double uy,
                            double uz, double bz1, double bz2, double bz3, int* __restrict__ counters,
                            double* __restrict__ records, int capacity) {
  const double xa = __ldg(p.x_axis + ix), ya = __ldg(p.y_axis + iy), za = __ldg(p.z_axis + mz);
  const size_t pl = ((size_t)mz * p.nxs + ix) * p.ny + iy;
  // ---- cond2: n(z) rejection (:403-421, :434)
  const double rr0 = sqrt(za * za + (xa * xa + ya * ya));
  const double z0c = interp1d(p.chi, p.zt, p.ntab, rr0 / p.h);
  const long long izn = llrint((z0c - p.dz_interp0) / p.delta_z);    // np.round: half to even
  double density = p.dn_cell[izn];
  {
    const double c = interp1d(p.coef_z, p.coef_v, p.ncoef, z0c);
    const double b0 = bias_qso(z0c);
    const double a1 = a_of_z(z0c, b0, p.z1, bz1) * p.sigma_p[0], a2 = a_of_z(z0c, b0, p.z2, bz2) * p.sigma_p[1],
                 a3 = a_of_z(z0c, b0, p.z3, bz3) * p.sigma_p[2];
    const double g1 = exp(a1 * a1 / 2), g2 = exp(a2 * a2 / 2), g3 = exp(a3 * a3 / 2);
    density /= (c * (g1 * (p.z2 - z0c) / (p.z2 - p.z1) + g2 * (z0c - p.z1) / (p.z2 - p.z1)) +
                (1 - c) * (g2 * (p.z3 - z0c) / (p.z3 - p.z2) + g3 * (z0c - p.z2) / (p.z3 - p.z2)));
  }
  if (!(p.density_max * u2 < density)) return;
  // ---- position in the cell, (ra, dec, R) (:437-445; box.py:201-237)
  const double X = xa + (-p.dx / 2 + (p.dx / 2 - -p.dx / 2) * ux);
  const double Y = ya + (-p.dy / 2 + (p.dy / 2 - -p.dy / 2) * uy);
  const double Z = za + (-p.dz / 2 + (p.dz / 2 - -p.dz / 2) * uz);
  const double numra = p.cr0 * X - p.sd0 * p.sr0 * Y + p.cd0 * p.sr0 * Z;
  const double denomra = -p.sr0 * X - p.sd0 * p.cr0 * Y + p.cd0 * p.cr0 * Z;
  const double numdec = p.cd0 * Y + p.sd0 * Z;
  const double RR = sqrt(X * X + Y * Y + Z * Z);
  const double PI = 3.141592653589793;
  double ra = 0.0;
  if (numra > 0 && denomra > 0) ra = atan(numra / denomra);
  else if (numra > 0 && denomra < 0) ra = atan(numra / denomra) + PI;
  else if (numra < 0 && denomra < 0) ra = atan(numra / denomra) + PI;
  else if (numra < 0 && denomra > 0) ra = atan(numra / denomra) + 2 * PI;
  else if (numra > 0 && denomra == 0) ra = PI / 2;
  else if (numra == 0 && denomra < 0) ra = PI;
  else if (numra < 0 && denomra == 0) ra = 3 * PI / 2;
  double dec = asin(numdec / RR);
  ra = ra * (180.0 / PI);                                             // np.degrees
  dec = dec * (180.0 / PI);
  const double zq = interp1d(p.chi, p.zt, p.ntab, RR / p.h);
  // ---- redshift-space shift of the quasar redshift (:450-457)
  double zrsd = zq;
  if (p.rsd) {
    const double vpar = (X * (double)p.velo[0][idx] + Y * (double)p.velo[1][idx] + Z * (double)p.velo[2][idx]) / RR;
    double rr_rsd = RR;
    if (zq < p.z_max + 1.0) rr_rsd += vpar * (1 + zq) * interp1d(p.dg_z, p.dg_v, p.ndg, zq) / (p.dgrowth0 * p.H0);
    zrsd = interp1d(p.chi, p.zt, p.ntab, rr_rsd / p.h);
  }
  // ---- cond3 (:459-462)
  if (!(diffmod(ra, p.ra0, 360.0) < p.dra) || !(diffmod(dec, p.dec0, 180.0) < p.ddec) || !(zrsd > p.z_min) ||
      !(zrsd < p.z_max))
    return;
  const int slot = atomicAdd(counters, 1);
  if (slot < capacity) {
    double* r = records + (size_t)slot * 8;
    r[0] = (double)pl;
    r[1] = zq; r[2] = zrsd; r[3] = ra; r[4] = dec; r[5] = X; r[6] = Y; r[7] = Z;
  }
}

// cond1 variate of cell (ix, iy, mz) in Philox mode: one Philox block per four consecutive z cells of a column, a
// 32-bit draw scaled in float32 -- the SAME variate in the exact and the production kernel, so that their selections
// differ only where the production kernel's float32 ptot rounds across the variate (tests/test_gpu_qso.py)
__device__ __forceinline__ void cond1_block(const smk_qso_params& p, int ix, int iy, int iz4, uint32_t (&c)[4]) {
  const uint64_t col = (uint64_t)(p.ix0 + ix) * p.ny + iy;
  c[0] = (uint32_t)col; c[1] = (uint32_t)(col >> 32); c[2] = (uint32_t)iz4; c[3] = 0x51u;
  philox4x32_10(c, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
}
__device__ __forceinline__ float cond1_uniform(uint32_t w) { return (float)w * 2.3283064365386963e-10f; }

// uniforms of a cond1 survivor in Philox mode: the reference shares one x offset per (plane, ix) and one y offset per
// (plane, iy), so those are keyed on the row; u2 and uz on the global cell
__device__ __forceinline__ void survivor_uniforms(const smk_qso_params& p, int ix, int iy, int mz, double& u2, double& ux,
                                                  double& uy, double& uz) {
  double d0, d1;
  const uint64_t gcell = ((uint64_t)(p.ix0 + ix) * p.ny + iy) * p.nz + mz;
  philox_uniform2(p.seed, (uint64_t)mz * p.nx_full + (p.ix0 + ix), 1u, ux, d0);
  philox_uniform2(p.seed, (uint64_t)mz * p.ny + iy, 2u, uy, d1);
  philox_uniform2(p.seed, gcell, 3u, uz, u2);
}

// ---- exact kernel: ptot in the reference's arithmetic (draw_qso.py:197-199, 237-249), one thread per cell.  Used with
// the caller's uniform arrays (parity with the reference's NumPy stream) and as the fallback of the Philox mode.
__global__ void __launch_bounds__(256) draw_qso_kernel(const smk_qso_params p, int* __restrict__ counters,
                                                       double* __restrict__ records, int capacity) {
  const size_t ncell = (size_t)p.nxs * p.ny * p.nz;
  const double bz1 = bias_qso(p.z1), bz2 = bias_qso(p.z2), bz3 = bias_qso(p.z3);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < ncell; idx += (size_t)gridDim.x * blockDim.x) {
    const int mz = (int)(idx % p.nz);
    const size_t t = idx / p.nz;
    const int iy = (int)(t % p.ny), ix = (int)(t / p.ny);
    const double xa = __ldg(p.x_axis + ix), ya = __ldg(p.y_axis + iy), za = __ldg(p.z_axis + mz);
    const double z_box = interp1d(p.chi, p.zt, p.ntab, sqrt((xa * xa + ya * ya) + za * za) / p.h);
    const float e1 = expf(p.boxln[0][idx]), e2 = expf(p.boxln[1][idx]), e3 = expf(p.boxln[2][idx]);
    const double bzb = bias_qso(z_box);
    const float p1 = (float)pow((double)e1, a_of_z(z_box, bzb, p.z1, bz1));
    const float p2 = (float)pow((double)e2, a_of_z(z_box, bzb, p.z2, bz2));
    const float p3 = (float)pow((double)e3, a_of_z(z_box, bzb, p.z3, bz3));
    const double p12 = (double)p1 * (p.z2 - z_box) / (p.z2 - p.z1) + (double)p2 * (z_box - p.z1) / (p.z2 - p.z1);
    const double p23 = (double)p2 * (p.z3 - z_box) / (p.z3 - p.z2) + (double)p3 * (z_box - p.z2) / (p.z3 - p.z2);
    const double ptot = interp1d(p.coef_z, p.coef_v, p.ncoef, z_box) * (p12 - p23) + p23;
    const size_t pl = ((size_t)mz * p.nxs + ix) * p.ny + iy;          // [plane][ix][iy], the reference's draw order
    double u1, u2, ux, uy, uz;
    if (p.u1) {
      u1 = p.u1[pl];
    } else {
      uint32_t c[4];
      cond1_block(p, ix, iy, mz >> 2, c);
      u1 = (double)cond1_uniform(c[mz & 3]);
    }
    if (!(u1 < p.norm * ptot)) continue;                               // cond1 (:427)
    atomicAdd(counters + 1, 1);                                        // "QSOs in the full box" (:430)
    if (p.u1) {
      u2 = p.u2[pl];
      ux = p.ux[(size_t)mz * p.nxs + ix];
      uy = p.uy[(size_t)mz * p.ny + iy];
      uz = p.uz[pl];
    } else {
      survivor_uniforms(p, ix, iy, mz, u2, ux, uy, uz);
    }
    finish_cell(p, idx, ix, iy, mz, u2, ux, uy, uz, bz1, bz2, bz3, counters, records, capacity);
  }
}

// ---- production kernel (Philox draws).  ptot = w1(R) exp(a1(R) g1) + w2(R) exp(a2(R) g2) + w3(R) exp(a3(R) g3) with
// g_k the lognormal boxes and w_k, a_k smooth functions of the cell's comoving distance R only: a table of
// (norm*w1, norm*w2, norm*w3, a1, a2, a3) on a uniform R grid is built on the device from the same 1-D tables (float64,
// qso_lut_kernel) and interpolated linearly in float32, so the per-cell work drops from three float64 pow, two binary
// searches and a float64 divide chain to three __expf.  Each thread takes four consecutive z cells (16-byte loads of
// the three boxes, one Philox call for their four cond1 variates); the ~4e-4 of the cells that pass cond1 go through
// the same float64 finish_cell() as the exact kernel.
#define SMK_QSO_LUT 2048
struct QsoLut {
  float r_lo, inv_dr;
  float pad[6];
  float e[SMK_QSO_LUT + 1][8];
};

__global__ void qso_lut_kernel(const smk_qso_params p, QsoLut* lut) {
  __shared__ double red[4];
  if (threadIdx.x == 0) {
    double x2lo = 1e300, x2hi = 0, y2lo = 1e300, y2hi = 0;
    for (int i = 0; i < p.nxs; ++i) { const double v = p.x_axis[i] * p.x_axis[i]; x2lo = fmin(x2lo, v); x2hi = fmax(x2hi, v); }
    for (int i = 0; i < p.ny; ++i) { const double v = p.y_axis[i] * p.y_axis[i]; y2lo = fmin(y2lo, v); y2hi = fmax(y2hi, v); }
    const double zlo = p.z_axis[0], zhi = p.z_axis[p.nz - 1];
    red[0] = sqrt(x2lo + y2lo + zlo * zlo) - 1.0;
    red[1] = sqrt(x2hi + y2hi + zhi * zhi) + 1.0;
    lut->r_lo = (float)red[0];
    lut->inv_dr = (float)(SMK_QSO_LUT / (red[1] - red[0]));
  }
  __syncthreads();
  const double r_lo = (double)(float)red[0], dr = 1.0 / (double)(float)(SMK_QSO_LUT / (red[1] - red[0]));
  const double bz1 = bias_qso(p.z1), bz2 = bias_qso(p.z2), bz3 = bias_qso(p.z3);
  for (int i = threadIdx.x; i <= SMK_QSO_LUT; i += blockDim.x) {
    const double z = interp1d(p.chi, p.zt, p.ntab, (r_lo + i * dr) / p.h);
    const double c = interp1d(p.coef_z, p.coef_v, p.ncoef, z), b = bias_qso(z);
    float* e = lut->e[i];
    e[0] = (float)(p.norm * c * (p.z2 - z) / (p.z2 - p.z1));
    e[1] = (float)(p.norm * (c * (z - p.z1) / (p.z2 - p.z1) + (1 - c) * (p.z3 - z) / (p.z3 - p.z2)));
    e[2] = (float)(p.norm * (1 - c) * (z - p.z2) / (p.z3 - p.z2));
    e[3] = 0.f;
    e[4] = (float)a_of_z(z, b, p.z1, bz1);
    e[5] = (float)a_of_z(z, b, p.z2, bz2);
    e[6] = (float)a_of_z(z, b, p.z3, bz3);
    e[7] = 0.f;
  }
}

// pd: device copy of the parameters (the out-of-line float64 path takes them by reference; a by-value kernel parameter
// would have to be copied to every thread's local memory for that)
__global__ void __launch_bounds__(256, 3) draw_qso_fast_kernel(const smk_qso_params* __restrict__ pd,
                                                               const QsoLut* __restrict__ lut, int* __restrict__ counters,
                                                               double* __restrict__ records, int capacity) {
  const smk_qso_params& p = *pd;
  const int nz4 = p.nz >> 2;
  const size_t nquad = (size_t)p.nxs * p.ny * nz4;
  const float r_lo = lut->r_lo, inv_dr = lut->inv_dr;
  const float4* __restrict__ b1 = reinterpret_cast<const float4*>(p.boxln[0]);
  const float4* __restrict__ b2 = reinterpret_cast<const float4*>(p.boxln[1]);
  const float4* __restrict__ b3 = reinterpret_cast<const float4*>(p.boxln[2]);
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquad; q += (size_t)gridDim.x * blockDim.x) {
    const int iz4 = (int)(q % nz4);
    const size_t t = q / nz4;
    const int iy = (int)(t % p.ny), ix = (int)(t / p.ny);
    const float4 g1 = __ldcs(b1 + q), g2 = __ldcs(b2 + q), g3 = __ldcs(b3 + q);
    const float xa = (float)__ldg(p.x_axis + ix), ya = (float)__ldg(p.y_axis + iy);
    const float xy2 = fmaf(xa, xa, ya * ya);
    uint32_t c[4];
    cond1_block(p, ix, iy, iz4, c);
    const float gg1[4] = {g1.x, g1.y, g1.z, g1.w}, gg2[4] = {g2.x, g2.y, g2.z, g2.w}, gg3[4] = {g3.x, g3.y, g3.z, g3.w};
    // (w_k, a_k) at the first and last cell of the quad from the table, linear in between (R is linear in the cell
    // index to 1e-3 Mpc/h over four cells, the tabulated functions to ~1e-5)
    float wq[2][3], aq[2][3];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float za = (float)__ldg(p.z_axis + 4 * iz4 + 3 * e);
      const float tt = (sqrtf(fmaf(za, za, xy2)) - r_lo) * inv_dr;
      int i = (int)tt;
      i = i < 0 ? 0 : (i > SMK_QSO_LUT - 1 ? SMK_QSO_LUT - 1 : i);
      const float f = tt - (float)i;
      const float4 w0 = *reinterpret_cast<const float4*>(lut->e[i]), a0 = *reinterpret_cast<const float4*>(lut->e[i] + 4);
      const float4 w1 = *reinterpret_cast<const float4*>(lut->e[i + 1]), a1 = *reinterpret_cast<const float4*>(lut->e[i + 1] + 4);
      wq[e][0] = fmaf(f, w1.x - w0.x, w0.x); wq[e][1] = fmaf(f, w1.y - w0.y, w0.y); wq[e][2] = fmaf(f, w1.z - w0.z, w0.z);
      aq[e][0] = fmaf(f, a1.x - a0.x, a0.x); aq[e][1] = fmaf(f, a1.y - a0.y, a0.y); aq[e][2] = fmaf(f, a1.z - a0.z, a0.z);
    }
    unsigned hit = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = j * (1.0f / 3.0f);
      const float pt = fmaf(fmaf(s, wq[1][0] - wq[0][0], wq[0][0]), __expf(fmaf(s, aq[1][0] - aq[0][0], aq[0][0]) * gg1[j]),
                            fmaf(fmaf(s, wq[1][1] - wq[0][1], wq[0][1]), __expf(fmaf(s, aq[1][1] - aq[0][1], aq[0][1]) * gg2[j]),
                                 fmaf(s, wq[1][2] - wq[0][2], wq[0][2]) * __expf(fmaf(s, aq[1][2] - aq[0][2], aq[0][2]) * gg3[j])));
      if (cond1_uniform(c[j]) < pt) hit |= 1u << j;                              // cond1: u < norm * ptot
    }
    if (!hit) continue;
    const double bz1 = bias_qso(p.z1), bz2 = bias_qso(p.z2), bz3 = bias_qso(p.z3);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      if (!((hit >> j) & 1)) continue;
      atomicAdd(counters + 1, 1);
      const int mz = 4 * iz4 + j;
      double u2, ux, uy, uz;
      survivor_uniforms(p, ix, iy, mz, u2, ux, uy, uz);
      finish_cell(p, 4 * q + j, ix, iy, mz, u2, ux, uy, uz, bz1, bz2, bz3, counters, records, capacity);
    }
  }
}

}  // namespace smk

extern "C" int smk_draw_qso(smk_ctx* ctx, const smk_qso_params* p, int* counters, double* records, int capacity) {
  using namespace smk;
  if (!p || !counters || !records || capacity < 0) { set_error("smk_draw_qso: null argument"); return SMK_ERR_ARG; }
  if (!p->boxln[0] || !p->boxln[1] || !p->boxln[2] || (p->rsd && (!p->velo[0] || !p->velo[1] || !p->velo[2]))) {
    set_error("smk_draw_qso: the three lognormal boxes (and the velocity boxes with rsd) are required");
    return SMK_ERR_ARG;
  }
  if (p->u1 && (!p->u2 || !p->ux || !p->uy || !p->uz)) { set_error("smk_draw_qso: incomplete uniform arrays"); return SMK_ERR_ARG; }
  if (!p->x_axis || !p->y_axis || !p->z_axis || !p->chi || !p->zt || !p->coef_z || !p->coef_v || !p->dn_cell || !p->dg_z ||
      !p->dg_v) {
    set_error("smk_draw_qso: missing table");
    return SMK_ERR_ARG;
  }
  const size_t ncell = (size_t)p->nxs * p->ny * p->nz;
  if (ncell == 0) return SMK_OK;
  size_t blocks = (ncell + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = smk_ctx_stream(ctx);
  const bool aligned = (((uintptr_t)p->boxln[0] | (uintptr_t)p->boxln[1] | (uintptr_t)p->boxln[2]) & 15) == 0;
  const bool exact = smk_option("qso_exact") != 0;    // parity tests: the reference-arithmetic kernel in Philox mode too
  if (!p->u1 && p->nz % 4 == 0 && aligned && !exact) {
    char* scratch = (char*)smk_ctx_scratch(ctx, sizeof(QsoLut) + sizeof(smk_qso_params));
    if (!scratch) return SMK_ERR_CUDA;
    QsoLut* lut = (QsoLut*)scratch;
    smk_qso_params* pd = (smk_qso_params*)(scratch + sizeof(QsoLut));
    SMK_CUDA_OK(cudaMemcpyAsync(pd, p, sizeof(smk_qso_params), cudaMemcpyHostToDevice, st));   // pageable source: staged
    qso_lut_kernel<<<1, 256, 0, st>>>(*p, lut);
    size_t nb = (ncell / 4 + 255) / 256;
    if (nb > 148 * 24) nb = 148 * 24;
    draw_qso_fast_kernel<<<(unsigned)nb, 256, 0, st>>>(pd, lut, counters, records, capacity);
  } else {
    draw_qso_kernel<<<(unsigned)blocks, 256, 0, st>>>(*p, counters, records, capacity);
  }
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}
