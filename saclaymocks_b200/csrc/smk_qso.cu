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
__device__ __forceinline__ void philox_uniform2(uint64_t seed, uint64_t ctr, uint32_t stream, double& u0, double& u1) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), stream, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  u0 = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6)) / 9007199254740992.0;
  u1 = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) / 9007199254740992.0;
}

__global__ void __launch_bounds__(256) draw_qso_kernel(const smk_qso_params p, int* __restrict__ counters,
                                                       double* __restrict__ records, int capacity) {
  const size_t ncell = (size_t)p.nxs * p.ny * p.nz;
  const double bz1 = bias_qso(p.z1), bz2 = bias_qso(p.z2), bz3 = bias_qso(p.z3);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < ncell; idx += (size_t)gridDim.x * blockDim.x) {
    const int mz = (int)(idx % p.nz);
    const size_t t = idx / p.nz;
    const int iy = (int)(t % p.ny), ix = (int)(t / p.ny);
    const double xa = __ldg(p.x_axis + ix), ya = __ldg(p.y_axis + iy), za = __ldg(p.z_axis + mz);
    // ---- ptot (draw_qso.py:197-199, 237-249)
    const double z_box = interp1d(p.chi, p.zt, p.ntab, sqrt((xa * xa + ya * ya) + za * za) / p.h);
    const float e1 = expf(p.boxln[0][idx]), e2 = expf(p.boxln[1][idx]), e3 = expf(p.boxln[2][idx]);
    const double bzb = bias_qso(z_box);
    const float p1 = (float)pow((double)e1, a_of_z(z_box, bzb, p.z1, bz1));
    const float p2 = (float)pow((double)e2, a_of_z(z_box, bzb, p.z2, bz2));
    const float p3 = (float)pow((double)e3, a_of_z(z_box, bzb, p.z3, bz3));
    const double p12 = (double)p1 * (p.z2 - z_box) / (p.z2 - p.z1) + (double)p2 * (z_box - p.z1) / (p.z2 - p.z1);
    const double p23 = (double)p2 * (p.z3 - z_box) / (p.z3 - p.z2) + (double)p3 * (z_box - p.z2) / (p.z3 - p.z2);
    const double ptot = interp1d(p.coef_z, p.coef_v, p.ncoef, z_box) * (p12 - p23) + p23;
    // ---- uniforms of this cell
    const size_t pl = ((size_t)mz * p.nxs + ix) * p.ny + iy;          // [plane][ix][iy], the reference's draw order
    double u1, u2, ux, uy, uz;
    if (p.u1) {
      u1 = p.u1[pl];
    } else {
      const uint64_t gcell = ((uint64_t)(p.ix0 + ix) * p.ny + iy) * p.nz + mz;
      philox_uniform2(p.seed, gcell, 0u, u1, u2);
    }
    if (!(u1 < p.norm * ptot)) continue;                               // cond1 (:427)
    atomicAdd(counters + 1, 1);                                        // "QSOs in the full box" (:430)
    if (p.u1) {
      u2 = p.u2[pl];
      ux = p.ux[(size_t)mz * p.nxs + ix];
      uy = p.uy[(size_t)mz * p.ny + iy];
      uz = p.uz[pl];
    } else {
      // the reference shares one x offset per (plane, ix) and one y offset per (plane, iy): key those on the row
      double d0, d1;
      philox_uniform2(p.seed, (uint64_t)mz * p.nx_full + (p.ix0 + ix), 1u, ux, d0);
      philox_uniform2(p.seed, (uint64_t)mz * p.ny + iy, 2u, uy, d1);
      const uint64_t gcell = ((uint64_t)(p.ix0 + ix) * p.ny + iy) * p.nz + mz;
      philox_uniform2(p.seed, gcell, 3u, uz, d0);
    }
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
    if (!(p.density_max * u2 < density)) continue;
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
      continue;
    const int slot = atomicAdd(counters, 1);
    if (slot < capacity) {
      double* r = records + (size_t)slot * 8;
      r[0] = (double)pl;
      r[1] = zq; r[2] = zrsd; r[3] = ra; r[4] = dec; r[5] = X; r[6] = Y; r[7] = Z;
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
  draw_qso_kernel<<<(unsigned)blocks, 256, 0, smk_ctx_stream(ctx)>>>(*p, counters, records, capacity);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}
