// Small-scale 1-D field and FGPA (bin/merge_spectra.py:303-339, py/SaclayMocks/util.py:421-433).
//
// smallscale: one CTA per quasar; white noise (given, or Philox) -> nfft-point r2c -> multiply by
// sqrt(max(P_miss(z_eff,k),0)/pixsize) -> c2r -> first npix samples -> * sigma_s(z)/sigma_s(z_eff).
// Both real transforms run as half-length complex FFTs entirely in shared memory (<= 32 KB).
#include <math.h>

#include "smk_fft.cuh"
#include "smk_internal.h"
#include "smk_philox.cuh"

namespace smk {

struct SmallScaleParams {
  int nqso, npix;
  const float* noise;      // [nqso][nfft] or null
  uint64_t seed;
  const float* filt;       // [nrows][nfft/2+1]
  const int* row_of_qso;   // [nqso]
  const float* sig_pix;    // [npix]
  const float* sig_eff;    // [nqso]
  float* delta_s;          // [nqso][npix]
  const long long* ids;    // [nqso] Philox stream id per quasar (null: the row index)
  const float2* tw;        // W_nfft
};

// All stages of plan P on the single line of the CTA through the (padded) load/store functors; the last stage
// re-sorts its outputs to natural order in place.  A barrier follows every stage.
template <class P, int S0, bool INV, int NT, class Load, class Store>
__device__ __forceinline__ void ss_fft(Load ld, Store st, const float2* __restrict__ tw) {
  if constexpr (S0 < P::S - 1) {
    dif_stage<P, S0, INV, 1, NT, OUT_INPLACE>(ld, st, tw, 2);
    __syncthreads();
    ss_fft<P, S0 + 1, INV, NT>(ld, st, tw);
  } else {
    dif_stage<P, P::S - 1, INV, 1, NT, OUT_RESORT>(ld, st, tw, 2);
    __syncthreads();
  }
}

// Shared-memory index of point p of the single line a CTA transforms: one float2 of padding after every 16 points.
// With one line per CTA the lanes of a warp run DIFFERENT butterflies, so the last radix-16 stage reads points
// 16 b + t for consecutive b: unpadded that is a 128-byte stride (all lanes on one bank pair, a 16-way conflict; the
// first version spent 85 % of its time in those replays); padded it is a 17-float2 stride, conflict free.
__device__ __forceinline__ int ss_idx(int p) { return p + (p >> 4); }

template <int M>
__global__ void __launch_bounds__(M >= 2048 ? 512 : (M >= 512 ? 256 : 64)) smallscale_kernel(SmallScaleParams p) {
  using P = typename PlanFor<M>::type;
  constexpr int NT = M >= 2048 ? 512 : (M >= 512 ? 256 : 64);
  extern __shared__ float2 sm[];   // [M + M/16 + 1]
  const int q = blockIdx.x;
  const int nfft = 2 * M;
  const int frow = p.row_of_qso[q];
  if (frow < 0) {                  // empty forest: delta_s = 0 (merge_spectra.py:327-330)
    for (int i = threadIdx.x; i < p.npix; i += NT) p.delta_s[(size_t)q * p.npix + i] = 0.f;
    return;
  }
  auto ld = [&](int, int pos, int, int) { return sm[ss_idx(pos)]; };
  auto st = [&](int, int pos, float2 val) { sm[ss_idx(pos)] = val; };
  // ---- white noise as M float2
  if (p.noise) {
    const float2* in2 = reinterpret_cast<const float2*>(p.noise + (size_t)q * nfft);
    for (int n = threadIdx.x; n < M; n += NT) sm[ss_idx(n)] = __ldg(in2 + n);
  } else {
    for (int c = threadIdx.x; c < M / 2; c += NT) {
      uint64_t id = p.ids ? (uint64_t)p.ids[q] : (uint64_t)q;
      float4 g = philox_normal4(p.seed ^ 0x5ca1ab1e5eedULL, id * (uint64_t)(M / 2) + c);
      sm[ss_idx(2 * c)] = make_float2(g.x, g.y);
      sm[ss_idx(2 * c + 1)] = make_float2(g.z, g.w);
    }
  }
  __syncthreads();
  // ---- forward r2c: M-point complex FFT of z[n] = x[2n] + i x[2n+1], output re-sorted to natural order
  ss_fft<P, 0, false, NT>(ld, st, p.tw);
  const float* filt = p.filt + (size_t)frow * (M + 1);
  // ---- post-process to X[k], multiply by the filter, and pre-process for the inverse, pair (k, M-k) at a time
  for (int k = threadIdx.x; k <= M / 2; k += NT) {
    float2 a = sm[ss_idx(k)], b = (k == 0) ? a : sm[ss_idx(M - k)];
    float2 w = __ldg(p.tw + k);
    float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
    float2 t = cmul(w, d);
    float2 mit = make_float2(t.y, -t.x);
    float2 xk = cadd(e, mit);
    float2 xm = make_float2(e.x - mit.x, -e.y + mit.y);
    float fk = __ldg(filt + k), fm = __ldg(filt + (M - k));
    xk = cscale(xk, fk);
    xm = cscale(xm, fm);
    if (k == 0) { xk.y = 0.f; xm.y = 0.f; }
    if (k == M - k) xm = xk;
    // inverse pre-step: Z[k] = A + iB, A = X[k] + conj(X[M-k]), B = (X[k] - conj(X[M-k])) conj(w)
    float2 wc = make_float2(w.x, -w.y);
    float2 A = make_float2(xk.x + xm.x, xk.y - xm.y);
    float2 B = cmul(make_float2(xk.x - xm.x, xk.y + xm.y), wc);
    sm[ss_idx(k)] = make_float2(A.x - B.y, A.y + B.x);
    if (k != 0 && k != M - k) sm[ss_idx(M - k)] = make_float2(A.x + B.y, -A.y + B.x);
  }
  __syncthreads();
  ss_fft<P, 0, true, NT>(ld, st, p.tw);
  // ---- numpy's irfft normalises by 1/nfft; keep the first npix samples; z-dependence of sigma_s
  const float inv_n = 1.0f / (float)nfft;
  const float se = p.sig_eff ? 1.0f / p.sig_eff[q] : 1.0f;
  const float* sm_f = reinterpret_cast<const float*>(sm);
  for (int i = threadIdx.x; i < p.npix; i += NT) {
    float v = (i < nfft) ? sm_f[2 * ss_idx(i >> 1) + (i & 1)] * inv_n : 0.f;
    if (p.sig_pix) v *= p.sig_pix[i] * se;
    p.delta_s[(size_t)q * p.npix + i] = v;
  }
}

template <int M>
static int launch_smallscale_t(const SmallScaleParams& p, cudaStream_t st) {
  constexpr int NT = M >= 2048 ? 512 : (M >= 512 ? 256 : 64);
  size_t smem = (size_t)(M + M / 16 + 1) * sizeof(float2);
  smallscale_kernel<M><<<p.nqso, NT, smem, st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

// ---- 1-D power spectrum of the rows (py/SaclayMocks/powerspectrum.py:204-238: P1D_1spectrum = |rfft(delta)|^2 DX / n,
// accumulated over spectra by ComputeP1D).  One CTA per row: the nfft pixels starting at first[q] (rows with fewer than
// nfft valid pixels are skipped), optionally turned into a contrast v / mean[q] - 1, nfft-point r2c in shared memory
// (the half-length complex FFT of smallscale_kernel), then sum P and sum P^2 per wavenumber with float64 atomics.
struct P1DParams {
  int nqso, npix;
  const float* rows;       // [nqso][npix]
  const int* first;        // [nqso] first pixel of the window (null: 0)
  const int* nvalid;       // [nqso] valid pixels from `first` on (null: npix)
  const float* mean;       // [nqso] (null: the rows are used as they are)
  float scale;             // DX / nfft
  double* sums;            // [2][nfft/2+1]
  unsigned long long* nused;
  const float2* tw;
};

template <int M>
__global__ void __launch_bounds__(M >= 2048 ? 512 : (M >= 512 ? 256 : 64)) p1d_kernel(P1DParams p) {
  using P = typename PlanFor<M>::type;
  constexpr int NT = M >= 2048 ? 512 : (M >= 512 ? 256 : 64);
  extern __shared__ float2 sm[];   // [M + M/16 + 1]
  const int q = blockIdx.x;
  const int nfft = 2 * M;
  const int i0 = p.first ? p.first[q] : 0;
  const int nv = p.nvalid ? p.nvalid[q] : p.npix - i0;
  if (i0 < 0 || nv < nfft || i0 + nfft > p.npix) return;
  auto ld = [&](int, int pos, int, int) { return sm[ss_idx(pos)]; };
  auto st = [&](int, int pos, float2 val) { sm[ss_idx(pos)] = val; };
  const float* row = p.rows + (size_t)q * p.npix + i0;
  const float inv_mean = p.mean ? 1.0f / p.mean[q] : 1.0f, off = p.mean ? 1.0f : 0.0f;
  for (int n = threadIdx.x; n < M; n += NT)
    sm[ss_idx(n)] = make_float2(row[2 * n] * inv_mean - off, row[2 * n + 1] * inv_mean - off);
  __syncthreads();
  ss_fft<P, 0, false, NT>(ld, st, p.tw);
  for (int k = threadIdx.x; k <= M / 2; k += NT) {
    float2 a = sm[ss_idx(k)], b = (k == 0) ? a : sm[ss_idx(M - k)];
    float2 w = __ldg(p.tw + k);
    float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
    float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
    float2 t = cmul(w, d);
    float2 mit = make_float2(t.y, -t.x);
    float2 xk = cadd(e, mit);                                       // X[k]
    float2 xm = make_float2(e.x - mit.x, -e.y + mit.y);             // X[M-k]
    const double pk = (double)((xk.x * xk.x + xk.y * xk.y) * p.scale);
    atomicAdd(p.sums + k, pk);
    atomicAdd(p.sums + (M + 1) + k, pk * pk);
    if (k != M - k) {
      const double pm = (double)((xm.x * xm.x + xm.y * xm.y) * p.scale);
      atomicAdd(p.sums + (M - k), pm);
      atomicAdd(p.sums + (M + 1) + (M - k), pm * pm);
    }
  }
  if (threadIdx.x == 0) atomicAdd(p.nused, 1ULL);
}

template <int M>
static int launch_p1d_t(const P1DParams& p, cudaStream_t st) {
  constexpr int NT = M >= 2048 ? 512 : (M >= 512 ? 256 : 64);
  size_t smem = (size_t)(M + M / 16 + 1) * sizeof(float2);
  p1d_kernel<M><<<p.nqso, NT, smem, st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

__global__ void fgpa_kernel(size_t n, int npix, const float* __restrict__ delta_l, const float* __restrict__ delta_s,
                            const float* __restrict__ eta, const float* __restrict__ G, const float* __restrict__ a,
                            const float* __restrict__ b, const float* __restrict__ c, float* __restrict__ flux) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; idx < n; idx += stride) {
    int i = (int)(idx % npix);
    float d = delta_l[idx];
    if (delta_s) d += delta_s[idx];
    if (eta) d += c[i] * eta[idx];
    float tau_over_a = expf(b[i] * G[i] * d);      // util.py:427
    flux[idx] = expf(-a[i] * tau_over_a);          // util.py:428
  }
}

}  // namespace smk

// twiddle tables for the 1-D transforms are cached per (device, nfft) in the library (tiny: <= 64 KB each)
#include <mutex>
static std::mutex g_tw1d_mutex;
static float2* g_tw1d[64][16] = {};

int smk_tw1d(int nfft, const float2** out) {
  int lg = 0, dev = 0;
  while ((1 << lg) < nfft) ++lg;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || lg >= 16) {
    smk::set_error("twiddle cache: bad device or transform length");
    return SMK_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lock(g_tw1d_mutex);
  if (!g_tw1d[dev][lg]) {
    float2* h = new float2[nfft];
    for (int k = 0; k < nfft; ++k) {
      double ang = -2.0 * M_PI * (double)k / (double)nfft;
      h[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    cudaError_t e = cudaMalloc(&g_tw1d[dev][lg], nfft * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(g_tw1d[dev][lg], h, nfft * sizeof(float2), cudaMemcpyHostToDevice);
    delete[] h;
    if (e != cudaSuccess) {
      smk::set_error(std::string("twiddle upload: ") + cudaGetErrorString(e));
      g_tw1d[dev][lg] = nullptr;
      return SMK_ERR_CUDA;
    }
  }
  *out = g_tw1d[dev][lg];
  return SMK_OK;
}

extern "C" int smk_smallscale(smk_ctx* ctx, int nqso, int nfft, int npix, const float* noise, uint64_t seed,
                              const float* filt_rows, const int* row_of_qso, const float* sig_pix,
                              const float* sig_eff, const long long* qso_ids, float* delta_s) {
  using namespace smk;
  if (nqso == 0) return SMK_OK;
  if (!filt_rows || !row_of_qso || !delta_s || npix > nfft) { set_error("smk_smallscale: bad argument"); return SMK_ERR_ARG; }
  SmallScaleParams p{nqso, npix, noise, seed, filt_rows, row_of_qso, sig_pix, sig_eff, delta_s, qso_ids, nullptr};
  int rc = smk_tw1d(nfft, &p.tw);
  if (rc) return rc;
  cudaStream_t st = smk_ctx_stream(ctx);
  switch (nfft) {
    case 256: return launch_smallscale_t<128>(p, st);
    case 512: return launch_smallscale_t<256>(p, st);
    case 1024: return launch_smallscale_t<512>(p, st);
    case 2048: return launch_smallscale_t<1024>(p, st);
    case 4096: return launch_smallscale_t<2048>(p, st);
    case 8192: return launch_smallscale_t<4096>(p, st);
  }
  set_error("smk_smallscale: nfft must be a power of two in [256, 8192]");
  return SMK_ERR_UNSUPPORTED;
}

extern "C" int smk_p1d(smk_ctx* ctx, int nqso, int npix, int nfft, const float* rows, const int* first, const int* nvalid,
                       const float* mean, double pixel, double* sums, unsigned long long* nused) {
  using namespace smk;
  if (nqso == 0) return SMK_OK;
  if (!rows || !sums || !nused || nfft > npix || pixel <= 0) { set_error("smk_p1d: bad argument"); return SMK_ERR_ARG; }
  P1DParams p{nqso, npix, rows, first, nvalid, mean, (float)(pixel / nfft), sums, nused, nullptr};
  int rc = smk_tw1d(nfft, &p.tw);
  if (rc) return rc;
  cudaStream_t st = smk_ctx_stream(ctx);
  switch (nfft) {
    case 256: return launch_p1d_t<128>(p, st);
    case 512: return launch_p1d_t<256>(p, st);
    case 1024: return launch_p1d_t<512>(p, st);
    case 2048: return launch_p1d_t<1024>(p, st);
    case 4096: return launch_p1d_t<2048>(p, st);
    case 8192: return launch_p1d_t<4096>(p, st);
  }
  set_error("smk_p1d: nfft must be a power of two in [256, 8192]");
  return SMK_ERR_UNSUPPORTED;
}

extern "C" int smk_fgpa(smk_ctx* ctx, int nqso, int npix, const float* delta_l, const float* delta_s,
                        const float* eta_par, const float* growthf, const float* a, const float* b, const float* c,
                        float* flux) {
  using namespace smk;
  size_t n = (size_t)nqso * npix;
  if (n == 0) return SMK_OK;
  if (!delta_l || !growthf || !a || !b || !flux || (eta_par && !c)) { set_error("smk_fgpa: null argument"); return SMK_ERR_ARG; }
  int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  fgpa_kernel<<<blocks, 256, 0, smk_ctx_stream(ctx)>>>(n, npix, delta_l, delta_s, eta_par, growthf, a, b, c, flux);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}
