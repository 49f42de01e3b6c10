// 3-D r2c / c2r FFT passes with fused prologues/epilogues for the GRF box synthesis
// (replaces pyfftw.FFTW(...).execute() + the numpy passes around it in bin/make_boxes.py:40-125,
// 242-431).  All passes are HBM-bound: each reads and writes the box once.
//
//   strided complex pass (x or y axis): tile = LINES consecutive kz columns x N points in shared
//     memory as [point][line]; the first DIF stage reads straight from global memory (rows of
//     LINES complex = 128 B) with the spectral-weight / k-factor multiply fused into the load, the
//     last stage writes straight to global memory in natural order.
//   contiguous pass (z axis): tile = LINES lines x (NZ/2+1) complex as [line][point] with odd
//     pitch; half-length complex FFT + real-transform pre/post step; the inverse fuses /N, sum and
//     sum of squares (for sigma) into the store, the forward fuses Philox noise into the load.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "smk_fft.cuh"
#include "smk_ztile.cuh"
#include "smk_internal.h"
#include "smk_philox.cuh"

#ifndef SMK_MID_LINES
#define SMK_MID_LINES 16  // kz columns per tile for x/y lengths 512 and 1024
#endif
#ifndef SMK_LINES_1024
#define SMK_LINES_1024 SMK_MID_LINES
#endif
#ifndef SMK_BIG_LINES
#define SMK_BIG_LINES 8   // kz columns per tile for x/y lengths above 1024 (shared memory: N * LINES * 8 B)
#endif

namespace smk {

// ------------------------------------------------------------------ strided complex pass
template <int N>
struct StridedTraits {
  static constexpr int LINES = (N > 1024) ? SMK_BIG_LINES : (N == 1024 ? SMK_LINES_1024 : (N >= 512 ? SMK_MID_LINES : 16));
  static constexpr int NT_ = LINES * N / 32;
  // 2560 = 16*16*2*5: 640 threads give 2 radix-16 butterflies per thread and stage (64 + 32 registers of payload)
  static constexpr int NT = (N == 2560) ? 640 : (NT_ < 64 ? 64 : (NT_ > 512 ? 512 : NT_));
  // resident CTAs per SM the register allocation should allow (shared memory: N*LINES*8 B per CTA)
  static constexpr int SMEM = N * LINES * 8;
  static constexpr int MINB_ = 220 * 1024 / SMEM;
  static constexpr int MINB__ = MINB_ < 1 ? 1 : (MINB_ * NT > 1536 ? 1536 / NT : MINB_);
  // a radix-32 first stage holds 64 payload registers per thread: at most 2 CTAs of 256 threads per SM (128 registers)
  static constexpr bool BIG_R0 = PlanFor<N>::type::radix(0) >= 32;
  static constexpr int MINB = (BIG_R0 && MINB__ * NT > 512) ? (512 / NT < 1 ? 1 : 512 / NT) : MINB__;
};

struct StridedParams {
  const float2* in;
  float2* out;
  PassAddr ain, aout;
  int ncols;       // multiple of LINES (the padded pitch)
  int wcols;       // valid columns of the weight table rows (nz/2+1)
  MulArgs mul;
  const float2* tw;
  // peer mode (fused exchange): output point k goes to rank k / aout.nsplit, into peer[rank] (that rank's receive
  // buffer, already offset to this rank's chunk) at (k % nsplit) * lo_stride + outer * outer_stride + column
  float2* peer[SMK_MAX_RANKS];
  int nouter;    // tiles = (ncols / LINES) x nouter
  int x_sms;     // fused-exchange pass only: run persistent on this many CTAs (0 = one CTA per tile)
};

// element offset of point n: plain stride, or the two-level [hi][lo] form left behind by an all-to-all
template <bool SPLIT>
__device__ __forceinline__ long long point_off(const PassAddr& a, int n) {
  if (!SPLIT) return (long long)n * a.lo_stride;
  int hi = a.nsplit >= 2 ? (int)__umulhi((unsigned)n, a.magic) : n;   // n / nsplit, see make_fastdiv()
  int lo = n - hi * a.nsplit;
  return hi * a.hi_stride + lo * a.lo_stride;
}

enum { OUT_PLAIN = 0, OUT_SPLIT = 1, OUT_PEER = 2 };

// Fused-multiply x passes (measured at 512 x 512 x 1536, tools/x_sweep.sh): the k-factor variants (eta, velocity) load
// their first-stage tasks two at a time, which fits 80 registers = 3 CTAs per SM (0.738 -> 0.702 ms); the table variant
// needs the weights as well, spills at 80 registers (0.870 -> 0.914 ms) and keeps all loads in flight at 2 CTAs per SM.
#ifndef SMK_VEL_F64
#define SMK_VEL_F64 0     // 1: velocity factor in real float64 arithmetic (the first implementation, kept for A/B runs)
#endif
#ifndef SMK_MUL_BATCH
#define SMK_MUL_BATCH 2   // first-stage tasks loaded together in the k-factor passes (0 = all)
#endif

// One tile (LINES kz columns x N points) of a strided pass: column tile tile_x of outer index `outer`.
template <int N, bool INV, int MUL, bool SPLIT_IN, int SPLIT_OUT>
__device__ __forceinline__ void strided_tile(const StridedParams& p, const int tile_x, const int outer) {
  using P = typename PlanFor<N>::type;
  constexpr int LINES = StridedTraits<N>::LINES;
  constexpr int NT = StridedTraits<N>::NT;
  constexpr int R0 = P::radix(0);
  constexpr int TPT0_ALL = (N / R0 * LINES + NT - 1) / NT;
  constexpr int B0 = (MUL == MUL_NONE || MUL == MUL_TABLE || P::S == 1) ? 0 : SMK_MUL_BATCH;   // first-stage batch
  constexpr int TPT0 = (B0 > 0 && B0 < TPT0_ALL) ? B0 : TPT0_ALL;
  extern __shared__ float2 sm[];   // [N][LINES]
  const int col = tile_x * LINES + threadIdx.x % LINES;   // this thread's kz column (fixed for the tile)
  const long long in_col = p.ain.tile_width ? (long long)(col / p.ain.tile_width) * p.ain.tile_stride + col % p.ain.tile_width
                                            : (long long)col;
  const float2* __restrict__ inl = p.in + (outer * p.ain.outer_stride + in_col);
  float2* __restrict__ outl = p.out + (outer * p.aout.outer_stride + col);

  // ---- fused multiply of make_boxes.py:247-429 (reference float32 rounding order), applied to the loaded element
  float wv[MUL == MUL_TABLE ? TPT0 : 1][MUL == MUL_TABLE ? R0 : 1];
  const float* __restrict__ wtl = nullptr;
  float2* __restrict__ sbl = nullptr;
  float ky = 0.f, kz = 0.f, ky2 = 0.f, kz2 = 0.f;
  if (MUL == MUL_TABLE) {
    wtl = p.mul.wt + (outer * p.mul.wt_outer_stride + min(col, p.wcols - 1));
    if (p.mul.store_back) sbl = p.mul.store_back + (outer * p.ain.outer_stride + col);
  } else if (MUL == MUL_ETA || MUL == MUL_VEL) {
    ky = __ldg(p.mul.ko + outer + p.mul.outer0);
    kz = col < p.wcols ? __ldg(p.mul.kc + col) : 0.f;
    ky2 = __fmul_rn(ky, ky);
    kz2 = __fmul_rn(kz, kz);
  }
  auto ld_g = [&](int, int n, int i, int t) {
    if (MUL == MUL_TABLE) wv[i][t] = __ldg(wtl + n * p.mul.wt_n_stride);
    // the plain inverse pass (y) reads data nobody reads again: evict-first keeps L2 for the chained z pass's slot
    if (INV && MUL == MUL_NONE) return __ldcs(inl + point_off<SPLIT_IN>(p.ain, n));
    return inl[point_off<SPLIT_IN>(p.ain, n)];
  };
  auto pre = [&](int, int n, int i, int t, float2 v) {
    if (MUL == MUL_NONE) return v;
    if (MUL == MUL_TABLE) {
      float w = col < p.wcols ? wv[i][t] : 0.f;
      v = make_float2(__fmul_rn(v.x, w), __fmul_rn(v.y, w));
      if (sbl) sbl[point_off<SPLIT_IN>(p.ain, n)] = v;
      return v;
    }
    const float kx = __ldg(p.mul.kn + n);
    float kk = __fadd_rn(__fadd_rn(__fmul_rn(kx, kx), ky2), kz2);      // (kx*kx + ky*ky) + kz*kz
    if (n == 0 && outer + p.mul.outer0 == 0 && col == 0) kk = 1.f;     // make_boxes.py:306
    const float ka = p.mul.fa == 0 ? kx : (p.mul.fa == 1 ? ky : kz);
    if (MUL == MUL_ETA) {
      const float kb = p.mul.fb == 0 ? kx : (p.mul.fb == 1 ? ky : kz);
      const float f = fdiv_fast(__fmul_rn(ka, kb), kk, rcp_approx(kk));
      return make_float2(__fmul_rn(v.x, f), __fmul_rn(v.y, f));
    }
    // MUL_VEL: boxk *= -1j*k/kk*H0*dgrowth0 -- float32 up to "*H0", then float64 (numpy promotes on the float64
    // scalar dgrowth0 and rounds the complex128 product back to complex64).  The float64 product is formed as an
    // unevaluated float32 pair instead (f = fh + fl to 2^-48, v * f = p + t), rounded once at the end: the same float32
    // as numpy's unless the exact product lies within ~2^-46 of a rounding boundary (none in 2e7 random samples), at
    // 11 FP32 instructions per element instead of 5 float32 <-> float64 conversions and 3 DMUL, which made this variant
    // issue-bound on the conversion pipe (0.90 ms against 0.64 ms for the eta variant).
#if SMK_VEL_F64
    const float f32 = __fmul_rn(fdiv_fast(-ka, kk, rcp_approx(kk)), 100.0f);
    const double f = (double)f32 * p.mul.vscale;
    return make_float2((float)(-(double)v.y * f), (float)((double)v.x * f));
#else
    const float f32 = __fmul_rn(fdiv_fast(-ka, kk, rcp_approx(kk)), 100.0f);
    const float fh = __fmul_rn(f32, p.mul.vs_hi);
    const float fl = __fmaf_rn(f32, p.mul.vs_lo, __fmaf_rn(f32, p.mul.vs_hi, -fh));
    auto times_f = [&](float a) {
      const float pr = __fmul_rn(a, fh);
      return __fadd_rn(pr, __fmaf_rn(a, fl, __fmaf_rn(a, fh, -pr)));
    };
    return make_float2(-times_f(v.y), times_f(v.x));
#endif
  };
  auto st_g = [&](int, int k, float2 val) { outl[point_off<SPLIT_OUT == OUT_SPLIT>(p.aout, k)] = val; };
  auto st_s = [&](int line, int pos, float2 val) { sm[pos * LINES + line] = val; };
  auto ld_s = [&](int line, int pos, int, int) { return sm[pos * LINES + line]; };

  if constexpr (SPLIT_OUT == OUT_PEER) {
    // ---- fused exchange: finish the transform in shared memory in natural order, then copy every destination
    //      rank's block (aout.nsplit = nx/R consecutive x rows of this tile = one contiguous run of the tiled receive
    //      layout [src][kz tile][y_l][x_l][LINES]) over NVLink with 16-byte stores, 512 B per warp instruction
    if constexpr (P::S == 1) {
      dif_stage<P, 0, INV, LINES, NT, OUT_RESORT>(ld_g, st_s, p.tw, 1, pre);
    } else {
      dif_stage<P, 0, INV, LINES, NT, OUT_INPLACE, decltype(ld_g), decltype(st_s), decltype(pre), B0>(ld_g, st_s, p.tw, 1, pre);
      __syncthreads();
      dif_stages_smem<P, 1, P::S - 1, INV, LINES, 1, LINES, NT>(sm, p.tw, 1);
      dif_stage<P, P::S - 1, INV, LINES, NT, OUT_RESORT>(ld_s, st_s, p.tw, 1);
    }
    __syncthreads();
    const int per_dest = p.aout.nsplit * (LINES / 2);                       // float4 per destination block
    const long long doff = tile_x * p.aout.tile_stride + outer * p.aout.outer_stride;
    const float4* sm4 = reinterpret_cast<const float4*>(sm);
    for (int d = 0; d < N / p.aout.nsplit; ++d) {
      float4* dst = reinterpret_cast<float4*>(p.peer[d] + doff);
      const float4* src = sm4 + d * per_dest;
      for (int e = threadIdx.x; e < per_dest; e += NT) dst[e] = src[e];
    }
  } else if constexpr (P::S == 1) {
    dif_stage<P, 0, INV, LINES, NT, OUT_NATURAL>(ld_g, st_g, p.tw, 1, pre);
  } else {
    dif_stage<P, 0, INV, LINES, NT, OUT_INPLACE, decltype(ld_g), decltype(st_s), decltype(pre), B0>(ld_g, st_s, p.tw, 1, pre);
    __syncthreads();
    dif_stages_smem<P, 1, P::S - 1, INV, LINES, 1, LINES, NT>(sm, p.tw, 1);
    dif_stage<P, P::S - 1, INV, LINES, NT, OUT_NATURAL>(ld_s, st_g, p.tw, 1);
  }
}

#define SMK_STRIDED_BOUNDS(N, MUL)                                                                        \
  __launch_bounds__(StridedTraits<N>::NT,                                                                 \
                    (MUL != MUL_TABLE || StridedTraits<N>::MINB == 1 || StridedTraits<N>::BIG_R0)        \
                        ? StridedTraits<N>::MINB                                                          \
                        : StridedTraits<N>::MINB - 1)

// one CTA per tile: grid (ncols / LINES, nouter)
template <int N, bool INV, int MUL, bool SPLIT_IN, int SPLIT_OUT>
__global__ void SMK_STRIDED_BOUNDS(N, MUL) c2c_strided_kernel(const __grid_constant__ StridedParams p) {
  strided_tile<N, INV, MUL, SPLIT_IN, SPLIT_OUT>(p, blockIdx.x, blockIdx.y);
}

// Persistent form for the fused-exchange x pass: a 1-D grid of SMK_X_SMS CTAs walks the tiles.  That pass is bound by
// NVLink, not by the SMs; capping its grid leaves the other SMs to the y / z passes of the previous product, which run
// on the second stream (with one CTA per tile the pass floods every SM and the two streams only time-slice, DESIGN.md
// section 6).  The tile body is called out of line so that its register allocation stays that of the plain kernel.
template <int N, bool INV, int MUL, bool SPLIT_IN, int SPLIT_OUT>
__device__ __noinline__ void strided_tile_call(const StridedParams& p, int tile_x, int outer) {
  strided_tile<N, INV, MUL, SPLIT_IN, SPLIT_OUT>(p, tile_x, outer);
}
template <int N, bool INV, int MUL, bool SPLIT_IN, int SPLIT_OUT>
__global__ void SMK_STRIDED_BOUNDS(N, MUL) c2c_strided_persistent_kernel(const __grid_constant__ StridedParams p) {
  const int gx = p.ncols / StridedTraits<N>::LINES;
  const int ntiles = gx * p.nouter;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    strided_tile_call<N, INV, MUL, SPLIT_IN, SPLIT_OUT>(p, tile % gx, tile / gx);
    __syncthreads();   // the next tile reuses the shared-memory tile
  }
}

template <int N, bool INV, int MUL, bool SPLIT_IN, int SPLIT_OUT>
static int launch_strided_t(const StridedParams& p, int nouter, cudaStream_t st) {
  constexpr int LINES = StridedTraits<N>::LINES;
  constexpr int NT = StridedTraits<N>::NT;
  size_t smem = (size_t)N * LINES * sizeof(float2);
  auto kern = c2c_strided_kernel<N, INV, MUL, SPLIT_IN, SPLIT_OUT>;
  if (smem > 48 * 1024) SMK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (p.ncols % LINES) { set_error("strided pass: column count must be a multiple of the tile width"); return SMK_ERR_ARG; }
  dim3 grid(p.ncols / LINES, nouter);
  if constexpr (SPLIT_OUT == OUT_PEER) {
    const int cap = p.x_sms;
    if (cap > 0 && (long long)cap < (long long)grid.x * grid.y) {
      auto pkern = c2c_strided_persistent_kernel<N, INV, MUL, SPLIT_IN, SPLIT_OUT>;
      if (smem > 48 * 1024) SMK_CUDA_OK(cudaFuncSetAttribute(pkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      pkern<<<cap, NT, smem, st>>>(p);
      SMK_CUDA_OK(cudaGetLastError());
      return SMK_OK;
    }
  }
  kern<<<grid, NT, smem, st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

template <int N>
static int launch_strided_n(bool inv, int mul, const StridedParams& p, int nouter, bool peer, cudaStream_t st) {
  const bool si = p.ain.nsplit < N, so = p.aout.nsplit < N;
  if (!inv) {
    if (mul != MUL_NONE || si || peer) { set_error("forward pass: unsupported variant"); return SMK_ERR_ARG; }
    return so ? launch_strided_t<N, false, MUL_NONE, false, OUT_SPLIT>(p, nouter, st)
              : launch_strided_t<N, false, MUL_NONE, false, OUT_PLAIN>(p, nouter, st);
  }
  if (peer) {
    if (si || !so) { set_error("peer pass: unsupported addressing"); return SMK_ERR_ARG; }
    switch (mul) {
      case MUL_TABLE: return launch_strided_t<N, true, MUL_TABLE, false, OUT_PEER>(p, nouter, st);
      case MUL_ETA: return launch_strided_t<N, true, MUL_ETA, false, OUT_PEER>(p, nouter, st);
      case MUL_VEL: return launch_strided_t<N, true, MUL_VEL, false, OUT_PEER>(p, nouter, st);
    }
    set_error("peer pass: bad multiplier mode");
    return SMK_ERR_ARG;
  }
  if (so) { set_error("inverse pass: unsupported variant"); return SMK_ERR_ARG; }
  switch (mul) {
    case MUL_NONE:
      return si ? launch_strided_t<N, true, MUL_NONE, true, OUT_PLAIN>(p, nouter, st)
                : launch_strided_t<N, true, MUL_NONE, false, OUT_PLAIN>(p, nouter, st);
    case MUL_TABLE: if (si) break; return launch_strided_t<N, true, MUL_TABLE, false, OUT_PLAIN>(p, nouter, st);
    case MUL_ETA: if (si) break; return launch_strided_t<N, true, MUL_ETA, false, OUT_PLAIN>(p, nouter, st);
    case MUL_VEL: if (si) break; return launch_strided_t<N, true, MUL_VEL, false, OUT_PLAIN>(p, nouter, st);
  }
  set_error("bad multiplier mode / addressing combination");
  return SMK_ERR_ARG;
}

#define SMK_STRIDED_SIZES(X) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(2560)

int strided_tile_width(int n) {
#define X(N_) if (n == N_) return StridedTraits<N_>::LINES;
  SMK_STRIDED_SIZES(X)
#undef X
  return 0;
}

bool strided_size_supported(int n) {
#define X(N_) if (n == N_) return true;
  SMK_STRIDED_SIZES(X)
#undef X
  return false;
}

int launch_c2c_strided(int N, bool inverse, int mul_mode, const float2* in, float2* out, PassAddr ain, PassAddr aout,
                       int nouter, int ncols, int wcols, const MulArgs& mul, const float2* tw, cudaStream_t st,
                       float2* const* peers, int npeers, int x_sms) {
  make_fastdiv(ain);
  make_fastdiv(aout);
  StridedParams p{in, out, ain, aout, ncols, wcols, mul, tw, {nullptr}, nouter, x_sms};
  p.mul.vs_hi = (float)mul.vscale;                                 // dgrowth0 as a float32 pair (MUL_VEL)
  p.mul.vs_lo = (float)(mul.vscale - (double)p.mul.vs_hi);
  if (peers) {
    if (npeers > SMK_MAX_RANKS) { set_error("too many ranks for the fused exchange"); return SMK_ERR_ARG; }
    for (int i = 0; i < npeers; ++i) p.peer[i] = peers[i];
  }
  switch (N) {
#define X(N_) case N_: return launch_strided_n<N_>(inverse, mul_mode, p, nouter, peers != nullptr, st);
    SMK_STRIDED_SIZES(X)
#undef X
  }
  set_error("unsupported x/y length " + std::to_string(N));
  return SMK_ERR_UNSUPPORTED;
}

struct R2CParams {
  const float* in;
  float2* out;
  long long nlines;
  int pitch;
  const float2* tw;   // W_NZ
  uint64_t seed;
  long long cell0;    // global index of the first cell of line 0 (Philox counter base)
};

template <int M, bool PHILOX>
// (as for c2r_z_kernel: four 8-line CTAs per SM at NZ = 1536 need 128 registers or fewer)
__global__ void __launch_bounds__(ZFwd<M>::NT, (M == 768) ? 32 / ZFwd<M>::LINES : 1) r2c_z_kernel(R2CParams p) {
  using ZT = ZFwd<M>;
  constexpr int LINES = ZFwd<M>::LINES, NT = ZFwd<M>::NT, LP = ZFwd<M>::LP;
  extern __shared__ float2 sm[];   // [LINES][LP]
  const long long line0 = (long long)blockIdx.x * LINES;
  // ---- load (or draw) the real lines as M float2 each
  if (PHILOX) {
    // one Philox call -> 4 normals = 2 float2 = cells 4c..4c+3 of the line (M is even)
    for (int idx = threadIdx.x; idx < LINES * (M / 2); idx += NT) {
      int line = idx / (M / 2), c = idx - line * (M / 2);
      if (line0 + line < p.nlines) {
        long long cell = p.cell0 + (line0 + line) * (2LL * M) + 4LL * c;
        float4 g = philox_normal4(p.seed, (uint64_t)cell >> 2);
        sm[line * LP + ZT::idx(2 * c)] = make_float2(g.x, g.y);
        sm[line * LP + ZT::idx(2 * c + 1)] = make_float2(g.z, g.w);
      }
    }
  } else {
    const float2* in2 = reinterpret_cast<const float2*>(p.in);
    for (int line = threadIdx.x >> 5; line < LINES; line += NT / 32) {      // one warp per line: no index division
      if (line0 + line < p.nlines) {
        const float2* src = in2 + (line0 + line) * M;
#pragma unroll 8
        for (int n = threadIdx.x & 31; n < M; n += 32) sm[line * LP + ZT::idx(n)] = __ldg(src + n);
      }
    }
  }
  __syncthreads();
  // ---- M-point complex FFT of z[n] = x[2n] + i x[2n+1], output re-sorted to natural order
  z_tile_fft<M, false>(sm, p.tw);
  // ---- store with the real-transform post-step fused in: X[k] = (Z[k]+conj(Z[M-k]))/2 - (i/2) w^k (Z[k]-conj(Z[M-k]))
  //      for the pair (k, M-k); one warp per line, both global writes coalesced; row pads are zeroed so that later
  //      passes may stream whole rows
  for (int line = threadIdx.x >> 5; line < LINES; line += NT / 32) {
    if (line0 + line < p.nlines) {
      const float2* row = sm + line * LP;
      float2* dst = p.out + (line0 + line) * p.pitch;
#pragma unroll 4
      for (int k = threadIdx.x & 31; k <= M / 2; k += 32) {
        float2 a = row[ZT::nat(k)], b = (k == 0) ? a : row[ZT::nat(M - k)];
        float2 w = __ldg(p.tw + k);                       // exp(-2 pi i k / NZ)
        float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));   // (a + conj b)/2
        float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));   // (a - conj b)/2
        float2 t = cmul(w, d);
        float2 mit = make_float2(t.y, -t.x);              // -i t
        dst[k] = cadd(e, mit);
        if (k != M - k) dst[M - k] = make_float2(e.x - mit.x, -e.y + mit.y);
      }
      for (int k = M + 1 + (threadIdx.x & 31); k < p.pitch; k += 32) dst[k] = make_float2(0.f, 0.f);
    }
  }
}

struct C2RParams {
  const float2* in;
  float* out;
  long long nlines;
  int pitch;
  const float2* tw;
  float norm;       // nx*ny*nz as float (the reference divides: box /= NX*NY*NZ)
  double* stats;    // [2] sum, sum of squares (atomically accumulated) or null
  int discard_in;   // chained mode: the input is an L2-resident scratch slot; drop its lines after reading (no write-back)
};

template <int M>
// two CTAs of (M >= 512: 98 KB) share an SM: the register allocation has to leave room for both
__global__ void __launch_bounds__(ZInv<M>::NT, (M >= 512) ? 32 / ZInv<M>::LINES : 1) c2r_z_kernel(C2RParams p) {
  using ZT = ZInv<M>;
  using CT = C2RTraits<M>;
  constexpr int LINES = ZInv<M>::LINES, NT = ZInv<M>::NT, LP = CT::LP;
  extern __shared__ float2 sm[];
  __shared__ double red[2][NT / 32];
  const long long line0 = (long long)blockIdx.x * LINES;
  // ---- load with the real-transform pre-step fused in: one warp per line reads X[k] (ascending) and X[M-k]
  //      (descending), both coalesced, and writes Z[k] = A + iB, Z[M-k] = conj(A) + i conj(B) with
  //      A = X[k] + conj(X[M-k]), B = (X[k] - conj(X[M-k])) w^-k.  The imaginary parts of the DC and Nyquist bins
  //      are ignored, as FFTW's / pocketfft's c2r do.
  constexpr int KI = (M / 2 + 1 + 31) / 32;   // (k, M-k) pairs per lane and line
  float2 w[KI];                               // w^k of the lane's k: the same for every line of the warp
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const int k = (threadIdx.x & 31) + 32 * i;
    w[i] = (k <= M / 2) ? __ldg(p.tw + k) : make_float2(0.f, 0.f);
  }
  for (int line = threadIdx.x >> 5; line < LINES; line += NT / 32) {
    const bool ok = line0 + line < p.nlines;
    const float2* src = p.in + (line0 + line) * p.pitch;
    float2* row = sm + line * LP;
    // all loads of the line first (2*KI independent 8-byte loads per lane in flight), arithmetic afterwards.  Issuing
    // the loads of BOTH lines of the warp before any arithmetic (104 registers) measured slower: 0.692 against 0.655 ms.
    float2 a[KI], b[KI];
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = (threadIdx.x & 31) + 32 * i;
      a[i] = b[i] = make_float2(0.f, 0.f);
      if (k <= M / 2 && ok) { a[i] = __ldg(src + k); b[i] = __ldg(src + M - k); }
    }
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = (threadIdx.x & 31) + 32 * i;
      if (k <= M / 2) {
        float2 av = a[i], bv = b[i];
        if (k == 0) { av.y = 0.f; bv.y = 0.f; }
        // with b' = conj(X[M-k]): A = X[k] + b', iB = (X[k] - b') * (i w^-k); Z[k] = A + iB, Z[M-k] = conj(A - iB)
        const float2 bc = make_float2(bv.x, -bv.y);
        const float2 iwc = make_float2(w[i].y, w[i].x);    // i exp(+2 pi i k / NZ) = (-sin, cos): the table's (cos, -sin) swapped
        const float2 A = cadd(av, bc);
        const float2 iB = cmul(csub(av, bc), iwc);
        row[CT::idx(k)] = cadd(A, iB);
        if (k != 0 && k != M - k) {
          const float2 T = csub(A, iB);
          row[CT::idx(M - k)] = make_float2(T.x, -T.y);
        }
      }
    }
  }
  __syncthreads();
  if (p.discard_in) {
    // every read of this CTA's input rows has completed (values are in shared memory): the rows are dead, so tell L2
    // to drop them instead of writing them back to HBM when they are evicted
    const long long nl = p.nlines - line0 < LINES ? p.nlines - line0 : LINES;
    const char* base = reinterpret_cast<const char*>(p.in + line0 * p.pitch);
    const int n128 = (int)(nl * p.pitch * 8 / 128);
    for (int i = threadIdx.x; i < n128; i += NT) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + 128LL * i) : "memory");
  }
  // ---- store x[2n], x[2n+1] = z[n] / N, accumulate sum and sum of squares
  float2 s1 = make_float2(0.f, 0.f), s2 = s1;   // even / odd cells of the thread's share, added up at the end
  const float rnorm = __frcp_rn(p.norm);
  float2* out2 = reinterpret_cast<float2*>(p.out);
  if constexpr (CT::TAIL) {
    using P = typename CT::P;
    c2r_stages<CT, 0, P::S - 2, LINES, NT, LP>(sm, p.tw);
    auto emit = [&](int line, int n, float2 val) {
      if (line0 + line < p.nlines) {
        const float2 z = fdiv_fast2(val, p.norm, rnorm);
        __stcs(out2 + (line0 + line) * M + n, z);      // written once, read much later: streaming store
        s1 = cadd(s1, z);
        s2 = cfma(z, z, s2);
      }
    };
    c2r_tail<CT, LINES, NT, LP>(sm, emit);
  } else if constexpr (CT::FUSE) {
    using P = typename CT::P;
    constexpr int RL = CT::RL, NB = CT::NB;
    c2r_stages<CT, 0, P::S - 1, LINES, NT, LP>(sm, p.tw);
    // last stage: lane = natural output index n0 (butterfly pos(n0) / RL), outputs n0 + q NB go straight to HBM
    for (int line = threadIdx.x >> 5; line < LINES; line += NT / 32) {
      if (line0 + line < p.nlines) {
        float2* dst = out2 + (line0 + line) * M;
        const float2* row = sm + line * LP;
#pragma unroll 2
        for (int n0 = threadIdx.x & 31; n0 < NB; n0 += 32) {
          float2 v[RL];
          c2r_last_butterfly<CT>(row, n0, v);
#pragma unroll
          for (int q = 0; q < RL; ++q) {
            const float2 z = fdiv_fast2(v[q], p.norm, rnorm);
            __stcs(dst + n0 + q * NB, z);            // written once, read much later: streaming store
            s1 = cadd(s1, z);
            s2 = cfma(z, z, s2);
          }
        }
      }
    }
  } else {
    z_tile_fft<M, true>(sm, p.tw);
    for (int line = threadIdx.x >> 5; line < LINES; line += NT / 32) {
      if (line0 + line < p.nlines) {
        float2* dst = out2 + (line0 + line) * M;
#pragma unroll 8
        for (int n = threadIdx.x & 31; n < M; n += 32) {
          const float2 z = fdiv_fast2(sm[line * LP + ZT::nat(n)], p.norm, rnorm);
          __stcs(dst + n, z);            // written once, read much later: streaming store
          s1 = cadd(s1, z);
          s2 = cfma(z, z, s2);
        }
      }
    }
  }
  if (p.stats != nullptr) {
    double d1 = (double)s1.x + (double)s1.y, d2 = (double)s2.x + (double)s2.y;
    for (int o = 16; o > 0; o >>= 1) {
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = d1; red[1][warp] = d2; }
    __syncthreads();
    if (warp == 0) {
      d1 = lane < NT / 32 ? red[0][lane] : 0.0;
      d2 = lane < NT / 32 ? red[1][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
      }
      if (lane == 0) { atomicAdd(p.stats, d1); atomicAdd(p.stats + 1, d2); }
    }
  }
}

#define SMK_Z_HALF_SIZES(X) X(4) X(8) X(12) X(16) X(32) X(48) X(64) X(128) X(256) X(512) X(768)

bool z_size_supported(int nz) {
#define X(M_) if (nz == 2 * M_) return true;
  SMK_Z_HALF_SIZES(X)
#undef X
  return false;
}

template <int M>
static int launch_r2c_t(const R2CParams& p, bool philox, cudaStream_t st) {
  size_t smem = (size_t)ZFwd<M>::LINES * ZFwd<M>::LP * sizeof(float2);
  unsigned grid = (unsigned)((p.nlines + ZFwd<M>::LINES - 1) / ZFwd<M>::LINES);
  if (philox) {
    auto kern = r2c_z_kernel<M, true>;
    if (smem > 48 * 1024) SMK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, ZFwd<M>::NT, smem, st>>>(p);
  } else {
    auto kern = r2c_z_kernel<M, false>;
    if (smem > 48 * 1024) SMK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, ZFwd<M>::NT, smem, st>>>(p);
  }
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

int launch_r2c_z(int NZ, const float* in, float2* out, long long nlines, int pitch, const float2* tw, bool philox,
                 uint64_t seed, long long cell0, cudaStream_t st) {
  R2CParams p{in, out, nlines, pitch, tw, seed, cell0};
  switch (NZ / 2) {
#define X(M_) case M_: return launch_r2c_t<M_>(p, philox, st);
    SMK_Z_HALF_SIZES(X)
#undef X
  }
  set_error("unsupported z length " + std::to_string(NZ));
  return SMK_ERR_UNSUPPORTED;
}

// (A persistent form of this pass with bulk-copy tile loads -- cp.async.bulk of the 50 KB tile into a staging buffer
// under an mbarrier while the previous tile's butterflies run, two CTAs of 192 threads per SM -- was built and measured
// in round 2: parity green, 0.958 ms against 0.772 ms.  The pass is bound by instruction issue and shared-memory
// traffic inside the SM, not by exposed DRAM latency: the staging hop and the smaller number of resident warps cost
// more than the overlap gains.  profiles/README.md.)
template <int M>
static int launch_c2r_t(const C2RParams& p, cudaStream_t st) {
  size_t smem = (size_t)ZInv<M>::LINES * C2RTraits<M>::LP * sizeof(float2);
  unsigned grid = (unsigned)((p.nlines + ZInv<M>::LINES - 1) / ZInv<M>::LINES);
  auto kern = c2r_z_kernel<M>;
  if (smem > 48 * 1024) SMK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, ZInv<M>::NT, smem, st>>>(p);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

int launch_c2r_z(int NZ, const float2* in, float* out, long long nlines, int pitch, const float2* tw, float norm,
                 double* stats, cudaStream_t st, bool discard_in) {
  C2RParams p{in, out, nlines, pitch, tw, norm, stats, discard_in ? 1 : 0};
  switch (NZ / 2) {
#define X(M_) case M_: return launch_c2r_t<M_>(p, st);
    SMK_Z_HALF_SIZES(X)
#undef X
  }
  set_error("unsupported z length " + std::to_string(NZ));
  return SMK_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------ stand-alone Philox fill
__global__ void philox_fill_kernel(float* out, long long ncells, uint64_t seed, long long cell0) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; 4 * q < ncells; q += stride) {
    float4 g = philox_normal4(seed, (uint64_t)(cell0 + 4 * q) >> 2);
    if (4 * q + 3 < ncells) {
      reinterpret_cast<float4*>(out)[q] = g;
    } else {
      float t[4] = {g.x, g.y, g.z, g.w};
      for (int i = 0; 4 * q + i < ncells; ++i) out[4 * q + i] = t[i];
    }
  }
}

int launch_philox_fill(float* out, long long ncells, uint64_t seed, long long cell0, cudaStream_t st) {
  if (cell0 % 4 != 0) { set_error("philox fill: slab offset must be a multiple of 4 cells"); return SMK_ERR_ARG; }
  long long nq = (ncells + 3) / 4;
  int blocks = (int)((nq + 255) / 256 < 148 * 16 ? (nq + 255) / 256 : 148 * 16);
  if (blocks < 1) blocks = 1;
  philox_fill_kernel<<<blocks, 256, 0, st>>>(out, ncells, seed, cell0);
  SMK_CUDA_OK(cudaGetLastError());
  return SMK_OK;
}

}  // namespace smk
