// Internal (C++) declarations shared by the translation units of libsmk.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/smk.h"

#define SMK_MAX_RANKS 16

namespace smk {

void set_error(const std::string& msg);

#define SMK_CUDA_OK(expr)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      smk::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
      return SMK_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

// Addressing of one strided complex pass.  A tile is LINES consecutive z-columns (col0..)
// of one `outer` index; point n of the transform sits at
//   base + outer*outer_stride + (n / nsplit)*hi_stride + (n % nsplit)*lo_stride + col
// (nsplit == N gives a plain stride; the two-level form addresses the [src][xl][yl][z]
// layout that an all-to-all leaves behind without an unpack pass).
struct PassAddr {
  long long outer_stride;
  long long hi_stride;
  long long lo_stride;
  int nsplit;
  unsigned magic = 0, shift = 0;   // n / nsplit == umulhi(n, magic) >> shift for 0 <= n < 65536
  // column addressing: plain rows (tile_width == 0: offset = column) or the tiled exchange layout, where groups of
  // tile_width kz columns are stored tile_stride elements apart: offset = (col / tile_width) * tile_stride + col % tile_width
  int tile_width = 0;
  long long tile_stride = 0;
};

// n / nsplit == umulhi(n, magic) for 0 <= n < 65536 and 2 <= nsplit <= 65536 (magic = floor(2^32 / d) + 1:
// the error n * (magic * d - 2^32) stays below 2^32)
inline void make_fastdiv(PassAddr& a) {
  a.shift = 0;
  a.magic = a.nsplit >= 2 ? (unsigned)((1ull << 32) / (unsigned)a.nsplit + 1) : 0u;
}

struct MulArgs {
  // spectral weight table W[n][outer][col] (float) with its own strides, or null
  const float* wt;
  long long wt_n_stride;
  long long wt_outer_stride;
  // optional store-back of (in * W) to this array (same addressing as the input)
  float2* store_back;
  // k tables for the k-factor products (float32, reference rounding): kn = k along the
  // transformed axis, ko = k along the outer axis, kc = k along the column (z) axis
  const float* kn;
  const float* ko;
  const float* kc;
  int outer0;      // global index of outer==0 (multi-GPU y offset)
  int fa, fb;      // axes of the factor: 0 = n axis (x), 1 = outer axis (y), 2 = column axis (z)
  double vscale;   // dgrowth0 of the velocity products (float64 scalar of make_boxes.py:321)
  float vs_hi, vs_lo;   // the same as an unevaluated float32 pair, filled in by launch_c2c_strided
};

enum MulMode { MUL_NONE = 0, MUL_TABLE = 1, MUL_ETA = 2, MUL_VEL = 3 };

int launch_c2c_strided(int N, bool inverse, int mul_mode, const float2* in, float2* out, PassAddr ain, PassAddr aout,
                       int nouter, int ncols, int wcols, const MulArgs& mul, const float2* tw, cudaStream_t st,
                       float2* const* peers = nullptr, int npeers = 0, int x_sms = 0);

int launch_r2c_z(int NZ, const float* in, float2* out, long long nlines, int pitch, const float2* tw,
                 bool philox, uint64_t seed, long long cell0, cudaStream_t st);

int launch_c2r_z(int NZ, const float2* in, float* out, long long nlines, int pitch, const float2* tw, float norm,
                 double* stats, cudaStream_t st, bool discard_in = false);

int launch_philox_fill(float* out, long long ncells, uint64_t seed, long long cell0, cudaStream_t st);

struct PkParams {
  const double* breaks;   // [nint+1]
  const double* coefs;    // [4][nint], highest power first (scipy PPoly convention)
  int nint;
  const float *kx, *ky, *kz;
  int nx, nyl, nzh, y0;
  float vcell;
  float* out;             // [nx][nyl][nzh]
};
int launch_pk_weights(const PkParams& p, cudaStream_t st);

int launch_pk_estimate(const float2* boxk, const float* kx, const float* ky, const float* kz, int nx, int nyl, int nzh,
                       int pitch, int y0, int nz, int nbins, double kmin, double kmax, double* out, cudaStream_t st);

bool strided_size_supported(int n);
int strided_tile_width(int n);   // kz columns per tile of the strided pass of length n
bool z_size_supported(int nz);

}  // namespace smk

// stream of a context (defined in smk_capi.cu; smk_ctx is opaque to the other translation units)
cudaStream_t smk_ctx_stream(const smk_ctx* ctx);
// persistent device scratch of at least `bytes` owned by the context (grown on demand; nullptr + error set on failure)
void* smk_ctx_scratch(smk_ctx* ctx, size_t bytes);
// device twiddle table W_nfft[k] = exp(-2 pi i k / nfft) of the 1-D transforms, cached per (device, nfft)
int smk_tw1d(int nfft, const float2** out);
// value of a library option (smk_set_option; 0 for an unknown name)
int smk_option(const char* name);
