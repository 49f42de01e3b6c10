// C ABI of libsmk.so (include/smk.h): context, plans, and the box-synthesis entry points.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef SMK_YZ_GROUP_DEFAULT
#define SMK_YZ_GROUP_DEFAULT 0
#endif
#ifndef SMK_YZ_STREAMS_DEFAULT
#define SMK_YZ_STREAMS_DEFAULT 2
#endif

#include <vector>

#include "smk_internal.h"
#include "smk_ztile.cuh"

namespace smk {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace smk

using namespace smk;

struct smk_ctx {
  int nx, ny, nz, nzh, pitch;
  int rank, nranks, nxl, nyl;
  double dcell;
  cudaStream_t stream;
  float2 *tw_x = nullptr, *tw_y = nullptr, *tw_z = nullptr;
  float *kx = nullptr, *ky = nullptr, *kz = nullptr;
  float2* work = nullptr;      // [nxl][ny][pitch] (== boxk size)
  double* stats = nullptr;     // scratch for the host wrapper
  size_t bytes = 0;
  // fused exchange (smk_exchange_*): receive buffers of this rank and the peers' views of theirs
  int nxbuf = 0;
  float2* xbuf[4] = {nullptr, nullptr, nullptr, nullptr};
  float2* xpeer[4][SMK_MAX_RANKS] = {};
  bool xconnected[4] = {false, false, false, false};
  int x_sms = 0;                // persistent fused-exchange x pass on this many CTAs (smk_exchange_set_sms; 0 = off)
  // y<->z chaining through L2 (see chain_setup): planes per group (0 = off), internal streams, ring of scratch slots
  int yz_group = 0, yz_streams = 0;
  bool yz_discard = true;
  float2* yz_scratch = nullptr;
  cudaStream_t yz_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t yz_fork = nullptr, yz_join[4] = {nullptr, nullptr, nullptr, nullptr};
  void* scratch = nullptr;      // small persistent scratch for kernels of other translation units (smk_ctx_scratch)
  size_t scratch_bytes = 0;
  // optional per-pass CUDA-event timing (smk_timing_*): events are recorded around every pass kernel
  bool timing = false;
  std::vector<cudaEvent_t> ev;       // pool
  std::vector<int> ev_pass;          // pass id of the interval that STARTS at event i (-1 = none)
  size_t ev_used = 0;
};

enum { PASS_R2C_Z = 0, PASS_FWD_Y, PASS_FWD_X, PASS_INV_X, PASS_INV_Y, PASS_C2R_Z, PASS_FWD_ZY, PASS_INV_YZ, PASS_COUNT };
static_assert(PASS_COUNT == SMK_NPASSES, "include/smk.h: SMK_NPASSES");

// record an event; `pass` names the kernel launched right after it (-1: closes the previous interval only)
static void tmark(smk_ctx* c, int pass) {
  if (!c->timing) return;
  if (c->ev_used == c->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev.push_back(e);
    c->ev_pass.push_back(-1);
  }
  cudaEventRecord(c->ev[c->ev_used], c->stream);
  c->ev_pass[c->ev_used] = pass;
  ++c->ev_used;
}

static int make_twiddles(int n, float2** dptr, size_t* bytes) {
  std::vector<float2> h(n);
  for (int k = 0; k < n; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)n;
    h[k] = make_float2((float)cos(a), (float)sin(a));
  }
  SMK_CUDA_OK(cudaMalloc(dptr, n * sizeof(float2)));
  SMK_CUDA_OK(cudaMemcpy(*dptr, h.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
  *bytes += n * sizeof(float2);
  return SMK_OK;
}

// W_NZ plus the z passes' compact first-stage table (smk_ztile.cuh)
static int make_z_twiddles(int nz, float2** dptr, size_t* bytes) {
  std::vector<float2> h;
  z_twiddle_table(nz, h);
  SMK_CUDA_OK(cudaMalloc(dptr, h.size() * sizeof(float2)));
  SMK_CUDA_OK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
  *bytes += h.size() * sizeof(float2);
  return SMK_OK;
}

// np.float32(np.fft.fftfreq(n) * 2 * k_ny) / rfftfreq -- make_boxes.py:299-304
static int make_ktable(int n, bool rfft, double dcell, float** dptr, size_t* bytes) {
  double k_ny = M_PI / dcell;
  int len = rfft ? n / 2 + 1 : n;
  std::vector<float> h(len);
  for (int i = 0; i < len; ++i) {
    double f;
    if (rfft) f = (double)i / ((double)n * 1.0);
    else f = (double)((i < (n + 1) / 2) ? i : i - n) / ((double)n * 1.0);
    h[i] = (float)(f * 2 * k_ny);
  }
  SMK_CUDA_OK(cudaMalloc(dptr, len * sizeof(float)));
  SMK_CUDA_OK(cudaMemcpy(*dptr, h.data(), len * sizeof(float), cudaMemcpyHostToDevice));
  *bytes += len * sizeof(float);
  return SMK_OK;
}

// ---- y<->z chaining through L2.  The y and z passes of one transform both work on whole x planes, so they can be run
// plane group by plane group with the intermediate (one group of [ny][pitch] planes) in a small scratch slot that is
// overwritten group after group and therefore stays resident in the 126 MB L2: the y pass's writes and the z pass's
// reads never reach HBM, and a 3-D transform moves 16 B per real cell instead of 24.  Groups are issued round robin on
// a few internal streams (each with its own slot) so that the tail of one group's kernels overlaps the next group's.
// Options yz_group = planes per group (0 disables: the default, the chained form measured slower than the plain passes),
// yz_streams = 1..4, yz_persist = 1 pins the slots in L2 with an access-policy window.
// ---- library options (smk_set_option): switches for parity tests and for the alternatives that lost their sweeps.
// Process-wide, read when a context is created (yz_*) or a call is made (skewers_kernel, qso_exact); no environment
// variable is read anywhere in the library.
#include <map>
#include <mutex>
static std::mutex g_opt_mutex;
static std::map<std::string, int>& options() {
  static std::map<std::string, int> o = {{"skewers_kernel", 0}, {"qso_exact", 0}, {"yz_group", SMK_YZ_GROUP_DEFAULT},
                                         {"yz_streams", SMK_YZ_STREAMS_DEFAULT}, {"yz_persist", 0}, {"yz_discard", 1}};
  return o;
}
int smk_option(const char* name) {
  std::lock_guard<std::mutex> lock(g_opt_mutex);
  auto it = options().find(name);
  return it == options().end() ? 0 : it->second;
}
static int env_int(const char* name, int) { return smk_option(name); }

static int chain_setup(smk_ctx* c) {
  int g = env_int("yz_group", 0);
  if (g <= 0) return SMK_OK;
  if (g > c->nxl) g = c->nxl;
  int ns = env_int("yz_streams", 0);
  ns = ns < 1 ? 1 : (ns > 4 ? 4 : ns);
  const size_t slot = (size_t)g * c->ny * c->pitch * sizeof(float2);
  SMK_CUDA_OK(cudaMalloc(&c->yz_scratch, slot * ns));
  c->bytes += slot * ns;
  SMK_CUDA_OK(cudaEventCreateWithFlags(&c->yz_fork, cudaEventDisableTiming));
  for (int s = 0; s < ns; ++s) {
    SMK_CUDA_OK(cudaStreamCreateWithFlags(&c->yz_stream[s], cudaStreamNonBlocking));
    SMK_CUDA_OK(cudaEventCreateWithFlags(&c->yz_join[s], cudaEventDisableTiming));
  }
  if (env_int("yz_persist", 0)) {
    int dev = 0, maxwin = 0, maxpersist = 0;
    SMK_CUDA_OK(cudaGetDevice(&dev));
    SMK_CUDA_OK(cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    SMK_CUDA_OK(cudaDeviceGetAttribute(&maxpersist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    size_t want = slot * ns;
    if (want > (size_t)maxpersist) want = (size_t)maxpersist;
    SMK_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    for (int s = 0; s < ns; ++s) {
      cudaStreamAttrValue v{};
      v.accessPolicyWindow.base_ptr = (char*)c->yz_scratch + slot * s;
      v.accessPolicyWindow.num_bytes = slot < (size_t)maxwin ? slot : (size_t)maxwin;
      v.accessPolicyWindow.hitRatio = 1.0f;
      v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      SMK_CUDA_OK(cudaStreamSetAttribute(c->yz_stream[s], cudaStreamAttributeAccessPolicyWindow, &v));
    }
  }
  c->yz_group = g;
  c->yz_streams = ns;
  c->yz_discard = env_int("yz_discard", 1) != 0;
  return SMK_OK;
}

static int chain_fork(smk_ctx* c) {
  SMK_CUDA_OK(cudaEventRecord(c->yz_fork, c->stream));
  for (int s = 0; s < c->yz_streams; ++s) SMK_CUDA_OK(cudaStreamWaitEvent(c->yz_stream[s], c->yz_fork, 0));
  return SMK_OK;
}

static int chain_join(smk_ctx* c) {
  for (int s = 0; s < c->yz_streams; ++s) {
    SMK_CUDA_OK(cudaEventRecord(c->yz_join[s], c->yz_stream[s]));
    SMK_CUDA_OK(cudaStreamWaitEvent(c->stream, c->yz_join[s], 0));
  }
  return SMK_OK;
}

// inverse: y pass (input addressing `ain`, outer = local x plane) -> slot -> z pass -> out_slab
static int chain_inverse_yz(smk_ctx* c, const float2* in, PassAddr ain, float* out_slab, double* stats) {
  const int G = c->yz_group;
  const size_t slot_elems = (size_t)G * c->ny * c->pitch;
  const float norm = (float)((double)c->nx * c->ny * c->nz);
  PassAddr aout{(long long)c->ny * c->pitch, 0, (long long)c->pitch, c->ny};
  MulArgs m{};
  tmark(c, PASS_INV_YZ);
  int rc = chain_fork(c);
  for (int x0 = 0, g = 0; x0 < c->nxl && rc == SMK_OK; x0 += G, ++g) {
    const int n = c->nxl - x0 < G ? c->nxl - x0 : G;
    const int s = g % c->yz_streams;
    float2* slot = c->yz_scratch + slot_elems * s;
    rc = launch_c2c_strided(c->ny, true, MUL_NONE, in + (long long)x0 * ain.outer_stride, slot, ain, aout, n, c->pitch,
                            c->nzh, m, c->tw_y, c->yz_stream[s]);
    if (rc) break;
    rc = launch_c2r_z(c->nz, slot, out_slab + (size_t)x0 * c->ny * c->nz, (long long)n * c->ny, c->pitch, c->tw_z, norm,
                      stats, c->yz_stream[s], c->yz_discard);
  }
  if (rc) return rc;
  rc = chain_join(c);
  tmark(c, -1);
  return rc;
}

// forward: z pass (noise or box_slab) -> slot -> y pass -> `out` with addressing `aout`
static int chain_forward_zy(smk_ctx* c, const float* box_slab, uint64_t seed, float2* out, PassAddr aout) {
  const int G = c->yz_group;
  const size_t slot_elems = (size_t)G * c->ny * c->pitch;
  const long long cell0 = (long long)c->rank * c->nxl * c->ny * c->nz;
  PassAddr ain{(long long)c->ny * c->pitch, 0, (long long)c->pitch, c->ny};
  MulArgs m{};
  tmark(c, PASS_FWD_ZY);
  int rc = chain_fork(c);
  for (int x0 = 0, g = 0; x0 < c->nxl && rc == SMK_OK; x0 += G, ++g) {
    const int n = c->nxl - x0 < G ? c->nxl - x0 : G;
    const int s = g % c->yz_streams;
    float2* slot = c->yz_scratch + slot_elems * s;
    const size_t off = (size_t)x0 * c->ny * c->nz;
    rc = launch_r2c_z(c->nz, box_slab ? box_slab + off : nullptr, slot, (long long)n * c->ny, c->pitch, c->tw_z,
                      box_slab == nullptr, seed, cell0 + (long long)off, c->yz_stream[s]);
    if (rc) break;
    rc = launch_c2c_strided(c->ny, false, MUL_NONE, slot, out + (long long)x0 * aout.outer_stride, ain, aout, n, c->pitch,
                            c->nzh, m, c->tw_y, c->yz_stream[s]);
  }
  if (rc) return rc;
  rc = chain_join(c);
  tmark(c, -1);
  return rc;
}

static int ctx_allocate(smk_ctx* c) {
  int rc;
  if ((rc = make_twiddles(c->nx, &c->tw_x, &c->bytes))) return rc;
  if ((rc = make_twiddles(c->ny, &c->tw_y, &c->bytes))) return rc;
  if ((rc = make_z_twiddles(c->nz, &c->tw_z, &c->bytes))) return rc;
  if ((rc = make_ktable(c->nx, false, c->dcell, &c->kx, &c->bytes))) return rc;
  if ((rc = make_ktable(c->ny, false, c->dcell, &c->ky, &c->bytes))) return rc;
  if ((rc = make_ktable(c->nz, true, c->dcell, &c->kz, &c->bytes))) return rc;
  size_t wbytes = (size_t)c->nxl * c->ny * c->pitch * sizeof(float2);
  SMK_CUDA_OK(cudaMalloc(&c->work, wbytes));
  c->bytes += wbytes;
  SMK_CUDA_OK(cudaMalloc(&c->stats, 2 * SMK_NPRODUCTS * sizeof(double)));
  return chain_setup(c);
}

// entry points that need the FFT plan refuse a light context (smk_ctx_create_light: stream + scratch only)
#define SMK_NEED_PLAN(c, what)                                                                  \
  do {                                                                                          \
    if (!(c) || (c)->nx <= 0) { set_error(what ": needs a context made by smk_ctx_create"); return SMK_ERR_ARG; } \
  } while (0)

extern "C" {

const char* smk_last_error(void) { return g_err.c_str(); }

int smk_set_option(const char* name, int value) {
  std::lock_guard<std::mutex> lock(g_opt_mutex);
  auto it = name ? options().find(name) : options().end();
  if (it == options().end()) { set_error(std::string("smk_set_option: unknown option ") + (name ? name : "(null)")); return SMK_ERR_ARG; }
  it->second = value;
  return SMK_OK;
}
int smk_version(void) { return 100; }

int smk_ctx_create(smk_ctx** out, int nx, int ny, int nz, double dcell, int rank, int nranks, void* stream) {
  if (!out || nx <= 0 || ny <= 0 || nz <= 0 || nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("smk_ctx_create: bad argument");
    return SMK_ERR_ARG;
  }
  if (!strided_size_supported(nx) || !strided_size_supported(ny) || !z_size_supported(nz)) {
    set_error("smk_ctx_create: unsupported box dimensions " + std::to_string(nx) + "x" + std::to_string(ny) + "x" +
              std::to_string(nz));
    return SMK_ERR_UNSUPPORTED;
  }
  if (nx % nranks || ny % nranks) {
    set_error("smk_ctx_create: nx and ny must be divisible by nranks");
    return SMK_ERR_ARG;
  }
  smk_ctx* c = new smk_ctx();
  c->nx = nx; c->ny = ny; c->nz = nz; c->nzh = nz / 2 + 1;
  c->pitch = (c->nzh + 15) / 16 * 16;
  c->rank = rank; c->nranks = nranks; c->nxl = nx / nranks; c->nyl = ny / nranks;
  c->dcell = dcell;
  c->stream = (cudaStream_t)stream;
  int rc = ctx_allocate(c);
  if (rc) {                  // nothing of a half-built context survives a failure (smk_ctx_destroy frees what exists)
    const std::string why = g_err;
    smk_ctx_destroy(c);
    set_error(why);
    return rc;
  }
  *out = c;
  return SMK_OK;
}

int smk_ctx_create_light(smk_ctx** out, void* stream) {
  if (!out) { set_error("smk_ctx_create_light: null argument"); return SMK_ERR_ARG; }
  smk_ctx* c = new smk_ctx();
  c->nx = c->ny = c->nz = c->nzh = c->pitch = c->nxl = c->nyl = 0;
  c->rank = 0; c->nranks = 1; c->dcell = 0.0;
  c->stream = (cudaStream_t)stream;
  *out = c;
  return SMK_OK;
}

int smk_ctx_destroy(smk_ctx* c) {
  if (!c) return SMK_OK;
  cudaFree(c->tw_x); cudaFree(c->tw_y); cudaFree(c->tw_z);
  cudaFree(c->kx); cudaFree(c->ky); cudaFree(c->kz);
  cudaFree(c->work); cudaFree(c->stats); cudaFree(c->yz_scratch); cudaFree(c->scratch);
  if (c->yz_fork) cudaEventDestroy(c->yz_fork);
  for (int s = 0; s < 4; ++s) {
    if (c->yz_join[s]) cudaEventDestroy(c->yz_join[s]);
    if (c->yz_stream[s]) cudaStreamDestroy(c->yz_stream[s]);
  }
  for (int b = 0; b < c->nxbuf; ++b) {
    for (int r = 0; r < c->nranks; ++r)
      if (c->xconnected[b] && r != c->rank && c->xpeer[b][r]) cudaIpcCloseMemHandle(c->xpeer[b][r]);
    cudaFree(c->xbuf[b]);
  }
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  delete c;
  return SMK_OK;
}

int smk_boxk_pitch(const smk_ctx* c) { return c->pitch; }
size_t smk_boxk_elems(const smk_ctx* c) { return (size_t)c->nx * c->nyl * c->pitch; }
size_t smk_box_elems(const smk_ctx* c) { return (size_t)c->nxl * c->ny * c->nz; }
size_t smk_workspace_bytes(const smk_ctx* c) { return c->bytes; }
int smk_set_stream(smk_ctx* c, void* stream) {
  c->stream = (cudaStream_t)stream;
  return SMK_OK;
}

int smk_sync(smk_ctx* c) {
  SMK_CUDA_OK(cudaStreamSynchronize(c->stream));
  return SMK_OK;
}

int smk_pk_weights(smk_ctx* c, const double* breaks, const double* coefs, int nint, float* wtable) {
  SMK_NEED_PLAN(c, "smk_pk_weights");
  if (!breaks || !coefs || nint < 1 || !wtable) { set_error("smk_pk_weights: bad argument"); return SMK_ERR_ARG; }
  PkParams p{breaks, coefs, nint, c->kx, c->ky, c->kz, c->nx, c->nyl, c->nzh, c->rank * c->nyl,
             (float)(c->dcell * c->dcell * c->dcell), wtable};
  return launch_pk_weights(p, c->stream);
}

int smk_pk_estimate(smk_ctx* c, const void* boxk, int nbins, double kmin, double kmax, double* sums) {
  SMK_NEED_PLAN(c, "smk_pk_estimate");
  if (!boxk || !sums) { set_error("smk_pk_estimate: null argument"); return SMK_ERR_ARG; }
  return launch_pk_estimate((const float2*)boxk, c->kx, c->ky, c->kz, c->nx, c->nyl, c->nzh, c->pitch, c->rank * c->nyl,
                            c->nz, nbins, kmin, kmax, sums, c->stream);
}

int smk_timing_enable(smk_ctx* c, int on) {
  c->timing = on != 0;
  c->ev_used = 0;
  return SMK_OK;
}

int smk_timing_collect(smk_ctx* c, double ms_sum[SMK_NPASSES], int count[SMK_NPASSES]) {
  for (int i = 0; i < PASS_COUNT; ++i) { ms_sum[i] = 0.0; count[i] = 0; }
  SMK_CUDA_OK(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i + 1 < c->ev_used; ++i) {
    int p = c->ev_pass[i];
    if (p < 0) continue;
    float ms = 0.f;
    SMK_CUDA_OK(cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
    ms_sum[p] += ms;
    count[p] += 1;
  }
  c->ev_used = 0;
  return SMK_OK;
}

int smk_noise_philox(smk_ctx* c, uint64_t seed, float* box_slab) {
  SMK_NEED_PLAN(c, "smk_noise_philox");
  long long ncells = (long long)c->nxl * c->ny * c->nz;
  return launch_philox_fill(box_slab, ncells, seed, (long long)c->rank * ncells, c->stream);
}

// ---- forward
int smk_fft_r2c_local(smk_ctx* c, const float* box_slab, uint64_t seed, void* sendbuf) {
  SMK_NEED_PLAN(c, "smk_fft_r2c_local");
  long long nlines = (long long)c->nxl * c->ny;
  long long cell0 = (long long)c->rank * nlines * c->nz;
  float2* tmp = (c->nranks == 1) ? (float2*)sendbuf : c->work;
  if (c->yz_group > 0) {
    PassAddr aout{(long long)c->nyl * c->pitch, (long long)c->nxl * c->nyl * c->pitch, (long long)c->pitch, c->nyl};
    return chain_forward_zy(c, box_slab, seed, (float2*)sendbuf, aout);
  }
  tmark(c, PASS_R2C_Z);
  int rc = launch_r2c_z(c->nz, box_slab, tmp, nlines, c->pitch, c->tw_z, box_slab == nullptr, seed, cell0, c->stream);
  if (rc) return rc;
  tmark(c, PASS_FWD_Y);
  // y pass: outer = local x plane; output into [dest][xl][yl][z] (dest = y / nyl)
  PassAddr ain{(long long)c->ny * c->pitch, 0, (long long)c->pitch, c->ny};
  PassAddr aout{(long long)c->nyl * c->pitch, (long long)c->nxl * c->nyl * c->pitch, (long long)c->pitch, c->nyl};
  MulArgs m{};
  rc = launch_c2c_strided(c->ny, false, MUL_NONE, tmp, (float2*)sendbuf, ain, aout, c->nxl, c->pitch, c->nzh, m, c->tw_y,
                          c->stream);
  tmark(c, -1);
  return rc;
}

int smk_fft_r2c_finish(smk_ctx* c, const void* recvbuf, void* boxk) {
  SMK_NEED_PLAN(c, "smk_fft_r2c_finish");
  // x pass on [nx][nyl][pitch]: outer = local y
  PassAddr a{(long long)c->pitch, 0, (long long)c->nyl * c->pitch, c->nx};
  MulArgs m{};
  tmark(c, PASS_FWD_X);
  int rc = launch_c2c_strided(c->nx, false, MUL_NONE, (const float2*)recvbuf, (float2*)boxk, a, a, c->nyl, c->pitch, c->nzh, m,
                              c->tw_x, c->stream);
  tmark(c, -1);
  return rc;
}

int smk_fft_r2c(smk_ctx* c, const float* box_slab, uint64_t seed, void* boxk) {
  if (c->nranks != 1) { set_error("smk_fft_r2c: single-rank entry point; use _local/_finish"); return SMK_ERR_ARG; }
  int rc = smk_fft_r2c_local(c, box_slab, seed, boxk);
  if (rc) return rc;
  return smk_fft_r2c_finish(c, boxk, boxk);
}

// ---- inverse
int smk_synth_c2r_local(smk_ctx* c, void* boxk, int product, const float* wtable, int store_p0, double dgrowth0,
                        void* sendbuf) {
  SMK_NEED_PLAN(c, "smk_synth_c2r_local");
  if (product < 0 || product >= SMK_NPRODUCTS) { set_error("bad product id"); return SMK_ERR_ARG; }
  MulArgs m{};
  int mode;
  if (product <= SMK_P0) {
    if (!wtable) { set_error("spectral weight table required for products PLN1..P0"); return SMK_ERR_ARG; }
    mode = MUL_TABLE;
    m.wt = wtable;
    m.wt_n_stride = (long long)c->nyl * c->nzh;
    m.wt_outer_stride = c->nzh;
    m.store_back = (product == SMK_P0 && store_p0) ? (float2*)boxk : nullptr;
  } else {
    static const int FA[] = {0, 1, 2, 0, 0, 1, 0, 1, 2};
    static const int FB[] = {0, 1, 2, 1, 2, 2, 0, 0, 0};
    mode = product >= SMK_VX ? MUL_VEL : MUL_ETA;
    m.kn = c->kx; m.ko = c->ky; m.kc = c->kz;
    m.outer0 = c->rank * c->nyl;
    m.fa = FA[product - SMK_ETA_XX];
    m.fb = FB[product - SMK_ETA_XX];
    m.vscale = dgrowth0;
  }
  PassAddr a{(long long)c->pitch, 0, (long long)c->nyl * c->pitch, c->nx};
  tmark(c, PASS_INV_X);
  int rc = launch_c2c_strided(c->nx, true, mode, (const float2*)boxk, (float2*)sendbuf, a, a, c->nyl, c->pitch, c->nzh, m,
                              c->tw_x, c->stream);
  tmark(c, -1);
  return rc;
}

// ---- fused exchange: the inverse x pass stores straight into the peers' receive buffers over NVLink
int smk_exchange_create(smk_ctx* c, int nbuf) {
  SMK_NEED_PLAN(c, "smk_exchange_create");
  if (nbuf < 1 || nbuf > 4 || c->nxbuf) { set_error("smk_exchange_create: nbuf must be 1..4, once per ctx"); return SMK_ERR_ARG; }
  size_t bytes = smk_boxk_elems(c) * sizeof(float2);
  for (int b = 0; b < nbuf; ++b) {
    SMK_CUDA_OK(cudaMalloc(&c->xbuf[b], bytes));
    SMK_CUDA_OK(cudaMemset(c->xbuf[b], 0, bytes));
    c->bytes += bytes;
  }
  c->nxbuf = nbuf;
  return SMK_OK;
}

int smk_exchange_handle(smk_ctx* c, int buf, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (buf < 0 || buf >= c->nxbuf) { set_error("smk_exchange_handle: bad buffer index"); return SMK_ERR_ARG; }
  cudaIpcMemHandle_t h;
  SMK_CUDA_OK(cudaIpcGetMemHandle(&h, c->xbuf[buf]));
  memcpy(handle, &h, 64);
  return SMK_OK;
}

int smk_exchange_connect(smk_ctx* c, int buf, const unsigned char* handles) {
  if (buf < 0 || buf >= c->nxbuf || !handles) { set_error("smk_exchange_connect: bad argument"); return SMK_ERR_ARG; }
  if (c->nranks > SMK_MAX_RANKS) { set_error("smk_exchange_connect: too many ranks"); return SMK_ERR_ARG; }
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank) { c->xpeer[buf][r] = c->xbuf[buf]; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    void* ptr = nullptr;
    SMK_CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->xpeer[buf][r] = (float2*)ptr;
  }
  c->xconnected[buf] = true;
  return SMK_OK;
}

int smk_exchange_set_sms(smk_ctx* c, int nsm) {
  if (nsm < 0) { set_error("smk_exchange_set_sms: negative CTA count"); return SMK_ERR_ARG; }
  c->x_sms = nsm;
  return SMK_OK;
}

void* smk_exchange_ptr(smk_ctx* c, int buf) { return (buf >= 0 && buf < c->nxbuf) ? c->xbuf[buf] : nullptr; }

int smk_synth_c2r_local_p2p(smk_ctx* c, void* boxk, int product, const float* wtable, int store_p0, double dgrowth0,
                            int buf) {
  SMK_NEED_PLAN(c, "smk_synth_c2r_local_p2p");
  if (buf < 0 || buf >= c->nxbuf || !c->xconnected[buf]) { set_error("smk_synth_c2r_local_p2p: exchange buffer not connected"); return SMK_ERR_ARG; }
  if (c->nranks < 2 || c->nxl < 2) { set_error("smk_synth_c2r_local_p2p: needs >= 2 ranks and >= 2 planes per rank"); return SMK_ERR_ARG; }
  if (product < 0 || product >= SMK_NPRODUCTS) { set_error("bad product id"); return SMK_ERR_ARG; }
  MulArgs m{};
  int mode;
  if (product <= SMK_P0) {
    if (!wtable) { set_error("spectral weight table required for products PLN1..P0"); return SMK_ERR_ARG; }
    mode = MUL_TABLE;
    m.wt = wtable;
    m.wt_n_stride = (long long)c->nyl * c->nzh;
    m.wt_outer_stride = c->nzh;
    m.store_back = (product == SMK_P0 && store_p0) ? (float2*)boxk : nullptr;
  } else {
    static const int FA[] = {0, 1, 2, 0, 0, 1, 0, 1, 2};
    static const int FB[] = {0, 1, 2, 1, 2, 2, 0, 0, 0};
    mode = product >= SMK_VX ? MUL_VEL : MUL_ETA;
    m.kn = c->kx; m.ko = c->ky; m.kc = c->kz;
    m.outer0 = c->rank * c->nyl;
    m.fa = FA[product - SMK_ETA_XX];
    m.fb = FB[product - SMK_ETA_XX];
    m.vscale = dgrowth0;
  }
  // tiled receive layout on rank d: [src][kz tile][y_l][x_l][LX] (LX = kz columns per x-pass tile); this rank is `src`.
  // One x-pass tile (one y_l, one kz tile) owns, per destination, a contiguous run of nxl*LX elements.
  const long long chunk = (long long)c->nxl * c->nyl * c->pitch;
  const int LX = strided_tile_width(c->nx);
  float2* peers[SMK_MAX_RANKS];
  for (int r = 0; r < c->nranks; ++r) peers[r] = c->xpeer[buf][r] + (long long)c->rank * chunk;
  PassAddr ain{(long long)c->pitch, 0, (long long)c->nyl * c->pitch, c->nx};
  PassAddr aout{(long long)c->nxl * LX, 0, 0, c->nxl};                        // nsplit = rows per destination
  aout.tile_width = LX;
  aout.tile_stride = (long long)c->nyl * c->nxl * LX;
  tmark(c, PASS_INV_X);
  int rc = launch_c2c_strided(c->nx, true, mode, (const float2*)boxk, nullptr, ain, aout, c->nyl, c->pitch, c->nzh, m,
                              c->tw_x, c->stream, peers, c->nranks, c->x_sms);
  tmark(c, -1);
  return rc;
}

int smk_synth_c2r_finish_p2p(smk_ctx* c, int buf, float* out_slab, double* stats) {
  SMK_NEED_PLAN(c, "smk_synth_c2r_finish_p2p");
  if (buf < 0 || buf >= c->nxbuf) { set_error("smk_synth_c2r_finish_p2p: bad buffer index"); return SMK_ERR_ARG; }
  // y pass reading the tiled receive layout [src][kz tile][y_l][x_l][LX]: y = src*nyl + y_l
  const long long chunk = (long long)c->nxl * c->nyl * c->pitch;
  const int LX = strided_tile_width(c->nx);
  PassAddr ain{(long long)LX, chunk, (long long)c->nxl * LX, c->nyl};
  ain.tile_width = LX;
  ain.tile_stride = (long long)c->nyl * c->nxl * LX;
  if (c->yz_group > 0) return chain_inverse_yz(c, c->xbuf[buf], ain, out_slab, stats);
  PassAddr aout{(long long)c->ny * c->pitch, 0, (long long)c->pitch, c->ny};
  MulArgs m{};
  tmark(c, PASS_INV_Y);
  int rc = launch_c2c_strided(c->ny, true, MUL_NONE, c->xbuf[buf], c->work, ain, aout, c->nxl, c->pitch, c->nzh, m,
                              c->tw_y, c->stream);
  if (rc) return rc;
  float norm = (float)((double)c->nx * c->ny * c->nz);
  tmark(c, PASS_C2R_Z);
  rc = launch_c2r_z(c->nz, c->work, out_slab, (long long)c->nxl * c->ny, c->pitch, c->tw_z, norm, stats, c->stream);
  tmark(c, -1);
  return rc;
}

int smk_synth_c2r_finish(smk_ctx* c, void* recvbuf, float* out_slab, double* stats) {
  SMK_NEED_PLAN(c, "smk_synth_c2r_finish");
  // y pass: input [src][xl][yl][z] (y = src*nyl + yl), output [xl][ny][pitch]
  PassAddr ain{(long long)c->nyl * c->pitch, (long long)c->nxl * c->nyl * c->pitch, (long long)c->pitch, c->nyl};
  if (c->yz_group > 0) return chain_inverse_yz(c, (const float2*)recvbuf, ain, out_slab, stats);
  PassAddr aout{(long long)c->ny * c->pitch, 0, (long long)c->pitch, c->ny};
  float2* tmp = (c->nranks == 1) ? (float2*)recvbuf : c->work;
  MulArgs m{};
  tmark(c, PASS_INV_Y);
  int rc = launch_c2c_strided(c->ny, true, MUL_NONE, (const float2*)recvbuf, tmp, ain, aout, c->nxl, c->pitch, c->nzh, m,
                              c->tw_y, c->stream);
  if (rc) return rc;
  float norm = (float)((double)c->nx * c->ny * c->nz);
  tmark(c, PASS_C2R_Z);
  rc = launch_c2r_z(c->nz, tmp, out_slab, (long long)c->nxl * c->ny, c->pitch, c->tw_z, norm, stats, c->stream);
  tmark(c, -1);
  return rc;
}

int smk_synth_c2r(smk_ctx* c, void* boxk, int product, const float* wtable, int store_p0, double dgrowth0,
                  float* out_slab, double* stats) {
  if (c->nranks != 1) { set_error("smk_synth_c2r: single-rank entry point; use _local/_finish"); return SMK_ERR_ARG; }
  int rc = smk_synth_c2r_local(c, boxk, product, wtable, store_p0, dgrowth0, c->work);
  if (rc) return rc;
  return smk_synth_c2r_finish(c, c->work, out_slab, stats);
}

int smk_make_boxes_host(smk_ctx* c, const float* noise_host, uint64_t seed, const float* const wtables_host[4],
                        double dgrowth0, float* const out_host[SMK_NPRODUCTS], double sigma_out[SMK_NPRODUCTS]) {
  SMK_NEED_PLAN(c, "smk_make_boxes_host");
  if (c->nranks != 1) { set_error("smk_make_boxes_host: single rank only"); return SMK_ERR_ARG; }
  size_t ncell = (size_t)c->nx * c->ny * c->nz, nk = smk_boxk_elems(c), nw = (size_t)c->nx * c->ny * c->nzh;
  // every resource of the call lives in this guard: any early return (SMK_CUDA_OK, argument errors) releases all of it
  struct Guard {
    float2* boxk = nullptr;
    float* wt = nullptr;
    float* box[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    ~Guard() {
      cudaFree(boxk); cudaFree(wt); cudaFree(box[0]); cudaFree(box[1]);
      if (copy_stream) cudaStreamDestroy(copy_stream);
      for (int i = 0; i < 2; ++i) {
        if (done[i]) cudaEventDestroy(done[i]);
        if (copied[i]) cudaEventDestroy(copied[i]);
      }
    }
  } g;
  float2*& boxk = g.boxk;
  float*& wt = g.wt;
  float** box = g.box;
  cudaStream_t& copy_stream = g.copy_stream;
  cudaEvent_t* done = g.done;
  cudaEvent_t* copied = g.copied;
  SMK_CUDA_OK(cudaMalloc(&boxk, nk * sizeof(float2)));
  SMK_CUDA_OK(cudaMalloc(&wt, nw * sizeof(float)));
  SMK_CUDA_OK(cudaMalloc(&box[0], ncell * sizeof(float)));
  SMK_CUDA_OK(cudaMalloc(&box[1], ncell * sizeof(float)));
  SMK_CUDA_OK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    SMK_CUDA_OK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    SMK_CUDA_OK(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
  }
  SMK_CUDA_OK(cudaMemsetAsync(boxk, 0, nk * sizeof(float2), c->stream));
  SMK_CUDA_OK(cudaMemsetAsync(c->stats, 0, 2 * SMK_NPRODUCTS * sizeof(double), c->stream));
  int rc = SMK_OK;
  if (noise_host) {
    SMK_CUDA_OK(cudaMemcpyAsync(box[0], noise_host, ncell * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    rc = smk_fft_r2c(c, box[0], seed, boxk);
  } else {
    rc = smk_fft_r2c(c, nullptr, seed, boxk);
  }
  int nb = 0;
  for (int p = 0; p < SMK_NPRODUCTS && rc == SMK_OK; ++p) {
    bool need_p0 = (p == SMK_P0);           // boxk must become boxk*P0 for every later product
    if (!out_host[p] && !need_p0) continue;
    if (p <= SMK_P0) {
      if (!wtables_host || !wtables_host[p]) { set_error("missing weight table"); rc = SMK_ERR_ARG; break; }
      SMK_CUDA_OK(cudaMemcpyAsync(wt, wtables_host[p], nw * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    int b = nb & 1;
    SMK_CUDA_OK(cudaStreamWaitEvent(c->stream, copied[b], 0));     // buffer free again?
    rc = smk_synth_c2r(c, boxk, p, wt, 1, dgrowth0, box[b], c->stats + 2 * p);
    if (rc) break;
    SMK_CUDA_OK(cudaEventRecord(done[b], c->stream));
    if (out_host[p]) {
      SMK_CUDA_OK(cudaStreamWaitEvent(copy_stream, done[b], 0));
      SMK_CUDA_OK(cudaMemcpyAsync(out_host[p], box[b], ncell * sizeof(float), cudaMemcpyDeviceToHost, copy_stream));
      SMK_CUDA_OK(cudaEventRecord(copied[b], copy_stream));
    }
    ++nb;
  }
  double hstats[2 * SMK_NPRODUCTS];
  // drain both streams before the guard frees anything, whatever happened above
  cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(copy_stream);
  if (rc) return rc;
  SMK_CUDA_OK(e1);
  SMK_CUDA_OK(e2);
  SMK_CUDA_OK(cudaMemcpy(hstats, c->stats, sizeof(hstats), cudaMemcpyDeviceToHost));
  for (int p = 0; p < SMK_NPRODUCTS; ++p) {
    double s1 = hstats[2 * p], s2 = hstats[2 * p + 1], n = (double)ncell;
    double var = s2 / n - (s1 / n) * (s1 / n);
    if (sigma_out) sigma_out[p] = var > 0 ? sqrt(var) : 0.0;
    if ((out_host[p] || p == SMK_P0) && (!(s2 > 0.0) || isnan(s2))) {
      set_error("box " + std::to_string(p) + " is null or NaN");   // make_boxes.py:100-105
      return SMK_ERR_NULL_BOX;
    }
  }
  return SMK_OK;
}

}  // extern "C"

cudaStream_t smk_ctx_stream(const smk_ctx* ctx) { return ctx ? ctx->stream : (cudaStream_t)0; }

void* smk_ctx_scratch(smk_ctx* c, size_t bytes) {
  if (c->scratch_bytes >= bytes) return c->scratch;
  if (c->scratch) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->scratch);
    c->scratch = nullptr;
    c->scratch_bytes = 0;
  }
  if (cudaMalloc(&c->scratch, bytes) != cudaSuccess) {
    smk::set_error("smk_ctx_scratch: cudaMalloc failed");
    c->scratch = nullptr;
    return nullptr;
  }
  c->scratch_bytes = bytes;
  return c->scratch;
}
