/* libsmk.so -- C ABI of the B200-native SaclayMocks hot path.
 *
 * The reference (igmhub/SaclayMocks) is pure Python and has no FFI; its seams for this path
 * are the in-Python calls listed next to each entry point below.
 * Conventions:
 *   - every function returns 0 on success or a negative SMK_ERR_* code; the message is
 *     available from smk_last_error() (thread-local); nothing throws across the boundary;
 *   - array arguments are caller-owned DEVICE pointers unless the name ends in _host;
 *     the library owns only the plan/workspaces inside smk_ctx;
 *   - one smk_ctx per GPU (per rank); calls are asynchronous on the ctx stream unless
 *     stated otherwise;
 *   - complex data are interleaved float pairs (numpy complex64 layout).
 *
 * Layout of the k-space box ("boxk"): [nx_k][ny_k][pitch] complex64 with
 * pitch = smk_boxk_pitch() >= nz/2+1 (rows padded to a multiple of 16 complex = 128 B).
 * Single rank: nx_k = nx, ny_k = ny.  With R ranks the real boxes are x-slabs
 * [nx/R][ny][nz] and boxk is kept transposed as y-slabs: [nx][ny/R][pitch].
 */
#ifndef SMK_H
#define SMK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMK_OK 0
#define SMK_ERR_ARG (-1)
#define SMK_ERR_CUDA (-2)
#define SMK_ERR_UNSUPPORTED (-3)
#define SMK_ERR_NULL_BOX (-4) /* all-zero or NaN output: make_boxes.py:63-70,100-105 raises ValueError */

typedef struct smk_ctx smk_ctx;

/* Products of make_boxes.py (bin/make_boxes.py:242-431), in the order the reference writes them. */
enum smk_product {
  SMK_PLN1 = 0, /* boxln_1: boxk * Pln1            make_boxes.py:247 */
  SMK_PLN2 = 1, /* boxln_2                         make_boxes.py:265 */
  SMK_PLN3 = 2, /* boxln_3                         make_boxes.py:274 */
  SMK_P0 = 3,   /* box (delta): boxk * P0, kept    make_boxes.py:289-291 */
  SMK_ETA_XX = 4, /* boxk*P0 * kx kx / k^2         make_boxes.py:324 */
  SMK_ETA_YY = 5,
  SMK_ETA_ZZ = 6,
  SMK_ETA_XY = 7,
  SMK_ETA_XZ = 8,
  SMK_ETA_YZ = 9,
  SMK_VX = 10, /* boxk*P0 * -i kx/k^2 * H0 * dD/dz(0)   make_boxes.py:403 */
  SMK_VY = 11,
  SMK_VZ = 12,
  SMK_NPRODUCTS = 13
};

const char* smk_last_error(void);
int smk_version(void);

/* Library options (process-wide).  The library reads no environment variable; these are the switches the parity tests
 * use and the alternatives that lost their measurements:
 *   "skewers_kernel"  0 = staged (TMA) gather with the global-memory kernel for what it hands back [default],
 *                     1 = global-memory gather for everything, 2 = one pixel per thread (no register blocking);
 *   "qso_exact"       1 = smk_draw_qso uses the reference-arithmetic kernel with Philox draws as well;
 *   "yz_group", "yz_streams", "yz_persist", "yz_discard"  y/z passes chained through L2 in groups of yz_group planes
 *                     (0 = off, default), read by smk_ctx_create.
 * Returns SMK_ERR_ARG for an unknown name. */
int smk_set_option(const char* name, int value);

/* ---- context / plan.  Replaces the pyfftw plan + wisdom handling of make_boxes.py:192-203.
 * dcell = cell size in Mpc/h (make_boxes.py -pixel).  stream = cudaStream_t (NULL = default stream).
 * rank/nranks describe the slab decomposition; nranks > 1 requires nx % nranks == 0 and ny % nranks == 0. */
int smk_ctx_create(smk_ctx** ctx, int nx, int ny, int nz, double dcell, int rank, int nranks, void* stream);
/* A context without an FFT plan: it carries a stream (and the library's small scratch) for the entry points that do not
 * transform boxes -- smk_skewers, smk_skewers_fgpa, smk_smallscale, smk_fgpa, smk_p1d, smk_draw_qso.  (The per-quasar
 * loop of make_spectra.py / merge_spectra.py has no plan either.)  The box entry points refuse it with SMK_ERR_ARG. */
int smk_ctx_create_light(smk_ctx** ctx, void* stream);
int smk_ctx_destroy(smk_ctx* ctx);
int smk_boxk_pitch(const smk_ctx* ctx);           /* complex elements per kz row */
size_t smk_boxk_elems(const smk_ctx* ctx);        /* complex elements of this rank's boxk */
size_t smk_box_elems(const smk_ctx* ctx);         /* float elements of this rank's real x-slab */
size_t smk_workspace_bytes(const smk_ctx* ctx);   /* bytes the ctx holds in HBM */
int smk_sync(smk_ctx* ctx);                       /* cudaStreamSynchronize(ctx stream) */
int smk_set_stream(smk_ctx* ctx, void* stream);   /* later calls launch on this cudaStream_t (host-side switch only) */

/* ---- per-pass device timing for the roofline report (bench.py).  When enabled, CUDA events are recorded on the
 * ctx stream around every FFT pass kernel; smk_timing_collect synchronises the stream and returns, per pass
 * (0 r2c-z, 1 forward-y, 2 forward-x, 3 inverse-x [with the fused multiply], 4 inverse-y, 5 c2r-z, 6 forward z+y
 * chained through L2, 7 inverse y+z chained through L2 [one interval per transform]), the summed duration in ms and
 * the number of intervals since the last collect. */
#define SMK_NPASSES 8
int smk_timing_enable(smk_ctx* ctx, int on);
int smk_timing_collect(smk_ctx* ctx, double ms_sum[SMK_NPASSES], int count[SMK_NPASSES]);

/* ---- spectral weight table of one product on the GPU (replaces the interpolate_pk.py / merge_pk.py stage,
 * bin/interpolate_pk.py:17-26, 63-77): wtable[x][y_local][kz] = float32(sqrt(float32(max(S(|k|), 0)) / Vcell)) with S
 * the cubic spline given in piecewise-polynomial form: breaks[nint+1] (ascending), coefs[4][nint] (highest power
 * first, scipy.interpolate.PPoly convention), both device float64; the end intervals extrapolate like FITPACK.
 * wtable: device float [nx][ny/R][nz/2+1], directly usable as the `wtable` argument of smk_synth_c2r. */
int smk_pk_weights(smk_ctx* ctx, const double* breaks, const double* coefs, int nint, float* wtable);

/* ---- float64 1-D FFT for LogNormalP (py/SaclayMocks/powerspectrum.py:145-200: P(k) -> xi(r) -> ln(1 + xi) -> P_ln(k),
 * two np.fft.fft of 2^20 and 2^19 points): out = unnormalised forward DFT of in, n a power of two; in / out / work are
 * distinct device arrays of n complex128 (interleaved doubles). */
int smk_fft1d_f64(smk_ctx* ctx, int n, const double* in, double* out, double* work);

/* ---- 3-D power spectrum estimator on the GPU (the estimator the statistical acceptance
 * tests use, P(k) = <|delta_k|^2> V / N^2): bins |boxk|^2 of this rank's k-slab of a forward transform (smk_fft_r2c of a
 * real box) into `nbins` equal bins of |k| in [kmin, kmax) (h/Mpc), weighting every mode with its Hermitian
 * multiplicity.  sums: device double [3][nbins], ACCUMULATED (zero first; all-reduce over ranks): sum of mult*|boxk|^2,
 * sum of mult, sum of mult*|k|. */
int smk_pk_estimate(smk_ctx* ctx, const void* boxk, int nbins, double kmin, double kmax, double* sums);

/* ---- white noise.  Replaces the np.random.normal plane loop of DrawGRF_boxk (make_boxes.py:46-48)
 * with Philox4x32-10 keyed by (seed, global cell index): identical for any slab decomposition. */
int smk_noise_philox(smk_ctx* ctx, uint64_t seed, float* box_slab);

/* ---- forward transform of DrawGRF_boxk (make_boxes.py:52-54): unnormalised r2c over (x,y,z).
 * box_slab: [nx/R][ny][nz] float (preserved).  If box_slab is NULL the noise is generated inside the
 * z-pass from (seed) and never touches HBM.  boxk: see layout above.
 * Multi-rank: the caller performs the slab exchange between smk_fft_r2c_local (z and y passes, output in
 * send layout) and smk_fft_r2c_finish (x pass); single rank: smk_fft_r2c does everything. */
int smk_fft_r2c(smk_ctx* ctx, const float* box_slab, uint64_t seed, void* boxk);
int smk_fft_r2c_local(smk_ctx* ctx, const float* box_slab, uint64_t seed, void* sendbuf);
int smk_fft_r2c_finish(smk_ctx* ctx, const void* recvbuf, void* boxk);

/* ---- one product of make_boxes.py: boxk *= W (or k-factor) followed by FFTandStore's arithmetic
 * (make_boxes.py:77-105): unnormalised c2r, /= nx*ny*nz, sum and sum of squares for sigma = np.std(box).
 * boxk is preserved, except that product SMK_P0 with store_p0 != 0 replaces it by boxk*P0 exactly like
 * make_boxes.py:289-291 (the eta and velocity products must then be fed that array).
 * wtable: float [nx_k][ny_k][nz/2+1] (unpadded rows, the HDU of P<NX>-<NY>-<NZ>.fits), required for
 * SMK_PLN1..SMK_P0, ignored otherwise.  dgrowth0 = dD/dz(z=0) (make_boxes.py:313), used by SMK_V*.
 * out_slab: [nx/R][ny][nz] float.  stats (device, 2 doubles, accumulated: zero them first): sum, sum of squares
 * of this rank's slab.
 * Multi-rank: smk_synth_c2r_local does the x pass into sendbuf; after the caller's all-to-all,
 * smk_synth_c2r_finish does the y and z passes from recvbuf. */
int smk_synth_c2r(smk_ctx* ctx, void* boxk, int product, const float* wtable, int store_p0, double dgrowth0,
                  float* out_slab, double* stats);
int smk_synth_c2r_local(smk_ctx* ctx, void* boxk, int product, const float* wtable, int store_p0, double dgrowth0,
                        void* sendbuf);
int smk_synth_c2r_finish(smk_ctx* ctx, void* recvbuf, float* out_slab, double* stats);

/* ---- fused exchange (multi-rank): instead of x pass -> send buffer -> all-to-all -> receive buffer, the x pass stores
 * every output row straight into the owning rank's receive buffer over NVLink (peer memory mapped through CUDA IPC).
 * smk_exchange_create allocates `nbuf` receive buffers (boxk-sized) inside the ctx; each rank publishes
 * smk_exchange_handle(buf) (a 64-byte cudaIpcMemHandle_t), gathers all ranks' handles (any host-side all-gather) and
 * calls smk_exchange_connect(buf, handles[nranks][64]).  smk_synth_c2r_local_p2p then replaces
 * smk_synth_c2r_local + all-to-all and smk_synth_c2r_finish_p2p replaces smk_synth_c2r_finish (the receive buffers use
 * a tiled layout [src][kz tile][y_l][x_l][tile width], in which each x-pass tile owns one contiguous run per
 * destination, so the peer stores are 512-B contiguous per warp).  The caller must order (a) a cross-rank barrier on
 * the streams between the two calls and (b) reuse of a buffer only after every rank has finished reading it (two
 * buffers used alternately with one barrier per product satisfy both). */
int smk_exchange_create(smk_ctx* ctx, int nbuf);
int smk_exchange_handle(smk_ctx* ctx, int buf, unsigned char handle[64]);
int smk_exchange_connect(smk_ctx* ctx, int buf, const unsigned char* handles);
void* smk_exchange_ptr(smk_ctx* ctx, int buf);
/* The fused x pass is bound by NVLink, not by the SMs: nsm > 0 runs it as a persistent kernel on nsm CTAs (several
 * tiles per CTA), which leaves the other SMs to the y / z passes of the previous product on another stream; 0 (default)
 * launches one CTA per tile. */
int smk_exchange_set_sms(smk_ctx* ctx, int nsm);
int smk_synth_c2r_local_p2p(smk_ctx* ctx, void* boxk, int product, const float* wtable, int store_p0, double dgrowth0,
                            int buf);
int smk_synth_c2r_finish_p2p(smk_ctx* ctx, int buf, float* out_slab, double* stats);

/* Host-buffer convenience wrapper of the whole DrawGRF_boxk + 13 x FFTandStore chain (the call a
 * make_boxes.py replacement makes): noise from `noise_host` ([nx][ny][nz] float) or Philox(seed) when NULL,
 * weight tables from host, every requested product copied back to out_host[p] (NULL = skip that product).
 * sigma_out[p] = std of product p.  Synchronous.  Single rank only. */
int smk_make_boxes_host(smk_ctx* ctx, const float* noise_host, uint64_t seed, const float* const wtables_host[4],
                        double dgrowth0, float* const out_host[SMK_NPRODUCTS], double sigma_out[SMK_NPRODUCTS]);

/* ---- skewers: batched ReadSpec (make_spectra.py:90-139) over all quasars of one x-slab.
 * fields[10]: delta, eta_xx, eta_yy, eta_zz, eta_xy, eta_xz, eta_yz, vx, vy, vz; each a float slab
 * [nxs][ny][nz] holding global x-planes [ix0, ix0+nxs) (halo planes included).  fields[1..6] may be NULL
 * when rsd == 0, fields[7..9] when dla == 0.
 * Pixels of quasar q: i in [0, npix_forest[q]) on the global grid rvec[npix] (Mpc/h); pixel i belongs to this
 * call iff xmin < rvec[i]*X/R <= xmax (make_spectra.py:443-448).  qso_xyzr: 4 doubles per quasar (X,Y,Z,R_QSO).
 * Outputs are [nqso][npix] float, written only at the pixels owned by this call: delta_l, eta_par,
 * vpar (before the make_spectra.py:510 rescale); pixels owned but outside the forest get -1e6 / 0 / 0
 * (make_spectra.py:99-101).  Neighbour indices are clamped to the slab (documented deviation: the reference
 * gather is unchecked).
 * With a context (ctx != NULL), nz % 4 == 0 and 16-byte aligned fields the gather is staged: every warp brings the
 * box of cells its 128 pixels can touch into shared memory with 3-D TMA loads (box sized from dir_x_max / dir_y_max);
 * segments whose windows reach over a box edge, or that are more oblique than the box allows, are handed back to the
 * global-memory kernel inside the same call. */
typedef struct smk_geom {
  int nx, ny, nz;       /* full box */
  double dx, dy, dz;    /* cell size */
  double r0;            /* h*R(z0): distance of the box centre */
  int dmax;             /* half width of the Gaussian window in cells (3) */
  double pixel_step;    /* spacing of rvec in Mpc/h (make_spectra.py -pixel, 0.2); 0 = unknown / non-uniform */
  double dir_x_max;     /* largest |X| / R_QSO and |Y| / R_QSO of the catalogue (direction cosines of the most oblique */
  double dir_y_max;     /* sightline): sizes the shared-memory box of the staged gather; 0 = unknown (0.25 assumed) */
} smk_geom;

int smk_skewers(smk_ctx* ctx, const smk_geom* g, const float* const fields[10], int ix0, int nxs, double xmin,
                double xmax, int rsd, int dla, int nqso, const double* qso_xyzr, const int* npix_forest,
                const double* rvec, int npix, float* delta_l, float* eta_par, float* vpar);

/* How the calling thread's last smk_skewers / smk_skewers_fgpa call ran (synchronises its stream): segments = pieces
 * of 128 pixels launched (0 when the staged kernel was not used), handed_back = those the staged kernel left to the
 * global-memory kernel, box = staged box extent in cells (x, y, z).  Any output pointer may be NULL. */
int smk_skewers_stats(smk_ctx* ctx, long long* segments, long long* handed_back, int box[3]);

/* ---- small-scale field (merge_spectra.py:308-324): delta_s = irfft(rfft(noise) * filt)[:npix] * zscale.
 * noise: [nqso][nfft] float white noise, or NULL to draw Philox(seed, quasar index).  filt_rows: [nrows][nfft/2+1]
 * float = sqrt(max(P_miss(z_row,k),0)/pixsize); row_of_qso[q] selects the row (nearest tabulated z to z_eff); a
 * negative row marks an empty forest, whose delta_s row is set to 0 (merge_spectra.py:327-330).
 * zscale: [nqso][npix] or NULL; when NULL, sig_pix[npix]/sig_eff[q] is used (sigma_s(z)/sigma_s(z_eff)).
 * qso_ids: [nqso] Philox stream id of each quasar (NULL: the row index), so that every rank that holds a piece of
 * a sightline regenerates the same delta_s.  nfft must be a power of two in [256, 8192]. */
int smk_smallscale(smk_ctx* ctx, int nqso, int nfft, int npix, const float* noise, uint64_t seed,
                   const float* filt_rows, const int* row_of_qso, const float* sig_pix, const float* sig_eff,
                   const long long* qso_ids, float* delta_s);

/* ---- FGPA (util.py:421-433): F = exp(-a exp(b G (delta_l + delta_s + c eta_par))).  a,b,c,G are per-pixel
 * vectors [npix] (constant when -zfix is used).  delta_s / eta_par may be NULL (treated as 0). */
int smk_fgpa(smk_ctx* ctx, int nqso, int npix, const float* delta_l, const float* delta_s, const float* eta_par,
             const float* growthf, const float* a, const float* b, const float* c, float* flux);

/* ---- 1-D power spectrum estimator on the GPU (py/SaclayMocks/powerspectrum.py:204-238):
 * P1D_1spectrum(delta, DX) = |rfft(delta)|^2 DX / n of the window of nfft pixels that starts at first[q] in row q of
 * rows [nqso][npix], accumulated over the rows like ComputeP1D (every spectrum has the same k grid here, so the profile
 * histogram of the reference is a per-wavenumber mean).  Rows with fewer than nfft valid pixels (nvalid[q], counted from
 * first[q]) are skipped.  mean != NULL: the window is first turned into the contrast rows / mean[q] - 1 (transmission
 * F -> delta_F).  first / nvalid NULL: window at 0 / all npix pixels valid.
 * sums: device double [2][nfft/2+1], ACCUMULATED: sum of P(k_j) and sum of P(k_j)^2, k_j = 2 pi j / (nfft pixel);
 * nused: device counter of the rows that contributed.  nfft: power of two in [256, 8192]. */
int smk_p1d(smk_ctx* ctx, int nqso, int npix, int nfft, const float* rows, const int* first, const int* nvalid,
            const float* mean, double pixel, double* sums, unsigned long long* nused);

/* ---- skewers with the FGPA fused into the gather's epilogue (make_spectra.py:90-139 followed by util.py:421-433 for
 * the pixels this slab owns): smk_skewers, then flux = exp(-a exp(b G (delta_l + delta_s + c eta_par))) from the values
 * still in registers -- the rows delta_l / eta_par are written once and not read back.  delta_s [nqso][npix] comes from
 * smk_smallscale run BEFORE this call (NULL: no small-scale term); growthf, a, b, c are the [npix] vectors of smk_fgpa.
 * Pixels past the forest get flux of delta_l = -1e6 (= 1 exactly); pixels of other slabs are left untouched. */
int smk_skewers_fgpa(smk_ctx* ctx, const smk_geom* g, const float* const fields[10], int ix0, int nxs, double xmin,
                     double xmax, int rsd, int dla, int nqso, const double* qso_xyzr, const int* npix_forest,
                     const double* rvec, int npix, float* delta_l, float* eta_par, float* vpar, const float* delta_s,
                     const float* growthf, const float* a, const float* b, const float* c, float* flux);

/* ---- quasar drawing on device-resident boxes: the cell loop of bin/draw_qso.py for one
 * x-slab -- ptot(z) from the three lognormal boxes (draw_qso.py:228-251), cond1 & cond2 & cond3, the random position
 * inside the cell, (ra, dec) and the redshift-space shift of the quasar redshift (draw_qso.py:394-480).  All pointers
 * are device pointers.  Boxes are float [nxs][ny][nz] slabs whose first plane is global plane ix0.  Tables (float64):
 * x_axis/y_axis/z_axis = cell centres (draw_qso.py:208-210); chi/zt[ntab] = comoving distance (Mpc) and redshift of
 * util.cosmo (r_2_z); coef_z/coef_v[ncoef] = util.qso_lognormal_coef; dn_cell = n(z) per cell on the grid
 * dz_interp0 + i*delta_z (draw_qso.py:300-334); dg_z/dg_v[ndg] = etc/dgrowth.fits.  Scalars as computed by the host
 * set-up (saclaymocks_b200/qso.py).  Uniform variates: u1,u2,uz [nz][nxs][ny], ux [nz][nxs], uy [nz][ny] (the
 * reference's legacy NumPy stream, draw_qso.py:425-445) or all NULL to draw Philox4x32-10(seed, global cell).
 * Output: counters[0] = number of quasars selected (may exceed `capacity`: then only the first `capacity` records
 * were stored and the call must be repeated with a larger buffer), counters[1] = cells passing cond1; records
 * [capacity][8] doubles: key = (plane*nxs + ix)*ny + iy, z, z_rsd, ra, dec (degrees), X, Y, Z (Mpc/h), in arbitrary
 * order (sort by key for the reference's np.where order).  counters must be zeroed by the caller. */
typedef struct smk_qso_params {
  int nxs, ny, nz, ix0, nx_full;
  const float* boxln[3];
  const float* velo[3];
  int rsd;
  const double *x_axis, *y_axis, *z_axis;
  double dx, dy, dz;
  const double *chi, *zt;
  int ntab;
  const double *coef_z, *coef_v;
  int ncoef;
  const double* dn_cell;
  double dz_interp0, delta_z;
  const double *dg_z, *dg_v;
  int ndg;
  double dgrowth0, H0, h;
  double z1, z2, z3;
  double sigma_p[3];
  double norm, density_max;
  double z_min, z_max;
  double ra0, dec0, dra, ddec;          /* degrees */
  double cr0, sr0, cd0, sd0;            /* cos/sin of ra0, dec0 (radians), computed by the host like numpy */
  const double *u1, *u2, *ux, *uy, *uz;
  uint64_t seed;
} smk_qso_params;

int smk_draw_qso(smk_ctx* ctx, const smk_qso_params* p, int* counters, double* records, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* SMK_H */
