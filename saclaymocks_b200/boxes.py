"""Host side of the GRF box synthesis: mirrors DrawGRF_boxk / FFTandStore of bin/make_boxes.py:40-125
on top of libsmk.so.  torch is used for device memory and streams only."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import tables

PRODUCTS = _lib.PRODUCT_NAMES
WEIGHT_OF = {"boxln_1": "Pln1", "boxln_2": "Pln2", "boxln_3": "Pln3", "box": "P0"}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class BoxSynth(object):
    """Plan + workspaces for one (NX, NY, NZ) box on one GPU (one rank of an x-slab decomposition).

    Real boxes are x-slabs [NX/R, NY, NZ] float32; boxk is [NX, NY/R, pitch] complex64 (y-slabs, kz rows
    padded to `pitch`), see include/smk.h.
    """

    def __init__(self, NX, NY, NZ, dcell, device=None, rank=0, nranks=1):
        if not torch.cuda.is_available():
            raise _lib.SmkError("BoxSynth needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.NX, self.NY, self.NZ, self.dcell = int(NX), int(NY), int(NZ), float(dcell)
        self.rank, self.nranks = rank, nranks
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.nzh = self.NZ // 2 + 1
        self.nxl, self.nyl = self.NX // nranks, self.NY // nranks
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream()
            h = C.c_void_p()
            _lib.check(self.lib.smk_ctx_create(C.byref(h), self.NX, self.NY, self.NZ, self.dcell, rank, nranks,
                                               C.c_void_p(self.stream.cuda_stream)))
        self.h = h
        self.pitch = self.lib.smk_boxk_pitch(self.h)
        self.dgrowth0 = float(tables.dgrowth()[1][0])          # make_boxes.py:313

    def close(self):
        if getattr(self, "h", None):
            self.lib.smk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ per-pass device timing (bench.py roofline)
    PASS_NAMES = ("r2c_z", "fwd_y", "fwd_x", "inv_x", "inv_y", "c2r_z", "fwd_zy", "inv_yz")

    def timing_enable(self, on=True):
        _lib.check(self.lib.smk_timing_enable(self.h, int(on)))

    def timing_collect(self):
        ms = (C.c_double * len(self.PASS_NAMES))()
        n = (C.c_int * len(self.PASS_NAMES))()
        _lib.check(self.lib.smk_timing_collect(self.h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PASS_NAMES)}

    # ------------------------------------------------------------------ allocation helpers
    def empty_boxk(self):
        return torch.zeros((self.NX, self.nyl, self.pitch), dtype=torch.complex64, device=self.device)

    def empty_box(self):
        return torch.empty((self.nxl, self.NY, self.NZ), dtype=torch.float32, device=self.device)

    def upload_weights(self, W):
        """W: float32 [NX, NY/R, NZ/2+1] (this rank's y-slab of an HDU of P<NX>-<NY>-<NZ>.fits)."""
        t = torch.as_tensor(np.ascontiguousarray(W, dtype=np.float32)) if not torch.is_tensor(W) else W
        t = t.to(self.device, non_blocking=True).contiguous()
        assert tuple(t.shape) == (self.NX, self.nyl, self.nzh), t.shape
        return t

    def weight_table(self, name):
        """One HDU of P<NX>-<NY>-<NZ>.fits (this rank's ky rows) evaluated on the GPU from the P(k) spline
        (interpolate_pk.py:17-26): name in Pln1, Pln2, Pln3, P0."""
        from . import pk
        br, co = pk.ppoly(name)
        br_d = torch.as_tensor(br, device=self.device)
        co_d = torch.as_tensor(co, device=self.device)
        W = torch.empty((self.NX, self.nyl, self.nzh), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.smk_pk_weights(self.h, _ptr(br_d), _ptr(co_d), int(co.shape[1]), _ptr(W)))
        torch.cuda.current_stream(self.device).synchronize()      # br_d / co_d may be released after this
        return W

    # ------------------------------------------------------------------ DrawGRF_boxk (make_boxes.py:40-72)
    def noise_philox(self, seed):
        box = self.empty_box()
        _lib.check(self.lib.smk_noise_philox(self.h, C.c_uint64(seed), _ptr(box)))
        return box

    def draw_grf_boxk(self, seed=0, noise=None):
        """White noise -> unnormalised r2c.  noise: device float32 [NX,NY,NZ] (e.g. the reference's MT19937
        array for parity runs) or None to draw Philox(seed) inside the z pass.  Single rank."""
        assert self.nranks == 1
        boxk = self.empty_boxk()
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous()
            assert tuple(noise.shape) == (self.NX, self.NY, self.NZ)
        _lib.check(self.lib.smk_fft_r2c(self.h, _ptr(noise), C.c_uint64(seed), _ptr(boxk)))
        return boxk

    # ------------------------------------------------------------------ multiply + FFTandStore arithmetic
    def synth(self, boxk, name, wtable=None, store_p0=True, out=None, stats=None):
        """box = irfftn(boxk * factor(name)) / N  (make_boxes.py:247-429 + :86-92).  Returns (box, stats) with
        stats a device tensor [sum, sum of squares].  For name == 'box' and store_p0, boxk becomes boxk*P0."""
        assert self.nranks == 1
        pid = _lib.PRODUCT_ID[name]
        if out is None:
            out = self.empty_box()
        if stats is None:
            stats = torch.zeros(2, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.smk_synth_c2r(self.h, _ptr(boxk), pid, _ptr(wtable), int(bool(store_p0)),
                                          C.c_double(self.dgrowth0), _ptr(out), _ptr(stats)))
        return out, stats

    def power_spectrum(self, box, nbins=30, kmin=0.05, kmax=None, group=None):
        """P(k) of a real box (this rank's x-slab) on the GPU: forward r2c with the library's own FFT, then
        smk_pk_estimate.  Returns (k_mean, P, n_modes) per bin with P = <|delta_k|^2> V / N^2 and n_modes the number of
        independent complex modes (the mode-count error of a Gaussian field is P / sqrt(n_modes)).  Single rank, or
        all ranks collectively with boxk exchanged by the caller's pipeline (use ChunkPipeline for multi-GPU)."""
        assert self.nranks == 1, "multi-rank: run the forward transform through ChunkPipeline.forward"
        boxk = self.empty_boxk()
        _lib.check(self.lib.smk_fft_r2c(self.h, _ptr(box), C.c_uint64(0), _ptr(boxk)))
        return self.power_spectrum_of_boxk(boxk, nbins, kmin, kmax)

    def power_spectrum_of_boxk(self, boxk, nbins=30, kmin=0.05, kmax=None, group=None):
        kmax = 0.9 * np.pi / self.dcell if kmax is None else kmax
        sums = torch.zeros((3, nbins), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.smk_pk_estimate(self.h, _ptr(boxk), nbins, C.c_double(kmin), C.c_double(kmax), _ptr(sums)))
        if self.nranks > 1:
            torch.distributed.all_reduce(sums, group=group)
        s = sums.cpu().numpy()
        N = float(self.NX) * self.NY * self.NZ
        vol = N * self.dcell ** 3
        with np.errstate(invalid="ignore", divide="ignore"):
            return s[2] / s[1], s[0] / s[1] * vol / N ** 2, s[1] / 2.0

    def sigma(self, stats, ncells=None):
        """np.std(box) from the fused sums (make_boxes.py:92); raises like make_boxes.py:100-105 on a null box."""
        s1, s2 = (float(v) for v in stats.cpu())
        n = float(ncells if ncells is not None else self.NX * self.NY * self.NZ)
        if not (s2 > 0.0) or np.isnan(s2):
            raise ValueError("box is null")
        return float(np.sqrt(max(s2 / n - (s1 / n) ** 2, 0.0)))

    def boxk_to_numpy(self, boxk):
        return boxk[:, :, :self.nzh].cpu().numpy()

    def boxk_from_numpy(self, a):
        boxk = self.empty_boxk()
        boxk[:, :, :self.nzh] = torch.as_tensor(np.ascontiguousarray(a, dtype=np.complex64)).to(self.device)
        return boxk

    # ------------------------------------------------------------------ whole chain through host buffers
    def make_boxes_host(self, W_host, seed=0, noise_host=None, products=PRODUCTS, out_host=None):
        """The C-ABI call a make_boxes.py replacement makes: host weight tables in, host boxes out.
        W_host: dict Pln1,Pln2,Pln3,P0 -> float32 [NX,NY,NZ/2+1] numpy (or pinned torch) arrays.
        Returns ({name: numpy float32 [NX,NY,NZ]}, {name: sigma})."""
        assert self.nranks == 1
        wt = (C.c_void_p * 4)()
        keep = []
        for i, k in enumerate(("Pln1", "Pln2", "Pln3", "P0")):
            w = W_host[k]
            w = w if torch.is_tensor(w) else torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
            keep.append(w)
            wt[i] = w.data_ptr()
        outs = (C.c_void_p * _lib.NPRODUCTS)()
        res = {}
        for name in products:
            if out_host is not None and name in out_host:
                t = out_host[name]
            else:
                t = torch.empty((self.NX, self.NY, self.NZ), dtype=torch.float32)
            res[name] = t
            outs[_lib.PRODUCT_ID[name]] = t.data_ptr()
        sig = (C.c_double * _lib.NPRODUCTS)()
        nz = None
        if noise_host is not None:
            nz = noise_host if torch.is_tensor(noise_host) else torch.from_numpy(
                np.ascontiguousarray(noise_host, dtype=np.float32))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.smk_make_boxes_host(self.h, _ptr(nz), C.c_uint64(seed), wt,
                                                    C.c_double(self.dgrowth0), outs, sig))
        return ({k: v.numpy() for k, v in res.items()}, {n: sig[_lib.PRODUCT_ID[n]] for n in products})
