"""One chunk on one or several GPUs: make_boxes -> (boxes stay in HBM) -> make_spectra -> FGPA.

Multi-GPU layout: every real product is x-slab sharded, rank g owning planes
[g*NX/G, (g+1)*NX/G) (the decomposition make_spectra.py:220-221 uses for its slices); boxk stays in the
transposed y-slab layout, so each 3-D transform needs exactly one all-to-all (torch.distributed / NCCL over
NVLink).  Skewer pixels are computed by the rank whose slab owns them (make_spectra.py:443-448) after a halo
exchange of dmax planes per side.  No data-path collective exists in the skewer stage besides that halo.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from . import slab
from . import spectra as sp
from .boxes import BoxSynth, PRODUCTS, WEIGHT_OF

_ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)   # noqa: E731


class ChunkPipeline(object):
    def __init__(self, NX, NY, NZ, dcell, device, rank=0, nranks=1, dmax=3, zfix=2.4, rsd=True, dla=True,
                 group=None):
        self.bs = BoxSynth(NX, NY, NZ, dcell, device=device, rank=rank, nranks=nranks)
        self.device, self.rank, self.nranks, self.group = device, rank, nranks, group
        self.geom = sp.SkewerGeometry(NX, NY, NZ, dcell, dmax=dmax)
        self.eng = sp.SkewerEngine(self.geom, device=device)
        self.fgpa = sp.FGPA(self.geom, zfix=zfix, device=device)
        self.rsd, self.dla, self.dmax = rsd, dla, dmax
        bs = self.bs
        self.boxk = bs.empty_boxk()
        # halo'd slabs: [halo_lo + nxl + halo_hi, NY, NZ]; global edges get no halo (make_spectra.py:220-221)
        self.hlo, self.hhi = slab.halo(rank, nranks, dmax)
        self.plane = bs.NY * bs.NZ
        self.fields = {}
        for name in PRODUCTS:
            h = (self.hlo + self.hhi) if name in sp.FIELDS else 0
            self.fields[name] = torch.empty((bs.nxl + h, bs.NY, bs.NZ), dtype=torch.float32, device=device)
        self.stats = torch.zeros((len(PRODUCTS), 2), dtype=torch.float64, device=device)
        if nranks > 1:
            # fused peer-store exchange (default) or the pipelined NCCL all-to-all (SMK_P2P=0); measurements of both in
            # profiles/README.md
            self.p2p = os.environ.get("SMK_P2P", "1") != "0"
            n = bs.NX * bs.nyl * bs.pitch
            # NCCL mode: two exchange buffer pairs (the all-to-all of product p overlaps the x pass of p+1 and the y/z
            # passes of p-1).  Fused mode: only the forward transform goes through NCCL -- one send buffer, and the
            # receive side is boxk itself (the forward x pass runs in place), 15 GB less per rank at the nominal size
            npair = 1 if self.p2p else 2
            self.sendbufs = [torch.empty(n, dtype=torch.complex64, device=device) for _ in range(npair)]
            self.recvbufs = ([self.boxk.view(-1)] if self.p2p else
                             [torch.empty(n, dtype=torch.complex64, device=device) for _ in range(npair)])
            self.sendbuf, self.recvbuf = self.sendbufs[0], self.recvbufs[0]
            if self.p2p:
                self._connect_exchange()
        self.W = None
        self.cat = None
        self.footprint = None
        self._qso_setup = None
        # kernels of this library per step: 3 forward passes + 13 x 3 inverse passes + small-scale field + per box class of
        # the catalogue (set_catalogue) the gather's four kernels: sentinels, work list, staged gather, hand-back walk
        self.launches_per_step = 3 + 13 * 3 + 1 + 4

    def _connect_exchange(self):
        """Fused exchange set-up: two receive buffers per rank, mapped into every peer through CUDA IPC."""
        import torch.distributed as dist
        L, h = self.bs.lib, self.bs.h
        _lib.check(L.smk_exchange_create(h, 2))
        for b in range(2):
            mine = (C.c_ubyte * 64)()
            _lib.check(L.smk_exchange_handle(h, b, mine))
            allh = [None] * self.nranks
            dist.all_gather_object(allh, bytes(mine), group=self.group)
            blob = (C.c_ubyte * (64 * self.nranks)).from_buffer_copy(b"".join(allh))
            _lib.check(L.smk_exchange_connect(h, b, blob))
        self._xptr = [L.smk_exchange_ptr(h, b) for b in range(2)]
        self._tok = torch.zeros(1, dtype=torch.float32, device=self.device)
        # x pass on a high-priority stream: its CTAs are placed first; with set_x_sms(n > 0) it runs persistent on n CTAs
        # and the y / z passes of the other stream fill the remaining SMs (SMK_X_SMS = initial value, default 0 = off)
        self._xstream = torch.cuda.Stream(device=self.device, priority=-1)
        # measured on 8 x B200 at the nominal size (bench.py --x-sms-sweep, profiles/README.md): boxes 184 ms with one CTA
        # per tile, 318 / 226 / 179 / 167 ms persistent on 32 / 48 / 64 / 96 CTAs -- the x pass is NVLink-bound there;
        # at 2 ranks it is HBM-bound and capping it loses
        default = 96 if self.nranks >= 8 else 0
        self.set_x_sms(int(os.environ.get("SMK_X_SMS", str(default)) or 0))
        dist.barrier(group=self.group)

    def set_x_sms(self, n):
        """CTAs of the persistent fused-exchange x pass (smk_exchange_set_sms); 0 = one CTA per tile."""
        _lib.check(self.bs.lib.smk_exchange_set_sms(self.bs.h, int(n)))
        self.x_sms = int(n)

    def _stream_barrier(self):
        """Cross-rank barrier in stream order (a one-element all-reduce): when it completes on this rank's stream,
        every rank's kernels enqueued before its own barrier have finished, so their peer stores are visible."""
        torch.distributed.all_reduce(self._tok, group=self.group)

    # ------------------------------------------------------------------ inputs
    def set_weights(self, W):
        """W: dict Pln1,Pln2,Pln3,P0 -> device float32 [NX, NY/R, NZ/2+1] (this rank's ky rows)."""
        self.W = {k: self.bs.upload_weights(v) for k, v in W.items()}

    def set_catalogue(self, ra, dec, z, ra0, dec0, ids=None):
        g = self.geom
        xyzr, nfor = sp.qso_lines_of_sight(g, ra, dec, z, ra0, dec0)
        keep = nfor >= 0
        xmin, xmax = slab.x_bounds(self.rank, self.nranks, g.LX)
        touch = slab.touching(xyzr, g.R_vec[0], g.R_vec[-1], xmin, xmax)
        sel = np.where(keep & touch)[0]
        # The staged gather sizes its shared-memory box for the most oblique sightline of a call (include/smk.h:
        # dir_x_max / dir_y_max).  Sightlines are therefore grouped by the box they need (extent along x and y of a
        # 128-pixel segment, in steps of two cells) and the gather is called once per group: the near-axis sightlines
        # of a wide chunk keep the small box and the high occupancy that comes with it.
        lseg = 127 * g.pixel
        ex = np.floor(lseg * np.abs(xyzr[sel, 0] / xyzr[sel, 3]) / g.DX).astype(np.int64) // 2
        ey = np.floor(lseg * np.abs(xyzr[sel, 1] / xyzr[sel, 3]) / g.DY).astype(np.int64) // 2
        key = ex * 64 + ey
        sel = sel[np.argsort(key, kind="stable")]
        key = np.sort(key, kind="stable")
        starts = np.concatenate(([0], np.where(np.diff(key) != 0)[0] + 1, [len(sel)])) if len(sel) else np.array([0, 0])
        groups = [(int(a), int(b)) for a, b in zip(starts[:-1], starts[1:]) if b > a]
        ids = np.arange(len(nfor), dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        self.cat = dict(sel=sel, xyzr=np.ascontiguousarray(xyzr[sel]), nfor=np.ascontiguousarray(nfor[sel]),
                        ids=ids[sel], z=np.asarray(z)[sel], xmin=xmin, xmax=xmax, n_total=len(nfor),
                        kept=np.where(keep)[0], xyzr_kept=np.ascontiguousarray(xyzr[keep]), groups=groups)
        nq, npix = len(sel), g.npixeltot
        self.cat["xyzr_d"] = torch.as_tensor(self.cat["xyzr"], device=self.device)
        self.cat["nfor_d"] = torch.as_tensor(self.cat["nfor"], device=self.device)
        self.cat["nf_merge"] = self.fgpa.forest_count(self.cat["z"])
        self.cat["prep"] = self.fgpa.prepare(self.cat["nf_merge"], self.cat["ids"])
        # rows are written only at the pixels this slab owns (smk_skewers: xmin < X <= xmax); everything else keeps this
        # NaN, which is what merging the slabs' pieces selects on (gather_rows; bin/make_spectra.py relies on the same mask)
        self.out = tuple(torch.full((nq, npix), float("nan"), dtype=torch.float32, device=self.device) for _ in range(4))
        self.delta_s = torch.zeros((nq, npix), dtype=torch.float32, device=self.device)
        # pixels this rank owns: xmin < X <= xmax and inside the forest (for the pixel-rate metric).  X = x_q r / R_q is
        # monotonic along the sightline, so the owned pixels of a sightline are one interval of the pixel grid
        cx, cR, cn = self.cat["xyzr"][:, 0], self.cat["xyzr"][:, 3], self.cat["nfor"].astype(np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            r_min, r_max = xmin * cR / cx, xmax * cR / cx
        lo = np.where(cx > 0, np.searchsorted(g.R_vec, r_min, side="right"), np.searchsorted(g.R_vec, r_max, side="left"))
        hi = np.where(cx > 0, np.searchsorted(g.R_vec, r_max, side="right"), np.searchsorted(g.R_vec, r_min, side="left"))
        zero = cx == 0
        lo = np.where(zero, 0, lo)
        hi = np.where(zero, npix if xmin < 0 <= xmax else 0, hi)
        own = int(np.clip(np.minimum(hi, cn) - lo, 0, None).sum())
        self.cat["own_pixels"] = own
        self.launches_per_step = 3 + 13 * 3 + 1 + 4 * len(groups) + (0 if os.environ.get("SMK_FUSED_FGPA", "1") != "0" else 1)

    def forest_pixels_total(self):
        t = torch.tensor([self.cat["own_pixels"]], dtype=torch.int64, device=self.device)
        if self.nranks > 1:
            torch.distributed.all_reduce(t, group=self.group)
        return int(t.item())

    # ------------------------------------------------------------------ boxes
    def _a2a(self):
        torch.distributed.all_to_all_single(self.recvbuf, self.sendbuf, group=self.group)

    def forward(self, seed, noise=None):
        bs, L = self.bs, self.bs.lib
        if self.nranks == 1:
            _lib.check(L.smk_fft_r2c(bs.h, _ptr(noise), C.c_uint64(seed), _ptr(self.boxk)))
        else:
            _lib.check(L.smk_fft_r2c_local(bs.h, _ptr(noise), C.c_uint64(seed), _ptr(self.sendbuf)))
            self._a2a()
            _lib.check(L.smk_fft_r2c_finish(bs.h, _ptr(self.recvbuf), _ptr(self.boxk)))

    def interior(self, name):
        """The owned planes of a halo'd field buffer (contiguous view)."""
        f = self.fields[name]
        lo = self.hlo if name in sp.FIELDS else 0
        return f[lo:lo + self.bs.nxl]

    def product(self, name):
        bs, L = self.bs, self.bs.lib
        pid = _lib.PRODUCT_ID[name]
        wt = self.W[WEIGHT_OF[name]] if name in WEIGHT_OF else None
        out = self.interior(name)
        st = self.stats[pid]
        if self.nranks == 1:
            _lib.check(L.smk_synth_c2r(bs.h, _ptr(self.boxk), pid, _ptr(wt), 1, C.c_double(bs.dgrowth0), _ptr(out),
                                       _ptr(st)))
        else:
            _lib.check(L.smk_synth_c2r_local(bs.h, _ptr(self.boxk), pid, _ptr(wt), 1, C.c_double(bs.dgrowth0),
                                             _ptr(self.sendbuf)))
            self._a2a()
            _lib.check(L.smk_synth_c2r_finish(bs.h, _ptr(self.recvbuf), _ptr(out), _ptr(st)))

    def step_boxes(self, seed=0, noise=None, products=PRODUCTS, on_product=None):
        """Forward transform + the inverse transform of every product.  on_product(name), if given, is called right
        after the last kernel of that product has been enqueued on the current stream (the end-to-end arm hangs its
        device-to-host copy of the finished box there)."""
        self.stats.zero_()
        self.forward(seed, noise)
        if self.nranks == 1:
            for name in products:
                self.product(name)
                if on_product is not None:
                    on_product(name)
            return
        # software pipeline over the independent inverse transforms (boxk is read-only after the 'box' product has
        # stored boxk*P0 back, and the products before it only read it): x pass of p | all-to-all of p-1 | y,z of p-2
        bs, L = self.bs, self.bs.lib
        if self.p2p:
            # fused exchange, two streams: X runs the x pass of product q (NVLink-bound peer stores) while M runs the
            # y and z passes of product q-1 (HBM-bound).  Per product, on X:  x(q) -> wait yz(q-2) -> barrier(q);
            # on M: wait barrier(q) -> y,z(q).  barrier(q) completes only when every rank has finished x(q) (its stores
            # into our buffer q&1 are visible) and yz(q-2) (buffer q&1 may be overwritten by x(q+2)... see below).
            main = torch.cuda.current_stream(self.device)
            X = self._xstream
            X.wait_stream(main)                         # boxk (forward transform) is ready
            ev_yz = [None, None]
            try:
                for q, name in enumerate(products):
                    slot = q & 1
                    pid = _lib.PRODUCT_ID[name]
                    wt = self.W[WEIGHT_OF[name]] if name in WEIGHT_OF else None
                    with torch.cuda.stream(X):
                        L.smk_set_stream(bs.h, C.c_void_p(X.cuda_stream))
                        # x(q) overwrites buffer `slot` on the peers: they finished reading it (yz(q-2)) before they
                        # entered barrier(q-1), which precedes x(q) on this stream
                        _lib.check(L.smk_synth_c2r_local_p2p(bs.h, _ptr(self.boxk), pid, _ptr(wt), 1,
                                                             C.c_double(bs.dgrowth0), slot))
                        if ev_yz[slot ^ 1] is not None:
                            X.wait_event(ev_yz[slot ^ 1])   # yz(q-1) done before this rank enters barrier(q)
                        self._stream_barrier()
                        ev_bar = torch.cuda.Event()
                        ev_bar.record(X)
                    L.smk_set_stream(bs.h, C.c_void_p(main.cuda_stream))
                    main.wait_event(ev_bar)
                    _lib.check(L.smk_synth_c2r_finish_p2p(bs.h, slot, _ptr(self.interior(name)), _ptr(self.stats[pid])))
                    ev_yz[slot] = torch.cuda.Event()
                    ev_yz[slot].record(main)
                    if on_product is not None:
                        on_product(name)
            finally:                                   # whatever happened, the ctx goes back to the caller's stream
                L.smk_set_stream(bs.h, C.c_void_p(main.cuda_stream))
                main.wait_stream(X)
            return
        pending = None            # (work, slot, name)
        for i, name in enumerate(products):
            slot = i & 1
            pid = _lib.PRODUCT_ID[name]
            wt = self.W[WEIGHT_OF[name]] if name in WEIGHT_OF else None
            _lib.check(L.smk_synth_c2r_local(bs.h, _ptr(self.boxk), pid, _ptr(wt), 1, C.c_double(bs.dgrowth0),
                                             _ptr(self.sendbufs[slot])))
            work = torch.distributed.all_to_all_single(self.recvbufs[slot], self.sendbufs[slot], group=self.group,
                                                       async_op=True)
            if pending is not None:
                self._finish(*pending)
                if on_product is not None:
                    on_product(pending[2])
            pending = (work, slot, name)
        self._finish(*pending)
        if on_product is not None:
            on_product(pending[2])

    def _finish(self, work, slot, name):
        work.wait()               # the compute stream waits for the exchange; the host does not block
        pid = _lib.PRODUCT_ID[name]
        _lib.check(self.bs.lib.smk_synth_c2r_finish(self.bs.h, _ptr(self.recvbufs[slot]), _ptr(self.interior(name)),
                                                    _ptr(self.stats[pid])))

    def sigmas(self):
        """np.std of every product over the whole box (all ranks)."""
        s = self.stats.clone()
        if self.nranks > 1:
            torch.distributed.all_reduce(s, group=self.group)
        s = s.cpu().numpy()
        n = float(self.bs.NX) * self.bs.NY * self.bs.NZ
        return {name: float(np.sqrt(max(s[i, 1] / n - (s[i, 0] / n) ** 2, 0.0))) for i, name in enumerate(PRODUCTS)}

    # ------------------------------------------------------------------ skewers
    def halo_exchange(self):
        """dmax planes per side of each of the 10 skewer fields to/from the x-neighbours."""
        if self.nranks == 1:
            return
        import torch.distributed as dist
        ops = []
        d, nxl = self.dmax, self.bs.nxl
        for name in sp.FIELDS:
            f = self.fields[name]
            if self.rank > 0:                       # lower neighbour
                ops.append(dist.P2POp(dist.isend, f[self.hlo:self.hlo + d], self.rank - 1, group=self.group))
                ops.append(dist.P2POp(dist.irecv, f[0:d], self.rank - 1, group=self.group))
            if self.rank < self.nranks - 1:         # upper neighbour
                ops.append(dist.P2POp(dist.isend, f[self.hlo + nxl - d:self.hlo + nxl], self.rank + 1, group=self.group))
                ops.append(dist.P2POp(dist.irecv, f[self.hlo + nxl:self.hlo + nxl + d], self.rank + 1, group=self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()

    def step_skewers(self, seed=0, noise=None):
        c = self.cat
        self.halo_exchange()
        nq, npix = len(c["sel"]), self.geom.npixeltot
        if nq == 0:
            return self.out
        dl, ep, vp, F = self.out
        fl = (C.c_void_p * 10)()
        for i, k in enumerate(sp.FIELDS):
            fl[i] = self.fields[k].data_ptr()
        ix0 = self.rank * self.bs.nxl - self.hlo
        nxs = self.bs.nxl + self.hlo + self.hhi
        L = self.bs.lib
        # small-scale field first, then the gather with the FGPA in its epilogue (delta_l / eta_par are not read back)
        ds = self.fgpa.small_scales(c["nf_merge"], noise=noise, seed=seed, prepared=c["prep"])
        timed = getattr(self, "gather_events", None)        # bench.py: CUDA events around the gather kernels
        if timed is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        fused = os.environ.get("SMK_FUSED_FGPA", "1") != "0"
        self.gather_stats = []
        for a, b in c["groups"]:                            # one call per box class (set_catalogue), rows [a, b)
            cg = self.geom.c_geom(c["xyzr"][a:b])
            row = lambda t: C.c_void_p(t.data_ptr() + a * t.stride(0) * t.element_size())
            args = (self.bs.h, C.byref(cg), fl, ix0, nxs, C.c_double(c["xmin"]), C.c_double(c["xmax"]), int(self.rsd),
                    int(self.dla), b - a, row(c["xyzr_d"]), row(c["nfor_d"]), _ptr(self.eng.rvec), npix, row(dl), row(ep),
                    row(vp))
            if fused:
                _lib.check(L.smk_skewers_fgpa(*args, row(ds), _ptr(self.fgpa.G), _ptr(self.fgpa.a), _ptr(self.fgpa.b),
                                              _ptr(self.fgpa.c), row(F)))
            else:
                _lib.check(L.smk_skewers(*args))
        if not fused:
            _lib.check(L.smk_fgpa(self.bs.h, nq, npix, _ptr(dl), _ptr(ds), _ptr(ep), _ptr(self.fgpa.G),
                                  _ptr(self.fgpa.a), _ptr(self.fgpa.b), _ptr(self.fgpa.c), _ptr(F)))
        if timed is not None:
            e1.record()
            timed.append((e0, e1))
        self.delta_s = ds
        return self.out

    def step(self, seed=0):
        self.step_boxes(seed)
        self.step_skewers(seed)

    def gather_rows(self):
        """Complete spectra rows on the home rank of every quasar (the slab that contains the quasar itself: the
        reference's HDU / slice id, draw_qso.py:495).  The reference stitches a sightline's pieces through the file
        system (make_spectra.py writes one piece per slice, merge_spectra.py:385-406 concatenates and sorts them); here
        each of delta_l, eta_par, vpar, flux, delta_s makes one all-to-all and the pieces are merged by position.
        Returns {"index": catalogue indices of this rank's quasars, "delta_l", "eta_par", "vpar", "flux", "delta_s":
        device tensors [n_home, npix]}; pixels outside the box keep NaN (they are in no slice of the reference either)."""
        c, g = self.cat, self.geom
        names = ("delta_l", "eta_par", "vpar", "flux")
        if self.nranks == 1:
            res = {"index": c["sel"]}
            res.update({k: t for k, t in zip(names, self.out)})
            res["delta_s"] = self.delta_s
            return res
        plan = slab.row_gather_plan(c["xyzr_kept"], g.R_vec[0], g.R_vec[-1], self.nranks, g.LX, self.rank)
        # the plan counts local rows in ascending catalogue order; the local rows are grouped by box class (set_catalogue)
        plan["send_rows"] = np.argsort(c["sel"], kind="stable")[plan["send_rows"]]
        res = {"index": c["kept"][plan["home_qso"]]}
        for k, t in zip(names + ("delta_s",), self.out + (self.delta_s,)):
            res[k] = slab.exchange_rows(t, plan, group=self.group)
        return res

    # ------------------------------------------------------------------ quasars (draw_qso.py on the resident boxes)
    def set_footprint(self, ra0, dec0, dra, ddec, zmin=1.8, zmax=3.6, chunk=1):
        """Window of the chunk (submit_mocks.py chunk_parameters) used by draw_qso."""
        self.footprint = dict(ra0=ra0, dec0=dec0, dra=dra, ddec=ddec, zmin=zmin, zmax=zmax, chunk=chunk)
        self._qso_setup = None

    def draw_qso(self, seed=0, uniforms=None):
        """Quasars of this rank's slab from the resident lognormal and velocity boxes (smk_draw_qso; the slab plays
        the role of the reference's slice `-i rank -Nslice nranks`).  sigma of the lognormal boxes comes from the sums
        the c2r pass accumulated (no extra pass).  Philox draws unless `uniforms` (legacy NumPy stream) is given."""
        from . import qso
        fp, bs = self.footprint, self.bs
        n = float(bs.nxl) * bs.NY * bs.NZ
        st3 = self.stats[[_lib.PRODUCT_ID[k] for k in ("boxln_1", "boxln_2", "boxln_3")]].cpu().numpy()
        sigma = tuple(float(np.float32(np.sqrt(max(s2 / n - (s1 / n) ** 2, 0.0)))) for s1, s2 in st3)
        if self._qso_setup is None or self._qso_setup.sigma_p != sigma:
            self._qso_setup = qso.QsoSetup(bs.nxl, bs.NY, bs.NZ, bs.NX, bs.dcell, self.rank, self.nranks, fp["ra0"],
                                           fp["dec0"], fp["dra"], fp["ddec"], fp["zmin"], fp["zmax"], sigma,
                                           dmax=self.dmax, rho_sum=qso.scaled_rho_sum(n))
        if getattr(self, "_qso_drawer", None) is None:
            self._qso_drawer = qso.QsoDrawer(bs)
        return self._qso_drawer.draw(self._qso_setup, [self.interior("boxln_%d" % k) for k in (1, 2, 3)],
                                     [self.interior(k) for k in ("vx", "vy", "vz")] if self.rsd else None,
                                     ix0=self.rank * bs.nxl, uniforms=uniforms, seed=seed, chunk=fp["chunk"], pmf=False)

    # ------------------------------------------------------------------ one chunk of the footprint, device resident
    def run_chunk(self, chunk=1, seed=0, cells=None, stripe_footprint=False):
        """The reference's per-chunk chain run_boxes-<c>.sh -> run_chunk-<c>.sh (submit_mocks.py:375-425: make_boxes,
        draw_qso, make_spectra, merge_spectra) without the files in between: boxes of this chunk (seed as given to
        make_boxes.py), quasars drawn on the resident lognormal boxes inside the chunk's window of chunk_parameters()
        (Philox draws, THING_ID = chunk*1e9 + slab*1e6 + n + 1 as draw_qso.py:495), their sightlines through the
        resident delta / eta / velocity boxes, small-scale field and FGPA.  Returns (catalogue of every rank's quasars
        as a dict of numpy columns, (delta_l, eta_par, vpar, flux) device rows of the quasars whose sightline touches
        this rank's slab -- row i belongs to catalogue entry self.cat["sel"][i])."""
        from . import chunks
        import time
        ra0, dra, dec0, ddec = chunks.chunk_window(cells if cells is not None else self.bs.NX, chunk, stripe_footprint)
        self.set_footprint(ra0, dec0, dra, ddec, chunk=int(chunk))
        tm, t0 = {}, time.time()

        def lap(name):                      # wall time of each phase (device work included: synchronised)
            nonlocal t0
            torch.cuda.synchronize(self.device)
            tm[name] = time.time() - t0
            t0 = time.time()
        self.step_boxes(seed)
        lap("boxes")
        mine = self.draw_qso(seed)
        lap("draw_qso")
        cols = ("RA", "DEC", "Z_QSO_NO_RSD", "Z_QSO_RSD", "HDU", "THING_ID")
        parts = [{c: mine[c] for c in cols}]
        if self.nranks > 1:                      # every rank needs the quasars of all slabs: a sightline crosses slabs
            parts = [None] * self.nranks
            torch.distributed.all_gather_object(parts, {c: mine[c] for c in cols}, group=self.group)
        cat = {c: np.concatenate([p[c] for p in parts]) for c in cols}
        lap("catalogue_allgather")
        z = cat["Z_QSO_RSD"] if self.rsd else cat["Z_QSO_NO_RSD"]              # make_spectra.py:417-420
        self.set_catalogue(cat["RA"], cat["DEC"], z, ra0, dec0, ids=cat["THING_ID"])
        lap("sightline_setup_host")
        self.step_skewers(seed)
        lap("skewers")
        self.last_chunk_timings = tm
        return cat, self.out

    # ------------------------------------------------------------------ end-to-end through host buffers
    # (any number of ranks: every rank stages its own x-slab of each box and its own spectra rows over its own PCIe link)
    STAGE_BYTES = 1 << 30      # pinned staging buffers of the end-to-end arm (two per rank)

    def make_host_buffers(self, W_dev):
        bs = self.bs
        from . import pk
        pp = {}
        for k in W_dev:                     # P(k) splines in piecewise-polynomial form: the real inputs of the chunk
            br, co = pk.ppoly(k)
            pp[k] = (torch.from_numpy(br).pin_memory(), torch.from_numpy(co).pin_memory(),
                     torch.empty(br.shape, dtype=torch.float64, device=self.device),
                     torch.empty(co.shape, dtype=torch.float64, device=self.device))
        slab = bs.nxl * bs.NY * bs.NZ
        # the two pinned buffers stand for the FITS writer's staging area: a slab larger than a buffer streams through
        # them piece by piece (every byte of every box still crosses PCIe inside the timed region)
        nstage = min(slab, self.STAGE_BYTES // 4)
        host = {"pp": pp,
                "box": [torch.empty(nstage, dtype=torch.float32).pin_memory() for _ in range(2)],
                "spec": [torch.empty(self.out[0].shape, dtype=torch.float32).pin_memory() for _ in range(4)],
                "copy_stream": torch.cuda.Stream(device=self.device), "npiece": 0}
        host["h2d_bytes"] = (sum(v[0].numel() * 8 + v[1].numel() * 8 for v in pp.values()) + self.cat["xyzr"].nbytes
                             + self.cat["nfor"].nbytes)
        host["d2h_bytes"] = len(PRODUCTS) * slab * 4 + 4 * self.out[0].numel() * 4
        return host

    def step_e2e_resident(self, host, seed=0):
        """The chunk as one pipeline: host inputs in (P(k) splines, sightline catalogue), boxes stay in HBM, quasars are
        drawn on the resident boxes (smk_draw_qso, Philox draws) and only the quasar table and the spectra rows go back
        to the host -- no box ever crosses PCIe.  (The skewers use the catalogue given to set_catalogue so that the
        workload is the one `value` is quoted on.)"""
        assert self.footprint is not None
        main = torch.cuda.current_stream(self.device)
        cs = host["copy_stream"]
        self.stats.zero_()
        for k, (br_h, co_h, br_d, co_d) in host["pp"].items():
            br_d.copy_(br_h, non_blocking=True)
            co_d.copy_(co_h, non_blocking=True)
            _lib.check(self.bs.lib.smk_pk_weights(self.bs.h, _ptr(br_d), _ptr(co_d), int(co_h.shape[1]),
                                                  _ptr(self.W[k])))
        c = self.cat
        c["xyzr_d"].copy_(torch.from_numpy(c["xyzr"]), non_blocking=True)
        c["nfor_d"].copy_(torch.from_numpy(c["nfor"]), non_blocking=True)
        self.step_boxes(seed)
        cat = self.draw_qso(seed)                 # synchronises (the table comes back to the host)
        self.step_skewers(seed)
        e = torch.cuda.Event()
        e.record(main)
        cs.wait_event(e)
        with torch.cuda.stream(cs):
            for h, d in zip(host["spec"], self.out):
                h.copy_(d, non_blocking=True)
        cs.synchronize()
        main.synchronize()
        return cat

    def step_e2e(self, host, seed=0):
        """Host inputs in (P(k) splines and the quasar catalogue, pinned), every box and every spectrum row out to
        host memory.  The spectral weight tables are evaluated on the GPU (smk_pk_weights) inside the step.  The boxes
        go through the same pipeline as the device-resident step (step_boxes: fused exchange on several ranks); the
        device-to-host copy of a finished box is enqueued on a copy stream and overlaps the next product's transforms.
        The step is bound by PCIe: every rank moves its own x-slabs and rows over its own link."""
        main = torch.cuda.current_stream(self.device)
        cs = host["copy_stream"]
        self.stats.zero_()
        for k, (br_h, co_h, br_d, co_d) in host["pp"].items():
            br_d.copy_(br_h, non_blocking=True)
            co_d.copy_(co_h, non_blocking=True)
            _lib.check(self.bs.lib.smk_pk_weights(self.bs.h, _ptr(br_d), _ptr(co_d), int(co_h.shape[1]),
                                                  _ptr(self.W[k])))
        c = self.cat
        c["xyzr_d"].copy_(torch.from_numpy(c["xyzr"]), non_blocking=True)
        c["nfor_d"].copy_(torch.from_numpy(c["nfor"]), non_blocking=True)

        def copy_out(name):
            e = torch.cuda.Event()
            e.record(main)
            cs.wait_event(e)
            flat = self.interior(name).reshape(-1)
            step = host["box"][0].numel()
            with torch.cuda.stream(cs):
                for o in range(0, flat.numel(), step):
                    n = min(step, flat.numel() - o)
                    host["box"][host["npiece"] & 1][:n].copy_(flat[o:o + n], non_blocking=True)
                    host["npiece"] += 1

        self.step_boxes(seed, on_product=copy_out)
        self.step_skewers(seed)
        e = torch.cuda.Event()
        e.record(main)
        cs.wait_event(e)
        with torch.cuda.stream(cs):
            for h, d in zip(host["spec"], self.out):
                h.copy_(d, non_blocking=True)
        cs.synchronize()
        main.synchronize()
        return host
