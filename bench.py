#!/usr/bin/env python
"""Benchmark of the hot path: one "step" = one chunk = Philox noise + r2c + 13 x (weight/k-factor + c2r + sigma)
+ full-density skewers (gather + small-scale field + FGPA) on synthetic input of the BASELINE.json shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--box NX] [--impl reference]

Prints ONE JSON line (see the driver contract in the task statement).  `value` = chunk cells / device time of the
whole step with inputs resident in HBM; `e2e` = same through host buffers (pinned inputs in, every box and
every spectrum row copied back); `roofline` = the kernel with the largest share of the step (FFT passes against the
measured HBM copy peak, the skewer gather against the FP32 peak and the HBM byte model), every hot kernel listed under
`roofline.kernels`; `cpu_baseline` = the CPU oracle (restated reference, scipy.fft + numba) on a bounded sample on this
host; `parity_selfcheck` (N > 1) = sharded vs single-GPU pipeline on the same seeded thin box, run before timing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def chunk_window(nx):
    """(ra0, dec0, half-width deg) of chunk 1 for a box of nx cells: chunk_parameters() of bin/submit_mocks.py:611-674
    (saclaymocks_b200/chunks.py); sizes outside its table (weak-scaling boxes) scale the 512-cell window."""
    from saclaymocks_b200 import chunks
    try:
        ra0, dra, dec0, _ = chunks.chunk_window(nx, 1)
        return ra0, dec0, dra
    except ValueError:
        return 190.0, 0.0, 6.4 * nx / 512.0


QSO_DENSITY = 89.8          # per deg^2 for 1.8 < z < 3.6 from etc/nz_qso_desi.dat
DCELL = 2.19
NZ = 1536


# ------------------------------------------------------------------------------------------------ synthetic input
def synthetic_qsos(nx, ny, seed=42):
    """Uniform in the chunk window, z from etc/nz_qso_desi.dat restricted to 1.8 < z < 3.6 (full density)."""
    from saclaymocks_b200 import tables
    ra0, dec0, half = chunk_window(nx)
    half_y = half * ny / nx
    rng = np.random.default_rng(seed)
    n = int(QSO_DENSITY * (2 * half) * (2 * half_y))
    nz = tables.nz_qso_desi()
    zc, w = 0.5 * (nz[:, 0] + nz[:, 1]), nz[:, 2].copy()
    w[(zc < 1.8) | (zc > 3.6)] = 0
    b = rng.choice(len(zc), size=n, p=w / w.sum())
    z = np.clip(nz[b, 0] + rng.uniform(0, 1, n) * (nz[b, 1] - nz[b, 0]), 1.8001, 3.5999).astype(np.float32)
    ra = (ra0 + rng.uniform(-half, half, n)).astype(np.float32)
    dec = (dec0 + rng.uniform(-half_y, half_y, n)).astype(np.float32)
    return ra, dec, z, ra0, dec0


def weight_tables_device(bs, device):
    """The four spectral weight tables of this rank's k-slab, evaluated on the GPU from the P(k) splines
    (smk_pk_weights = GPU interpolate_pk; input preparation, outside the timed region)."""
    return {name: bs.weight_table(name) for name in ("Pln1", "Pln2", "Pln3", "P0")}


# The configuration BASELINE.json names for each GPU count: config 2 (one 512 x 512 x 1536 chunk) on one GPU, config 3
# (box-size 1024) on 2 and 4 GPUs, config 4 (the nominal 2560 x 2560 x 1536 chunk) on 8 GPUs.  `--weak` selects the
# round-1 boxes with a fixed number of cells per GPU instead.
CONFIG_BOX = {1: (512, 512), 2: (1024, 1024), 4: (1024, 1024), 8: (2560, 2560)}
WEAK_BOX = {1: (512, 512), 2: (1024, 512), 4: (1024, 1024), 8: (2048, 1024)}
FP32_FLOP_PER_PIXEL = 343 * 11 * 2      # literal ReadSpec: 343 weights x (10 fields + the weight sum), SURVEY.md 8(d)
SKEWER_BYTES_PER_PIXEL = 191.0          # 17.9 B/pixel/field x 10 fields + 12 B of outputs, SURVEY.md 8(d)


def box_for(world, args):
    if args.box:
        return args.box, args.box
    return (WEAK_BOX if args.weak else CONFIG_BOX)[world]


def workload_config(nx, ny, world):
    """The `config` object of the JSON line: identical in the GPU arm and the reference arm."""
    nqso = len(synthetic_qsos(nx, ny)[0])
    return {"workload": workload_name(nx, ny, nqso), "box": [nx, ny, NZ], "nqso": int(nqso),
            "cells_per_gpu": nx * ny * NZ // world,
            "l2": "every pass streams >= 1.6 GB per launch, far above the 126 MB L2; no flush needed",
            "parallelism": "x-slabs over %d GPU(s), one exchange per 3-D transform" % world}


def workload_name(nx, ny, nqso):
    return ("single chunk %dx%dx%d: Philox noise + r2c + 13 products (3 lognormal, delta, 6 eta, 3 velocity) + %d "
            "full-density skewers (gather 10 fields, delta_s, FGPA)" % (nx, ny, NZ, nqso))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock and throttle reasons sampled through NVML every 20 ms while the timed region runs."""

    def __init__(self, gpu_index=0):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.idx = gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                self._stop.wait(0.02)
        except Exception as e:          # NVML missing: fall back to one nvidia-smi query
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=clocks.sm,clocks.max.sm",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                v = o.stdout.strip().split(",")
                self.samples.append(float(v[0]))
                self.max_mhz = float(v[1])
            except Exception:
                self.reasons.add("clock query failed: %s" % e)

    def __enter__(self):
        self.t.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ CPU oracle timing
def _skewer_worker(args):
    (NX, NY, NZ_, dcell, seed, nq) = args
    from oracle import spectra as osp
    rng = np.random.default_rng(seed)
    boxes = {k: rng.standard_normal((NX, NY, NZ_), dtype=np.float32) for k in osp.FIELDS}
    geom = osp.Geometry(NX, NY, NZ_, dcell)
    ra, dec, z, ra0, dec0 = synthetic_qsos(NX, NY, seed)
    q = np.zeros(nq, dtype=[("RA", "f4"), ("DEC", "f4"), ("Z_QSO_NO_RSD", "f4"), ("Z_QSO_RSD", "f4"),
                            ("THING_ID", "i8"), ("HDU", "i4")])
    sel = rng.choice(len(ra), nq, replace=False)
    q["RA"], q["DEC"], q["Z_QSO_RSD"], q["Z_QSO_NO_RSD"] = ra[sel], dec[sel], z[sel], z[sel]
    osp.make_spectra_slice(geom, boxes, [q[:1]], 0, 1, ra0, dec0)          # numba warm-up, not timed
    t0 = time.time()
    pieces = osp.make_spectra_slice(geom, boxes, [q], 0, 1, ra0, dec0)
    dt = time.time() - t0
    return sum(p["npix_forest"] for p in pieces), dt


def cpu_oracle_step(nx, workers, nq_per_worker=24, skewer_procs=None):
    """One bounded sample of the chunk on the CPU oracle: a nx x nx x 1536 box through all 13 products plus skewers
    at full density, the skewer time extrapolated from nq_per_worker quasars per core running on all cores."""
    from multiprocessing import get_context
    from oracle import boxes as ob
    dcell = DCELL
    t0 = time.time()
    rng = np.random.default_rng(1)
    W = {k: rng.random((nx, nx, NZ // 2 + 1), dtype=np.float32) for k in ("Pln1", "Pln2", "Pln3", "P0")}   # synthetic
    t1 = time.time()
    ob.make_boxes(nx, nx, NZ, dcell, 42, W, workers=workers)
    t_boxes = time.time() - t1
    procs = skewer_procs or workers
    with get_context("fork").Pool(procs) as pool:
        res = pool.map(_skewer_worker, [(32, 32, NZ, dcell, 100 + i, nq_per_worker) for i in range(procs)])
    npx = sum(r[0] for r in res)
    rate = sum(r[0] / r[1] for r in res)                 # aggregate pixels/s with `procs` processes
    ra, dec, z, _, _ = synthetic_qsos(nx, nx)
    from oracle import spectra as osp
    geom = osp.Geometry(nx, nx, NZ, dcell)
    npx_full = int(np.searchsorted(geom.lambda_vec, 1215.67 * (1 + z.astype(np.float64))).sum())
    t_skew = npx_full / rate if rate > 0 else 0.0
    cells = nx * nx * NZ
    return {"cells": cells, "t_boxes": t_boxes, "t_skewers_extrapolated": t_skew, "skewer_pixels_per_s": rate,
            "npix_sampled": npx, "npix_full": npx_full, "setup_s": t1 - t0,
            "value": cells / (t_boxes + t_skew)}


REF_SAMPLE_NX = 256          # the CPU sample: a 256 x 256 x 1536 box (the size that is kinder to the CPU than 128)
REF_TIME_BUDGET_S = 200.0    # the reference arm stops timing new steps once this much wall time has been spent


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (restated as the oracle: the reference is
    Python and its FFTW/fitsio/healpy wheels are absent) on all host cores.  Every step is the SAME bounded
    sample -- a 256 x 256 x 1536 box through all 13 products plus full-density skewers -- whatever --steps says; the
    number of timed steps is capped by a wall-time budget instead of shrinking the sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    nx = REF_SAMPLE_NX
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    bx, by = box_for(world, args)
    t_start = time.time()
    cpu_oracle_step(32, cores, nq_per_worker=4, skewer_procs=min(cores, 4))        # imports, numba compilation
    nwarm = min(args.warmup, 1)
    vals, times, steps_timed = [], [], 0
    for i in range(nwarm + args.steps):
        t0 = time.time()
        r = cpu_oracle_step(nx, cores)
        if i >= nwarm:
            vals.append(r["value"])
            times.append(time.time() - t0)
            steps_timed += 1
            if time.time() - t_start + times[-1] > REF_TIME_BUDGET_S:
                break
    v = float(np.mean(vals))
    sample = ("each step = a %dx%dx1536 box of the same workload: 13 products (scipy.fft float32, workers=%d) + numba "
              "ReadSpec on 24 quasars/core extrapolated to that box's full-density catalogue; cells/s of the sample "
              "stand for the workload's (FFT cost per cell grows only logarithmically with the box); pocketfft stands "
              "in for FFTW (pyfftw is not installable); %d of the %d requested steps timed (wall-time budget %.0f s, "
              "%d warm-up)" % (nx, nx, cores, steps_timed, args.steps, REF_TIME_BUDGET_S, nwarm))
    line = {"impl": "reference", "metric": "grf_cells_per_s", "value": v, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "steps_timed": steps_timed,
            "ms_per_step": 1e3 * float(np.mean(times)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(bx, by, world),
            "cpu_baseline": {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                             "sample_box": [nx, nx, NZ]},
            "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def quiet_stdout():
    """Reserve stdout for the ONE JSON line: whatever a library writes to file descriptor 1 meanwhile (NCCL prints its
    version banner there when NCCL_DEBUG is set) is sent to stderr.  Returns the descriptor of the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def emit(real_fd, text):
    sys.stdout.flush()
    os.write(real_fd, (text + "\n").encode())


def parity_selfcheck(world, rank, dev, nx):
    """Sharded pipeline against the single-GPU pipeline on the same seeded input, run by every rank before the timed
    region (N > 1): a thin box with the benched x and z transform lengths and the real geometry (nx x 64 x 1536 cells
    of 2.19 Mpc/h, Philox noise, GPU weight tables), all 13 products and the skewer rows of 96 sightlines spread over
    every slab (halo exchange included).  The single-GPU result is itself held
    to the CPU oracle by tests/ (-m gpu); this check carries that parity over to the sharded configuration the SCALE
    line is measured on.  Returns the worst relative L2 over products and ranks and the worst row difference."""
    import torch
    import torch.distributed as dist
    from saclaymocks_b200.boxes import PRODUCTS
    from saclaymocks_b200.chunk import ChunkPipeline
    ny, nz, dcell = 64, NZ, DCELL
    rng = np.random.default_rng(5)
    nq = 96
    res = {}
    outs = []
    for nr, rk in ((world, rank), (1, 0)):
        pipe = ChunkPipeline(nx, ny, nz, dcell, device=dev, rank=rk, nranks=nr)
        pipe.set_weights({k: pipe.bs.weight_table(k) for k in ("Pln1", "Pln2", "Pln3", "P0")})
        if not outs:
            g = pipe.geom
            half = np.degrees(np.arctan((g.LX / 2 - 4 * dcell) / (g.R0 + g.LZ / 2))) * 0.95
            half_y = np.degrees(np.arctan((g.LY / 2 - 4 * dcell) / (g.R0 + g.LZ / 2))) * 0.9
            ra = (190.0 + rng.uniform(-half, half, nq)).astype("f4")
            dec = rng.uniform(-half_y, half_y, nq).astype("f4")
            z = rng.uniform(2.0, 3.55, nq).astype("f4")
        pipe.set_catalogue(ra, dec, z, 190.0, 0.0)
        pipe.step_boxes(seed=1234)
        pipe.step_skewers(seed=1234)
        torch.cuda.synchronize()
        outs.append(pipe)
    sh, one = outs
    nxl = nx // world
    worst = 0.0
    for name in PRODUCTS:
        a = sh.interior(name).double()
        b = one.interior(name)[rank * nxl:(rank + 1) * nxl].double()
        worst = max(worst, float(((a - b) ** 2).sum().sqrt() / (b ** 2).sum().sqrt()))
    rows = 0.0
    npx = 0
    sel_one = list(one.cat["sel"])
    for i, q in enumerate(sh.cat["sel"]):
        j = sel_one.index(q)
        for k in (0, 1, 3):                                   # delta_l, eta_par, flux
            a, b = sh.out[k][i], one.out[k][j]
            m = ~torch.isnan(a) & (one.out[0][j] > -1e5)
            if bool(m.any()):
                rows = max(rows, float((a[m] - b[m]).abs().max()))
                npx += int(m.sum()) if k == 0 else 0
    t = torch.tensor([worst, rows], dtype=torch.float64, device=dev)
    n = torch.tensor([npx], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n)
    del sh, one, outs
    torch.cuda.empty_cache()
    worst, rows = (float(v) for v in t.cpu())
    return {"what": "sharded (%d ranks, fused exchange) vs single-GPU pipeline, same Philox seed" % world,
            "box": [nx, ny, nz], "products": len(PRODUCTS), "max_rel_l2_boxes": worst, "max_abs_rows": rows,
            "forest_pixels_compared": int(n.item()), "tol_boxes": 1e-5, "tol_rows": 1e-5,
            "ok": bool(worst < 1e-5 and rows < 1e-5 and int(n.item()) > 1000)}


def run_gpu(args):
    real_stdout = quiet_stdout()
    import torch
    import torch.distributed as dist
    from saclaymocks_b200.chunk import ChunkPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nx, ny = box_for(world, args)
    selfcheck = None
    if world > 1 and not args.no_selfcheck:
        selfcheck = parity_selfcheck(world, rank, dev, nx)
    pipe = ChunkPipeline(nx, ny, NZ, DCELL, device=dev, rank=rank, nranks=world)
    W = weight_tables_device(pipe.bs, dev)
    ra, dec, z, ra0, dec0 = synthetic_qsos(nx, ny)
    pipe.set_catalogue(ra, dec, z, ra0, dec0)
    pipe.set_weights(W)
    half = chunk_window(nx)[2]
    pipe.set_footprint(ra0, dec0, half, half * ny / nx)
    cells = nx * ny * NZ
    if world > 1 and args.x_sms is not None:
        pipe.set_x_sms(args.x_sms)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def per_rank(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(v)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    # ---- device-resident arm
    for _ in range(args.warmup):
        pipe.step(seed=42)
    barrier()
    pipe.bs.timing_enable(True)
    pipe.gather_events = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    with ClockSampler(local) as clk:
        barrier()
        t0 = time.time()
        for i in range(args.steps):
            ev[3 * i].record()
            pipe.step_boxes(seed=42 + i)
            ev[3 * i + 1].record()
            pipe.step_skewers(seed=42 + i)
            ev[3 * i + 2].record()
        ev[3 * args.steps].record()
        barrier()
        wall = time.time() - t0
    passes = pipe.bs.timing_collect()
    pipe.bs.timing_enable(False)
    from saclaymocks_b200 import _lib
    seg, back, sbox = _lib.skewers_stats() if pipe.gather_events else (0, 0, (0, 0, 0))
    t_gather = float(np.mean([a.elapsed_time(b) for a, b in pipe.gather_events])) if pipe.gather_events else 0.0
    pipe.gather_events = None
    t_box = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)) / args.steps
    t_skw = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)) / args.steps
    t_tot = ev[0].elapsed_time(ev[3 * args.steps]) / args.steps
    # per-rank stage times: a rank's skewer interval starts when ITS boxes are done, so max(boxes) + max(skewers) may
    # exceed the step; the per-rank lists show who waits for whom
    box_r, skw_r, gat_r = per_rank(t_box), per_rank(t_skw), per_rank(t_gather)
    pix_r = [int(v) for v in per_rank(pipe.cat["own_pixels"])]
    nq_r = [int(v) for v in per_rank(len(pipe.cat["sel"]))]
    t_tot = max(per_rank(t_tot))
    npx = int(sum(pix_r))

    # ---- tuning aid: the box stage with the fused x pass persistent on n CTAs (smk_exchange_set_sms)
    sweep = {}
    if world > 1 and args.x_sms_sweep:
        keep = pipe.x_sms
        for n in [int(v) for v in args.x_sms_sweep.split(",")]:
            pipe.set_x_sms(n)
            pipe.step_boxes(seed=1)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(3):
                pipe.step_boxes(seed=2 + i)
            e1.record()
            barrier()
            sweep[str(n)] = max(per_rank(e0.elapsed_time(e1) / 3))
        pipe.set_x_sms(keep)

    # ---- quasar drawing on the resident boxes (reported beside the step, not part of `value`)
    pipe.draw_qso(seed=1)
    barrier()
    tq = time.time()
    nq_drawn = 0
    for i in range(max(1, min(args.steps, 3))):
        nq_drawn = len(pipe.draw_qso(seed=2 + i)["RA"])
    barrier()
    t_qso = 1e3 * (time.time() - tq) / max(1, min(args.steps, 3))

    # ---- end-to-end arm: pinned host inputs in, all boxes + all spectra rows back to host, every step, through the
    #      same pipelined step as above.  With several ranks every rank stages its own x-slabs and rows over its own
    #      PCIe link; the time is taken between two barriers (= the slowest rank), byte counts are summed over ranks.
    e2e = None
    if not args.no_e2e:
        host = pipe.make_host_buffers(W)
        pipe.step_e2e(host, seed=7)
        barrier()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.time()
        for i in range(n_e2e):
            pipe.step_e2e(host, seed=8 + i)
        barrier()
        dt = (time.time() - t0) / n_e2e
        pipe.step_e2e_resident(host, seed=7)
        barrier()
        t0 = time.time()
        for i in range(n_e2e):
            pipe.step_e2e_resident(host, seed=8 + i)
        barrier()
        dtr = (time.time() - t0) / n_e2e
        tt = torch.tensor([dt, dtr], dtype=torch.float64, device=dev)
        bb = torch.tensor([host["h2d_bytes"], host["d2h_bytes"], 4 * pipe.out[0].numel() * 4 + 64 * nq_drawn],
                          dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bb)
        dt, dtr = (float(v) for v in tt.cpu())
        h2d, d2h, d2h_res = (int(v) for v in bb.cpu())
        e2e = {"value": cells / dt, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * dt, "steps": n_e2e, "skewer_pixels_per_s": npx / dt,
               "what": "P(k) splines + sightline catalogue in from pinned host memory (weight tables evaluated on the "
                       "GPU inside the step), all 13 boxes and the four spectra arrays of every rank copied back to "
                       "pinned host memory through two 1 GiB staging buffers per rank (PCIe-bound); same pipelined "
                       "step as the device-resident arm (fused exchange at N > 1)"}
        e2e["resident"] = {"value": cells / dtr, "unit": "cells/s", "ms_per_step": 1e3 * dtr,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_res,
                           "what": "same chunk with the boxes kept in HBM: P(k) splines + sightlines in, quasars drawn on "
                                   "the resident boxes (smk_draw_qso), quasar table + spectra rows out (the boxes, which "
                                   "make_boxes.py writes to disk, do NOT reach the host in this variant)"}
        del host

    # ---- roofline: every hot kernel against its bound, the one with the largest share of the step as the headline
    peak, how = measured_peak_hbm()
    nk_bytes = pipe.bs.NX * pipe.bs.nyl * pipe.bs.nzh * 8          # this rank's half-complex box
    nr_bytes = pipe.bs.nxl * pipe.bs.NY * pipe.bs.NZ * 4           # this rank's real box
    alg = {"inv_x": 2 * nk_bytes + (4.0 / 13.0) * nk_bytes / 2, "inv_y": 2 * nk_bytes, "c2r_z": nk_bytes + nr_bytes,
           "r2c_z": nk_bytes, "fwd_y": 2 * nk_bytes, "fwd_x": 2 * nk_bytes}
    per = {k: (ms / n if n else 0.0) for k, (ms, n) in passes.items()}
    kernels = {}
    for k in ("inv_x", "inv_y", "c2r_z"):
        if per[k]:
            g = alg[k] / (per[k] * 1e-3) / 1e9
            kernels[k] = {"bound": "hbm", "ms_per_launch": per[k], "launches_per_step": 13, "ms_per_step": 13 * per[k],
                          "achieved": g, "peak": peak, "unit": "GB/s", "frac": g / peak}
    traffic_all, traffic_src = {}, None
    for tag in ("r02",):
        tf = os.path.join(ROOT, "profiles", "traffic_%s.json" % tag)
        if os.path.isfile(tf) and (nx, ny, world) == (512, 512, 1):
            traffic_all, traffic_src = json.load(open(tf)), "profiles/traffic_%s.json (ncu --set full of this workload)" % tag
    sm_mhz_max = clk.summary()["sm_max_mhz"] or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz_max * 1e6 / 1e12            # TFLOP/s: SMs x FMA lanes x 2 x max SM clock
    ig = int(np.argmax(gat_r))              # the rank whose gather takes longest (the fullest slab)
    if gat_r[ig] > 0:
        px0, t_gather = pix_r[ig], gat_r[ig]
        gbs = SKEWER_BYTES_PER_PIXEL * px0 / (t_gather * 1e-3) / 1e9
        tfl = FP32_FLOP_PER_PIXEL * px0 / (t_gather * 1e-3) / 1e12
        kernels["skewers"] = {"bound": "fp32", "ms_per_launch": t_gather, "launches_per_step": 1, "ms_per_step": t_gather,
                              "achieved": tfl, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tfl / fp32_peak,
                              "peak_source": "nominal: 148 SMs x 128 FMA lanes x 2 x %.0f MHz" % sm_mhz_max,
                              "flop_per_pixel": FP32_FLOP_PER_PIXEL, "pixels": int(px0),
                              "hbm_model": {"bytes_per_pixel": SKEWER_BYTES_PER_PIXEL, "achieved": gbs, "peak": peak,
                                            "unit": "GB/s", "frac": gbs / peak}}
    names = {"c2r_z": "c2r_z_kernel", "inv_y": "c2c_strided_kernel<inverse y>", "inv_x": "c2c_strided_kernel<inverse x, fused multiply>",
             "skewers": "skewers gather (smk_skewers_fgpa)"}
    for k, v in kernels.items():
        v["traffic"] = traffic_all.get(k)
    if not kernels:
        kernels["none"] = {"bound": "hbm", "ms_per_step": 0.0, "achieved": 0.0, "peak": peak, "unit": "GB/s", "frac": 0.0}
        names["none"] = "no kernel timed"
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    roofline = dict(kernels[dom])
    roofline.update({"kernel": dom, "kernel_name": names[dom], "peak_source": kernels[dom].get("peak_source", how),
                     "traffic_source": traffic_src, "kernels": kernels,
                     "passes_ms": per, "passes_gbs": {k: (alg[k] / (per[k] * 1e-3) / 1e9 if per.get(k) else None) for k in alg},
                     "chunk_gbs": (332.0 * cells / world) / (max(box_r) * 1e-3) / 1e9,
                     "chunk_frac": (332.0 * cells / world) / (max(box_r) * 1e-3) / 1e9 / peak})
    if world > 1:
        # serial HBM + NVLink model of the box stage: 332 B/cell of HBM traffic per rank at the measured
        # copy peak + 14 exchanges of (N-1)/N of this rank's half-complex box at the measured 770 GB/s per direction
        t_hbm = 332.0 * cells / world / (peak * 1e9) * 1e3
        t_nvl = 14 * (world - 1) / world * nk_bytes / 770e9 * 1e3
        # A rank's box interval also holds its wait for the ranks that are still in the previous step's skewers (the box
        # stage runs in lock step): the rank with the fullest slab never waits, so the smallest interval is the box stage
        t_box_true = min(box_r)
        t_skw_model = SKEWER_BYTES_PER_PIXEL * max(pix_r) / (peak * 1e9) * 1e3
        roofline["boxes_model_ms"] = {"hbm": t_hbm, "nvlink": t_nvl, "serial": t_hbm + t_nvl, "overlapped": max(t_hbm, t_nvl),
                                      "t_boxes_ms_without_waiting": t_box_true,
                                      "frac_serial": (t_hbm + t_nvl) / t_box_true,
                                      "frac_overlapped": max(t_hbm, t_nvl) / t_box_true}
        roofline["chunk_model_ms"] = {"boxes_serial": t_hbm + t_nvl, "skewers_hbm_fullest_slab": t_skw_model,
                                      "frac_of_step": (t_hbm + t_nvl + t_skw_model) / t_tot,
                                      "what": "combined HBM/NVLink roofline of the chunk (serial model of the box stage + "
                                              "byte model of the fullest slab's skewers) over the measured step"}
        roofline["note"] = ("N > 1: the x pass runs on a second stream against the y / z passes of the previous product, "
                            "so the per-pass intervals overlap and each is inflated by the other stream's kernels; the "
                            "single-GPU line carries the per-kernel roofline")

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:          # the CPU baseline is reported at N=1 only
            cores = os.cpu_count()
            r = cpu_oracle_step(REF_SAMPLE_NX, cores)
            cpu = {"value": r["value"], "unit": "cells/s", "cores": cores, "kind": "port",
                   "sample": "256x256x1536 box (1/4 of the cells), 13 products with scipy.fft float32 workers=%d "
                             "(pocketfft stands in for FFTW: pyfftw is not installable) + numba ReadSpec on 24 "
                             "quasars/core extrapolated to the box's %d forest pixels" % (cores, r["npix_full"]),
                   "sample_box": [REF_SAMPLE_NX, REF_SAMPLE_NX, NZ],
                   "t_boxes_s": r["t_boxes"], "t_skewers_s": r["t_skewers_extrapolated"],
                   "skewer_pixels_per_s": r["skewer_pixels_per_s"], "grf_cells_per_s_boxes": r["cells"] / r["t_boxes"]}
        line = {"metric": "grf_cells_per_s", "value": cells / (t_tot * 1e-3), "unit": "cells/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_tot,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(nx, ny, world), "forest_pixels": npx,
                "grf_cells_per_s_boxes": cells / (max(box_r) * 1e-3), "box_cells_per_s": 13 * cells / (max(box_r) * 1e-3),
                "skewer_pixels_per_s": npx / (max(skw_r) * 1e-3), "t_boxes_ms": max(box_r), "t_skewers_ms": max(skw_r),
                "t_gather_ms": max(gat_r),
                "gather_staging_rank0": {"segments_of_128_pixels": seg, "handed_back_to_global_memory_kernel": back,
                                         "box_cells_xyz": list(sbox)},
                "per_rank": {"t_boxes_ms": box_r, "t_skewers_ms": skw_r, "t_gather_ms": gat_r, "forest_pixels": pix_r,
                             "sightlines": nq_r,
                             "skewer_pixels_max_over_mean": max(pix_r) / (sum(pix_r) / world) if npx else None,
                             "t_skewers_max_over_mean": max(skw_r) / (sum(skw_r) / world) if sum(skw_r) else None},
                "t_draw_qso_ms": t_qso, "t_draw_qso_kernel_ms": pipe._qso_drawer.last_kernel_ms,
                "nqso_drawn_rank0": int(nq_drawn),
                "wall_s": wall, "clocks": clk.summary(), "e2e": e2e, "gpu_launches": pipe.launches_per_step * args.steps,
                "roofline": roofline, "cpu_baseline": cpu, "parity_selfcheck": selfcheck,
                "x_sms": getattr(pipe, "x_sms", None), "x_sms_sweep_boxes_ms": sweep or None}
        emit(real_stdout, json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--box", type=int, default=0, help="NX=NY override (default: the BASELINE.json configuration of "
                                                       "this GPU count: 512 / 1024 / 1024 / 2560 at 1 / 2 / 4 / 8)")
    ap.add_argument("--weak", action="store_true", help="round-1 boxes with a fixed number of cells per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the sharded-vs-single-GPU parity check (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--x-sms", type=int, default=None, help="N > 1: run the fused-exchange x pass persistent on this many CTAs")
    ap.add_argument("--x-sms-sweep", default="", help="N > 1: after the timed region, time the box stage for each of these "
                                                      "comma-separated CTA counts of the persistent x pass (0 = one CTA per tile)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
