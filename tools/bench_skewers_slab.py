#!/usr/bin/env python
"""The skewer gather of ONE rank of the nominal 8-GPU chunk on one GPU: the x-slab of rank `--rank` of a
2560 x 2560 x 1536 box (random fields: the gather's cost does not depend on the values) with the sightlines of the
full-density synthetic catalogue that touch it, grouped by box class like ChunkPipeline does.
    python tools/bench_skewers_slab.py [--rank 2] [--nx 2560] [--ranks 8]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, default=2)
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--nx", type=int, default=2560)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import bench
    from saclaymocks_b200 import _lib, slab
    from saclaymocks_b200 import spectra as sp
    dev = torch.device("cuda:0")
    nx, nz, dcell = a.nx, 1536, 2.19
    g = sp.SkewerGeometry(nx, nx, nz, dcell)
    hlo, hhi = slab.halo(a.rank, a.ranks, 3)
    nxl = nx // a.ranks
    nxs = nxl + hlo + hhi
    ix0 = a.rank * nxl - hlo
    torch.manual_seed(0)
    fields = [torch.randn((nxs, nx, nz), dtype=torch.float32, device=dev) for _ in range(10)]
    ra, dec, z, ra0, dec0 = bench.synthetic_qsos(nx, nx)
    xyzr, nfor = sp.qso_lines_of_sight(g, ra, dec, z, ra0, dec0)
    xmin, xmax = slab.x_bounds(a.rank, a.ranks, g.LX)
    sel = np.where((nfor >= 0) & slab.touching(xyzr, g.R_vec[0], g.R_vec[-1], xmin, xmax))[0]
    lseg = 127 * g.pixel
    ex = np.floor(lseg * np.abs(xyzr[sel, 0] / xyzr[sel, 3]) / g.DX).astype(np.int64) // 2
    ey = np.floor(lseg * np.abs(xyzr[sel, 1] / xyzr[sel, 3]) / g.DY).astype(np.int64) // 2
    key = ex * 64 + ey
    order = np.argsort(key, kind="stable")
    sel, key = sel[order], key[order]
    starts = np.concatenate(([0], np.where(np.diff(key) != 0)[0] + 1, [len(sel)]))
    groups = [(int(s), int(e)) for s, e in zip(starts[:-1], starts[1:]) if e > s]
    nq, npix = len(sel), g.npixeltot
    q_d = torch.as_tensor(np.ascontiguousarray(xyzr[sel]), device=dev)
    nf_d = torch.as_tensor(np.ascontiguousarray(nfor[sel]), device=dev)
    rvec = torch.as_tensor(g.R_vec, dtype=torch.float64, device=dev)
    out = [torch.full((nq, npix), float("nan"), dtype=torch.float32, device=dev) for _ in range(3)]
    ctx = _lib.StreamCtx(dev)
    L = _lib.lib()
    fl = (C.c_void_p * 10)(*[f.data_ptr() for f in fields])

    def run(record=None):
        for s, e in groups:
            cg = g.c_geom(xyzr[sel[s:e]])
            row = lambda t: C.c_void_p(t.data_ptr() + s * t.stride(0) * t.element_size())
            _lib.check(L.smk_skewers(ctx.handle(), C.byref(cg), fl, ix0, nxs, C.c_double(xmin), C.c_double(xmax), 1, 1,
                                     e - s, row(q_d), row(nf_d), C.c_void_p(rvec.data_ptr()), npix, row(out[0]), row(out[1]),
                                     row(out[2])))
            if record is not None:
                torch.cuda.synchronize()
                record.append((e - s,) + _lib.skewers_stats())
    stats = []
    run(stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    own = int((~torch.isnan(out[0]) & (out[0] > -1e5)).sum())
    print(json.dumps({"rank": a.rank, "ranks": a.ranks, "box": [nx, nx, nz], "sightlines": nq, "forest_pixels_owned": own,
                      "gather_ms": ms, "pixels_per_s": own / (ms * 1e-3), "groups": len(groups),
                      "classes": [{"sightlines": n, "segments": s_, "handed_back": b, "box": list(bx)} for n, s_, b, bx in stats]}))


if __name__ == "__main__":
    main()
