#!/usr/bin/env python
"""Extract the static calibration tables of the reference's etc/ directory into one npz.

Run in the build container only (needs /root/reference, which is absent on the GPU box):
    python tests/golden/make_etc_tables.py
Writes saclaymocks_b200/data/etc_tables.npz.  These are *data* (physical calibration
tables read at run time through $SACLAYMOCKS_BASE/etc by the reference), not source code:
  etc/params.fits       z,a,b,c            (bin/merge_spectra.py:58-75)
  etc/dgrowth.fits      Z,dD/dz, key OM    (bin/make_boxes.py:307-313)
  etc/PlanckDR12.fits   K,PK, key ZREF     (py/SaclayMocks/powerspectrum.py:68-88)
  etc/p1dmiss_z*.fits   k,P1DmissRSD       (bin/merge_spectra.py:105-118)
  etc/nz_qso_desi.dat   n(z)               (bench / synthetic QSO catalogues only)
"""
import os, re, sys
import numpy as np

REF = os.environ.get("SACLAYMOCKS_BASE", "/root/reference") + "/etc/"


def read_bintable(path):
    b = open(path, "rb").read()
    off, hdus = 0, []
    while off < len(b):
        hd = {}
        done = False
        while not done:
            blk = b[off:off + 2880]
            off += 2880
            for i in range(0, 2880, 80):
                c = blk[i:i + 80].decode("ascii")
                if c.startswith("END"):
                    done = True
                    break
                m = re.match(r"(HIERARCH\s+)?([\w/\-]+)\s*=\s*('[^']*'|[^/]*)", c)
                if m:
                    hd[m.group(2)] = m.group(3).strip().strip("'").strip()
        nax = int(hd.get("NAXIS", 0))
        size = abs(int(hd.get("BITPIX", 8))) // 8 if nax else 0
        for i in range(1, nax + 1):
            size *= int(hd["NAXIS%d" % i])
        data = b[off:off + size]
        off += (size + 2879) // 2880 * 2880
        hdus.append((hd, data))
    hd, data = hdus[1]
    nf = int(hd["TFIELDS"])
    assert all(hd["TFORM%d" % (i + 1)] == "D" for i in range(nf))
    arr = np.frombuffer(data, dtype=">f8").reshape(int(hd["NAXIS2"]), nf).astype("f8")
    cols = {hd["TTYPE%d" % (i + 1)]: arr[:, i].copy() for i in range(nf)}
    return hd, cols


def main():
    out = {}
    hd, c = read_bintable(REF + "params.fits")
    for k in "zabc":
        out["params_" + k] = c[k]
    hd, c = read_bintable(REF + "dgrowth.fits")
    out["dgrowth_Z"], out["dgrowth_dDdz"], out["dgrowth_OM"] = c["Z"], c["dD/dz"], np.float64(hd["OM"])
    hd, c = read_bintable(REF + "PlanckDR12.fits")
    out["planck_K"], out["planck_PK"], out["planck_ZREF"] = c["K"], c["PK"], np.float64(hd["ZREF"])
    zs = [1.8, 2.2, 2.6, 3.0, 3.6]
    pk = []
    for z in zs:
        hd, c = read_bintable(REF + "p1dmiss_z%.1f.fits" % z)
        out["p1dmiss_k"] = c["k"]
        pk.append(c["P1DmissRSD"])
    out["p1dmiss_z"] = np.array(zs)
    out["p1dmiss_pk"] = np.array(pk)
    nz = np.loadtxt(REF + "nz_qso_desi.dat")
    out["nz_qso_desi"] = nz
    out["qso_lognormal_coef"] = np.loadtxt(REF + "qso_lognormal_coef.txt")[:, :2]     # z, coef (util.py:520-536)
    dst = os.path.join(os.path.dirname(__file__), "..", "..", "saclaymocks_b200", "data", "etc_tables.npz")
    np.savez_compressed(dst, **out)
    print("wrote", os.path.abspath(dst), os.path.getsize(dst), "bytes")
    for k, v in out.items():
        print(k, np.shape(v), np.ravel(v)[:3])


if __name__ == "__main__":
    main()
