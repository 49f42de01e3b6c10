"""Minimal FITS reader/writer with the subset of the `fitsio` API the hot-path scripts use.

The reference does all its file I/O through the `fitsio` wheel (cfitsio), which is not
installable in this image.  The on-disk contracts of SURVEY.md Appendix A only need:
IMAGE HDUs (primary or extension) of float32/float64/int, BINTABLE HDUs with scalar
columns (E, D, J, K, I, B, L, nA), header keywords, and transparent ``.gz``.

API mirrored (same names / argument meaning as fitsio):
    FITS(filename, mode='r'|'rw', clobber=False)      bin/make_boxes.py:111, bin/make_spectra.py:567
    FITS.write(data, header=..., extname=..., names=...)   bin/make_boxes.py:113, bin/make_spectra.py:574-596
    FITS[i|name].read(), .read_header(), .write_key(name, value, comment=)   bin/make_boxes.py:115-116
    read(filename, ext=), read_header(filename, ext=)     bin/make_spectra.py:196,224
Files are assembled in memory and flushed on close(); that is enough for the slab /
spectra files of this pipeline (each at most a few hundred MB).
"""
import gzip
import io
import os
import re

import numpy as np

BLOCK = 2880

_BITPIX = {np.dtype("u1"): 8, np.dtype(">i2"): 16, np.dtype(">i4"): 32, np.dtype(">i8"): 64,
           np.dtype(">f4"): -32, np.dtype(">f8"): -64}
_BITPIX_INV = {8: "u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_TFORM = {"f4": "E", "f8": "D", "i4": "J", "i8": "K", "i2": "I", "u1": "B", "b1": "L"}
_TFORM_INV = {"E": ">f4", "D": ">f8", "J": ">i4", "K": ">i8", "I": ">i2", "B": "u1", "L": "u1"}


class FITSHDR(dict):
    """Case-insensitive keyword dictionary (fitsio.FITSHDR look-alike)."""

    def __init__(self, *a, **k):
        super().__init__()
        self.comments = {}
        for key, val in dict(*a, **k).items():
            self[key] = val

    def __setitem__(self, key, val):
        super().__setitem__(key.upper(), val)

    def __getitem__(self, key):
        return super().__getitem__(key.upper())

    def __contains__(self, key):
        return super().__contains__(key.upper())

    def get(self, key, default=None):
        return super().get(key.upper(), default)


def _fmt_value(val):
    if isinstance(val, (bool, np.bool_)):
        return "%20s" % ("T" if val else "F")
    if isinstance(val, (int, np.integer)):
        return "%20d" % int(val)
    if isinstance(val, (float, np.floating)):
        v = float(val)
        s = repr(v).upper() if np.isfinite(v) else "'%s'" % v
        if "E" not in s and "." not in s and np.isfinite(v):
            s += "."
        return "%20s" % s
    s = str(val).replace("'", "''")
    return "'%-8s'" % s


def _card(name, val, comment=None):
    name = str(name)
    if len(name) > 8 or not re.match(r"^[A-Z0-9_\-]+$", name.upper()):
        head = "HIERARCH %s = " % name
    else:
        head = "%-8s= " % name.upper()
    c = head + _fmt_value(val)
    if comment:
        c += " / " + str(comment)
    return c[:80].ljust(80)


def _parse_value(s):
    s = s.strip()
    if not s:
        return ""
    if s[0] == "'":
        m = re.match(r"'((?:[^']|'')*)'", s)
        return m.group(1).replace("''", "'").rstrip() if m else s.strip("'").rstrip()
    s = s.split("/")[0].strip()
    if s == "T":
        return True
    if s == "F":
        return False
    try:
        return int(s)
    except ValueError:
        pass
    try:
        return float(s.replace("D", "E"))
    except ValueError:
        return s


def _parse_header(buf, off):
    hdr = FITSHDR()
    order = []
    while True:
        blk = buf[off:off + BLOCK]
        if len(blk) < BLOCK:
            raise IOError("truncated FITS header")
        off += BLOCK
        end = False
        for i in range(0, BLOCK, 80):
            c = blk[i:i + 80].decode("ascii", "replace")
            key = c[:8].strip()
            if key == "END":
                end = True
                break
            if key in ("COMMENT", "HISTORY", ""):
                continue
            if c.startswith("HIERARCH"):
                m = re.match(r"HIERARCH\s+(.+?)\s*=\s*(.*)", c)
                if m:
                    hdr[m.group(1).strip()] = _parse_value(m.group(2))
                    order.append(m.group(1).strip().upper())
                continue
            if c[8:10] == "= ":
                rest = c[10:]
                hdr[key] = _parse_value(rest)
                order.append(key)
                if "/" in rest and not rest.strip().startswith("'"):
                    hdr.comments[key] = rest.split("/", 1)[1].strip()
        if end:
            break
    return hdr, off


class _HDU(object):
    """One header-data unit, either parsed from a file (lazy data) or pending a write."""

    def __init__(self, header, raw=None, data=None, kind="IMAGE", extra_keys=None):
        self.header = header
        self._raw = raw          # memoryview of the data bytes (read side)
        self._data = data        # ndarray (write side or decoded)
        self.kind = kind
        self.extra_keys = extra_keys if extra_keys is not None else []   # [(name,val,comment)]

    # ---- fitsio-like API
    def read_header(self):
        h = FITSHDR(self.header)
        for n, v, c in self.extra_keys:
            h[n] = v
        return h

    def write_key(self, name, value, comment=None):
        self.extra_keys.append((name, value, comment))

    def get_extname(self):
        return str(self.header.get("EXTNAME", "")).strip()

    def read(self, columns=None):
        if self._data is not None and self._raw is None:
            return self._data
        h = self.header
        if self.kind == "IMAGE":
            nax = int(h.get("NAXIS", 0))
            if nax == 0:
                return None
            shape = tuple(int(h["NAXIS%d" % i]) for i in range(nax, 0, -1))
            dt = np.dtype(_BITPIX_INV[int(h["BITPIX"])])
            arr = np.frombuffer(self._raw, dtype=dt, count=int(np.prod(shape))).reshape(shape)
            out = arr.astype(dt.newbyteorder("="))
            bzero, bscale = h.get("BZERO", 0), h.get("BSCALE", 1)
            if bzero != 0 or bscale != 1:
                out = out * bscale + bzero
            return out
        # BINTABLE
        nrow = int(h["NAXIS2"])
        nf = int(h["TFIELDS"])
        names, fmts_be, fmts_ne = [], [], []
        for i in range(1, nf + 1):
            tf = str(h["TFORM%d" % i]).strip()
            m = re.match(r"(\d*)([A-Z])", tf)
            rep = int(m.group(1)) if m.group(1) else 1
            code = m.group(2)
            nm = str(h["TTYPE%d" % i]).strip()
            names.append(nm)
            if code == "A":
                fmts_be.append("S%d" % rep)
                fmts_ne.append("S%d" % rep)
            else:
                base = np.dtype(_TFORM_INV[code])
                ne = base.newbyteorder("=")
                if rep == 1:
                    fmts_be.append(base)
                    fmts_ne.append(ne)
                else:
                    fmts_be.append((base, (rep,)))
                    fmts_ne.append((ne, (rep,)))
        dt_be = np.dtype({"names": names, "formats": fmts_be})
        assert dt_be.itemsize == int(h["NAXIS1"]), "unsupported BINTABLE layout"
        arr = np.frombuffer(self._raw, dtype=dt_be, count=nrow)
        out = np.empty(nrow, dtype=np.dtype({"names": names, "formats": fmts_ne}))
        for n in names:
            out[n] = arr[n]
        if columns is not None:
            return out[columns]
        return out

    def __getitem__(self, key):
        return self.read()[key]

    # ---- serialisation
    def _cards(self, primary):
        d = self._data
        cards = []
        if self.kind == "IMAGE":
            if d is None:
                bitpix, shape = 16, ()
            else:
                bitpix, shape = _BITPIX[d.dtype], d.shape
            if primary:
                cards.append(_card("SIMPLE", True, "file does conform to FITS standard"))
            else:
                cards.append(_card("XTENSION", "IMAGE", "IMAGE extension"))
            cards.append(_card("BITPIX", bitpix, "number of bits per data pixel"))
            cards.append(_card("NAXIS", len(shape), "number of data axes"))
            for i, n in enumerate(reversed(shape)):
                cards.append(_card("NAXIS%d" % (i + 1), int(n), "length of data axis %d" % (i + 1)))
            if primary:
                cards.append(_card("EXTEND", True, "FITS dataset may contain extensions"))
            else:
                cards.append(_card("PCOUNT", 0, "required keyword; must = 0"))
                cards.append(_card("GCOUNT", 1, "required keyword; must = 1"))
        else:
            cards.append(_card("XTENSION", "BINTABLE", "binary table extension"))
            cards.append(_card("BITPIX", 8, "8-bit bytes"))
            cards.append(_card("NAXIS", 2, "2-dimensional binary table"))
            cards.append(_card("NAXIS1", d.dtype.itemsize, "width of table in bytes"))
            cards.append(_card("NAXIS2", len(d), "number of rows in table"))
            cards.append(_card("PCOUNT", 0, "size of special data area"))
            cards.append(_card("GCOUNT", 1, "one data group (required keyword)"))
            cards.append(_card("TFIELDS", len(d.dtype.names), "number of fields in each row"))
            for i, n in enumerate(d.dtype.names):
                ft = d.dtype[n]
                if ft.kind == "S":
                    tf = "%dA" % ft.itemsize
                else:
                    tf = _TFORM[ft.newbyteorder("=").str[1:]]
                cards.append(_card("TTYPE%d" % (i + 1), n, "label for field %3d" % (i + 1)))
                cards.append(_card("TFORM%d" % (i + 1), tf))
        for n, v in self.header.items():
            cards.append(_card(n, v, self.header.comments.get(n.upper())))
        for n, v, c in self.extra_keys:
            cards.append(_card(n, v, c))
        cards.append("END".ljust(80))
        s = "".join(cards)
        s += " " * (-len(s) % BLOCK)
        return s.encode("ascii")

    def _tobytes(self, primary):
        out = [self._cards(primary)]
        if self._data is not None:
            b = self._data.tobytes()
            out.append(b)
            out.append(b"\0" * (-len(b) % BLOCK))
        return out


def _to_big_endian_image(a):
    a = np.asarray(a)
    if a.dtype == np.bool_:
        a = a.astype("u1")
    if a.dtype.kind == "f":
        dt = ">f4" if a.dtype.itemsize == 4 else ">f8"
    elif a.dtype.kind in "iu":
        dt = {1: "u1", 2: ">i2", 4: ">i4", 8: ">i8"}[a.dtype.itemsize]
    else:
        raise TypeError("unsupported image dtype %s" % a.dtype)
    return np.ascontiguousarray(a, dtype=dt)


def _to_table(data, names):
    if isinstance(data, np.ndarray) and data.dtype.names:
        names = list(data.dtype.names)
        cols = [data[n] for n in names]
    else:
        cols = [np.asarray(c) for c in data]
    fmts = []
    for c in cols:
        if c.dtype.kind in "SU":
            w = max(1, (c.dtype.itemsize // (4 if c.dtype.kind == "U" else 1)))
            fmts.append("S%d" % w)
        elif c.dtype.kind == "f":
            fmts.append(">f4" if c.dtype.itemsize == 4 else ">f8")
        elif c.dtype.kind in "iu":
            fmts.append({1: "u1", 2: ">i2", 4: ">i4", 8: ">i8"}[c.dtype.itemsize])
        elif c.dtype.kind == "b":
            fmts.append("u1")
        else:
            raise TypeError("unsupported column dtype %s" % c.dtype)
    n = len(cols[0]) if cols else 0
    out = np.zeros(n, dtype=np.dtype({"names": list(names), "formats": fmts}))
    for nm, c in zip(names, cols):
        out[nm] = c.astype("S") if c.dtype.kind == "U" else c
    return out


def _header_from(header):
    h = FITSHDR()
    if header is None:
        return h
    if isinstance(header, dict):
        for k, v in header.items():
            h[k] = v
        return h
    for rec in header:            # list of {'name','value','comment'}
        h[rec["name"]] = rec["value"]
        if rec.get("comment"):
            h.comments[rec["name"].upper()] = rec["comment"]
    return h


class FITS(object):
    def __init__(self, filename, mode="r", clobber=False, **_):
        self.filename = filename
        self.mode = mode
        self.hdus = []
        self._dirty = False
        if mode in ("r", "readonly") or (os.path.exists(filename) and not clobber):
            self._load()

    # ---- read side
    def _load(self):
        opener = gzip.open if self.filename.endswith(".gz") else open
        with opener(self.filename, "rb") as f:
            buf = f.read()
        mv = memoryview(buf)
        off = 0
        while off < len(buf):
            if not bytes(mv[off:off + 8]).strip():
                break
            hdr, off = _parse_header(buf, off)
            nax = int(hdr.get("NAXIS", 0))
            size = 0
            if nax:
                size = abs(int(hdr["BITPIX"])) // 8
                for i in range(1, nax + 1):
                    size *= int(hdr["NAXIS%d" % i])
            size += int(hdr.get("PCOUNT", 0))
            kind = "BINTABLE" if str(hdr.get("XTENSION", "")).strip() == "BINTABLE" else "IMAGE"
            self.hdus.append(_HDU(hdr, raw=mv[off:off + size], kind=kind))
            off += size + (-size % BLOCK)

    def __len__(self):
        return len(self.hdus)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self.hdus[key]
        for h in self.hdus:
            if h.get_extname().upper() == str(key).upper():
                return h
        raise KeyError("no HDU named %s in %s" % (key, self.filename))

    def __iter__(self):
        return iter(self.hdus)

    # ---- write side
    def write(self, data, header=None, extname=None, names=None, **_):
        hdr = _header_from(header)
        if extname is not None:
            hdr["EXTNAME"] = extname
        is_table = names is not None or (isinstance(data, np.ndarray) and data.dtype.names is not None)
        if is_table:
            if not self.hdus:
                self.hdus.append(_HDU(FITSHDR(), data=None, kind="IMAGE"))
            self.hdus.append(_HDU(hdr, data=_to_table(data, names), kind="BINTABLE"))
        else:
            self.hdus.append(_HDU(hdr, data=_to_big_endian_image(data), kind="IMAGE"))
        self._dirty = True

    def close(self):
        if self._dirty and self.mode != "r":
            chunks = []
            for i, h in enumerate(self.hdus):
                if h._raw is not None and h._data is None:
                    h._data = _to_table(h.read(), None) if h.kind == "BINTABLE" else (
                        None if h.read() is None else _to_big_endian_image(h.read()))
                    h._raw = None
                chunks += h._tobytes(primary=(i == 0))
            if self.filename.endswith(".gz"):
                with gzip.open(self.filename, "wb", compresslevel=1) as f:
                    for c in chunks:
                        f.write(c)
            else:
                with open(self.filename, "wb") as f:
                    for c in chunks:
                        f.write(c)
            self._dirty = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read(filename, ext=0, columns=None, header=False):
    f = FITS(filename, "r")
    hdu = f[ext]
    data = hdu.read() if columns is None else hdu.read(columns=columns)
    if header:
        return data, hdu.read_header()
    return data


def read_header(filename, ext=0):
    return FITS(filename, "r")[ext].read_header()


def write(filename, data, header=None, extname=None, names=None, clobber=True):
    f = FITS(filename, "rw", clobber=clobber)
    f.write(data, header=header, extname=extname, names=names)
    f.close()
