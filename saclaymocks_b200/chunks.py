"""Chunk layout of the mocks: sky window (ra0, dra, dec0, ddec), chunk ids and slab count per box size.

Restates the table of bin/submit_mocks.py:611-674 (`chunk_parameters`), which the orchestrator uses to generate the
draw_qso.py / make_spectra.py command lines (submit_mocks.py:1246, 883-923).  The nominal size is 2560 cells with the
7 chunks of the DESI footprint; the other sizes are the reference's debugging boxes.  Values are kept as the decimal
strings the reference pastes into its shell scripts, so that float(...) of them is exactly what the scripts parse.
"""
import numpy as np

# cells -> (nslice, [(chunkid, ra0, dra, dec0, ddec), ...])
_H = "32.2413248675"      # half-width of a nominal chunk
_LAYOUT = {
    2560: (512, [("1", "125.5", _H, "20", _H), ("2", "189.982649735", _H, "20", _H), ("3", "254.46529947", _H, "20", _H),
                 ("4", "140", "52.6529106357", "64.4956624338", "12.2543375662"),
                 ("5", "245.305821271", "52.6529106357", "64.4956624338", "12.2543375662"),
                 ("6", "-24", _H, "7", _H), ("7", "40.4826497351", _H, "7", _H)]),
    256: (8, [(str(i + 1), r, "3.17", "20", "3.17") for i, r in enumerate(
        ("189.982649735", "183.642649735", "177.302649735", "170.962649735", "196.322649735", "202.662649735"))]),
    512: (8, [(str(i + 1), r, "6.4", d, "6.4") for i, (r, d) in enumerate(
        [(r, d) for d in ("0", "12.8") for r in ("190", "202.8", "177.2", "215.6")])]),
    1024: (32, [("1", "190", "12.7", "0", "12.7")]),
    128: (8, [("1", "190", "1.6", "0", "1.6")]),
    32: (2, [("1", "190", "1.6", "0", "1.6")]),
}
_STRIPE = (512, [("1", "-25", "25", "0", "2"), ("2", "25", "25", "0", "2")])


def chunk_parameters(cells, stripe_footprint=False):
    """(ra0, dra, dec0, ddec, chunkid, nslice) as numpy arrays of strings (nslice: 0-d integer array), exactly what
    submit_mocks.py:611-674 returns.  An unknown box size raises (the reference fails with UnboundLocalError)."""
    if stripe_footprint:
        nslice, rows = _STRIPE
    elif cells in _LAYOUT:
        nslice, rows = _LAYOUT[cells]
    else:
        raise ValueError("chunk_parameters: no chunk layout for a box of %r cells (known: %s)"
                         % (cells, sorted(_LAYOUT)))
    cid, ra0, dra, dec0, ddec = (np.array(c) for c in zip(*rows))
    return ra0, dra, dec0, ddec, cid, np.array(nslice)


def chunk_window(cells, chunk=1, stripe_footprint=False):
    """(ra0, dra, dec0, ddec) of one chunk as floats (what draw_qso.py / make_spectra.py parse from their CLI)."""
    ra0, dra, dec0, ddec, cid, _ = chunk_parameters(cells, stripe_footprint)
    i = list(cid).index(str(chunk))
    return float(ra0[i]), float(dra[i]), float(dec0[i]), float(ddec[i])


def chunk_ids(cells, stripe_footprint=False):
    return [int(c) for c in chunk_parameters(cells, stripe_footprint)[4]]
