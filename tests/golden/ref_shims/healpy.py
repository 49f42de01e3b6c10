"""Shim: healpy is absent; only ang2pix is reached on the hot path (util.radec2pix)."""
from saclaymocks_b200.healpix import ang2pix   # noqa


class pixelfunc(object):
    pass
