"""Shim: the reference imports `fitsio` (cfitsio wheel, absent here); route to fitsio_lite."""
from saclaymocks_b200.fitsio_lite import *          # noqa
from saclaymocks_b200.fitsio_lite import FITS, FITSHDR, read, read_header, write   # noqa
