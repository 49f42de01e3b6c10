"""Small helpers shared by the bin/ scripts (same behaviour as py/SaclayMocks/util.py)."""
import sys

from .cosmo import fgrowth  # noqa: F401
from .healpix import radec2pix  # noqa: F401


def str2bool(v):
    """util.py:93-100."""
    if str(v).lower() in ("yes", "true", "t", "y", "1"):
        return True
    if str(v).lower() in ("no", "false", "f", "n", "0"):
        return False
    print("boolean value expected, got {!r}".format(v))
    sys.exit(1)
