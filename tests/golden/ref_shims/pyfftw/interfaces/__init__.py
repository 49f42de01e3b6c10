from . import numpy_fft   # noqa
