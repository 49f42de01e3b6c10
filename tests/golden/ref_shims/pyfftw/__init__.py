"""Shim: pyFFTW (FFTW 3.3.8 wheel) is absent; same call surface on scipy.fft (pocketfft),
float32 in -> complex64 out, unnormalised forward and *unnormalised* backward like FFTW."""
import numpy as np
import scipy.fft as _sfft
from . import interfaces   # noqa


def empty_aligned(shape, dtype="float32", **_):
    return np.empty(shape, dtype=dtype)


def import_wisdom(w):
    return (True, True, True)


def export_wisdom():
    return (b"", b"", b"")


class FFTW(object):
    def __init__(self, a, b, axes=(-1,), direction="FFTW_FORWARD", threads=1, flags=(), **_):
        self.a, self.b, self.axes, self.direction, self.threads = a, b, axes, direction, threads

    def execute(self):
        if self.direction == "FFTW_FORWARD":
            self.b[...] = _sfft.rfftn(self.a, axes=self.axes, workers=self.threads)
        else:
            s = [self.b.shape[i] for i in self.axes]
            self.b[...] = _sfft.irfftn(self.a, s=s, axes=self.axes, workers=self.threads, norm="forward")

    __call__ = execute
