import scipy.fft as _sfft


def rfftn(a, s=None, axes=None, threads=1, **_):
    return _sfft.rfftn(a, s=s, axes=axes, workers=threads)


def irfftn(a, s=None, axes=None, threads=1, **_):
    return _sfft.irfftn(a, s=s, axes=axes, workers=threads)


def rfft(a, n=None, axis=-1, threads=1, **_):
    return _sfft.rfft(a, n=n, axis=axis, workers=threads)


def irfft(a, n=None, axis=-1, threads=1, **_):
    return _sfft.irfft(a, n=n, axis=axis, workers=threads)


def fft(a, n=None, axis=-1, threads=1, **_):
    return _sfft.fft(a, n=n, axis=axis, workers=threads)
