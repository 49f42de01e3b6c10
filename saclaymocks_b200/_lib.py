"""ctypes binding of libsmk.so (include/smk.h).  There is no CPU fallback: importing a compute entry
point without the CUDA library raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMK_LIB_PATH", os.path.join(_HERE, "libsmk.so"))   # override: kernel-variant experiments

NPRODUCTS = 13
PRODUCT_NAMES = ("boxln_1", "boxln_2", "boxln_3", "box", "eta_xx", "eta_yy", "eta_zz", "eta_xy", "eta_xz", "eta_yz",
                 "vx", "vy", "vz")
PRODUCT_ID = {n: i for i, n in enumerate(PRODUCT_NAMES)}

EXPORTS = ("smk_last_error", "smk_version", "smk_ctx_create", "smk_ctx_destroy", "smk_boxk_pitch", "smk_boxk_elems",
           "smk_box_elems", "smk_workspace_bytes", "smk_sync", "smk_noise_philox", "smk_fft_r2c", "smk_fft_r2c_local",
           "smk_fft_r2c_finish", "smk_synth_c2r", "smk_synth_c2r_local", "smk_synth_c2r_finish",
           "smk_make_boxes_host", "smk_skewers", "smk_skewers_fgpa", "smk_smallscale", "smk_fgpa", "smk_timing_enable", "smk_timing_collect", "smk_pk_weights", "smk_exchange_create", "smk_exchange_handle",
           "smk_exchange_connect", "smk_exchange_ptr", "smk_synth_c2r_local_p2p", "smk_synth_c2r_finish_p2p", "smk_set_stream", "smk_draw_qso", "smk_pk_estimate", "smk_ctx_create_light", "smk_exchange_set_sms", "smk_skewers_stats", "smk_p1d", "smk_fft1d_f64", "smk_set_option")


class SmkError(RuntimeError):
    pass


class Geom(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("dz", C.c_double), ("r0", C.c_double), ("dmax", C.c_int), ("pixel_step", C.c_double),
                ("dir_x_max", C.c_double), ("dir_y_max", C.c_double)]


_lib = None


def lib():
    """Load libsmk.so (built by saclaymocks_b200.build / __graft_entry__.build); raise if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SmkError("libsmk.so not built (%s): run `python -m saclaymocks_b200.build`; there is no CPU fallback"
                       % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, d, u64, sz = C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_size_t
    L.smk_last_error.restype = C.c_char_p
    L.smk_version.restype = i
    L.smk_ctx_create.argtypes = [C.POINTER(vp), i, i, i, d, i, i, vp]
    L.smk_set_option.argtypes = [C.c_char_p, i]
    L.smk_ctx_create_light.argtypes = [C.POINTER(vp), vp]
    L.smk_ctx_destroy.argtypes = [vp]
    L.smk_boxk_pitch.argtypes = [vp]
    L.smk_boxk_elems.argtypes = [vp]
    L.smk_boxk_elems.restype = sz
    L.smk_box_elems.argtypes = [vp]
    L.smk_box_elems.restype = sz
    L.smk_workspace_bytes.argtypes = [vp]
    L.smk_workspace_bytes.restype = sz
    L.smk_sync.argtypes = [vp]
    L.smk_set_stream.argtypes = [vp, vp]
    L.smk_noise_philox.argtypes = [vp, u64, vp]
    L.smk_fft_r2c.argtypes = [vp, vp, u64, vp]
    L.smk_fft_r2c_local.argtypes = [vp, vp, u64, vp]
    L.smk_fft_r2c_finish.argtypes = [vp, vp, vp]
    L.smk_synth_c2r.argtypes = [vp, vp, i, vp, i, d, vp, vp]
    L.smk_synth_c2r_local.argtypes = [vp, vp, i, vp, i, d, vp]
    L.smk_synth_c2r_finish.argtypes = [vp, vp, vp, vp]
    L.smk_make_boxes_host.argtypes = [vp, vp, u64, C.POINTER(vp), d, C.POINTER(vp), C.POINTER(d)]
    L.smk_skewers.argtypes = [vp, C.POINTER(Geom), C.POINTER(vp), i, i, d, d, i, i, i, vp, vp, vp, i, vp, vp, vp]
    L.smk_skewers_fgpa.argtypes = L.smk_skewers.argtypes + [vp, vp, vp, vp, vp, vp]
    L.smk_skewers_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(i)]
    L.smk_smallscale.argtypes = [vp, i, i, i, vp, u64, vp, vp, vp, vp, vp, vp]
    L.smk_p1d.argtypes = [vp, i, i, i, vp, vp, vp, vp, d, vp, vp]
    L.smk_fgpa.argtypes = [vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    L.smk_pk_weights.argtypes = [vp, vp, vp, i, vp]
    L.smk_exchange_create.argtypes = [vp, i]
    L.smk_exchange_handle.argtypes = [vp, i, vp]
    L.smk_exchange_connect.argtypes = [vp, i, vp]
    L.smk_exchange_set_sms.argtypes = [vp, i]
    L.smk_exchange_ptr.argtypes = [vp, i]
    L.smk_exchange_ptr.restype = vp
    L.smk_synth_c2r_local_p2p.argtypes = [vp, vp, i, vp, i, d, i]
    L.smk_synth_c2r_finish_p2p.argtypes = [vp, i, vp, vp]
    L.smk_draw_qso.argtypes = [vp, vp, vp, vp, i]
    L.smk_fft1d_f64.argtypes = [vp, i, vp, vp, vp]
    L.smk_pk_estimate.argtypes = [vp, vp, i, d, d, vp]
    L.smk_timing_enable.argtypes = [vp, i]
    L.smk_timing_collect.argtypes = [vp, C.POINTER(d), C.POINTER(i)]
    for name in EXPORTS:
        if name not in ("smk_last_error", "smk_boxk_elems", "smk_box_elems", "smk_workspace_bytes", "smk_exchange_ptr"):
            getattr(L, name).restype = i
    _lib = L
    return L


class StreamCtx(object):
    """A light library context (smk_ctx_create_light) that follows torch's current stream: `handle()` re-binds it to
    the stream that is current on `device` at the time of the call, so kernels are ordered like any other torch work
    (passing NULL instead would put them on the legacy default stream)."""

    def __init__(self, device):
        import torch
        self._torch = torch
        self.device = device
        self.lib = lib()
        h = C.c_void_p()
        check(self.lib.smk_ctx_create_light(C.byref(h), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        self.h = h

    def handle(self):
        self.lib.smk_set_stream(self.h, C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream))
        return self.h

    def close(self):
        if getattr(self, "h", None):
            self.lib.smk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def skewers_stats(ctx=None):
    """(segments launched by the staged gather, segments handed back to the global-memory kernel, staged box (x, y, z))
    of this thread's last smk_skewers / smk_skewers_fgpa call."""
    seg, back, box = C.c_longlong(), C.c_longlong(), (C.c_int * 3)()
    check(lib().smk_skewers_stats(ctx, C.byref(seg), C.byref(back), box))
    return int(seg.value), int(back.value), tuple(box)


class option(object):
    """with _lib.option("skewers_kernel", 1): ...  -- a library option (smk_set_option) for the duration of a block."""
    DEFAULTS = {"skewers_kernel": 0, "qso_exact": 0, "yz_group": 0, "yz_streams": 2, "yz_persist": 0, "yz_discard": 1}

    def __init__(self, name, value):
        self.name, self.value = name, int(value)

    def __enter__(self):
        check(lib().smk_set_option(self.name.encode(), self.value))
        return self

    def __exit__(self, *a):
        check(lib().smk_set_option(self.name.encode(), self.DEFAULTS[self.name]))


def check(rc):
    if rc != 0:
        raise SmkError("libsmk error %d: %s" % (rc, lib().smk_last_error().decode()))
