"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/smk.h declares
(no compute call is made here)."""
import ctypes
import os
import re

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "smk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(smk_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_loads_and_exports_the_header():
    from saclaymocks_b200 import build, _lib
    lib = ctypes.CDLL(build.build())
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsmk.so does not export %s declared in include/smk.h" % n
    for n in _lib.EXPORTS:
        assert n in names, "%s is bound by _lib.py but not declared in include/smk.h" % n
    assert lib.smk_version() >= 100


def test_python_binding_matches_header():
    from saclaymocks_b200 import _lib
    L = _lib.lib()
    for n in declared_functions():
        assert getattr(L, n) is not None


def test_geom_layout_matches_the_header(tmp_path):
    """struct smk_geom: the ctypes mirror has the size and field offsets gcc gives the C declaration."""
    import subprocess
    from saclaymocks_b200._lib import Geom
    src = tmp_path / "szg.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu",'
                   'sizeof(smk_geom),offsetof(smk_geom,dx),offsetof(smk_geom,dmax),offsetof(smk_geom,pixel_step),'
                   'offsetof(smk_geom,dir_y_max));return 0;}\n'
                   % os.path.join(os.path.abspath(ROOT), "include", "smk.h"))
    exe = str(tmp_path / "szg")
    subprocess.check_call(["gcc", str(src), "-o", exe])
    got = [int(v) for v in subprocess.check_output([exe]).split()]
    assert got == [ctypes.sizeof(Geom), Geom.dx.offset, Geom.dmax.offset, Geom.pixel_step.offset, Geom.dir_y_max.offset]


def test_qso_params_layout_matches_the_header(tmp_path):
    """struct smk_qso_params: the ctypes mirror has the size and field offsets gcc gives the C declaration."""
    import subprocess
    from saclaymocks_b200.qso import QsoParams
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu",'
                   'sizeof(smk_qso_params),offsetof(smk_qso_params,x_axis),offsetof(smk_qso_params,sigma_p),'
                   'offsetof(smk_qso_params,u1),offsetof(smk_qso_params,seed));return 0;}\n'
                   % os.path.join(os.path.abspath(ROOT), "include", "smk.h"))
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", str(src), "-o", exe])
    got = [int(v) for v in subprocess.check_output([exe]).split()]
    want = [ctypes.sizeof(QsoParams), QsoParams.x_axis.offset, QsoParams.sigma_p.offset, QsoParams.u1.offset,
            QsoParams.seed.offset]
    assert got == want


def test_product_refuses_to_run_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from saclaymocks_b200 import _lib
    from saclaymocks_b200.boxes import BoxSynth
    with pytest.raises(_lib.SmkError):
        BoxSynth(16, 16, 24, 2.19)


def test_options_and_light_context_need_no_gpu():
    """smk_set_option / smk_ctx_create_light are host-side only: usable (and checked) without a device."""
    from saclaymocks_b200 import _lib
    L = _lib.lib()
    assert L.smk_set_option(b"skewers_kernel", 1) == 0 and L.smk_set_option(b"skewers_kernel", 0) == 0
    assert L.smk_set_option(b"no_such_option", 1) != 0
    assert b"unknown option" in L.smk_last_error()
    with _lib.option("qso_exact", 1):
        pass
    h = ctypes.c_void_p()
    assert L.smk_ctx_create_light(ctypes.byref(h), None) == 0 and h.value
    assert L.smk_boxk_pitch(h) == 0                      # no FFT plan behind a light context
    assert L.smk_exchange_create(h, 2) != 0              # ... and the box entry points refuse it
    assert L.smk_ctx_destroy(h) == 0
