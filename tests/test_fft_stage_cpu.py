"""The FFT passes' shared-memory stage code, run thread by thread on the host.

tests/fft_stage_host.cpp includes the kernels' own stage code (saclaymocks_b200/csrc/smk_ztile.cuh, smk_fft.cuh) through a
stand-in cuda_runtime.h and compares every line with a float64 inverse DFT, for every tile shape the C ABI dispatches
(NZ/2 = 32 ... 2048) and for both twiddle sources (SMK_Z_TW=2: compact first-stage table + compile-time constants, 1: split from the W table, 0: table loads).  It
also counts shared-memory wavefronts per request (2.0 = free of bank conflicts) and the L1 wavefronts of the twiddle loads.
The same harness runs the stage sequence of the strided x / y passes (tile shapes of StridedTraits, up to the 2560-point
plan 16.16.2.5).  The transforms replace FFTW's r2c / c2r of make_boxes.py:53, 87 (reference); numerics on the GPU are
covered by test_gpu_boxes.py / test_gpu_sizes.py.
"""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("z_tw", [2, 1, 0])
def test_c2r_stages_on_host(tmp_path, z_tw):
    exe = str(tmp_path / "fft_stage_host")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas", "-DSMK_Z_TW=%d" % z_tw,
                           "-I" + os.path.join(HERE, "host_stub"), "-I" + os.path.join(ROOT, "saclaymocks_b200", "csrc"),
                           "-o", exe, os.path.join(HERE, "fft_stage_host.cpp")])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout.splitlines()
    strided = [l for l in out if l.startswith("strided N=")]   # x / y passes: first stage from global, last stage back
    assert len(strided) == 9 and all(l.endswith(" ok") for l in strided), res.stdout
    tiles = [l for l in out if l.startswith("tile M=")]        # forward tiles and the inverse tiles that are not fused
    assert len(tiles) == 17 and all(l.endswith(" ok") for l in tiles), res.stdout
    checked = [l for l in out if l.startswith("M=")]           # fused inverse tiles (c2r_stages + c2r_tail / c2r_last_butterfly)
    assert len(checked) == 7 and all(l.endswith(" ok") for l in checked), res.stdout
    for l in checked:
        wf = [float(x) for x in re.search(r"lds_wavefronts_per_request=([\d.]+),([\d.]+)", l).groups()]
        # no bank conflicts in any tile shape (the table-load build keeps the old task order, which conflicts two-way
        # in the second stage of the 4-line tile: one more reason it is not the default)
        assert wf[0] == 2.0 and (wf[1] in (0.0, 2.0) or z_tw == 0), l
    (m768,) = [l for l in checked if l.startswith("M=768 ")]
    tw_wf = sum(int(x) for x in re.search(r"twiddle_wavefronts_per_tile=(\d+),(\d+)", m768).groups())
    data_wf = int(re.search(r"tile data: (\d+)", m768).group(1))    # wavefronts of one pass over the tile
    # NZ = 1536: twiddle table loads cost several passes' worth of wavefronts; the compact table a fraction of one
    assert tw_wf / data_wf < (0.2, 2.5, 5.0)[2 - z_tw] and (z_tw or tw_wf / data_wf > 2.5), m768
    assert "tail=1 jfast=1" in m768 or z_tw == 0        # the last two stages run in registers (c2r_tail)
