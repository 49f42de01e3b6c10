"""Host-side check of the FFT core (tests/fft_butterfly_host.cu, compiled with nvcc and run on the CPU): radix
butterflies and every plan -- the default ones and the tuning variants -- against a naive DFT in double."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.parametrize("defines", [(), ("-DSMK_PLAN_512=3", "-DSMK_PLAN_1024=3", "-DSMK_PLAN_2560=3", "-DSMK_PLAN_768=2")])
def test_butterflies_and_plans_against_naive_dft(tmp_path, defines):
    if not (os.path.isfile(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft_host")
    r = subprocess.run([NVCC, "-std=c++17", "-O1", *defines, os.path.join(HERE, "fft_butterfly_host.cu"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "worst plan" in r.stdout
    if defines:         # the three-stage fall-backs of 512 / 1024 and the three-stage 2560 variant
        assert "plan 512 (3 stages, first radix 8)" in r.stdout and "plan 2560 (3 stages, first radix 32)" in r.stdout
        assert "plan 768 (2 stages, first radix 32)" in r.stdout
    else:
        assert "plan 512 (2 stages, first radix 32)" in r.stdout and "plan 1024 (2 stages, first radix 32)" in r.stdout
