"""Philox4x32-10 (saclaymocks_b200/csrc/smk_philox.cuh) against the Random123 known-answer vectors, through the shipped
header compiled for the host, and against an independent Python restatement on random inputs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def philox4x32_10(ctr, key):
    """Salmon et al. 2011, section 3.3 / Random123 philox.h: 10 rounds of the 4x32 S-P network, key bumped by the
    Weyl constants between rounds."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(v) for v in ctr]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k1) & 0xffffffff,
             p0 & 0xffffffff]
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    return tuple(c)


def test_python_restatement_matches_random123_vectors():
    for ctr, key, out in KAT:
        assert philox4x32_10(ctr, key) == out


def test_shipped_header_matches_known_answers(tmp_path):
    if not (os.path.isfile(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "philox_kat")
    r = subprocess.run([NVCC, "-std=c++17", "-O1", os.path.join(HERE, "philox_kat_host.cu"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "0 failed" in r.stdout
    rng = np.random.default_rng(7)
    for _ in range(16):            # random counters / keys: header == Python restatement
        w = rng.integers(0, 2 ** 32, 6, dtype=np.uint64)
        r = subprocess.run([exe] + ["%x" % int(v) for v in w], capture_output=True, text=True)
        got = tuple(int(v, 16) for v in r.stdout.split())
        assert got == philox4x32_10(w[:4], w[4:]), w
