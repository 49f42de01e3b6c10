"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A numpy / scipy.fft / numba restatement of the reference's numerics for the hot path
(bin/make_boxes.py, bin/interpolate_pk.py, bin/make_spectra.py, FGPA half of
bin/merge_spectra.py, and the util/box/constant helpers they call).  Each function cites
the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (saclaymocks_b200/) never does and fails loudly when
its CUDA library is missing.

Pinning: the reference ships no golden vectors.  The oracle is pinned
against outputs of the UNMODIFIED reference scripts executed in the build container with
stand-ins for the absent third-party wheels (tests/golden/run_reference_shimmed.py ->
tests/golden/ref_small.npz, ref_ref32.npz); tests/test_oracle_vs_reference.py checks every
stage.  Third-party arithmetic that is not under /root/reference: pyFFTW>=0.11.1 / FFTW
3.3.8 (requirements.txt:9) is replaced by scipy.fft (pocketfft, float32, same unnormalised
conventions, same c2c-then-c2r order) both in the golden run and here; this is the one
unpinned link and is stated next to every CPU number.
"""
