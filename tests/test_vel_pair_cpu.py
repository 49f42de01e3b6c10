"""The velocity products' factor: make_boxes.py:321 multiplies complex64 boxk by `-1j*k/kk*H0*dgrowth0` with a float64
scalar dgrowth0, i.e. numpy forms the product in complex128 and rounds back to complex64.  The x pass
(saclaymocks_b200/csrc/smk_boxes.cu, MUL_VEL) carries f32 * dgrowth0 as an unevaluated float32 pair and rounds v * f
once; this restates that arithmetic operation by operation in numpy (fused multiply-adds through float64, exact for
float32 operands) and checks it against the float64 product on random operands.  Reference: bin/make_boxes.py:403-429."""
import numpy as np


def fma32(a, b, c):
    # a * b is exact in float64 for float32 operands; the sum is rounded once to float64 and once more to float32 (a
    # double rounding that differs from a true float32 FMA only on ties of the second rounding: not hit at these sizes)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def times_pair(v, f32, vscale):
    hi = np.float32(vscale)
    lo = np.float32(vscale - float(hi))
    his = np.full(v.shape, hi, np.float32)
    los = np.full(v.shape, lo, np.float32)
    fh = f32 * his                                   # __fmul_rn(f32, vs_hi)
    fl = fma32(f32, los, fma32(f32, his, -fh))       # __fmaf_rn(f32, vs_lo, __fmaf_rn(f32, vs_hi, -fh))
    pr = v * fh                                      # __fmul_rn(a, fh)
    return pr + fma32(v, fl, fma32(v, fh, -pr))      # __fadd_rn(pr, __fmaf_rn(a, fl, __fmaf_rn(a, fh, -pr)))


def test_float32_pair_reproduces_the_float64_product():
    rng = np.random.default_rng(7)
    n = 4_000_000
    for vscale in (-0.51383754, 0.5318290371, -1.0 / 3.0):     # etc/dgrowth.fits row 0 and two other scalars
        v = (rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)).astype(np.float32)
        f32 = (rng.standard_normal(n) * 10.0 ** rng.uniform(-4, 2, n)).astype(np.float32)
        ref = (v.astype(np.float64) * (f32.astype(np.float64) * vscale)).astype(np.float32)    # numpy's complex128 route
        got = times_pair(v, f32, vscale)
        assert np.count_nonzero(got != ref) == 0
        plain = v * (f32 * np.float32(vscale))       # what a float32-only product would give: wrong in the last bit
        assert np.count_nonzero(plain != ref) > n // 10
