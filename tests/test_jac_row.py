"""track.cuh `jacRowMixed`: rows 3 and 4 of the tracking Jacobian without FP64.

The reference evaluates  v[3] = (-Wx*Wy*z_sqr)*gx + (-(1.0 + Wy*Wy*z_sqr))*gy  with a double literal (SE3Tracker.cpp:1284):
(float)((double)a + (-(1.0 + (double)b)) * (double)g).  The kernel forms the same value from error-free float transformations
(FP64 issues at a small fraction of the FP32 rate on B200).  This test restates that float sequence with numpy float32 scalars
(the FMA emulated exactly in float64: a product of two floats fits in 53 bits) and compares it with the reference expression."""
import numpy as np

F = np.float32


def fma32(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))      # exact product, one rounding: an FMA


def jac_row_mixed(a, b, g):
    ph = F(b * g)
    pl = fma32(b, g, -ph)
    s = F(a - g)
    sv = F(s - a)
    se = F(F(a - F(s - sv)) + F(-g - sv))
    t = F(s - ph)
    tv = F(t - s)
    te = F(F(s - F(t - tv)) + F(-ph - tv))
    return F(t + F(F(se + te) - pl))


def reference_row(a, b, g):
    return F(np.float64(a) + (-(1.0 + np.float64(b))) * np.float64(g))


def test_error_free_float_row_equals_the_reference_double_expression():
    rng = np.random.default_rng(11)
    n = 200000
    # magnitudes of the tracker: a = (Wx Wy / Wz^2) gx, b = (Wy/Wz)^2 in [0, ~0.5], g = focal * image gradient up to a few 1e4
    a = (rng.normal(size=n) * 10.0 ** rng.uniform(-3, 4, n)).astype(F)
    b = (rng.random(n) ** 2 * 0.6).astype(F)
    g = (rng.normal(size=n) * 10.0 ** rng.uniform(-2, 4.5, n)).astype(F)
    bad = 0
    with np.errstate(all="ignore"):
        for i in range(n):
            x, y = jac_row_mixed(a[i], b[i], g[i]), reference_row(a[i], b[i], g[i])
            if x != y:
                bad += 1
                assert abs(np.float64(x) - np.float64(y)) <= np.spacing(abs(y))      # never more than one ulp
    # differences can only come from the reference's two 53-bit roundings (double rounding): ~2^-29 per value
    assert bad <= 2, bad
    # the other row: v[4] = (1.0 + Wx*Wx*z_sqr)*gx + (Wx*Wy*z_sqr)*gy == -jacRowMixed(-a2, b2, gx)
    for i in range(2000):
        x = -jac_row_mixed(F(-a[i]), b[i], g[i])
        y = F((1.0 + np.float64(b[i])) * np.float64(g[i]) + np.float64(a[i]))
        assert x == y or abs(np.float64(x) - np.float64(y)) <= np.spacing(abs(y))
