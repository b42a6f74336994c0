"""Pin the oracle's Eigen-free SE3 / LDLT restatement: the property tests the vendored Sophus suite states
(thirdparty/Sophus/sophus/test_se3.cpp:38-60 fixtures; tests.hpp:70-133 exp/log, group action) and an
independent matrix-exponential check (scipy)."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import expm


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def hat6(a):
    u, w = a[:3], a[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = u
    return M


def se3_from_rotvec_t(L, rotvec, t):
    a = np.zeros(6)
    a[3:] = rotvec
    o = np.zeros(7)
    L.lsdo_se3d_exp(_d(a), _d(o))
    o[4:] = t
    return o


def mat(L, qt):
    R, t = np.zeros(9), np.zeros(3)
    L.lsdo_se3d_matrix(_d(qt), _d(R), _d(t))
    M = np.eye(4)
    M[:3, :3] = R.reshape(3, 3)
    M[:3, 3] = t
    return M


def mul(L, a, b):
    o = np.zeros(7)
    L.lsdo_se3d_mul(_d(a), _d(b), _d(o))
    return o


@pytest.fixture(scope="module")
def se3_vec(oracle):
    """the se3_vec fixtures of test_se3.cpp:38-60"""
    L = oracle.lib()
    P = np.pi
    v = [se3_from_rotvec_t(L, [0.2, 0.5, 0.0], [0, 0, 0]),
         se3_from_rotvec_t(L, [0.2, 0.5, -1.0], [10, 0, 0]),
         se3_from_rotvec_t(L, [0., 0., 0.], [0, 100, 5]),
         se3_from_rotvec_t(L, [0., 0., 0.00001], [0, 0, 0]),
         se3_from_rotvec_t(L, [0., 0., 0.00001], [0, -0.00000001, 0.0000000001]),
         se3_from_rotvec_t(L, [0., 0., 0.00001], [0.01, 0, 0]),
         se3_from_rotvec_t(L, [P, 0, 0], [4, -5, 0])]
    v.append(mul(L, mul(L, se3_from_rotvec_t(L, [0.2, 0.5, 0.0], [0, 0, 0]), se3_from_rotvec_t(L, [P, 0, 0], [0, 0, 0])),
                 se3_from_rotvec_t(L, [-0.2, -0.5, -0.0], [0, 0, 0])))
    v.append(mul(L, mul(L, se3_from_rotvec_t(L, [0.3, 0.5, 0.1], [2, 0, -7]), se3_from_rotvec_t(L, [P, 0, 0], [0, 0, 0])),
                 se3_from_rotvec_t(L, [-0.3, -0.5, -0.1], [0, 6, 0])))
    return v


def test_exp_matches_matrix_exponential(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for scale in (1e-12, 1e-6, 1e-2, 1.0, 3.0):
        for _ in range(10):
            a = rng.standard_normal(6) * scale
            qt = np.zeros(7)
            L.lsdo_se3d_exp(_d(a), _d(qt))
            assert np.allclose(mat(L, qt), expm(hat6(a)), atol=1e-12 * max(1, scale))


def test_exp_log_roundtrip(oracle, se3_vec):
    """tests.hpp:70-86 (expLogTest): exp(log(T)) == T"""
    L = oracle.lib()
    for T in se3_vec:
        lg, back = np.zeros(6), np.zeros(7)
        L.lsdo_se3d_log(_d(T), _d(lg))
        L.lsdo_se3d_exp(_d(lg), _d(back))
        assert np.allclose(mat(L, back), mat(L, T), atol=1e-9)


def test_group_action_and_inverse(oracle, se3_vec):
    """tests.hpp:104-133 (groupActionTest / multiplication vs matrices) + inverse"""
    L = oracle.lib()
    for A in se3_vec:
        inv = np.zeros(7)
        L.lsdo_se3d_inverse(_d(A), _d(inv))
        assert np.allclose(mat(L, mul(L, A, inv)), np.eye(4), atol=1e-9)
        for B in se3_vec:
            assert np.allclose(mat(L, mul(L, A, B)), mat(L, A) @ mat(L, B), atol=1e-9 * 100)


def test_float_flavour_tracks_double(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = (rng.standard_normal(6) * 0.05)
        qd, qf = np.zeros(7), np.zeros(7, np.float32)
        af = a.astype(np.float32)
        L.lsdo_se3d_exp(_d(a), _d(qd))
        L.lsdo_se3f_exp(_f(af), _f(qf))
        assert np.allclose(qf, qd, atol=2e-7)
        b = (rng.standard_normal(6) * 0.05)
        bd, bf = np.zeros(7), np.zeros(7, np.float32)
        L.lsdo_se3d_exp(_d(b), _d(bd))
        L.lsdo_se3f_exp(_f(b.astype(np.float32)), _f(bf))
        pd, pf = np.zeros(7), np.zeros(7, np.float32)
        L.lsdo_se3d_mul(_d(qd), _d(bd), _d(pd))
        L.lsdo_se3f_mul(_f(qf), _f(bf), _f(pf))
        assert np.allclose(pf, pd, atol=5e-7)


def test_ldlt6_solves_spd_systems(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    for _ in range(50):
        J = rng.standard_normal((40, 6)) * np.array([100, 100, 30, 5, 5, 20])
        A = (J.T @ J / 40).astype(np.float32)
        b = rng.standard_normal(6).astype(np.float32) * 10
        x = np.zeros(6, np.float32)
        L.lsdo_ldlt6_solve(_f(np.ascontiguousarray(A)), _f(b), _f(x))
        ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        assert np.allclose(x, ref, rtol=2e-3, atol=1e-5)


def test_mat3_inverse(oracle):
    L = oracle.lib()
    K = np.array([[525, 0, 319.5], [0, 525, 239.5], [0, 0, 1]], np.float32)
    Ki = np.zeros(9, np.float32)
    L.lsdo_mat3_inverse(_f(np.ascontiguousarray(K).reshape(9)), _f(Ki))
    assert np.allclose(Ki.reshape(3, 3) @ K, np.eye(3), atol=1e-5)
