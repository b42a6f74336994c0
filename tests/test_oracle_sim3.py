"""Pin the oracle's Sim3 restatement: the Sophus Sim3 group maps against an independent matrix
exponential on the fixtures of thirdparty/Sophus/sophus/test_sim3.cpp:43-84, the 7x7 LDLT against numpy, the roll
constants of Sim3Tracker.cpp:451-460 against their closed form, and trackFrameSim3 (Sim3Tracker.cpp:149-382) on a
synthetic keyframe pair with a known relative pose and a known scale factor."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import expm

from lsd_slam_b200 import synth


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def hat7(a):
    u, w, s = a[:3], a[3:6], a[6]
    M = np.zeros((4, 4))
    M[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) + s * np.eye(3)
    M[:3, 3] = u
    return M


def quat_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def mat(qts):
    M = np.eye(4)
    M[:3, :3] = qts[7] * quat_R(qts[:4])
    M[:3, 3] = qts[4:7]
    return M


TANGENTS = [[0, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0], [0, 1, 0, 1, 0, 0, 0.1], [0, 0, 1, 0, 1, 0, 0.1],
            [-1, 1, 0, 0, 0, 1, -0.1], [20, -1, 0, -1, 1, 0, -0.1], [30, 5, -1, 20, -1, 0, 1.5],
            [0.1, 0.2, 0.3, 0, 0, 1e-12, 0.2], [0.1, 0.2, 0.3, 0.3, -0.2, 0.1, 1e-13]]


def test_sim3_exp_matches_matrix_exponential(oracle):
    L = oracle.lib()
    for a in TANGENTS:
        a = np.array(a, np.float64)
        o = np.zeros(8)
        L.lsdo_sim3d_exp(_d(a), _d(o))
        assert np.allclose(mat(o), expm(hat7(a)), rtol=1e-9, atol=1e-9), a
        assert abs(np.linalg.norm(o[:4]) - 1) < 1e-12


def test_sim3_mul_inverse(oracle):
    L = oracle.lib()
    gs = []
    for a in TANGENTS:
        o = np.zeros(8)
        L.lsdo_sim3d_exp(_d(np.array(a, np.float64)), _d(o))
        gs.append(o)
    for a in gs:
        inv = np.zeros(8)
        L.lsdo_sim3d_inverse(_d(a), _d(inv))
        assert np.allclose(mat(inv) @ mat(a), np.eye(4), atol=1e-9)
        for b in gs:
            ab = np.zeros(8)
            L.lsdo_sim3d_mul(_d(a), _d(b), _d(ab))
            assert np.allclose(mat(ab), mat(a) @ mat(b), rtol=1e-9, atol=1e-8)


def test_ldlt7_matches_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(5)
    for _ in range(20):
        J = rng.normal(size=(40, 7)) * rng.uniform(0.1, 30, size=7)
        A = (J.T @ J).astype(np.float32)
        b = rng.normal(size=7).astype(np.float32)
        x = np.zeros(7, np.float32)
        L.lsdo_ldlt7_solve(_f(np.ascontiguousarray(A).reshape(49)), _f(b), _f(x))
        ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
        assert np.allclose(x, ref, rtol=2e-3, atol=1e-5 * np.abs(ref).max())


def test_pose_constants_roll(oracle):
    """pure rotation about the optical axis by phi -> rollMat = that rotation; scale multiplies rotMat only"""
    L = oracle.lib()
    for phi, s in ((0.0, 1.0), (0.3, 1.0), (-0.7, 1.3)):
        qts = np.array([0, 0, np.sin(phi / 2), np.cos(phi / 2), 0.1, -0.2, 0.3, s])
        R, t, roll = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(4, np.float32)
        L.lsdo_sim3_pose_constants(_d(qts), _f(R), _f(t), _f(roll))
        assert np.allclose(roll, [np.cos(phi), -np.sin(phi), np.sin(phi), np.cos(phi)], atol=1e-6)
        assert np.allclose(R.reshape(3, 3), s * quat_R(qts[:4]), atol=1e-6)
        assert np.allclose(t, qts[4:7], atol=1e-7)
    # a tilt about x has no roll component: the 2x2 block stays close to identity in its first row
    qts = np.array([np.sin(0.1), 0, 0, np.cos(0.1), 0, 0, 0, 1.0])
    R, t, roll = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(4, np.float32)
    L.lsdo_sim3_pose_constants(_d(qts), _f(R), _f(t), _f(roll))
    assert abs(roll[0] - 1) < 1e-6 and abs(roll[1]) < 1e-6 and abs(roll[2]) < 1e-6


def _keyframes(oracle, scale_b=1.0, w=320, h=240, ka=0, kb=6):
    seq = synth.Sequence(w, h, seed=1234)
    fs = []
    for k, sc in ((ka, 1.0), (kb, scale_b)):
        img, z = seq.render(k)
        f = oracle.Frame(k, img, seq.K)
        f.setDepthFromGroundTruth((z * sc).astype(np.float32))
        fs.append(f)
    return seq, fs[0], fs[1]


def test_track_frame_sim3_recovers_pose_and_scale(oracle):
    seq, A, B = _keyframes(oracle)
    gt = seq.frame_to_ref_qt(6, 0)
    init = np.concatenate([gt, [1.0]])
    init[4:7] += [0.01, -0.005, 0.004]
    init[7] = 1.03
    r = oracle.sim3_track(A, B, init, 4, 1)
    est = np.array(r.frameToRef_qts)
    assert not r.diverged
    assert abs(est[7] - 1.0) < 2e-3
    assert np.linalg.norm(est[4:7] - gt[4:7]) < 0.05 * np.linalg.norm(gt[4:7]) + 1e-3
    assert abs(abs(np.dot(est[:4], gt[:4])) - 1) < 1e-6
    assert r.numCalcResidualCalls[4] > 0 and r.numCalcResidualCalls[1] > 0 and r.numCalcResidualCalls[0] == 0
    H = np.array(r.lastSim3Hessian).reshape(7, 7)
    assert np.allclose(H, H.T, rtol=1e-5) and (np.diag(H) > 0).all()
    assert r.lastResidual > 0 and r.lastDepthResidual > 0 and r.lastPhotometricResidual > 0
    assert 0.8 < r.pointUsage <= 1.0


def test_track_frame_sim3_estimates_scale(oracle):
    """frame B carries a map 1.05x larger -> frameToReference.scale() = 1/1.05"""
    seq, A, B = _keyframes(oracle, scale_b=1.05)
    gt = seq.frame_to_ref_qt(6, 0)
    r = oracle.sim3_track(A, B, np.concatenate([gt, [1.0]]), 4, 1)
    assert not r.diverged
    assert abs(r.frameToRef_qts[7] - 1 / 1.05) < 5e-3


def test_track_frame_sim3_diverges_without_overlap(oracle):
    seq, A, B = _keyframes(oracle)
    far = np.array([0, 0, 0, 1, 50.0, 0, 0, 1.0])             # 50 units sideways: nothing projects into the image
    r = oracle.sim3_track(A, B, far, 4, 1)
    assert r.diverged
    assert np.allclose(r.frameToRef_qts, [0, 0, 0, 1, 0, 0, 0, 1])
    assert not np.any(np.array(r.lastSim3Hessian))


def test_sim3_eval_consistency(oracle):
    """LGS7 = LGS6 + remapped LGS4 (LGSX.h:422-441): symmetric, num_constraints = 2 * warped size"""
    seq, A, B = _keyframes(oracle)
    gt = seq.frame_to_ref_qt(6, 0)
    inv = np.zeros(8)
    oracle.lib().lsdo_sim3d_inverse(_d(np.concatenate([gt, [1.0]])), _d(inv))
    e = oracle.sim3_eval(A, B, 2, inv)
    A7 = np.array(e.A).reshape(7, 7)
    assert np.allclose(A7, A7.T, rtol=1e-6)
    assert e.num_constraints == 2 * e.warpedSize and e.numTermsP == e.warpedSize and 0 < e.numTermsD <= e.numTermsP
    assert A7[6, 6] > 0 and A7[6, 0] == 0 and A7[6, 1] == 0 and A7[6, 5] == 0      # sigma only couples with rows 2,3,4
    assert abs(e.mean - (e.sumResD + e.sumResP) / (e.numTermsD + e.numTermsP)) < 1e-6 * e.mean


def test_sse_flavour_agrees_with_scalar(oracle, seq_small, frames_small):
    """the reference's SSE variants of the weights / LGS loops (rcp approximations, N mod 4 tail dropped) give the same
    tracking up to their approximation error"""
    seq, A, B = _keyframes(oracle)
    init = np.concatenate([seq.frame_to_ref_qt(6, 0), [1.02]])
    init[4:7] += [0.01, -0.005, 0.004]
    ref = oracle.sim3_track(A, B, init, 4, 1)
    oracle.set_globals(useSSE=1)
    try:
        sse = oracle.sim3_track(A, B, init, 4, 1)
    finally:
        oracle.set_globals()
    a, b = np.array(ref.frameToRef_qts), np.array(sse.frameToRef_qts)
    assert np.abs(a - b).max() < 2e-4 and not sse.diverged
    assert abs(sse.lastResidual - ref.lastResidual) < 5e-3 * ref.lastResidual
    Hs, Hr = np.array(sse.lastSim3Hessian), np.array(ref.lastSim3Hessian)
    assert np.abs(Hs - Hr).max() < 5e-3 * np.abs(Hr).max()
