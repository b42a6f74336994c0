"""The product's host/device math header (lsd_slam_b200/csrc/hostmath.h: SE3 / Sim3 group maps, the pose constants of the Sim3
kernel, pivoted LDL^T, 3x3 inverse) compiled for the HOST with g++ (-ffp-contract=off) and compared bit for bit with the
oracle's C restatement of Sophus / Eigen -- the part of the CUDA product that can be executed without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def prog(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hm") / "hostmath_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "hostmath_check.cpp")], check=True)
    return exe


def _run(exe, lines):
    r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    return [np.array([float(t) for t in ln.split()]) for ln in r.stdout.strip().splitlines()]


def _fmt(op, *arrs):
    return op + " " + " ".join(repr(float(v)) for a in arrs for v in np.asarray(a).ravel())


def test_hostmath_header_matches_oracle_bit_for_bit(oracle, prog):
    L = oracle.lib()
    rng = np.random.default_rng(17)
    fp, dp = oracle._fp, oracle._dp
    queries, wants = [], []

    def f32(n, s=1.0):
        return (rng.normal(0, s, n)).astype(np.float32)

    def unit_qt32():
        o = np.zeros(7, np.float32)
        L.lsdo_se3f_exp(fp(f32(6, 0.5)), fp(o))
        return o
    for _ in range(20):
        a = f32(6, 0.3)
        o = np.zeros(7, np.float32); L.lsdo_se3f_exp(fp(a), fp(o))
        queries.append(_fmt("se3exp", a)); wants.append(o.astype(np.float64))
        p, q = unit_qt32(), unit_qt32()
        o = np.zeros(7, np.float32); L.lsdo_se3f_mul(fp(p), fp(q), fp(o))
        queries.append(_fmt("se3mul", p, q)); wants.append(o.astype(np.float64))
        o = np.zeros(7, np.float32); L.lsdo_se3f_inverse(fp(p), fp(o))
        queries.append(_fmt("se3inv", p)); wants.append(o.astype(np.float64))
        t7 = rng.normal(0, [0.5, 0.5, 0.5, 0.3, 0.3, 0.3, 0.2])
        g = np.zeros(8); L.lsdo_sim3d_exp(dp(t7), dp(g))
        queries.append(_fmt("sim3exp", t7)); wants.append(g.copy())
        h = np.zeros(8); L.lsdo_sim3d_exp(dp(rng.normal(0, 0.3, 7)), dp(h))
        o = np.zeros(8); L.lsdo_sim3d_mul(dp(g), dp(h), dp(o))
        queries.append(_fmt("sim3mul", g, h)); wants.append(o.copy())
        o = np.zeros(8); L.lsdo_sim3d_inverse(dp(g), dp(o))
        queries.append(_fmt("sim3inv", g)); wants.append(o.copy())
        R, t, roll = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(4, np.float32)
        L.lsdo_sim3_pose_constants(dp(g), fp(R), fp(t), fp(roll))
        queries.append(_fmt("sim3pose", g)); wants.append(np.concatenate([R, t, roll]).astype(np.float64))
        for n, name in ((6, "lsdo_ldlt6_solve"), (7, "lsdo_ldlt7_solve")):
            J = rng.normal(size=(30, n)) * rng.uniform(0.1, 40, size=n)
            A = np.ascontiguousarray((J.T @ J).astype(np.float32).reshape(n * n))
            b = f32(n, 3.0)
            x = np.zeros(n, np.float32)
            getattr(L, name)(fp(A), fp(b), fp(x))
            queries.append(_fmt("ldlt%d" % n, A, b)); wants.append(x.astype(np.float64))
        m = (np.eye(3) * rng.uniform(100, 600) + rng.normal(0, 30, (3, 3))).astype(np.float32).reshape(9)
        r = np.zeros(9, np.float32); L.lsdo_mat3_inverse(fp(m), fp(r))
        queries.append(_fmt("mat3inv", m)); wants.append(r.astype(np.float64))
    got = _run(prog, queries)
    assert len(got) == len(wants)
    for q, g, w in zip(queries, got, wants):
        assert g.tobytes() == w.tobytes(), (q.split()[0], np.abs(g - w).max())
