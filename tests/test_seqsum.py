"""csrc/seqsum.cuh: the reference's sequential `float sumIdepth += idepth_smoothed` (DepthMap.cpp:1286-1293) computed in
parallel, bit for bit.  The integer model is checked on the host (nvcc-built, no GPU); the kernels on adversarial arrays
against numpy's sequential float32 cumulative sum."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integer_model_of_float_addition_on_the_host():
    exe = os.path.join(tempfile.mkdtemp(prefix="seqsum_"), "seqsum_host_test")
    r = subprocess.run(["nvcc", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tests", "native", "seqsum_host_test.cu")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def seq_sum(x, valid=None):
    x = np.asarray(x, np.float32).ravel()
    if valid is not None:
        x = x[np.asarray(valid).ravel() != 0]
    if x.size == 0:
        return np.float32(0), 0
    with np.errstate(all="ignore"):
        return np.cumsum(x, dtype=np.float32)[-1], int(x.size)


def cases():
    rng = np.random.default_rng(7)
    yield "one", np.array([0.37], np.float32), None
    yield "thirty-one", rng.uniform(0.1, 3, 31).astype(np.float32), None
    yield "run", rng.uniform(0.1, 3, 256).astype(np.float32), None
    yield "run+1", rng.uniform(0.1, 3, 257).astype(np.float32), None
    yield "idepth-like 640x480", rng.gamma(4.0, 0.25, 640 * 480).astype(np.float32), (rng.random(640 * 480) < 0.4)
    yield "idepth-like 1280x1024", rng.gamma(4.0, 0.25, 1280 * 1024).astype(np.float32), (rng.random(1280 * 1024) < 0.4)
    # ties: terms that are odd multiples of half an ulp of the running sum in most binades
    t = (rng.integers(1, 64, 200000) * 2.0 ** -9).astype(np.float32)
    yield "ties 2^-9 grid", t, None
    t = (rng.integers(1, 9, 300000) * 2.0 ** rng.integers(-12, -2, 300000)).astype(np.float32)
    yield "ties mixed grids", t, None
    yield "all equal 1/3", np.full(500000, 1.0 / 3.0, np.float32), None
    yield "all equal 1", np.ones(400000, np.float32), None
    yield "tiny then big", np.concatenate([np.full(1000, 1e-30, np.float32), np.full(1000, 7.25, np.float32)]), None
    yield "denormals", np.full(70000, 1e-41, np.float32), None
    yield "jumps over binades", (2.0 ** rng.integers(-20, 14, 5000)).astype(np.float32), None
    yield "beyond the table (sum > 2^24)", np.full(300000, 100.0, np.float32), None
    yield "absorbed terms", np.concatenate([[2.0 ** 20], np.full(5000, 0.03, np.float32)]).astype(np.float32), None
    x = rng.normal(0.5, 1.0, 100000).astype(np.float32)
    yield "signed", x, None
    x = rng.uniform(0.1, 3, 50000).astype(np.float32)
    x[[100, 20000, 20001, 49999]] = [-5.0, 0.0, -0.0, -1e-3]
    yield "mostly positive, a few specials", x, None
    x = rng.uniform(0.1, 3, 4000).astype(np.float32)
    x[3000] = np.inf
    yield "inf", x, None
    x = rng.uniform(0.1, 3, 4000).astype(np.float32)
    x[1234] = np.nan
    yield "nan", x, None
    yield "zeros", np.zeros(1000, np.float32), None
    yield "nothing valid", np.ones(1000, np.float32), np.zeros(1000, np.uint8)
    yield "cancels to zero and restarts", np.array([1.5, -1.5] * 300 + [0.25] * 700, np.float32), None


@pytest.mark.gpu
def test_sequential_sum_kernels_match_the_sequential_loop_bit_for_bit():
    from lsd_slam_b200 import abi
    ctx = abi.Context(64, 48, np.array([[50, 0, 32], [0, 50, 24], [0, 0, 1]], np.float32), max_frames=2)
    try:
        for name, x, valid in cases():
            want, n = seq_sum(x, valid)
            got, cnt = ctx.seq_sum_f32(x, valid)
            assert cnt == n, name
            assert np.float32(got).tobytes() == np.float32(want).tobytes() or (np.isnan(got) and np.isnan(want)), (name, got, want)
        assert ctx.seq_sum_f32(np.zeros(0, np.float32)) == (np.float32(0), 0)
    finally:
        ctx.close()
