"""Parity at BASELINE.json's full sizes (640x480 and 1280x1024): the track + map loop of both sides, plus the
size-independent properties the domain offers (zero-motion -> identity, determinism run to run)."""
import numpy as np
import pytest

from lsd_slam_b200 import abi, synth
from lsd_slam_b200.stream import GpuStream
from tests.util import IDENT, pose_err

pytestmark = pytest.mark.gpu


def _loop_both(oracle, w, h, n, kf_every=0):
    seq = synth.Sequence(w, h, seed=1234)
    frames = [seq.render(k) for k in range(n)]
    ctx = abi.Context(w, h, seq.K, max_frames=8)
    gs = GpuStream(ctx, mode=1, kf_every=kf_every)
    gs.init_gt(0, frames[0][0], frames[0][1])
    okf = oracle.Frame(0, frames[0][0], seq.K)
    okf.setDepthFromGroundTruth(frames[0][1])
    odm = oracle.DepthMap(w, h, seq.K)
    odm.initializeFromGTDepth(okf)
    last = IDENT
    worst = (0.0, 0.0)
    for k in range(1, n):
        pg = gs.step(k, frames[k][0])
        of = oracle.Frame(k, frames[k][0], seq.K)
        oracle.lib().lsdo_frame_set_depthHasBeenUpdatedFlag(okf.ptr, 0)
        r = oracle.se3_track(okf, of, last)
        last = np.array(r.frameToRef_qt)
        odm.updateKeyframe([of])
        dt, ang = pose_err(pg, last)
        worst = (max(worst[0], dt), max(worst[1], ang))
        assert list(gs.tracker.last.numCalcResidualCalls) == list(r.numCalcResidualCalls), k
    a, b = gs.map.current(), odm.current().copy()     # copy: the oracle view dies with odm
    ctx.close()
    return worst, a, b


@pytest.mark.parametrize("size", [(640, 480), (1280, 1024)])
def test_full_size_loop_parity(oracle, size):
    w, h = size
    n = 6 if w == 640 else 4
    worst, a, b = _loop_both(oracle, w, h, n)
    assert worst[0] <= 1e-4 and worst[1] <= 1e-6, worst                 # pose <= 1e-4 relative (north_star)
    va, vb = a["isValid"] > 0, b["isValid"] > 0
    assert (va != vb).mean() <= 1e-3
    both = va & vb
    rel = np.abs(a["idepth_smoothed"][both] - b["idepth_smoothed"][both]) / np.abs(b["idepth_smoothed"][both])
    assert (rel <= 1e-3).mean() >= 0.999, float((rel <= 1e-3).mean())   # inverse depth <= 1e-3 relative per pixel


def test_full_size_determinism_and_zero_motion():
    seq = synth.Sequence(640, 480, seed=1234)
    f0, d0 = seq.render(0)
    f1, _ = seq.render(1)
    outs = []
    for _ in range(2):
        ctx = abi.Context(640, 480, seq.K, max_frames=4)
        gs = GpuStream(ctx, mode=1, kf_every=0)
        gs.init_gt(0, f0, d0)
        p1 = gs.step(1, f1)
        p0 = gs.step(2, f0)                      # the keyframe image itself, initialised from frame 1's pose
        outs.append((p1, p0, gs.map.current().copy()))
        ctx.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])   # no float atomics anywhere
    assert outs[0][2].tobytes() == outs[1][2].tobytes()
    assert np.abs(outs[0][1][:3]).max() < 2e-4 and np.abs(outs[0][1][4:]).max() < 2e-3          # back at the keyframe
