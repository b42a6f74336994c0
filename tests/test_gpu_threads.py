"""Thread contract of the boundary (SlamSystem.cpp:111, :206: trackFrame on the caller's thread, the DepthMap methods on the
mapping thread, both on one context): every entry point of the C ABI holds the context's mutex, so calls from several threads
serialise and each sees a consistent context.  The two threads below work on disjoint state (the tracker never reads the
hypotheses; observeDepth / regularize never write the keyframe's tracking depth; they use different frame slots), so their
results must equal, bit for bit, those of the same calls made one after the other."""
import threading

import numpy as np
import pytest

from lsd_slam_b200 import abi

pytestmark = pytest.mark.gpu

ROUNDS = 6


def _setup(seq, frames):
    ctx = abi.Context(seq.w, seq.h, seq.K, device=0, max_frames=20)
    ctx.upload(0, frames[0][0])
    ctx.set_depth_gt(0, frames[0][1])
    dm = abi.DepthMap(ctx)
    dm.initializeFromGTDepth(0)
    # the tracker gets its own copy of the keyframe (id 50): trackFrame counts numFramesTrackedOnThis on the keyframe it tracks on
    # (SE3Tracker.cpp:479-480) and observeDepth reads that counter of the ACTIVE keyframe (DepthMap.cpp:454) -- shared state in the
    # reference as well, which would make the interleaving visible in nextStereoFrameMinID
    ctx.upload(50, frames[0][0])
    ctx.set_depth_gt(50, frames[0][1])
    for k in range(1, 9):
        ctx.upload(k, frames[k][0])                        # tracked
        ctx.upload(100 + k, frames[k][0])                  # mapped (own slots: the tracker rewrites the pose of the frame it tracks)
        ctx.set_pose(100 + k, np.concatenate([seq.frame_to_ref_qt(k), [1.0]]), 0, 0.0)
    return ctx, dm


def _track_all(ctx, seq):
    trk = abi.SE3Tracker(ctx, mode=1)
    out = []
    for _ in range(ROUNDS):
        for k in range(1, 9):
            out.append(np.array(trk.trackFrame(50, k, seq.frame_to_ref_qt(k - 1))))
            ctx.clear_good_mask(k)
    return np.array(out)


def _map_all(dm):
    for _ in range(ROUNDS):
        for k in range(1, 9):
            dm.observeDepth([100 + k])
            dm.regularizeFillHoles()
            dm.regularize(False, 24)
    return dm.current().copy()


def test_tracking_and_mapping_threads_share_one_context(seq_small, frames_small):
    ctx, dm = _setup(seq_small, frames_small)
    poses_seq = _track_all(ctx, seq_small)
    map_seq = _map_all(dm)
    ctx.close()

    ctx, dm = _setup(seq_small, frames_small)
    res, errs = {}, []

    def run(name, fn, *a):
        try:
            res[name] = fn(*a)
        except Exception as e:           # noqa: BLE001 -- surfaced below
            errs.append((name, repr(e)))

    ta = threading.Thread(target=run, args=("poses", _track_all, ctx, seq_small))
    tb = threading.Thread(target=run, args=("map", _map_all, dm))
    ta.start()
    tb.start()
    ta.join(300)
    tb.join(300)
    assert not ta.is_alive() and not tb.is_alive()
    assert not errs, errs
    ctx.close()
    assert np.array_equal(res["poses"], poses_seq)
    a, b = res["map"], map_seq
    assert np.array_equal(a["isValid"], b["isValid"]) and np.array_equal(a["blacklisted"], b["blacklisted"])
    v = a["isValid"] != 0
    assert v.sum() > 10000
    for f in ("idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed"):
        assert np.array_equal(a[f][v].view(np.uint32), b[f][v].view(np.uint32)), f
    for f in ("validity_counter", "nextStereoFrameMinID"):
        assert np.array_equal(a[f][v], b[f][v]), f
