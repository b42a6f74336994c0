"""The CUDA path against the COMMITTED golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py:
reference_320x240.npz and the Sim3 entries are outputs of the reference's own sources compiled unmodified -- oracle/_ref --,
the keyframe-output and undistorter entries come from the C oracle): bit-exact for the integer / per-pixel outputs, north_star tolerances for the tracked poses and the map that follows
from them.  Complements the live oracle comparisons of the other -m gpu tests."""
import os
import zlib

import numpy as np
import pytest

from lsd_slam_b200 import abi
from tests.util import IDENT, pose_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_hot_path_against_golden(gpu_ctx_small, seq_small, frames_small):
    gold = np.load(os.path.join(HERE, "golden", "reference_320x240.npz"))
    ctx = gpu_ctx_small
    ctx.upload(0, frames_small[0][0])
    ctx.set_depth_gt(0, frames_small[0][1])
    dm = abi.DepthMap(ctx)
    dm.initializeFromGTDepth(0)
    trk = abi.SE3Tracker(ctx, mode=1)
    assert np.array_equal(ctx.download(0, abi.BUF_MAXGRAD, 0)[100], gold["maxgrad0_row100"])
    assert np.array_equal(ctx.download(0, abi.BUF_IDEPTH, 2), gold["idepth_l2"])
    assert np.array_equal(ctx.download(0, abi.BUF_IDEPTH_VAR, 3), gold["idepthvar_l3"])
    last = IDENT
    for k in range(1, 6):                       # the call sequence of tests/golden/make_golden.py
        ctx.upload(k, frames_small[k][0])
        trk.importFrame(0)
        last = trk.trackFrame(0, k, last)
        dt, ang = pose_err(last, gold["poses_1_5"][k - 1])
        assert dt <= 1e-4 and ang <= 1e-4, (k, dt, ang)
        dm.updateKeyframe([k])
        if k == 5:
            mask = ctx.download(5, abi.BUF_GOODMASK)
            assert (mask != gold["goodmask_f5"]).mean() <= 2e-3
            qts = dm.createKeyFrame(5)
            assert np.abs(qts - gold["new_kf_pose_qts"]).max() <= 1e-4
    cur = dm.current()
    sl = (slice(60, 180, 3), slice(80, 240, 3))
    va, vb = cur["isValid"][sl] > 0, gold["hyp_isValid"] > 0
    assert (va != vb).mean() <= 2e-3
    both = va & vb
    for f in ("idepth", "idepth_smoothed"):
        rel = np.abs(cur[f][sl][both] - gold["hyp_" + f][both]) / np.abs(gold["hyp_" + f][both])
        assert (rel <= 1e-3).mean() >= 0.998, f


def test_8f_rows_against_golden(seq_small, frames_small):
    gold = np.load(os.path.join(HERE, "golden", "oracle_8f_320x240.npz"))
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, max_frames=6)
    for k in (0, 4):
        ctx.upload(k, frames_small[k][0])
        ctx.set_depth_gt(k, frames_small[k][1])
    init = np.concatenate([seq_small.frame_to_ref_qt(4, 0), [1.02]])
    init[4:7] += [0.01, -0.005, 0.004]
    trk = abi.Sim3Tracker(ctx)
    est = trk.trackFrameSim3(0, 4, init, 4, 1)
    g = gold["sim3_frameToRef_qts"]
    dt, ang = pose_err(est[:7], g[:7])
    assert dt <= 1e-4 and ang <= 1e-4 and abs(est[7] - g[7]) <= 1e-4 * g[7]
    assert np.array_equal(np.array([list(trk.last.numCalcResidualCalls), list(trk.last.numCalcWarpUpdateCalls)]), gold["sim3_calls"])
    H = np.array(trk.last.lastSim3Hessian)
    assert np.abs(H - gold["sim3_hessian"]).max() <= 1e-3 * np.abs(gold["sim3_hessian"]).max()
    dm = abi.DepthMap(ctx)
    dm.initializeFromGTDepth(0)
    dm.finalizeKeyFrame()
    a, b, c = ctx.reactivation_data(0)
    assert np.array_equal(b[120], gold["react_var_row120"])
    assert np.array_equal(np.bincount(c.ravel(), minlength=256), gold["react_validity_hist"])
    assert zlib.crc32(ctx.pack_pointcloud(0, 1).tobytes()) == int(gold["pointcloud_l1_crc"][0])
    u = abi.UndistorterPTAM([0.535719308086809, 0.669566858850269, 0.493248545285398, 0.500408664348414, 0.897966326944875],
                            (seq_small.w, seq_small.h), "crop", (seq_small.w, seq_small.h))
    assert np.array_equal(u.K, gold["undist_K"]) and np.array_equal(u.remapX[60], gold["undist_remapX_row60"])
    ctx2 = abi.Context(seq_small.w, seq_small.h, u.getK(), max_frames=4)
    u.install(ctx2)
    assert zlib.crc32(ctx2.undistort(frames_small[0][0]).tobytes()) == int(gold["undist_image_crc"][0])
    ctx2.close()
    ctx.close()
