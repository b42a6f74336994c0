"""Generates tests/golden/oracle_320x240.npz: outputs of the CPU oracle (oracle/lsd_oracle.c, parity flavour)
on the seeded 320x240 synthetic stream.  The reference itself cannot be built or imported in this image
(SURVEY 8c), so these vectors pin the ORACLE (and, through the GPU parity tests, the CUDA path) against
regressions; they are not reference outputs.  Run:  python -m tests.golden.make_golden
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
IDENT = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def compute(oracle, seq, frames):
    out = {}
    img0, d0 = frames[0]
    kf = oracle.Frame(0, img0, seq.K)
    kf.setDepthFromGroundTruth(d0)
    dm = oracle.DepthMap(seq.w, seq.h, seq.K)
    dm.initializeFromGTDepth(kf)
    out["input_frame0_u8_crc"] = np.array([int(img0.astype(np.uint64).sum()), int((img0.astype(np.uint64) * np.arange(img0.size).reshape(img0.shape) % 65521).sum())], np.uint64)
    out["maxgrad0_row100"] = kf.maxGradients(0)[100].copy()
    out["idepth_l2"] = kf.idepth(2).copy()
    out["idepthvar_l3"] = kf.idepthVar(3).copy()
    last = IDENT
    poses = []
    for k in range(1, 6):
        f = oracle.Frame(k, frames[k][0], seq.K)
        r = oracle.se3_track(kf, f, last)
        last = np.array(r.frameToRef_qt)
        poses.append(last)
        kf.L.lsdo_frame_set_depthHasBeenUpdatedFlag(kf.ptr, 0)
        dm.updateKeyframe([f])
        if k == 5:
            out["goodmask_f5"] = f.refPixelWasGood().copy()
            dm.createKeyFrame(f)
            out["new_kf_pose_qts"] = f.thisToParent()
    out["poses_1_5"] = np.array(poses)
    cur = dm.current()
    for n in ("isValid", "blacklisted", "validity_counter", "idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed"):
        out["hyp_" + n] = cur[n][60:180:3, 80:240:3].copy()
    return out


def compute_8f(oracle, seq, frames):
    """Sim3 tracking, keyframe output formats, UndistorterPTAM"""
    import zlib
    out = {}
    kfs = {}
    for k in (0, 4):
        f = oracle.Frame(k, frames[k][0], seq.K)
        f.setDepthFromGroundTruth(frames[k][1])
        kfs[k] = f
    init = np.concatenate([seq.frame_to_ref_qt(4, 0), [1.02]])
    init[4:7] += [0.01, -0.005, 0.004]
    r = oracle.sim3_track(kfs[0], kfs[4], init, 4, 1)
    out["sim3_frameToRef_qts"] = np.array(r.frameToRef_qts)
    out["sim3_hessian"] = np.array(r.lastSim3Hessian, np.float32)
    out["sim3_residuals"] = np.array([r.lastResidual, r.lastDepthResidual, r.lastPhotometricResidual, r.pointUsage,
                                      r.affineEstimation_a, r.affineEstimation_b], np.float32)
    out["sim3_calls"] = np.array([list(r.numCalcResidualCalls), list(r.numCalcWarpUpdateCalls)], np.int32)
    dm = oracle.DepthMap(seq.w, seq.h, seq.K)
    dm.initializeFromGTDepth(kfs[0])
    dm.finalizeKeyFrame()
    a, b, c = kfs[0].reactivation_data()
    out["react_var_row120"] = b[120].copy()
    out["react_validity_hist"] = np.bincount(c.ravel(), minlength=256).astype(np.int64)
    out["pointcloud_l1_crc"] = np.array([zlib.crc32(kfs[0].pack_pointcloud(1).tobytes())], np.uint64)
    u = oracle.UndistorterPTAM([0.535719308086809, 0.669566858850269, 0.493248545285398, 0.500408664348414, 0.897966326944875],
                               (seq.w, seq.h), "crop", (seq.w, seq.h))
    out["undist_K"] = u.K.copy()
    out["undist_remapX_row60"] = u.remapX[60].copy()
    out["undist_image_crc"] = np.array([zlib.crc32(u.undistort(frames[0][0]).tobytes())], np.uint64)
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from lsd_slam_b200 import synth
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.set_globals()
    seq = synth.Sequence(320, 240, seed=1234)
    frames = {k: seq.render(k) for k in range(0, 6)}
    res = compute(pyoracle, seq, frames)
    np.savez_compressed(os.path.join(HERE, "oracle_320x240.npz"), **res)
    print({k: v.shape for k, v in res.items()})
    res = compute_8f(pyoracle, seq, frames)
    np.savez_compressed(os.path.join(HERE, "oracle_8f_320x240.npz"), **res)
    print({k: v.shape for k, v in res.items()})
