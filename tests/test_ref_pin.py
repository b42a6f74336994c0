"""PIN of the oracle: the hand-written C restatement (oracle/lsd_oracle.c) against oracle/_ref/ -- the reference's OWN
sources (DepthMap.cpp, SE3Tracker.cpp, Sim3Tracker.cpp, TrackingReference.cpp, Frame.cpp, ... and the vendored Sophus)
compiled unmodified by oracle/ref_build.py -- on identical inputs.

* Frame builders, point cloud, stereo constants, and the WHOLE DepthMap (observe / line stereo / fill holes /
  regularise / propagate / createKeyFrame / finalizeKeyFrame) must agree BIT FOR BIT, every field of every pixel.
* Tracking (fp32 sums in another association order on neither side: both are sequential) must agree to fp32
  round-off in every reported quantity, and poses within 1e-5.
The reference-compiled library is prebuilt here (where /root/reference exists) and travels to the GPU box."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import IDENT, pose_err

pytestmark = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built and /root/reference absent")

FIELDS = ("isValid", "blacklisted", "nextStereoFrameMinID", "validity_counter", "idepth", "idepth_var",
          "idepth_smoothed", "idepth_var_smoothed")


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_hyp_identical(a, b, what):
    """isValid and blacklisted on every pixel; every other field, as bit patterns, on every valid pixel.  (The reference's
    DepthMapPixelHypothesis() constructor sets only isValid and blacklisted -- DepthMapPixelHypothesis.h:63-64 -- so the
    other fields of a never-valid pixel are uninitialised heap memory in the reference and carry no information.)"""
    va, vb = a["isValid"] != 0, b["isValid"] != 0
    assert int((va != vb).sum()) == 0, f"{what}: isValid differs on {int((va != vb).sum())} pixels"
    n = int((a["blacklisted"] != b["blacklisted"]).sum())
    assert n == 0, f"{what}: blacklisted differs on {n} pixels"
    for f in FIELDS[2:]:
        x, y = _bits(np.ascontiguousarray(a[f]))[va], _bits(np.ascontiguousarray(b[f]))[va]
        n = int((x != y).sum())
        assert n == 0, f"{what}: field {f} differs on {n} of {int(va.sum())} valid pixels"


class Twin:
    """the same call sequence on the C restatement (fl=False) and on the reference-compiled library (fl='ref')"""

    def __init__(self, seq, frames, init="gt", flavours=(False, "ref")):
        self.seq, self.frames, self.fl = seq, frames, flavours
        self.kf, self.dm, self.fr = {}, {}, {}
        libc = C.CDLL(None)
        for fl in flavours:
            po.set_globals(fl)
            img0, d0 = frames[0]
            kf = po.Frame(0, img0, seq.K, fast=fl)
            dm = po.DepthMap(seq.w, seq.h, seq.K, fast=fl)
            if init == "gt":
                kf.setDepthFromGroundTruth(d0)
                dm.initializeFromGTDepth(kf)
            else:
                libc.srand(1)                       # DepthMap.cpp:898 draws from glibc rand(): same stream on both sides
                dm.initializeRandomly(kf)
            self.kf[fl], self.dm[fl], self.fr[fl] = kf, dm, {0: kf}

    def add_frame(self, k, qts, itr=0.0, mask=None, parent=0):
        for fl in self.fl:
            f = po.Frame(k, self.frames[k][0], self.seq.K, fast=fl)
            f.set_thisToParent(qts, self.fr[fl][parent])
            f.L.lsdo_frame_set_initialTrackedResidual(f.ptr, float(itr))
            if mask is not None:
                f.refPixelWasGood(create=True)[:] = mask
            self.fr[fl][k] = f

    def call(self, fn):
        for fl in self.fl:
            fn(self.dm[fl], self.fr[fl], fl)

    def check(self, what):
        a, b = (self.dm[fl].current() for fl in self.fl)
        assert_hyp_identical(a, b, what)
        return int((a["isValid"] > 0).sum())


def _gt_qts(seq, k):
    return np.concatenate([seq.frame_to_ref_qt(k), [1.0]])


def test_vendored_sophus_suite_passes_on_the_shim():
    """thirdparty/Sophus/sophus/test_{so3,se3,sim3,rxso3}.cpp, compiled unmodified against oracle/ref_shim/Eigen"""
    import os
    import subprocess
    from oracle import ref_build
    ref_build.build()
    for t in ref_build.SOPHUS_TESTS:
        r = subprocess.run([os.path.join(ref_build.OUT, f"test_{t}")], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-500:]
        assert "passed" in r.stderr and "failed" not in r.stderr


def test_frame_builders_bit_exact(seq_small, frames_small):
    """Frame.cpp:35-54, 491-630, 643-680, 690-767, 245-293, 775-877 compiled from the reference vs the restatement"""
    img, d0 = frames_small[0]
    a = po.Frame(0, img, seq_small.K, fast=False)
    b = po.Frame(0, img, seq_small.K, fast="ref")
    a.setDepthFromGroundTruth(d0)
    b.setDepthFromGroundTruth(d0)
    for lvl in range(5):
        assert np.array_equal(a.image(lvl), b.image(lvl)), lvl
        assert np.array_equal(a.gradients(lvl).view(np.uint32), b.gradients(lvl).view(np.uint32)), lvl
        assert np.array_equal(a.idepth(lvl).view(np.uint32), b.idepth(lvl).view(np.uint32)), lvl
        assert np.array_equal(a.idepthVar(lvl).view(np.uint32), b.idepthVar(lvl).view(np.uint32)), lvl
        Ka, Kia = a.K(lvl)
        Kb, Kib = b.K(lvl)
        assert np.array_equal(Ka, Kb) and np.array_equal(Kia.view(np.uint32), Kib.view(np.uint32)), lvl
    # buildMaxGradients: rows 1 and h-2 read never-written pool memory in the reference; interior is defined
    ma, mb = a.maxGradients(0), b.maxGradients(0)
    assert np.array_equal(ma[2:-2].view(np.uint32), mb[2:-2].view(np.uint32))
    assert a.L.lsdo_frame_numPoints(a.ptr) == b.L.lsdo_frame_numPoints(b.ptr)
    assert np.float32(a.L.lsdo_frame_meanIdepth(a.ptr)) == np.float32(b.L.lsdo_frame_meanIdepth(b.ptr))


def test_point_cloud_bit_exact(seq_small, frames_small):
    """TrackingReference::makePointCloud (TrackingReference.cpp:96-147)"""
    img, d0 = frames_small[0]
    a = po.Frame(0, img, seq_small.K, fast=False)
    b = po.Frame(0, img, seq_small.K, fast="ref")
    a.setDepthFromGroundTruth(d0)
    b.setDepthFromGroundTruth(d0)
    for lvl in (1, 2, 3, 4):
        pa, pb = a.point_cloud(lvl), b.point_cloud(lvl)
        assert len(pa[0]) == len(pb[0]) > 100
        for x, y in zip(pa, pb):
            assert np.array_equal(_bits(x), _bits(y)), lvl


def test_prepare_for_stereo_bit_exact(seq_small, frames_small):
    """Frame::prepareForStereoWith (Frame.cpp:295-317): Sim3 inverse, K*R*s, columns of thisToOther_R"""
    rng = np.random.default_rng(5)
    K = np.ascontiguousarray(seq_small.K, np.float32).reshape(9)
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    for trial in range(8):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        qts = np.concatenate([q, rng.normal(size=3) * 0.3, [np.exp(rng.normal() * 0.2) if trial else 1.0]])
        outs = []
        for fl in (False, "ref"):
            kf = po.Frame(0, frames_small[0][0], seq_small.K, fast=fl)
            f = po.Frame(3, frames_small[3][0], seq_small.K, fast=fl)
            o = np.zeros(30, np.float32)
            f.L.lsdo_ref_prepareForStereoWith(f.ptr, kf.ptr, qts.ctypes.data_as(dp), K.ctypes.data_as(fp), o.ctypes.data_as(fp))
            outs.append(o)
        assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32)), (trial, outs[0] - outs[1])


def test_depthmap_update_sequence_bit_exact(seq_small, frames_small):
    """updateKeyframe x6 (1, 2 and 3 reference frames, with and without the tracking mask) -- DepthMap.cpp:111-473,
    1072-1213, 1442-1972, 656-880 and Frame::setDepth, every field of every pixel"""
    t = Twin(seq_small, frames_small, "gt")
    rng = np.random.default_rng(11)
    w1, h1 = seq_small.w // 2, seq_small.h // 2
    groups = [[1], [2, 3], [4], [5, 6, 7], [8], [9]]
    for gi, ks in enumerate(groups):
        for k in ks:
            mask = (rng.random((h1, w1)) > 0.15).astype(np.uint8) if gi % 2 == 0 else None
            t.add_frame(k, _gt_qts(seq_small, k), itr=0.05 * gi, mask=mask)
        t.call(lambda dm, fr, fl: dm.updateKeyframe([fr[k] for k in ks]))
        n = t.check(f"updateKeyframe {ks}")
        # the keyframe's exported depth (Frame::setDepth + pyramids) as the tracker will read it
        for lvl in range(5):
            a, b = (t.kf[fl].idepth(lvl) for fl in t.fl)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
            a, b = (t.kf[fl].idepthVar(lvl) for fl in t.fl)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert n > 15000


def test_depthmap_random_init_every_branch_bit_exact(seq_small, frames_small):
    """random hypotheses drive doLineStereo / observeDepthUpdate through their failure, inconsistency and skip branches"""
    t = Twin(seq_small, frames_small, "random")
    t.check("initializeRandomly")
    for k in (4, 6, 8, 10, 12):
        t.add_frame(k, _gt_qts(seq_small, k))
        t.call(lambda dm, fr, fl: dm.updateKeyframe([fr[k]]))
        n = t.check(f"random-init updateKeyframe [{k}]")
    assert n > 5000


def test_depthmap_individual_passes_bit_exact(seq_small, frames_small):
    """observeDepth, regularizeDepthMapFillHoles, regularizeDepthMap<true/false> one by one (private members of the
    reference class called as they are)"""
    t = Twin(seq_small, frames_small, "gt")
    t.add_frame(2, _gt_qts(seq_small, 2))
    t.add_frame(3, _gt_qts(seq_small, 3))
    t.call(lambda dm, fr, fl: dm.observeDepth([fr[2], fr[3]]))
    t.check("observeDepth")
    t.call(lambda dm, fr, fl: dm.regularizeFillHoles())
    t.check("fillHoles")
    t.call(lambda dm, fr, fl: dm.regularize(False, 24))
    t.check("regularize<false>")
    t.call(lambda dm, fr, fl: dm.regularize(True, 24))
    t.check("regularize<true>")
    ia, ib = (t.dm[fl].integral() for fl in t.fl)
    assert np.array_equal(ia, ib)


def test_create_keyframe_and_finalize_bit_exact(seq_small, frames_small):
    """finalizeKeyFrame + createKeyFrame (propagateDepth with the tracking mask, 2x regularise, fill holes, rescale,
    pose scale) -- DepthMap.cpp:475-653, 1222-1327, 1363-1395; then mapping continues on the new keyframe, including with
    frames that were tracked on the PREVIOUS keyframe (:1085-1099).
    The scene is rescaled so that keyframe 0 already has mean inverse depth 1, as every keyframe after the first has in a real
    run: doLineStereo only accepts a reference frame whose depth ratio to the keyframe is within 0.7..1.4 (:119 `rescaleFactor`),
    so with a 2x scale jump between keyframes the off-parent frames would be rejected pixel by pixel and test nothing."""
    d0 = frames_small[0][1]
    sc = float(np.mean(1.0 / d0[d0 > 0]))
    frames = {k: (frames_small[k][0], (frames_small[k][1] * sc).astype(np.float32)) for k in range(13)}

    def qts(k, ref=0):
        q = seq_small.frame_to_ref_qt(k, ref=ref).copy()
        q[4:7] *= sc
        return np.concatenate([q, [1.0]])
    t = Twin(seq_small, frames, "gt")
    rng = np.random.default_rng(3)
    w1, h1 = seq_small.w // 2, seq_small.h // 2
    for k in (1, 2, 3):
        t.add_frame(k, qts(k), itr=0.1)
        t.call(lambda dm, fr, fl: dm.updateKeyframe([fr[k]]))
    t.check("before finalize")
    t.call(lambda dm, fr, fl: dm.finalizeKeyFrame())
    t.check("finalizeKeyFrame")
    mask = (rng.random((h1, w1)) > 0.1).astype(np.uint8)
    t.add_frame(9, qts(9), itr=0.1, mask=mask)
    t.call(lambda dm, fr, fl: dm.createKeyFrame(fr[9]))
    t.check("createKeyFrame")
    pa, pb = (t.fr[fl][9].thisToParent() for fl in t.fl)          # rescaled Sim3 (DepthMap.cpp:1305)
    assert np.allclose(pa, pb, rtol=0, atol=1e-12), pa - pb
    assert 0.8 < pa[7] < 1.25
    # frames tracked on the PREVIOUS keyframe, mapped on the new one: refToKf comes from the chained absolute poses (:1099), and the
    # tracking-mask gate of :245 / :322 is off for them
    for k in (4, 5):
        t.add_frame(k, qts(k), itr=0.07, mask=mask, parent=0)
    before = t.dm[False].current().copy()
    t.call(lambda dm, fr, fl: dm.updateKeyframe([fr[4], fr[5]]))
    n = t.check("updateKeyframe with frames tracked on the previous keyframe")
    after = t.dm[False].current()
    both = (before["isValid"] != 0) & (after["isValid"] != 0)
    assert n > 10000 and int((before["idepth"][both] != after["idepth"][both]).sum()) > 2000     # they did observe
    # frames tracked on the NEW keyframe
    for k in (10, 11):
        q = qts(k, ref=9)
        q[4:7] /= pa[7]                                           # the new keyframe's units
        t.add_frame(k, q, itr=0.1, parent=9)
        t.call(lambda dm, fr, fl: dm.updateKeyframe([fr[k]]))
        t.check(f"updateKeyframe on new keyframe [{k}]")
    q = qts(12, ref=9)
    q[4:7] /= pa[7]
    t.add_frame(12, q, itr=0.1, parent=9)
    t.add_frame(6, qts(6), itr=0.05, mask=mask, parent=0)
    t.call(lambda dm, fr, fl: dm.updateKeyframe([fr[6], fr[12]]))           # mixed parents in one call
    t.check("updateKeyframe with mixed tracking parents")


def test_se3_single_evaluation_matches_reference_compiled(seq_small, frames_small):
    """calcResidualAndBuffers + calcWeightsAndResidual + calculateWarpUpdate (SE3Tracker.cpp:885-1029, 749-790,
    1258-1299, LGSX.h:184-402), the reference's private members called directly"""
    for fl in (False, "ref"):
        po.set_globals(fl)
    frames = {}
    for fl in (False, "ref"):
        kf = po.Frame(0, frames_small[0][0], seq_small.K, fast=fl)
        kf.setDepthFromGroundTruth(frames_small[0][1])
        frames[fl] = (kf, po.Frame(3, frames_small[3][0], seq_small.K, fast=fl))
    inv = np.zeros(7)
    g = np.ascontiguousarray(seq_small.frame_to_ref_qt(3), np.float64)
    po.lib(False).lsdo_se3d_inverse(g.ctypes.data_as(C.POINTER(C.c_double)), inv.ctypes.data_as(C.POINTER(C.c_double)))
    pose = inv.astype(np.float32)
    for lvl in (4, 3, 2, 1):
        ra = po.se3_eval(*frames[False], lvl, pose, 1.0, 0.0, po.default_track_settings(False), write_mask=(lvl == 1))
        rb = po.se3_eval(*frames["ref"], lvl, pose, 1.0, 0.0, po.default_track_settings("ref"), write_mask=(lvl == 1))
        assert ra.warpedSize == rb.warpedSize > 50
        assert (ra.goodCount, ra.badCount) == (rb.goodCount, rb.badCount)
        assert ra.pointUsage == rb.pointUsage
        for f in ("meanUnweightedRes", "meanWeightedRes", "lsError", "meanRes", "affine_a_lastIt", "affine_b_lastIt"):
            x, y = getattr(ra, f), getattr(rb, f)
            assert abs(x - y) <= 2e-6 * max(abs(y), 1.0), (lvl, f, x, y)
        A, B = np.array(ra.A), np.array(rb.A)
        assert np.abs(A - B).max() <= 2e-6 * np.abs(B).max(), lvl
        assert np.abs(np.array(ra.b) - np.array(rb.b)).max() <= 2e-6 * np.abs(np.array(rb.b)).max() + 1e-9, lvl
    ma, mb = frames[False][1].refPixelWasGood(create=False), frames["ref"][1].refPixelWasGood(create=False)
    assert np.array_equal(ma != 0, mb != 0)


@pytest.mark.parametrize("pair", [(False, "ref"), (True, "ref_sse")], ids=["scalar", "sse"])
def test_se3_track_matches_reference_compiled(seq_small, frames_small, pair):
    """SE3Tracker::trackFrame over a tracked + mapped sequence: pose, every reported statistic, the good-mask.
    'sse' compares the oracle's restatement of the SSE loops (SE3Tracker.cpp:492-575, 1033-1130, LGSX.h:205-386) with the
    stock ENABLE_SSE build of the reference."""
    fa, fb = pair
    res = {}
    for fl in pair:
        po.set_globals(fl, useSSE=1) if fl is True else po.set_globals(fl)
        kf = po.Frame(0, frames_small[0][0], seq_small.K, fast=fl)
        kf.setDepthFromGroundTruth(frames_small[0][1])
        dm = po.DepthMap(seq_small.w, seq_small.h, seq_small.K, fast=fl)
        dm.initializeFromGTDepth(kf)
        st = po.default_track_settings(fl)
        last, out, keep = IDENT, [], []
        for k in range(1, 7):
            f = po.Frame(k, frames_small[k][0], seq_small.K, fast=fl)
            keep.append(f)
            kf.L.lsdo_frame_set_depthHasBeenUpdatedFlag(kf.ptr, 0)
            r = po.se3_track(kf, f, last, st)
            last = np.array(r.frameToRef_qt)
            out.append((last.copy(), r.lastResidual, r.pointUsage, r.lastGoodCount, r.lastBadCount, r.lastMeanRes,
                        r.affineEstimation_a, r.affineEstimation_b, r.diverged, r.trackingWasGood, r.initialTrackedResidual,
                        f.refPixelWasGood(create=False).copy() != 0))
            dm.updateKeyframe([f])
        res[fl] = out
    for k, (x, y) in enumerate(zip(res[fa], res[fb])):
        if fa is False:
            # scalar path, strict IEEE on both sides: the restatement reproduces the reference-compiled tracker BIT FOR BIT
            # (pose as doubles, every statistic, the mask) -- same sequential sums, same LDLT, same Sophus arithmetic
            assert np.array_equal(x[0], y[0]), (k, x[0] - y[0])
            for i in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
                assert np.float32(x[i]) == np.float32(y[i]), (k, i, x[i], y[i])
            assert np.array_equal(x[11], y[11])
            continue
        # timing flavours: both are -O3 builds with FMA contraction left to the compiler, so only fp32 round-off agreement
        dt, ang = pose_err(x[0], y[0])
        assert dt <= 1e-4 and ang <= 1e-6, (k, dt, ang)
        for i in (1, 2, 6, 10):
            assert abs(x[i] - y[i]) <= 2e-4 * max(abs(y[i]), 1.0), (k, i, x[i], y[i])
        assert abs(x[5] - y[5]) <= 1e-2 and abs(x[7] - y[7]) <= 1e-2    # mean signed residual, affine b (grey levels): fp32 cancellation
        assert abs(x[3] - y[3]) <= 3 and abs(x[4] - y[4]) <= 3, (k, x[3:5], y[3:5])       # good / bad counts
        assert x[8] == y[8] and x[9] == y[9]
        assert (x[11] != y[11]).mean() <= 1e-3


def test_sse_vs_scalar_gap_of_the_reference_itself(seq_small, frames_small):
    """The stock build defines ENABLE_SSE (CMakeLists.txt:37-43): _mm_rcp_ps weights, N mod 4 points dropped.  This
    MEASURES how far the reference's two code paths are apart on the same input (documented in DESIGN.md section 2):
    the scalar path is the parity target of the CUDA kernels, the SSE path the timing baseline."""
    out = {}
    for fl in ("ref", "ref_sse"):
        po.set_globals(fl)
        kf = po.Frame(0, frames_small[0][0], seq_small.K, fast=fl)
        kf.setDepthFromGroundTruth(frames_small[0][1])
        f = po.Frame(3, frames_small[3][0], seq_small.K, fast=fl)
        out[fl] = np.array(po.se3_track(kf, f, IDENT, po.default_track_settings(fl)).frameToRef_qt)
    dt, ang = pose_err(out["ref_sse"], out["ref"])
    gt = seq_small.frame_to_ref_qt(3)
    e_scalar, _ = pose_err(out["ref"], gt)
    e_sse, _ = pose_err(out["ref_sse"], gt)
    print(f"reference SSE vs reference scalar: translation {dt:.2e} rel, rotation {ang:.2e} rad; "
          f"vs ground truth: scalar {e_scalar:.2e}, SSE {e_sse:.2e}")
    assert 1e-4 < dt < 5e-2             # both minimise the same cost; they are NOT within 1e-4 of each other
    assert abs(e_scalar - e_sse) < 2e-2 # and neither is closer to the ground truth than the other by more than that


def test_bench_loop_with_keyframe_changes_bit_exact(seq_small, frames_small):
    """the loop bench.py and the full-size GPU tests run (oracle/cpu_stream.py: track, map, forced finalizeKeyFrame +
    createKeyFrame every 10 frames), C oracle vs reference-compiled: every pose equal as doubles, final map identical"""
    from oracle.cpu_stream import CpuStream
    out = {}
    for fl in (False, "ref"):
        cs = CpuStream(seq_small, fl, kf_every=10)
        cs.init_gt(0, *frames_small[0])
        for k in range(1, 26):
            cs.step(k, frames_small[k][0])
        out[fl] = (np.array(cs.poses), cs.dm.current().copy(), cs.kf_changes, cs)
    assert out[False][2] == out["ref"][2] == [10, 20]
    assert np.array_equal(out[False][0], out["ref"][0])
    assert_hyp_identical(out[False][1], out["ref"][1], "after 25 frames and 2 keyframe changes")
