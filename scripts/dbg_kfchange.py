"""Diagnostic (GPU): where do the CUDA path and the oracle part ways over a keyframe change at full size?"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from lsd_slam_b200 import abi, synth
from lsd_slam_b200.stream import GpuStream
from oracle import pyoracle as po
from oracle.cpu_stream import CpuStream
from tests.util import pose_err, hyp_equal_report

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
seq = synth.Sequence(W, H, seed=1234)
N = 22
fr = [seq.render(k) for k in range(N)]
po.set_globals(False)
ctx = abi.Context(W, H, seq.K, max_frames=8)
gs = GpuStream(ctx, mode=1, kf_every=0, fused_call=False)
gs.init_gt(0, *fr[0])
cs = CpuStream(seq, False, kf_every=0)
cs.init_gt(0, *fr[0])


def cmp(tag):
    a, b = gs.map.current(), cs.dm.current().copy()
    rep = hyp_equal_report(a, b)
    va, vb = a["isValid"] > 0, b["isValid"] > 0
    both = va & vb
    rel = np.abs(a["idepth"][both] - b["idepth"][both]) / np.abs(b["idepth"][both])
    print(f"{tag}: valid_mismatch {rep['valid_mismatch']} bl {rep['blacklist_mismatch']} vc {rep['validity_mismatch']} idepth bitdiff {rep['idepth_bitdiff']} "
          f"frac<=1e-3 {(rel <= 1e-3).mean():.6f} frac<=1e-5 {(rel <= 1e-5).mean():.6f} max {rel.max():.3e} n {both.sum()}", flush=True)


for k in range(1, 20):
    pg = gs.step(k, fr[k][0])
    pc, _ = cs.step(k, fr[k][0])
    r = cs.results[-1]
    same = list(gs.tracker.last.numCalcResidualCalls) == list(r.numCalcResidualCalls)
    print(k, "pose err", pose_err(pg, pc), "counts same", same, flush=True)
    if k in (1, 5, 10, 15, 19):
        cmp(f"after frame {k}")

# frame 20 by hand on both sides
k = 20
ctx.upload(k, fr[k][0])
gs.tracker.importFrame(0)
pg = gs.tracker.trackFrame(0, k, gs.last_pose)
f = po.Frame(k, fr[k][0], seq.K)
cs.L.lsdo_frame_set_depthHasBeenUpdatedFlag(cs.kf.ptr, 0)
r = po.se3_track(cs.kf, f, cs.last, cs.st)
pc = np.array(r.frameToRef_qt)
print("frame 20 pose err", pose_err(pg, pc))
mg, mo = ctx.download(k, abi.BUF_GOODMASK), f.refPixelWasGood()
print("mask mismatch", int((mg != (mo != 0)).sum()), "of", mg.size, "zeros g/o", int((mg == 0).sum()), int((mo == 0).sum()))
gs.map.finalizeKeyFrame(); cs.dm.finalizeKeyFrame()
cmp("after finalizeKeyFrame")
# (A) each side's own state: propagateDepth only
bo = cs.dm.current().copy(); bg = gs.map.current().copy()
cs.dm.propagateDepth(f); gs.map.propagateDepth(k)
cmp("own state: after propagateDepth")
# (B) identical state + identical pose + identical mask: GPU state := oracle state
cs.dm.set_current(bo)
gs.map.setHypotheses(0, bo, do_set_depth=False)
ctx.set_pose(k, f.thisToParent(), 0, r.initialTrackedResidual)
cs.dm.propagateDepth(f); gs.map.propagateDepth(k)
cmp("oracle state+pose, own masks: after propagateDepth")
cs.dm.set_current(bo)
gs.map.setHypotheses(0, bo, do_set_depth=False)
cs.dm.createKeyFrame(f)
q = gs.map.createKeyFrame(k)
print("rescale", q[7], f.thisToParent()[7], abs(q[7] - f.thisToParent()[7]) / q[7])
cmp("oracle state+pose: after createKeyFrame")
ctx.close()
