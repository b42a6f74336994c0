import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lsd_slam_b200 import abi, synth
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_adapter import _sim3_mul, _sim3_inv
seq = synth.Sequence(320, 240, seed=1234)
frames = [seq.render(k) for k in range(12)]
ctx = abi.Context(seq.w, seq.h, seq.K, max_frames=12)
trk, dm = abi.SE3Tracker(ctx, mode=1), abi.DepthMap(ctx)
ctx.upload(0, frames[0][0]); ctx.set_depth_gt(0, frames[0][1]); dm.initializeFromGTDepth(0)
ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
kf, prev, last = 0, None, ident
c2w = {0: np.array([0, 0, 0, 1, 0, 0, 0, 1.0])}
for k in range(1, 8):
    ctx.upload(k, frames[k][0])
    if ctx.depth_updated_flag(kf): trk.importFrame(kf)
    pose = np.array(trk.trackFrame(kf, k, last))
    c2w[k] = _sim3_mul(c2w[kf], np.concatenate([pose, [1.0]]))
    if k % 5 == 0:
        dm.finalizeKeyFrame(); q = dm.createKeyFrame(k)
        c2w[k] = _sim3_mul(c2w[kf], q); kf, last = k, ident
        before = dm.current().copy()
        r2k = _sim3_mul(_sim3_inv(c2w[kf]), c2w[prev])
        print("refToKf", r2k)
        dm.updateKeyframe([(prev, r2k)])
        after = dm.current().copy()
        v = (before["isValid"] != 0) & (after["isValid"] != 0)
        print("late map: valid", v.sum(), "idepth changed", (before["idepth"][v] != after["idepth"][v]).sum(), "validity changed", (before["isValid"] != after["isValid"]).sum(),
              "flag", ctx.depth_updated_flag(kf))
        prev = None
    else:
        dm.updateKeyframe([k]); ctx.clear_good_mask(k); last, prev = pose, k
    print(k, pose)
