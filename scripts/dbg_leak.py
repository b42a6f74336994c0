"""does an idle, never-closed context disturb mode-1 tracking on other contexts? (r2n flake hunt)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pytest
from lsd_slam_b200 import abi, synth

seq = synth.Sequence(320, 240, seed=1234)
frames = [seq.render(k) for k in range(3)]
leaked = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    ctx = abi.Context(seq.w, seq.h, seq.K, device=0, max_frames=20)
    ctx.upload(0, frames[0][0]); ctx.set_depth_gt(0, frames[0][1])
    dm = abi.DepthMap(ctx); dm.initializeFromGTDepth(0)
    ctx.upload(1, frames[1][0])
    trk = abi.SE3Tracker(ctx, mode=1)
    trk.trackFrame(0, 1, seq.frame_to_ref_qt(0))
    leaked.append((ctx, dm, trk))
sys.exit(pytest.main(["-q", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_track.py"), "-x"]))
