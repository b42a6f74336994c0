"""LSDGPU_SIM3_DEBUG=1 python scripts/dbg_sim3.py : per-phase cycle totals of one Sim3 tracking per cluster size"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from lsd_slam_b200 import abi, synth
seq = synth.Sequence(640, 480, seed=1234)
ctx = abi.Context(640, 480, seq.K, max_frames=4)
trk = abi.Sim3Tracker(ctx)
for k in (0, 6):
    img, z = seq.render(k)
    ctx.upload(k, img); ctx.set_depth_gt(k, z)
init = np.concatenate([seq.frame_to_ref_qt(6, 0), [1.02]]); init[4:7] += [0.01, -0.005, 0.004]
for cs in (1, 8):
    os.environ["LSDGPU_SIM3_CLUSTER"] = str(cs)
    for _ in range(2):
        trk.trackFrameSim3(0, 6, init, 4, 1)
