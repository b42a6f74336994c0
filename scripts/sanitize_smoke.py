"""Short run of every kernel for compute-sanitizer (memcheck / racecheck): 160x112 stream, keyframe change included."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from lsd_slam_b200 import abi, synth
from lsd_slam_b200.stream import GpuStream
seq = synth.Sequence(160, 112, seed=7)
fr = [seq.render(k) for k in range(8)]
for mode in (1, 0):
    ctx = abi.Context(160, 112, seq.K, max_frames=8)
    gs = GpuStream(ctx, mode=mode, kf_every=4, fused_call=(mode == 1))
    gs.init_gt(0, fr[0][0], fr[0][1])
    for k in range(1, 8):
        gs.step(k, fr[k][0])
    gs.map.current(); gs.map.integral()
    ctx.download(gs.kf_id, abi.BUF_GOODMASK); ctx.depth_stats(gs.kf_id)
    ctx.close()
print("sanitize run ok")
