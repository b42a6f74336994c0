import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["LSDGPU_TRACK_DEBUG"]="1"
from lsd_slam_b200 import abi, synth
from lsd_slam_b200.stream import GpuStream
seq=synth.Sequence(640,480)
fr=[seq.render(k) for k in range(8)]
ctx=abi.Context(640,480,seq.K,max_frames=8)
gs=GpuStream(ctx,mode=1,kf_every=20)
gs.init_gt(0,fr[0][0],fr[0][1])
for k in range(1,8):
    gs.step(k,fr[k][0])
    print(list(gs.tracker.last.numCalcResidualCalls))
