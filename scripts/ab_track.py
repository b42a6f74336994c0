"""A/B of tracker builds on one box, one process per library: median device time of trackFrame + the debug cycle table.
usage: python scripts/ab_track.py <path to liblsdgpu.so> [frames]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from lsd_slam_b200 import abi
abi.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from lsd_slam_b200 import synth
from lsd_slam_b200.stream import GpuStream, IDENT
seq = synth.Sequence(640, 480)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fr = [seq.render(k) for k in range(N)]
ctx = abi.Context(640, 480, seq.K, max_frames=8)
ctx.stage_reserve(N)
for k in range(N):
    ctx.stage_put(k, fr[k][0])
gs = GpuStream(ctx, mode=1, kf_every=0, fused_call=False)
gs.init_gt(0, fr[0][0], fr[0][1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
T = {k: [] for k in ("build", "track", "map", "total")}
last = IDENT
for k in range(1, N):
    flush.fill_(k & 255); torch.cuda.synchronize()
    ctx.timer_begin(0); ctx.timer_begin(1)
    ctx.frame_from_stage(k, k)
    ctx.timer_end(1)
    if ctx.depth_updated_flag(0):
        gs.tracker.importFrame(0)
    ctx.timer_begin(3)
    last = gs.tracker.trackFrame(0, k, last)
    ctx.timer_end(3); ctx.timer_begin(4)
    gs.map.updateKeyframe([k]); ctx.clear_good_mask(k)
    ctx.timer_end(4); ctx.timer_end(0)
    for name, slot in (("build", 1), ("track", 3), ("map", 4), ("total", 0)):
        T[name].append(ctx.timer_ms(slot) * 1e3)
    if k > 1:
        ctx.release(k - 1)
print(os.path.basename(sys.argv[1]), " ".join(f"{n} {np.median(v[5:]):.1f}us" for n, v in T.items()), "pose", np.round(last[4:], 6))
