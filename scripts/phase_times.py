"""Per-phase device time of one frame step (CUDA events on the context's stream), median over frames."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from lsd_slam_b200 import abi, synth
from lsd_slam_b200.stream import GpuStream, IDENT
seq = synth.Sequence(640, 480)
N = 40
fr = [seq.render(k) for k in range(N)]
ctx = abi.Context(640, 480, seq.K, max_frames=8)
ctx.stage_reserve(N)
for k in range(N):
    ctx.stage_put(k, fr[k][0])
gs = GpuStream(ctx, mode=1, kf_every=0, fused_call=False)
gs.init_gt(0, fr[0][0], fr[0][1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
T = {k: [] for k in ("build", "import", "track", "map", "total")}
last = IDENT
for k in range(1, N):
    flush.fill_(k & 255); torch.cuda.synchronize()
    ctx.timer_begin(0); ctx.timer_begin(1)
    ctx.frame_from_stage(k, k)
    ctx.timer_end(1); ctx.timer_begin(2)
    if ctx.depth_updated_flag(0):
        gs.tracker.importFrame(0)
    ctx.timer_end(2); ctx.timer_begin(3)
    last = gs.tracker.trackFrame(0, k, last)
    ctx.timer_end(3); ctx.timer_begin(4)
    gs.map.updateKeyframe([k]); ctx.clear_good_mask(k)
    ctx.timer_end(4); ctx.timer_end(0)
    for name, slot in (("build", 1), ("import", 2), ("track", 3), ("map", 4), ("total", 0)):
        T[name].append(ctx.timer_ms(slot) * 1e3)
    if k > 1:
        ctx.release(k - 1)
for n, v in T.items():
    print(f"{n:7s} median {np.median(v[5:]):7.1f} us   p90 {np.percentile(v[5:], 90):7.1f} us")
