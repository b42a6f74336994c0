"""Throughput of the batched permaRef path vs the CPU oracle, 640x480 (level 4 = 40x30)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from lsd_slam_b200 import abi, synth
from oracle import pyoracle as po

seq = synth.Sequence(640, 480, seed=1234)
ctx = abi.Context(640, 480, seq.K, max_frames=8)
trk = abi.SE3Tracker(ctx)
po.set_globals()
prs = {}
for k in (0, 6, 12):
    img, d = seq.render(k)
    ctx.upload(k, img); ctx.set_depth_gt(k, d); trk.setPermaRef(k)
    okf = po.Frame(k, img, seq.K); okf.setDepthFromGroundTruth(d); prs[k] = (po.PermaRef(okf), okf)
img, _ = seq.render(9)
ctx.upload(9, img)
of = po.Frame(9, img, seq.K)
ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
out = {}
for n in (1, 64, 296, 1024, 4096):
    ids = [(0, 6, 12)[i % 3] for i in range(n)]
    inits = np.tile(ident, (n, 1))
    for _ in range(3):
        trk.trackFrameOnPermaref(ids, 9, inits)
    ctx.timer_begin(1)
    reps = 5
    for _ in range(reps):
        res = trk.trackFrameOnPermaref(ids, 9, inits)
    ctx.timer_end(1)
    ms = ctx.timer_ms(1) / reps
    evals = sum(r.numCalcResidualCalls[4] for r in res)
    out[n] = {"ms_per_batch": ms, "candidates_per_s": n / (ms * 1e-3), "evaluations": evals, "us_per_candidate": 1e3 * ms / n}
    for _ in range(3):
        trk.checkPermaRefOverlap(ids, inits)
    ctx.timer_begin(2)
    for _ in range(reps):
        trk.checkPermaRefOverlap(ids, inits)
    ctx.timer_end(2)
    out[n]["overlap_ms_per_batch"] = ctx.timer_ms(2) / reps
t0 = time.perf_counter(); m = 0
while time.perf_counter() - t0 < 3.0:
    for k in (0, 6, 12):
        prs[k][0].track(of, ident); m += 1
cpu = (time.perf_counter() - t0) / m
print(json.dumps({"gpu": out, "cpu_oracle_ms_per_candidate": cpu * 1e3, "cpu_candidates_per_s_1thread": 1 / cpu}))
