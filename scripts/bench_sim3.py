"""Latency / throughput of the batched Sim3 tracker vs the CPU oracle, 640x480, levels 4..1.
A batch is built from the ordered pairs of three keyframes (both directions, as SlamSystem::tryTrackSim3 runs them)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from lsd_slam_b200 import abi, synth
from oracle import pyoracle as po

W, H = 640, 480
seq = synth.Sequence(W, H, seed=1234)
ctx = abi.Context(W, H, seq.K, max_frames=8)
trk = abi.Sim3Tracker(ctx)
po.set_globals()
po.set_globals(fast=True)
of, off = {}, {}
for k in (0, 6, 12):
    img, z = seq.render(k)
    ctx.upload(k, img); ctx.set_depth_gt(k, z)
    for d, fast in ((of, False), (off, True)):
        f = po.Frame(k, img, seq.K, fast=fast); f.setDepthFromGroundTruth(z); d[k] = f
pairs = [(a, b) for a in of for b in of if a != b]


def init(ref, fr):
    q = np.concatenate([seq.frame_to_ref_qt(fr, ref), [1.02]])
    q[4:7] += [0.01, -0.005, 0.004]
    return q


out = {}
for n in (1, 2, 6, 18, 36, 148, 296):
    ps = [pairs[i % len(pairs)] for i in range(n)]
    refs, frs = [p[0] for p in ps], [p[1] for p in ps]
    inits = np.array([init(*p) for p in ps])
    for _ in range(20 if n == 1 else 3):          # the first batches after process start see the clock ramp
        res = trk.trackFrameSim3Batch(refs, frs, inits, 4, 1)
    reps = 10 if n <= 36 else 4
    ctx.timer_begin(1)
    for _ in range(reps):
        res = trk.trackFrameSim3Batch(refs, frs, inits, 4, 1)
    ctx.timer_end(1)
    ms = ctx.timer_ms(1) / reps
    evals = sum(sum(r.numCalcResidualCalls) for r in res)
    out[n] = {"ms_per_batch": ms, "trackings_per_s": n / (ms * 1e-3), "evaluations": evals, "us_per_tracking": 1e3 * ms / n}
    if n == 1:
        for cs in (1, 2, 4, 8):
            os.environ["LSDGPU_SIM3_CLUSTER"] = str(cs)
            for _ in range(3):
                trk.trackFrameSim3Batch(refs, frs, inits, 4, 1)
            ctx.timer_begin(2)
            for _ in range(reps):
                trk.trackFrameSim3Batch(refs, frs, inits, 4, 1)
            ctx.timer_end(2)
            out[n][f"ms_cluster{cs}"] = ctx.timer_ms(2) / reps
        del os.environ["LSDGPU_SIM3_CLUSTER"]
# CPU, one thread (the constraint-search thread of the reference runs trackFrameSim3 single-threaded): strict-IEEE scalar
# flavour, -O3 scalar, and -O3 with the reference's SSE variants of calcSim3WeightsAndResidual / calcSim3LGS
# (Sim3Tracker.cpp:611-736, 860-973; calcSim3Buffers is scalar in the reference as well) = the reference's default build
cpu = {}
for name, d, fast, sse in (("parity_scalar", of, False, 0), ("O3_scalar", off, True, 0), ("O3_sse", off, True, 1)):
    po.set_globals(fast=fast, useSSE=sse)
    t0 = time.perf_counter(); m = 0
    while time.perf_counter() - t0 < 3.0:
        for ref, fr in pairs:
            po.sim3_track(d[ref], d[fr], init(ref, fr), 4, 1); m += 1
    cpu[name] = (time.perf_counter() - t0) / m * 1e3
    po.set_globals(fast=fast)
print(json.dumps({"gpu": out, "cpu_oracle_ms_per_tracking_1thread": cpu}))
