"""BASELINE config 5: ONE 640x480 stream, SE3 tracking point-sharded over N GPUs, the 40 sums of every evaluation
all-reduced with NCCL (160 bytes, latency only).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py
Prints one JSON line on rank 0: ms per trackFrame sharded (max over ranks) next to the single-GPU host-driven (mode 0) and
device-resident (mode 1) trackers on rank 0."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.getcwd())
os.environ.setdefault("NCCL_DEBUG", "WARN")
from lsd_slam_b200 import abi, sharded, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
seq = synth.Sequence(640, 480, seed=1234)
ctx = abi.Context(640, 480, seq.K, device=local, max_frames=6)
img0, d0 = seq.render(0)
ctx.upload(0, img0); ctx.set_depth_gt(0, d0)
frames = [3, 4, 5]
for k in frames:
    ctx.upload(k, seq.render(k)[0])
ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def timed(trk, reps):
    for k in frames:
        trk.trackFrame(0, k, ident)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in frames:
            pose = trk.trackFrame(0, k, ident)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / (reps * len(frames))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), pose, sum(trk.last.numCalcResidualCalls)


sh_ms, sh_pose, evals = timed(sharded.ShardedSE3Tracker(ctx, rank, world, sharded.make_allreduce(dist, device=f"cuda:{local}")), 10)
out = {"config": "single 640x480 stream, point-sharded SE3 tracking + NCCL all-reduce of 40 sums per evaluation",
       "n_gpus": world, "sharded_ms_per_trackFrame": sh_ms, "evaluations_per_frame": evals}
# single-GPU references on every rank (identical work; keeps the barriers of timed() aligned)
m0, p0, _ = timed(abi.SE3Tracker(ctx, mode=0), 10)
m1, p1, _ = timed(abi.SE3Tracker(ctx, mode=1), 10)
if rank == 0:
    out["single_gpu_mode0_ms"] = m0
    out["single_gpu_mode1_ms"] = m1
    out["pose_diff_vs_mode0"] = float(np.abs(sh_pose - p0).max())
    out["note"] = "wall clock around the blocking ABI call (the all-reduce callback runs on the host between kernels), max over ranks"
    print(json.dumps(out), flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
