"""BASELINE config 5 without the host: ONE 640x480 stream whose SE3 tracking is sharded over N GPUs INSIDE the device-resident
tracking kernel -- the per-pass sums and the level-1 tracking mask cross NVLink through peer-mapped memory (csrc/track_persistent.cuh,
gridExchangeRanks).  Every rank runs the same track + map loop on the same frames (mapping is replicated, bit-identical).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_peer_sharded.py [--frames 40]
Rank 0 prints ONE JSON line: per-frame times of the sharded loop (device events, max over ranks) next to the same loop on one GPU,
and the checks: all ranks identical bit for bit (poses, maps, masks), poses within 1e-5 of the single-GPU run."""
import argparse, hashlib, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsd_slam_b200 import abi, sharded, synth
from lsd_slam_b200.stream import GpuStream

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=40)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--kf-every", type=int, default=20)
ap.add_argument("--same-gpu", action="store_true", help="functional check on ONE GPU: all ranks on cuda:0 (time-sliced), gloo rendezvous")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if args.same_gpu:
    local = 0
torch.cuda.set_device(local)
if args.same_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DEV = "cpu" if args.same_gpu else "cuda"
seq = synth.Sequence(args.width, args.height, seed=1234)
frames = [seq.render(k) for k in range(args.frames + 1)]


def run(attach):
    ctx = abi.Context(args.width, args.height, seq.K, device=local, max_frames=8)
    if attach:
        sharded.attach_peers(ctx, dist)
    gs = GpuStream(ctx, mode=1, kf_every=args.kf_every, fused_call=True)
    gs.init_gt(0, frames[0][0], frames[0][1])
    ms = []
    for k in range(1, args.frames + 1):
        dist.barrier()
        torch.cuda.synchronize()
        ctx.timer_begin(0)
        gs.step(k, frames[k][0])
        ctx.timer_end(0)
        ms.append(ctx.timer_ms(0))
    poses = np.array(gs.poses)
    cur = gs.map.current().copy()
    h = hashlib.sha256(cur.tobytes()).hexdigest()
    # tracking alone: the last frame against the current keyframe, 30 times
    trk = abi.SE3Tracker(ctx, mode=1)
    tms = []
    for _ in range(33):
        dist.barrier()
        torch.cuda.synchronize()
        ctx.timer_begin(1)
        trk.trackFrame(gs.kf_id, args.frames, gs.last_pose)
        ctx.timer_end(1)
        tms.append(ctx.timer_ms(1))
    out = dict(ms=np.array(ms), poses=poses, map_hash=h, n_valid=int((cur["isValid"] != 0).sum()), track_ms=np.array(tms[3:]))
    if attach:
        sharded.detach_peers(ctx, dist)
    ctx.close()
    return out


single = run(False)
dist.barrier()
multi = run(True)
t = torch.tensor(multi["ms"], dtype=torch.float64, device=DEV)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms_multi = t.cpu().numpy()
t = torch.tensor(multi["track_ms"], dtype=torch.float64, device=DEV)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
track_multi = t.cpu().numpy()
# identical on every rank?
blob = [None] * world
dist.all_gather_object(blob, (multi["poses"].tobytes(), multi["map_hash"]))
if rank == 0:
    same = all(b == blob[0] for b in blob)
    dp = np.abs(multi["poses"] - single["poses"])
    tr = np.linalg.norm(multi["poses"][:, 4:7] - single["poses"][:, 4:7], axis=1) / np.maximum(np.linalg.norm(single["poses"][:, 4:7], axis=1), 1e-12)
    warm = 5
    out = {"config": f"single {args.width}x{args.height} stream, track+map, tracking sharded {world}-way inside the kernel (peer-memory exchange over NVLink)",
           "n_gpus": world, "frames": args.frames, "kf_every": args.kf_every,
           "sharded_ms_per_step": float(np.median(ms_multi[warm:])), "single_gpu_ms_per_step": float(np.median(single["ms"][warm:])),
           "sharded_trackFrame_ms": float(np.median(track_multi)), "single_gpu_trackFrame_ms": float(np.median(single["track_ms"])),
           "sharded_mean_ms": float(np.mean(ms_multi[warm:])), "single_gpu_mean_ms": float(np.mean(single["ms"][warm:])),
           "all_ranks_bit_identical": bool(same), "max_pose_rel_vs_single_gpu": float(tr.max()), "max_pose_abs_diff": float(dp.max()),
           "maps_equal_single_gpu": bool(multi["map_hash"] == single["map_hash"]), "n_valid": multi["n_valid"],
           "timing": "CUDA events on the context's stream around each step (lsdgpu_timer_*), max over ranks, median over frames"}
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
