"""Worker for the world_size-2 tests (launched by torch.distributed.run, gloo backend).

  cpu : host-side multi-rank logic only (shard partition, all-reduce callback, max-over-ranks timing reduce)
  gpu : point-sharded SE3 tracking of one frame, both ranks on cuda:0, sums all-reduced over gloo
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from lsd_slam_b200 import sharded
    out = {"rank": rank, "world": world}
    if mode == "cpu":
        from lsd_slam_b200 import abi
        n = 320 * 240 // 4
        mine = sharded.shard_chunks(n, rank, world)
        cnt = torch.zeros(n, dtype=torch.int32)
        cnt[mine] = 1
        dist.all_reduce(cnt)
        out["partition_ok"] = bool((cnt == 1).all())
        cb = sharded.make_allreduce(dist)
        buf = (np.arange(abi.EVAL_NSUMS, dtype=np.float32) + 100 * rank)
        import ctypes as C
        cb(None, buf.ctypes.data_as(C.POINTER(C.c_float)), abi.EVAL_NSUMS)
        expect = sum(np.arange(abi.EVAL_NSUMS, dtype=np.float32) + 100 * r for r in range(world))
        out["allreduce_ok"] = bool(np.array_equal(buf, expect))
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["max_ok"] = float(t.item()) == 10.0 + world - 1
    else:
        from lsd_slam_b200 import abi, synth
        seq = synth.Sequence(320, 240, seed=1234)
        img0, d0 = seq.render(0)
        img3, _ = seq.render(3)
        ctx = abi.Context(seq.w, seq.h, seq.K, device=0, max_frames=4)
        ctx.upload(0, img0)
        ctx.set_depth_gt(0, d0)
        ctx.upload(3, img3)
        trk = sharded.ShardedSE3Tracker(ctx, rank, world, sharded.make_allreduce(dist))
        pose = trk.trackFrame(0, 3, np.array([0, 0, 0, 1, 0, 0, 0], np.float64))
        out["pose"] = pose.tolist()
        out["residual"] = trk.lastResidual
        out["calls"] = list(trk.last.numCalcResidualCalls)
        ctx.close()
    print("RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
