"""Single stream, SE3 tracking point-sharded across ranks (SURVEY 8e / BASELINE config 5).

Every rank holds the same keyframe and frame; rank r evaluates every world-th 32-pixel chunk of each level and
the 40 partial sums of every evaluation (21 A + 6 b + 13 statistics) are all-reduced (NCCL on GPUs, gloo for
the CPU-side tests) before the LM decision, so all ranks stay in lock step.  The exchange is 160 bytes per
evaluation: latency-bound, expected to be SLOWER than one GPU (reported as such in DESIGN.md).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


def shard_chunks(n_pixels: int, rank: int, world: int) -> np.ndarray:
    """pixel indices owned by `rank` (32-pixel chunks dealt round-robin) -- mirrors the kernel's rule"""
    i = np.arange(n_pixels)
    return i[((i >> 5) % world) == rank]


def make_allreduce(dist, device=None):
    """ctypes callback summing the LSDGPU_EVAL_NSUMS floats over all ranks with torch.distributed"""
    import torch

    def _cb(_user, ptr, n):
        a = np.ctypeslib.as_array(ptr, shape=(n,))
        t = torch.from_numpy(a.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        a[:] = t.cpu().numpy()

    return abi.ALLREDUCE_FN(_cb)


class ShardedSE3Tracker(abi.SE3Tracker):
    def __init__(self, ctx: abi.Context, rank: int, world: int, allreduce_cb):
        super().__init__(ctx, mode=0)
        self.rank, self.world = rank, world
        self._cb = allreduce_cb            # keep the ctypes callback alive

    def trackFrame(self, kf_id: int, frame_id: int, frameToReference_initialEstimate) -> np.ndarray:
        q = np.ascontiguousarray(frameToReference_initialEstimate, np.float64)
        r = abi.TrackResult()
        cb = C.cast(self._cb, C.c_void_p) if self._cb is not None else None
        self.ctx._ck(self.ctx.L.lsdgpu_se3_track_sharded(self.ctx.ptr, kf_id, frame_id, q.ctypes.data_as(C.POINTER(C.c_double)),
                                                         C.byref(self.settings), self.rank, self.world, cb, None, C.byref(r)))
        self.last = r
        self.pointUsage, self.lastGoodCount, self.lastBadCount = r.pointUsage, r.lastGoodCount, r.lastBadCount
        self.lastMeanRes, self.lastResidual = r.lastMeanRes, r.lastResidual
        self.affineEstimation_a, self.affineEstimation_b = r.affineEstimation_a, r.affineEstimation_b
        self.diverged, self.trackingWasGood = bool(r.diverged), bool(r.trackingWasGood)
        return np.array(r.frameToRef_qt, np.float64)


def attach_peers(ctx: abi.Context, dist, group=None):
    """Device-side exchange (include/lsdgpu.h, "one stream over several GPUs"): gather the CUDA IPC handles of all ranks' arenas,
    map them, and switch this context's device-resident tracker to sharded operation.  Collective: every rank calls it."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.peer_export(), group=group)
    ctx.peer_attach(rank, handles)
    dist.barrier(group)


def detach_peers(ctx: abi.Context, dist, group=None):
    dist.barrier(group)
    ctx.peer_detach()
    dist.barrier(group)
