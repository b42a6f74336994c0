"""Sequential track + map loop over one image stream -- the part of SlamSystem that sits directly above the hot
path, in ``dataset_slam _hz:=0`` (blockUntilMapped) order:

    SlamSystem::trackFrame        SlamSystem.cpp:890-1040   new Frame -> importFrame if the KF depth changed
                                                            (:907-912) -> tracker->trackFrame (:932)
    SlamSystem::doMappingIteration SlamSystem.cpp:739-828   updateKeyframe (:794 -> :542-614), or on a keyframe
                                                            change finalizeKeyFrame (:400) + createKeyFrame (:473)

Keyframe SELECTION (SlamSystem.cpp:998-1020), the pose graph and loop closures are out of scope: a new keyframe
is forced every ``kf_every`` frames.  This driver only issues C-ABI calls; it is what bench.py
times and what the full-loop parity tests run.
"""
from __future__ import annotations

import numpy as np

from . import abi

IDENT = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


class GpuStream:
    def __init__(self, ctx: abi.Context, mode: int = 1, kf_every: int = 20):
        self.ctx = ctx
        self.tracker = abi.SE3Tracker(ctx, mode=mode)
        self.map = abi.DepthMap(ctx)
        self.kf_every = kf_every
        self.kf_id = None
        self.last_pose = IDENT.copy()
        self.n_tracked = 0
        self.prev_ids: list[int] = []
        self.poses: list[np.ndarray] = []

    def init_gt(self, fid: int, image_u8: np.ndarray, depth: np.ndarray):
        """SlamSystem::gtDepthInit, SlamSystem.cpp:831-854"""
        self.ctx.upload(fid, image_u8)
        self.ctx.set_depth_gt(fid, depth)
        self.map.initializeFromGTDepth(fid)
        self.kf_id = fid
        self.last_pose = IDENT.copy()

    def step(self, fid: int, image_u8: np.ndarray | None = None, stage_index: int | None = None) -> np.ndarray:
        """one frame: Frame construction (host image, or a pre-staged device image), trackFrame, mapping"""
        ctx = self.ctx
        if stage_index is not None:
            ctx.frame_from_stage(fid, stage_index)
        else:
            ctx.upload(fid, image_u8)
        if ctx.depth_updated_flag(self.kf_id):
            self.tracker.importFrame(self.kf_id)
        pose = self.tracker.trackFrame(self.kf_id, fid, self.last_pose)
        self.n_tracked += 1
        if self.tracker.diverged:
            raise RuntimeError(f"tracking diverged on frame {fid}")
        if self.kf_every and self.n_tracked % self.kf_every == 0:
            self.map.finalizeKeyFrame()
            self.map.createKeyFrame(fid)
            old_kf, self.kf_id = self.kf_id, fid
            self.last_pose = IDENT.copy()
            for i in self.prev_ids + [old_kf]:
                ctx.release(i)
            self.prev_ids = []
        else:
            self.map.updateKeyframe([fid])
            self.last_pose = pose
            ctx.clear_good_mask(fid)                 # SlamSystem.cpp:573
            self.prev_ids.append(fid)
            while len(self.prev_ids) > 1:
                ctx.release(self.prev_ids.pop(0))
        self.poses.append(pose)
        return pose
