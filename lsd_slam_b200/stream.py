"""Sequential track + map loop over one image stream -- the part of SlamSystem that sits directly above the hot
path, in ``dataset_slam _hz:=0`` (blockUntilMapped) order:

    SlamSystem::trackFrame        SlamSystem.cpp:890-1040   new Frame -> importFrame if the KF depth changed
                                                            (:907-912) -> tracker->trackFrame (:932)
    SlamSystem::doMappingIteration SlamSystem.cpp:739-828   updateKeyframe (:794 -> :542-614), or on a keyframe
                                                            change finalizeKeyFrame (:400) + createKeyFrame (:473)

Keyframe SELECTION (SlamSystem.cpp:998-1020), the pose graph and loop closures are out of scope: a new keyframe
is forced every ``kf_every`` frames.  This driver only issues C-ABI calls; it is what bench.py
times and what the full-loop parity tests run.  ``fused_call=True`` issues the whole frame as ONE ABI call
(``lsdgpu_track_and_map``, same call sequence inside the library); ``False`` issues the individual calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

IDENT = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


class GpuStream:
    def __init__(self, ctx: abi.Context, mode: int = 1, kf_every: int = 20, fused_call: bool = True):
        self.ctx = ctx
        self.tracker = abi.SE3Tracker(ctx, mode=mode)
        self.map = abi.DepthMap(ctx)
        self.kf_every = kf_every
        self.fused_call = fused_call
        self.kf_id = None
        self.last_pose = IDENT.copy()
        self.n_tracked = 0
        self.prev_ids: list[int] = []
        self.poses: list[np.ndarray] = []
        self._res = abi.TrackResult()
        self._qts = np.zeros(8, np.float64)

    def init_gt(self, fid: int, image_u8: np.ndarray, depth: np.ndarray):
        """SlamSystem::gtDepthInit, SlamSystem.cpp:831-854"""
        self.ctx.upload(fid, image_u8)
        self.ctx.set_depth_gt(fid, depth)
        self.map.initializeFromGTDepth(fid)
        self.kf_id = fid
        self.last_pose = IDENT.copy()

    def _track_and_map_fused(self, fid, image_u8, stage_index, kf_change):
        ctx, trk = self.ctx, self.tracker
        img = None
        if image_u8 is not None:
            img = np.ascontiguousarray(image_u8, np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
        q = np.ascontiguousarray(self.last_pose, np.float64)
        r = self._res
        ctx._ck(ctx.L.lsdgpu_track_and_map(ctx.ptr, self.kf_id, fid, img, -1 if stage_index is None else stage_index,
                                           q.ctypes.data_as(C.POINTER(C.c_double)), C.byref(trk.settings), trk.mode,
                                           int(kf_change), C.byref(r), self._qts.ctypes.data_as(C.POINTER(C.c_double))))
        trk.last = r
        trk.pointUsage, trk.lastGoodCount, trk.lastBadCount = r.pointUsage, r.lastGoodCount, r.lastBadCount
        trk.lastMeanRes, trk.lastResidual = r.lastMeanRes, r.lastResidual
        trk.affineEstimation_a, trk.affineEstimation_b = r.affineEstimation_a, r.affineEstimation_b
        trk.diverged, trk.trackingWasGood = bool(r.diverged), bool(r.trackingWasGood)
        return np.array(r.frameToRef_qt, np.float64)

    def step(self, fid: int, image_u8: np.ndarray | None = None, stage_index: int | None = None) -> np.ndarray:
        """one frame: Frame construction (host image, or a pre-staged device image), trackFrame, mapping"""
        ctx = self.ctx
        kf_change = bool(self.kf_every) and (self.n_tracked + 1) % self.kf_every == 0
        if self.fused_call:
            pose = self._track_and_map_fused(fid, image_u8, stage_index, kf_change)
        else:
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
        if kf_change:
            if not self.fused_call:
                self.map.finalizeKeyFrame()
                self.map.createKeyFrame(fid)
            old_kf, self.kf_id = self.kf_id, fid
            self.last_pose = IDENT.copy()
            for i in self.prev_ids + [old_kf]:
                ctx.release(i)
            self.prev_ids = []
        else:
            if not self.fused_call:
                self.map.updateKeyframe([fid])
                ctx.clear_good_mask(fid)             # SlamSystem.cpp:573
            self.last_pose = pose
            self.prev_ids.append(fid)
            while len(self.prev_ids) > 1:
                ctx.release(self.prev_ids.pop(0))
        self.poses.append(pose)
        return pose
