// track_persistent.cuh -- device-resident LM loop of SE3Tracker::trackFrame (mode 1).  Placeholder: the
// first milestone routes mode 1 through the host-driven loop; replaced by the persistent kernel next.
#pragma once
#include "internal.cuh"
#include "track.cuh"

struct TrackState {
    float dummy[64];
};
static cudaError_t trackPersistentSetup(lsdgpu_ctx*) { return cudaSuccess; }
static int trackHostLM(lsdgpu_ctx* ctx, FrameSlot* kf, FrameSlot* fr, const double init_qt[7],
                       const lsdgpu_track_settings* st, lsdgpu_track_result* out);
static int trackPersistent(lsdgpu_ctx* ctx, FrameSlot* kf, FrameSlot* fr, const double init_qt[7],
                           const lsdgpu_track_settings* st, lsdgpu_track_result* out)
{
    return trackHostLM(ctx, kf, fr, init_qt, st, out);
}
