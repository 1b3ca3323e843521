// ImageFunc.h -- the reference's tracker entry point (src/ImageFunc.h:31), same signature.
#pragma once

#include <vector>

#include "DepthPropagation.h"
#include "Frame.h"

std::vector<float> GetImagePoseEstimate(frame* prev_frame, frame* current_frame, int frame_num, depthMap* currDepthMap,
                                        frame* tminus1_prev_frame, float* initial_pose_estimate,
                                        bool fromLoopClosure = false, bool homo = false);

namespace ellc_host {
// One B200 context (ellc_handle: own CUDA streams, own slot pools) per calling host thread, created lazily from
// util::configure()'s values -- the main thread and the loop-closure thread (src/GlobalOptimize.cpp:241, :566-568) track
// concurrently without sharing any mutable state.  shutdown() destroys all of them (call it when no tracker call is in flight).
void shutdown();
// Batched form for the loop-closure thread: n independent (keyframe, frame) pairs in one launch
// (what src/GlobalOptimize.cpp:480-610 does one call at a time).  init_poses: n x 6; returns n x 6 relative poses.
std::vector<float> TrackPairsBatched(const std::vector<frame*>& keyframes, const std::vector<depthMap*>& depthMaps,
                                     const std::vector<frame*>& frames, const std::vector<float>& init_poses);
const char* last_error();                 // of the calling thread
int context_count();                      // contexts (one per calling host thread) currently holding a device handle
}  // namespace ellc_host
