// ImageFunc.h -- the reference's tracker entry point (src/ImageFunc.h:31), same signature.
#pragma once

#include <vector>

#include "DepthPropagation.h"
#include "Frame.h"

std::vector<float> GetImagePoseEstimate(frame* prev_frame, frame* current_frame, int frame_num, depthMap* currDepthMap,
                                        frame* tminus1_prev_frame, float* initial_pose_estimate,
                                        bool fromLoopClosure = false, bool homo = false);

namespace ellc_host {
// One B200 context per process (created lazily from util::configure()'s values); closes at exit.
void shutdown();
// Batched form for the loop-closure thread: n independent (keyframe, frame) pairs in one launch
// (what src/GlobalOptimize.cpp:480-610 does one call at a time).  init_poses: n x 6; returns n x 6 relative poses.
std::vector<float> TrackPairsBatched(const std::vector<frame*>& keyframes, const std::vector<depthMap*>& depthMaps,
                                     const std::vector<frame*>& frames, const std::vector<float>& init_poses);
const char* last_error();
}  // namespace ellc_host
