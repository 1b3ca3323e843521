// DepthPropagation.h -- host shim of the slice of `class depthMap` (src/DepthPropagation.h:40-78) the tracker reads:
// the per-level depth / variance arrays handed over by the depth module (producer is out of scope).
#pragma once

#include <vector>

#include "ExternVariable.h"
#include "Frame.h"

class depthMap {
public:
    depthMap();
    frame* keyFrame;
    frame* currentFrame;
    float* deptharrptr[util::MAX_PYRAMID_LEVEL];          // stride ORIG_COLS >> level; level 0 uses -1 for invalid
    float* depthvararrptr[util::MAX_PYRAMID_LEVEL];       // -1 = invalid
    // Tell the shim that the keyframe's depth / variance pyramids changed (after updateDepthImage in the reference):
    // the next GetImagePoseEstimate re-uploads and re-selects.
    void markDepthUpdated() { ++stamp; }
    unsigned long long stamp;

private:
    std::vector<float> depth_store_[util::MAX_PYRAMID_LEVEL], var_store_[util::MAX_PYRAMID_LEVEL];
};
