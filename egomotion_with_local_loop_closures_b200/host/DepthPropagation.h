// DepthPropagation.h -- host shim of the slice of `class depthMap` (src/DepthPropagation.h:40-78) the tracker reads:
// the per-level depth / variance arrays handed over by the depth module (producer is out of scope).
#pragma once

#include <vector>

#include "ExternVariable.h"
#include "Frame.h"

// The members of the reference's depthhypothesis (src/DepthHypothesis.h) that updateDepthImage / calculate_no_of_Seeds read.
struct depthhypothesis {
    bool isValid;
    float invDepthSmoothed;
    float varianceSmoothed;
    depthhypothesis() : isValid(false), invDepthSmoothed(-1.0f), varianceSmoothed(-1.0f) {}
};

class depthMap {
public:
    depthMap();
    depthhypothesis* currentDepthHypothesis;              // ORIG_COLS x ORIG_ROWS, filled by the caller's depth module
    // src/DepthPropagation.cpp:1254-1315 (+ buildInvVarDepth :1637-1719, mapDepthArr2Mat :1721-1746) on the B200: uploads the
    // hypotheses, builds the keyframe's depth / variance pyramids there, and mirrors them (and the border-invalidated isValid
    // flags) back into keyFrame->depth, keyFrame->depth_pyramid[], deptharrptr[], depthvararrptr[].
    void updateDepthImage(bool fromKeyFrameCreation = false);
    float calculate_no_of_Seeds(bool calculate_on_current = true);      // :1804-1830
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
    std::vector<depthhypothesis> hyp_store_;
};
