// PixelWisePyramid.h -- host shim of `class PixelWisePyramid` (src/PixelWisePyramid.h:38-107): same public members, the
// per-iteration work runs on the B200 (one ellc_gn_evaluate + ellc_solve_update per calculatePixelWiseParallel()).
// GetImagePoseEstimate does NOT go through this class (it runs the whole coarse-to-fine schedule in one launch); the
// class exists for callers that drive iterations themselves, as src/ImageFunc.cpp:163-253 does.
#pragma once

#include "DepthPropagation.h"
#include "Frame.h"

class PixelWisePyramid {
public:
    int pyrlevel;
    int nRows, nCols;
    float* pose;                 // caller-owned float[6], assigned after construction (src/ImageFunc.cpp:183)
    float weightedPose;
    float prevPose[6];
    Mat hessianInv;              // not materialised by the GPU path (left empty)
    Mat deltapose;               // 1x6
    Mat sd_param;                // 1x6
    Mat hessian;                 // 6x6
    Mat display_weightimg;       // filled when FLAG_DISPLAY_IMAGES-style consumers ask for it (want_weight_image)
    bool want_weight_image;
    float residualSum;           // sum w r^2 of the last evaluation (Pyramid.cpp:682 definition)
    frame* prev_frame;
    frame* current_frame;
    depthMap* currentDepthMap;

    PixelWisePyramid(frame* prevframe, frame* currentframe, float* pose, depthMap* currDepthMap);
    void putPreviousPose(frame* tminus1_prev_frame);
    void updatePose();
    void calculatePixelWiseParallel();
    ~PixelWisePyramid();
};
