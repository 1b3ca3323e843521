// PixelWisePyramid.h -- host shim of `class PixelWisePyramid` (src/PixelWisePyramid.h:38-107): the same public members and
// methods; the per-iteration work runs on the B200 (one ellc_gn_iterate per calculatePixelWiseParallel() /
// calculatePixelWiseParallelInvCompositional()).  GetImagePoseEstimate does NOT go through this class (it runs the whole
// coarse-to-fine schedule in one launch); the class exists for callers that drive the iterations themselves, as
// src/ImageFunc.cpp:163-253 does.
//
// Members the device does not materialise: steepestDescent / weightedSteepestDescent (6 x N scratch of the inverse-compositional
// variant: its device form is the per-keyframe record list built by ellc_prepare_keyframes_lc), saveImg, the per-thread partial
// sums (hessian_thread*, sd_param_thread*), test_img, covarianceMatrixInv / motionPrior (dead code in the reference,
// FLAG_USE_MOTION_PRIOR = false).  They are declared so that code naming them compiles, and stay empty.
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
    float covarianceDiagonalWts[6];

    Mat covarianceMatrixInv, motionPrior;                 // unused (motion prior off)
    Mat hessianInv;              // 6x6, hessian.inv() of the last iteration (:451, :939)
    Mat saveImg;
    Mat deltapose;               // 1x6
    Mat sd_param;                // 1x6
    Mat hessian;                 // 6x6
    Mat steepestDescent, weightedSteepestDescent;         // not materialised (see above)
    Mat savedWarpedPointsX, savedWarpedPointsY;           // -2: no depth, -1: warped out of the image (:217-218, :277-278)

    frame* prev_frame;
    frame* current_frame;
    depthMap* currentDepthMap;

    // for display (src/PixelWisePyramid.cpp:195-283, :332, :361): filled by every calculatePixelWiseParallel() while
    // fill_display is true (the reference always fills them; switch it off when only the normal equations are wanted --
    // the display planes force the bit-faithful STRICT arithmetic and six image read-backs per iteration)
    Mat display_warpedimg, display_templateimg, display_2bewarpedimg, display_iterationres, display_origres, display_weightimg;
    bool fill_display;
    bool want_weight_image;      // display_weightimg only (enough for saveWeights); implied by fill_display

    Mat hessian_thread1, hessian_thread2, hessian_thread3, sd_param_thread1, sd_param_thread2, sd_param_thread3, test_img;

    float residualSum;           // sum w r^2 of the last evaluation (the definition of src/Pyramid.cpp:682)

    PixelWisePyramid(frame* prevframe, frame* currentframe, float* pose, depthMap* currDepthMap);
    void putPreviousPose(frame* tminus1_prev_frame);
    void updatePose();                                    // :460-491 from the members hessian / sd_param (hessianInv is refreshed)
    void saveWeights(bool useAverageWeights = false);     // :500-552
    void calculatePixelWiseParallel();                    // :416-455
    void calculatePixelWiseParallelInvCompositional(int iter);   // :917-974
    ~PixelWisePyramid();
};
