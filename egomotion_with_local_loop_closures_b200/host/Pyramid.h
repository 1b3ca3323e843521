// Pyramid.h -- host shim of `class Pyramid` (src/Pyramid.h:34-89), the matrix-form sibling of PixelWisePyramid: Jacobian at the
// WARPED pixel and transformed depth (src/Pyramid.cpp:99-130), weight of an out-of-bounds pixel not zeroed (:629-651).  In the
// reference only performPrecomputation() is ever called (src/ImageFunc.cpp:169-176; performIterationSteps has no caller); both
// are provided, on top of ellc_gn_iterate(ELLC_VARIANT_PYRAMID) -- always the bit-faithful STRICT arithmetic on the device.
//
// The N-long per-point arrays (N = prev_frame->no_nonZeroDepthPts, raster order of the selected pixels) that the device produces
// are mirrored: weights, residual, warpedImage, warpedPoints (2 x N; -1 where the warp left the image).  worldPoints,
// transformedWorldPoints, steepestDescent, warpedGradientx / y, saveImg stay empty (intermediate values that never leave the
// kernel's registers); covarianceMatrixInv / motionPrior belong to the dead motion-prior code (:741-774).
#pragma once

#include "DepthPropagation.h"
#include "Frame.h"

class Pyramid {
public:
    int level;
    float* pose;                 // caller-owned float[6] (the constructor takes it, src/Pyramid.cpp:13)
    float lastErr, error, pointUsage, weightedPose;
    float prevPose[6];
    float covarianceDiagonalWts[6];

    Mat steepestDescent, hessianInv, worldPoints, transformedWorldPoints, saveImg, warpedPoints, warpedImage, residual,
        warpedGradientx, warpedGradienty, weights, covarianceMatrixInv, motionPrior, deltapose;
    Mat hessian, sd_param;       // (locals of calculateHessianInv / updatePose in the reference; kept as members here)

    frame* prev_frame;
    frame* current_frame;
    depthMap* currentDepthMap;

    Pyramid(frame* prevframe, frame* currentframe, float* pose, depthMap* currDepthMap);
    void performPrecomputation();                         // :700-711: lastErr = sum w r^2 / N at the current pose
    float performIterationSteps();                        // :714-726: update, re-evaluate, returns error / lastErr
    void calculateHessianInv();                           // :153-207
    void updatePose();                                    // :528-553
    void putPreviousPose(frame* tminus1_prev_frame);
    float calResidualAndWeights();                        // :558-694: evaluates at *pose, returns sum w r^2 / N
    // stages of the reference that have no separate device step: they are all part of one evaluation
    void calculateSteepestDescent() {}
    void calculateWorldPoints() {}
    void calculateWarpedPoints() {}
    void calculateWarpedImage() {}
    void calCovarianceMatrixInv(float*) {}
    void calMotionPrior() {}

private:
    float evaluate(bool update);
};
