// Frame.h -- host shim of the reference's `class frame` (src/Frame.h:32-397), restricted to the members the tracking
// path touches.  Same member names and meanings; the work behind constructImagePyramids / calculateGradient /
// calculateNonZeroDepthPts / updationOnPyrChange is done by the B200 library and read back on demand.
#pragma once

#include <vector>

#include "ExternVariable.h"
#include "Mat.h"

using ellc_host::Mat;
using ellc_host::uchar;

class frame {
public:
    frame();
    // replaces frame(VideoCapture): the caller hands over the grayscale, undistorted, resized level-0 image
    // (src/Frame.cpp:45-75 is host I/O and stays with the caller)
    frame(const unsigned char* gray, int width, int height);
    ~frame();

    int frameId;
    int parentKeyframeId;
    bool isKeyframe;
    int numWeightsAdded[util::MAX_PYRAMID_LEVEL];
    int width, height;

    Mat image;                                            // black and white image
    Mat image_pyramid[util::MAX_PYRAMID_LEVEL];
    Mat depth;
    Mat weight_pyramid[util::MAX_PYRAMID_LEVEL];
    Mat depth_pyramid[util::MAX_PYRAMID_LEVEL];
    Mat gradientx, gradienty;
    Mat mask;                                             // 255: non-zero depth, 0: zero depth
    int currentRows, currentCols;
    int pyrLevel;
    int no_nonZeroDepthPts;

    float poseWrtOrigin[6];                               // wrt KF
    float poseWrtWorld[6];                                // wrt first frame
    float rescaleFactor;

    void constructImagePyramids();                        // src/Frame.cpp:170-182   (GPU: pyrdown_u8_kernel)
    void calculateGradient();                             // src/Frame.cpp:185-285   (GPU: pack_tex_kernel)
    void calculateNonZeroDepthPts();                      // src/Frame.cpp:295-301   (GPU: select_count_kernel)
    void updationOnPyrChange(int level, bool isPrevious = true);   // src/Frame.cpp:316-327
    void initializePose();
    void calculatePoseWrtOrigin(frame* prev_image, float* poseChangeWrtPrevframe, bool frmhomo = false);   // :329-348
    void calculatePoseWrtWorld(frame* prev_image, float* poseChangeWrtPrevframe, bool frmhomo = false);    // :352-372
    void concatenateRelativePose(float* src_1wrt2, float* src_2wrt3, float* dest_1wrt3);                   // :503-530
    void concatenateOriginPose(float* src_1wrt0, float* src_2wrt0, float* dest_1wrt2);                     // :534-562
    void finaliseWeights();                               // src/Frame.cpp:678-695 (+ loop-closure records on the GPU)

    // residency in the B200 library (not part of the reference surface)
    int gpu_frame_slot, gpu_kf_slot;
    unsigned long long gpu_kf_stamp;                      // bumped by depthMap::markDepthUpdated()
    bool gpu_lc_ready;                                    // weights finalised and loop-closure records built
    static int numberOfInstances;
};
