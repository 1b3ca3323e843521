// ellc_common.cuh -- shared device/host types of the B200 GN tracker (internal; the public ABI is include/ellc_gn.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ellc_gn.h"

namespace ellc {

constexpr int kLevels = ELLC_LEVELS;

// One selected keyframe pixel (mask != 0, src/Frame.cpp:298), produced once per keyframe level by the selection
// kernels and streamed by every GN iteration as one coalesced 16-byte load.
struct __align__(16) SelRec {
    uint32_t xy;        // x | (y << 16)
    float    depth;     // prev_frame->depth_pyramid[level](y,x)        (src/PixelWisePyramid.cpp:192)
    float    var;       // currentDepthMap->depthvararrptr[level][idx]  (src/PixelWisePyramid.cpp:348)
    uint32_t ikf;       // prev_frame->image_pyramid[level](y,x)        (src/PixelWisePyramid.cpp:189), low byte
};
static_assert(sizeof(SelRec) == 16, "SelRec must be 16 bytes");

// Packed texel of the current frame at one pyramid level: everything one bilinear tap needs in one 32-bit word.
//   bits  0.. 7  intensity I (unsigned)
//   bits  8..19  2*gradx, 12-bit two's complement  (gradx of src/Frame.cpp:185-285 is a multiple of 0.5 in [-255, 255])
//   bits 20..31  2*grady, 12-bit two's complement
// Interpolating the doubled integer gradients and halving the result is bit-identical to interpolating the
// reference's f32 gradient maps (scaling by 2 is exact in binary floating point).  The all-zero word is an
// out-of-bounds tap (pixVal = 0, src/Frame.h:211-215).
__host__ __device__ inline uint32_t tex_pack(int I, int gx2, int gy2) {
    return (uint32_t)I | (((uint32_t)gx2 & 0xfffu) << 8) | (((uint32_t)gy2 & 0xfffu) << 20);
}
__host__ __device__ inline int tex_I(uint32_t w) { return (int)(w & 0xffu); }
__host__ __device__ inline int tex_gx2(uint32_t w) { return ((int)(w << 12)) >> 20; }
__host__ __device__ inline int tex_gy2(uint32_t w) { return ((int)w) >> 20; }

// Geometry of the pyramid for one configuration.
struct Geometry {
    int width, height;
    int pyr_w[kLevels], pyr_h[kLevels];        // cv::pyrDown dims: ceil-halving (image_pyramid[l])
    int cols[kLevels], rows[kLevels];          // currentCols/currentRows = width>>l, height>>l (src/Frame.cpp:321-322)
    int64_t img_off[kLevels + 1];              // element offsets of the u8 pyramid levels inside a slot
    int64_t win_off[kLevels + 1];              // element offsets of (cols x rows) level arrays (depth/var/mask/tex/recs)
};

// Per-level intrinsics, GetIntrinsic (src/UserDefinedFunc.cpp:33-49)
struct LevelK {
    float fx, fy, cx, cy;
    float ifx, ify;            // fast-mode reciprocals
    float fy_ifx, fx_ify;      // fy/fx, fx/fy
};

struct TrackParams {
    Geometry geo;
    LevelK K[kLevels];
    int max_iter[kLevels];
    float huber_half;          // HUBER_D / 2
    float noise2;              // CAMERA_PIXEL_NOISE_2
    float weight[6];
    float stop_threshold;
    int jacobian_at_warped;
    // pools
    const uint32_t* tex_pool;  int64_t tex_slot_stride;       // packed texels, per frame slot
    const SelRec* rec_pool;    int64_t rec_slot_stride;       // selection records, per keyframe slot
    const int* count_pool;                                    // [kf_slot][kLevels]
    // work
    const ellc_pair* pairs;
    ellc_result* results;
    ellc_iter_trace* trace;    // optional [pair][level][ELLC_MAX_TRACE_ITERS]
    int n_pairs;
    // evaluate-only mode (ellc_gn_evaluate): run `level_hi..level_lo`, `iter_limit` iterations, no pose update
    int level_hi, level_lo;
    int iter_limit;            // 0 = use max_iter
    int no_update;
    float* weight_out;         // optional display_weightimg of the evaluated level (cols x rows), evaluate-only mode
};

}  // namespace ellc
