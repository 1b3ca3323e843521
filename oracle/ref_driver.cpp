// ref_driver.cpp -- builds the reference's OWN tracking sources into oracle/_ref/libellc_ref.so (TEST INFRASTRUCTURE ONLY).
//
// The reference (IITD-COMPUTER-VISION-GROUP/Egomotion_with_Local_Loop_Closures) needs OpenCV 3.0, Eigen 3.2.5 and Boost 1.59,
// none of which exist in this image, and it ships no tests or golden vectors.  To pin the oracle restatement on the
// reference's own code anyway, this translation unit #includes the UNMODIFIED reference sources where they lie
// (-I/root/reference/src; nothing is copied into this repository) and compiles them against the stand-in headers under
// oracle/shim/ (a minimal cv::Mat / Eigen::Matrix / boost::thread_group; the third-party arithmetic in them is our
// restatement of the published algorithms, pinned separately against cv2 / scipy).  What runs here is therefore the
// reference's per-pixel code (src/PixelWisePyramid.cpp), its samplers (src/Frame.h), pyramid / gradient / mask / pose
// bookkeeping (src/Frame.cpp), intrinsics (src/UserDefinedFunc.cpp) and its driver GetImagePoseEstimate (src/ImageFunc.cpp).
// The reference hard-codes its image size and intrinsics (src/ExternVariable.h:39-59: 480x270, fx = 410.6, fy = 409.0);
// ellc_ref_dims reports them and the tests run the oracle and the CUDA path at exactly that configuration.
//
// One translation unit on purpose: util::FLAG_DO_UNDISTORTION is a `static bool` defined in a header (one copy per
// translation unit, src/ExternVariable.h:61); SURVEY 8d runs without lens distortion, and only code in the same unit as
// Frame.cpp can switch that copy off.
#include "Frame.cpp"
#include "EigenInitialization.cpp"
#include "UserDefinedFunc.cpp"
#include "PixelWisePyramid.cpp"
#include "Pyramid.cpp"
#include "ImageFunc.cpp"
#include "DepthPropagation.cpp"       // updateDepthImage / buildInvVarDepth / mapDepthArr2Mat / calculate_no_of_Seeds
// (the gating helpers are private members of globalOptimize; the keyword is redefined for this one header only -- every
// standard header it pulls in has already been included above -- so that the driver can call them without touching the source)
#define private public
#include "GlobalOptimize.cpp"         // calculateImageHistogram / compareImageHistogram / calculateRotationStats
#undef private

// ---- definitions that live in src/main.cpp (:34-60), which is not part of the tracking path --------------------------------
int util::MAX_ITER[] = {4, 7, 9, 12};
bool util::FLAG_ALTERNATE_GN_RA = false;
int util::FLAG_IS_BOOTSTRAP = false;
int util::BATCH_START_ID = 0;
int util::BATCH_SIZE = 0;
int util::NUM_GN_PROPAGATION = 0;
int util::NUM_RA_PROPAGATION = 0;
bool util::GAUSS_NEWTON_FLAG_ON = false;
bool util::ROTATION_AVERAGING_FLAG_ON = false;
bool util::FLAG_DO_LOOP_CLOSURE = false;
bool util::FLAG_REPLICATE_NEW_DEPTH = false;
bool util::FLAG_INITIALIZE_NONZERO_POSE = false;
bool util::FLAG_SAVE_MATS = false;
bool util::FLAG_DO_PARALLEL_SHORT_LOOP_CLOSURE = false;
bool util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION = false;
bool util::FLAG_DO_PARALLEL_CONST_WEIGHT_POSE_EST = false;
bool util::FLAG_USE_LOOP_CLOSURE_TRIGGER = false;
bool util::EXIT_CONDITION = false;

// ---- src/DisplayFunc.cpp (highgui windows): not on the path, FLAG_DISPLAY_IMAGES is false ------------------------------------
FILE* STATE_FILE = nullptr;
void MY_ASSERT_FUNC(bool) {}
void DEBUG_COND_FUNC(bool) {}
void DisplayIterationRes(frame*, Mat, String, int, bool) {}
void DisplayInitialRes(frame*, frame*, String, int, bool) {}
void DisplayWarpedImg(Mat, frame*, String, int, bool) {}
void DisplayWeights(frame*, Mat, String, int, bool) {}
void DisplayOriginalImg(Mat, frame*, String, int, bool) {}
void DisplayColouredDepth(Mat, Mat) {}
void DisplayIterationResPixelWise(Mat, frame*, String, int, bool) {}
void DisplayInitialResPixelWise(Mat, frame*, String, int, bool) {}
void DisplayWarpedImgPxelWise(Mat, frame*, String, int, bool) {}
void DisplayWeightsPixelWise(Mat, frame*, String, int, bool) {}
void DisplayOriginalImgPixelWise(Mat, frame*, String, int, bool) {}

namespace {

const int W0 = util::ORIG_COLS, H0 = util::ORIG_ROWS;

// frame::frame(VideoCapture) (src/Frame.cpp:34-119) resizes the camera image by RESIZE_FACTOR = 1/4 and converts BGR -> gray:
// feed it the 4x pixel-replicated, 3-channel version of the wanted image, which both steps return exactly.
frame* make_frame(const uint8_t* gray) {
    const int F = (int)util::DIM_FACTOR;
    Mat big(H0 * F, W0 * F, CV_8UC3);
    for (int y = 0; y < H0 * F; ++y) {
        uchar* row = big.ptr<uchar>(y);
        for (int x = 0; x < W0 * F; ++x) {
            const uchar v = gray[(y / F) * W0 + (x / F)];
            row[3 * x] = row[3 * x + 1] = row[3 * x + 2] = v;
        }
    }
    VideoCapture cap(&big);
    return new frame(cap);
}

// depthMap is 10 MB of fixed arrays with a constructor in src/DepthPropagation.cpp (the depth module, out of scope).  The
// tracker only reads depthvararrptr[level] (src/PixelWisePyramid.cpp:348), so an all-zero object with those pointers set is
// everything it needs; no member function of depthMap is ever called.
struct DepthHolder {
    depthMap* dm;
    std::vector<float> var[4];
    DepthHolder() : dm((depthMap*)std::calloc(1, sizeof(depthMap))) {}
    ~DepthHolder() { std::free(dm); }
};

void set_keyframe_depth(frame* kf, DepthHolder& dh, const float* const* depth, const float* const* var) {
    for (int l = 0; l < 4; ++l) {
        const int r = H0 >> l, c = W0 >> l;
        kf->depth_pyramid[l] = Mat::zeros(r, c, CV_32FC1);
        for (int y = 0; y < r; ++y) std::memcpy(kf->depth_pyramid[l].ptr<float>(y), depth[l] + (size_t)y * c, (size_t)c * sizeof(float));
        dh.var[l].assign(var[l], var[l] + (size_t)r * c);
        dh.dm->depthvararrptr[l] = dh.var[l].data();
    }
    kf->depth = kf->depth_pyramid[0];
}

void init_once() {
    util::FLAG_DO_UNDISTORTION = false;       // no lens distortion in the synthetic configurations
}

}  // namespace

extern "C" {

struct ellc_ref_iter {
    float H[36];
    float b[6];
    float weighted_pose;
    float pose_after[6];
    float pad_;
};
struct ellc_ref_trace {
    int n_selected[4];
    int n_iters[4];
    ellc_ref_iter it[4][12];
    float final_pose[6];
};

// the reference's compile-time configuration (src/ExternVariable.h:39-62, :76, :148-149)
void ellc_ref_dims(int* w, int* h, float* fx, float* fy, float* cx, float* cy) {
    *w = W0; *h = H0; *fx = util::ORIG_FX; *fy = util::ORIG_FY; *cx = util::ORIG_CX; *cy = util::ORIG_CY;
}

// frame construction + updationOnPyrChange(level): image_pyramid[level] (pyrDown dims), gradientx / gradienty, and -- when a
// depth level is given -- mask and no_nonZeroDepthPts (src/Frame.cpp:170-327).  Any output may be NULL.
int ellc_ref_frame_level(const uint8_t* gray, int level, const float* depth_level, uint8_t* pyr_image, int* pyr_w, int* pyr_h,
                         float* gradx, float* grady, uint8_t* mask, int* count) {
    init_once();
    frame* f = make_frame(gray);
    const int r = H0 >> level, c = W0 >> level;
    if (depth_level) {
        for (int y = 0; y < r; ++y) std::memcpy(f->depth_pyramid[level].ptr<float>(y), depth_level + (size_t)y * c, (size_t)c * sizeof(float));
    }
    f->updationOnPyrChange(level, depth_level != nullptr);
    const Mat& im = f->image_pyramid[level];
    if (pyr_w) *pyr_w = im.cols;
    if (pyr_h) *pyr_h = im.rows;
    if (pyr_image) for (int y = 0; y < im.rows; ++y) std::memcpy(pyr_image + (size_t)y * im.cols, im.ptr<uchar>(y), (size_t)im.cols);
    if (gradx) for (int y = 0; y < r; ++y) std::memcpy(gradx + (size_t)y * c, f->gradientx.ptr<float>(y), (size_t)c * sizeof(float));
    if (grady) for (int y = 0; y < r; ++y) std::memcpy(grady + (size_t)y * c, f->gradienty.ptr<float>(y), (size_t)c * sizeof(float));
    if (depth_level && mask) for (int y = 0; y < r; ++y) std::memcpy(mask + (size_t)y * c, f->mask.ptr<uchar>(y), (size_t)c);
    if (depth_level && count) *count = f->no_nonZeroDepthPts;
    delete f;
    return 0;
}

// frame::getInterpolatedElement(x, y, int) on the level image and (x, y, "gradx" / "grady") on the gradients (src/Frame.h:181-394)
int ellc_ref_interpolate(const uint8_t* gray, int level, int n, const float* xs, const float* ys, float* intensity, float* gx, float* gy) {
    init_once();
    frame* f = make_frame(gray);
    f->updationOnPyrChange(level, false);
    for (int i = 0; i < n; ++i) {
        intensity[i] = f->getInterpolatedElement(xs[i], ys[i], 1);
        gx[i] = f->getInterpolatedElement(xs[i], ys[i], "gradx");
        gy[i] = f->getInterpolatedElement(xs[i], ys[i], "grady");
    }
    delete f;
    return 0;
}

// frame::concatenateRelativePose / concatenateOriginPose (src/Frame.cpp:503-562)
void ellc_ref_concat_relative(const float a[6], const float b[6], float dest[6]) {
    frame f;
    float x[6], y[6];
    std::memcpy(x, a, sizeof(x)); std::memcpy(y, b, sizeof(y));
    f.concatenateRelativePose(x, y, dest);
}
void ellc_ref_concat_origin(const float a[6], const float b[6], float dest[6]) {
    frame f;
    float x[6], y[6];
    std::memcpy(x, a, sizeof(x)); std::memcpy(y, b, sizeof(y));
    f.concatenateOriginPose(x, y, dest);
}

// The reference's driver, untouched: GetImagePoseEstimate(keyframe, frame, ..., t-1 frame, ...) (src/ImageFunc.cpp:49-315).
// The keyframe sits at the world origin; the t-1 frame carries tminus1_pose_wrt_world, from which the driver derives its
// initial pose (:97-108).  Returns the relative pose and the frame's poseWrtOrigin / poseWrtWorld post-conditions (:305-307).
int ellc_ref_get_image_pose_estimate(const uint8_t* kf_gray, const uint8_t* cur_gray, const float* const* depth, const float* const* var,
                                     const float tminus1_pose_wrt_world[6], float pose_out[6], float pose_wrt_origin[6],
                                     float pose_wrt_world[6]) {
    init_once();
    frame* kf = make_frame(kf_gray);
    frame* cur = make_frame(cur_gray);
    frame* tm1 = make_frame(cur_gray);
    DepthHolder dh;
    set_keyframe_depth(kf, dh, depth, var);
    for (int i = 0; i < 6; ++i) tm1->poseWrtWorld[i] = tminus1_pose_wrt_world[i];
    float unused[6] = {0, 0, 0, 0, 0, 0};
    std::vector<float> p = GetImagePoseEstimate(kf, cur, 1, dh.dm, tm1, unused, false, false);
    for (int i = 0; i < 6; ++i) { pose_out[i] = p[i]; pose_wrt_origin[i] = cur->poseWrtOrigin[i]; pose_wrt_world[i] = cur->poseWrtWorld[i]; }
    delete kf; delete cur; delete tm1;
    return 0;
}

// The same level / iteration schedule (src/ImageFunc.cpp:150-299) driven from here, so that the per-iteration state of the
// reference's PixelWisePyramid object can be recorded: hessian, sd_param, weightedPose, pose (public members,
// src/PixelWisePyramid.h:38-107).  weight_l0 (may be NULL): display_weightimg of the last executed level-0 iteration (:361).
int ellc_ref_track_trace(const uint8_t* kf_gray, const uint8_t* cur_gray, const float* const* depth, const float* const* var,
                         const float init_pose[6], ellc_ref_trace* out, float* weight_l0) {
    init_once();
    std::memset(out, 0, sizeof(*out));
    frame* kf = make_frame(kf_gray);
    frame* cur = make_frame(cur_gray);
    DepthHolder dh;
    set_keyframe_depth(kf, dh, depth, var);
    float pose[6];
    for (int i = 0; i < 6; ++i) pose[i] = init_pose[i];
    for (int level = util::MAX_PYRAMID_LEVEL - 1; level >= 0; --level) {
        kf->updationOnPyrChange(level);
        cur->updationOnPyrChange(level, false);
        PixelWisePyramid wp(kf, cur, pose, dh.dm);
        wp.putPreviousPose(cur);
        wp.pose = pose;
        out->n_selected[level] = kf->no_nonZeroDepthPts;
        for (int iter = 0; iter < util::MAX_ITER[level]; ++iter) {
            wp.calculatePixelWiseParallel();
            ellc_ref_iter& r = out->it[level][iter];
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) r.H[i * 6 + j] = wp.hessian.at<float>(i, j);
            for (int i = 0; i < 6; ++i) { r.b[i] = wp.sd_param.at<float>(0, i); r.pose_after[i] = pose[i]; }
            r.weighted_pose = wp.weightedPose;
            out->n_iters[level] = iter + 1;
            if (level == 0 && weight_l0)
                for (int y = 0; y < H0; ++y) std::memcpy(weight_l0 + (size_t)y * W0, wp.display_weightimg.ptr<float>(y), (size_t)W0 * sizeof(float));
            if (wp.weightedPose < 1.0f) break;                     // src/ImageFunc.cpp:251-252
        }
    }
    for (int i = 0; i < 6; ++i) out->final_pose[i] = pose[i];
    delete kf; delete cur;
    return 0;
}

// The reference's constant-weight loop-closure flow, end to end, with its own driver:
//   1. FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION on: every sequential GetImagePoseEstimate() saves the weights of the last iteration
//      of each level into the keyframe (saveWeights(true), src/ImageFunc.cpp:280-288, src/PixelWisePyramid.cpp:500-552);
//   2. frame::finaliseWeights() when the keyframe is retired (src/main.cpp:431-434, src/Frame.cpp:678-695);
//   3. GetImagePoseEstimate(..., fromLoopClosure = true): calculatePixelWiseParallelInvCompositional (:917-974).
// parallel selects FLAG_DO_PARALLEL_CONST_WEIGHT_POSE_EST (3 precompute bands + 2 iteration bands vs one).
// Outputs: the finalised weight pyramid + counts, the sequential poses, the loop-closure pose from the driver, and the trace of
// the same loop-closure track stepped from here (hessian, sd_param, weightedPose, pose per iteration).
int ellc_ref_lc_flow(const uint8_t* kf_gray, int n_seq, const uint8_t* const* seq_gray, const float* seq_tminus1, const float* const* depth,
                     const float* const* var, const uint8_t* lc_gray, const float lc_tminus1[6], int parallel, float* const* weight_out,
                     int counts[4], float* seq_poses, float lc_pose[6], ellc_ref_trace* lc_trace) {
    init_once();
    util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION = true;
    util::FLAG_DO_PARALLEL_CONST_WEIGHT_POSE_EST = parallel != 0;
    frame* kf = make_frame(kf_gray);
    DepthHolder dh;
    set_keyframe_depth(kf, dh, depth, var);
    float unused[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_seq; ++i) {
        frame* cur = make_frame(seq_gray[i]);
        frame* tm1 = make_frame(seq_gray[i]);
        for (int k = 0; k < 6; ++k) tm1->poseWrtWorld[k] = seq_tminus1[i * 6 + k];
        std::vector<float> p = GetImagePoseEstimate(kf, cur, i + 1, dh.dm, tm1, unused, false, false);
        for (int k = 0; k < 6; ++k) seq_poses[i * 6 + k] = p[k];
        delete cur; delete tm1;
    }
    kf->finaliseWeights();
    for (int l = 0; l < 4; ++l) {
        const int r = H0 >> l, c = W0 >> l;
        counts[l] = kf->numWeightsAdded[l];
        for (int y = 0; y < r; ++y) std::memcpy(weight_out[l] + (size_t)y * c, kf->weight_pyramid[l].ptr<float>(y), (size_t)c * sizeof(float));
    }
    {
        frame* cur = make_frame(lc_gray);
        frame* tm1 = make_frame(lc_gray);
        for (int k = 0; k < 6; ++k) tm1->poseWrtWorld[k] = lc_tminus1[k];
        std::vector<float> p = GetImagePoseEstimate(kf, cur, n_seq + 1, dh.dm, tm1, unused, true, false);
        for (int k = 0; k < 6; ++k) lc_pose[k] = p[k];
        delete cur; delete tm1;
    }
    if (lc_trace) {
        std::memset(lc_trace, 0, sizeof(*lc_trace));
        frame* cur = make_frame(lc_gray);
        float pose[6] = {0, 0, 0, 0, 0, 0}, zero[6] = {0, 0, 0, 0, 0, 0}, tm1w[6];
        for (int k = 0; k < 6; ++k) tm1w[k] = lc_tminus1[k];
        kf->concatenateOriginPose(tm1w, zero, pose);                            // src/ImageFunc.cpp:106
        for (int level = util::MAX_PYRAMID_LEVEL - 1; level >= 0; --level) {
            kf->updationOnPyrChange(level);
            cur->updationOnPyrChange(level, false);
            PixelWisePyramid wp(kf, cur, pose, dh.dm);
            wp.putPreviousPose(cur);
            wp.pose = pose;
            lc_trace->n_selected[level] = kf->no_nonZeroDepthPts;
            for (int iter = 0; iter < util::MAX_ITER[level]; ++iter) {
                wp.calculatePixelWiseParallelInvCompositional(iter);
                ellc_ref_iter& r = lc_trace->it[level][iter];
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) r.H[i * 6 + j] = wp.hessian.at<float>(i, j);
                for (int i = 0; i < 6; ++i) { r.b[i] = wp.sd_param.at<float>(0, i); r.pose_after[i] = pose[i]; }
                r.weighted_pose = wp.weightedPose;
                lc_trace->n_iters[level] = iter + 1;
                if (wp.weightedPose < 1.0f) break;
            }
        }
        for (int i = 0; i < 6; ++i) lc_trace->final_pose[i] = pose[i];
        delete cur;
    }
    delete kf;
    util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION = false;
    util::FLAG_DO_PARALLEL_CONST_WEIGHT_POSE_EST = false;
    return 0;
}

// depthMap::updateDepthImage (src/DepthPropagation.cpp:1254-1315: level-0 depth from the hypotheses with the 3-pixel border rule,
// then buildInvVarDepth :1637-1719 and mapDepthArr2Mat :1722-1746) and calculate_no_of_Seeds (:1804-1830, evaluated on the
// input flags).  Outputs per level: keyFrame->depth_pyramid (Mat convention, 0 = invalid), deptharrptr / depthvararrptr
// (array convention, -1 = invalid); valid_out: isValid after the call.
int ellc_ref_update_depth_image(const uint8_t* kf_gray, const uint8_t* valid, const float* inv_depth_smoothed, const float* variance_smoothed,
                                float* const* depth_mat, float* const* depth_arr, float* const* var_arr, uint8_t* valid_out, float* occupancy) {
    init_once();
    depthMap* dm = new depthMap();
    frame* kf = make_frame(kf_gray);
    dm->keyFrame = kf;
    dm->currentFrame = kf;
    const int n = W0 * H0;
    for (int i = 0; i < n; ++i) {
        dm->currentDepthHypothesis[i].isValid = valid[i] != 0;
        dm->currentDepthHypothesis[i].invDepthSmoothed = inv_depth_smoothed[i];
        dm->currentDepthHypothesis[i].varianceSmoothed = variance_smoothed[i];
    }
    *occupancy = dm->calculate_no_of_Seeds(true);
    dm->updateDepthImage(false);
    for (int i = 0; i < n; ++i) valid_out[i] = dm->currentDepthHypothesis[i].isValid ? 1 : 0;
    for (int l = 0; l < 4; ++l) {
        const int r = H0 >> l, c = W0 >> l;
        for (int y = 0; y < r; ++y) std::memcpy(depth_mat[l] + (size_t)y * c, kf->depth_pyramid[l].ptr<float>(y), (size_t)c * sizeof(float));
        std::memcpy(depth_arr[l], dm->deptharrptr[l], (size_t)r * c * sizeof(float));
        std::memcpy(var_arr[l], dm->depthvararrptr[l], (size_t)r * c * sizeof(float));
    }
    delete kf;
    delete dm;
    return 0;
}

// globalOptimize::calculateImageHistogram (src/GlobalOptimize.cpp:40-100) and compareImageHistogram (:116-122) on two images,
// calculateRotationStats (:419-452) on two poses.
int ellc_ref_gating(const uint8_t* gray_a, const uint8_t* gray_b, const float pose_a[6], const float pose_b[6], float hist_a[256],
                    float hist_b[256], double* kl_ab, float* rms_error, float* relative_view_angle) {
    init_once();
    globalOptimize* go = new globalOptimize("/dev/null");
    frame* fa = make_frame(gray_a);
    frame* fb = make_frame(gray_b);
    go->calculateImageHistogram(fa);
    Mat ha = go->currentLoopFrame.image_histogram.clone();
    go->calculateImageHistogram(fb);
    Mat hb = go->currentLoopFrame.image_histogram.clone();
    for (int i = 0; i < 256; ++i) { hist_a[i] = ha.at<float>(i); hist_b[i] = hb.at<float>(i); }
    *kl_ab = go->compareImageHistogram(ha, hb);
    float pa[6], pb[6];
    std::memcpy(pa, pose_a, sizeof(pa)); std::memcpy(pb, pose_b, sizeof(pb));
    go->calculateRotationStats(pa, pb);
    *rms_error = go->rms_error;
    *relative_view_angle = go->relative_view_angle;
    delete fa; delete fb; delete go;
    return 0;
}

// The matrix-form tracker of src/Pyramid.cpp.  The reference constructs it at every level but never iterates it
// (performIterationSteps is not called from src/ImageFunc.cpp:192-226); here it is driven the way its own comments describe:
// performPrecomputation() at the given pose (:700-711), then `iters` x performIterationSteps() (:714-726).
// Outputs: n = selected pixels; weights / residual (1 x n, selection order) and last_err = sum w r^2 / n at the INPUT pose (:682);
// hessian_inv of the first iteration; pose after each iteration; error ratios returned by performIterationSteps.
int ellc_ref_pyramid_run(const uint8_t* kf_gray, const uint8_t* cur_gray, const float* const* depth, const float* const* var, int level,
                         const float pose_in[6], int iters, int* n_sel, float* weights, float* residual, float* last_err,
                         float hessian_inv[36], float* poses_after, float* ratios) {
    init_once();
    frame* kf = make_frame(kf_gray);
    frame* cur = make_frame(cur_gray);
    DepthHolder dh;
    set_keyframe_depth(kf, dh, depth, var);
    kf->updationOnPyrChange(level);
    cur->updationOnPyrChange(level, false);
    float pose[6];
    for (int i = 0; i < 6; ++i) pose[i] = pose_in[i];
    Pyramid wp(kf, cur, pose, dh.dm);
    wp.putPreviousPose(cur);
    wp.pose = pose;
    wp.performPrecomputation();
    const int n = kf->no_nonZeroDepthPts;
    *n_sel = n;
    *last_err = wp.lastErr;
    for (int i = 0; i < n; ++i) { weights[i] = wp.weights.ptr<float>(0)[i]; residual[i] = wp.residual.ptr<float>(0)[i]; }
    for (int it = 0; it < iters; ++it) {
        ratios[it] = wp.performIterationSteps();
        if (it == 0)
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) hessian_inv[i * 6 + j] = wp.hessianInv.at<float>(i, j);
        for (int i = 0; i < 6; ++i) poses_after[it * 6 + i] = pose[i];
    }
    delete kf; delete cur;
    return 0;
}

}  // extern "C"
