// HostShim.cpp -- C++ host side of the drop-in: the reference's frame / depthMap / PixelWisePyramid /
// GetImagePoseEstimate call surface (same names, argument meaning and post-conditions) implemented on top of the C-ABI
// of include/ellc_gn.h.  No numerical work happens here: every image, gradient, mask, normal-equation and pose update
// is produced by the CUDA library; this file only moves buffers and keeps the reference's bookkeeping
// (src/ImageFunc.cpp:92-138 initial pose, :305-307 pose write-back, level-0 post-conditions).
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ellc_gn.h"
#include "DepthPropagation.h"
#include "ExternVariable.h"
#include "Frame.h"
#include "ImageFunc.h"
#include "PixelWisePyramid.h"

// ---- util:: definitions (src/main.cpp:34-60 defaults) ----------------------------------------------------------------
namespace util {
int ORIG_COLS = 480, ORIG_ROWS = 270;                                    // src/ExternVariable.h:50-51 defaults
float ORIG_FX = 1642.405612f / 4, ORIG_FY = 1636.148027f / 4, ORIG_CX = 240.0f, ORIG_CY = 135.0f;
int MAX_ITER[4] = {4, 7, 9, 12};
bool FLAG_DO_PARALLEL_POSE_ESTIMATION = true;
bool FLAG_INITIALIZE_NONZERO_POSE = false;
bool FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION = false;
bool FLAG_DO_LOOP_CLOSURE = false;
int BATCH_START_ID = 0;
void configure(int cols, int rows, float fx, float fy, float cx, float cy) {
    ORIG_COLS = cols; ORIG_ROWS = rows; ORIG_FX = fx; ORIG_FY = fy; ORIG_CX = cx; ORIG_CY = cy;
}
}  // namespace util

namespace {

const int kFrameSlots = 64, kKfSlots = 48;     // the reference keeps a ring of 43 keyframes (src/ExternVariable.h:161-162)

struct Context {
    ellc_handle* h = nullptr;
    std::mutex mu;                              // main thread + loop-closure thread share one context
    std::vector<frame*> frame_owner, kf_owner;
    std::vector<unsigned long long> kf_stamp;
    int next_frame = 0, next_kf = 0;
    std::string err;
    int cols = 0, rows = 0;
};
Context g;

void fail(const std::string& what) {
    g.err = what + ": " + (g.h ? ellc_last_error_string(g.h) : ellc_last_error_string(nullptr));
    throw std::runtime_error(g.err);
}

ellc_handle* ctx() {
    if (g.h && (g.cols != util::ORIG_COLS || g.rows != util::ORIG_ROWS)) { ellc_destroy(g.h); g.h = nullptr; }
    if (!g.h) {
        ellc_config c;
        ellc_default_config(&c, util::ORIG_COLS, util::ORIG_ROWS);
        c.fx = util::ORIG_FX; c.fy = util::ORIG_FY; c.cx = util::ORIG_CX; c.cy = util::ORIG_CY;
        for (int l = 0; l < 4; ++l) c.max_iter[l] = util::MAX_ITER[l];
        c.huber_d = util::HUBER_D; c.camera_pixel_noise_2 = util::CAMERA_PIXEL_NOISE_2;
        for (int i = 0; i < 6; ++i) c.weight[i] = util::weight[i];
        c.max_frames = kFrameSlots; c.max_keyframes = kKfSlots;
        if (ellc_create(&c, &g.h) != ELLC_OK) fail("ellc_create");
        g.frame_owner.assign(kFrameSlots, nullptr);
        g.kf_owner.assign(kKfSlots, nullptr);
        g.kf_stamp.assign(kKfSlots, 0);
        g.cols = util::ORIG_COLS; g.rows = util::ORIG_ROWS;
    }
    return g.h;
}

int frame_slot(frame* f) {
    ellc_handle* h = ctx();
    if (f->gpu_frame_slot >= 0 && g.frame_owner[f->gpu_frame_slot] == f) return f->gpu_frame_slot;
    const int s = g.next_frame;
    g.next_frame = (g.next_frame + 1) % kFrameSlots;
    if (g.frame_owner[s]) g.frame_owner[s]->gpu_frame_slot = -1;
    g.frame_owner[s] = f; f->gpu_frame_slot = s;
    if (ellc_upload_frame(h, s, f->image.ptr<uchar>(0)) != ELLC_OK) fail("ellc_upload_frame");
    return s;
}

int keyframe_slot(frame* f, depthMap* dm) {
    ellc_handle* h = ctx();
    const bool resident = f->gpu_kf_slot >= 0 && g.kf_owner[f->gpu_kf_slot] == f;
    if (resident && !dm) return f->gpu_kf_slot;                            // mask / count queries reuse what is there
    const unsigned long long stamp = dm ? dm->stamp : 0;
    if (resident && g.kf_stamp[f->gpu_kf_slot] == stamp) return f->gpu_kf_slot;
    int s = f->gpu_kf_slot;
    if (s < 0 || g.kf_owner[s] != f) {
        s = g.next_kf;
        g.next_kf = (g.next_kf + 1) % kKfSlots;
        if (g.kf_owner[s]) { g.kf_owner[s]->gpu_kf_slot = -1; g.kf_owner[s]->gpu_lc_ready = false; }
        g.kf_owner[s] = f; f->gpu_kf_slot = s;
        f->gpu_lc_ready = false;                                             // a fresh slot holds no weights
        if (ellc_reset_keyframe_weights(h, s) != ELLC_OK) fail("ellc_reset_keyframe_weights");
    }
    const float* dptr[4]; const float* vptr[4];
    std::vector<float> novar[4];
    for (int l = 0; l < 4; ++l) {
        dptr[l] = f->depth_pyramid[l].ptr<float>(0);                         // Mats: 0 = invalid at every level
        if (dm && dm->depthvararrptr[l]) vptr[l] = dm->depthvararrptr[l];
        else { novar[l].assign((size_t)(g.cols >> l) * (g.rows >> l), -1.0f); vptr[l] = novar[l].data(); }
    }
    if (ellc_upload_keyframe(h, s, f->image.ptr<uchar>(0), dptr, vptr) != ELLC_OK) fail("ellc_upload_keyframe");
    if (ellc_synchronize(h) != ELLC_OK) fail("ellc_synchronize");             // novar[] dies at scope exit
    g.kf_stamp[s] = stamp;
    if (f->gpu_lc_ready) {                                                   // new depth: the loop-closure records follow it
        const int32_t ks = s;
        if (ellc_prepare_keyframes_lc(h, 1, &ks) != ELLC_OK) fail("ellc_prepare_keyframes_lc");
    }
    return s;
}

void level_dims(int level, int& pw, int& ph, int& cols, int& rows) {
    if (ellc_level_dims(ctx(), level, &pw, &ph, &cols, &rows) != ELLC_OK) fail("ellc_level_dims");
}

}  // namespace

namespace ellc_host {
void shutdown() {
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.h) { ellc_destroy(g.h); g.h = nullptr; }
}
const char* last_error() { return g.err.c_str(); }
}  // namespace ellc_host

// ---- frame --------------------------------------------------------------------------------------------------------------
int frame::numberOfInstances = 0;

frame::frame() : frameId(0), parentKeyframeId(0), isKeyframe(false), width(0), height(0), currentRows(0), currentCols(0),
                 pyrLevel(0), no_nonZeroDepthPts(0), rescaleFactor(1.0f), gpu_frame_slot(-1), gpu_kf_slot(-1), gpu_kf_stamp(0), gpu_lc_ready(false) {
    for (int i = 0; i < 6; ++i) poseWrtOrigin[i] = poseWrtWorld[i] = 0.0f;
    for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l) numWeightsAdded[l] = 0;
}

frame::frame(const unsigned char* gray, int w, int h) : frame() {
    frameId = ++numberOfInstances;                                           // src/Frame.cpp:37
    width = w; height = h;
    image = Mat(h, w, ellc_host::CV_8UC1);
    std::memcpy(image.ptr<uchar>(0), gray, (size_t)w * h);
    pyrLevel = 0; currentCols = w; currentRows = h;                          // :81-84
    constructImagePyramids();                                                // :103
    calculateGradient();                                                     // :104
    depth = Mat::zeros(h, w, ellc_host::CV_32FC1);                            // :107-117
    for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l) {
        depth_pyramid[l] = Mat::zeros(h >> l, w >> l, ellc_host::CV_32FC1);
        weight_pyramid[l] = Mat::zeros(h >> l, w >> l, ellc_host::CV_32FC1);
    }
}

// frame::finaliseWeights, src/Frame.cpp:678-695 (called when the keyframe is retired, src/main.cpp:431-434): average the saved
// weights, then build the keyframe's loop-closure records so that later loop-closure pairs on it run the constant-weight tracker.
// weight_pyramid[] is read back so that callers looking at the member see what the reference would hold.
void frame::finaliseWeights() {
    std::lock_guard<std::mutex> lk(g.mu);
    if (gpu_kf_slot < 0) { std::printf("\nWeights cannot be averaged!!! "); return; }
    const int32_t ks = gpu_kf_slot;
    if (ellc_finalise_weights(ctx(), ks) != ELLC_OK) fail("ellc_finalise_weights");
    if (ellc_prepare_keyframes_lc(ctx(), 1, &ks) != ELLC_OK) fail("ellc_prepare_keyframes_lc");
    for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l)
        if (ellc_read_keyframe_weights(ctx(), ks, l, weight_pyramid[l].ptr<float>(0), nullptr) != ELLC_OK) fail("ellc_read_keyframe_weights");
    gpu_lc_ready = true;
}

frame::~frame() {
    std::lock_guard<std::mutex> lk(g.mu);
    if (gpu_frame_slot >= 0 && gpu_frame_slot < (int)g.frame_owner.size() && g.frame_owner[gpu_frame_slot] == this) g.frame_owner[gpu_frame_slot] = nullptr;
    if (gpu_kf_slot >= 0 && gpu_kf_slot < (int)g.kf_owner.size() && g.kf_owner[gpu_kf_slot] == this) g.kf_owner[gpu_kf_slot] = nullptr;
}

void frame::constructImagePyramids() {
    std::lock_guard<std::mutex> lk(g.mu);
    const int s = frame_slot(this);
    for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l) {
        int pw, ph, c, r;
        level_dims(l, pw, ph, c, r);
        image_pyramid[l] = Mat(ph, pw, ellc_host::CV_8UC1);
        if (ellc_read_frame_level(ctx(), s, l, image_pyramid[l].ptr<uchar>(0), nullptr, nullptr) != ELLC_OK) fail("ellc_read_frame_level");
    }
}

void frame::calculateGradient() {
    std::lock_guard<std::mutex> lk(g.mu);
    const int s = frame_slot(this);
    gradientx = Mat(currentRows, currentCols, ellc_host::CV_32FC1);
    gradienty = Mat(currentRows, currentCols, ellc_host::CV_32FC1);
    if (ellc_read_frame_level(ctx(), s, pyrLevel, nullptr, gradientx.ptr<float>(0), gradienty.ptr<float>(0)) != ELLC_OK) fail("ellc_read_frame_level");
}

void frame::calculateNonZeroDepthPts() {
    std::lock_guard<std::mutex> lk(g.mu);
    const int s = keyframe_slot(this, nullptr);
    int pw, ph, c, r;
    level_dims(pyrLevel, pw, ph, c, r);
    mask = Mat(r, c, ellc_host::CV_8UC1);
    int count = 0;
    if (ellc_read_keyframe_level(ctx(), s, pyrLevel, nullptr, mask.ptr<uchar>(0), &count) != ELLC_OK) fail("ellc_read_keyframe_level");
    no_nonZeroDepthPts = count;
}

void frame::updationOnPyrChange(int level, bool isPrevious) {
    pyrLevel = level;
    currentRows = height >> level;                                           // height / pow(2, level), src/Frame.cpp:321-322
    currentCols = width >> level;
    if (isPrevious) calculateNonZeroDepthPts();
    calculateGradient();
}

void frame::initializePose() { for (int i = 0; i < 6; ++i) poseWrtOrigin[i] = 0.0f; }

void frame::concatenateRelativePose(float* a, float* b, float* dest) { ellc_concat_relative(a, b, dest); }
void frame::concatenateOriginPose(float* a, float* b, float* dest) { ellc_concat_origin(a, b, dest); }

void frame::calculatePoseWrtOrigin(frame* prev_image, float* d, bool frmhomo) {
    if (!frmhomo) concatenateRelativePose(d, prev_image->poseWrtOrigin, poseWrtOrigin);
    else for (int i = 0; i < 3; ++i) poseWrtOrigin[i] = prev_image->poseWrtOrigin[i] + d[i];
}
void frame::calculatePoseWrtWorld(frame* prev_image, float* d, bool frmhomo) {
    if (!frmhomo) concatenateRelativePose(d, prev_image->poseWrtWorld, poseWrtWorld);
    else for (int i = 0; i < 3; ++i) poseWrtWorld[i] = prev_image->poseWrtWorld[i] + d[i];
}

// ---- depthMap -----------------------------------------------------------------------------------------------------------
depthMap::depthMap() : keyFrame(nullptr), currentFrame(nullptr), stamp(1) {
    hyp_store_.assign((size_t)util::ORIG_COLS * util::ORIG_ROWS, depthhypothesis());
    currentDepthHypothesis = hyp_store_.data();
    for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l) {
        const size_t n = (size_t)(util::ORIG_COLS >> l) * (util::ORIG_ROWS >> l);
        depth_store_[l].assign(n, l == 0 ? -1.0f : 0.0f);
        var_store_[l].assign(n, -1.0f);
        deptharrptr[l] = depth_store_[l].data();
        depthvararrptr[l] = var_store_[l].data();
    }
}

void depthMap::updateDepthImage(bool /*fromKeyFrameCreation*/) {
    const int w = util::ORIG_COLS, h = util::ORIG_ROWS;
    const size_t n = (size_t)w * h;
    std::vector<unsigned char> valid(n), vout(n);
    std::vector<float> idep(n), vs(n);
    for (size_t i = 0; i < n; ++i) {
        valid[i] = currentDepthHypothesis[i].isValid ? 1 : 0;
        idep[i] = currentDepthHypothesis[i].invDepthSmoothed;
        vs[i] = currentDepthHypothesis[i].varianceSmoothed;
    }
    std::lock_guard<std::mutex> lk(g.mu);
    ellc_handle* hd = ctx();
    frame* f = keyFrame;
    int s = f->gpu_kf_slot;
    if (s < 0 || g.kf_owner[s] != f) {                                       // same slot policy as keyframe_slot()
        s = g.next_kf;
        g.next_kf = (g.next_kf + 1) % kKfSlots;
        if (g.kf_owner[s]) { g.kf_owner[s]->gpu_kf_slot = -1; g.kf_owner[s]->gpu_lc_ready = false; }
        g.kf_owner[s] = f; f->gpu_kf_slot = s; f->gpu_lc_ready = false;
        if (ellc_reset_keyframe_weights(hd, s) != ELLC_OK) fail("ellc_reset_keyframe_weights");
    }
    if (ellc_upload_keyframe_hypotheses(hd, s, f->image.ptr<uchar>(0), valid.data(), idep.data(), vs.data(), vout.data()) != ELLC_OK)
        fail("ellc_upload_keyframe_hypotheses");
    for (size_t i = 0; i < n; ++i) currentDepthHypothesis[i].isValid = vout[i] != 0;       // :1279-1282
    for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l) {
        if (ellc_read_keyframe_depth(hd, s, l, f->depth_pyramid[l].ptr<float>(0), depthvararrptr[l]) != ELLC_OK) fail("ellc_read_keyframe_depth");
        const size_t nl = (size_t)(w >> l) * (h >> l);
        const float* d = f->depth_pyramid[l].ptr<float>(0);
        for (size_t i = 0; i < nl; ++i) deptharrptr[l][i] = (l == 0 && depthvararrptr[0][i] < 0) ? -1.0f : d[i];   // deptharrpyr0 uses -1
    }
    std::memcpy(f->depth.ptr<float>(0), f->depth_pyramid[0].ptr<float>(0), n * sizeof(float));
    ++stamp;
    g.kf_stamp[s] = stamp;                                                   // the device copy IS the current one: no re-upload
    if (f->gpu_lc_ready) { const int32_t ks = s; if (ellc_prepare_keyframes_lc(hd, 1, &ks) != ELLC_OK) fail("ellc_prepare_keyframes_lc"); }
}

float depthMap::calculate_no_of_Seeds(bool /*calculate_on_current*/) {
    float count = 0;
    for (int i = 0; i < util::ORIG_COLS * util::ORIG_ROWS; ++i) count += float(currentDepthHypothesis[i].isValid);
    return count / (util::ORIG_COLS * util::ORIG_ROWS) * 100;
}

// ---- PixelWisePyramid ---------------------------------------------------------------------------------------------------
PixelWisePyramid::PixelWisePyramid(frame* prevframe, frame* currentframe, float* /*pose ignored, as the reference*/, depthMap* dm)
    : pyrlevel(prevframe->pyrLevel), nRows(prevframe->currentRows), nCols(prevframe->currentCols), pose(nullptr),
      weightedPose(0.f), want_weight_image(false), residualSum(0.f), prev_frame(prevframe), current_frame(currentframe),
      currentDepthMap(dm) {
    for (int i = 0; i < 6; ++i) prevPose[i] = 0.f;
    hessian = Mat::zeros(6, 6, ellc_host::CV_32FC1);
    sd_param = Mat::zeros(1, 6, ellc_host::CV_32FC1);
    deltapose = Mat::zeros(1, 6, ellc_host::CV_32FC1);
}
PixelWisePyramid::~PixelWisePyramid() {}

void PixelWisePyramid::putPreviousPose(frame* t) { t->concatenateOriginPose(t->poseWrtWorld, prev_frame->poseWrtWorld, prevPose); }

void PixelWisePyramid::calculatePixelWiseParallel() {
    {
        std::lock_guard<std::mutex> lk(g.mu);
        const int ks = keyframe_slot(prev_frame, currentDepthMap), fs = frame_slot(current_frame);
        ellc_iter_trace it;
        float* wimg = nullptr;
        if (want_weight_image) { display_weightimg = Mat::zeros(nRows, nCols, ellc_host::CV_32FC1); wimg = display_weightimg.ptr<float>(0); }
        if (ellc_gn_evaluate(ctx(), ks, fs, pyrlevel, pose, &it, wimg) != ELLC_OK) fail("ellc_gn_evaluate");
        std::memcpy(hessian.ptr<float>(0), it.H, sizeof(it.H));
        std::memcpy(sd_param.ptr<float>(0), it.b, sizeof(it.b));
        residualSum = it.res_sum;
    }
    updatePose();                                                            // hessian.inv() + updatePose(), :451-453
}

void PixelWisePyramid::updatePose() {
    std::lock_guard<std::mutex> lk(g.mu);
    float out[6], d[6], wp = 0.f;
    if (ellc_solve_update(ctx(), hessian.ptr<float>(0), sd_param.ptr<float>(0), pose, out, d, &wp) != ELLC_OK) fail("ellc_solve_update");
    for (int i = 0; i < 6; ++i) { pose[i] = out[i]; deltapose.ptr<float>(0)[i] = d[i]; }
    weightedPose = wp;
}

// ---- GetImagePoseEstimate -------------------------------------------------------------------------------------------------
static void initial_relative_pose(frame* prev_frame, frame* tminus1, float* initial_pose_estimate, bool fromLoopClosure, float pose[6]) {
    // src/ImageFunc.cpp:97-138
    for (int i = 0; i < 6; ++i) pose[i] = 0.0f;
    if (!util::FLAG_INITIALIZE_NONZERO_POSE || fromLoopClosure) {
        prev_frame->concatenateOriginPose(tminus1->poseWrtWorld, prev_frame->poseWrtWorld, pose);
    } else {
        prev_frame->concatenateOriginPose(initial_pose_estimate, prev_frame->poseWrtWorld, pose);
        float pose_trans[6];
        prev_frame->concatenateOriginPose(tminus1->poseWrtWorld, prev_frame->poseWrtWorld, pose_trans);
        pose[3] = pose_trans[3]; pose[4] = pose_trans[4]; pose[5] = pose_trans[5];
    }
}

std::vector<float> GetImagePoseEstimate(frame* prev_frame, frame* current_frame, int /*frame_num*/, depthMap* currDepthMap,
                                        frame* tminus1_prev_frame, float* initial_pose_estimate, bool fromLoopClosure, bool /*homo*/) {
    float pose[6];
    initial_relative_pose(prev_frame, tminus1_prev_frame, initial_pose_estimate, fromLoopClosure, pose);
    {
        std::lock_guard<std::mutex> lk(g.mu);
        ellc_pair pr;
        pr.kf_slot = keyframe_slot(prev_frame, currDepthMap);
        pr.frame_slot = frame_slot(current_frame);
        // src/ImageFunc.cpp:241-244: loop-closure pairs use the constant-weight inverse-compositional tracker once the
        // keyframe's weights are final; :280-288: sequential tracks save their last weights into the keyframe's pyramid
        const bool cw = util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION;
        pr.flags = (cw && fromLoopClosure && prev_frame->gpu_lc_ready) ? ELLC_PAIR_CONST_WEIGHT
                 : (cw && !fromLoopClosure) ? ELLC_PAIR_SAVE_WEIGHTS : ELLC_PAIR_DEFAULT;
        for (int i = 0; i < 6; ++i) pr.init_pose[i] = pose[i];
        ellc_result res;
        if (ellc_track_batch(ctx(), 1, &pr, &res, nullptr) != ELLC_OK) fail("ellc_track_batch");
        for (int i = 0; i < 6; ++i) pose[i] = res.pose[i];
        if (pr.flags == ELLC_PAIR_SAVE_WEIGHTS) {                              // saveWeights(true), src/PixelWisePyramid.cpp:546-548
            const int32_t fs = pr.frame_slot;
            if (ellc_accumulate_weights(ctx(), pr.kf_slot, 1, &fs) != ELLC_OK) fail("ellc_accumulate_weights");
            for (int l = 0; l < util::MAX_PYRAMID_LEVEL; ++l) prev_frame->numWeightsAdded[l]++;
        }
    }
    // post-conditions the depth module relies on (src/ImageFunc.cpp:158-159 end at level 0)
    prev_frame->updationOnPyrChange(0);
    current_frame->updationOnPyrChange(0, false);
    current_frame->calculatePoseWrtOrigin(prev_frame, pose);                 // :305
    current_frame->calculatePoseWrtWorld(prev_frame, pose);                  // :306
    return std::vector<float>(pose, pose + 6);                               // :311-313
}

std::vector<float> ellc_host::TrackPairsBatched(const std::vector<frame*>& keyframes, const std::vector<depthMap*>& depthMaps,
                                                const std::vector<frame*>& frames, const std::vector<float>& init_poses) {
    const size_t n = frames.size();
    std::vector<ellc_pair> pairs(n);
    std::vector<ellc_result> res(n);
    std::lock_guard<std::mutex> lk(g.mu);
    for (size_t i = 0; i < n; ++i) {
        pairs[i].kf_slot = keyframe_slot(keyframes[i], depthMaps[i]);
        pairs[i].frame_slot = frame_slot(frames[i]);
        pairs[i].flags = (util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION && keyframes[i]->gpu_lc_ready) ? ELLC_PAIR_CONST_WEIGHT : ELLC_PAIR_DEFAULT;
        for (int k = 0; k < 6; ++k) pairs[i].init_pose[k] = init_poses[i * 6 + k];
    }
    if (n && ellc_track_batch(ctx(), (int)n, pairs.data(), res.data(), nullptr) != ELLC_OK) fail("ellc_track_batch");
    std::vector<float> out(n * 6);
    for (size_t i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) out[i * 6 + k] = res[i].pose[k];
    return out;
}
