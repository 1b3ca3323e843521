// Drives the C++ host shim exactly the way the reference's main loop drives its tracker (src/main.cpp:199-384):
// one keyframe, consecutive frames, GetImagePoseEstimate initialised from the previous frame's world pose.
// Input blob (written by tests/test_host_shim.py): int32 w, h, n_frames; f32 fx fy cx cy; u8 kf image; 4 depth levels;
// 4 variance levels; n_frames u8 images.  Prints one line per result; the pytest side compares with the CPU oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "DepthPropagation.h"
#include "ExternVariable.h"
#include "Frame.h"
#include "ImageFunc.h"
#include "PixelWisePyramid.h"
#include "PoseFiles.h"

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: test_shim blob [--link-only]\n"); return 2; }
    if (argc > 2 && std::string(argv[2]) == "--posefiles") {
        // host-only: the text files the reference's callers write (no GPU involved)
        util::BATCH_START_ID = 101;
        frame kf, f;
        kf.frameId = 1; kf.rescaleFactor = 0.98765432f;
        f.frameId = 7;
        const float w[6] = {0.0123456789f, -1.5e-5f, 3.0f, 123456.789f, -0.000123456f, 1e-10f};
        for (int i = 0; i < 6; ++i) { f.poseWrtWorld[i] = w[i]; f.poseWrtOrigin[i] = -w[5 - i]; }
        std::ofstream o1(std::string(argv[1]) + ".orig"), o2(std::string(argv[1]) + ".match");
        ellc_host::write_orig_pose(o1, &f, &kf, 37.123456f);
        ellc_host::write_match_pose(o2, &f, &kf, 37.123456f);
        ellc_host::write_match_pose(o2, &f, &kf, 12.5f, 0.0712345f, 9.87654321f, 4.5f);
        o1.close(); o2.close();
        std::ifstream in(std::string(argv[1]) + ".orig");
        int id, kid; float p[6];
        in >> id; in.seekg(0);
        std::ifstream init(std::string(argv[1]) + ".orig");
        int fn; init >> fn >> kid;                       // skip the two ids, then read six floats back
        std::stringstream ss; ss << fn << " "; for (int i = 0; i < 6; ++i) { float v; init >> v; ss << v << " "; }
        int fno; const bool ok = ellc_host::read_initial_pose(ss, fno, p);
        printf("posefiles ok %d %d %.9g\n", ok ? 1 : 0, fno, p[3]);
        return 0;
    }
    if (argc > 2) { printf("link ok\n"); return 0; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int w, h, n; float k[4];
    rd(f, &w, 4); rd(f, &h, 4); rd(f, &n, 4); rd(f, k, 16);
    util::configure(w, h, k[0], k[1], k[2], k[3]);
    std::vector<unsigned char> img((size_t)w * h);
    rd(f, img.data(), img.size());
    try {
        frame kf(img.data(), w, h);
        depthMap dm;
        dm.keyFrame = &kf;
        for (int l = 0; l < 4; ++l) rd(f, kf.depth_pyramid[l].ptr<float>(0), (size_t)(w >> l) * (h >> l) * 4);
        for (int l = 0; l < 4; ++l) rd(f, dm.depthvararrptr[l], (size_t)(w >> l) * (h >> l) * 4);
        dm.markDepthUpdated();
        printf("pyr %d %d %d %d\n", kf.image_pyramid[1].cols, kf.image_pyramid[1].rows, kf.image_pyramid[3].cols, kf.image_pyramid[3].rows);
        frame* prev = &kf;
        std::vector<frame*> frames;
        for (int i = 0; i < n; ++i) {
            rd(f, img.data(), img.size());
            frame* cur = new frame(img.data(), w, h);
            frames.push_back(cur);
            float init[6] = {0, 0, 0, 0, 0, 0};
            std::vector<float> p = GetImagePoseEstimate(&kf, cur, i + 2, &dm, prev, init);
            printf("pose %d %.9g %.9g %.9g %.9g %.9g %.9g\n", i, p[0], p[1], p[2], p[3], p[4], p[5]);
            printf("world %d %.9g %.9g %.9g %.9g %.9g %.9g\n", i, cur->poseWrtWorld[0], cur->poseWrtWorld[1], cur->poseWrtWorld[2],
                   cur->poseWrtWorld[3], cur->poseWrtWorld[4], cur->poseWrtWorld[5]);
            printf("post %d %d %d %d %d %d\n", i, kf.pyrLevel, cur->pyrLevel, kf.no_nonZeroDepthPts, cur->gradientx.cols, kf.mask.rows);
            prev = cur;
        }
        // caller-driven iterations through the PixelWisePyramid surface (src/ImageFunc.cpp:163-253) at level 2
        kf.updationOnPyrChange(2);
        frames[0]->updationOnPyrChange(2, false);
        float pose[6] = {0, 0, 0, 0, 0, 0};
        PixelWisePyramid pw(&kf, frames[0], pose, &dm);
        pw.pose = pose;
        for (int it = 0; it < 3; ++it) {
            pw.calculatePixelWiseParallel();
            printf("iter %d %.9g %.9g %.9g %.9g %.9g %.9g wp %.9g res %.9g H00 %.9g\n", it, pose[0], pose[1], pose[2], pose[3], pose[4], pose[5],
                   pw.weightedPose, pw.residualSum, pw.hessian.ptr<float>(0)[0]);
        }
        printf("count2 %d\n", kf.no_nonZeroDepthPts);
        // constant-weight loop-closure flow (FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION): sequential tracks save their weights
        // (src/ImageFunc.cpp:280-288), the keyframe is finalised (src/main.cpp:431-434), then a loop-closure pair runs the
        // inverse-compositional tracker (src/ImageFunc.cpp:241-244)
        // keyframe from depth hypotheses (depthMap::updateDepthImage on the device): rebuild the same depth from 1/depth hypotheses
        {
            frame kf2(kf.image.ptr<unsigned char>(0), w, h);
            depthMap dm2;
            dm2.keyFrame = &kf2;
            const float* d0 = kf.depth_pyramid[0].ptr<float>(0);
            for (int i = 0; i < w * h; ++i) {
                dm2.currentDepthHypothesis[i].isValid = d0[i] > 0;
                dm2.currentDepthHypothesis[i].invDepthSmoothed = d0[i] > 0 ? 1.0f / d0[i] : -1.0f;
                dm2.currentDepthHypothesis[i].varianceSmoothed = dm.depthvararrptr[0][i];
            }
            const float occ_in = dm2.calculate_no_of_Seeds();
            dm2.updateDepthImage();
            int c1 = 0, c3 = 0;
            for (int i = 0; i < (w >> 1) * (h >> 1); ++i) c1 += kf2.depth_pyramid[1].ptr<float>(0)[i] > 0;
            for (int i = 0; i < (w >> 3) * (h >> 3); ++i) c3 += kf2.depth_pyramid[3].ptr<float>(0)[i] > 0;
            float init[6] = {0, 0, 0, 0, 0, 0};
            std::vector<float> p = GetImagePoseEstimate(&kf2, frames[0], 2, &dm2, &kf2, init);
            printf("hyp %.9g %.9g %d %d %.9g %.9g %.9g\n", occ_in, dm2.calculate_no_of_Seeds(), c1, c3, p[0], p[1], p[2]);
        }
        util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION = true;
        prev = &kf;
        for (int i = 0; i < n; ++i) {
            float init[6] = {0, 0, 0, 0, 0, 0};
            GetImagePoseEstimate(&kf, frames[i], i + 2, &dm, prev, init);
            prev = frames[i];
        }
        kf.finaliseWeights();
        printf("nweights %d %d\n", kf.numWeightsAdded[0], kf.numWeightsAdded[3]);
        {
            double ws = 0;
            const float* wp0 = kf.weight_pyramid[1].ptr<float>(0);
            for (int i = 0; i < (w >> 1) * (h >> 1); ++i) ws += wp0[i];
            printf("wsum1 %.9g\n", ws);
            float init[6] = {0, 0, 0, 0, 0, 0};
            std::vector<float> p = GetImagePoseEstimate(&kf, frames[n - 1], 99, &dm, frames[0], init, true);
            printf("lcpose %.9g %.9g %.9g %.9g %.9g %.9g\n", p[0], p[1], p[2], p[3], p[4], p[5]);
        }
        for (frame* c : frames) delete c;
    } catch (const std::exception& e) {
        fprintf(stderr, "shim error: %s\n", e.what());
        ellc_host::shutdown();
        return 3;
    }
    ellc_host::shutdown();
    fclose(f);
    return 0;
}
