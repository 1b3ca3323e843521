// Drives the C++ host shim exactly the way the reference's main loop drives its tracker (src/main.cpp:199-384):
// one keyframe, consecutive frames, GetImagePoseEstimate initialised from the previous frame's world pose.
// Input blob (written by tests/test_host_shim.py): int32 w, h, n_frames; f32 fx fy cx cy; u8 kf image; 4 depth levels;
// 4 variance levels; n_frames u8 images.  Prints one line per result; the pytest side compares with the CPU oracle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
#include "Pyramid.h"

#include <thread>

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
    if (argc > 2 && std::string(argv[2]) == "--config") {
        // host-only: main()'s argument handling and the batch parameters of config.txt (src/main.cpp:80-101, :132-137)
        std::string msg;
        const char* a1[] = {"ELLC"};
        printf("cfg0 %d\n", ellc_host::configure_from_args(1, a1, &msg));
        const char* a2[] = {"ELLC", "LC"};
        int r = ellc_host::configure_from_args(2, a2, &msg);
        printf("cfg1 %d %s\n", r, msg.c_str());
        const char* a3[] = {"ELLC", "LC", "/nonexistent/config.txt"};
        r = ellc_host::configure_from_args(3, a3, &msg);
        printf("cfg2 %d %s\n", r, msg.c_str());
        const char* a4[] = {"ELLC", "LC", argv[1]};
        const int rc = ellc_host::configure_from_args(3, a4, &msg);
        printf("cfg3 %d %d %d %d %d\n", rc, util::BATCH_START_ID, util::BATCH_SIZE, util::FLAG_IS_BOOTSTRAP ? 1 : 0, util::FLAG_ALTERNATE_GN_RA ? 1 : 0);
        return 0;
    }
    const bool surface = argc > 2 && std::string(argv[2]) == "--surface";
    if (argc > 2 && !surface) { printf("link ok\n"); return 0; }
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
        if (surface) {
            // ---- the rest of the reference's class surface ----------------------------------------------------
            std::vector<frame*> fr;
            for (int i = 0; i < n; ++i) { rd(f, img.data(), img.size()); fr.push_back(new frame(img.data(), w, h)); }
            const int L = 1;
            kf.updationOnPyrChange(L);
            fr[0]->updationOnPyrChange(L, false);
            // (a) frame::getInterpolatedElement, both overloads, inside / on the border / outside
            const float pts[][2] = {{10.25f, 20.75f}, {0.0f, 0.0f}, {-0.5f, 3.2f}, {(float)(w >> L) - 1.0f, 5.5f}, {(float)(w >> L) - 0.5f, 5.5f},
                                    {7.0f, (float)(h >> L) - 0.25f}, {-3.0f, -3.0f}, {33.999f, 41.001f}, {(float)(w >> L) + 2.0f, 1.0f}};
            for (const auto& q : pts)
                printf("interp %.9g %.9g %.9g %.9g %.9g %.9g\n", q[0], q[1], fr[0]->getInterpolatedElement(q[0], q[1], 1), fr[0]->getInterpolatedElement(q[0], q[1], 0),
                       fr[0]->getInterpolatedElement(q[0], q[1], "gradx"), fr[0]->getInterpolatedElement(q[0], q[1], "grady"));
            // (b) PixelWisePyramid with every display member, hessianInv, saveWeights(true / false)
            float pose[6] = {0.002f, -0.001f, 0.0015f, 0.003f, -0.002f, 0.001f};
            float pose_in[6];
            for (int i = 0; i < 6; ++i) pose_in[i] = pose[i];
            PixelWisePyramid pw(&kf, fr[0], pose, &dm);
            pw.pose = pose;
            pw.calculatePixelWiseParallel();
            printf("pw_pose %.9g %.9g %.9g %.9g %.9g %.9g wp %.9g\n", pose[0], pose[1], pose[2], pose[3], pose[4], pose[5], pw.weightedPose);
            double sw = 0, swarp = 0, sres = 0, sorig = 0; long n2 = 0, n1 = 0, tsum = 0, bsum = 0;
            for (int y = 0; y < pw.nRows; ++y)
                for (int x = 0; x < pw.nCols; ++x) {
                    sw += pw.display_weightimg.ptr<float>(y)[x]; swarp += pw.display_warpedimg.ptr<float>(y)[x];
                    sres += pw.display_iterationres.ptr<float>(y)[x]; sorig += pw.display_origres.ptr<float>(y)[x];
                    n2 += pw.savedWarpedPointsX.ptr<float>(y)[x] == -2.0f; n1 += pw.savedWarpedPointsX.ptr<float>(y)[x] == -1.0f;
                    tsum += pw.display_templateimg.ptr<unsigned char>(y)[x]; bsum += pw.display_2bewarpedimg.ptr<unsigned char>(y)[x];
                }
            printf("pw_disp %.9g %.9g %.9g %.9g %ld %ld %ld %ld\n", sw, swarp, sres, sorig, n2, n1, tsum, bsum);
            // a few individual pixels of the planes (selected pixels in raster order) for an independent check on the Python side
            int shown = 0;
            for (int y = 0; y < pw.nRows && shown < 6; y += 7)
                for (int x = 0; x < pw.nCols && shown < 6; x += 11)
                    if (kf.mask.ptr<unsigned char>(y)[x]) {
                        printf("pw_px %d %d %.9g %.9g %.9g %.9g %.9g\n", x, y, pw.savedWarpedPointsX.ptr<float>(y)[x], pw.savedWarpedPointsY.ptr<float>(y)[x],
                               pw.display_warpedimg.ptr<float>(y)[x], pw.display_iterationres.ptr<float>(y)[x], pw.display_weightimg.ptr<float>(y)[x]);
                        ++shown;
                    }
            double hh = 0;                                   // hessian * hessianInv ~ I
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += (double)pw.hessian.ptr<float>(i)[k] * pw.hessianInv.ptr<float>(k)[j];
                    hh = std::max(hh, std::fabs(s - (i == j ? 1.0 : 0.0)));
                }
            printf("pw_hinv %.3g\n", hh);
            pw.saveWeights(true);
            pw.saveWeights(true);
            double ws = 0;
            for (int i = 0; i < pw.nRows * pw.nCols; ++i) ws += kf.weight_pyramid[L].ptr<float>(0)[i];
            printf("pw_save %d %.9g %.9g\n", kf.numWeightsAdded[L], ws, 2 * sw);
            pw.saveWeights(false);
            double cs = 0;
            for (int i = 0; i < pw.nRows * pw.nCols; ++i) cs += fr[0]->weight_pyramid[L].ptr<float>(0)[i];
            printf("pw_scatter %.9g\n", cs);
            // (c) inverse-compositional constant-weight iterations through the class (weights = what saveWeights left, finalised)
            kf.finaliseWeights();                            // (prints the reference's "Weights cannot be averaged" for the empty levels)
            printf("\n");
            float lpose[6] = {0, 0, 0, 0, 0, 0};
            PixelWisePyramid pl(&kf, fr[1], lpose, &dm);
            pl.pose = lpose;
            for (int it = 0; it < 3; ++it) {
                pl.calculatePixelWiseParallelInvCompositional(it);
                printf("lc_iter %d %.9g %.9g %.9g %.9g %.9g %.9g wp %.9g H00 %.9g\n", it, lpose[0], lpose[1], lpose[2], lpose[3], lpose[4], lpose[5],
                       pl.weightedPose, pl.hessian.ptr<float>(0)[0]);
            }
            // (d) class Pyramid (matrix form): performPrecomputation + two performIterationSteps
            float ppose[6];
            for (int i = 0; i < 6; ++i) ppose[i] = pose_in[i];
            Pyramid py(&kf, fr[0], ppose, &dm);
            py.performPrecomputation();
            printf("pyr_pre %.9g %d\n", py.lastErr, py.weights.cols);
            for (int it = 0; it < 2; ++it) {
                const float ratio = py.performIterationSteps();
                printf("pyr_iter %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g ratio %.9g err %.9g\n", it, ppose[0], ppose[1], ppose[2], ppose[3], ppose[4], ppose[5],
                       py.weightedPose, ratio, py.error);
            }
            // (e) calculateRandT post-condition of GetImagePoseEstimate (src/ImageFunc.cpp:307)
            float init0[6] = {0, 0, 0, 0, 0, 0};
            util::FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION = false;
            std::vector<float> p0 = GetImagePoseEstimate(&kf, fr[0], 2, &dm, &kf, init0);
            printf("randt %.9g %.9g %.9g %.9g %.9g %.9g\n", fr[0]->SE3_R[1], fr[0]->SE3_R[5], fr[0]->SE3_T[0], fr[0]->SE3_T[2], fr[0]->SE3_Pose[15], fr[0]->Sim3_R[0]);
            printf("world0 %.9g %.9g %.9g %.9g %.9g %.9g\n", fr[0]->poseWrtWorld[0], fr[0]->poseWrtWorld[1], fr[0]->poseWrtWorld[2], fr[0]->poseWrtWorld[3],
                   fr[0]->poseWrtWorld[4], fr[0]->poseWrtWorld[5]);
            // (f) two host threads, each with its own copies of the frames (the loop-closure thread, src/GlobalOptimize.cpp:181, :566-568):
            //     both must reproduce the single-threaded poses bit for bit, concurrently, on separate contexts
            std::vector<std::vector<float> > want;
            for (int i = 0; i < n; ++i) { float z[6] = {0, 0, 0, 0, 0, 0}; want.push_back(GetImagePoseEstimate(&kf, fr[i], i + 2, &dm, &kf, z)); }
            int bad[2] = {0, 0};
            auto worker = [&](int t) {
                try {
                    frame kfc(kf);                                   // `new frame(*currentframe)`
                    depthMap dmc;
                    dmc.keyFrame = &kfc;
                    for (int l = 0; l < 4; ++l) std::memcpy(dmc.depthvararrptr[l], dm.depthvararrptr[l], (size_t)(w >> l) * (h >> l) * 4);
                    dmc.markDepthUpdated();
                    for (int rep = 0; rep < 6; ++rep)
                        for (int i = 0; i < n; ++i) {
                            frame cur(*fr[i]);
                            float z[6] = {0, 0, 0, 0, 0, 0};
                            std::vector<float> p = GetImagePoseEstimate(&kfc, &cur, i + 2, &dmc, &kfc, z);
                            for (int k = 0; k < 6; ++k) bad[t] += p[k] != want[i][k];
                        }
                } catch (const std::exception& e) { fprintf(stderr, "thread %d: %s\n", t, e.what()); bad[t] += 1000; }
            };
            std::thread t0(worker, 0), t1(worker, 1);
            t0.join(); t1.join();
            printf("threads %d %d contexts %d\n", bad[0], bad[1], ellc_host::context_count());
            // (g) a batch with more distinct frames than there are frame slots (64): split into sub-batches, nothing overwritten
            {
                std::vector<frame*> copies, kfs; std::vector<depthMap*> dms; std::vector<float> inits;
                for (int i = 0; i < 70; ++i) { copies.push_back(new frame(*fr[i % n])); kfs.push_back(&kf); dms.push_back(&dm); for (int k = 0; k < 6; ++k) inits.push_back(0.f); }
                std::vector<float> out = ellc_host::TrackPairsBatched(kfs, dms, copies, inits);
                // (the 64-pair sub-batch runs with 4 CTAs per pair, a pair alone with 8: another summation tree, so the poses agree to
                // rounding, not bit for bit; a recycled slot would show up as a pose of another frame, 1e-3 away)
                double worst = 0;
                for (int i = 0; i < 70; ++i) for (int k = 0; k < 6; ++k) worst = std::max(worst, (double)std::fabs(out[i * 6 + k] - want[i % n][k]));
                printf("bigbatch %.3g\n", worst);
                for (frame* c : copies) delete c;
            }
            for (frame* c : fr) delete c;
            ellc_host::shutdown();
            fclose(f);
            return 0;
        }
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
