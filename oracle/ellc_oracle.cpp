/*
 * ellc_oracle.cpp -- CPU ORACLE (test infrastructure, NOT the product; see ellc_oracle.h).
 *
 * Restates, function by function, the reference's CPU tracker.  File:line citations refer to
 * /root/reference (IITD-COMPUTER-VISION-GROUP/Egomotion_with_Local_Loop_Closures).  The arithmetic
 * mirrors the reference's types: fp32 everywhere, except the sub-expressions that C++11 promotes to
 * double through std::pow(float,int) (src/PixelWisePyramid.cpp:296,300,306,308,311-312), and the
 * cv::gemm double accumulators (OpenCV GEMMSingleMul<float,double>).  Build with
 * `-std=c++11 -O3 -ffp-contract=off` (the reference's CMakeLists.txt:11,20 flags; x86-64 baseline has
 * no FMA, -ffp-contract=off makes that explicit).
 *
 * Third-party arithmetic that is NOT in the reference tree is restated from the published algorithms:
 *   - cv::pyrDown (OpenCV 3.0.0 imgproc/pyramids.cpp): separable [1 4 6 4 1], REFLECT_101, (s+128)>>8.
 *   - cv::Mat::inv DECOMP_LU (OpenCV 3.0.0 core/lapack.cpp LUImpl<float>, eps = FLT_EPSILON*10).
 *   - cv::gemm small-matrix path (double accumulator, result rounded to float).
 *   - Eigen 3.2.5 unsupported/MatrixFunctions MatrixExponential<float> (Pade 3/5/7 + squaring).
 *   - Eigen .log(): replaced by the exact SE(3) logarithm evaluated in double and rounded to float
 *     (Eigen's float Schur-based log agrees with it to a few ulp; documented in DESIGN.md).
 */
#include "ellc_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

// UNZERO macro, src/ExternVariable.h:232 (double constants, result rounded to float on assignment)
inline float unzero(float v) {
    double r = (v < 0 ? (v > -1e-10 ? -1e-10 : (double)v) : (v < 1e-10 ? 1e-10 : (double)v));
    return (float)r;
}

struct Intr { float fx, fy, cx, cy; };

// GetIntrinsic, src/UserDefinedFunc.cpp:33-49: float / pow(2,level) evaluated in double, stored as float.
inline Intr level_intrinsics(const ellc_oracle_config* c, int level) {
    Intr k;
    double s = std::pow(2.0, level);
    k.fx = (float)(c->fx / s);
    k.fy = (float)(c->fy / s);
    k.cx = (float)(c->cx / s);
    k.cy = (float)(c->cy / s);
    return k;
}

// ------------------------------------------------------------------------------------------------
// 4x4 float matrix toolbox for the Eigen restatement
// ------------------------------------------------------------------------------------------------
struct M4 { float a[16]; };

inline M4 m4_identity() { M4 r; std::memset(r.a, 0, sizeof(r.a)); r.a[0] = r.a[5] = r.a[10] = r.a[15] = 1.f; return r; }
inline M4 m4_mul(const M4& x, const M4& y) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = x.a[i * 4 + 0] * y.a[0 * 4 + j];
            for (int k = 1; k < 4; ++k) s += x.a[i * 4 + k] * y.a[k * 4 + j];
            r.a[i * 4 + j] = s;
        }
    return r;
}
// r = c2*X + c1*Y + c0*I (used for the Pade polynomials)
inline M4 m4_lin(float cx, const M4* X, float cy, const M4* Y, float cz, const M4* Z, float ci) {
    M4 r;
    for (int i = 0; i < 16; ++i) {
        float s = 0.f;
        bool first = true;
        if (X) { s = cx * X->a[i]; first = false; }
        if (Y) { s = first ? cy * Y->a[i] : s + cy * Y->a[i]; first = false; }
        if (Z) { s = first ? cz * Z->a[i] : s + cz * Z->a[i]; first = false; }
        float id = (i % 5 == 0) ? ci : 0.f;
        r.a[i] = first ? id : s + id;
    }
    return r;
}
// Solve A X = B (n x n, float) with partial-pivot LU, as Eigen's PartialPivLU::solve.
void lu_solve_f32(int n, const float* A_in, const float* B_in, float* X) {
    std::vector<float> A(A_in, A_in + n * n), B(B_in, B_in + n * n);
    for (int i = 0; i < n; ++i) {
        int p = i;
        for (int r = i + 1; r < n; ++r)
            if (std::fabs(A[r * n + i]) > std::fabs(A[p * n + i])) p = r;
        if (p != i) {
            for (int c = 0; c < n; ++c) { std::swap(A[i * n + c], A[p * n + c]); std::swap(B[i * n + c], B[p * n + c]); }
        }
        float piv = A[i * n + i];
        for (int r = i + 1; r < n; ++r) {
            float f = A[r * n + i] / piv;
            A[r * n + i] = f;
            for (int c = i + 1; c < n; ++c) A[r * n + c] -= f * A[i * n + c];
            for (int c = 0; c < n; ++c) B[r * n + c] -= f * B[i * n + c];
        }
    }
    for (int c = 0; c < n; ++c)
        for (int i = n - 1; i >= 0; --i) {
            float s = B[i * n + c];
            for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * X[k * n + c];
            X[i * n + c] = s / A[i * n + i];
        }
}

// hat(): src/PixelWisePyramid.cpp:153 -- [[0,-wz,wy,vx],[wz,0,-wx,vy],[-wy,wx,0,vz],[0,0,0,0]]
inline M4 se3_hat(const float p[6]) {
    M4 m; std::memset(m.a, 0, sizeof(m.a));
    m.a[1] = -p[2]; m.a[2] = p[1];  m.a[3] = p[3];
    m.a[4] = p[2];  m.a[6] = -p[0]; m.a[7] = p[4];
    m.a[8] = -p[1]; m.a[9] = p[0];  m.a[11] = p[5];
    return m;
}

// Eigen 3.2.5 MatrixExponential<MatrixXf>::compute -- Pade(3|5|7) selected on the L1 norm, float thresholds.
M4 mat_exp_f32(const M4& M) {
    float l1 = 0.f;
    for (int j = 0; j < 4; ++j) {
        float s = 0.f;
        for (int i = 0; i < 4; ++i) s += std::fabs(M.a[i * 4 + j]);
        l1 = std::max(l1, s);
    }
    M4 U, V;
    int squarings = 0;
    if (l1 < 4.258730016922831e-001f) {
        const float b[] = {120.f, 60.f, 12.f, 1.f};
        M4 A2 = m4_mul(M, M);
        M4 t = m4_lin(b[3], &A2, 0, nullptr, 0, nullptr, b[1]);
        U = m4_mul(M, t);
        V = m4_lin(b[2], &A2, 0, nullptr, 0, nullptr, b[0]);
    } else if (l1 < 1.880152677804762e+000f) {
        const float b[] = {30240.f, 15120.f, 3360.f, 420.f, 30.f, 1.f};
        M4 A2 = m4_mul(M, M), A4 = m4_mul(A2, A2);
        M4 t = m4_lin(b[5], &A4, b[3], &A2, 0, nullptr, b[1]);
        U = m4_mul(M, t);
        V = m4_lin(b[4], &A4, b[2], &A2, 0, nullptr, b[0]);
    } else {
        const float maxnorm = 3.925724783138660f;
        std::frexp(l1 / maxnorm, &squarings);
        if (squarings < 0) squarings = 0;
        M4 A = M;
        float sc = std::pow(2.0f, (float)squarings);
        for (int i = 0; i < 16; ++i) A.a[i] = M.a[i] / sc;
        const float b[] = {17297280.f, 8648640.f, 1995840.f, 277200.f, 25200.f, 1512.f, 56.f, 1.f};
        M4 A2 = m4_mul(A, A), A4 = m4_mul(A2, A2), A6 = m4_mul(A4, A2);
        M4 t = m4_lin(b[7], &A6, b[5], &A4, b[3], &A2, b[1]);
        U = m4_mul(A, t);
        V = m4_lin(b[6], &A6, b[4], &A4, b[2], &A2, b[0]);
    }
    M4 num, den, R;
    for (int i = 0; i < 16; ++i) { num.a[i] = U.a[i] + V.a[i]; den.a[i] = -U.a[i] + V.a[i]; }
    lu_solve_f32(4, den.a, num.a, R.a);
    for (int s = 0; s < squarings; ++s) R = m4_mul(R, R);
    return R;
}

// Eigen .inverse() on a dynamic 4x4 float (PartialPivLU based), src/Frame.cpp:550.
M4 m4_inverse(const M4& T) {
    M4 I = m4_identity(), R;
    lu_solve_f32(4, T.a, I.a, R.a);
    return R;
}

// log of a rigid transform (Eigen .log(), src/Frame.cpp:521/553), evaluated in double from the float
// entries and rounded to float; extraction order of src/Frame.cpp:523-528.
void rigid_log(const M4& T, float out[6]) {
    const double R00 = T.a[0], R01 = T.a[1], R02 = T.a[2];
    const double R10 = T.a[4], R11 = T.a[5], R12 = T.a[6];
    const double R20 = T.a[8], R21 = T.a[9], R22 = T.a[10];
    const double t[3] = {T.a[3], T.a[7], T.a[11]};
    double ax = 0.5 * (R21 - R12), ay = 0.5 * (R02 - R20), az = 0.5 * (R10 - R01);   // sin(th) * n
    double s = std::sqrt(ax * ax + ay * ay + az * az);
    double c = 0.5 * (R00 + R11 + R22 - 1.0);
    double th = std::atan2(s, c);
    double k;                                   // omega = k * (ax,ay,az)
    if (s < 1e-7) k = (c > 0) ? 1.0 + th * th / 6.0 : 0.0;   // (theta ~ pi is outside this tracker's domain)
    else k = th / s;
    double w[3] = {k * ax, k * ay, k * az};
    // V^-1 = I - 1/2 W + coef W^2, coef = (1 - th*sin/(2(1-cos)))/th^2
    double coef;
    if (th < 1e-4) coef = 1.0 / 12.0 + th * th / 720.0;
    else coef = (1.0 - (th * std::sin(th)) / (2.0 * (1.0 - std::cos(th)))) / (th * th);
    // W t  and  W (W t)
    double wt[3] = {w[1] * t[2] - w[2] * t[1], w[2] * t[0] - w[0] * t[2], w[0] * t[1] - w[1] * t[0]};
    double wwt[3] = {w[1] * wt[2] - w[2] * wt[1], w[2] * wt[0] - w[0] * wt[2], w[0] * wt[1] - w[1] * wt[0]};
    out[0] = (float)w[0]; out[1] = (float)w[1]; out[2] = (float)w[2];
    for (int i = 0; i < 3; ++i) out[3 + i] = (float)(t[i] - 0.5 * wt[i] + coef * wwt[i]);
}

// ------------------------------------------------------------------------------------------------
// the per-band pixel loop: PixelWisePyramid::calculatePixelWise, src/PixelWisePyramid.cpp:58-413
// ------------------------------------------------------------------------------------------------
struct BandAcc {
    float H[36];
    float b[6];
    double bd[6];        // double accumulator used only by the Pyramid.cpp variant (cv::gemm 1xN * Nx6)
    float res_f32;
    double res_f64;
    int n_oob;
    double Hd[36];       // same fp32 products, double accumulation (diagnostic)
    double bsum[6];
};

struct LevelView {
    int rows, cols;                  // currentRows/currentCols = height/2^L, width/2^L
    const uint8_t* kf; int kf_stride;
    const uint8_t* cur; int cur_stride;
    const float* gx; const float* gy;   // current frame gradients (cols stride)
    const uint8_t* mask;             // cols stride
    const float* depth;              // stride cols (depth_pyramid / deptharrptr share W>>L)
    const float* var;                // depthvararrptr[L], stride cols
};

void pixelwise_band(const ellc_oracle_config* cfg, const Intr& K, const LevelView& lv, const float SE3v[12],
                    int ymin, int ymax, BandAcc* acc, float* weight_img) {
    std::memset(acc, 0, sizeof(*acc));
    const float fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;
    const float tx = SE3v[3], ty = SE3v[7], tz = SE3v[11];          // :176-178
    const bool at_warped = cfg->jacobian_at_warped != 0;
    const float huber_half = cfg->huber_d / 2;

    for (int y = ymin; y < ymax; ++y) {
        for (int x = 0; x < lv.cols; ++x) {
            const int idx = x + lv.cols * y;
            if (lv.mask[idx] == 0) { if (weight_img) weight_img[idx] = 0.f; continue; }     // :209-221
            const float dep = lv.depth[idx];
            // back-projection :236-238
            float wX = (x - cx) * dep / fx;
            float wY = (y - cy) * dep / fy;
            float wZ = dep;
            // rigid transform :242-264 (both textual branches are the same fp32 op sequence)
            float tX = ((SE3v[0] * wX) + (SE3v[1] * wY) + (SE3v[2] * wZ) + (SE3v[3]));
            float tY = ((SE3v[4] * wX) + (SE3v[5] * wY) + (SE3v[6] * wZ) + (SE3v[7]));
            float tZ = ((SE3v[8] * wX) + (SE3v[9] * wY) + (SE3v[10] * wZ) + (SE3v[11]));
            tZ = unzero(tZ);
            float u = ((tX / tZ) * fx) + cx;
            float v = ((tY / tZ) * fy) + cy;
            // sampling :271,291-292
            float Iw = ellc_oracle_interp_u8(lv.cur, lv.cur_stride, lv.rows, lv.cols, u, v, 1);
            float gradx = ellc_oracle_interp_f32(lv.gx, lv.cols, lv.rows, lv.cols, u, v);
            float grady = ellc_oracle_interp_f32(lv.gy, lv.cols, lv.rows, lv.cols, u, v);
            // Jacobian :296-320 (PixelWise: keyframe pixel & depth) / Pyramid.cpp:99-130 (warped pixel & Z')
            float jb[6], jt[6], J[6];
            if (!at_warped) {
                const float yc = -cy + y, xc = -cx + x;         // float + int -> float
                const double idep = std::pow((double)dep, -1);
                jb[0] = (float)(grady * (-(fy + (std::pow((double)yc, 2) / fy))));
                jt[0] = gradx * (-(yc * xc) / fy);
                jb[1] = grady * ((yc * xc) / fx);
                jt[1] = (float)(gradx * (fx + (std::pow((double)xc, 2) / fx)));
                jb[2] = grady * ((fy * xc) / fx);
                jt[2] = gradx * (-(fx * yc / fy));
                jb[3] = 0;
                jt[3] = (float)(gradx * (fx * idep));
                jb[4] = (float)(grady * (fy * idep));
                jt[4] = 0;
                jb[5] = (float)(grady * (-yc * idep));
                jt[5] = (float)(gradx * (-xc * idep));
            } else {
                const float yc = -cy + v, xc = -cx + u;
                const double idep = std::pow((double)tZ, -1);
                jb[0] = (float)(grady * (-(fy + (std::pow((double)yc, 2) / fy))));
                jt[0] = gradx * (-(yc * xc) / fy);
                jb[1] = grady * ((yc * xc) / fx);
                jt[1] = (float)(gradx * (fx + (std::pow((double)xc, 2) / fx)));
                jb[2] = grady * ((fy * xc) / fx);
                jt[2] = gradx * (-(fx * yc / fy));
                jb[3] = 0;
                jt[3] = (float)(gradx * (fx * idep));
                jb[4] = (float)(grady * (fy * idep));
                jt[4] = 0;
                jb[5] = (float)(grady * (-yc * idep));
                jt[5] = (float)(gradx * (-xc * idep));
            }
            for (int i = 0; i < 6; ++i) J[i] = jt[i] + jb[i];                      // :315-320
            // residual :325-330
            const bool oob = (Iw == -1);
            float residual = oob ? 0.0f : Iw - float(lv.kf[x + lv.kf_stride * y]);
            // weight :334-359 (Pyramid.cpp:629-651 does not zero the weight of OOB pixels)
            float res_weight;
            if (oob && !at_warped) {
                res_weight = 0;
            } else {
                float px = tX, py = tY, pz = tZ;
                float d = 1.0f / dep;
                float rp = residual;
                float gxs = fx * gradx;
                float gys = fy * grady;
                float s = 1.0f * lv.var[idx];
                float g0 = (tx * pz - tz * px) / (pz * pz * d);
                float g1 = (ty * pz - tz * py) / (pz * pz * d);
                float drpdd = gxs * g0 + gys * g1;
                float w_p = 1.0f / (cfg->camera_pixel_noise_2 + s * drpdd * drpdd);
                float weighted_rp = std::fabs(rp * sqrtf(w_p));
                float wh = std::fabs(weighted_rp < huber_half ? 1 : huber_half / weighted_rp);
                res_weight = wh * w_p;
            }
            if (oob) acc->n_oob++;
            if (weight_img) weight_img[idx] = res_weight;                            // :361
            // accumulation :364-374  (cv::gemm 6x1*1x6 = rounded fp32 products; += is fp32)
            float wJ[6];
            for (int i = 0; i < 6; ++i) wJ[i] = J[i] * res_weight;
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) { const float pr = wJ[i] * J[j]; acc->H[i * 6 + j] += pr; acc->Hd[i * 6 + j] += (double)pr; }
            const float rw = residual * res_weight;
            for (int i = 0; i < 6; ++i) acc->bsum[i] += (double)(J[i] * rw);
            if (!at_warped) {
                for (int i = 0; i < 6; ++i) acc->b[i] += J[i] * rw;
            } else {
                for (int i = 0; i < 6; ++i) acc->bd[i] += (double)rw * (double)J[i];  // Pyramid.cpp:531-533
            }
            // residual statistic (definition of Pyramid.cpp:682; commented out at PixelWisePyramid.cpp:357)
            const float term = res_weight * residual * residual;
            acc->res_f32 += term;
            acc->res_f64 += (double)term;
        }
    }
    if (at_warped) for (int i = 0; i < 6; ++i) acc->b[i] = (float)acc->bd[i];
}

void build_level_mask(const float* depth, int n, std::vector<uint8_t>& mask, int* count) {
    mask.resize(n);
    *count = ellc_oracle_mask_count(depth, n, mask.data());
}

// calculatePixelWiseParallel minus inversion/update: src/PixelWisePyramid.cpp:416-442
void evaluate_level(const ellc_oracle_config* cfg, int level, const LevelView& lv, const float pose[6],
                    ellc_oracle_iter* out, float* weight_img) {
    const Intr K = level_intrinsics(cfg, level);
    M4 T = mat_exp_f32(se3_hat(pose));                                       // :153-159
    float SE3v[12];
    for (int i = 0; i < 12; ++i) SE3v[i] = T.a[i];                           // :162-173

    const int nb = cfg->jacobian_at_warped ? 1 : std::max(1, cfg->num_bands);
    std::vector<BandAcc> acc(nb);
    const int inc = lv.rows / nb;                                            // :426
    auto run = [&](int t) {
        int y0 = t * inc, y1 = (t == nb - 1) ? lv.rows : (t + 1) * inc;       // :432-434
        pixelwise_band(cfg, K, lv, SE3v, y0, y1, &acc[t], weight_img);
    };
    if (cfg->use_threads && nb > 1) {
        std::vector<std::thread> th;
        for (int t = 0; t < nb; ++t) th.emplace_back(run, t);
        for (auto& t : th) t.join();
    } else {
        for (int t = 0; t < nb; ++t) run(t);
    }
    // :441-442  H = H1 + H2 + H3 (left to right, fp32)
    std::memcpy(out->H, acc[0].H, sizeof(out->H));
    std::memcpy(out->b, acc[0].b, sizeof(out->b));
    out->res_sum_f32 = acc[0].res_f32;
    out->res_sum_f64 = acc[0].res_f64;
    out->n_oob = acc[0].n_oob;
    for (int i = 0; i < 36; ++i) out->H_f64[i] = acc[0].Hd[i];
    for (int i = 0; i < 6; ++i) out->b_f64[i] = acc[0].bsum[i];
    for (int t = 1; t < nb; ++t) {
        for (int i = 0; i < 36; ++i) out->H_f64[i] += acc[t].Hd[i];
        for (int i = 0; i < 6; ++i) out->b_f64[i] += acc[t].bsum[i];
        for (int i = 0; i < 36; ++i) out->H[i] = out->H[i] + acc[t].H[i];
        for (int i = 0; i < 6; ++i) out->b[i] = out->b[i] + acc[t].b[i];
        out->res_sum_f32 += acc[t].res_f32;
        out->res_sum_f64 += acc[t].res_f64;
        out->n_oob += acc[t].n_oob;
    }
}

void track_impl(const ellc_oracle_config* cfg, const uint8_t* const* kf_pyr, const uint8_t* const* cur_pyr,
                const float* const* depth, const float* const* var, const float init_pose[6],
                float out_pose[6], ellc_oracle_trace* trace, float* const* weight_last = nullptr) {
    float pose[6];
    for (int i = 0; i < 6; ++i) pose[i] = init_pose[i];
    if (trace) std::memset(trace, 0, sizeof(*trace));

    int pw[ELLC_ORACLE_LEVELS];
    pw[0] = cfg->width;
    for (int l = 1; l < ELLC_ORACLE_LEVELS; ++l) pw[l] = (pw[l - 1] + 1) / 2;     // cv::pyrDown output width

    std::vector<float> gx, gy;
    std::vector<uint8_t> mask;
    for (int level = ELLC_ORACLE_LEVELS - 1; level >= 0; --level) {               // src/ImageFunc.cpp:150
        LevelView lv;
        lv.rows = (int)(cfg->height / std::pow(2.0, level));                      // src/Frame.cpp:321-322
        lv.cols = (int)(cfg->width / std::pow(2.0, level));
        lv.kf = kf_pyr[level]; lv.kf_stride = pw[level];
        lv.cur = cur_pyr[level]; lv.cur_stride = pw[level];
        lv.depth = depth[level]; lv.var = var[level];
        int count = 0;
        build_level_mask(depth[level], lv.rows * lv.cols, mask, &count);          // ImageFunc.cpp:158
        gx.assign((size_t)lv.rows * lv.cols, 0.f); gy.assign((size_t)lv.rows * lv.cols, 0.f);
        ellc_oracle_gradient(lv.cur, lv.cur_stride, lv.rows, lv.cols, gx.data(), gy.data());   // :159
        lv.gx = gx.data(); lv.gy = gy.data(); lv.mask = mask.data();
        if (trace) trace->n_selected[level] = count;

        int executed = 0;
        for (int iter = 0; iter < cfg->max_iter[level]; ++iter) {                 // ImageFunc.cpp:192
            ellc_oracle_iter local;
            ellc_oracle_iter* rec = (trace && iter < ELLC_ORACLE_MAX_ITERS) ? &trace->it[level][iter] : &local;
            // display_weightimg is rewritten by every iteration; what saveWeights(true) adds at the end of the level
            // (src/ImageFunc.cpp:280-288) is the image of the last executed one
            evaluate_level(cfg, level, lv, pose, rec, weight_last ? weight_last[level] : nullptr);
            float Hinv[36];
            ellc_oracle_invert6(rec->H, Hinv);                                    // PixelWisePyramid.cpp:451
            ellc_oracle_update_pose(cfg, Hinv, rec->b, pose, rec->delta, &rec->weighted_pose);   // :453
            for (int i = 0; i < 6; ++i) rec->pose_after[i] = pose[i];
            ++executed;
            if (rec->weighted_pose < cfg->stop_threshold) break;                  // ImageFunc.cpp:251-252
        }
        if (trace) trace->n_iters[level] = executed;
    }
    for (int i = 0; i < 6; ++i) out_pose[i] = pose[i];
    if (trace) for (int i = 0; i < 6; ++i) trace->final_pose[i] = pose[i];
}

// ------------------------------------------------------------------------------------------------
// inverse-compositional constant-weight variant, src/PixelWisePyramid.cpp:561-974
// ------------------------------------------------------------------------------------------------
struct LcBand { float b[6]; double bd[6]; float res_f32; double res_f64; int n_oob; };

// iteratePixelWiseInvCompositional, :687-913, rows [ymin, ymax)
void lc_iterate_band(const Intr& K, const LevelView& lv, const float* weight, const float* J /* 6 x (rows*cols) */,
                     const float SE3v[12], int ymin, int ymax, LcBand* acc) {
    std::memset(acc, 0, sizeof(*acc));
    const float fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;
    const size_t np = (size_t)lv.rows * lv.cols;
    for (int y = ymin; y < ymax; ++y) {
        for (int x = 0; x < lv.cols; ++x) {
            const int idx = x + lv.cols * y;
            if (lv.mask[idx] == 0) continue;                                          // :803-812
            const float dep = lv.depth[idx];
            float wX = (x - cx) * dep / fx;                                           // :824-826
            float wY = (y - cy) * dep / fy;
            float wZ = dep;
            float tX = ((SE3v[0] * wX) + (SE3v[1] * wY) + (SE3v[2] * wZ) + (SE3v[3]));   // :833-855
            float tY = ((SE3v[4] * wX) + (SE3v[5] * wY) + (SE3v[6] * wZ) + (SE3v[7]));
            float tZ = ((SE3v[8] * wX) + (SE3v[9] * wY) + (SE3v[10] * wZ) + (SE3v[11]));
            tZ = unzero(tZ);
            float u = ((tX / tZ) * fx) + cx;
            float v = ((tY / tZ) * fy) + cy;
            float Iw = ellc_oracle_interp_u8(lv.cur, lv.cur_stride, lv.rows, lv.cols, u, v, 1);   // :862
            const bool oob = (Iw == -1);
            float residual = oob ? 0.0f : Iw - float(lv.kf[x + lv.kf_stride * y]);    // :873-878
            if (oob) acc->n_oob++;
            const float rw = residual * weight[idx];                                  // :890 residual*weight_ptr[x]
            for (int i = 0; i < 6; ++i) {
                const float t = J[i * np + idx] * rw;                                 // steepestdescentMat.mul(...)
                acc->b[i] += t;
                acc->bd[i] += (double)t;
            }
            const float term = weight[idx] * residual * residual;
            acc->res_f32 += term;
            acc->res_f64 += (double)term;
        }
    }
}

void track_lc_impl(const ellc_oracle_config* cfg, const uint8_t* const* kf_pyr, const uint8_t* const* cur_pyr,
                   const float* const* depth, const float* const* weight, const float init_pose[6],
                   float out_pose[6], ellc_oracle_trace* trace) {
    float pose[6];
    for (int i = 0; i < 6; ++i) pose[i] = init_pose[i];
    if (trace) std::memset(trace, 0, sizeof(*trace));
    int pw[ELLC_ORACLE_LEVELS];
    pw[0] = cfg->width;
    for (int l = 1; l < ELLC_ORACLE_LEVELS; ++l) pw[l] = (pw[l - 1] + 1) / 2;

    std::vector<float> gx, gy, J;
    std::vector<uint8_t> mask;
    for (int level = ELLC_ORACLE_LEVELS - 1; level >= 0; --level) {
        LevelView lv;
        lv.rows = (int)(cfg->height / std::pow(2.0, level));
        lv.cols = (int)(cfg->width / std::pow(2.0, level));
        lv.kf = kf_pyr[level]; lv.kf_stride = pw[level];
        lv.cur = cur_pyr[level]; lv.cur_stride = pw[level];
        lv.depth = depth[level]; lv.var = nullptr;
        const size_t np = (size_t)lv.rows * lv.cols;
        int count = 0;
        build_level_mask(depth[level], (int)np, mask, &count);
        lv.mask = mask.data();
        // prev_frame->gradientx/y at this level (updationOnPyrChange -> calculateGradient, src/Frame.cpp:316-327)
        gx.assign(np, 0.f); gy.assign(np, 0.f);
        ellc_oracle_gradient(lv.kf, lv.kf_stride, lv.rows, lv.cols, gx.data(), gy.data());
        lv.gx = nullptr; lv.gy = nullptr;
        if (trace) trace->n_selected[level] = count;
        const Intr K = level_intrinsics(cfg, level);
        const float fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;

        // ---- precomputePixelWiseInvCompositional :561-680 (its three row bands write disjoint pixels) --------
        J.assign(6 * np, 0.f);
        double Hd[36];
        for (int i = 0; i < 36; ++i) Hd[i] = 0.0;
        for (int y = 0; y < lv.rows; ++y) {
            for (int x = 0; x < lv.cols; ++x) {
                const int idx = x + lv.cols * y;
                if (mask[idx] == 0) continue;                                         // :613-631 (zeros)
                const float gradx = gx[idx], grady = gy[idx], dep = lv.depth[idx];
                const float yc = -cy + y, xc = -cx + x;
                const double idep = std::pow((double)dep, -1);
                float jb[6], jt[6];
                jb[0] = (float)(grady * (-(fy + (std::pow((double)yc, 2) / fy))));     // :639-657
                jt[0] = gradx * (-(yc * xc) / fy);
                jb[1] = grady * ((yc * xc) / fx);
                jt[1] = (float)(gradx * (fx + (std::pow((double)xc, 2) / fx)));
                jb[2] = grady * ((fy * xc) / fx);
                jt[2] = gradx * (-(fx * yc / fy));
                jb[3] = 0;
                jt[3] = (float)(gradx * (fx * idep));
                jb[4] = (float)(grady * (fy * idep));
                jt[4] = 0;
                jb[5] = (float)(grady * (-yc * idep));
                jt[5] = (float)(gradx * (-xc * idep));
                float Jp[6], wJ[6];
                for (int i = 0; i < 6; ++i) { Jp[i] = jt[i] + jb[i]; J[i * np + idx] = Jp[i]; }     // :661-666
                for (int i = 0; i < 6; ++i) wJ[i] = Jp[i] * weight[level][idx];                       // :668-673
                // hessian = weightedSteepestDescent * steepestDescent.t() (:938): cv::gemm on CV_32F accumulates the
                // float x float products in double (GEMMSingleMul<float,double> / GEMMBlockMul) and rounds once
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) Hd[i * 6 + j] += (double)wJ[i] * (double)Jp[j];
            }
        }
        float H[36], Hinv[36];
        for (int i = 0; i < 36; ++i) H[i] = (float)Hd[i];
        ellc_oracle_invert6(H, Hinv);                                                 // :939

        int executed = 0;
        for (int iter = 0; iter < cfg->max_iter[level]; ++iter) {
            ellc_oracle_iter local;
            ellc_oracle_iter* rec = (trace && iter < ELLC_ORACLE_MAX_ITERS) ? &trace->it[level][iter] : &local;
            std::memset(rec, 0, sizeof(*rec));
            M4 T = mat_exp_f32(se3_hat(pose));                                        // :768-790
            float SE3v[12];
            for (int i = 0; i < 12; ++i) SE3v[i] = T.a[i];
            LcBand band[2];
            int nb = 1;
            if (cfg->lc_parallel) {
                const int inc = lv.rows / 3;                                          // :925 NUM_CONST_WT_POSE_EST_THREADS = 3
                lc_iterate_band(K, lv, weight[level], J.data(), SE3v, 0, inc, &band[0]);          // :945
                lc_iterate_band(K, lv, weight[level], J.data(), SE3v, inc, lv.rows, &band[1]);    // :946
                nb = 2;
            } else {
                lc_iterate_band(K, lv, weight[level], J.data(), SE3v, 0, lv.rows, &band[0]);      // :970
            }
            for (int i = 0; i < 6; ++i) { rec->b[i] = band[0].b[i]; rec->b_f64[i] = band[0].bd[i]; }
            rec->res_sum_f32 = band[0].res_f32; rec->res_sum_f64 = band[0].res_f64; rec->n_oob = band[0].n_oob;
            if (nb == 2) {
                for (int i = 0; i < 6; ++i) { rec->b[i] = rec->b[i] + band[1].b[i]; rec->b_f64[i] += band[1].bd[i]; }   // :951
                rec->res_sum_f32 += band[1].res_f32; rec->res_sum_f64 += band[1].res_f64; rec->n_oob += band[1].n_oob;
            }
            for (int i = 0; i < 36; ++i) { rec->H[i] = H[i]; rec->H_f64[i] = Hd[i]; }
            ellc_oracle_update_pose(cfg, Hinv, rec->b, pose, rec->delta, &rec->weighted_pose);       // :954
            for (int i = 0; i < 6; ++i) rec->pose_after[i] = pose[i];
            ++executed;
            if (rec->weighted_pose < cfg->stop_threshold) break;                      // src/ImageFunc.cpp:251-252
        }
        if (trace) trace->n_iters[level] = executed;
    }
    for (int i = 0; i < 6; ++i) out_pose[i] = pose[i];
    if (trace) for (int i = 0; i < 6; ++i) trace->final_pose[i] = pose[i];
}

void build_pyramid(const uint8_t* img0, int w, int h, std::vector<std::vector<uint8_t> >& store, const uint8_t* ptr[4]) {
    store.resize(3);
    ptr[0] = img0;
    int cw = w, ch = h;
    const uint8_t* src = img0;
    for (int l = 1; l < 4; ++l) {
        int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
        store[l - 1].resize((size_t)nw * nh);
        ellc_oracle_pyrdown_u8(src, cw, ch, cw, store[l - 1].data());
        ptr[l] = store[l - 1].data();
        src = ptr[l]; cw = nw; ch = nh;
    }
}

}  // namespace

// ==================================================================================================
// exported C API
// ==================================================================================================
extern "C" {

void ellc_oracle_default_config(ellc_oracle_config* c, int width, int height) {
    std::memset(c, 0, sizeof(*c));
    c->width = width; c->height = height;
    // same ratios as src/ExternVariable.h:53-59 (fx ~ 0.855*W there; the synthetic benchmark uses 0.8*W)
    c->fx = 0.8f * width; c->fy = 0.8f * width; c->cx = width / 2.0f; c->cy = height / 2.0f;
    c->max_iter[0] = 4; c->max_iter[1] = 7; c->max_iter[2] = 9; c->max_iter[3] = 12;   // src/main.cpp:34
    c->huber_d = 3.0f; c->camera_pixel_noise_2 = 4.0f * 4.0f;                          // ExternVariable.h:148-149
    c->weight[0] = c->weight[1] = c->weight[2] = 100000.0f;                            // :76
    c->weight[3] = c->weight[4] = c->weight[5] = 10000.0f;
    c->stop_threshold = 1.0f;
    c->num_bands = 3; c->use_threads = 0; c->jacobian_at_warped = 0; c->lc_parallel = 1;
}

void ellc_oracle_pyrdown_u8(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst) {
    static const int k[5] = {1, 4, 6, 4, 1};
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    std::vector<int> row((size_t)5 * dw);
    for (int y = 0; y < dh; ++y) {
        // horizontal pass of the 5 contributing source rows
        for (int i = 0; i < 5; ++i) {
            const uint8_t* s = src + (size_t)reflect101(2 * y + i - 2, h) * src_stride;
            int* r = row.data() + (size_t)i * dw;
            for (int x = 0; x < dw; ++x) {
                int acc = 0;
                for (int j = 0; j < 5; ++j) acc += k[j] * s[reflect101(2 * x + j - 2, w)];
                r[x] = acc;
            }
        }
        for (int x = 0; x < dw; ++x) {
            int acc = 0;
            for (int i = 0; i < 5; ++i) acc += k[i] * row[(size_t)i * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
        }
    }
}

void ellc_oracle_gradient(const uint8_t* img, int stride, int rows, int cols, float* gx, float* gy) {
    // src/Frame.cpp:206-283: central half-difference inside, one-sided un-halved difference on the border.
    for (int y = 0; y < rows; ++y) {
        const uint8_t* r = img + (size_t)y * stride;
        const uint8_t* up = img + (size_t)(y > 0 ? y - 1 : y) * stride;
        const uint8_t* dn = img + (size_t)(y < rows - 1 ? y + 1 : y) * stride;
        for (int x = 0; x < cols; ++x) {
            float dx;
            if (x == 0) dx = (float(r[x + 1]) - float(r[x]));
            else if (x == cols - 1) dx = (float(r[x]) - float(r[x - 1]));
            else dx = 0.5f * (float(r[x + 1]) - float(r[x - 1]));
            float dy;
            if (y == 0) dy = (float(dn[x]) - float(r[x]));
            else if (y == rows - 1) dy = (float(r[x]) - float(up[x]));
            else dy = 0.5f * (float(dn[x]) - float(up[x]));
            gx[(size_t)y * cols + x] = dx;
            gy[(size_t)y * cols + x] = dy;
        }
    }
}

int ellc_oracle_mask_count(const float* depth, int n, uint8_t* mask) {
    int c = 0;
    for (int i = 0; i < n; ++i) {
        bool sel = depth[i] > 0.0f;                 // NaN -> unselected (src/Frame.cpp:298)
        mask[i] = sel ? 255 : 0;
        c += sel;
    }
    return c;
}

float ellc_oracle_interp_u8(const uint8_t* img, int stride, int rows, int cols, float x1, float y1, int check_oob) {
    // src/Frame.h:181-279
    const int nCols = cols - 1, nRows = rows - 1;
    int oob = 0;
    const float wy = y1 - std::floor(y1);
    const float wx = x1 - std::floor(x1);
    uint8_t p1, p2;
    float y = std::floor(y1), x = std::floor(x1);
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) { p1 = 0; oob++; }
    else p1 = img[(size_t)(int)y * stride + (int)x];
    x = x1;
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) { p2 = 0; oob++; }
    else p2 = img[(size_t)(int)y * stride + (int)std::ceil(x)];
    const float top = ((1 - wx) * p1) + (wx * p2);
    y = y1; x = std::floor(x1);
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) { p1 = 0; oob++; }
    else p1 = img[(size_t)(int)std::ceil(y) * stride + (int)x];
    x = x1;
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) { p2 = 0; oob++; }
    else p2 = img[(size_t)(int)std::ceil(y) * stride + (int)std::ceil(x)];
    if (oob == 4 && check_oob == 1) return -1.0f;
    const float btm = ((1 - wx) * p1) + (wx * p2);
    return ((1 - wy) * top) + (wy * btm);
}

float ellc_oracle_interp_f32(const float* img, int stride, int rows, int cols, float x1, float y1) {
    // src/Frame.h:283-394
    const int nCols = cols - 1, nRows = rows - 1;
    const float wy = y1 - std::floor(y1);
    const float wx = x1 - std::floor(x1);
    float p1, p2;
    float y = std::floor(y1), x = std::floor(x1);
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) p1 = 0;
    else p1 = img[(size_t)(int)y * stride + (int)x];
    x = x1;
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) p2 = 0;
    else p2 = img[(size_t)(int)y * stride + (int)std::ceil(x)];
    const float top = ((1 - wx) * p1) + (wx * p2);
    y = y1; x = std::floor(x1);
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) p1 = 0;
    else p1 = img[(size_t)(int)std::ceil(y) * stride + (int)x];
    x = x1;
    if ((x < 0) || (x > nCols) || (y < 0) || (y > nRows)) p2 = 0;
    else p2 = img[(size_t)(int)std::ceil(y) * stride + (int)std::ceil(x)];
    const float btm = ((1 - wx) * p1) + (wx * p2);
    return ((1 - wy) * top) + (wy * btm);
}

void ellc_oracle_build_depth_pyramid(int w, int h, float* const* depth, float* const* var) {
    // src/DepthPropagation.cpp:1637-1719
    for (int i = 1; i < ELLC_ORACLE_LEVELS; ++i) {
        const int width = w >> i, height = h >> i, sw = 2 * width;
        const float* vs = var[i - 1]; const float* ds = depth[i - 1];
        float* vd = var[i]; float* dd = depth[i];
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) {
                const int idx = 2 * (x + y * sw);
                const int off[4] = {0, 1, sw, sw + 1};
                float idepthSum = 0, ivarSum = 0; int num = 0;
                for (int k = 0; k < 4; ++k) {
                    float v = vs[idx + off[k]];
                    if (v > 0) {
                        float ivar = 1.0f / v;
                        ivarSum += ivar;
                        idepthSum += ivar * 1.0f / ds[idx + off[k]];
                        num++;
                    }
                }
                if (num > 0) { dd[x + y * width] = ivarSum / idepthSum; vd[x + y * width] = num / ivarSum; }
                else { dd[x + y * width] = 0.0f; vd[x + y * width] = -1.0f; }
            }
    }
}

int ellc_oracle_update_depth_image(int w, int h, uint8_t* valid, const float* idepth, const float* var_s,
                                   float* depth_mat, float* depth_arr, float* var_arr, float* occupancy) {
    // depthMap::calculate_no_of_Seeds, src/DepthPropagation.cpp:1804-1830 (float counter, /(W*H)*100)
    float count = 0;
    for (int i = 0; i < w * h; ++i) count += float(valid[i] != 0);
    if (occupancy) *occupancy = count / (w * h) * 100;
    // src/DepthPropagation.cpp:1273-1300
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int i = x + y * w;
            if (y < 3 || y >= h - 3 || x < 3 || x >= w - 3) valid[i] = 0;
            if (valid[i] && idepth[i] >= -0.05f) {
                depth_mat[i] = (1 / idepth[i]);
                depth_arr[i] = (1 / idepth[i]);
                var_arr[i] = var_s[i];
            } else {
                depth_mat[i] = 0.0f;
                depth_arr[i] = -1.0f;
                var_arr[i] = -1.0f;
            }
        }
    return (int)count;
}

void ellc_oracle_image_histogram(const uint8_t* img, int n_pixels, float hist[256]) {
    for (int i = 0; i < 256; ++i) hist[i] = 0.f;
    for (int i = 0; i < n_pixels; ++i) hist[img[i]] += 1.0f;          // cv::calcHist: integer counts stored as float
    float sum = 0;
    for (int i = 0; i < 256; ++i) sum += hist[i];                     // src/GlobalOptimize.cpp:78-83
    for (int i = 0; i < 256; ++i) hist[i] /= sum;                     // :85-88
}

double ellc_oracle_hist_kl_div(const float h1[256], const float h2[256]) {
    double result = 0;
    for (int j = 0; j < 256; ++j) {
        double p = h1[j], q = h2[j];
        if (std::fabs(p) <= DBL_EPSILON) continue;
        if (std::fabs(q) <= DBL_EPSILON) q = 1e-10;
        result += p * std::log(p / q);
    }
    return result;
}

void ellc_oracle_rotation_stats(const float p1[6], const float p2[6], float* rms_error, float* relative_view_angle) {
    // src/GlobalOptimize.cpp:424-437; pow(float, int/double) promotes to double, the assignment rounds to float
    *rms_error = (float)std::pow(std::pow((double)(p1[0] - p2[0]), 2) + std::pow((double)(p1[1] - p2[1]), 2) + std::pow((double)(p1[2] - p2[2]), 2), 0.5);
    float v1[3], v2[3];
    M4 T1 = mat_exp_f32(se3_hat(p1)), T2 = mat_exp_f32(se3_hat(p2));   // calculateViewVec :439-452: third row of R
    for (int i = 0; i < 3; ++i) { v1[i] = T1.a[8 + i]; v2[i] = T2.a[8 + i]; }
    const float mag1 = (float)std::pow((double)(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]), 0.5);
    const float mag2 = (float)std::pow((double)(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]), 0.5);
    float a = std::acos((v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / (mag1 * mag2));
    a = (a * 180) / 3.14f;
    *relative_view_angle = a;
}

void ellc_oracle_se3_exp(const float pose[6], float T[16]) {
    M4 r = mat_exp_f32(se3_hat(pose));
    std::memcpy(T, r.a, sizeof(r.a));
}

void ellc_oracle_se3_log(const float T[16], float pose[6]) {
    M4 m; std::memcpy(m.a, T, sizeof(m.a));
    rigid_log(m, pose);
}

void ellc_oracle_concat_relative(const float a[6], const float b[6], float dest[6]) {
    M4 A = mat_exp_f32(se3_hat(a)), B = mat_exp_f32(se3_hat(b));
    float out[6];
    rigid_log(m4_mul(A, B), out);
    for (int i = 0; i < 6; ++i) dest[i] = out[i];
}

void ellc_oracle_concat_origin(const float a[6], const float b[6], float dest[6]) {
    M4 A = mat_exp_f32(se3_hat(a)), B = mat_exp_f32(se3_hat(b));
    float out[6];
    rigid_log(m4_mul(A, m4_inverse(B)), out);
    for (int i = 0; i < 6; ++i) dest[i] = out[i];
}

int ellc_oracle_invert6(const float Hin[36], float Hinv[36]) {
    // OpenCV 3.0.0 LUImpl<float>(A, m=6, b=I, n=6), eps = FLT_EPSILON*10; on failure dst = 0.
    const int m = 6;
    float A[36], B[36];
    std::memcpy(A, Hin, sizeof(A));
    for (int i = 0; i < 36; ++i) B[i] = (i % 7 == 0) ? 1.f : 0.f;
    const float eps = FLT_EPSILON * 10;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j)
            if (std::abs(A[j * m + i]) > std::abs(A[k * m + i])) k = j;
        if (std::abs(A[k * m + i]) < eps) { std::memset(Hinv, 0, 36 * sizeof(float)); return 0; }
        if (k != i) {
            for (int j = i; j < m; ++j) std::swap(A[i * m + j], A[k * m + j]);
            for (int j = 0; j < m; ++j) std::swap(B[i * m + j], B[k * m + j]);
        }
        float d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; ++j) {
            float alpha = A[j * m + i] * d;
            for (int c = i + 1; c < m; ++c) A[j * m + c] += alpha * A[i * m + c];
            for (int c = 0; c < m; ++c) B[j * m + c] += alpha * B[i * m + c];
        }
        A[i * m + i] = -d;
    }
    for (int i = m - 1; i >= 0; --i)
        for (int j = 0; j < m; ++j) {
            float s = B[i * m + j];
            for (int c = i + 1; c < m; ++c) s -= A[i * m + c] * B[c * m + j];
            B[i * m + j] = s * A[i * m + i];
        }
    std::memcpy(Hinv, B, sizeof(B));
    return 1;
}

void ellc_oracle_update_pose(const ellc_oracle_config* cfg, const float Hinv[36], const float b[6],
                             float pose[6], float delta[6], float* weighted_pose) {
    // src/PixelWisePyramid.cpp:466-470: cv::gemm 6x6 * 6x1 (double accumulator), transpose, negate
    for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += (double)Hinv[i * 6 + k] * (double)b[k];
        delta[i] = -(float)s;
    }
    // :479
    float wp = std::abs(delta[0] * cfg->weight[0]) + std::abs(delta[1] * cfg->weight[1]) + std::abs(delta[2] * cfg->weight[2]) +
               std::abs(delta[3] * cfg->weight[3]) + std::abs(delta[4] * cfg->weight[4]) + std::abs(delta[5] * cfg->weight[5]);
    *weighted_pose = wp;
    ellc_oracle_concat_relative(delta, pose, pose);                              // :484
}

void ellc_oracle_gn_evaluate(const ellc_oracle_config* cfg, int level,
                             const uint8_t* kf_img, int kf_stride, const uint8_t* cur_img, int cur_stride,
                             const float* depth, const float* var, const float pose[6],
                             ellc_oracle_iter* out, float* weight_img) {
    LevelView lv;
    lv.rows = (int)(cfg->height / std::pow(2.0, level));
    lv.cols = (int)(cfg->width / std::pow(2.0, level));
    lv.kf = kf_img; lv.kf_stride = kf_stride; lv.cur = cur_img; lv.cur_stride = cur_stride;
    lv.depth = depth; lv.var = var;
    std::vector<uint8_t> mask; int count = 0;
    build_level_mask(depth, lv.rows * lv.cols, mask, &count);
    std::vector<float> gx((size_t)lv.rows * lv.cols), gy((size_t)lv.rows * lv.cols);
    ellc_oracle_gradient(cur_img, cur_stride, lv.rows, lv.cols, gx.data(), gy.data());
    lv.gx = gx.data(); lv.gy = gy.data(); lv.mask = mask.data();
    std::memset(out, 0, sizeof(*out));
    evaluate_level(cfg, level, lv, pose, out, weight_img);
}

void ellc_oracle_track_prebuilt(const ellc_oracle_config* cfg, const uint8_t* const* kf_pyr, const uint8_t* const* cur_pyr,
                                const float* const* depth, const float* const* var,
                                const float init_pose[6], float out_pose[6], ellc_oracle_trace* trace) {
    track_impl(cfg, kf_pyr, cur_pyr, depth, var, init_pose, out_pose, trace);
}

void ellc_oracle_track_weights_prebuilt(const ellc_oracle_config* cfg, const uint8_t* const* kf_pyr, const uint8_t* const* cur_pyr,
                                        const float* const* depth, const float* const* var, const float init_pose[6],
                                        float out_pose[6], ellc_oracle_trace* trace, float* const* weight_last) {
    track_impl(cfg, kf_pyr, cur_pyr, depth, var, init_pose, out_pose, trace, weight_last);
}

void ellc_oracle_track_lc_prebuilt(const ellc_oracle_config* cfg, const uint8_t* const* kf_pyr, const uint8_t* const* cur_pyr,
                                   const float* const* depth, const float* const* weight, const float init_pose[6],
                                   float out_pose[6], ellc_oracle_trace* trace) {
    track_lc_impl(cfg, kf_pyr, cur_pyr, depth, weight, init_pose, out_pose, trace);
}

void ellc_oracle_track(const ellc_oracle_config* cfg, const uint8_t* kf_img0, const uint8_t* cur_img0,
                       const float* const* depth, const float* const* var,
                       const float init_pose[6], float out_pose[6], ellc_oracle_trace* trace) {
    std::vector<std::vector<uint8_t> > ks, cs;
    const uint8_t* kp[4]; const uint8_t* cp[4];
    build_pyramid(kf_img0, cfg->width, cfg->height, ks, kp);       // src/Frame.cpp:170-182
    build_pyramid(cur_img0, cfg->width, cfg->height, cs, cp);
    track_impl(cfg, kp, cp, depth, var, init_pose, out_pose, trace);
}

double ellc_oracle_track_many(const ellc_oracle_config* cfg, int n_pairs, int n_workers,
                              const int* kf_idx, const int* fr_idx,
                              const uint8_t* const* kf_imgs, const uint8_t* const* fr_imgs,
                              const float* const* kf_depth, const float* const* kf_var,
                              const float* init_poses, float* out_poses) {
    // keyframe image pyramids are built once per keyframe at frame construction in the reference
    // (src/Frame.cpp:103) and are cached on the GPU side too -> prebuilt, untimed.
    int n_kf = 0;
    for (int i = 0; i < n_pairs; ++i) n_kf = std::max(n_kf, kf_idx[i] + 1);
    std::vector<std::vector<std::vector<uint8_t> > > kstore(n_kf);
    std::vector<const uint8_t*> kptr((size_t)n_kf * 4, nullptr);
    std::vector<char> have(n_kf, 0);
    for (int i = 0; i < n_pairs; ++i) {
        int k = kf_idx[i];
        if (!have[k]) { build_pyramid(kf_imgs[k], cfg->width, cfg->height, kstore[k], &kptr[(size_t)k * 4]); have[k] = 1; }
    }
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_pairs) break;
            std::vector<std::vector<uint8_t> > cs; const uint8_t* cp[4];
            build_pyramid(fr_imgs[fr_idx[i]], cfg->width, cfg->height, cs, cp);
            int k = kf_idx[i];
            track_impl(cfg, &kptr[(size_t)k * 4], cp, &kf_depth[(size_t)k * 4], &kf_var[(size_t)k * 4],
                       init_poses + (size_t)i * 6, out_poses + (size_t)i * 6, nullptr);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (n_workers <= 1) worker();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_workers; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
