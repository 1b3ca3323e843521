// ellc_lie.cuh -- SE(3) pose algebra and the 6x6 solve, shared by the device kernels and the host-side pose helpers.
//
// Replaces (citations relative to the reference repo):
//   - Eigen .exp() on hat(pose)            src/PixelWisePyramid.cpp:153-159   -> closed-form Rodrigues, evaluated in double
//   - concatenateRelativePose               src/Frame.cpp:503-530             -> log(exp(a) exp(b)), double, rounded to f32
//   - concatenateOriginPose                 src/Frame.cpp:534-562             -> log(exp(a) exp(b)^-1)
//   - cv::Mat::inv() (DECOMP_LU, CV_32F)    src/PixelWisePyramid.cpp:451      -> same partial-pivot LU in fp32, eps 10*FLT_EPSILON,
//                                                                               singular => all-zero inverse (zero step)
//   - updatePose()                          src/PixelWisePyramid.cpp:460-491
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace ellc {

#ifdef __CUDA_ARCH__
#define ELLC_MUL(a, b) __fmul_rn((a), (b))
#define ELLC_ADD(a, b) __fadd_rn((a), (b))
#define ELLC_SUB(a, b) __fsub_rn((a), (b))
#define ELLC_DIV(a, b) __fdiv_rn((a), (b))
#define ELLC_DMUL(a, b) __dmul_rn((a), (b))
#define ELLC_DADD(a, b) __dadd_rn((a), (b))
#else
#define ELLC_MUL(a, b) ((a) * (b))
#define ELLC_ADD(a, b) ((a) + (b))
#define ELLC_SUB(a, b) ((a) - (b))
#define ELLC_DIV(a, b) ((a) / (b))
#define ELLC_DMUL(a, b) ((a) * (b))
#define ELLC_DADD(a, b) ((a) + (b))
#endif

// exp(hat(p)) -> R (row-major 3x3) and t, in double.
__host__ __device__ inline void se3_exp_d(const double p[6], double R[9], double t[3]) {
    const double wx = p[0], wy = p[1], wz = p[2];
    const double th2 = wx * wx + wy * wy + wz * wz;
    double A, B, C;
    if (th2 < 1e-12) {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
        C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        const double th = sqrt(th2);
        double s, c;
#ifdef __CUDA_ARCH__
        sincos(th, &s, &c);
#else
        s = sin(th); c = cos(th);
#endif
        A = s / th;
        B = (1.0 - c) / th2;
        C = (th - s) / (th2 * th);
    }
    // W = hat3(w), W2 = W*W
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    const double W2[9] = {-(wy * wy + wz * wz), wx * wy, wx * wz,
                          wx * wy, -(wx * wx + wz * wz), wy * wz,
                          wx * wz, wy * wz, -(wx * wx + wy * wy)};
    double V[9];
    for (int i = 0; i < 9; ++i) {
        const double id = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = id + A * W[i] + B * W2[i];
        V[i] = id + B * W[i] + C * W2[i];
    }
    for (int i = 0; i < 3; ++i) t[i] = V[i * 3 + 0] * p[3] + V[i * 3 + 1] * p[4] + V[i * 3 + 2] * p[5];
}

// log of a rigid transform (R row-major, t) -> 6-vector, in double.  Entry extraction as src/Frame.cpp:523-528.
__host__ __device__ inline void se3_log_d(const double R[9], const double t[3], double out[6]) {
    const double ax = 0.5 * (R[7] - R[5]), ay = 0.5 * (R[2] - R[6]), az = 0.5 * (R[3] - R[1]);   // sin(th) n
    const double s = sqrt(ax * ax + ay * ay + az * az);
    const double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    const double th = atan2(s, c);
    double k;
    if (s < 1e-7) k = (c > 0) ? 1.0 + th * th / 6.0 : 0.0;
    else k = th / s;
    const double w[3] = {k * ax, k * ay, k * az};
    double coef;
    if (th < 1e-4) coef = 1.0 / 12.0 + th * th / 720.0;
    else coef = (1.0 - (th * sin(th)) / (2.0 * (1.0 - cos(th)))) / (th * th);
    const double wt[3] = {w[1] * t[2] - w[2] * t[1], w[2] * t[0] - w[0] * t[2], w[0] * t[1] - w[1] * t[0]};
    const double wwt[3] = {w[1] * wt[2] - w[2] * wt[1], w[2] * wt[0] - w[0] * wt[2], w[0] * wt[1] - w[1] * wt[0]};
    out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
    for (int i = 0; i < 3; ++i) out[3 + i] = t[i] - 0.5 * wt[i] + coef * wwt[i];
}

// dest = log(exp(a) exp(b))   -- frame::concatenateRelativePose, src/Frame.cpp:503-530
__host__ __device__ inline void concat_relative_f(const float a[6], const float b[6], float dest[6]) {
    double pa[6], pb[6], Ra[9], ta[3], Rb[9], tb[3], R[9], t[3], o[6];
    for (int i = 0; i < 6; ++i) { pa[i] = a[i]; pb[i] = b[i]; }
    se3_exp_d(pa, Ra, ta);
    se3_exp_d(pb, Rb, tb);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = Ra[i * 3 + 0] * Rb[0 * 3 + j] + Ra[i * 3 + 1] * Rb[1 * 3 + j] + Ra[i * 3 + 2] * Rb[2 * 3 + j];
        t[i] = Ra[i * 3 + 0] * tb[0] + Ra[i * 3 + 1] * tb[1] + Ra[i * 3 + 2] * tb[2] + ta[i];
    }
    se3_log_d(R, t, o);
    for (int i = 0; i < 6; ++i) dest[i] = (float)o[i];
}

// dest = log(exp(a) exp(b)^-1) -- frame::concatenateOriginPose, src/Frame.cpp:534-562
__host__ __device__ inline void concat_origin_f(const float a[6], const float b[6], float dest[6]) {
    double pa[6], pb[6], Ra[9], ta[3], Rb[9], tb[3], R[9], t[3], o[6];
    for (int i = 0; i < 6; ++i) { pa[i] = a[i]; pb[i] = b[i]; }
    se3_exp_d(pa, Ra, ta);
    se3_exp_d(pb, Rb, tb);
    // inv(Tb) = [Rb^T, -Rb^T tb]
    double Ri[9], ti[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Ri[i * 3 + j] = Rb[j * 3 + i];
    }
    for (int i = 0; i < 3; ++i) ti[i] = -(Ri[i * 3 + 0] * tb[0] + Ri[i * 3 + 1] * tb[1] + Ri[i * 3 + 2] * tb[2]);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = Ra[i * 3 + 0] * Ri[0 * 3 + j] + Ra[i * 3 + 1] * Ri[1 * 3 + j] + Ra[i * 3 + 2] * Ri[2 * 3 + j];
        t[i] = Ra[i * 3 + 0] * ti[0] + Ra[i * 3 + 1] * ti[1] + Ra[i * 3 + 2] * ti[2] + ta[i];
    }
    se3_log_d(R, t, o);
    for (int i = 0; i < 6; ++i) dest[i] = (float)o[i];
}

// exp(hat(pose)) rounded to fp32: SE3_vec[12] = r11 r12 r13 t1 r21 ... (src/PixelWisePyramid.cpp:162-173)
__host__ __device__ inline void pose_to_rt_f(const float pose[6], float Rt[12]) {
    double p[6], R[9], t[3];
    for (int i = 0; i < 6; ++i) p[i] = pose[i];
    se3_exp_d(p, R, t);
    for (int i = 0; i < 3; ++i) {
        Rt[i * 4 + 0] = (float)R[i * 3 + 0];
        Rt[i * 4 + 1] = (float)R[i * 3 + 1];
        Rt[i * 4 + 2] = (float)R[i * 3 + 2];
        Rt[i * 4 + 3] = (float)t[i];
    }
}

// cv::Mat::inv() for a 6x6 CV_32F: partial-pivot LU on [A | I] in fp32 (each op individually rounded, as OpenCV's
// scalar LUImpl<float>), pivot threshold FLT_EPSILON*10, back-substitution multiplying by the stored reciprocal.
// Returns false and an all-zero inverse when singular (=> deltapose = 0, weightedPose = 0 < 1 => level ends).
__host__ __device__ inline bool invert6_lu_f(const float Hin[36], float Hinv[36]) {
    const int m = 6;
    float A[36], B[36];
    for (int i = 0; i < 36; ++i) { A[i] = Hin[i]; B[i] = (i % 7 == 0) ? 1.f : 0.f; }
    const float eps = 1.1920929e-07f * 10;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j)
            if (fabsf(A[j * m + i]) > fabsf(A[k * m + i])) k = j;
        if (fabsf(A[k * m + i]) < eps) {
            for (int q = 0; q < 36; ++q) Hinv[q] = 0.f;
            return false;
        }
        if (k != i) {
            for (int j = i; j < m; ++j) { float tmp = A[i * m + j]; A[i * m + j] = A[k * m + j]; A[k * m + j] = tmp; }
            for (int j = 0; j < m; ++j) { float tmp = B[i * m + j]; B[i * m + j] = B[k * m + j]; B[k * m + j] = tmp; }
        }
        const float d = ELLC_DIV(-1.f, A[i * m + i]);
        for (int j = i + 1; j < m; ++j) {
            const float alpha = ELLC_MUL(A[j * m + i], d);
            for (int c = i + 1; c < m; ++c) A[j * m + c] = ELLC_ADD(A[j * m + c], ELLC_MUL(alpha, A[i * m + c]));
            for (int c = 0; c < m; ++c) B[j * m + c] = ELLC_ADD(B[j * m + c], ELLC_MUL(alpha, B[i * m + c]));
        }
        A[i * m + i] = -d;
    }
    for (int i = m - 1; i >= 0; --i)
        for (int j = 0; j < m; ++j) {
            float s = B[i * m + j];
            for (int c = i + 1; c < m; ++c) s = ELLC_SUB(s, ELLC_MUL(A[i * m + c], B[c * m + j]));
            B[i * m + j] = ELLC_MUL(s, A[i * m + i]);
        }
    for (int q = 0; q < 36; ++q) Hinv[q] = B[q];
    return true;
}

// hessianInv = hessian.inv(); updatePose()  -- src/PixelWisePyramid.cpp:451-491.
// delta_i = -(sum_k Hinv[i][k] b[k]) with cv::gemm's double accumulator; weightedPose = sum |delta_i * weight_i|;
// pose <- log(exp(delta) exp(pose)).  Returns false if the hessian was singular.
__host__ __device__ inline bool solve_update_f(const float H[36], const float b[6], const float weight[6],
                                               float pose[6], float delta[6], float* weighted_pose) {
    float Hinv[36];
    const bool ok = invert6_lu_f(H, Hinv);
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s = ELLC_DADD(s, ELLC_DMUL((double)Hinv[i * 6 + k], (double)b[k]));
        delta[i] = -(float)s;
    }
    float wp = fabsf(ELLC_MUL(delta[0], weight[0]));
    for (int i = 1; i < 6; ++i) wp = ELLC_ADD(wp, fabsf(ELLC_MUL(delta[i], weight[i])));
    *weighted_pose = wp;
    float np[6];
    concat_relative_f(delta, pose, np);
    for (int i = 0; i < 6; ++i) pose[i] = np[i];
    return ok;
}

}  // namespace ellc
