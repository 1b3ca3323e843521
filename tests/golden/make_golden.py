"""Generates the committed golden fixtures under tests/golden/ (run in the BUILD container, which has cv2 + scipy).

The reference ships no tests or golden vectors and cannot be compiled here (needs OpenCV/Eigen/Boost C++), so the
oracle's third-party arithmetic is pinned against the libraries that ARE available:
  - cv2.pyrDown            (bit-exact restatement target; reference call src/Frame.cpp:175-179)
  - cv2.invert DECOMP_LU   (reference call src/PixelWisePyramid.cpp:451; cv2 4.13's SIMD build differs from OpenCV
                            3.0.0's scalar loop in the last bits, so this is a tolerance fixture)
  - scipy.linalg.expm/logm (Eigen .exp()/.log(), src/Frame.cpp:511-521), float64
plus one end-to-end oracle trace on a tiny synthetic pair (regression pin for oracle AND GPU).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def hat(p):
    return np.array([[0, -p[2], p[1], p[3]], [p[2], 0, -p[0], p[4]], [-p[1], p[0], 0, p[5]], [0, 0, 0, 0]], np.float64)


def main():
    import cv2
    import scipy.linalg as sl

    import oracle
    from egomotion_with_local_loop_closures_b200 import synth

    rng = np.random.default_rng(20240917)
    out = {}
    # --- pyrDown
    for i, (h, w) in enumerate([(48, 64), (29, 37), (33, 61), (16, 16), (135, 240)]):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"pyr_in_{i}"] = img
        out[f"pyr_out_{i}"] = cv2.pyrDown(img)
    # --- 6x6 LU inverse
    Hs, His = [], []
    for i in range(8):
        J = (rng.standard_normal((300, 6)) * np.array([800, 800, 800, 90, 90, 90])).astype(np.float32)
        H = (J.T @ J).astype(np.float32)
        ok, Hi = cv2.invert(H, flags=cv2.DECOMP_LU)
        assert ok != 0
        Hs.append(H); His.append(Hi)
    out["inv_in"] = np.stack(Hs); out["inv_out"] = np.stack(His)
    ok, z = cv2.invert(np.zeros((6, 6), np.float32), flags=cv2.DECOMP_LU)
    out["inv_singular_ok"] = np.array([ok]); out["inv_singular_out"] = z
    # --- expm / logm
    poses = np.concatenate([rng.standard_normal((12, 6)) * np.array([0.02, 0.02, 0.02, 0.05, 0.05, 0.05]),
                            rng.standard_normal((6, 6)) * np.array([0.4, 0.4, 0.4, 0.5, 0.5, 0.5]),
                            rng.standard_normal((4, 6)) * np.array([1.2, 1.2, 1.2, 2.0, 2.0, 2.0]),
                            np.zeros((1, 6))]).astype(np.float32)
    out["se3_poses"] = poses
    out["se3_expm"] = np.stack([sl.expm(hat(p.astype(np.float64))) for p in poses])
    a, b = poses[:10], poses[5:15]
    out["concat_rel"] = np.stack([np.real(sl.logm(sl.expm(hat(x.astype(np.float64))) @ sl.expm(hat(y.astype(np.float64)))))[[2, 0, 1, 0, 1, 2], [1, 2, 0, 3, 3, 3]] for x, y in zip(a, b)])
    out["concat_org"] = np.stack([np.real(sl.logm(sl.expm(hat(x.astype(np.float64))) @ np.linalg.inv(sl.expm(hat(y.astype(np.float64))))))[[2, 0, 1, 0, 1, 2], [1, 2, 0, 3, 3, 3]] for x, y in zip(a, b)])
    np.savez_compressed(os.path.join(HERE, "library_vectors.npz"), **out)

    # --- tiny end-to-end oracle trace (inputs + outputs committed)
    w, h = 160, 120
    scene = synth.SynthScene(w, h, wavelength_px=28.0)
    kf = scene.keyframe(noise_seed=3)
    gt = np.array([0.004, -0.006, 0.003, 0.006, -0.004, 0.005], np.float32)
    cur = scene.render(synth.se3_exp(gt), noise_seed=4)
    k = synth.intrinsics(w, h)
    cfg = oracle.default_config(w, h, fx=float(k["fx"]), fy=float(k["fy"]), cx=float(k["cx"]), cy=float(k["cy"]))
    pose, tr = oracle.track(cfg, kf["image"], cur, kf["depth"], kf["var"], np.zeros(6, np.float32))
    e2e = dict(width=np.array([w]), height=np.array([h]), kf_image=kf["image"], cur_image=cur, gt=gt, pose=pose,
               n_selected=np.array(tr["n_selected"]), n_iters=np.array(tr["n_iters"]))
    for l in range(4):
        e2e[f"depth{l}"] = kf["depth"][l]; e2e[f"var{l}"] = kf["var"][l]
        e2e[f"res_f64_{l}"] = np.array([it["res_sum_f64"] for it in tr["levels"][l]])
        e2e[f"wp_{l}"] = np.array([it["weighted_pose"] for it in tr["levels"][l]])
        e2e[f"pose_after_{l}"] = np.stack([it["pose_after"] for it in tr["levels"][l]])
        e2e[f"H_{l}"] = np.stack([it["H"] for it in tr["levels"][l]]); e2e[f"b_{l}"] = np.stack([it["b"] for it in tr["levels"][l]])
    np.savez_compressed(os.path.join(HERE, "oracle_track_160x120.npz"), **e2e)
    print("wrote", os.listdir(HERE))


def main_gating():
    """Loop-closure gating (src/GlobalOptimize.cpp:40-122, :424-452): cv2.calcHist / cv2.compareHist(KL_DIV) vectors and
    float64 view angles; written to their own file so that the older fixtures stay byte-identical."""
    import cv2
    import scipy.linalg as sl
    rng = np.random.default_rng(20261018)
    base = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    imgs = [base, np.clip(base.astype(int) + rng.integers(-25, 26, base.shape), 0, 255).astype(np.uint8),
            (base // 2).astype(np.uint8), np.full((60, 80), 77, np.uint8), rng.integers(100, 140, (60, 80), dtype=np.uint8)]
    hists = []
    for im in imgs:
        hst = cv2.calcHist([im], [0], None, [256], [0, 256]).ravel().astype(np.float32)
        s_ = np.float32(0)
        for v in hst:
            s_ = np.float32(s_ + v)
        hists.append((hst / s_).astype(np.float32))
    kl = np.array([[cv2.compareHist(a.reshape(-1, 1), b.reshape(-1, 1), cv2.HISTCMP_KL_DIV) for b in hists] for a in hists], np.float64)
    poses = (rng.standard_normal((8, 6)) * np.array([0.08, 0.08, 0.08, 0.3, 0.3, 0.3])).astype(np.float32)
    ang = np.zeros((8, 8)); rms = np.zeros((8, 8))
    for i in range(8):
        for j in range(8):
            v1 = sl.expm(hat(poses[i].astype(np.float64)))[2, :3]; v2 = sl.expm(hat(poses[j].astype(np.float64)))[2, :3]
            ang[i, j] = np.degrees(np.arccos(np.clip(v1 @ v2 / (np.linalg.norm(v1) * np.linalg.norm(v2)), -1, 1))) * (np.pi / 3.14)
            rms[i, j] = np.linalg.norm(poses[i, :3].astype(np.float64) - poses[j, :3].astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "gating_vectors.npz"), images=np.stack(imgs), hists=np.stack(hists), kl=kl, poses=poses,
                        view_angle_deg=ang, rms=rms)
    print("wrote gating_vectors.npz")


def main_matops():
    """Semantics of the cv::Mat expressions the tracker relies on, recorded from cv2 (own file; older fixtures stay byte-identical):
      * gemm on CV_32F (hessian = weightedSteepestDescent * steepestDescent.t(), src/PixelWisePyramid.cpp:938; the 6x1 * 1x6
        products :373; hessianInv * sd_param.t() :466): double accumulator, rounded to float once;
      * Mat / scalar (frame::finaliseWeights, src/Frame.cpp:688) is MatOp_AddEx with alpha = 1./s evaluated by convertTo, whose
        CV_32F kernel multiplies by (float)alpha -- observed here through cv2.normalize(NORM_MINMAX), which calls
        src.convertTo(dst, CV_32F, scale, shift) with scale = alpha when the input spans exactly [0, 1]."""
    import cv2
    rng = np.random.default_rng(20261019)
    out = {}
    for tag, n in (("small", 7), ("large", 4000)):
        A = (rng.standard_normal((6, n)) * np.array([[800, 800, 800, 90, 90, 90]]).T).astype(np.float32)
        B = rng.standard_normal((n, 6)).astype(np.float32)
        out[f"gemm_A_{tag}"], out[f"gemm_B_{tag}"] = A, B
        out[f"gemm_C_{tag}"] = cv2.gemm(A, B, 1.0, None, 0.0)
    x = rng.uniform(0, 1, (1, 3000)).astype(np.float32)
    x[0, 0], x[0, 1] = 0.0, 1.0
    out["scale_x"] = x
    for n in (3, 5, 6, 7, 8):
        out[f"scale_y_{n}"] = cv2.normalize(x, None, alpha=1.0 / n, beta=0.0, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    np.savez_compressed(os.path.join(HERE, "matop_vectors.npz"), **out)
    print("wrote matop_vectors.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gating":
        main_gating()
    elif len(sys.argv) > 1 and sys.argv[1] == "matops":
        main_matops()
    else:
        main()
        main_gating()
        main_matops()
