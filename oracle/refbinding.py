"""ctypes binding of oracle/_ref/libellc_ref.so: the REFERENCE'S OWN tracking sources (src/Frame.cpp, PixelWisePyramid.cpp,
Pyramid.cpp, UserDefinedFunc.cpp, ImageFunc.cpp, compiled where they lie under /root/reference against the stand-in headers
of oracle/shim/; recipe: oracle/Makefile, driver: oracle/ref_driver.cpp).  Test infrastructure only: it pins the oracle
restatement on the reference's own per-pixel code and driver.  The library is git-ignored and travels to the GPU box as a
built file; /root/reference is needed only to (re)build it."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = "/root/reference/src"
LEVELS = 4
MAX_ITERS = 12


# Two builds of the same sources: the reference's compiled-in camera (480x270) and the metric resolution (640x480, generated
# ExternVariable.h: see the Makefile).  select() switches the library every function of this module talks to.
_VARIANTS = {"default": ("libellc_ref.so", "ref"), "640x480": ("libellc_ref_640x480.so", "ref640")}
_variant = "default"


def _so(variant=None):
    return os.path.join(_HERE, "_ref", _VARIANTS[variant or _variant][0])


def select(variant="default"):
    """Use the build `variant` ("default" = 480x270 as shipped, "640x480") from now on; returns the previous selection."""
    global _variant
    assert variant in _VARIANTS
    prev, _variant = _variant, variant
    return prev


def available(variant=None):
    """True if the library exists or can be built here (the reference sources are mounted)."""
    return os.path.exists(_so(variant)) or os.path.isdir(REFERENCE_SRC)


def build(force=False, variant=None):
    so = _so(variant)
    if not os.path.isdir(REFERENCE_SRC):
        if os.path.exists(so):
            return so
        raise RuntimeError(so + " is missing and /root/reference is not mounted")
    deps = [os.path.join(_HERE, "ref_driver.cpp"), os.path.join(_HERE, "Makefile")]
    for root, _, files in os.walk(os.path.join(_HERE, "shim")):
        deps += [os.path.join(root, f) for f in files]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, _VARIANTS[variant or _variant][1]], stdout=subprocess.DEVNULL)
    return so


class Iter(C.Structure):
    _fields_ = [("H", C.c_float * 36), ("b", C.c_float * 6), ("weighted_pose", C.c_float), ("pose_after", C.c_float * 6),
                ("pad_", C.c_float)]


class Trace(C.Structure):
    _fields_ = [("n_selected", C.c_int * LEVELS), ("n_iters", C.c_int * LEVELS), ("it", (Iter * MAX_ITERS) * LEVELS),
                ("final_pose", C.c_float * 6)]


_libs = {}


def lib():
    if _variant not in _libs:
        _libs[_variant] = C.CDLL(build())
    return _libs[_variant]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f6(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(6))


def dims():
    """The reference's compile-time size and intrinsics (src/ExternVariable.h:39-59)."""
    w, h = C.c_int(), C.c_int()
    fx, fy, cx, cy = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    lib().ellc_ref_dims(C.byref(w), C.byref(h), C.byref(fx), C.byref(fy), C.byref(cx), C.byref(cy))
    return dict(width=w.value, height=h.value, fx=np.float32(fx.value), fy=np.float32(fy.value), cx=np.float32(cx.value),
                cy=np.float32(cy.value))


def frame_level(image, level, depth_level=None):
    """frame construction + updationOnPyrChange(level): pyramid image, gradients, (mask, count) if a depth level is given."""
    d = dims()
    image = np.ascontiguousarray(image, np.uint8)
    assert image.shape == (d["height"], d["width"])
    r, c = d["height"] >> level, d["width"] >> level
    pyr = np.zeros((d["height"], d["width"]), np.uint8)
    pw, ph, cnt = C.c_int(), C.c_int(), C.c_int()
    gx = np.zeros((r, c), np.float32); gy = np.zeros((r, c), np.float32); mask = np.zeros((r, c), np.uint8)
    dl = None if depth_level is None else np.ascontiguousarray(depth_level, np.float32)
    lib().ellc_ref_frame_level(_p(image), level, _p(dl), _p(pyr), C.byref(pw), C.byref(ph), _p(gx), _p(gy), _p(mask), C.byref(cnt))
    out = dict(image=pyr.reshape(-1)[:pw.value * ph.value].reshape(ph.value, pw.value).copy(), gradx=gx, grady=gy)
    if dl is not None:
        out.update(mask=mask, count=cnt.value)
    return out


def interpolate(image, level, xs, ys):
    xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
    image = np.ascontiguousarray(image, np.uint8)
    n = len(xs)
    i = np.zeros(n, np.float32); gx = np.zeros(n, np.float32); gy = np.zeros(n, np.float32)
    lib().ellc_ref_interpolate(_p(image), level, n, _p(xs), _p(ys), _p(i), _p(gx), _p(gy))
    return i, gx, gy


def concat_relative(a, b):
    out = np.empty(6, np.float32)
    lib().ellc_ref_concat_relative(_p(_f6(a)), _p(_f6(b)), _p(out))
    return out


def concat_origin(a, b):
    out = np.empty(6, np.float32)
    lib().ellc_ref_concat_origin(_p(_f6(a)), _p(_f6(b)), _p(out))
    return out


def _pyr_ptrs(arrs):
    a = [np.ascontiguousarray(x, np.float32) for x in arrs]
    return a, (C.c_void_p * LEVELS)(*[_p(x) for x in a])


def get_image_pose_estimate(kf_image, cur_image, depth_pyr, var_pyr, tminus1_pose_wrt_world):
    """The reference's own driver (src/ImageFunc.cpp:49-315).  Returns (relative pose, poseWrtOrigin, poseWrtWorld)."""
    kf_image = np.ascontiguousarray(kf_image, np.uint8); cur_image = np.ascontiguousarray(cur_image, np.uint8)
    d, dp = _pyr_ptrs(depth_pyr); v, vp = _pyr_ptrs(var_pyr)
    pose = np.empty(6, np.float32); po = np.empty(6, np.float32); pw = np.empty(6, np.float32)
    lib().ellc_ref_get_image_pose_estimate(_p(kf_image), _p(cur_image), dp, vp, _p(_f6(tminus1_pose_wrt_world)), _p(pose), _p(po), _p(pw))
    return pose, po, pw


def track_trace(kf_image, cur_image, depth_pyr, var_pyr, init_pose, want_weights=False):
    """Level / iteration schedule of src/ImageFunc.cpp:150-299 on the reference's PixelWisePyramid object, recording its public
    per-iteration state (hessian, sd_param, weightedPose, pose)."""
    kf_image = np.ascontiguousarray(kf_image, np.uint8); cur_image = np.ascontiguousarray(cur_image, np.uint8)
    d, dp = _pyr_ptrs(depth_pyr); v, vp = _pyr_ptrs(var_pyr)
    tr = Trace()
    dm = dims()
    w = np.zeros((dm["height"], dm["width"]), np.float32) if want_weights else None
    lib().ellc_ref_track_trace(_p(kf_image), _p(cur_image), dp, vp, _p(_f6(init_pose)), C.byref(tr), _p(w))
    levels = []
    for l in range(LEVELS):
        its = []
        for k in range(tr.n_iters[l]):
            it = tr.it[l][k]
            its.append(dict(H=np.array(it.H, np.float32).reshape(6, 6), b=np.array(it.b, np.float32),
                            weighted_pose=np.float32(it.weighted_pose), pose_after=np.array(it.pose_after, np.float32)))
        levels.append(its)
    out = dict(n_selected=list(tr.n_selected), n_iters=list(tr.n_iters), levels=levels, final_pose=np.array(tr.final_pose, np.float32))
    if want_weights:
        out["weights_l0"] = w
    return out


def _trace_dict(tr):
    levels = []
    for l in range(LEVELS):
        its = []
        for k in range(tr.n_iters[l]):
            it = tr.it[l][k]
            its.append(dict(H=np.array(it.H, np.float32).reshape(6, 6), b=np.array(it.b, np.float32),
                            weighted_pose=np.float32(it.weighted_pose), pose_after=np.array(it.pose_after, np.float32)))
        levels.append(its)
    return dict(n_selected=list(tr.n_selected), n_iters=list(tr.n_iters), levels=levels, final_pose=np.array(tr.final_pose, np.float32))


def lc_flow(kf_image, seq_images, seq_tminus1, depth_pyr, var_pyr, lc_image, lc_tminus1, parallel=True):
    """The reference's constant-weight loop-closure flow with its own driver: sequential tracks saving weights,
    frame::finaliseWeights, then GetImagePoseEstimate(fromLoopClosure=true).  Returns dict(weights, counts, seq_poses, lc_pose,
    lc_trace)."""
    dm = dims()
    kf_image = np.ascontiguousarray(kf_image, np.uint8)
    seq = [np.ascontiguousarray(a, np.uint8) for a in seq_images]
    sp = (C.c_void_p * len(seq))(*[_p(a) for a in seq])
    tm1 = np.ascontiguousarray(np.asarray(seq_tminus1, np.float32).reshape(len(seq), 6))
    d, dp = _pyr_ptrs(depth_pyr); v, vp = _pyr_ptrs(var_pyr)
    lc_image = np.ascontiguousarray(lc_image, np.uint8)
    w = [np.zeros((dm["height"] >> l, dm["width"] >> l), np.float32) for l in range(LEVELS)]
    wp = (C.c_void_p * LEVELS)(*[_p(a) for a in w])
    counts = (C.c_int * LEVELS)()
    seq_poses = np.zeros((len(seq), 6), np.float32)
    lc_pose = np.zeros(6, np.float32)
    tr = Trace()
    lib().ellc_ref_lc_flow(_p(kf_image), len(seq), sp, _p(tm1), dp, vp, _p(lc_image), _p(_f6(lc_tminus1)), int(bool(parallel)), wp, counts,
                           _p(seq_poses), _p(lc_pose), C.byref(tr))
    return dict(weights=w, counts=list(counts), seq_poses=seq_poses, lc_pose=lc_pose, lc_trace=_trace_dict(tr))


def update_depth_image(kf_image, valid, inv_depth_smoothed, variance_smoothed):
    """depthMap::updateDepthImage + buildInvVarDepth + mapDepthArr2Mat + calculate_no_of_Seeds (src/DepthPropagation.cpp)."""
    dm = dims()
    h, w = dm["height"], dm["width"]
    kf_image = np.ascontiguousarray(kf_image, np.uint8)
    valid = np.ascontiguousarray(valid, np.uint8); idep = np.ascontiguousarray(inv_depth_smoothed, np.float32)
    vs = np.ascontiguousarray(variance_smoothed, np.float32)
    mk = lambda: [np.zeros((h >> l, w >> l), np.float32) for l in range(LEVELS)]
    dmat, darr, varr = mk(), mk(), mk()
    ptrs = lambda a: (C.c_void_p * LEVELS)(*[_p(x) for x in a])
    vout = np.zeros((h, w), np.uint8)
    occ = C.c_float()
    lib().ellc_ref_update_depth_image(_p(kf_image), _p(valid), _p(idep), _p(vs), ptrs(dmat), ptrs(darr), ptrs(varr), _p(vout), C.byref(occ))
    return dict(valid_out=vout, depth=dmat, depth_arr=darr, var=varr, occupancy=occ.value)


def gating(image_a, image_b, pose_a, pose_b):
    """calculateImageHistogram / compareImageHistogram / calculateRotationStats (src/GlobalOptimize.cpp:40-122, :419-452)."""
    ha = np.zeros(256, np.float32); hb = np.zeros(256, np.float32)
    kl = C.c_double(); rms = C.c_float(); ang = C.c_float()
    lib().ellc_ref_gating(_p(np.ascontiguousarray(image_a, np.uint8)), _p(np.ascontiguousarray(image_b, np.uint8)), _p(_f6(pose_a)),
                          _p(_f6(pose_b)), _p(ha), _p(hb), C.byref(kl), C.byref(rms), C.byref(ang))
    return dict(hist_a=ha, hist_b=hb, kl=kl.value, rms_error=rms.value, relative_view_angle=ang.value)


def pyramid_run(kf_image, cur_image, depth_pyr, var_pyr, level, pose, iters=1):
    """src/Pyramid.cpp driven as its comments describe: performPrecomputation at `pose`, then `iters` performIterationSteps."""
    dm = dims()
    kf_image = np.ascontiguousarray(kf_image, np.uint8); cur_image = np.ascontiguousarray(cur_image, np.uint8)
    d, dp = _pyr_ptrs(depth_pyr); v, vp = _pyr_ptrs(var_pyr)
    cap = (dm["height"] >> level) * (dm["width"] >> level)
    w = np.zeros(cap, np.float32); r = np.zeros(cap, np.float32)
    n = C.c_int(); le = C.c_float()
    hinv = np.zeros(36, np.float32); poses = np.zeros((iters, 6), np.float32); ratios = np.zeros(iters, np.float32)
    lib().ellc_ref_pyramid_run(_p(kf_image), _p(cur_image), dp, vp, level, _p(_f6(pose)), iters, C.byref(n), _p(w), _p(r), C.byref(le),
                               _p(hinv), _p(poses), _p(ratios))
    return dict(n=n.value, weights=w[:n.value].copy(), residual=r[:n.value].copy(), last_err=le.value, hessian_inv=hinv.reshape(6, 6),
                poses_after=poses, ratios=ratios)
