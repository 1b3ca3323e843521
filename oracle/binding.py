"""ctypes binding of oracle/libellc_oracle.so (test infrastructure only; see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libellc_oracle.so")
LEVELS = 4
MAX_ITERS = 64


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++ -std=c++11 -O3, the reference's flags)."""
    src = os.path.join(_HERE, "ellc_oracle.cpp")
    hdr = os.path.join(_HERE, "ellc_oracle.h")
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _SO


class Config(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("max_iter", C.c_int * LEVELS),
                ("huber_d", C.c_float), ("camera_pixel_noise_2", C.c_float),
                ("weight", C.c_float * 6), ("stop_threshold", C.c_float),
                ("num_bands", C.c_int), ("use_threads", C.c_int), ("jacobian_at_warped", C.c_int),
                ("lc_parallel", C.c_int)]


class Iter(C.Structure):
    _fields_ = [("H", C.c_float * 36), ("b", C.c_float * 6), ("delta", C.c_float * 6),
                ("weighted_pose", C.c_float), ("pose_after", C.c_float * 6),
                ("res_sum_f32", C.c_float), ("res_sum_f64", C.c_double), ("n_oob", C.c_int), ("pad_", C.c_int),
                ("H_f64", C.c_double * 36), ("b_f64", C.c_double * 6)]


class Trace(C.Structure):
    _fields_ = [("n_selected", C.c_int * LEVELS), ("n_iters", C.c_int * LEVELS),
                ("it", (Iter * MAX_ITERS) * LEVELS), ("final_pose", C.c_float * 6)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ellc_oracle_interp_u8.restype = C.c_float
        _lib.ellc_oracle_interp_f32.restype = C.c_float
        _lib.ellc_oracle_interp_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        _lib.ellc_oracle_interp_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        _lib.ellc_oracle_mask_count.restype = C.c_int
        _lib.ellc_oracle_invert6.restype = C.c_int
        _lib.ellc_oracle_track_many.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f6(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(6))


def default_config(width, height, **over):
    cfg = Config()
    lib().ellc_oracle_default_config(C.byref(cfg), int(width), int(height))
    for k, v in over.items():
        if k in ("max_iter", "weight"):
            arr = getattr(cfg, k)
            for i, x in enumerate(v):
                arr[i] = x
        else:
            setattr(cfg, k, v)
    return cfg


def pyr_dims(w, h):
    """Image-pyramid dims as cv::pyrDown produces them: ((w+1)//2, (h+1)//2) per level."""
    dims = [(w, h)]
    for _ in range(1, LEVELS):
        w, h = (w + 1) // 2, (h + 1) // 2
        dims.append((w, h))
    return dims


def pyrdown(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().ellc_oracle_pyrdown_u8(_p(img), w, h, w, _p(out))
    return out


def image_pyramid(img0):
    pyr = [np.ascontiguousarray(img0, dtype=np.uint8)]
    for _ in range(1, LEVELS):
        pyr.append(pyrdown(pyr[-1]))
    return pyr


def gradient(img, rows=None, cols=None):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    rows = img.shape[0] if rows is None else rows
    cols = img.shape[1] if cols is None else cols
    gx = np.empty((rows, cols), np.float32)
    gy = np.empty((rows, cols), np.float32)
    lib().ellc_oracle_gradient(_p(img), img.shape[1], rows, cols, _p(gx), _p(gy))
    return gx, gy


def mask_count(depth):
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    mask = np.empty(depth.shape, np.uint8)
    n = lib().ellc_oracle_mask_count(_p(depth), depth.size, _p(mask))
    return mask, n


def interp_u8(img, x, y, check_oob=1, rows=None, cols=None):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    rows = img.shape[0] if rows is None else rows
    cols = img.shape[1] if cols is None else cols
    return lib().ellc_oracle_interp_u8(_p(img), img.shape[1], rows, cols, float(np.float32(x)), float(np.float32(y)), check_oob)


def interp_f32(img, x, y):
    img = np.ascontiguousarray(img, dtype=np.float32)
    return lib().ellc_oracle_interp_f32(_p(img), img.shape[1], img.shape[0], img.shape[1], float(np.float32(x)), float(np.float32(y)))


def build_depth_pyramid(depth0_arr, var0):
    """depth0_arr / var0: level-0 arrays in the updateDepthImage convention (var>0 marks valid)."""
    h, w = depth0_arr.shape
    d = [np.ascontiguousarray(depth0_arr, np.float32)]
    v = [np.ascontiguousarray(var0, np.float32)]
    for l in range(1, LEVELS):
        d.append(np.zeros((h >> l, w >> l), np.float32))
        v.append(np.zeros((h >> l, w >> l), np.float32))
    dp = (C.c_void_p * LEVELS)(*[_p(a) for a in d])
    vp = (C.c_void_p * LEVELS)(*[_p(a) for a in v])
    lib().ellc_oracle_build_depth_pyramid(w, h, dp, vp)
    return d, v


def update_depth_image(valid, inv_depth_smoothed, variance_smoothed):
    """Level-0 depth image / arrays from per-pixel hypotheses + the 4-level pyramids (updateDepthImage -> buildInvVarDepth ->
    mapDepthArr2Mat).  Returns dict(valid_out, depth [4 levels, Mat convention: 0 = invalid], var [4 levels, -1 = invalid],
    n_valid, occupancy)."""
    valid = np.ascontiguousarray(valid, np.uint8).copy()
    h, w = valid.shape
    idep = np.ascontiguousarray(inv_depth_smoothed, np.float32)
    vs = np.ascontiguousarray(variance_smoothed, np.float32)
    dmat = np.zeros((h, w), np.float32); darr = np.zeros((h, w), np.float32); varr = np.zeros((h, w), np.float32)
    occ = C.c_float()
    _l = lib()
    _l.ellc_oracle_update_depth_image.restype = C.c_int
    n = _l.ellc_oracle_update_depth_image(w, h, _p(valid), _p(idep), _p(vs), _p(dmat), _p(darr), _p(varr), C.byref(occ))
    d, v = build_depth_pyramid(darr, varr)
    d[0] = dmat                                             # mapDepthArr2Mat: depth_pyramid[0] = keyFrame->depth
    return dict(valid_out=valid, depth=d, var=v, n_valid=n, occupancy=occ.value)


def image_histogram(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(256, np.float32)
    lib().ellc_oracle_image_histogram(_p(img), img.size, _p(out))
    return out


def hist_kl_div(h1, h2):
    _l = lib()
    _l.ellc_oracle_hist_kl_div.restype = C.c_double
    return _l.ellc_oracle_hist_kl_div(_p(np.ascontiguousarray(h1, np.float32)), _p(np.ascontiguousarray(h2, np.float32)))


def rotation_stats(pose1, pose2):
    rms, ang = C.c_float(), C.c_float()
    lib().ellc_oracle_rotation_stats(_p(_f6(pose1)), _p(_f6(pose2)), C.byref(rms), C.byref(ang))
    return rms.value, ang.value


def se3_exp(pose):
    T = np.empty(16, np.float32)
    lib().ellc_oracle_se3_exp(_p(_f6(pose)), _p(T))
    return T.reshape(4, 4)


def se3_log(T):
    T = np.ascontiguousarray(np.asarray(T, np.float32).reshape(16))
    out = np.empty(6, np.float32)
    lib().ellc_oracle_se3_log(_p(T), _p(out))
    return out


def concat_relative(a, b):
    out = np.empty(6, np.float32)
    lib().ellc_oracle_concat_relative(_p(_f6(a)), _p(_f6(b)), _p(out))
    return out


def concat_origin(a, b):
    out = np.empty(6, np.float32)
    lib().ellc_oracle_concat_origin(_p(_f6(a)), _p(_f6(b)), _p(out))
    return out


def invert6(H):
    H = np.ascontiguousarray(np.asarray(H, np.float32).reshape(36))
    out = np.empty(36, np.float32)
    ok = lib().ellc_oracle_invert6(_p(H), _p(out))
    return out.reshape(6, 6), ok


def update_pose(cfg, Hinv, b, pose):
    Hinv = np.ascontiguousarray(np.asarray(Hinv, np.float32).reshape(36))
    b = _f6(b)
    pose = _f6(pose).copy()
    delta = np.empty(6, np.float32)
    wp = C.c_float()
    lib().ellc_oracle_update_pose(C.byref(cfg), _p(Hinv), _p(b), _p(pose), _p(delta), C.byref(wp))
    return pose, delta, wp.value


def iter_to_dict(it):
    return dict(H=np.array(it.H, np.float32).reshape(6, 6), b=np.array(it.b, np.float32),
                delta=np.array(it.delta, np.float32), weighted_pose=it.weighted_pose,
                pose_after=np.array(it.pose_after, np.float32), res_sum_f32=it.res_sum_f32,
                res_sum_f64=it.res_sum_f64, n_oob=it.n_oob,
                H_f64=np.array(it.H_f64, np.float64).reshape(6, 6), b_f64=np.array(it.b_f64, np.float64))


def gn_evaluate(cfg, level, kf_img, cur_img, depth, var, pose, want_weights=False):
    """H, b and residual statistics at `pose` on `level` (no pose update)."""
    kf_img = np.ascontiguousarray(kf_img, np.uint8)
    cur_img = np.ascontiguousarray(cur_img, np.uint8)
    depth = np.ascontiguousarray(depth, np.float32)
    var = np.ascontiguousarray(var, np.float32)
    it = Iter()
    wimg = np.zeros(depth.shape, np.float32) if want_weights else None
    lib().ellc_oracle_gn_evaluate(C.byref(cfg), level, _p(kf_img), kf_img.shape[1], _p(cur_img), cur_img.shape[1],
                                  _p(depth), _p(var), _p(_f6(pose)), C.byref(it), _p(wimg) if want_weights else None)
    d = iter_to_dict(it)
    if want_weights:
        d["weights"] = wimg
    return d


def trace_to_dict(tr):
    out = dict(n_selected=list(tr.n_selected), n_iters=list(tr.n_iters), final_pose=np.array(tr.final_pose, np.float32), levels=[])
    for l in range(LEVELS):
        out["levels"].append([iter_to_dict(tr.it[l][i]) for i in range(tr.n_iters[l])])
    return out


def track(cfg, kf_img0, cur_img0, depth_pyr, var_pyr, init_pose, want_trace=True):
    kf_img0 = np.ascontiguousarray(kf_img0, np.uint8)
    cur_img0 = np.ascontiguousarray(cur_img0, np.uint8)
    d = [np.ascontiguousarray(a, np.float32) for a in depth_pyr]
    v = [np.ascontiguousarray(a, np.float32) for a in var_pyr]
    dp = (C.c_void_p * LEVELS)(*[_p(a) for a in d])
    vp = (C.c_void_p * LEVELS)(*[_p(a) for a in v])
    out = np.empty(6, np.float32)
    tr = Trace() if want_trace else None
    lib().ellc_oracle_track(C.byref(cfg), _p(kf_img0), _p(cur_img0), dp, vp, _p(_f6(init_pose)), _p(out),
                            C.byref(tr) if want_trace else None)
    return out, (trace_to_dict(tr) if want_trace else None)


def _pyr_ptrs(arrs, dtype):
    keep = [np.ascontiguousarray(a, dtype) for a in arrs]
    return keep, (C.c_void_p * LEVELS)(*[_p(a) for a in keep])


def track_with_weights(cfg, kf_img0, cur_img0, depth_pyr, var_pyr, init_pose):
    """Forward track + per level display_weightimg of the last executed iteration (what saveWeights(true) accumulates)."""
    w, h = cfg.width, cfg.height
    kfp, kfpp = _pyr_ptrs(image_pyramid(kf_img0), np.uint8)
    cup, cupp = _pyr_ptrs(image_pyramid(cur_img0), np.uint8)
    d, dp = _pyr_ptrs(depth_pyr, np.float32)
    v, vp = _pyr_ptrs(var_pyr, np.float32)
    wl = [np.zeros((h >> l, w >> l), np.float32) for l in range(LEVELS)]
    wlp = (C.c_void_p * LEVELS)(*[_p(a) for a in wl])
    out = np.empty(6, np.float32)
    tr = Trace()
    lib().ellc_oracle_track_weights_prebuilt(C.byref(cfg), kfpp, cupp, dp, vp, _p(_f6(init_pose)), _p(out), C.byref(tr), wlp)
    return out, trace_to_dict(tr), wl


def accumulate_weights(weight_pyr, counts, weight_last):
    """saveWeights(true), src/PixelWisePyramid.cpp:546-548: weight_pyramid[l] += display_weightimg; numWeightsAdded[l]++."""
    for l in range(LEVELS):
        weight_pyr[l] = (weight_pyr[l].astype(np.float32) + weight_last[l].astype(np.float32)).astype(np.float32)
        counts[l] += 1


def finalise_weights(weight_pyr, counts):
    """frame::finaliseWeights, src/Frame.cpp:678-695: weight_pyramid[l] = weight_pyramid[l] / numWeightsAdded[l] when > 0.
    cv::Mat / scalar is a multiplication by the reciprocal (matop.cpp: MatOp_AddEx with alpha = 1./s, applied by the CV_32F
    convertTo kernel as a float multiply by (float)alpha), not a division."""
    return [(weight_pyr[l].astype(np.float32) * np.float32(1.0 / counts[l])).astype(np.float32) if counts[l] > 0 else weight_pyr[l]
            for l in range(LEVELS)]


def track_lc(cfg, kf_img0, cur_img0, depth_pyr, weight_pyr, init_pose):
    """Inverse-compositional constant-weight tracker (loop-closure pairs)."""
    kfp, kfpp = _pyr_ptrs(image_pyramid(kf_img0), np.uint8)
    cup, cupp = _pyr_ptrs(image_pyramid(cur_img0), np.uint8)
    d, dp = _pyr_ptrs(depth_pyr, np.float32)
    wt, wp = _pyr_ptrs(weight_pyr, np.float32)
    out = np.empty(6, np.float32)
    tr = Trace()
    lib().ellc_oracle_track_lc_prebuilt(C.byref(cfg), kfpp, cupp, dp, wp, _p(_f6(init_pose)), _p(out), C.byref(tr))
    return out, trace_to_dict(tr)


def track_many(cfg, kf_idx, fr_idx, kf_imgs, fr_imgs, kf_depth, kf_var, init_poses, n_workers=1):
    """CPU-baseline driver. kf_depth/kf_var: list (per keyframe) of 4-level lists. Returns (poses, seconds)."""
    n = len(kf_idx)
    kf_idx = np.ascontiguousarray(kf_idx, np.int32)
    fr_idx = np.ascontiguousarray(fr_idx, np.int32)
    kf_imgs = [np.ascontiguousarray(a, np.uint8) for a in kf_imgs]
    fr_imgs = [np.ascontiguousarray(a, np.uint8) for a in fr_imgs]
    kd = [np.ascontiguousarray(a, np.float32) for pyr in kf_depth for a in pyr]
    kv = [np.ascontiguousarray(a, np.float32) for pyr in kf_var for a in pyr]
    kip = (C.c_void_p * len(kf_imgs))(*[_p(a) for a in kf_imgs])
    fip = (C.c_void_p * len(fr_imgs))(*[_p(a) for a in fr_imgs])
    kdp = (C.c_void_p * len(kd))(*[_p(a) for a in kd])
    kvp = (C.c_void_p * len(kv))(*[_p(a) for a in kv])
    init = np.ascontiguousarray(np.asarray(init_poses, np.float32).reshape(n, 6))
    out = np.empty((n, 6), np.float32)
    secs = lib().ellc_oracle_track_many(C.byref(cfg), n, int(n_workers), _p(kf_idx), _p(fr_idx), kip, fip, kdp, kvp, _p(init), _p(out))
    return out, secs
