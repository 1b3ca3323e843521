"""Seeded synthetic inputs for the frame-to-keyframe GN tracker.

The reference ships no data, so tests and bench.py render their own: a textured height-field surface seen by a
pinhole camera at known poses.  Everything here is numpy data generation -- it is not on the product path.

Conventions (same as the reference, src/PixelWisePyramid.cpp:153,236-264): a pose is the Lie-algebra 6-vector
``[wx wy wz vx vy vz]``; ``exp(hat(pose))`` maps keyframe-camera points to current-camera points.  Keyframe depth
images use 0 for "no depth" (src/DepthPropagation.cpp:1290-1296) and variance arrays use -1.
"""
from __future__ import annotations

import numpy as np

LEVELS = 4
MIN_ABS_GRAD_DECREASE = 5.0      # src/ExternVariable.h:82
BORDER = 3                       # src/DepthPropagation.cpp:1279-1282


# ---------------------------------------------------------------------------------------------------------
# SE(3) helpers (float64, closed form) -- ground-truth pose algebra for the generator and the tests
# ---------------------------------------------------------------------------------------------------------
def hat3(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)


def se3_exp(p):
    p = np.asarray(p, np.float64)
    w, v = p[:3], p[3:]
    th = np.linalg.norm(w)
    W = hat3(w)
    if th < 1e-8:
        R = np.eye(3) + W + 0.5 * W @ W
        V = np.eye(3) + 0.5 * W + W @ W / 6.0
    else:
        A, B, Cc = np.sin(th) / th, (1 - np.cos(th)) / th**2, (th - np.sin(th)) / th**3
        R = np.eye(3) + A * W + B * W @ W
        V = np.eye(3) + B * W + Cc * W @ W
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ v
    return T


def se3_log(T):
    R, t = T[:3, :3], T[:3, 3]
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-8:
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    else:
        w = th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    W = hat3(w)
    if th < 1e-6:
        Vinv = np.eye(3) - 0.5 * W + W @ W / 12.0
    else:
        Vinv = np.eye(3) - 0.5 * W + (1 - th * np.sin(th) / (2 * (1 - np.cos(th)))) / th**2 * W @ W
    return np.concatenate([w, Vinv @ t])


def relative_pose(T_cur_w, T_kf_w):
    """Lie-algebra pose of `cur` w.r.t. `kf`: log(T_cur_w T_kf_w^-1) (src/Frame.cpp:534-562 semantics)."""
    return se3_log(T_cur_w @ np.linalg.inv(T_kf_w))


# ---------------------------------------------------------------------------------------------------------
def intrinsics(width, height):
    """fx = fy = 0.8 W, principal point at the image centre (ratios of src/ExternVariable.h:53-59)."""
    return dict(fx=np.float32(0.8 * width), fy=np.float32(0.8 * width), cx=np.float32(width / 2.0), cy=np.float32(height / 2.0))


def _bandpass_field(rng, shape, wavelength, rel_bw=0.35):
    """Unit-variance real random field whose spectrum is a Gaussian ring at 1/wavelength cycles/texel."""
    h, w = shape
    noise = rng.standard_normal(shape)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    f = np.sqrt(fx * fx + fy * fy)
    f0 = 1.0 / wavelength
    filt = np.exp(-0.5 * ((f - f0) / (rel_bw * f0)) ** 2)
    field = np.fft.irfft2(np.fft.rfft2(noise) * filt, s=shape)
    field -= field.mean()
    field /= field.std()
    return field


class SynthScene:
    """Textured height-field Z = h(X, Y) in world coordinates, rendered by exact ray/surface intersection."""

    def __init__(self, width, height, seed_tex=1234, edge_gain=12.0, wavelength_px=56.0, k=None):
        self.width, self.height = int(width), int(height)
        k = k or intrinsics(width, height)                 # k: intrinsics override (the reference's compiled-in camera)
        self.fx, self.fy, self.cx, self.cy = (float(k[n]) for n in ("fx", "fy", "cx", "cy"))
        rng = np.random.default_rng(seed_tex)
        # surface: tilted plane + 4 low-frequency bumps, Z in ~[1.0, 2.2] over the visible region
        self.z0, self.slope = 1.55, np.array([0.16, -0.11])
        self.bump_amp = np.array([0.07, 0.05, 0.04, 0.03])
        self.bump_k = rng.uniform(-1, 1, (4, 2))
        self.bump_k *= (2 * np.pi / np.array([2.3, 1.7, 1.4, 1.1]))[:, None] / np.linalg.norm(self.bump_k, axis=1, keepdims=True)
        self.bump_phi = rng.uniform(0, 2 * np.pi, 4)
        # texture: one texel ~ one level-0 pixel footprint at Z = z0
        self.texel = self.z0 / self.fx
        half_w = 0.5 * self.width / self.fx * 2.3 + 0.35
        half_h = 0.5 * self.height / self.fy * 2.3 + 0.35
        tw, th = int(2 * half_w / self.texel) + 2, int(2 * half_h / self.texel) + 2
        self.tex_origin = np.array([-half_w, -half_h])
        edges = _bandpass_field(rng, (th, tw), wavelength_px)
        shade = _bandpass_field(rng, (th, tw), 9.0 * wavelength_px)
        fine = _bandpass_field(rng, (th, tw), 0.3 * wavelength_px, rel_bw=0.5)
        # plateaus separated by soft edges -> semi-dense gradient structure; faint fine grain keeps plateaus non-constant
        self.tex = (128.0 + 78.0 * np.tanh(edge_gain * edges) + 14.0 * shade + 0.8 * fine).astype(np.float32)

    # -- surface ------------------------------------------------------------------------------------------
    def _h(self, X, Y):
        z = self.z0 + self.slope[0] * X + self.slope[1] * Y
        for a, k, p in zip(self.bump_amp, self.bump_k, self.bump_phi):
            z = z + a * np.sin(k[0] * X + k[1] * Y + p)
        return z

    def _dh(self, X, Y):
        gx = np.full_like(X, self.slope[0])
        gy = np.full_like(X, self.slope[1])
        for a, k, p in zip(self.bump_amp, self.bump_k, self.bump_phi):
            c = a * np.cos(k[0] * X + k[1] * Y + p)
            gx = gx + c * k[0]
            gy = gy + c * k[1]
        return gx, gy

    def _intersect(self, T_cw):
        """Per-pixel ray parameter s (= camera-frame depth) and world hit point for camera pose T_cw."""
        R, t = T_cw[:3, :3], T_cw[:3, 3]
        o = -R.T @ t
        u, v = np.meshgrid(np.arange(self.width, dtype=np.float64), np.arange(self.height, dtype=np.float64))
        dc = np.stack([(u - self.cx) / self.fx, (v - self.cy) / self.fy, np.ones_like(u)], -1)
        d = dc @ R                                    # = R^T dc per pixel
        s = np.full(u.shape, self.z0)
        for _ in range(8):                            # Newton on f(s) = o_z + s d_z - h(o_xy + s d_xy)
            X, Y = o[0] + s * d[..., 0], o[1] + s * d[..., 1]
            gx, gy = self._dh(X, Y)
            f = o[2] + s * d[..., 2] - self._h(X, Y)
            s = s - f / (d[..., 2] - gx * d[..., 0] - gy * d[..., 1])
        return s, o[0] + s * d[..., 0], o[1] + s * d[..., 1]

    def _sample_tex(self, X, Y):
        tx = (X - self.tex_origin[0]) / self.texel
        ty = (Y - self.tex_origin[1]) / self.texel
        th, tw = self.tex.shape
        tx = np.clip(tx, 0, tw - 1.001)
        ty = np.clip(ty, 0, th - 1.001)
        x0, y0 = tx.astype(np.int64), ty.astype(np.int64)
        ax, ay = tx - x0, ty - y0
        t = self.tex
        return ((1 - ay) * ((1 - ax) * t[y0, x0] + ax * t[y0, x0 + 1]) + ay * ((1 - ax) * t[y0 + 1, x0] + ax * t[y0 + 1, x0 + 1]))

    # -- products -----------------------------------------------------------------------------------------
    def render(self, T_cw=None, noise_seed=None, noise_sigma=1.0):
        """u8 image seen from camera pose T_cw (4x4, camera <- world).  Optional seeded sensor noise."""
        T_cw = np.eye(4) if T_cw is None else np.asarray(T_cw, np.float64)
        _, X, Y = self._intersect(T_cw)
        img = self._sample_tex(X, Y)
        if noise_seed is not None:
            img = img + noise_sigma * np.random.default_rng(noise_seed).standard_normal(img.shape)
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    def depth(self, T_cw=None):
        T_cw = np.eye(4) if T_cw is None else np.asarray(T_cw, np.float64)
        return self._intersect(T_cw)[0]

    def keyframe(self, T_cw=None, seed_depth=5678, noise_seed=None, idepth_noise=0.02, variance=0.01):
        """Keyframe bundle: u8 image + semi-dense depth/variance pyramids as the depth module would hand them over.

        Level 0 (src/DepthPropagation.cpp:1254-1315): valid where the smeared max-gradient >= MIN_ABS_GRAD_DECREASE
        (rule of src/Frame.cpp:618-674) and >= 3 px from the border; depth = 1/(rho_gt (1 + 0.02 N(0,1))).
        Levels 1..3: buildInvVarDepth (src/DepthPropagation.cpp:1637-1719).
        """
        T_cw = np.eye(4) if T_cw is None else np.asarray(T_cw, np.float64)
        img = self.render(T_cw, noise_seed=noise_seed)
        z = self.depth(T_cw)
        valid = select_semidense(img)
        rng = np.random.default_rng(seed_depth)
        idepth = (1.0 / z) * (1.0 + idepth_noise * rng.standard_normal(z.shape))
        depth0 = np.where(valid, 1.0 / idepth, 0.0).astype(np.float32)
        var0 = np.where(valid, variance, -1.0).astype(np.float32)
        depth, var = build_inv_var_depth(depth0, var0)
        return dict(image=img, depth=depth, var=var, gt_depth=z.astype(np.float32))


def image_gradient(img):
    """frame::calculateGradient (src/Frame.cpp:185-285) in numpy (used only to pick semi-dense pixels)."""
    f = img.astype(np.float32)
    gx = np.empty_like(f)
    gy = np.empty_like(f)
    gx[:, 1:-1] = 0.5 * (f[:, 2:] - f[:, :-2])
    gx[:, 0] = f[:, 1] - f[:, 0]
    gx[:, -1] = f[:, -1] - f[:, -2]
    gy[1:-1, :] = 0.5 * (f[2:, :] - f[:-2, :])
    gy[0, :] = f[1, :] - f[0, :]
    gy[-1, :] = f[-1, :] - f[-2, :]
    return gx, gy


def select_semidense(img, thresh=MIN_ABS_GRAD_DECREASE, border=BORDER):
    """Pixels whose 3x3-smeared gradient magnitude >= thresh (src/Frame.cpp:618-674), minus a 3 px border."""
    gx, gy = image_gradient(img)
    g = np.sqrt(gx * gx + gy * gy)
    t = np.zeros_like(g)
    t[1:-1, :] = np.maximum(np.maximum(g[1:-1, :], g[:-2, :]), g[2:, :])
    m = g.copy()
    m[1:-1, 1:-1] = np.maximum(np.maximum(t[1:-1, :-2], t[1:-1, 1:-1]), t[1:-1, 2:])
    valid = m >= thresh
    valid[:border, :] = valid[-border:, :] = False
    valid[:, :border] = valid[:, -border:] = False
    return valid


def build_inv_var_depth(depth0, var0):
    """depthMap::buildInvVarDepth (src/DepthPropagation.cpp:1637-1719), vectorised fp32 with the same op order.

    depth0: level-0 depth image (0 = invalid); var0: level-0 variance (-1 = invalid).  Children are read only where
    var > 0, so the 0-vs(-1) convention of the level-0 depth array does not matter.
    """
    depth, var = [np.ascontiguousarray(depth0, np.float32)], [np.ascontiguousarray(var0, np.float32)]
    h, w = depth0.shape
    one = np.float32(1.0)
    for l in range(1, LEVELS):
        ds, vs = depth[-1], var[-1]
        hh, ww = h >> l, w >> l
        ivar_sum = np.zeros((hh, ww), np.float32)
        idep_sum = np.zeros((hh, ww), np.float32)
        num = np.zeros((hh, ww), np.int32)
        with np.errstate(divide="ignore", invalid="ignore"):
            for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
                v = vs[dy:2 * hh:2, dx:2 * ww:2][:hh, :ww]
                d = ds[dy:2 * hh:2, dx:2 * ww:2][:hh, :ww]
                ok = v > 0
                ivar = np.where(ok, one / v, np.float32(0)).astype(np.float32)
                ivar_sum = (ivar_sum + ivar).astype(np.float32)
                idep_sum = (idep_sum + np.where(ok, (ivar * one) / d, np.float32(0)).astype(np.float32)).astype(np.float32)
                num += ok
            dn = np.where(num > 0, ivar_sum / idep_sum, np.float32(0)).astype(np.float32)
            vn = np.where(num > 0, num.astype(np.float32) / ivar_sum, np.float32(-1)).astype(np.float32)
        depth.append(np.ascontiguousarray(dn))
        var.append(np.ascontiguousarray(vn))
    return depth, var


# ---------------------------------------------------------------------------------------------------------
# motion models
# ---------------------------------------------------------------------------------------------------------
def smooth_trajectory(n_frames, seed_pose=91011, rot_step=np.deg2rad(0.4), trans_step=0.008):
    """World poses T_cw[k] of a smooth 6-DoF motion: <= rot_step rad and <= trans_step units per frame."""
    rng = np.random.default_rng(seed_pose)
    dirs = rng.standard_normal((3, 6))
    T = [np.eye(4)]
    for k in range(1, n_frames):
        ph = 2 * np.pi * k / 37.0
        xi = dirs[0] + 0.5 * np.sin(ph) * dirs[1] + 0.5 * np.cos(0.7 * ph) * dirs[2]
        w = xi[:3] / max(np.linalg.norm(xi[:3]), 1e-9) * rot_step * (0.6 + 0.4 * np.sin(0.31 * k) ** 2)
        v = xi[3:] / max(np.linalg.norm(xi[3:]), 1e-9) * trans_step * (0.6 + 0.4 * np.cos(0.23 * k) ** 2)
        T.append(se3_exp(np.concatenate([w, v])) @ T[-1])
    return T


def random_pose(rng, rot=np.deg2rad(1.5), trans=0.02):
    w = rng.standard_normal(3)
    v = rng.standard_normal(3)
    return np.concatenate([w / np.linalg.norm(w) * rot * rng.uniform(0.3, 1.0), v / np.linalg.norm(v) * trans * rng.uniform(0.3, 1.0)])
