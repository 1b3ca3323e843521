"""The C++ host shim (reference class surface on top of the C-ABI): builds on CPU, runs and matches the oracle on GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "egomotion_with_local_loop_closures_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")


def build_shim():
    import __graft_entry__ as g
    g.build()
    src = [os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), os.path.join(PKG, "host", "HostShim.cpp"),
           os.path.join(PKG, "host", "PoseFiles.cpp")]
    deps = src + [os.path.join(PKG, "libellc_gn.so")]
    if not os.path.exists(EXE) or any(os.path.getmtime(s) > os.path.getmtime(EXE) for s in deps):
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-pthread", "-I", os.path.join(PKG, "host"), "-o", EXE] + src +
                              ["-L", PKG, "-lellc_gn", "-Wl,-rpath," + PKG])
    return EXE


def test_shim_builds_and_links():
    exe = build_shim()
    out = subprocess.check_output([exe, "none", "--link-only"], text=True)
    assert "link ok" in out


def test_pose_files_match_the_reference_format(tmp_path):
    """poses_orig.txt / matchframes*.txt (src/main.cpp:373,382; src/GlobalOptimize.cpp:580): the C++ writers use the reference's own
    stream expressions; the Python writers must produce the same bytes, and the files must parse back."""
    from egomotion_with_local_loop_closures_b200 import posefiles
    exe = build_shim()
    base = str(tmp_path / "poses")
    out = subprocess.check_output([exe, base, "--posefiles"], text=True)
    assert out.startswith("posefiles ok 1 107")
    w = np.array([0.0123456789, -1.5e-5, 3.0, 123456.789, -0.000123456, 1e-10], np.float32)
    o = -w[::-1]
    orig = open(base + ".orig").read()
    match = open(base + ".match").read().splitlines(keepends=True)
    assert orig == posefiles.orig_pose_line(7, 1, w, 0.98765432, 37.123456, batch_start_id=101)
    assert orig.split()[:4] == ["107", "101", "0.0123457", "-1.5e-05"]                  # 6 significant digits, ids offset by BATCH_START_ID-1
    assert match[0] == posefiles.match_pose_line(7, 1, o, 0.98765432, 37.123456, batch_start_id=101)
    assert match[1] == posefiles.match_pose_line(7, 1, o, 0.98765432, 12.5, 0.0712345, 9.87654321, 4.5, batch_start_id=101)
    assert match[0].rstrip().endswith(" 0 0 0")
    rows = posefiles.read_pose_file(base + ".match")
    assert rows.shape == (2, 13) and rows[1, 10] == pytest.approx(0.0712345, rel=1e-5)


@pytest.mark.gpu
def test_shim_tracks_like_the_reference_driver(tmp_path, oracle_mod):
    from egomotion_with_local_loop_closures_b200 import synth
    from tests.helpers import oracle_config
    exe = build_shim()
    w, h, n = 320, 240, 3
    scene = synth.SynthScene(w, h)
    kf = scene.keyframe(noise_seed=5)
    T = synth.smooth_trajectory(n + 1, seed_pose=3)
    frames = [scene.render(T[i + 1], noise_seed=50 + i) for i in range(n)]
    k = synth.intrinsics(w, h)
    blob = tmp_path / "case.bin"
    with open(blob, "wb") as f:
        f.write(struct.pack("<iii4f", w, h, n, float(k["fx"]), float(k["fy"]), float(k["cx"]), float(k["cy"])))
        f.write(kf["image"].tobytes())
        for l in range(4):
            f.write(np.ascontiguousarray(kf["depth"][l], np.float32).tobytes())
        for l in range(4):
            f.write(np.ascontiguousarray(kf["var"][l], np.float32).tobytes())
        for im in frames:
            f.write(im.tobytes())
    out = subprocess.check_output([exe, str(blob)], text=True)
    lines = [l.split() for l in out.strip().splitlines()]
    got_pose = {int(l[1]): np.array(l[2:8], np.float64) for l in lines if l[0] == "pose"}
    got_world = {int(l[1]): np.array(l[2:8], np.float64) for l in lines if l[0] == "world"}
    post = [l for l in lines if l[0] == "post"]
    assert [l for l in lines if l[0] == "pyr"][0][1:] == [str(w // 2), str(h // 2), str(w // 8), str(h // 8)]

    case = dict(width=w, height=h)
    ocfg = oracle_config(oracle_mod, case)
    kf_world = np.zeros(6, np.float32)
    prev_world = kf_world.copy()
    for i in range(n):
        init = oracle_mod.concat_origin(prev_world, kf_world)                 # src/ImageFunc.cpp:106
        opose, otr = oracle_mod.track(ocfg, kf["image"], frames[i], kf["depth"], kf["var"], init)
        world = oracle_mod.concat_relative(opose, kf_world)                   # src/ImageFunc.cpp:306
        assert np.abs(got_pose[i] - opose).max() < 1e-4 and np.abs(got_pose[i] - opose).max() < 2e-6
        assert np.abs(got_world[i] - world).max() < 2e-6
        gt = synth.relative_pose(T[i + 1], np.eye(4))
        assert np.abs(got_pose[i] - gt).max() < 3e-3
        assert post[i][2:] == ["0", "0", str(otr["n_selected"][0]), str(w), str(h)]          # level-0 post-conditions
        prev_world = world
    # caller-driven iterations through PixelWisePyramid at level 2
    kpyr = oracle_mod.image_pyramid(kf["image"]); cpyr = oracle_mod.image_pyramid(frames[0])
    pose = np.zeros(6, np.float32)
    iters = [l for l in lines if l[0] == "iter"]
    for it in range(3):
        o = oracle_mod.gn_evaluate(ocfg, 2, kpyr[2], cpyr[2], kf["depth"][2], kf["var"][2], pose)
        Hinv, _ = oracle_mod.invert6(o["H"])
        pose, delta, wp = oracle_mod.update_pose(ocfg, Hinv, o["b"], pose)
        g = iters[it]
        assert np.abs(np.array(g[2:8], np.float64) - pose).max() < 2e-6
        assert abs(float(g[11]) - o["res_sum_f64"]) <= 2.5e-5 * o["res_sum_f64"]
        assert abs(float(g[13]) - o["H_f64"][0, 0]) <= 2.5e-5 * o["H_f64"][0, 0]
    assert [l for l in lines if l[0] == "count2"][0][1] == str(int((kf["depth"][2] > 0).sum()))
    # keyframe rebuilt from 1/depth hypotheses through depthMap::updateDepthImage (device pyramids)
    d0 = kf["depth"][0]
    with np.errstate(divide="ignore"):
        ref = oracle_mod.update_depth_image((d0 > 0).astype(np.uint8), np.where(d0 > 0, np.float32(1) / d0, np.float32(-1)).astype(np.float32), kf["var"][0])
    hyp = [l for l in lines if l[0] == "hyp"][0]
    assert abs(float(hyp[1]) - ref["occupancy"]) < 1e-4
    assert abs(float(hyp[2]) - 100.0 * float((ref["valid_out"] != 0).sum()) / (w * h)) < 1e-3
    assert int(hyp[3]) == int((ref["depth"][1] > 0).sum()) and int(hyp[4]) == int((ref["depth"][3] > 0).sum())
    opose, _ = oracle_mod.track(ocfg, kf["image"], frames[0], ref["depth"], ref["var"], np.zeros(6, np.float32))
    assert np.abs(np.array(hyp[5:8], np.float64) - opose[:3]).max() < 2e-6
    # constant-weight loop-closure flow: weights saved by the sequential tracks, finalised, then one loop-closure pair
    h_, w_ = h, w
    wp = [np.zeros((h_ >> l, w_ >> l), np.float32) for l in range(4)]
    cnt = [0] * 4
    prev_world = kf_world.copy()
    worlds = []
    for i in range(n):
        init = oracle_mod.concat_origin(prev_world, kf_world)
        opose, _, wl = oracle_mod.track_with_weights(ocfg, kf["image"], frames[i], kf["depth"], kf["var"], init)
        oracle_mod.accumulate_weights(wp, cnt, wl)
        prev_world = oracle_mod.concat_relative(opose, kf_world)
        worlds.append(prev_world)
    wf = oracle_mod.finalise_weights(wp, cnt)
    assert [l for l in lines if l[0] == "nweights"][0][1:] == [str(n), str(n)]
    assert abs(float([l for l in lines if l[0] == "wsum1"][0][1]) - float(wf[1].sum(dtype=np.float64))) <= 1e-4 * float(wf[1].sum(dtype=np.float64))
    init = oracle_mod.concat_origin(worlds[0], kf_world)                      # t-1 frame of the loop-closure call = frame 0
    opose, otr = oracle_mod.track_lc(ocfg, kf["image"], frames[n - 1], kf["depth"], wf, init)
    got = np.array([l for l in lines if l[0] == "lcpose"][0][1:7], np.float64)
    assert np.abs(got - opose).max() < 1e-4


@pytest.mark.gpu
def test_shim_against_the_reference_driver_fixture(tmp_path):
    """The drop-in C++ surface (our frame / depthMap / GetImagePoseEstimate on the GPU) next to the REFERENCE'S OWN driver on
    the same inputs: tests/golden/reference_track_480x270.npz holds what the reference's GetImagePoseEstimate returned and the
    poseWrtWorld it left behind (src/ImageFunc.cpp:305-307) for a keyframe at the origin and a t-1 frame at the origin."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_track_480x270.npz"))
    exe = build_shim()
    w, h = int(g["width"][0]), int(g["height"][0])
    blob = tmp_path / "ref_case.bin"
    with open(blob, "wb") as f:
        f.write(struct.pack("<iii4f", w, h, 1, *(float(v) for v in g["intr"])))
        f.write(g["kf_image"].tobytes())
        for l in range(4):
            f.write(np.ascontiguousarray(g[f"depth{l}"], np.float32).tobytes())
        for l in range(4):
            f.write(np.ascontiguousarray(g[f"var{l}"], np.float32).tobytes())
        f.write(g["frames"][0].tobytes())
    out = subprocess.check_output([exe, str(blob)], text=True)
    lines = [l.split() for l in out.strip().splitlines()]
    pose = np.array([l for l in lines if l[0] == "pose"][0][2:8], np.float64)
    world = np.array([l for l in lines if l[0] == "world"][0][2:8], np.float64)
    post = [l for l in lines if l[0] == "post"][0]
    assert np.abs(pose - g["p0_driver_pose"]).max() < 2e-6
    assert np.abs(world - g["p0_driver_pose_wrt_world"]).max() < 2e-6
    assert post[2:5] == ["0", "0", str(int(g["p0_n_selected"][0]))]          # both frames back at level 0, level-0 mask count
    assert [l for l in lines if l[0] == "pyr"][0][1:] == ["240", "135", "60", "34"]      # pyrDown dims at 480x270: (h+1)/2 rows


def test_config_txt_reader(tmp_path):
    """config.txt of the batch ("LC") mode and main()'s argument handling (src/main.cpp:80-101, :132-137): three integers
    BATCH_START_ID BATCH_SIZE FLAG_IS_BOOTSTRAP, read only when argv[1] == "LC"; the reference's two error exits."""
    exe = build_shim()
    cfg = tmp_path / "config.txt"
    cfg.write_text("101 12 1\n")
    out = subprocess.check_output([exe, str(cfg), "--config"], text=True).splitlines()
    assert out[0] == "cfg0 0"
    assert out[1].startswith("cfg1 -1 Either Config. file or loop closure flag missing!")
    assert out[2].startswith("cfg2 -1 Unable to open Config. file!")
    assert out[3] == "cfg3 1 101 12 1 1"


def _write_blob(path, w, h, k, kf, frames):
    with open(path, "wb") as f:
        f.write(struct.pack("<iii4f", w, h, len(frames), float(k["fx"]), float(k["fy"]), float(k["cx"]), float(k["cy"])))
        f.write(kf["image"].tobytes())
        for l in range(4):
            f.write(np.ascontiguousarray(kf["depth"][l], np.float32).tobytes())
        for l in range(4):
            f.write(np.ascontiguousarray(kf["var"][l], np.float32).tobytes())
        for im in frames:
            f.write(im.tobytes())


@pytest.mark.gpu
def test_shim_class_surface(tmp_path, oracle_mod):
    """The rest of the reference's class surface through the C++ shim on the GPU, against the oracle:
    frame::getInterpolatedElement (both overloads, per-tap bound quirks), PixelWisePyramid with all display members / hessianInv /
    saveWeights(true|false) / calculatePixelWiseParallelInvCompositional, class Pyramid (performPrecomputation,
    performIterationSteps), calculateRandT, two host threads on separate contexts, a batch with more frames than slots."""
    from egomotion_with_local_loop_closures_b200 import synth
    from tests.helpers import oracle_config
    exe = build_shim()
    w, h, n = 320, 240, 3
    scene = synth.SynthScene(w, h)
    kf = scene.keyframe(noise_seed=5)
    T = synth.smooth_trajectory(n + 1, seed_pose=3)
    frames = [scene.render(T[i + 1], noise_seed=50 + i) for i in range(n)]
    k = synth.intrinsics(w, h)
    blob = tmp_path / "case.bin"
    _write_blob(blob, w, h, k, kf, frames)
    out = subprocess.check_output([exe, str(blob), "--surface"], text=True)
    lines = [l.split() for l in out.strip().splitlines() if l.strip()]
    get = lambda tag: [l for l in lines if l[0] == tag]
    L = 1
    ocfg = oracle_config(oracle_mod, dict(width=w, height=h))
    kpyr = oracle_mod.image_pyramid(kf["image"])
    cpyr = [oracle_mod.image_pyramid(f) for f in frames]
    rows, cols = h >> L, w >> L
    ogx, ogy = oracle_mod.gradient(cpyr[0][L], rows=rows, cols=cols)
    # (a) samplers: bit-identical to the oracle's restatement (which is bit-identical to the reference's own inline functions)
    for l in get("interp"):
        x, y = np.float32(l[1]), np.float32(l[2])
        want = [oracle_mod.interp_u8(cpyr[0][L], x, y, 1, rows=rows, cols=cols), oracle_mod.interp_u8(cpyr[0][L], x, y, 0, rows=rows, cols=cols),
                oracle_mod.interp_f32(ogx, x, y), oracle_mod.interp_f32(ogy, x, y)]
        assert [np.float32(v) for v in l[3:7]] == [np.float32(v) for v in want], l
    # (b) one forward iteration through the class, every display member
    pose_in = np.array([0.002, -0.001, 0.0015, 0.003, -0.002, 0.001], np.float32)
    o = oracle_mod.gn_evaluate(ocfg, L, kpyr[L], cpyr[0][L], kf["depth"][L], kf["var"][L], pose_in, want_weights=True)
    Hinv, _ = oracle_mod.invert6(o["H"])
    opose, _, owp = oracle_mod.update_pose(ocfg, Hinv, o["b"], pose_in)
    g = get("pw_pose")[0]
    assert np.abs(np.array(g[1:7], np.float64) - opose).max() < 2e-6 and abs(float(g[8]) - owp) <= 1e-4 * owp
    sel = kf["depth"][L] > 0
    d = get("pw_disp")[0]
    wsum = float(o["weights"].astype(np.float64).sum())
    assert abs(float(d[1]) - wsum) <= 1e-6 * wsum                                            # display_weightimg (STRICT: bit-identical weights)
    assert int(d[5]) == int((~sel).sum())                                                    # savedWarpedPoints == -2 where there is no depth
    assert int(d[6]) == o["n_oob"]                                                           # == -1 where the warp left the image
    assert int(d[7]) == int(cpyr[0][L][:rows, :cols][sel].astype(np.int64).sum())            # display_templateimg = CURRENT image under the mask
    assert int(d[8]) == int(kpyr[L][:rows, :cols][sel].astype(np.int64).sum())               # display_2bewarpedimg = keyframe image
    orig = (cpyr[0][L][:rows, :cols].astype(np.int64) - kpyr[L][:rows, :cols].astype(np.int64))[sel].sum()
    assert abs(float(d[4]) - orig) < 0.5                                                     # display_origres
    # individual pixels: warped coordinates by the reference's fp32 operation sequence (numpy float32 = individually rounded),
    # warped intensity by the oracle's sampler, residual = warped - keyframe intensity, weight = the oracle's
    Tm = oracle_mod.se3_exp(pose_in).astype(np.float32)
    fx, fy, cx, cy = (np.float32(float(k[n_]) / 2 ** L) for n_ in ("fx", "fy", "cx", "cy"))
    f32 = np.float32
    for l in get("pw_px"):
        x, y = int(l[1]), int(l[2])
        dep = f32(kf["depth"][L][y, x])
        X = f32(f32(f32(x) - cx) * dep) / fx; Y = f32(f32(f32(y) - cy) * dep) / fy
        tx = f32(f32(f32(Tm[0, 0] * X) + f32(Tm[0, 1] * Y)) + f32(Tm[0, 2] * dep)) + Tm[0, 3]
        ty = f32(f32(f32(Tm[1, 0] * X) + f32(Tm[1, 1] * Y)) + f32(Tm[1, 2] * dep)) + Tm[1, 3]
        tz = f32(f32(f32(Tm[2, 0] * X) + f32(Tm[2, 1] * Y)) + f32(Tm[2, 2] * dep)) + Tm[2, 3]
        u = f32(f32(tx / tz) * fx) + cx; v = f32(f32(ty / tz) * fy) + cy
        assert f32(l[3]) == f32(u) and f32(l[4]) == f32(v), l
        iw = oracle_mod.interp_u8(cpyr[0][L], u, v, 1, rows=rows, cols=cols)
        assert f32(l[5]) == f32(iw) and f32(l[6]) == f32(iw) - f32(kpyr[L][y, x]), l
        assert f32(l[7]) == o["weights"][y, x], l
    assert float(get("pw_hinv")[0][1]) < 1e-3                                                # hessian * hessianInv ~ I
    s = get("pw_save")[0]
    assert int(s[1]) == 2 and abs(float(s[2]) - float(s[3])) <= 1e-6 * float(s[3])           # saveWeights(true) twice: sums add, count 2
    assert float(get("pw_scatter")[0][1]) > 0                                                # saveWeights(false): scattered to the warped positions
    # (c) constant-weight iterations: only level L carries weights (finalised: (w + w) / 2 = w); the oracle runs that level alone
    wf = [np.zeros((h >> l, w >> l), np.float32) for l in range(4)]
    wf[L] = o["weights"]
    lcfg = oracle_config(oracle_mod, dict(width=w, height=h), max_iter=[0, 3, 0, 0], stop_threshold=-1.0)
    _, otr = oracle_mod.track_lc(lcfg, kf["image"], frames[1], kf["depth"], wf, np.zeros(6, np.float32))
    for it, l in enumerate(get("lc_iter")):
        ol = otr["levels"][L][it]
        assert np.abs(np.array(l[2:8], np.float64) - ol["pose_after"]).max() < 5e-6, (it, l)
        assert abs(float(l[9]) - ol["weighted_pose"]) <= 1e-3 * max(1.0, ol["weighted_pose"])
        assert abs(float(l[11]) - ol["H_f64"][0, 0]) <= 1e-5 * ol["H_f64"][0, 0]
    # (d) class Pyramid: the matrix-form variant (Jacobian at the warped pixel), free-running for two iterations
    pcfg = oracle_config(oracle_mod, dict(width=w, height=h), jacobian_at_warped=1)
    po = oracle_mod.gn_evaluate(pcfg, L, kpyr[L], cpyr[0][L], kf["depth"][L], kf["var"][L], pose_in)
    nsel = int(sel.sum())
    pre = get("pyr_pre")[0]
    assert int(pre[2]) == nsel and abs(float(pre[1]) - po["res_sum_f64"] / nsel) <= 1e-5 * po["res_sum_f64"] / nsel
    pose = pose_in
    last = po["res_sum_f64"] / nsel
    for it, l in enumerate(get("pyr_iter")):
        Hinv, _ = oracle_mod.invert6(po["H"])
        pose, _, owp = oracle_mod.update_pose(pcfg, Hinv, po["b"], pose)
        po = oracle_mod.gn_evaluate(pcfg, L, kpyr[L], cpyr[0][L], kf["depth"][L], kf["var"][L], pose)
        err = po["res_sum_f64"] / nsel
        assert np.abs(np.array(l[2:8], np.float64) - pose).max() < 2e-6, (it, l)
        assert abs(float(l[10]) - err / last) <= 1e-4 and abs(float(l[12]) - err) <= 1e-4 * err
        last = err
    # (e) calculateRandT: SE3_Pose = exp(hat(poseWrtWorld))
    world = np.array(get("world0")[0][1:7], np.float32)
    Tw = oracle_mod.se3_exp(world)
    r = [float(v) for v in get("randt")[0][1:7]]
    assert np.allclose(r, [Tw[0, 1], Tw[1, 2], Tw[0, 3], Tw[2, 3], 1.0, Tw[0, 0]], rtol=0, atol=2e-7)
    # (f) two threads, two contexts, bit-identical poses; (g) oversized batch split without overwriting slots
    t = get("threads")[0]
    assert t[1] == "0" and t[2] == "0" and int(t[4]) >= 3, t
    assert float(get("bigbatch")[0][1]) < 1e-6
