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
