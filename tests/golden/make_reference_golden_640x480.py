"""Generates tests/golden/reference_track_640x480.npz from THE REFERENCE'S OWN CODE at the METRIC resolution (640x480).

The reference fixes its camera at compile time (src/ExternVariable.h:50-59); oracle/_ref/libellc_ref_640x480.so is the same
unmodified sources built from a generated build directory whose ONLY generated file is ExternVariable.h with the four
"change with correct value" lines set to 640x480, fx = fy = 512 (recipe: `make -C oracle ref640`; nothing is copied into the
repository).  The fixture holds two seeded pairs (one started at the identity, one from a perturbed pose) and what the
reference produced for them: selected-pixel counts, iteration counts, hessian / sd_param / weightedPose / pose of every
iteration, the driver's result and its poseWrtWorld post-condition.  tests/test_reference_pin.py checks the oracle against
it (CPU), tests/test_gpu_parity.py the CUDA path (GPU box, where /root/reference does not exist).

To keep the file small only the level-0 keyframe depth is stored; the depth / variance pyramids are rebuilt from it by the
generator's own buildInvVarDepth restatement (synth.build_inv_var_depth, plain float32 numpy) and their SHA-256 digests are
stored, so a reader can tell that it rebuilt exactly the arrays the reference was given.

    python tests/golden/make_reference_golden_640x480.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)


def rebuild_pyramids(depth0, variance=0.01):
    from egomotion_with_local_loop_closures_b200 import synth
    var0 = np.where(depth0 > 0, np.float32(variance), np.float32(-1)).astype(np.float32)
    return synth.build_inv_var_depth(depth0, var0)


def main():
    from egomotion_with_local_loop_closures_b200 import synth
    from make_reference_golden import digest
    from oracle import refbinding as ref
    prev = ref.select("640x480")
    try:
        k = ref.dims()
        assert (k["width"], k["height"]) == (640, 480)
        kk = dict(fx=k["fx"], fy=k["fy"], cx=k["cx"], cy=k["cy"])
        scene = synth.SynthScene(640, 480, k=kk)
        kf = scene.keyframe(noise_seed=61)
        depth, var = rebuild_pyramids(kf["depth"][0])
        for l in range(4):
            assert np.array_equal(depth[l], kf["depth"][l]) and np.array_equal(var[l], kf["var"][l]), l
        rng = np.random.default_rng(61)
        gt = [synth.random_pose(rng, rot=np.deg2rad(1.2), trans=0.015).astype(np.float32), synth.random_pose(rng, rot=np.deg2rad(2.5), trans=0.03).astype(np.float32)]
        frames = [scene.render(synth.se3_exp(p), noise_seed=6100 + i) for i, p in enumerate(gt)]
        inits = [np.zeros(6, np.float32), (gt[1] * 0.6).astype(np.float32)]
        out = dict(width=np.array([640]), height=np.array([480]), intr=np.array([k["fx"], k["fy"], k["cx"], k["cy"]], np.float32),
                   kf_image=kf["image"], depth0=kf["depth"][0], frames=np.stack(frames), gt=np.stack(gt), init=np.stack(inits),
                   depth_sha=np.array([digest(a) for a in depth]), var_sha=np.array([digest(a) for a in var]))
        for i, (fr, init) in enumerate(zip(frames, inits)):
            tr = ref.track_trace(kf["image"], fr, depth, var, init)
            out[f"p{i}_n_selected"] = np.array(tr["n_selected"], np.int32)
            out[f"p{i}_n_iters"] = np.array(tr["n_iters"], np.int32)
            out[f"p{i}_pose"] = tr["final_pose"]
            for l in range(4):
                its = tr["levels"][l]
                out[f"p{i}_H_{l}"] = np.stack([it["H"] for it in its])
                out[f"p{i}_b_{l}"] = np.stack([it["b"] for it in its])
                out[f"p{i}_wp_{l}"] = np.array([it["weighted_pose"] for it in its], np.float32)
                out[f"p{i}_pose_{l}"] = np.stack([it["pose_after"] for it in its])
            pose, po, pw = ref.get_image_pose_estimate(kf["image"], fr, depth, var, init)
            out[f"p{i}_driver_pose"] = pose
            out[f"p{i}_driver_pose_wrt_world"] = pw
        path = os.path.join(HERE, "reference_track_640x480.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")
    finally:
        ref.select(prev)


if __name__ == "__main__":
    main()
