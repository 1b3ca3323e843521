"""Shared builders for the parity tests: seeded synthetic cases and oracle/GPU configuration twins."""
import numpy as np

from egomotion_with_local_loop_closures_b200 import synth


def make_case(width, height, n_frames=3, seed=11, rot=np.deg2rad(1.2), trans=0.015):
    """One keyframe at the world origin and n_frames frames at seeded random small poses."""
    scene = synth.SynthScene(width, height)
    kf = scene.keyframe(noise_seed=seed)
    rng = np.random.default_rng(seed)
    frames, gt = [], []
    for i in range(n_frames):
        p = synth.random_pose(rng, rot=rot, trans=trans)
        frames.append(scene.render(synth.se3_exp(p), noise_seed=seed * 100 + i))
        gt.append(p.astype(np.float32))
    return dict(width=width, height=height, scene=scene, kf=kf, frames=frames, gt=gt)


def oracle_config(oracle, case, **over):
    k = synth.intrinsics(case["width"], case["height"])
    return oracle.default_config(case["width"], case["height"], fx=float(k["fx"]), fy=float(k["fy"]),
                                 cx=float(k["cx"]), cy=float(k["cy"]), **over)


def gpu_config(capi, case, **over):
    k = synth.intrinsics(case["width"], case["height"])
    return capi.default_config(case["width"], case["height"], fx=float(k["fx"]), fy=float(k["fy"]),
                               cx=float(k["cx"]), cy=float(k["cy"]), **over)


def full_H(h21):
    H = np.zeros((6, 6), np.float64)
    k = 0
    for i in range(6):
        for j in range(i, 6):
            H[i, j] = H[j, i] = h21[k]
            k += 1
    return H


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---- measured-deviation log ----------------------------------------------------------------------------------------------
# The GPU parity tests append what they MEASURED (worst deviation per quantity) to gpurun_out/parity_metrics.jsonl when that
# directory exists, so that the tolerances written in the tests can be read next to the numbers a run actually produced
# (DESIGN.md section 2 quotes them).
def record(name, value, **extra):
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if not os.path.isdir(d):
        return
    row = dict(name=name, value=float(value), **extra)
    with open(os.path.join(d, "parity_metrics.jsonl"), "a") as f:
        f.write(json.dumps(row) + "\n")


def iters_match(got, want, arith):
    """Executed GN iterations per level: identical in both flavours.  (SURVEY.md section 7 would tolerate +-1 for the FAST flavour,
    whose K5 uses the closed-form small-angle exp / log; measured on B200 the counts are identical on every test case, and the
    kernel is run-to-run deterministic, so the tests assert equality.)"""
    return [int(v) for v in got] == [int(v) for v in want]
