"""Diagnostic (GPU box): iteration-by-iteration comparison of a full track with the oracle trace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from egomotion_with_local_loop_closures_b200 import capi
from tests.helpers import make_case, oracle_config, gpu_config

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
case = make_case(w, h, n_frames=2, seed=23)
ocfg = oracle_config(oracle, case)
for arith in (1, 0):
    t = capi.Tracker(gpu_config(capi, case, max_keyframes=1, max_frames=2, arithmetic=arith))
    t.upload_keyframe(0, case["kf"]["image"], case["kf"]["depth"], case["kf"]["var"])
    for i, f in enumerate(case["frames"]):
        t.upload_frame(i, f)
    res, tr = t.track_batch(t.make_pairs([0, 0], [0, 1]), want_trace=True)
    print(f"=== arith={'strict' if arith else 'fast'}")
    for i in range(2):
        opose, otr = oracle.track(ocfg, case["kf"]["image"], case["frames"][i], case["kf"]["depth"], case["kf"]["var"], np.zeros(6, np.float32))
        print(f"pair {i}: gpu iters {list(res[i]['n_iters'])} oracle {otr['n_iters']}  |pose diff| {np.abs(res[i]['pose']-opose).max():.2e}")
        for l in (3, 2, 1, 0):
            for k, o in enumerate(otr["levels"][l]):
                g = tr[i, l, k]
                if not g["executed"]:
                    print(f"  L{l} it{k}: gpu did not execute"); continue
                gH = np.array(g["H"], np.float64).reshape(6, 6)
                bs = np.sqrt(np.diag(o["H_f64"]) * o["res_sum_f64"])
                print(f"  L{l} it{k}: res rel {abs(float(g['res_sum'])-o['res_sum_f64'])/o['res_sum_f64']:.1e}  H {np.abs(gH-o['H_f64']).max()/np.abs(o['H_f64']).max():.1e}"
                      f"  b {(np.abs(np.array(g['b'],np.float64)-o['b_f64'])/bs).max():.1e}  delta {np.abs(g['delta']-o['delta']).max():.1e} (|d| {np.abs(o['delta']).max():.1e})"
                      f"  pose {np.abs(g['pose_after']-o['pose_after']).max():.1e}  wp {float(g['weighted_pose']):.3f}/{o['weighted_pose']:.3f} oob {int(g['n_oob'])}/{o['n_oob']}")
    t.close()
