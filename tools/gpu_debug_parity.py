"""Diagnostic (GPU box): per-level comparison of the CUDA normal equations with the oracle at forced poses."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from egomotion_with_local_loop_closures_b200 import capi
from tests.helpers import make_case, oracle_config, gpu_config

w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (320, 240)
case = make_case(w, h, n_frames=2, seed=11)
ocfg = oracle_config(oracle, case)
kpyr = oracle.image_pyramid(case["kf"]["image"])
for arith in (1, 0):
    for cluster in (1, 4):
        t = capi.Tracker(gpu_config(capi, case, max_keyframes=1, max_frames=2, arithmetic=arith, ctas_per_pair=cluster))
        t.upload_keyframe(0, case["kf"]["image"], case["kf"]["depth"], case["kf"]["var"])
        for i, f in enumerate(case["frames"]):
            t.upload_frame(i, f)
        print(f"=== arith={'strict' if arith else 'fast'} cluster={cluster}")
        for fi in range(1):
            cpyr = oracle.image_pyramid(case["frames"][fi])
            for level in range(4):
                for name, pose in (("zero", np.zeros(6, np.float32)), ("gt", case["gt"][fi]), ("6gt", (case["gt"][fi] * 6).astype(np.float32))):
                    o = oracle.gn_evaluate(ocfg, level, kpyr[level], cpyr[level], case["kf"]["depth"][level], case["kf"]["var"][level], pose, want_weights=True)
                    g, gw = t.gn_evaluate(0, fi, level, pose, want_weights=True)
                    gH = np.array(g["H"], np.float64).reshape(6, 6); gb = np.array(g["b"], np.float64)
                    bscale = np.sqrt(np.diag(o["H_f64"]) * o["res_sum_f64"])
                    dW = np.abs(gw - o["weights"]); iw = np.unravel_index(dW.argmax(), dW.shape)
                    print(f"L{level} {name:4s} H/f64 {np.abs(gH-o['H_f64']).max()/np.abs(o['H_f64']).max():.1e} (orc {np.abs(o['H']-o['H_f64']).max()/np.abs(o['H_f64']).max():.1e})"
                          f" b/f64 {(np.abs(gb-o['b_f64'])/bscale).max():.1e} (orc {(np.abs(o['b']-o['b_f64'])/bscale).max():.1e})"
                          f" res {abs(float(g['res_sum'])-o['res_sum_f64'])/max(o['res_sum_f64'],1e-30):.1e} oob {int(g['n_oob'])}/{o['n_oob']}"
                          f" wmax {dW.max()/max(o['weights'].max(),1e-30):.1e} at {iw} nz {int((dW>0).sum())}/{int((o['weights']>0).sum())}")
        t.close()
