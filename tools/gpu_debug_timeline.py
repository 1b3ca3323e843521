"""Diagnostic (GPU box): device timeline of the pipelined e2e loop (events on the library's own streams)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
nk, nf = 16, 256
wl = bench.build_workload(nk, nf, 9, seed=0)
import torch
from egomotion_with_local_loop_closures_b200 import capi, synth
k = synth.intrinsics(640, 480)
trk = capi.Tracker(capi.default_config(640, 480, fx=float(k["fx"]), fy=float(k["fy"]), cx=float(k["cx"]), cy=float(k["cy"]), max_keyframes=2 * nk, max_frames=2 * nf))
dev = torch.device("cuda", 0)
S = [torch.cuda.ExternalStream(trk.stream_of(i), device=dev) for i in range(3)]
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep = []; frames = []; kimg = []; kd = []; kv = []
for f in wl["frames"]:
    t, a = pin(f); keep.append(t); frames.append(a)
for i in range(nk):
    t, a = pin(wl["kf_images"][i]); keep.append(t); kimg.append(a)
    d, v = [], []
    for l in range(4):
        t, a = pin(wl["kf_depth"][i][l]); keep.append(t); d.append(a)
        t, a = pin(wl["kf_var"][i][l]); keep.append(t); v.append(a)
    kd.append(d); kv.append(v)
def upload(half):
    for i in range(nk): trk.upload_keyframe(half * nk + i, kimg[i], kd[i], kv[i])
    for i in range(nf): trk.upload_frame(half * nf + i, frames[i])
pairs = [trk.make_pairs(wl["kf_idx"] + h * nk, wl["fr_idx"] + h * nf, wl["init"]) for h in (0, 1)]
n = len(pairs[0])
upload(0); upload(1); trk.synchronize(); trk.track_batch(pairs[0]); trk.track_batch(pairs[1])
ev = lambda: torch.cuda.Event(enable_timing=True)
base = ev(); base.record(S[0]); torch.cuda.synchronize()
rows = []; pend = None
for s in range(6):
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record(S[1]); upload(s & 1); b.record(S[1])
    c.record(S[0]); dp = trk.track_batch_async(pairs[s & 1]); d.record(S[0])
    if pend is not None: trk.results_download(pend, n)
    pend = dp; rows.append((a, b, c, d))
trk.results_download(pend, n); torch.cuda.synchronize()
for s, (a, b, c, d) in enumerate(rows):
    print(f"step {s}: upload [{base.elapsed_time(a):7.2f} .. {base.elapsed_time(b):7.2f}]  compute-stream reaches step at {base.elapsed_time(c):7.2f}, track done {base.elapsed_time(d):7.2f}  kernel {trk.last_track_kernel_ms():.2f}")
