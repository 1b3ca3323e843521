"""Diagnostic (GPU box): does the upload of step k+1 overlap the track kernel of step k?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
wl = bench.build_workload(16, 256, 9, seed=0)
import torch
from egomotion_with_local_loop_closures_b200 import capi, synth
k = synth.intrinsics(640, 480)
nk, nf = 16, 256
trk = capi.Tracker(capi.default_config(640, 480, fx=float(k["fx"]), fy=float(k["fy"]), cx=float(k["cx"]), cy=float(k["cy"]), max_keyframes=2 * nk, max_frames=2 * nf))
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
nbytes = sum(a.nbytes for a in frames) + sum(a.nbytes for a in kimg) + sum(a.nbytes for d in kd for a in d) * 2
upload(0); upload(1); trk.synchronize(); trk.track_batch(pairs[0]); trk.track_batch(pairs[1])
# 1. upload only
t0 = time.perf_counter(); upload(0); t1 = time.perf_counter(); trk.synchronize(); t2 = time.perf_counter()
print(f"upload only: host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms, {nbytes/1e6:.0f} MB -> {nbytes/(t2-t0)/1e9:.1f} GB/s")
# 2. track only (resident, slots prepared)
t0 = time.perf_counter(); trk.track_batch(pairs[0]); t1 = time.perf_counter()
print(f"track only: {1e3*(t1-t0):.2f} ms, kernel {trk.last_track_kernel_ms():.2f} ms")
# 3. serial: upload + track each step
t0 = time.perf_counter()
for s in range(4):
    upload(s & 1); trk.track_batch(pairs[s & 1])
t1 = time.perf_counter()
print(f"serial upload+track: {1e3*(t1-t0)/4:.2f} ms/step, kernel {trk.last_track_kernel_ms():.2f} ms")
# 4. pipelined
pend = None; t0 = time.perf_counter(); marks = []
for s in range(6):
    a = time.perf_counter(); upload(s & 1); b = time.perf_counter()
    d = trk.track_batch_async(pairs[s & 1]); c = time.perf_counter()
    if pend is not None: trk.results_download(pend, n)
    e = time.perf_counter(); pend = d
    marks.append((1e3*(b-a), 1e3*(c-b), 1e3*(e-c)))
trk.results_download(pend, n); t1 = time.perf_counter()
print(f"pipelined: {1e3*(t1-t0)/6:.2f} ms/step, kernel {trk.last_track_kernel_ms():.2f} ms; per step (upload-enqueue, launch, wait-prev): " + " ".join(f"({x:.1f},{y:.1f},{z:.1f})" for x, y, z in marks))
