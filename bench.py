#!/usr/bin/env python
"""bench.py -- frame-keyframe GN tracks/sec at 640x480 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N --steps K --warmup W]            # our arm (CUDA path through the C-ABI)
    python bench.py --impl reference [...]                       # reference arm: the CPU tracker on the host cores

A "step" is one batch of the hot path on one GPU: `frames` new frames (pyramid + gradient texels), `keyframes`
keyframes (pyramid + mask/count/selection) and `pairs` frame-keyframe tracks (each frame against its own keyframe and
K-1 local-loop-closure candidates), i.e. BASELINE config 5 / config 3 at 640x480.  Weak scaling: every rank gets the same
amount of work; ranks shard a global pair list by keyframe affinity and all-gather the 256-byte result records (NCCL).

Timed regions
  value : inputs resident in HBM (level-0 u8 images, keyframe depth/variance pyramids, pair list on the host);
          prepare_frames + prepare_keyframes + track_batch (+ result D2H, + NCCL all-gather for N>1).
  e2e   : the same through the reference-facing C-ABI with HOST (pinned) buffers: H2D of every image and depth/variance
          pyramid and D2H of the results inside the timed region.
Inputs (hundreds of MB per step) are larger than the 126 MB L2, so no explicit L2 flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Keep stdout for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1 behind Python's back, so fd 1 is
# pointed at stderr for the whole run and the result line goes to a private duplicate of the original stdout.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


from egomotion_with_local_loop_closures_b200 import synth  # noqa: E402

W, H = 640, 480
METRIC = "frame-keyframe GN tracks/sec at 640x480"
UNIT = "tracks/s"


# ----------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------
def _render_job(args):
    kind, seed_scene, T, seed = args
    scene = _render_job.scenes.get(seed_scene)
    if scene is None:
        scene = _render_job.scenes[seed_scene] = synth.SynthScene(W, H, seed_tex=seed_scene)
    if kind == "kf":
        kf = scene.keyframe(T, seed_depth=seed, noise_seed=seed + 1)
        return kf["image"], kf["depth"], kf["var"]
    return scene.render(T, noise_seed=seed)


_render_job.scenes = {}


def build_workload(n_kf, n_frames, pairs_per_frame, seed, workers=None):
    """Seeded pool of keyframes / frames / pairs for one rank.  Rendering is numpy (untimed setup)."""
    rng = np.random.default_rng(seed)
    T_kf = [synth.se3_exp(synth.random_pose(rng, rot=np.deg2rad(2.0), trans=0.04)) for _ in range(n_kf)]
    primary = np.arange(n_frames) % n_kf
    T_fr = [synth.se3_exp(synth.random_pose(rng, rot=np.deg2rad(1.0), trans=0.015)) @ T_kf[primary[i]] for i in range(n_frames)]
    jobs = [("kf", 1234 + seed, T_kf[k], 5678 + 17 * k + seed) for k in range(n_kf)]
    jobs += [("fr", 1234 + seed, T_fr[i], 91011 + i + 1000 * seed) for i in range(n_frames)]
    workers = workers or min(os.cpu_count() or 1, 32)
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            out = pool.map(_render_job, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
    else:
        out = [_render_job(j) for j in jobs]
    kfs, frames = out[:n_kf], out[n_kf:]
    kf_idx, fr_idx, init = [], [], []
    for i in range(n_frames):
        others = [k for k in range(n_kf) if k != primary[i]]
        cand = [int(primary[i])] + list(rng.choice(others, size=min(pairs_per_frame - 1, len(others)), replace=False))
        for k in cand:
            rel = synth.relative_pose(T_fr[i], T_kf[k])
            # init = pose of the "previous frame" (src/ImageFunc.cpp:106): ground truth perturbed by a small motion
            init.append(rel + synth.random_pose(rng, rot=np.deg2rad(0.3), trans=0.004))
            kf_idx.append(k)
            fr_idx.append(i)
    perm = rng.permutation(len(kf_idx))
    return dict(kf_images=[k[0] for k in kfs], kf_depth=[k[1] for k in kfs], kf_var=[k[2] for k in kfs], frames=frames,
                kf_idx=np.array(kf_idx, np.int32)[perm], fr_idx=np.array(fr_idx, np.int32)[perm],
                init=np.array(init, np.float32)[perm])


def algorithmic_bytes(res):
    """per GN iteration at level L, B_iter(L) = 2 P_L + 9 N_L; summed over executed iterations of all tracks."""
    P = np.array([(W >> l) * (H >> l) for l in range(4)], np.float64)
    it = res["n_iters"].astype(np.float64)
    n = res["n_selected"].astype(np.float64)
    return float((it * (2.0 * P[None, :] + 9.0 * n)).sum())


def setup_bytes(n_frames, n_kf):
    """B_setup: keyframe depth read + mask write (5 P_L) and pyramid build (P_{L-1} + P_L) per frame / keyframe."""
    P = np.array([(W >> l) * (H >> l) for l in range(4)], np.float64)
    pyr = float(sum(P[l - 1] + P[l] for l in range(1, 4)))
    return n_frames * pyr + n_kf * (pyr + 5.0 * P.sum())


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread (5 ms period, no start-up latency);
    falls back to `nvidia-smi -lms 50` when pynvml is unavailable.  mark() brackets the timed region; samples outside are dropped."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.stop_flag, self.t0, self.t1, self.nvml = gpu_index, [], None, False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.nvml[1], pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, hd = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(hd))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(hd))
                self.rows.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        inside = lambda t: (self.t0 is None or t >= self.t0) and (self.t1 is None or t <= self.t1)
        if self.nvml:
            self.thread.join(timeout=1)
            rows = [r for r in self.rows if inside(r[0])] or self.rows[-3:]
            reasons = sorted({name for _, _, m in rows for bit, name in self.REASONS.items() if m & bit})
            sm = [r[1] for r in rows]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm), "reasons": reasons,
                    "source": "nvml, 5 ms period, timed region only"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 50 (warm-up + timed region)"}


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Reference arm: the reference's CPU tracker (the oracle restatement -- the real one cannot be compiled here,
    DESIGN.md) on all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    n_pairs = max(cores * 64, 256)                        # ~3 s of CPU work per step
    n_frames = max(8, n_pairs // args.pairs_per_frame)
    wl = build_workload(min(args.keyframes, 8), n_frames, args.pairs_per_frame, seed=0)
    n_pairs = min(n_pairs, len(wl["kf_idx"]))
    k = synth.intrinsics(W, H)
    cfg = oracle.default_config(W, H, fx=float(k["fx"]), fy=float(k["fy"]), cx=float(k["cx"]), cy=float(k["cy"]))
    times = []
    for step in range(args.warmup + args.steps):
        _, secs = oracle.track_many(cfg, wl["kf_idx"][:n_pairs], wl["fr_idx"][:n_pairs], wl["kf_images"], wl["frames"],
                                    wl["kf_depth"], wl["kf_var"], wl["init"][:n_pairs], n_workers=cores)
        if step >= args.warmup:
            times.append(secs)
    total = float(sum(times))
    value = n_pairs * len(times) / total
    sample = f"{n_pairs} pairs/step of the pair_sweep_640x480 workload, {cores} worker threads (1 pair per thread, bands sequential)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "pair_sweep_640x480", "width": W, "height": H, "pairs_per_step": n_pairs,
                       "note": "CPU oracle restatement of the reference tracker (g++ -std=c++11 -O3); omits the reference's per-pixel cv::Mat/cv::String overhead, so it is faster than the real binary"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--keyframes", type=int, default=32, help="keyframes per GPU per step")
    ap.add_argument("--frames", type=int, default=512, help="frames per GPU per step")
    ap.add_argument("--pairs-per-frame", type=int, default=9, help="1 sequential + K=8 loop-closure candidates")
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"])
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--lc-mode", default="forward", choices=["forward", "const_weight"],
                    help="const_weight: the K-1 loop-closure pairs of every frame run the reference's constant-weight "
                         "inverse-compositional tracker (FLAG_DO_CONST_WEIGHT_POSE_ESTIMATION); not the headline workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # ---- data first (fork-based rendering must precede CUDA initialisation)
    t_setup = time.time()
    # every rank renders its own segment: share the host cores between the ranks of the node
    wl = build_workload(args.keyframes, args.frames, args.pairs_per_frame, seed=rank,
                        workers=max(1, min(32, (os.cpu_count() or 1) // max(1, world))))
    n_pairs = len(wl["kf_idx"])

    import torch
    import torch.distributed as dist
    from egomotion_with_local_loop_closures_b200 import capi
    from egomotion_with_local_loop_closures_b200.sharding import gather_results, shard_pairs

    torch.cuda.set_device(local_rank)
    if world > 1:
        # The only collective is the gather of 256-byte result records (1.2 MB per rank and step): latency-bound, and it runs
        # WHILE the tracking kernel of the next step owns the SMs.  NCCL's defaults (many channels, Simple protocol) park a CTA on
        # many SMs while ranks wait for each other, and every such SM loses one of its two resident tracking CTAs (measured at
        # N = 2: tracking kernel 14.1 -> 16.9 ms, 92 % scaling).  One channel + the low-latency protocol: 14.1 ms, 98-99 %.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        os.environ.setdefault("NCCL_PROTO", "LL")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    k = synth.intrinsics(W, H)
    cfg = capi.default_config(W, H, fx=float(k["fx"]), fy=float(k["fy"]), cx=float(k["cx"]), cy=float(k["cy"]),
                              max_keyframes=2 * args.keyframes, max_frames=2 * args.frames, device=local_rank,
                              arithmetic=capi.ARITH_STRICT if args.arith == "strict" else capi.ARITH_FAST,
                              ctas_per_pair=args.cluster)
    trk = capi.Tracker(cfg)
    stream = torch.cuda.ExternalStream(trk.stream(), device=torch.device("cuda", local_rank))

    # Global pair list = the sequence segments of all ranks (keyframe / frame ids offset per segment).  shard_pairs deals whole
    # connected components (segments) to ranks, so every rank gets one segment back -- the production path for a mixed list.
    if world > 1:
        g_kf = np.concatenate([wl["kf_idx"].astype(np.int64) + r * args.keyframes for r in range(world)])
        g_fr = np.concatenate([wl["fr_idx"].astype(np.int64) + r * args.frames for r in range(world)])
        shards = shard_pairs(g_kf, g_fr, world)
        seg = [int(g_kf[s][0] // args.keyframes) for s in shards]
        assert sorted(seg) == list(range(world)) and all(len(s) == n_pairs for s in shards)
        my_idx = shards[seg.index(rank)]                       # this rank rendered segment `rank`
    else:
        my_idx = np.arange(n_pairs)
    n_total = n_pairs * world

    # pinned host copies (e2e uploads) -----------------------------------------------------------------------------
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = []
    h_frames, h_kf_img, h_kf_depth, h_kf_var = [], [], [], []
    for f in wl["frames"]:
        t, a = pin(f); keep.append(t); h_frames.append(a)
    for i in range(args.keyframes):
        t, a = pin(wl["kf_images"][i]); keep.append(t); h_kf_img.append(a)
        d, v = [], []
        for l in range(4):
            t, a = pin(wl["kf_depth"][i][l]); keep.append(t); d.append(a)
            t, a = pin(wl["kf_var"][i][l]); keep.append(t); v.append(a)
        h_kf_depth.append(d); h_kf_var.append(v)
    h2d_bytes = sum(a.nbytes for a in h_frames) + sum(a.nbytes for a in h_kf_img) + sum(a.nbytes for d in h_kf_depth for a in d) + \
        sum(a.nbytes for d in h_kf_var for a in d) + n_pairs * capi.PAIR_DTYPE.itemsize
    d2h_bytes = n_pairs * capi.RESULT_DTYPE.itemsize

    def upload_all(half=0):
        ko, fo = half * args.keyframes, half * args.frames
        for i in range(args.keyframes):
            trk.upload_keyframe(ko + i, h_kf_img[i], h_kf_depth[i], h_kf_var[i])
        for i in range(args.frames):
            trk.upload_frame(fo + i, h_frames[i])

    lc = args.lc_mode == "const_weight"
    primary = (wl["kf_idx"] == (wl["fr_idx"] % args.keyframes))
    flags = np.where(primary, capi.PAIR_SAVE_WEIGHTS, capi.PAIR_CONST_WEIGHT).astype(np.int32) if lc else 0
    pairs = trk.make_pairs(wl["kf_idx"], wl["fr_idx"], wl["init"], flags=flags)
    pairs_half = [pairs, trk.make_pairs(wl["kf_idx"] + args.keyframes, wl["fr_idx"] + args.frames, wl["init"])]
    if lc:
        args.no_e2e = True                                     # an upload invalidates the loop-closure records: resident mode only
        args.no_cpu_baseline = True                            # (the forward CPU baseline is not this workload)
    fr_slots = np.arange(args.frames, dtype=np.int32)
    kf_slots = np.arange(args.keyframes, dtype=np.int32)

    upload_all()
    if not lc:
        upload_all(1)                                          # the forward workload is resident twice: steps alternate between the halves
    trk.synchronize()
    if lc:
        # untimed setup, as in the reference's flow: every keyframe's weight pyramid = average of the last-iteration weights of
        # the sequential tracks of its frames (saveWeights / finaliseWeights), then the loop-closure records
        seq = trk.make_pairs(wl["kf_idx"][primary], wl["fr_idx"][primary], wl["init"][primary], flags=capi.PAIR_SAVE_WEIGHTS)
        trk.track_batch(seq)
        for kslot in range(args.keyframes):
            trk.reset_keyframe_weights(kslot)
            fr = np.sort(wl["fr_idx"][primary][wl["kf_idx"][primary] == kslot])
            for lo in range(0, len(fr), 64):
                trk.accumulate_weights(kslot, fr[lo:lo + 64])
            trk.finalise_weights(kslot)
        trk.prepare_keyframes_lc(np.arange(args.keyframes, dtype=np.int32))
        trk.synchronize()
    setup_s = time.time() - t_setup

    kernel_ms = []

    # N > 1: the NCCL all-gather of the 256-byte result records of step k runs (torch's stream) while the kernels of step k+1 run
    # (the library's compute stream); the library keeps the records of the last two batches.  The shard sizes are exchanged once.
    mg = {"pending": None, "last": None, "rec": None, "counts": None, "idx": None}

    def gather_step(pend):
        dptr, ev = pend
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)                                    # batch k is complete before its records are read
        from cuda import cudart  # cuda-python is in the image
        err, = cudart.cudaMemcpyAsync(mg["rec"].data_ptr(), dptr, n_pairs * 256, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, cur.cuda_stream)
        assert int(err) == 0
        gathered = gather_results(mg["rec"], my_idx, n_total, counts=mg["counts"])
        return gathered[mg["idx"]].cpu().numpy().view(capi.RESULT_DTYPE).reshape(-1)

    # Forward workload: steps alternate between two resident copies of the inputs (slot halves), so that the preparation of step
    # k+1 (pyramids, texels, selection lists: ellc_prepare_async, low-priority stream) overlaps the tracking kernel of step k, and
    # the records of step k are fetched while step k+1 runs.  Every step still does all of its own work.
    pipe = {"k": 0, "pending": None, "last": None}

    def finish_previous(pend):
        if world > 1:
            res_prev = gather_step(pend)
        else:
            res_prev = trk.results_download(pend[0], n_pairs)
        kernel_ms.append(trk.batch_kernel_ms(1))              # the batch before the one just launched: its events are complete
        return res_prev

    def step_resident_pipelined():
        half = pipe["k"] & 1
        trk.prepare_async(fr_slots + half * args.frames, kf_slots + half * args.keyframes)
        if world > 1 and mg["rec"] is None:
            mg["rec"] = torch.empty((n_pairs, 256), dtype=torch.uint8, device="cuda")
            mg["idx"] = torch.as_tensor(my_idx, device="cuda")
            cnt = torch.tensor([n_pairs], dtype=torch.int64, device="cuda")
            allc = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(allc, cnt)
            mg["counts"] = [int(c.item()) for c in allc]
        dptr = trk.track_batch_async(pairs_half[half])
        ev = torch.cuda.Event()
        ev.record(stream)
        if pipe["pending"] is not None:
            pipe["last"] = finish_previous(pipe["pending"])
        pipe["pending"] = (dptr, ev)
        pipe["k"] += 1
        return pipe["last"]

    def drain_resident_pipelined():
        if pipe["pending"] is not None:
            if world > 1:
                pipe["last"] = gather_step(pipe["pending"])
            else:
                pipe["last"] = trk.results_download(pipe["pending"][0], n_pairs)
            kernel_ms.append(trk.batch_kernel_ms(0))
            pipe["pending"] = None
        return pipe["last"]

    def step_resident():
        trk.prepare_frames(fr_slots)
        trk.prepare_keyframes(kf_slots)
        if lc:
            trk.prepare_keyframes_lc(kf_slots)                 # the per-keyframe Jacobians / hessians are part of the step
        if world > 1:
            if mg["rec"] is None:
                mg["rec"] = torch.empty((n_pairs, 256), dtype=torch.uint8, device="cuda")
                mg["idx"] = torch.as_tensor(my_idx, device="cuda")
                cnt = torch.tensor([n_pairs], dtype=torch.int64, device="cuda")
                allc = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
                dist.all_gather(allc, cnt)
                mg["counts"] = [int(c.item()) for c in allc]
            dptr = trk.track_batch_async(pairs)
            ev = torch.cuda.Event()
            ev.record(stream)
            if mg["pending"] is not None:
                mg["last"] = gather_step(mg["pending"])
            mg["pending"] = (dptr, ev)
            return mg["last"]
        res = trk.track_batch(pairs)
        kernel_ms.append(trk.last_track_kernel_ms())
        return res

    def drain_resident():
        if world > 1 and mg["pending"] is not None:
            mg["last"] = gather_step(mg["pending"])
            mg["pending"] = None
            kernel_ms.append(trk.last_track_kernel_ms())
        return mg["last"] if world > 1 else None

    # e2e: every step uploads its inputs from pinned host memory and downloads its result records.  Steps alternate
    # between two halves of the slot pools, so the uploads of step k+1 (copy stream) overlap the kernels of step k;
    # the results of step k are fetched while step k+1 runs.
    e2e_state = {"k": 0, "pending": None, "last": None}

    def step_e2e():
        half = e2e_state["k"] & 1
        t0 = time.perf_counter()
        upload_all(half)
        t1 = time.perf_counter()
        dptr = trk.track_batch_async(pairs_half[half])
        e2e_state["host_upload_ms"] = e2e_state.get("host_upload_ms", 0.0) + 1e3 * (t1 - t0)
        e2e_state["host_launch_ms"] = e2e_state.get("host_launch_ms", 0.0) + 1e3 * (time.perf_counter() - t1)
        if e2e_state["pending"] is not None:
            e2e_state["last"] = trk.results_download(e2e_state["pending"], n_pairs)
        e2e_state["pending"] = dptr
        e2e_state["k"] += 1
        return e2e_state["last"]

    def drain_e2e():
        if e2e_state["pending"] is not None:
            e2e_state["last"] = trk.results_download(e2e_state["pending"], n_pairs)
            e2e_state["pending"] = None
        return e2e_state["last"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False, drain=None):
        res = None
        clk = ClockSampler(local_rank) if sample_clocks else None
        if clk:
            clk.start()                                   # sampled over warm-up + timed steps (the same load)
        for _ in range(warmup):
            res = fn()
        if drain:
            r2 = drain()
            res = r2 if r2 is not None else res
        barrier()
        kernel_ms.clear()
        trk.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if clk:
            clk.mark_begin()
        e0.record(stream)
        for _ in range(steps):
            res = fn()
        if drain:
            r2 = drain()
            res = r2 if r2 is not None else res
        stream.wait_stream(torch.cuda.current_stream())   # collectives / copies issued on torch's stream are inside the timed region
        e1.record(stream)
        barrier()
        if clk:
            clk.mark_end()
        ms = e0.elapsed_time(e1)
        launches = trk.launch_count()
        clocks = clk.stop() if clk else None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, launches, clocks

    if lc:
        ms, res, launches, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True, drain=drain_resident)
    else:
        ms, res, launches, clocks = timed(step_resident_pipelined, args.steps, args.warmup, sample_clocks=True, drain=drain_resident_pipelined)
    k_ms = float(np.mean(kernel_ms))
    value = n_total * args.steps / (ms * 1e-3)
    alg = algorithmic_bytes(res)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg / (k_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel: one `ncu --set full` capture of this command (profiles/r01_traffic.json, made by
    # tools/profile_summary.py), bytes per launch like the algorithmic figure; scaled by pair count if the workload differs.
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * n_pairs / tj["pairs_per_launch"]
        traffic_src = "profiles/r01_traffic.json (%s; dram__bytes_read + write of one launch of %d pairs%s)" % (
            tj["report"], tj["pairs_per_launch"], "" if n_pairs == tj["pairs_per_launch"] else ", scaled by pair count")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "bytes per launch (compare with algorithmic_bytes_per_launch)", "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel": "gn_track_kernel", "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg,
                "kernel_share_of_step": k_ms / (ms / args.steps),
                "mean_iters_per_level": [float(x) for x in res["n_iters"].mean(axis=0)],
                "mean_selected_per_level": [float(x) for x in res["n_selected"].mean(axis=0)]}

    e2e = None
    if not args.no_e2e:
        ems, eres, _, _ = timed(step_e2e, args.steps, max(1, args.warmup), drain=drain_e2e)
        e2e = {"value": n_total * args.steps / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes) * world,
               "d2h_bytes_per_step": int(d2h_bytes) * world, "ms_per_step": ems / args.steps,
               "host_enqueue_ms_per_step": {"uploads": e2e_state.get("host_upload_ms", 0.0) / max(1, e2e_state["k"]),
                                            "track_launch": e2e_state.get("host_launch_ms", 0.0) / max(1, e2e_state["k"])},
               "pipelining": "2 slot halves alternate; uploads of step k+1 on the copy stream overlap the kernels of step k"}
        assert np.array_equal(eres["pose"], res["pose"]), "e2e and resident paths disagree"

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        cores = os.cpu_count() or 1
        ns = min(n_pairs, max(cores * 256, 1024))           # ~10-15 s of CPU work on the box's cores
        ocfg = oracle.default_config(W, H, fx=float(k["fx"]), fy=float(k["fy"]), cx=float(k["cx"]), cy=float(k["cy"]))
        oposes, secs = oracle.track_many(ocfg, wl["kf_idx"][:ns], wl["fr_idx"][:ns], wl["kf_images"], wl["frames"], wl["kf_depth"],
                                         wl["kf_var"], wl["init"][:ns], n_workers=cores)
        perr = float(np.abs(oposes - res["pose"][:ns]).max())
        cpu_baseline = {"value": ns / secs, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {ns} pairs of this step's pair list, {cores} worker threads (1 pair per thread); "
                                  f"max |pose_gpu - pose_cpu| on the sample = {perr:.2e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "pair_sweep_640x480" + ("_lc_const_weight" if lc else ""), "width": W, "height": H, "keyframes_per_gpu": args.keyframes,
                           "frames_per_gpu": args.frames, "pairs_per_gpu_per_step": n_pairs, "pairs_per_frame": args.pairs_per_frame,
                           "arithmetic": args.arith,
                           "pipelining": ("none (loop-closure mode: one set of slots)" if lc else
                                          "resident inputs held twice; steps alternate between the two sets of slots: preparation of step k+1 on a "
                                          "low-priority stream overlaps the tracking kernel of step k, records of step k fetched during step k+1"),
                           "parallelism": f"pair list sharded by connected components (sequence segments) x{world}, NCCL all-gather of 256 B result records",
                           "l2": "inputs (%.0f MB per step per GPU) exceed the 126 MB L2; no flush" % (h2d_bytes / 1e6),
                           "setup_bytes_per_step": setup_bytes(args.frames, args.keyframes) * world, "setup_seconds": setup_s},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        emit(line)
    trk.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
