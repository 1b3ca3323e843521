"""CPU: the parts of bench.py's output contract that can be checked without a GPU -- the reference arm (`--impl reference`) runs the
CPU oracle on a bounded sample of the workload and must print the keys the driver reads, with a `config` that names the WORKLOAD
only (the same dict our arm prints for the same flags: bench.workload_config)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "6", "--keyframes", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tracks/s" and line["higher_is_better"] is True
    assert line["metric"] == "frame-keyframe GN tracks/sec at 640x480"
    assert line["gpu_launches"] == 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert cb["reference_faithful"]["cores"] == 3 and cb["reference_faithful"]["value"] > 0      # NUM_POSE_THREADS, src/ExternVariable.h:224
    # the workload, and nothing but the workload
    assert set(line["config"]) == {"workload", "width", "height", "keyframes_per_gpu", "frames_per_gpu", "pairs_per_gpu_per_step",
                                   "pairs_per_frame", "l2"}
    assert line["config"]["workload"] == "pair_sweep_640x480" and (line["config"]["width"], line["config"]["height"]) == (640, 480)
    assert line["implementation"]["sample_pairs_per_step"] <= line["config"]["pairs_per_gpu_per_step"]

    sys.path.insert(0, ROOT)
    import bench
    bench.W, bench.H = 640, 480
    want = bench.workload_config(bench.CONFIGS["pair_sweep"], 2, 6, line["config"]["pairs_per_frame"], line["config"]["pairs_per_gpu_per_step"])
    assert want == line["config"]
    # inputs of the default workload: 512 frames + 32 keyframes (image, 4-level depth and variance) + 4608 pair records
    assert bench.input_bytes_per_step(512, 32, 4608) == 271730688
