"""Turn ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/profile_summary.py kernel <rep.ncu-rep> "<title>" "<command>"     -> markdown table on stdout
    python tools/profile_summary.py launches <launches.csv> "<command>"             -> markdown table on stdout
    python tools/profile_summary.py traffic <rep.ncu-rep> <pairs> "<command>" [version]  -> json on stdout (bench.py reads it)
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    mode = sys.argv[1]
    if mode == "kernel":
        rep, title, cmd = sys.argv[2:5]
        val, unit = raw(rep)
        print(f"## {title}\n\n`{cmd}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in val:
                print(f"| {k} | {val[k]} | {unit[k]} |")
        print()
    elif mode == "launches":
        path, cmd = sys.argv[2:4]
        rows = [r for r in csv.reader(open(path)) if len(r) > 5]
        ix = {h: i for i, h in enumerate(rows[0])}
        agg, cnt = collections.Counter(), collections.Counter()
        for r in rows[1:]:
            try:
                v = float(r[ix["Metric Value"]].replace(",", ""))
            except ValueError:
                continue
            name = r[ix["Kernel Name"]].split("(")[0]
            agg[name] += v
            cnt[name] += 1
        tot = sum(agg.values())
        print(f"`{cmd}`\n\n| kernel | launches | total us | share |\n|---|---|---|---|")
        for k, v in agg.most_common():
            print(f"| {k} | {cnt[k]} | {v / 1e3:.1f} | {v / tot:.3f} |")
        print()
    elif mode == "traffic":
        rep, pairs, cmd = sys.argv[2], int(sys.argv[3]), sys.argv[4]
        val, unit = raw(rep)
        rd = to_bytes(val["dram__bytes_read.sum"], unit["dram__bytes_read.sum"])
        wr = to_bytes(val["dram__bytes_write.sum"], unit["dram__bytes_write.sum"])
        # the library the capture ran with: the box's copy of the repo is this tree (gpurun snapshots it), so the version string of
        # the in-tree library -- a hash of the kernel sources -- identifies it; bench.py refuses a capture of another build
        import ctypes
        import os
        lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "egomotion_with_local_loop_closures_b200", "libellc_gn.so")
        L = ctypes.CDLL(lib)
        L.ellc_version.restype = ctypes.c_char_p
        version = sys.argv[5] if len(sys.argv) > 5 else L.ellc_version().decode()
        print(json.dumps({"kernel": "gn_track_kernel", "library_version": version, "workload": "pair_sweep_640x480",
                          "pairs_per_launch": pairs, "dram_bytes_read": rd, "dram_bytes_write": wr,
                          "inst_executed": float(val["smsp__inst_executed.sum"]),
                          "issue_active_pct": float(val["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                          "kernel_ms_under_ncu": float(val["gpu__time_duration.sum"]), "l2_hit_rate_pct": float(val["lts__t_sector_hit_rate.pct"]),
                          "command": cmd, "report": rep.split("/")[-1]}, indent=1))


if __name__ == "__main__":
    main()
