#!/usr/bin/env python
"""Per-opcode histogram of the loops of one kernel, from `cuobjdump -sass` (no GPU needed).

    python tools/sass_histogram.py [--lib path/to/libellc_gn.so] [--kernel gn_track_kernelILb0E] [--min-len 40] [--md out.md]

A loop is a backward branch (BRA to a lower address); its body is the address range [target, branch].  Innermost loops (no other
loop nested inside) are reported, longest first.  The pixel loop of a pyramid level of the forward kernel is unrolled twice
(ping-pong tap sets), so its body holds TWO pixels: the table prints instructions per pixel = body / pixels_per_body.

Pipe classes (B200_PROFILING.md / tools/ubench/pipes.cu): fma = FP32 FFMA/FMUL/FADD (1 issue cycle per warp instruction),
alu = integer / logic / compare / select / conversions-free moves, xu = MUFU and F2I / I2F conversions, lsu = memory,
cbu = branches / barriers / votes, uniform = uniform datapath.
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PIPE = [
    ("fma", r"^(FFMA|FMUL|FADD|FFMA2|FMUL2|FADD2|HFMA2|HADD2|HMUL2|IMAD|IMAD\.WIDE)"),
    ("fp64", r"^(DFMA|DMUL|DADD|DSETP)"),
    ("xu", r"^(MUFU|F2I|I2F|F2F|FRND|I2I|POPC|FLO|BREV)"),
    ("lsu", r"^(LDG|STG|LDS|STS|LDL|STL|LDGSTS|LDGDEPBAR|LD|ST|ATOM|RED|LDSM|LDC|LDCU|CCTL|MEMBAR|DEPBAR|ERRBAR|LDTM|UBLKCP)"),
    ("cbu", r"^(BRA|BRX|JMP|EXIT|BSSY|BSYNC|BAR|WARPSYNC|VOTE|VOTEU|CALL|RET|BREAK|NANOSLEEP|YIELD|BMOV|UCGABAR)"),
    ("shfl", r"^(SHFL|MATCH|REDUX)"),
    ("uniform", r"^(U[A-Z0-9]+|R2UR|S2UR)"),
    ("alu", r"^(LOP3|LOP|IADD3|IADD|LEA|SHF|SHL|SHR|PRMT|ISETP|FSETP|FSEL|SEL|MOV|FMNMX|IMNMX|VIMNMX|FCHK|IABS|PLOP3|P2R|R2P|CS2R|S2R|FSET|SGXT|BMSK|VABSDIFF|IDP|FMNMX3|VIADD|VIADDMNMX|FSWZADD|NOP)"),
]


def classify(op):
    base = op.split(".")[0]
    if base == "IMAD":
        # IMAD.MOV / IMAD.IADD / IMAD.SHL are integer work issued on the FMA pipe ("fmaheavy"); keep them visible
        return "imad"
    for name, pat in PIPE:
        if re.match(pat, base):
            return name
    return "other"


def disassemble(lib, kernel):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    hits = [f for f in funcs if kernel in f]
    if not hits:
        sys.exit(f"kernel matching {kernel!r} not found; have: {[f for f in funcs if 'ellc' in f][:20]}")
    return hits[0], funcs[hits[0]]


def opcode(text):
    t = text
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    return t.split(None, 1)[0]


def loops_of(ins):
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        op = opcode(t)
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    loops.append((addr_index[tgt], i))
    inner = [l for l in loops if not any((o[0] >= l[0] and o[1] <= l[1] and o != l) for o in loops)]
    return sorted(inner, key=lambda l: l[0] - l[1])


def rare_blocks(ins, lo, hi):
    """Instruction indices inside [lo, hi] that a forward conditional branch with an in-loop target jumps over: in the pixel loops
    these are the rarely executed blocks (the slow path of the exact division, the per-tap out-of-bounds gather)."""
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    skipped = set()
    for i in range(lo, hi + 1):
        a, t = ins[i]
        if t.startswith("@") and opcode(t).startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = addr_index.get(int(m.group(1), 16))
                if tgt is not None and i < tgt <= hi:
                    prev = ins[tgt - 1][1]
                    if not prev.startswith("@") and opcode(prev) == "BRA":
                        continue                                    # layout 2 below: this branch jumps TO the rare block
                    skipped.update(range(i + 1, tgt))
    # the other layout of the same thing: `@!P BRA slow; <usual path>; BRA join; slow: ...; join:` -- the block between an
    # unconditional forward BRA and its target is rare when a conditional branch of the loop jumps to its first instruction
    cond_targets = set()
    for i in range(lo, hi + 1):
        a, t = ins[i]
        if t.startswith("@") and opcode(t).startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) in addr_index:
                cond_targets.add(addr_index[int(m.group(1), 16)])
    for i in range(lo, hi + 1):
        a, t = ins[i]
        if not t.startswith("@") and opcode(t) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", t)
            tgt = addr_index.get(int(m.group(1), 16)) if m else None
            if tgt is not None and i + 1 < tgt <= hi and (i + 1) in cond_targets:
                skipped.update(range(i + 1, tgt))
    return skipped


def histogram(ins, lo, hi, skip=()):
    ops = collections.Counter()
    pipes = collections.Counter()
    for i, (a, t) in enumerate(ins[lo:hi + 1], start=lo):
        if i in skip:
            continue
        op = opcode(t)
        key = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "F2I", "I2F", "LDG", "LDS", "IMAD", "LDGSTS")) else op.split(".")[0]
        ops[key] += 1
        pipes[classify(op)] += 1
    return ops, pipes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "egomotion_with_local_loop_closures_b200", "libellc_gn.so"))
    ap.add_argument("--kernel", default="gn_track_kernelILb0E")
    ap.add_argument("--min-len", type=int, default=60)
    ap.add_argument("--pixels-per-body", type=int, default=2)
    ap.add_argument("--md", default=None)
    ap.add_argument("--json", default=None, help="machine-readable summary (bench.py reads profiles/r02_sass_histogram.json)")
    ap.add_argument("--require", default="LDGSTS", help="only loops containing this opcode (the pixel loops stream records with LDGSTS); '' = all")
    ap.add_argument("--fast-path", type=int, default=1, help="1: leave out blocks that an in-loop forward conditional branch jumps over")
    args = ap.parse_args()
    name, ins = disassemble(args.lib, args.kernel)
    lines = [f"# SASS opcode histogram: `{name}`", "",
             f"`cuobjdump -sass {os.path.relpath(args.lib, ROOT)}`; {len(ins)} instructions in the kernel; innermost loops with at least "
             f"{args.min_len} instructions, {args.pixels_per_body} pixel(s) per loop body.", ""]
    for lo, hi in loops_of(ins):
        n = hi - lo + 1
        if n < args.min_len:
            continue
        if args.require and not any(opcode(t).startswith(args.require) for _, t in ins[lo:hi + 1]):
            continue
        skip = rare_blocks(ins, lo, hi) if args.fast_path else set()
        ops, pipes = histogram(ins, lo, hi, skip)
        ppb = args.pixels_per_body
        n_exec = n - len(skip)
        lines.append(f"## loop 0x{ins[lo][0]:x} .. 0x{ins[hi][0]:x}: {n} instructions in the body, {n_exec} on the usual path "
                     f"({len(skip)} in rarely executed blocks that a forward branch skips) = {n_exec / ppb:.1f} per pixel")
        lines.append("")
        lines.append("| pipe class | per body | per pixel |")
        lines.append("|---|---|---|")
        for k, v in sorted(pipes.items(), key=lambda kv: -kv[1]):
            lines.append(f"| {k} | {v} | {v / ppb:.1f} |")
        lines.append("")
        lines.append("| opcode | per body | per pixel |")
        lines.append("|---|---|---|")
        for k, v in sorted(ops.items(), key=lambda kv: -kv[1]):
            lines.append(f"| {k} | {v} | {v / ppb:.1f} |")
        lines.append("")
    if args.json:
        import ctypes
        import json
        L = ctypes.CDLL(args.lib)
        L.ellc_version.restype = ctypes.c_char_p
        per_loop = []
        for lo, hi in loops_of(ins):
            n = hi - lo + 1
            if n < args.min_len or (args.require and not any(opcode(t).startswith(args.require) for _, t in ins[lo:hi + 1])):
                continue
            skip = rare_blocks(ins, lo, hi) if args.fast_path else set()
            ops, pipes = histogram(ins, lo, hi, skip)
            per_loop.append({"start": "0x%x" % ins[lo][0], "body": n, "usual_path": n - len(skip), "per_pixel": (n - len(skip)) / args.pixels_per_body,
                             "writes_weight_image": bool(ops.get("ST", 0)), "pipes_per_pixel": {k: v / args.pixels_per_body for k, v in pipes.items()}})
        plain = [l for l in per_loop if not l["writes_weight_image"]] or per_loop
        out = {"library_version": L.ellc_version().decode(), "kernel": name, "loops": per_loop,
               "inst_per_pixel_level0": min(l["per_pixel"] for l in plain) if plain else None,
               "note": "pixel loops of the four pyramid levels (x with / without display_weightimg output); the forward loop body holds two pixels"}
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)
    text = "\n".join(lines)
    if args.md:
        with open(args.md, "w") as f:
            f.write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
