"""Instruction / stall-sample share per pipeline phase (function) of a kernel, from an ncu source-level capture.
Phases are found by scanning the source files for function headers, so line numbers never go stale.

    python scripts/ncu_phases.py gpurun_out/prof.ncu-rep
"""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "point_cloud_toolbox_b200", "csrc")
HEAD = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:PCT_HD(?:_NOINLINE)?|__device__ __forceinline__|__global__|static|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
STRUCT = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?struct\s+([A-Za-z_][A-Za-z0-9_]*)")


def outline(path):
    """line -> enclosing 'struct::function' label (coarse: last header seen above the line)."""
    labels, cur_s, cur_f, depth_s = {}, None, None, None
    depth = 0
    for n, line in enumerate(open(path), 1):
        m = STRUCT.match(line)
        if m and "{" in line and ";" not in line.split("{")[0]:
            cur_s, depth_s = m.group(1), depth
        m = HEAD.match(line)
        if m and m.group(1) not in ("if", "for", "while", "return", "sizeof"):
            cur_f = m.group(1)
        labels[n] = (cur_s + "::" if cur_s else "") + (cur_f or "?")
        depth += line.count("{") - line.count("}")
        if cur_s is not None and depth <= depth_s and "}" in line:
            cur_s = None
    return labels


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    outlines, cur, hdr, agg = {}, None, None, {}
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "File Path":
            cur = os.path.basename(r[1])
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10 or r[2] != "-" or cur is None:
            continue
        try:
            line, samples, inst = int(r[0]), int(r[6]), int(r[7])
        except ValueError:
            continue
        if cur not in outlines:
            p = os.path.join(CSRC, cur)
            outlines[cur] = outline(p) if os.path.exists(p) else {}
        label = cur + "  " + outlines[cur].get(line, "?")
        a = agg.setdefault(label, [0, 0])
        a[0] += inst
        a[1] += samples
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print(f"# warp instructions {ti}, stall samples {ts}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if v[0] / ti > 0.002:
            print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / ts:5.1f}% smp  {k}")


if __name__ == "__main__":
    main()
