"""Per-source-line share of executed instructions and stall samples from an ncu --set full --import-source capture.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "File Path":
            cur = r[1]
            continue
        if r and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10 or r[2] != "-":
            continue
        try:
            line, samples, inst = int(r[0]), int(r[6]), int(r[7])
            thr = int(r[8])
        except ValueError:
            continue
        agg.append((inst, samples, thr, cur.split("/")[-1], line, r[1][:110]))
    ti, ts = sum(a[0] for a in agg), sum(a[1] for a in agg)
    print(f"# warp instructions {ti}, stall samples {ts}")
    files = {}
    for a in agg:
        f = files.setdefault(a[3], [0, 0])
        f[0] += a[0]
        f[1] += a[1]
    for f, v in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"# {100 * v[0] / ti:5.1f}% inst {100 * v[1] / ts:5.1f}% smp  {f}")
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{100 * a[0] / ti:5.1f}% inst {100 * a[1] / ts:5.1f}% smp  thr/inst {a[2] / max(a[0], 1):5.1f}  {a[3]}:{a[4]}  {a[5]}")


if __name__ == "__main__":
    main()
