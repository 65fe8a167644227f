"""GPU experiment: device and host time of repeated whole-cloud index builds (looking for spikes)."""
import os
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from point_cloud_toolbox_b200 import GridIndex  # noqa: E402
from scripts.tune import torus  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    with_query = len(sys.argv) > 2 and sys.argv[2] == "q"
    pts = torus(n)
    print("cpus", os.cpu_count(), "load", os.getloadavg(), flush=True)
    nosync = len(sys.argv) > 3 and sys.argv[3] == "nosync"
    evs = []
    last = None
    for rep in range(14):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record()
        ix = GridIndex(pts, k_hint=20)
        e1.record()
        t1 = time.perf_counter()
        fit = ix.curvature_knn(20, want_coeffs=False) if with_query else None
        e2.record()
        if last is not None:
            last[0].close()
        last = (ix, fit)
        if nosync:
            evs.append((e0, e1, e2, t0, t1))
            continue
        torch.cuda.synchronize()
        print(f"rep{rep} build dev={e0.elapsed_time(e1):.2f}ms host={1e3 * (t1 - t0):.2f}ms query dev={e1.elapsed_time(e2):.2f}ms", flush=True)
    torch.cuda.synchronize()
    for rep, (e0, e1, e2, t0, t1) in enumerate(evs):
        print(f"rep{rep} nosync build dev={e0.elapsed_time(e1):.2f}ms host={1e3 * (t1 - t0):.2f}ms query dev={e1.elapsed_time(e2):.2f}ms", flush=True)


if __name__ == "__main__":
    main()
