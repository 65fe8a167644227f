"""GPU micro-benchmark of the fused query call: per-repetition device times (CUDA events), one index.

    python scripts/qbench.py [N] [k,k,...] [reps] [k_hint,k_hint,...]
"""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from point_cloud_toolbox_b200 import GridIndex  # noqa: E402
from scripts.tune import torus  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    ks = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [20]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    hints = [int(a) for a in sys.argv[4].split(",")] if len(sys.argv) > 4 else [None]
    pts = torus(n)
    for k, hint in [(k, h) for k in ks for h in hints]:
        ix = GridIndex(pts, k_hint=hint or k)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.5:  # clocks up
            ix.curvature_knn(k, want_coeffs=False)
            torch.cuda.synchronize()
        times = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix.curvature_knn(k, want_coeffs=False)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        st = ix.last_stats()
        ts = sorted(times)
        print(f"N={n} k={k} hint={hint} ppc={n / ix.info().cells_level0:.2f} min={ts[0]:.2f}ms med={ts[len(ts) // 2]:.2f}ms max={ts[-1]:.2f}ms ({n / ts[len(ts) // 2] / 1e3:.1f} Mq/s) "
              f"retries={st.level1_retries} exact={st.exact_path} unstaged={st.unstaged} all={[round(t, 1) for t in times]}", flush=True)
        ix.close()


if __name__ == "__main__":
    main()
