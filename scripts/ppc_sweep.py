"""GPU experiment: fine sweep of the points-per-cell target (through k_hint) with per-repetition timing."""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from point_cloud_toolbox_b200 import GridIndex  # noqa: E402
from scripts.tune import torus  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    hints = [int(a) for a in sys.argv[3].split(",")]
    pts = torus(n)
    ix = GridIndex(pts, k_hint=k)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:
        ix.curvature_knn(k, want_coeffs=False)
        torch.cuda.synchronize()
    ix.close()
    for hint in hints:
        ix = GridIndex(pts, k_hint=hint)
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix.curvature_knn(k, want_coeffs=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        st = ix.last_stats()
        info = ix.info()
        print(f"N={n} k={k} hint={hint} ppc={n / info.cells_level0:.2f} query={min(ts):.2f}ms retries={st.level1_retries} unstaged={st.unstaged}", flush=True)
        ix.close()


if __name__ == "__main__":
    main()
