"""GPU experiment: sweep the pre-collection gain (PCT_CUT_GAIN) of the staged kNN kernel."""
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from point_cloud_toolbox_b200 import GridIndex  # noqa: E402
from scripts.tune import timed, torus  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    gains = [float(g) for g in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 3, 4, 5, 6.5, 8, 10]
    pts = torus(n)
    for g in gains:
        os.environ["PCT_CUT_GAIN"] = str(g)
        ix = GridIndex(pts, k_hint=k)
        t_q, _ = timed(lambda: ix.curvature_knn(k, want_coeffs=False))
        st = ix.last_stats()
        print(f"N={n} k={k} gain={g} fused={t_q:.2f}ms ({n / t_q / 1e3:.1f} Mq/s) retries={st.level1_retries} "
              f"exact={st.exact_path} unstaged={st.unstaged}", flush=True)
        ix.close()


if __name__ == "__main__":
    main()
