"""GPU experiment: throughput of the epsilon-ball entry points (count, CSR fill, fused fit)."""
import math
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from point_cloud_toolbox_b200 import GridIndex  # noqa: E402
from scripts.tune import torus  # noqa: E402


def timeit(fn, reps=4):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), out


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    mean_count = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
    pts = torus(n)
    rho = n / (4 * math.pi ** 2 / 3)
    radius = math.sqrt(mean_count / (math.pi * rho))
    ix = GridIndex(pts, cell_hint=radius * 1.001)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.4:
        ix.ball_count(radius)
        torch.cuda.synchronize()
    tc, counts = timeit(lambda: ix.ball_count(radius))
    tf, fit = timeit(lambda: ix.curvature_ball(radius))
    tl, csr = timeit(lambda: ix.ball(radius), reps=2)
    print(f"N={n} radius={radius:.3e} mean count={counts.float().mean().item():.1f} max={int(counts.max())} "
          f"count={tc:.2f}ms ({n / tc / 1e3:.0f} Mq/s) fused={tf:.2f}ms ({n / tf / 1e3:.0f} Mq/s) csr={tl:.2f}ms", flush=True)
    ix.close()


if __name__ == "__main__":
    main()
