"""GPU experiments: cell-size sweep and stage timing of the device path (run under gpurun)."""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from point_cloud_toolbox_b200 import GridIndex  # noqa: E402


def torus(n, seed=3, dev="cuda"):
    gen = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * (2 * math.pi)
    v = torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * (2 * math.pi)
    w = 1.0 + torch.cos(v) / 3.0
    return torch.stack((w * torch.cos(u), w * torch.sin(u), torch.sin(v) / 3.0), 1).float().contiguous()


def timed(fn, reps=3, warm_s=0.4):
    # the idle GPU sits at 120 MHz: spin until the clocks are up before timing anything
    t0 = time.perf_counter()
    while True:
        fn()
        torch.cuda.synchronize()
        if time.perf_counter() - t0 > warm_s:
            break
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    ks = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [20]
    pts = torus(n)
    for k in ks:
        for hint in sorted({max(4, int(k * f)) for f in (0.5, 0.65, 0.8, 1.0, 1.25, 1.5)}):
            t_build, ix = timed(lambda: GridIndex(pts, k_hint=hint), reps=2)
            t_q, _ = timed(lambda: ix.curvature_knn(k, want_coeffs=False))
            t_l, _ = timed(lambda: ix.knn(k), reps=1)
            st = ix.last_stats()
            info = ix.info()
            ix.curvature_knn(k, want_coeffs=False)
            st = ix.last_stats()
            print(f"N={n} k={k} k_hint={hint} ppc={n / info.cells_level0:.2f} build={t_build:.2f}ms fused={t_q:.2f}ms "
                  f"({n / t_q / 1e3:.1f} Mq/s) lists={t_l:.2f}ms retries={st.level1_retries} exact={st.exact_path} unstaged={st.unstaged}", flush=True)
            ix.close()


if __name__ == "__main__":
    main()
