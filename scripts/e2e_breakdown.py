"""Wall-clock breakdown of the public-API path at bench size (run under gpurun)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from point_cloud_toolbox_b200 import PointCloud, engine  # noqa: E402
from scripts.tune import torus  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    k = 20
    dev_pts = torus(n)
    host = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
    host.copy_(dev_pts)
    torch.cuda.synchronize()
    del dev_pts
    pts = host.numpy()
    empty = np.zeros((n, 0), np.float32)
    for rep in range(4):
        t = [time.perf_counter()]

        def lap():
            torch.cuda.synchronize()
            t.append(time.perf_counter())

        pc = PointCloud(points=pts, normals=empty, k_neighbors=k); lap()
        d = pc._upload(); lap()
        ix = engine.GridIndex(d, k_hint=k); lap()
        fit = ix.curvature_knn(k, want_coeffs=False); lap()
        bad = bool((fit.status & 8).any()); lap()
        kh = engine.to_host(fit.curv[:, :2].t()); lap()
        ix.close(); lap()
        names = ["init", "h2d", "build", "fused", "status", "d2h", "close"]
        print(rep, " ".join(f"{nm}={1e3 * (b - a):.1f}ms" for nm, a, b in zip(names, t, t[1:])), f"total={1e3 * (t[-1] - t[0]):.1f}ms", flush=True)
        t0 = time.perf_counter()
        pc2 = PointCloud(points=pts, normals=empty, k_neighbors=k)
        pc2.plant_kdtree(k)
        K, H = pc2.compute_pointwise_explicit_quadratic_curvature()
        torch.cuda.synchronize()
        print(rep, f"api total={1e3 * (time.perf_counter() - t0):.1f}ms", flush=True)
        del pc2, K, H


if __name__ == "__main__":
    main()
