"""Golden vectors for degenerate geometry, made by running the UNMODIFIED reference (dev container only).

TEST INFRASTRUCTURE ONLY.      python -m oracle.make_golden_degenerate

np.linalg.lstsq (ref :359) returns the minimum-norm solution when the design matrix of a neighbourhood is rank
deficient -- collinear scan lines, coincident points, points that lie on a conic -- so the reference produces
finite numbers there.  Cases (each a small cloud, k chosen so that whole neighbourhoods are degenerate):

  line        points on one straight line (+ a second, far away line)            rank 3 designs, z = 0
  coincident  every point of a random cloud repeated 24 times (k = 20)            all neighbours equal the query: X = [0 0 0 0 0 1]
  plane_grid  a regular lattice in the plane z = 0.25, cut to 3 rows in y          rank < 6 near the ends, z = 0
  two_lines   two parallel lines with a parabolic profile z = 0.01 x^2, y in {0, 1} b^2 = b on every row: rank 5, z != 0

Stored: the fp32 cloud, k, and the reference's own neighbor_indices, quadratic_coefficients, K, H for every point.
"""
from __future__ import annotations

import os

import numpy as np

from .make_golden import GOLDEN_DIR, import_reference


def clouds():
    rng = np.random.default_rng(12)
    t = np.arange(60, dtype=np.float32) * np.float32(0.125)
    line = np.concatenate((np.stack((t, 0 * t, 0 * t), 1), np.stack((t, 0 * t + 50, 0 * t + 7), 1))).astype(np.float32)
    base = rng.normal(size=(12, 3)).astype(np.float32)
    coincident = np.repeat(base, 24, axis=0)
    gx, gy = np.meshgrid(np.arange(40, dtype=np.float32), np.arange(3, dtype=np.float32), indexing="ij")
    plane = np.stack((gx.ravel(), gy.ravel(), np.full(gx.size, 0.25, np.float32)), 1).astype(np.float32)
    x = np.arange(-20, 21, dtype=np.float32)
    two = np.concatenate((np.stack((x, 0 * x, np.float32(0.01) * x * x), 1), np.stack((x, 0 * x + 1, np.float32(0.01) * x * x), 1))).astype(np.float32)
    return {"line": (line, 20), "coincident": (coincident, 20), "plane_grid": (plane, 20), "two_lines": (two, 12)}


def main():
    ref = import_reference()
    out = {}
    for name, (pts, k) in clouds().items():
        pc = ref.PointCloud(points=pts, normals=np.zeros((len(pts), 0), np.float32), k_neighbors=k)
        pc.plant_kdtree(k)
        with np.errstate(all="ignore"):
            K, H = pc.compute_pointwise_explicit_quadratic_curvature()
        coeffs = np.stack([np.asarray(c, np.float32) for c in pc.quadratic_coefficients])
        out[name + "_points"] = pts
        out[name + "_k"] = np.int32(k)
        out[name + "_neighbor_indices"] = pc.neighbor_indices
        out[name + "_coeffs"] = coeffs
        out[name + "_K"] = np.asarray(K, np.float32)
        out[name + "_H"] = np.asarray(H, np.float32)
        print(f"{name}: N={len(pts)} k={k} finite K: {np.isfinite(K).mean():.3f}  |K|max={np.nanmax(np.abs(K)):.3g} |H|max={np.nanmax(np.abs(H)):.3g} "
              f"|coeffs|max={np.nanmax(np.abs(coeffs)):.3g}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "degenerate.npz"), **out)


if __name__ == "__main__":
    main()
