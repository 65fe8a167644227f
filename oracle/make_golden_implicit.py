"""Golden vectors for the implicit 10-coefficient quadric, made by running the UNMODIFIED reference (dev container only).

TEST INFRASTRUCTURE ONLY.      python -m oracle.make_golden_implicit

* ``calculate_implicit_quadric_curvatures`` (ref :435-480) is a closed formula: its outputs on random coefficient
  vectors pin the product's ``pct_implicit_quadric_curvature``.
* ``fit_implicit_quadric_surface`` (ref :363-396) hands min |A c|^2, |c| = 1 to scipy's SLSQP from the all-ones start.
  Stored for a few bunny neighbourhoods (the point and its k - 1 nearest, centred, exactly as ref :617-633 builds them):
  the reference's coefficients, its objective value, and the spectrum of A^T A -- the evidence that SLSQP stops far
  from the minimiser (objective many orders above the smallest eigenvalue, small overlap with its eigenvector), which
  is why the product's fit is declared "parity unpinned" (DESIGN.md section 9).
"""
from __future__ import annotations

import os

import numpy as np

from .make_golden import GOLDEN_DIR, import_reference


def design(points):
    p = np.asarray(points)
    return np.column_stack((p[:, 0] ** 2, p[:, 1] ** 2, p[:, 2] ** 2, p[:, 0] * p[:, 1], p[:, 0] * p[:, 2], p[:, 1] * p[:, 2],
                            p[:, 0], p[:, 1], p[:, 2], np.ones(len(p))))


def main():
    ref = import_reference()
    rng = np.random.default_rng(21)
    coeffs = rng.normal(size=(256, 10))
    coeffs[:8, 6:9] *= 1e-3                                   # small gradients
    curv = np.array([ref.PointCloud.calculate_implicit_quadric_curvatures(c) for c in coeffs])
    bunny = np.load(os.path.join(GOLDEN_DIR, "bunny_points.npz"))["points"]
    from scipy.spatial import cKDTree

    tree = cKDTree(bunny)
    k = 30
    rows = [100, 5000, 12345, 20000, 30000]
    nbhd, ref_c, ref_obj, eig_w, eig_v = [], [], [], [], []
    for i in rows:
        _, nb = tree.query(bunny[i], k)                        # ref :624 (k, the point itself included)
        pts = bunny[nb] - bunny[i]                             # ref :627
        c = ref.PointCloud.fit_implicit_quadric_surface(pts)   # SLSQP, unmodified
        A = design(pts)
        w, v = np.linalg.eigh(A.T @ A)
        nbhd.append(pts); ref_c.append(c); ref_obj.append(np.sum((A @ c) ** 2)); eig_w.append(w); eig_v.append(v[:, 0])
        print(f"row {i}: objective(SLSQP) = {ref_obj[-1]:.3e}  smallest eigenvalue = {w[0]:.3e}  |c| = {np.linalg.norm(c):.6f}  "
              f"overlap with the minimiser = {abs(np.dot(c / np.linalg.norm(c), v[:, 0])):.3f}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "implicit.npz"), coeffs=coeffs, curv=curv, k=np.int32(k), rows=np.asarray(rows),
                        neighbourhoods=np.asarray(nbhd, np.float32), slsqp_coeffs=np.asarray(ref_c), slsqp_objective=np.asarray(ref_obj),
                        eigenvalues=np.asarray(eig_w), minimiser=np.asarray(eig_v))


if __name__ == "__main__":
    main()
