"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the dev container (``/root/reference`` is not
present on the GPU box):

    python -m oracle.make_golden

The reference file is imported byte-for-byte; only its top-level imports that
are irrelevant to the hot path and not installed here are satisfied with empty
modules (matplotlib, pymesh, pyvista, memory_profiler -- pointCloudToolbox.py:7,
:11, :16-17, :22).  For each case the reference's own public calls are made:

    pc = PointCloud(file_path)  /  PointCloud(points=..., normals=...)
    pc.plant_kdtree(k)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()

and a seeded sample of rows of its outputs is stored together with the fp32
cloud it saw (``pc.points``, i.e. after the loader's max-shift).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ROWS_PER_CASE = 2048


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "pymesh", "pyvista", "memory_profiler"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.patches"].Patch = object
    sys.modules["memory_profiler"].profile = lambda f: f
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pointCloudToolbox  # noqa: E402  (the reference, unmodified)
    return pointCloudToolbox


def run_case(ref, name, k, file_path=None, points=None, seed=0):
    if file_path is not None:
        pc = ref.PointCloud(file_path, k_neighbors=k)
    else:
        pc = ref.PointCloud(points=points, normals=np.zeros((len(points), 0), np.float32), k_neighbors=k)
    pc.plant_kdtree(k)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    n = len(pc.points)
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.choice(n, size=min(ROWS_PER_CASE, n), replace=False))
    coeffs = np.stack([pc.quadratic_coefficients[i] for i in rows]).astype(np.float32)
    # a few raw outputs of the static methods, for direct function-level checks
    stat_rows = rows[:64]
    rotated = np.stack([
        ref.PointCloud.get_best_fit_plane_and_rotate(pc.points[pc.neighbor_indices[i]] - pc.points[i])
        for i in stat_rows
    ])
    out = {
        "k": np.int32(k),
        "rows": rows.astype(np.int64),
        "neighbor_indices": pc.neighbor_indices[rows],
        "dists": pc.dists[rows],
        "quadratic_coefficients": coeffs,
        "K_quadratic": np.asarray(K, np.float32)[rows],
        "H_quadratic": np.asarray(H, np.float32)[rows],
        "K_H_sq_quadratic": np.asarray(pc.K_H_sq_quadratic, np.float32)[rows],
        "static_rows": stat_rows.astype(np.int64),
        "static_rotated": rotated,
        "num_points": np.int64(n),
        "nan_count": np.int64(np.count_nonzero(~np.isfinite(np.asarray(K)))),
    }
    path = os.path.join(GOLDEN_DIR, f"{name}_k{k}.npz")
    np.savez_compressed(path, **out)
    print(f"{name} k={k}: N={n} rows={len(rows)} NaN(K)={out['nan_count']} -> {path}")
    return pc


def main():
    from oracle import datasets

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = import_reference()

    bunny = os.path.join(REFERENCE_ROOT, "sample_scans", "bunny.txt")
    egg = os.path.join(REFERENCE_ROOT, "sample_scans", "egg_carton.txt")

    pc = run_case(ref, "bunny", 20, file_path=bunny, seed=20)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "bunny_points.npz"), points=pc.points)
    run_case(ref, "bunny", 30, file_path=bunny, seed=30)

    pc = run_case(ref, "egg_carton", 20, file_path=egg, seed=21)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "egg_carton_points.npz"), points=pc.points)

    # C1 stand-in (sample_scans/torus.txt is absent): the loader is exercised
    # through a real text file, the cloud it produced is stored.
    p64, _, _ = datasets.torus_grid(317)
    tmp = "/tmp/torus_c1.txt"
    np.savetxt(tmp, p64, fmt="%.6f")
    pc = run_case(ref, "torus_c1", 20, file_path=tmp, seed=22)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "torus_c1_points.npz"), points=pc.points)
    assert np.array_equal(pc.points, datasets.torus_c1()[0]), "datasets.torus_c1() must equal the loader's output"

    # loader fixture: a tiny 6-column file (points + normals) and what the reference made of it
    rng = np.random.default_rng(5)
    table = np.round(rng.normal(size=(50, 6)), 5)
    small = "/tmp/loader_case.txt"
    np.savetxt(small, table, fmt="%.5f")
    pc = ref.PointCloud(small, k_neighbors=5)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, "loader_case.npz"),
        table=table, points=pc.points, normals=pc.normals,
        x_domain=np.asarray(pc.x_domain), y_domain=np.asarray(pc.y_domain), z_domain=np.asarray(pc.z_domain),
        l1_norm=pc.l1_norm, l2_norm=pc.l2_norm, infinity_norm=pc.infinity_norm,
    )
    neighbor_study_case(ref, bunny)
    print("done")


def neighbor_study_case(ref, bunny_path):
    """Return values of the reference's own explicit_quadratic_neighbor_study (ref :732-800) for seeded samples."""
    pc = ref.PointCloud(bunny_path, k_neighbors=20)
    pc.plant_kdtree(20)
    seeds, tols, sample_size = [0, 1], [1e-7, 5.0, 50.0, 500.0, 5000.0], 48
    results = np.zeros((len(seeds), len(tols)), np.int64)
    samples = np.zeros((len(seeds), sample_size), np.int64)
    for a, seed in enumerate(seeds):
        for b, tol in enumerate(tols):
            np.random.seed(seed)
            results[a, b] = pc.explicit_quadratic_neighbor_study(tol=tol, sample_size=sample_size)
        np.random.seed(seed)
        samples[a] = np.random.randint(0, len(pc.points), sample_size)  # the sample ref :751 drew
    np.savez_compressed(os.path.join(GOLDEN_DIR, "neighbor_study.npz"), seeds=np.asarray(seeds), tols=np.asarray(tols),
                        sample_size=np.int64(sample_size), results=results, samples=samples)
    print("neighbor study:", results.tolist())


if __name__ == "__main__":
    main()
