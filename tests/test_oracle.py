"""The oracle against the reference's own outputs (tests/golden, made by oracle/make_golden.py)."""
import io

import numpy as np
import pytest

import oracle
from oracle import datasets
from conftest import load_golden

CASES = [("bunny", 20), ("bunny", 30), ("egg_carton", 20), ("torus_c1", 20)]


def _cloud(name):
    return load_golden(name + "_points")["points"]


@pytest.mark.parametrize("name,k", CASES)
def test_ranking_key_is_scipys(name, k):
    """dists stored by the reference == sqrt of the restated fp64 key, bit for bit."""
    g = load_golden(f"{name}_k{k}")
    pts = _cloud(name).astype(np.float64)
    d2 = oracle.squared_distance_key(pts[g["neighbor_indices"]], pts[g["rows"]][:, None, :])
    assert np.array_equal(np.sqrt(d2).astype(np.float32), g["dists"])


@pytest.mark.parametrize("name,k", CASES)
def test_knn_canonical_matches_reference_up_to_ties(name, k):
    g = load_golden(f"{name}_k{k}")
    pts = _cloud(name)
    idx, dist, d2 = oracle.knn_canonical(pts, k, rows=g["rows"])
    # distances are tie-independent: must be bit-equal row by row
    assert np.array_equal(dist, g["dists"])
    ref_idx = g["neighbor_indices"]
    differs = (idx != ref_idx).any(axis=1)
    # every difference must sit inside a group of exactly equal keys
    p64 = pts.astype(np.float64)
    d2_ref = oracle.squared_distance_key(p64[ref_idx], p64[g["rows"]][:, None, :])
    assert np.array_equal(d2_ref, d2)
    for r in np.nonzero(differs)[0]:
        bad = idx[r] != ref_idx[r]
        for c in np.nonzero(bad)[0]:
            same_key = d2[r] == d2[r, c]
            boundary = c == k - 1 or same_key.sum() > 1 or d2[r, c] == d2[r, -1]
            assert boundary, (name, k, r, c)
    # rows without any exact tie are identical, order included
    no_tie = np.array([len(np.unique(row)) == k for row in d2]) & (d2[:, -1] < np.inf)
    strict = no_tie & ~differs
    assert strict.sum() >= 0.9 * no_tie.sum()
    if name == "bunny":
        assert not differs.any()  # no boundary ties on bunny


@pytest.mark.parametrize("name,k", CASES)
def test_per_point_functions_reproduce_reference_bitwise(name, k):
    """Fed the reference's own neighbour rows, the restatement gives the reference's numbers exactly."""
    g = load_golden(f"{name}_k{k}")
    pts = _cloud(name)
    sel = np.arange(0, len(g["rows"]), 4)
    res = oracle.curvature_from_neighbors(pts, g["neighbor_indices"][sel], g["rows"][sel])
    assert np.array_equal(res["coeffs"], g["quadratic_coefficients"][sel])
    assert np.array_equal(res["K"], g["K_quadratic"][sel])
    assert np.array_equal(res["H"], g["H_quadratic"][sel])
    assert np.array_equal(res["H2"], g["K_H_sq_quadratic"][sel])


@pytest.mark.parametrize("name,k", [("bunny", 20), ("egg_carton", 20)])
def test_static_rotation_reproduces_reference_bitwise(name, k):
    g = load_golden(f"{name}_k{k}")
    pts = _cloud(name)
    pos = {int(r): j for j, r in enumerate(g["rows"])}
    for j, i in enumerate(g["static_rows"]):
        nb = g["neighbor_indices"][pos[int(i)]]
        rot, normal = oracle.best_fit_plane_and_rotate(pts[nb] - pts[i], return_normal=True)
        assert np.array_equal(rot, g["static_rotated"][j])
        # the rotation takes the oriented normal to +z (or is the identity)
        assert abs(np.linalg.norm(normal) - 1) < 1e-12


@pytest.mark.parametrize("name,k", [("bunny", 30), ("torus_c1", 20)])
def test_batched_equals_per_point(name, k):
    g = load_golden(f"{name}_k{k}")
    pts = _cloud(name)
    sel = np.arange(0, len(g["rows"]), 2)
    a = oracle.curvature_from_neighbors(pts, g["neighbor_indices"][sel], g["rows"][sel])
    b = oracle.curvature_from_neighbors_batched(pts, g["neighbor_indices"][sel], g["rows"][sel])
    r_k = g["dists"][sel, -1].astype(np.float64)
    rep = oracle.compare.curvature_report(b, a, r_k)
    assert rep["violations"] == 0, rep
    assert rep["tight_fraction"] > 0.99, rep
    assert np.allclose(a["margin"], b["margin"], atol=1e-9)


def test_loader_matches_reference():
    g = load_golden("loader_case")
    buf = io.StringIO()
    np.savetxt(buf, g["table"], fmt="%.5f")
    buf.seek(0)
    pts, nrm = oracle.load_points(buf)
    assert pts.dtype == np.float32 and np.array_equal(pts, g["points"])
    assert np.array_equal(nrm, g["normals"])
    assert pts[:, 0].max() == 0 and pts[:, 1].max() == 0


def test_torus_c1_is_regenerable(torus_c1):
    pts, K, H = datasets.torus_c1()
    assert np.array_equal(pts, torus_c1)
    assert len(pts) == 317 * 317


def test_ball_membership_is_inclusive_squared_radius(bunny):
    pts = bunny[:6000]
    r = 3.8e-3
    rows = np.arange(0, 6000, 7)
    off, idx, dist = oracle.ball_canonical(pts, r, rows=rows)
    p64 = pts.astype(np.float64)
    for j, i in enumerate(rows[:200]):
        d2 = oracle.squared_distance_key(p64, p64[i])
        want = np.nonzero((d2 <= r * r) & (np.arange(len(pts)) != i))[0]
        got = idx[off[j]:off[j + 1]]
        assert set(got.tolist()) == set(want.tolist())
        key = list(zip(d2[got].tolist(), got.tolist()))
        assert key == sorted(key)
    # a radius that hits a neighbour distance exactly: boundary is inclusive
    i = 17
    d2 = oracle.squared_distance_key(p64, p64[i])
    target = np.sort(d2)[10]
    r_exact = float(np.sqrt(target))
    if r_exact * r_exact == target:
        off, idx, _ = oracle.ball_canonical(pts, r_exact, rows=[i])
        assert np.count_nonzero(d2 <= target) - 1 == off[1]


def test_closed_form_sphere():
    pts, K, H = datasets.sphere_fibonacci(20000)
    rows = np.arange(0, 20000, 40)
    res = oracle.knn_curvature(pts, 20, rows=rows)
    rep = oracle.compare.closed_form_report(res["K"], res["H"], K[rows], H[rows])
    assert rep["K_abs_p99"] < 2e-2 and rep["H_abs_p99"] < 1e-2, rep


def test_knn_errors():
    pts = np.random.default_rng(0).normal(size=(10, 3)).astype(np.float32)
    with pytest.raises(IndexError):
        oracle.knn_canonical(pts, 10)
    with pytest.raises(ValueError):
        oracle.best_fit_plane_and_rotate(np.array([[0, 0, np.nan], [1, 0, 0], [0, 1, 0]], np.float32))


def test_neighbor_study_reproduces_reference_returns(bunny):
    """oracle.neighbor_study on the samples the reference drew gives the reference's own return values."""
    g = load_golden("neighbor_study")
    from scipy.spatial import cKDTree

    tree = cKDTree(bunny)
    for a in range(len(g["seeds"])):
        for b, tol in enumerate(g["tols"]):
            if tol in (1e-7, 500.0):  # the two ends: nothing converges / most converge early (keeps the CPU suite short)
                assert oracle.neighbor_study(bunny, g["samples"][a], tol=float(tol), tree=tree) == int(g["results"][a, b])


# ---------------------------------------------------------------------------
# either side of the path: the restatements of oracle/around_path.py reproduce what the
# unmodified reference functions produced (tests/golden/io_energy_pca.npz, oracle/make_golden_io.py)
# ---------------------------------------------------------------------------
def test_around_path_oracle_is_pinned_to_the_reference():
    from oracle import around_path as ap

    g = load_golden("io_energy_pca")
    for tag in ("f64", "f32"):
        assert ap.points_ply_bytes(g[f"points_ply_in_{tag}"]) == g[f"points_ply_bytes_{tag}"].tobytes()
    assert ap.curvature_ply_bytes(g["curv_ply_points"], g["curv_ply_K"], g["curv_ply_H"]) == g["curv_ply_bytes"].tobytes()
    p = ap.parse_ply(g["parse_ply_text"].tobytes().decode())
    assert p.dtype == np.float32 and np.array_equal(p.view(np.uint32), g["parse_ply_points"].view(np.uint32))
    e = ap.mesh_energies(g["energy_vertices"], g["energy_triangles"], g["energy_K"], g["energy_H"])
    assert np.allclose(e, g["energy_result"], rtol=1e-14, atol=1e-15)
    e0 = ap.mesh_energies(g["energy_vertices"], g["energy_triangles"])
    assert np.allclose(e0, g["energy_result_no_curvature"], rtol=1e-14, atol=0)
    l1, l2, K, H, d = ap.pca_principal_curvatures(g["pca_points"], int(g["pca_k"]))
    assert np.array_equal(l1, g["pca_l1"]) and np.array_equal(l2, g["pca_l2"])
    assert np.array_equal(K, g["pca_K"]) and np.array_equal(H, g["pca_H"]) and np.array_equal(d, g["pca_directions"])
    # the vectorised form on canonical kNN rows gives the same numbers (what the GPU tests compare against)
    idx = oracle.knn_canonical(g["pca_points"], int(g["pca_k"]))[0]
    vals, dirs = ap.pca_from_rows(g["pca_points"], idx)
    assert np.allclose(vals[:, 0], l1, rtol=1e-9, atol=1e-18) and np.allclose(vals[:, 1], l2, rtol=1e-9, atol=1e-18)
    assert np.allclose(vals[:, 3], K, rtol=1e-9, atol=1e-30) and np.allclose(vals[:, 4], H, rtol=1e-9, atol=1e-18)


def test_slsqp_does_not_reach_the_minimiser():
    """Why the implicit quadric fit is "parity unpinned" (DESIGN.md section 9): on bunny neighbourhoods the
    UNMODIFIED reference's fit_implicit_quadric_surface (SLSQP from the all-ones start, ref :363-396; outputs stored by
    oracle/make_golden_implicit.py) ends with an objective 1e7 .. 1e11 times the minimum of the problem it poses and
    with a small overlap with the minimiser, the smallest eigenvector of A^T A."""
    g = load_golden("implicit")
    for pts, c, obj, w, v in zip(g["neighbourhoods"], g["slsqp_coeffs"], g["slsqp_objective"], g["eigenvalues"], g["minimiser"]):
        p = pts.astype(np.float32)
        A = np.column_stack((p[:, 0] ** 2, p[:, 1] ** 2, p[:, 2] ** 2, p[:, 0] * p[:, 1], p[:, 0] * p[:, 2], p[:, 1] * p[:, 2],
                             p[:, 0], p[:, 1], p[:, 2], np.ones(len(p))))
        assert abs(np.linalg.norm(c) - 1) < 1e-6                       # the constraint holds ...
        assert np.isclose(np.sum((A @ c) ** 2), obj, rtol=1e-9)
        assert np.allclose(np.linalg.eigvalsh(A.T @ A), w, rtol=1e-3, atol=1e-15 * w[-1])   # (tiny eigenvalues sit at the accuracy limit of eigh: eps * the largest)
        assert obj > 1e6 * max(w[0], 1e-300)                           # ... but the objective is nowhere near its minimum
        assert abs(np.dot(c, v)) < 0.6                                 # and the direction is another one
        assert w[3] < 1e-9 * w[-1]                                     # four nearly flat directions: (n.x) * (a.x + b) vanishes near a plane
