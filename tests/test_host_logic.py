"""CPU checks of the device code's logic and of the C ABI (no GPU needed).

The per-query search and fit routines of the CUDA library are __host__ __device__
(csrc/pct_grid.cuh, csrc/pct_math.cuh); tests/host_harness compiles them for the CPU so
their exactness logic is exercised against the oracle here, before any GPU time is spent.
This harness is test infrastructure -- the product never loads it.
"""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import oracle
from oracle import compare
from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "tests", "host_harness"))


@pytest.fixture(scope="module")
def harness():
    import build as hb

    lib = ctypes.CDLL(hb.build())
    lib.h_build.restype = ctypes.c_void_p
    lib.h_build.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_float]
    lib.h_destroy.argtypes = [ctypes.c_void_p]
    lib.h_knn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7
    lib.h_knn_staged.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 7
    lib.h_set_slab.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_float] * 4
    lib.h_fit_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int] + [ctypes.c_void_p] * 5
    lib.h_ball.argtypes = [ctypes.c_void_p, ctypes.c_double] + [ctypes.c_void_p] * 5
    return lib


def P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run_knn(lib, pts, k, h, max_fast_level=1, staged_u=0, cap_pts=4096, slab=None):
    n = len(pts)
    pts = np.ascontiguousarray(pts, np.float32)
    ix = lib.h_build(P(pts), n, float(h))
    if slab is not None:
        lib.h_set_slab(ix, *slab)
    out = dict(idx=np.zeros((n, k), np.int32), dist=np.zeros((n, k), np.float32), code=np.full(n, -9, np.int32),
               normal=np.zeros((n, 3), np.float32), coeffs=np.zeros((n, 6), np.float32), curv=np.zeros((n, 5), np.float32),
               status=np.zeros(n, np.uint8))
    args = (P(out["idx"]), P(out["dist"]), P(out["code"]), P(out["normal"]), P(out["coeffs"]), P(out["curv"]), P(out["status"]))
    if staged_u:
        lib.h_knn_staged(ix, k, max_fast_level, staged_u, cap_pts, *args)
    else:
        lib.h_knn(ix, k, max_fast_level, *args)
    lib.h_destroy(ix)
    out.update(K=out["curv"][:, 0], H=out["curv"][:, 1], k1=out["curv"][:, 2], k2=out["curv"][:, 3])
    return out


@pytest.mark.parametrize("name,k,stride", [("bunny", 20, 3), ("bunny", 30, 5), ("egg_carton", 20, 9), ("torus_c1", 20, 9)])
def test_search_and_fit_logic_against_oracle(harness, name, k, stride):
    pts = load_golden(name + "_points")["points"][::stride]
    ref = oracle.knn_curvature(pts, k)
    h = 1.25 * float(np.median(ref["dist"][:, -1]))
    got = run_knn(harness, pts, k, h)
    assert (got["code"] >= 0).all()
    assert compare.neighbor_rows_differing(got["idx"], ref["idx"]) == 0
    assert np.array_equal(got["dist"], ref["dist"])
    rep = compare.curvature_report(got, ref, ref["dist"][:, -1])
    assert rep["violations"] == 0, rep
    assert rep["tight_fraction"] > 0.999, rep
    # the frame is the reference's Rodrigues frame, so even the coefficients agree
    scale = np.abs(ref["coeffs"]).max(axis=1, keepdims=True)
    assert np.quantile(np.abs(got["coeffs"] - ref["coeffs"]) / scale, 0.999) < 1e-5


@pytest.mark.parametrize("staged_u", [1, 2])
def test_staged_source_gives_the_same_rows(harness, bunny, staged_u):
    """The staged kernel's region tables (host emulation of its steps A-D) feed the same selection."""
    pts = bunny[::3]
    k = 20
    ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
    h = 1.25 * float(np.median(ref_dist[:, -1]))
    got = run_knn(harness, pts, k, h, staged_u=staged_u)
    assert (got["code"] >= 0).all()
    assert np.mean(got["code"] == 50) > 0.8, np.bincount(got["code"])  # most queries are answered out of the staged copy
    assert compare.neighbor_rows_differing(got["idx"], ref_idx) == 0
    assert np.array_equal(got["dist"], ref_dist)
    # a staging buffer that is too small sends chunks to the L1/L2 path; nothing else changes
    small = run_knn(harness, pts, k, h, staged_u=staged_u, cap_pts=1200)
    assert 0.0 < np.mean(small["code"] == 50) < np.mean(got["code"] == 50)
    assert np.array_equal(small["idx"], got["idx"])


def _random_cloud(rng, case):
    """Small clouds with the traits that stress the search (the GPU suite runs the same families at larger sizes)."""
    n = int(rng.integers(30, 1200))
    kind = case % 5
    if kind == 0:
        p = rng.normal(size=(n, 3))
    elif kind == 1:
        p = np.stack((rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.02 * rng.normal(size=n)), 1)
    elif kind == 2:
        t = rng.uniform(0, 6, n)
        p = np.stack((np.cos(t), np.sin(t), 0.3 * t), 1) + 0.01 * rng.normal(size=(n, 3))
    elif kind == 3:
        c = rng.normal(size=(4, 3)) * 3
        s = np.array([0.01, 0.1, 0.5, 1.0])
        w = rng.integers(0, 4, n)
        p = c[w] + rng.normal(size=(n, 3)) * s[w, None]
    else:
        v = rng.normal(size=(n, 3))
        p = v / np.linalg.norm(v, axis=1, keepdims=True)
        p[rng.integers(0, n, n // 15)] = p[rng.integers(0, n, n // 15)]
    scale = 10.0 ** rng.integers(-3, 4)
    offset = rng.normal(size=3) * scale * 10.0 ** rng.integers(0, 3)
    return (p * scale + offset).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1])
def test_random_clouds_host_logic(harness, seed):
    """The device code's selection on randomised clouds (scales 1e-3 .. 1e3, offsets, sheets, filaments, clusters,
    duplicates), from the sorted cloud and from staged regions, with cell sizes from far too small to far too large:
    rows and distances bit-exact whatever path answers."""
    rng = np.random.default_rng(300 + seed)
    for case in range(10):
        pts = _random_cloud(rng, case)
        n = len(pts)
        k = int(min(n - 2, rng.choice([1, 4, 9, 20, 33])))
        ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
        r_k = float(np.median(ref_dist[:, -1]))
        if not r_k > 0:
            continue
        for factor, staged_u in ((1.25, 0), (1.25, 2), (0.4, 2), (4.0, 1)):
            got = run_knn(harness, pts, k, factor * r_k, max_fast_level=2, staged_u=staged_u)
            tag = (seed, case, n, k, factor, staged_u)
            assert (got["code"] >= 0).all(), tag
            assert compare.neighbor_rows_differing(got["idx"], ref_idx) == 0, tag
            assert np.array_equal(got["dist"], ref_dist), tag


def test_staged_source_on_ties_duplicates_and_tiny_clouds(harness):
    rng = np.random.default_rng(5)
    g = np.arange(9, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    base = rng.normal(size=(1200, 3)).astype(np.float32)
    dup = np.concatenate((base, base[:150]))
    tiny = rng.normal(size=(23, 3)).astype(np.float32)
    for name, pts, k, h in (("lattice", lattice, 12, 1.3), ("dup", dup, 10, 0.3), ("tiny", tiny, 5, 0.7)):
        ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
        for u in (1, 2):
            got = run_knn(harness, pts, k, h, staged_u=u)
            assert compare.neighbor_rows_differing(got["idx"], ref_idx) == 0, (name, u)
            assert np.array_equal(got["dist"], ref_dist), (name, u)


@pytest.mark.parametrize("axis", [0, 2])
def test_slab_indices_reproduce_the_whole_cloud_answer(harness, bunny, axis):
    """Multi-GPU partition: every slab index holds only its own points plus a margin, never lets a search
    radius cross the margin, and still gives the rows of the WHOLE cloud for the queries it owns."""
    pts = np.ascontiguousarray(bunny[::2])
    k = 16
    ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
    h = 1.25 * float(np.median(ref_dist[:, -1]))
    x = pts[:, axis]
    cuts = np.quantile(x, [0.0, 0.3, 0.55, 1.0]).astype(np.float32)
    cuts[0], cuts[-1] = -np.inf, np.inf
    margin = np.float32(4.5 * h)
    answered = np.zeros(len(pts), bool)
    unresolved = 0
    for g in range(3):
        own_lo, own_hi = cuts[g], cuts[g + 1]
        c_lo, c_hi = np.float32(own_lo - margin), np.float32(own_hi + margin)
        sel = np.flatnonzero((x >= c_lo) & (x <= c_hi))
        assert len(sel) < 0.75 * len(pts)  # a real subset
        for staged_u in (0, 2):
            got = run_knn(harness, pts[sel], k, h, max_fast_level=2, staged_u=staged_u,
                          slab=(axis, float(c_lo), float(c_hi), float(own_lo), float(own_hi)))
            own = (x[sel] >= own_lo) & (x[sel] < own_hi)
            assert np.array_equal(got["code"] != -2, own)
            fast = own & (got["code"] >= 0) & (got["code"] % 50 <= 2)   # answered by a grid level, not by the harness's brute force
            assert fast[own].mean() > 0.995
            assert np.array_equal(sel[got["idx"][fast]], ref_idx[sel[fast]])
            assert np.array_equal(got["dist"][fast], ref_dist[sel[fast]])
        answered[sel[own]] = True
        unresolved += int((own & ~fast).sum())
    assert answered.all() and unresolved < 0.005 * len(pts)


def test_cell_size_never_changes_the_answer(harness, bunny):
    pts = bunny[::7]
    ref = oracle.knn_canonical(pts, 12)
    for h in (5e-4, 2e-3, 8e-3, 5e-2, 1.0):
        got = run_knn(harness, pts, 12, h, max_fast_level=2)
        assert compare.neighbor_rows_differing(got["idx"], ref[0]) == 0, h
        assert np.array_equal(got["dist"], ref[1]), h


def test_ties_duplicates_and_outliers(harness):
    rng = np.random.default_rng(11)
    g = np.arange(7, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    base = rng.normal(size=(1500, 3)).astype(np.float32)
    dup = np.concatenate((base, base[:100], base[:20]))
    far = np.concatenate((base * 0.01, rng.normal(size=(4, 3)).astype(np.float32) * 300))
    for name, pts, k, h in (("lattice", lattice, 12, 1.3), ("dup", dup, 10, 0.25), ("far", far, 16, 0.004)):
        ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
        got = run_knn(harness, pts, k, h)
        assert compare.neighbor_rows_differing(got["idx"], ref_idx) == 0, name
        assert np.array_equal(got["dist"], ref_dist), name


def test_fit_rows_on_reference_rows(harness):
    g = load_golden("bunny_k20")
    pts = load_golden("bunny_points")["points"]
    nq, k = g["neighbor_indices"].shape
    normal = np.zeros((nq, 3), np.float32); coeffs = np.zeros((nq, 6), np.float32)
    curv = np.zeros((nq, 5), np.float32); status = np.zeros(nq, np.uint8)
    idx = np.ascontiguousarray(g["neighbor_indices"], np.int32)
    qids = np.ascontiguousarray(g["rows"], np.int32)
    harness.h_fit_rows(P(pts), P(idx), nq, k, P(qids), P(normal), P(coeffs), P(curv), P(status))
    assert not status.any()
    assert np.mean(coeffs == g["quadratic_coefficients"]) > 0.9
    assert np.allclose(curv[:, 0], g["K_quadratic"], rtol=1e-4, atol=1e-2)
    assert np.allclose(curv[:, 1], g["H_quadratic"], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("k", [4, 5])  # k = 2: the normal is arbitrary; k = 3: three points are coplanar, z is rounding noise
def test_fewer_rows_than_coefficients_give_lstsq_minimum_norm(harness, bunny, k):
    """ref :359 on k < 6 rows is underdetermined; lstsq's minimum-norm solution is what the neighbour study sees."""
    pts = np.ascontiguousarray(bunny[::5])
    idx, _, _ = oracle.knn_canonical(pts, k)
    rows = np.arange(0, len(pts), 3)
    sub = np.ascontiguousarray(idx[rows], np.int32)
    ref = oracle.curvature_from_neighbors(pts, sub, rows=rows)
    nq = len(rows)
    normal = np.zeros((nq, 3), np.float32); coeffs = np.zeros((nq, 6), np.float32)
    curv = np.zeros((nq, 5), np.float32); status = np.zeros(nq, np.uint8)
    qids = np.ascontiguousarray(rows, np.int32)
    harness.h_fit_rows(P(pts), P(sub), nq, k, P(qids), P(normal), P(coeffs), P(curv), P(status))
    ok = status == 0
    assert ok.mean() > 0.99
    scale = np.abs(ref["coeffs"]).max(axis=1, keepdims=True)
    err = np.abs(coeffs - ref["coeffs"]) / scale
    assert np.quantile(err[ok], 0.99) < 1e-4, np.quantile(err[ok], [0.5, 0.9, 0.99, 1.0])


def test_rank_deficient_rows_give_lstsq_minimum_norm(harness):
    """Collinear, coincident, planar-lattice and two-parallel-lines neighbourhoods: the reference's lstsq (ref :359)
    returns the minimum-norm solution; fit_min_norm (Givens QR + Jacobi SVD, rcond = eps * max(M, N)) must land on the
    reference's own numbers (tests/golden/degenerate.npz, made by oracle/make_golden_degenerate.py)."""
    g = load_golden("degenerate")
    for name in ("line", "coincident", "plane_grid", "two_lines"):
        pts = np.ascontiguousarray(g[name + "_points"])
        k = int(g[name + "_k"])
        idx = np.ascontiguousarray(g[name + "_neighbor_indices"].astype(np.int32))
        n = len(pts)
        normal = np.zeros((n, 3), np.float32); coeffs = np.zeros((n, 6), np.float32)
        curv = np.zeros((n, 5), np.float32); status = np.zeros(n, np.uint8)
        harness.h_fit_rows(P(pts), P(idx), n, k, None, P(normal), P(coeffs), P(curv), P(status))
        assert np.isfinite(curv).all(), name
        assert np.allclose(curv[:, 0], g[name + "_K"], rtol=1e-4, atol=1e-7), name
        assert np.allclose(curv[:, 1], g[name + "_H"], rtol=1e-4, atol=1e-6), name
        scale = np.maximum(np.abs(g[name + "_coeffs"]).max(axis=1, keepdims=True), 1e-6)
        assert (np.abs(coeffs - g[name + "_coeffs"]) / scale).max() < 1e-3, name
        if name in ("line", "coincident"):
            assert (status == 4).all()       # informational bit: the design was rank deficient
    # random rank-deficient designs against numpy's lstsq on the same rows (the normal is not at stake: planar rows)
    rng = np.random.default_rng(3)
    for trial in range(20):
        k = 12
        a = rng.integers(-4, 5, size=k).astype(np.float32)
        b = rng.integers(0, 2, size=k).astype(np.float32)          # b^2 = b: rank 5
        z = np.zeros(k, np.float32)
        rows_pts = np.stack((a, b, z), 1)
        pts = np.concatenate((np.zeros((1, 3), np.float32), rows_pts))
        idx = np.arange(1, k + 1, dtype=np.int32)[None]
        normal = np.zeros((1, 3), np.float32); coeffs = np.zeros((1, 6), np.float32)
        curv = np.zeros((1, 5), np.float32); status = np.zeros(1, np.uint8)
        qids = np.zeros(1, np.int32)
        harness.h_fit_rows(P(pts), P(idx), 1, k, P(qids), P(normal), P(coeffs), P(curv), P(status))
        assert np.isfinite(coeffs).all() and np.abs(coeffs).max() < 1e-6    # z = 0 in the plane: w = 0


def test_ball_logic(harness, bunny):
    pts = np.ascontiguousarray(bunny[::4])
    n = len(pts)
    radius = 6e-3
    roff, ridx, _ = oracle.ball_canonical(pts, radius)
    for h in (radius * 1.001, radius / 3):  # index built for the radius, or reused from a finer grid
        ix = harness.h_build(P(pts), n, float(h))
        counts = np.zeros(n, np.int32); normal = np.zeros((n, 3), np.float32); coeffs = np.zeros((n, 6), np.float32)
        curv = np.zeros((n, 5), np.float32); status = np.zeros(n, np.uint8)
        harness.h_ball(ix, radius, P(counts), P(normal), P(coeffs), P(curv), P(status))
        harness.h_destroy(ix)
        assert np.array_equal(counts, np.diff(roff))
    rows = np.arange(0, n, 9)
    sub_off = np.concatenate(([0], np.cumsum(np.diff(roff)[rows])))
    sub_idx = np.concatenate([ridx[roff[r]:roff[r + 1]] for r in rows])
    ref = oracle.curvature_from_csr(pts, sub_off, sub_idx, rows=rows)
    got = dict(normal=normal[rows], K=curv[rows, 0], H=curv[rows, 1], k1=curv[rows, 2], k2=curv[rows, 3])
    rep = compare.curvature_report(got, ref, np.full(len(rows), radius), rows_ok=np.diff(roff)[rows] >= 8)
    assert rep["violations"] == 0, rep


@pytest.mark.parametrize("seed", [0, 1])
def test_random_clouds_ball_host_logic(harness, seed):
    """The device code's inclusive radius test (fp32 brackets, fp64 inside them) on randomised clouds: member counts
    equal scipy's query_ball_point at any grid level, empty and crowded balls included."""
    rng = np.random.default_rng(400 + seed)
    for case in range(10):
        pts = _random_cloud(rng, case)
        n = len(pts)
        _, d8, _ = oracle.knn_canonical(pts, min(8, n - 2))
        radius = float(np.quantile(d8[:, -1], rng.choice([0.1, 0.5, 0.9])) * rng.choice([0.5, 1.0, 2.0]))
        if not radius > 0:
            continue
        roff, _, _ = oracle.ball_canonical(pts, radius)
        for h in (radius * 1.001, radius / 2.5):
            ix = harness.h_build(P(pts), n, float(h))
            counts = np.zeros(n, np.int32); normal = np.zeros((n, 3), np.float32); coeffs = np.zeros((n, 6), np.float32)
            curv = np.zeros((n, 5), np.float32); status = np.zeros(n, np.uint8)
            harness.h_ball(ix, radius, P(counts), P(normal), P(coeffs), P(curv), P(status))
            harness.h_destroy(ix)
            assert np.array_equal(counts, np.diff(roff)), (seed, case, n, radius, h)


# ---------------------------------------------------------------------------
# C ABI: the shipped library loads here (no GPU) and exports what the header declares
# ---------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from point_cloud_toolbox_b200 import _lib

    header = open(os.path.join(ROOT, "include", "pct_b200.h")).read()
    declared = set(re.findall(r"\b(pct_[a-z0-9_]+)\s*\(", header))
    declared -= {"pct_index", "pct_index_info", "pct_query_stats"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pct_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib.pct_version() == 200


def test_no_cpu_fallback_without_a_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    import point_cloud_toolbox_b200 as m

    pts = np.random.default_rng(0).normal(size=(100, 3)).astype(np.float32)
    pc = m.PointCloud(points=pts, normals=np.zeros((100, 0), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pc.plant_kdtree(5)
    with pytest.raises(ValueError, match="Either file_path or points and normals"):
        m.PointCloud()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "point_cloud_toolbox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


# ---------------------------------------------------------------------------
# text loader (host code of the library; ref :51-53 np.loadtxt + astype(float32))
# ---------------------------------------------------------------------------
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64 if a.dtype == np.float64 else np.uint32)


def _load_f64(path, threads=0):
    from point_cloud_toolbox_b200 import _lib

    r, c = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(_lib.lib.pct_text_shape(str(path).encode(), ctypes.byref(r), ctypes.byref(c)))
    out = np.empty((r.value, c.value), np.float64)
    _lib.check(_lib.lib.pct_text_load(str(path).encode(), r.value, c.value, P(out), threads))
    return out


@pytest.mark.parametrize("fmt,threads", [("%.18e", 0), ("%.6f", 3), ("%.9g", 1)])
def test_text_loader_equals_loadtxt_bit_for_bit(tmp_path, fmt, threads):
    from point_cloud_toolbox_b200 import engine

    rng = np.random.default_rng(5)
    table = rng.standard_normal((60_000, 6)) * 10.0 ** rng.integers(-20, 20, (60_000, 6))
    path = tmp_path / "t.txt"
    np.savetxt(path, table, fmt=fmt)
    want = np.loadtxt(path)
    got = _load_f64(path, threads)
    assert got.shape == want.shape and np.array_equal(_bits(got), _bits(want))
    got32 = engine.load_text_f32(path, threads).numpy()
    assert got32.dtype == np.float32 and np.array_equal(_bits(got32), _bits(want.astype(np.float32)))


def test_text_loader_format_corners(tmp_path):
    from point_cloud_toolbox_b200 import engine

    path = tmp_path / "c.txt"
    path.write_text("# header\n1 2 3\n\n   \n  4\t5 6 # trailing comment\r\n+7 -8e-400 1e400\nnan inf -Infinity")
    want = np.loadtxt(path)
    got = _load_f64(path)
    assert np.array_equal(_bits(got), _bits(want))           # signed zero, inf, nan payloads included
    path.write_text("1 2 3\n4 5\n")
    with pytest.raises(ValueError, match="row 2"):
        _load_f64(path)
    with pytest.raises(ValueError):
        np.loadtxt(path)
    for bad in ("1,2,3\n", "1 2 x\n", "1 2 3abc\n"):
        path.write_text(bad)
        with pytest.raises(ValueError):
            engine.load_text_f32(path)
        with pytest.raises(ValueError):
            np.loadtxt(path)
    path.write_text("1 2 3\n")                                 # np.loadtxt squeezes one row
    assert engine.load_text_f32(path).shape == (3,) and np.loadtxt(path).shape == (3,)
    path.write_text("")
    assert engine.load_text_f32(path).shape == (0,)
    with pytest.raises(FileNotFoundError):
        engine.load_text_f32(tmp_path / "missing.txt")


def test_text_loader_on_the_reference_scan_fixture():
    # the fp32 cloud of the golden fixture came from the unmodified reference's np.loadtxt (oracle/make_golden.py)
    g = load_golden("loader_case")
    import tempfile

    from point_cloud_toolbox_b200 import engine

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "cloud.txt")
        np.savetxt(path, g["table"], fmt="%.5f")
        t = engine.load_text_f32(path).numpy()
    pts = t[:, 0:3].copy()
    pts[:, 0] -= pts[:, 0].max()
    pts[:, 1] -= pts[:, 1].max()
    assert np.array_equal(pts, g["points"]) and np.array_equal(t[:, 3:6], g["normals"])


# ---------------------------------------------------------------------------
# PLY reader and the two writers (host code of the library): byte- / bit-identical to the reference
# ---------------------------------------------------------------------------
def test_ply_writers_and_reader_reproduce_the_reference_files(tmp_path):
    from point_cloud_toolbox_b200 import utils as U

    g = load_golden("io_energy_pca")
    for tag in ("f64", "f32"):
        path = tmp_path / f"p_{tag}.ply"
        U.save_points_to_ply(g[f"points_ply_in_{tag}"], str(path))
        assert path.read_bytes() == g[f"points_ply_bytes_{tag}"].tobytes()
    path = tmp_path / "c.ply"
    U.save_curvatures_to_ply(g["curv_ply_points"], g["curv_ply_K"], g["curv_ply_H"], str(path))
    assert path.read_bytes() == g["curv_ply_bytes"].tobytes()
    path = tmp_path / "in.ply"
    path.write_bytes(g["parse_ply_text"].tobytes())
    got = U.parse_ply(str(path))
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), g["parse_ply_points"].view(np.uint32))
    # the reference's error behaviour: None for a missing file, a short line, no header
    assert U.parse_ply(str(tmp_path / "missing.ply")) is None
    path.write_text("ply\nend_header\n1 2 3\n4 5\n")
    assert U.parse_ply(str(path)) is None
    path.write_text("1 2 3\n")
    assert U.parse_ply(str(path)) is None
    path.write_text("ply\nend_header\n")
    assert U.parse_ply(str(path)).shape == (0,)


def test_float_formatting_against_python_on_random_bit_patterns(tmp_path):
    from oracle import around_path as ap
    from point_cloud_toolbox_b200 import utils as U

    rng = np.random.default_rng(8)
    n = 70_001                                            # more than one block of rows per thread
    bits = rng.integers(0, 2 ** 32, (n, 5), dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32)
    vals[:, 3] = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 17, n)).astype(np.float32)   # the fixed-notation band
    path = tmp_path / "c.ply"
    U.save_curvatures_to_ply(vals[:, :3], vals[:, 3], vals[:, 4], str(path))
    assert path.read_bytes() == ap.curvature_ply_bytes(vals[:, :3], vals[:, 3], vals[:, 4])
    pts = (rng.standard_normal((n, 3)) * 10.0 ** rng.integers(-9, 12, (n, 3)))
    pts[::97] = np.round(pts[::97] * 2e6) / 2e6           # exact ties of the sixth decimal
    path = tmp_path / "p.ply"
    U.save_points_to_ply(pts, str(path))
    assert path.read_bytes() == ap.points_ply_bytes(pts)
    back = U.parse_ply(str(path))                          # and the reader on what the writer wrote
    assert np.array_equal(back, ap.parse_ply(path.read_text()))


# ---------------------------------------------------------------------------
# energy integration and PCA rows: the kernels' __host__ __device__ arithmetic against the oracle
# ---------------------------------------------------------------------------
def test_energy_terms_against_reference_numbers(harness):
    from oracle import around_path as ap

    harness.h_mesh_energies.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    g = load_golden("io_energy_pca")
    v, t = np.ascontiguousarray(g["energy_vertices"]), np.ascontiguousarray(g["energy_triangles"])
    K, H = np.ascontiguousarray(g["energy_K"]), np.ascontiguousarray(g["energy_H"])
    out = np.zeros(4)
    harness.h_mesh_energies(P(v), len(v), P(t), len(t), P(K), P(H), P(out))
    assert out[3] == 0 and np.allclose(out[:3], g["energy_result"], rtol=1e-13, atol=1e-14)   # summation order only
    harness.h_mesh_energies(P(v), len(v), P(t), len(t), None, None, P(out))
    assert np.allclose(out[:3], g["energy_result_no_curvature"], rtol=1e-13, atol=0)
    # per-triangle terms are exact: one triangle at a time equals the oracle bit for bit
    for i in range(0, len(t), 37):
        ti = np.ascontiguousarray(t[i:i + 1])
        harness.h_mesh_energies(P(v), len(v), P(ti), 1, P(K), P(H), P(out))
        want = ap.mesh_energies(v, ti, K, H)
        assert tuple(out[:3]) == tuple(float(x) for x in want)
    bad = np.array([[0, 1, len(v)], [0, 1, -1], [0, 1, -len(v) - 1]], np.int32)
    harness.h_mesh_energies(P(v), len(v), P(bad), 3, P(K), P(H), P(out))
    assert out[3] == 2


def test_pca_rows_against_reference_numbers(harness):
    harness.h_pca_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p]
    from oracle import around_path as ap

    g = load_golden("io_energy_pca")
    pts, k = np.ascontiguousarray(g["pca_points"]), int(g["pca_k"])
    idx = np.ascontiguousarray(oracle.knn_canonical(pts, k)[0])
    vals, dirs = np.zeros((len(pts), 6)), np.zeros((len(pts), 3, 2))
    harness.h_pca_rows(P(pts), P(idx), len(pts), k, 0, P(vals), P(dirs))
    scale = g["pca_l1"]                                   # errors of an eigenvalue are relative to the largest
    assert np.all(np.abs(vals[:, 0] - g["pca_l1"]) <= 1e-12 * scale)
    assert np.all(np.abs(vals[:, 1] - g["pca_l2"]) <= 1e-12 * scale)
    assert np.allclose(vals[:, 3], g["pca_K"], rtol=1e-9, atol=0) and np.allclose(vals[:, 4], g["pca_H"], rtol=1e-12, atol=0)
    # directions up to sign, where the eigenvalues are separated
    ref = g["pca_directions"]
    gap = np.minimum(vals[:, 0] - vals[:, 1], vals[:, 1] - vals[:, 2]) / vals[:, 0]
    ok = gap > 1e-3
    dots = np.abs(np.einsum("nij,nij->nj", dirs, ref))
    assert ok.mean() > 0.9 and np.all(dots[ok] > 1 - 1e-9)
    want, _ = ap.pca_from_rows(pts, idx, include_self=True)
    harness.h_pca_rows(P(pts), P(idx), len(pts), k, 1, P(vals), P(dirs))
    assert np.all(np.abs(vals[:, :3] - want[:, :3]) <= 1e-12 * want[:, :1])
    assert np.allclose(vals[:, 5], want[:, 5], rtol=1e-7, atol=1e-12)


def test_text_loader_fuzz_against_loadtxt(tmp_path):
    """Random small files in the dialect np.loadtxt reads by default: the table is the same bit for bit, or both refuse."""
    from point_cloud_toolbox_b200 import engine

    rng = np.random.default_rng(21)

    def token():
        kind = rng.integers(0, 9)
        v = float(rng.standard_normal() * 10.0 ** rng.integers(-12, 12))
        if kind == 0:
            return str(int(rng.integers(-10 ** 6, 10 ** 6)))
        if kind == 1:
            return f"{v:.{rng.integers(0, 18)}e}"
        if kind == 2:
            return f"{v:.{rng.integers(0, 12)}f}"
        if kind == 3:
            return repr(float(np.float32(v)))
        if kind == 4:
            return "+" + repr(abs(v))
        if kind == 5:
            return rng.choice(["nan", "inf", "-inf", "NaN", "Infinity", "1e400", "-1e-400", "0", "-0.0", ".5", "5.", "1E5"])
        if kind == 6:
            return f"{v:.3g}".upper()
        return repr(v)

    path = tmp_path / "f.txt"
    agree = 0
    for case in range(150):
        cols = int(rng.integers(1, 7))
        rows = int(rng.integers(2, 40))
        lines = []
        for r in range(rows):
            sep = rng.choice([" ", "  ", "\t", " \t "])
            line = sep.join(token() for _ in range(cols))
            if rng.random() < 0.1:
                line = "   " + line
            if rng.random() < 0.1:
                line += "  # note " + token()
            if rng.random() < 0.08:
                lines.append(rng.choice(["", "   ", "# only a comment", "\t"]))
            lines.append(line)
        if case % 10 == 9:                                  # a broken file now and then
            lines[int(rng.integers(0, len(lines)))] += rng.choice([" 1", " x", ",2"])
        eol = "\r\n" if case % 7 == 3 else "\n"
        text = eol.join(lines) + (eol if case % 3 else "")
        path.write_bytes(text.encode())
        try:
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = np.loadtxt(path, ndmin=2)
        except ValueError:
            with pytest.raises(ValueError):
                engine.load_text_f32(path)
            continue
        got = _load_f64(path)
        assert got.shape == want.shape, (case, text)
        assert np.array_equal(_bits(got), _bits(want)), (case, text)
        agree += 1
    assert agree > 100


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/pct_b200.h compiles as C99 and a C program links against the library."""
    import shutil
    import subprocess

    from point_cloud_toolbox_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "pct_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    long long rows = -1, cols = -1;\n'
                   '    if (pct_version() != 200) return 1;\n'
                   '    if (pct_text_shape("/nonexistent/file", (int64_t*)&rows, (int64_t*)&cols) == PCT_OK) return 2;\n'
                   '    printf("%s\\n", pct_last_error());\n'
                   '    return 0;\n}\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lpct_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "cannot open" in out


def test_io_code_under_address_and_ub_sanitizers(tmp_path):
    """csrc/pct_io.cu built alone with -fsanitize=address,undefined; random bit patterns through the writers and readers,
    garbage files through the parsers (tests/host_harness/io_sanitize)."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    here = os.path.join(ROOT, "tests", "host_harness", "io_sanitize")
    exe = tmp_path / "io_sanitize"
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-x", "c++", "-D__uint_as_float(x)=(0.f)", "-I", cuda_inc, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "point_cloud_toolbox_b200", "csrc", "pct_io.cu"), os.path.join(here, "stub.cpp"),
           os.path.join(here, "main.cpp"), "-o", str(exe), "-pthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr and "cannot find" in res.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert res.returncode == 0, res.stderr[-2000:]
    run = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert run.returncode == 0 and "asan run done" in run.stdout, (run.stdout[-500:], run.stderr[-3000:])


def test_pinned_result_buffers_are_leased_until_every_view_is_gone(monkeypatch):
    """engine._PinnedPool: a recycled host buffer must never be handed out while an array (or a view of a view)
    of the previous result is alive; the lease is tied to the numpy array with weakref.finalize."""
    import gc
    import weakref

    import torch

    from point_cloud_toolbox_b200 import engine

    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "pin_memory"}))
    pool = engine._PinnedPool()
    h, lease = pool.take((10, 2), torch.float32)
    arr = h.numpy()
    weakref.finalize(arr, engine._PinnedPool.release, lease)
    view = arr[:, 0]
    inner = view[2:5]
    del arr, h
    gc.collect()
    assert lease[1] and pool.take((10, 2), torch.float32)[1] is not lease      # a second buffer, not the leased one
    del view
    gc.collect()
    assert lease[1]                                                             # a view of a view still sees the data
    del inner
    gc.collect()
    assert not lease[1]
    assert pool.take((10, 2), torch.float32)[1] is lease                        # recycled now
