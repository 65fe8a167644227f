"""GPU parity: the CUDA path (through the C ABI, via the PointCloud drop-in) against the
oracle and the reference's golden vectors.  Everything here needs a B200: -m gpu.

Tolerances are the policy of oracle/compare.py (relative 1e-3 + absolute floor scaled by
r_k, neighbour sets bit-exact); the observed agreement is asserted much tighter where the
path is expected to reproduce the reference's fp64 steps (tight_fraction).
"""
import os

import numpy as np
import pytest

import oracle
from oracle import compare, datasets
from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def pct():
    import point_cloud_toolbox_b200 as m

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return m


def _cloud(name):
    return load_golden(name + "_points")["points"]


def _empty_normals(n):
    return np.zeros((n, 0), np.float32)


def _gpu_dict(pc):
    return dict(normal=pc.normals_quadratic, K=pc.K_quadratic, H=pc.H_quadratic, k1=pc.k1_quadratic, k2=pc.k2_quadratic)


CASES = [("bunny", 20), ("bunny", 30), ("egg_carton", 20), ("torus_c1", 20)]


@pytest.mark.parametrize("name,k", CASES)
def test_knn_lists_bit_exact(pct, name, k):
    pts = _cloud(name)
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
    pc.plant_kdtree(k)
    idx, dist = pc.neighbor_indices, pc.dists
    assert idx.dtype == np.int32 and dist.dtype == np.float32 and idx.shape == (len(pts), k)
    ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
    assert compare.neighbor_rows_differing(idx, ref_idx) == 0
    assert np.array_equal(dist, ref_dist)
    # and the reference's own rows (scipy order) wherever no exact tie is involved
    g = load_golden(f"{name}_k{k}")
    same = (idx[g["rows"]] == g["neighbor_indices"]).all(axis=1)
    assert np.array_equal(dist[g["rows"]], g["dists"])
    if name == "bunny":
        assert same.all()
    else:
        assert same.mean() > 0.9


@pytest.mark.parametrize("name,k", CASES)
def test_fused_curvature_matches_oracle(pct, name, k):
    pts = _cloud(name)
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
    pc.plant_kdtree(k)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    assert K.dtype == np.float32 and K.shape == (len(pts),)
    ref = oracle.knn_curvature(pts, k)
    rep = compare.curvature_report(_gpu_dict(pc), ref, ref["dist"][:, -1])
    assert rep["violations"] == 0, rep
    assert rep["tight_fraction"] > 0.999, rep
    assert not np.isnan(K).any() and not np.isnan(H).any()
    # golden: the reference's own K / H on rows where its neighbour row equals the canonical one
    g = load_golden(f"{name}_k{k}")
    canon = (ref["idx"][g["rows"]] == g["neighbor_indices"]).all(axis=1)
    Kg, Hg = g["K_quadratic"][canon], g["H_quadratic"][canon]
    rk = g["dists"][canon, -1].astype(np.float64)
    assert np.all(np.abs(K[g["rows"]][canon] - Kg) <= 1e-3 * np.abs(Kg) + 1e-5 / rk ** 2)
    safe = ref["margin"][g["rows"]][canon] >= compare.MARGIN
    dH = np.abs(H[g["rows"]][canon] - Hg)
    assert np.all(dH[safe] <= 1e-3 * np.abs(Hg[safe]) + 1e-5 / rk[safe])
    st = pc.kdtree.index.last_stats()
    assert st.queries == len(pts)


@pytest.mark.parametrize("name,k", [("bunny", 20), ("egg_carton", 20)])
def test_fit_from_reference_rows_reproduces_reference(pct, name, k):
    """pct_fit_from_neighbors on the reference's own neighbour rows vs the reference's coefficients."""
    g = load_golden(f"{name}_k{k}")
    pts = _cloud(name)
    d_pts = torch.from_numpy(pts).cuda()
    idx = torch.from_numpy(g["neighbor_indices"]).cuda()
    qids = torch.from_numpy(g["rows"].astype(np.int32)).cuda()
    out = pct.fit_from_neighbors(d_pts, idx, qids)
    coeffs = out.coeffs.cpu().numpy()
    ref_c = g["quadratic_coefficients"]
    scale = np.abs(ref_c).max(axis=1, keepdims=True)
    assert np.max(np.abs(coeffs - ref_c) / scale) < 2e-5
    assert np.mean(coeffs == ref_c) > 0.9  # the same fp32 numbers, not merely close ones
    curv = out.curv.cpu().numpy()
    rk = g["dists"][:, -1].astype(np.float64)
    assert np.all(np.abs(curv[:, 0] - g["K_quadratic"]) <= 1e-4 * np.abs(g["K_quadratic"]) + 1e-6 / rk ** 2)
    assert np.all(np.abs(curv[:, 1] - g["H_quadratic"]) <= 1e-4 * np.abs(g["H_quadratic"]) + 1e-6 / rk)
    assert np.allclose(curv[:, 4], g["K_H_sq_quadratic"], rtol=1e-4, atol=1e-6)


def test_list_path_equals_fused_path(pct, bunny):
    k = 20
    a = pct.PointCloud(points=bunny, normals=_empty_normals(len(bunny)))
    a.plant_kdtree(k)
    Ka, Ha = a.compute_pointwise_explicit_quadratic_curvature()
    b = pct.PointCloud(points=bunny, normals=_empty_normals(len(bunny)))
    b.plant_kdtree(k)
    _ = b.neighbor_indices  # materialise -> the fit now runs from the stored rows (ref :640)
    b.fit_explicit_quadratic_surfaces_to_neighborhoods()
    Kb, Hb = b.calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
    assert np.allclose(Ka, Kb, rtol=1e-5, atol=1e-3) and np.allclose(Ha, Hb, rtol=1e-5, atol=1e-4)
    assert np.allclose(b.K_H_sq_quadratic, Hb * Hb, rtol=1e-6)
    assert b.quadratic_coefficients.shape == (len(bunny), 6)
    assert np.allclose(a.quadratic_coefficients, b.quadratic_coefficients, rtol=1e-4, atol=1e-6)


def test_validate_shape_call_sequence(pct, bunny):
    """utils.py:484-501: plant(100), fit, re-plant(k), curvature -> curvature belongs to the k=100 fit."""
    pts = bunny[:12000]
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)))
    pc.plant_kdtree(100)
    pc.fit_explicit_quadratic_surfaces_to_neighborhoods()
    pc.plant_kdtree(23)
    K, H = pc.calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
    rows = np.arange(0, len(pts), 6)
    ref = oracle.knn_curvature(pts, 100, rows=rows)
    got = dict(normal=pc.normals_quadratic[rows], K=K[rows], H=H[rows], k1=pc.k1_quadratic[rows], k2=pc.k2_quadratic[rows])
    rep = compare.curvature_report(got, ref, ref["dist"][:, -1])
    assert rep["violations"] == 0, rep
    assert pc.k_neighbors == 23


def test_static_methods(pct):
    g = load_golden("bunny_k20")
    pts = _cloud("bunny")
    pos = {int(r): j for j, r in enumerate(g["rows"])}
    for j, i in enumerate(g["static_rows"][:16]):
        nb = g["neighbor_indices"][pos[int(i)]]
        centered = pts[nb] - pts[i]
        rot = pct.PointCloud.get_best_fit_plane_and_rotate(centered)
        assert rot.dtype == np.float64 and rot.shape == (20, 3)
        assert np.allclose(rot, g["static_rotated"][j], rtol=0, atol=1e-12 * np.abs(centered).max() * 1e3)
        coeffs = pct.PointCloud.fit_quadratic_surface(g["static_rotated"][j])
        want = oracle.fit_quadratic_surface(g["static_rotated"][j])
        assert coeffs.dtype == np.float32
        assert np.allclose(coeffs, want, rtol=2e-6, atol=1e-9)
        got = pct.PointCloud.calculate_explicit_quadratic_curvatures(want)
        ref = oracle.explicit_quadratic_curvatures(want)
        assert np.allclose(got, ref, rtol=2e-6, atol=0)
    with pytest.raises(ValueError):
        pct.PointCloud.get_best_fit_plane_and_rotate(np.array([[0, 0, np.inf], [1, 0, 0], [0, 1, 0]], np.float32))
    with pytest.raises(ValueError):
        pct.PointCloud.fit_quadratic_surface(np.zeros((5, 2)))


@pytest.mark.parametrize("radius", [3.8e-3, 7.6e-3])
def test_ball_query_matches_scipy(pct, bunny, radius):
    pc = pct.PointCloud(points=bunny, normals=_empty_normals(len(bunny)))
    pc.plant_ball(radius)
    off, idx, dist = pc.ball_neighbors()
    roff, ridx, rdist = oracle.ball_canonical(bunny, radius)
    assert compare.csr_equal(off, idx, roff, ridx)
    assert np.array_equal(dist, rdist)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    rows = np.arange(0, len(bunny), 5)
    sub_off = np.concatenate(([0], np.cumsum(np.diff(roff)[rows])))
    sub_idx = np.concatenate([ridx[roff[r]:roff[r + 1]] for r in rows])
    ref = oracle.curvature_from_csr(bunny, sub_off, sub_idx, rows=rows)
    enough = np.diff(roff)[rows] >= 8
    r_k = np.full(len(rows), radius)
    got = dict(normal=pc.normals_quadratic[rows], K=K[rows], H=H[rows], k1=pc.k1_quadratic[rows], k2=pc.k2_quadratic[rows])
    rep = compare.curvature_report(got, ref, r_k, rows_ok=enough)
    assert rep["violations"] == 0, rep
    assert rep["rows"] > 0.9 * len(rows)


def test_exact_path_duplicates_lattice_outliers(pct):
    rng = np.random.default_rng(7)
    # 3-D integer lattice: huge exact tie groups at the k-th distance
    g = np.arange(9, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    # duplicated points + far outliers + a dense cluster
    base = rng.normal(size=(3000, 3)).astype(np.float32)
    dup = np.concatenate((base, base[:200], base[:50]))
    outliers = np.concatenate((base * 0.01, rng.normal(size=(5, 3)).astype(np.float32) * 500))
    for name, pts, k in (("lattice", lattice, 12), ("dup", dup, 10), ("outliers", outliers, 20), ("tiny", base[:9], 8)):
        pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)))
        pc.plant_kdtree(k)
        ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
        assert compare.neighbor_rows_differing(pc.neighbor_indices, ref_idx) == 0, name
        assert np.array_equal(pc.dists, ref_dist), name
    st = pc.kdtree.index.last_stats()
    assert st.queries == 9


def test_neighbor_study_matches_reference_and_oracle(pct, bunny):
    """explicit_quadratic_neighbor_study (ref :732-800): same sample, same probes, same return value."""
    g = load_golden("neighbor_study")
    pc = pct.PointCloud(points=bunny, normals=np.zeros((len(bunny), 0), np.float32), k_neighbors=20)
    pc.plant_kdtree(20)
    size = int(g["sample_size"])
    for a, seed in enumerate(g["seeds"]):
        for b, tol in enumerate(g["tols"]):
            np.random.seed(int(seed))
            got = pc.explicit_quadratic_neighbor_study(tol=float(tol), sample_size=size)
            assert got == int(g["results"][a, b]), (seed, tol, got, int(g["results"][a, b]))
    # per-point converged counts against the oracle on another sample and bounds
    rng = np.random.default_rng(3)
    sample = rng.integers(0, len(bunny), 40)
    want, counts = oracle.neighbor_study(bunny, sample, tol=80.0, lower_bound=4, upper_bound=60, return_counts=True)
    state = np.random.get_state()
    np.random.seed(123)
    draw = np.random.randint(0, len(bunny), 40)
    np.random.seed(123)
    got = pc.explicit_quadratic_neighbor_study(tol=80.0, sample_size=40, lower_bound=4, upper_bound=60)
    np.random.set_state(state)
    assert got == oracle.neighbor_study(bunny, draw, tol=80.0, lower_bound=4, upper_bound=60)
    assert 4 <= want - 1 <= 60 and len(counts) == 40


def test_knn_points_rows_equal_full_lists(pct, bunny):
    pts = bunny[::2]
    d = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    ix = pct.GridIndex(d, k_hint=20)
    ids = np.random.default_rng(0).integers(0, len(pts), 300).astype(np.int32)
    for k in (7, 40, 100):
        ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k, rows=ids)
        idx, dist = ix.knn_points(ids, k)
        assert np.array_equal(idx.cpu().numpy(), ref_idx)
        assert np.array_equal(dist.cpu().numpy(), ref_dist)
    with pytest.raises(IndexError):
        ix.knn_points(np.array([len(pts)], np.int32), 5)
    ix.close()


@pytest.mark.parametrize("world", [2, 5])
def test_slab_mode_equals_whole_cloud_index(pct, world):
    """Multi-GPU slab partition, its ranks run one after the other on this GPU: every point is answered by
    exactly one slab index and the answers are those of the whole-cloud index."""
    from point_cloud_toolbox_b200 import distributed as pdist

    pts, _, _ = datasets.torus_random(200_000, seed=4)
    k = 20
    cloud = torch.from_numpy(pts).cuda()
    whole = pct.GridIndex(cloud, k_hint=k)
    ref = whole.curvature_knn(k, want_coeffs=False).records
    got = torch.full_like(ref, float("nan"))
    seen = torch.zeros(len(pts), dtype=torch.int32, device="cuda")
    for rank in range(world):
        part = pdist.curvature_knn_slab(cloud, k, rank, world)
        assert part.index.n < 0.8 * len(pts)            # a slab index holds its slab and a margin, not the cloud
        assert part.unresolved == 0
        got[part.ids] = part.records
        seen[part.ids] += 1
        part.index.close()
    assert bool((seen == 1).all())
    ok = ~torch.isnan(ref[:, 3])
    assert bool(ok.float().mean() > 0.9999)
    rel = ((got[:, 3:7] - ref[:, 3:7]).abs() / ref[:, 3:7].abs().clamp_min(1e-3))[ok]
    assert float(rel.max()) < 1e-4, float(rel.max())     # same neighbours, sums in another order
    assert bool(((got[:, :3] * ref[:, :3]).sum(1)[ok] > 0.99999).all())
    # neighbour rows of a slab index are the whole cloud's rows
    part = pdist.curvature_knn_slab(cloud, k, 1, world)
    idx_local, dist_local = part.index.knn(k)
    sel, own = pdist.slab_select(cloud[:, part.axis], part.bounds)
    idx_whole, dist_whole = whole.knn(k)
    assert idx_local.shape[0] == int(own.sum())            # compact rows: one per owned point, in ascending original index
    owned = sel[own]
    assert torch.equal(sel[idx_local.long()], idx_whole[owned].long())
    assert torch.equal(dist_local, dist_whole[owned])
    part.index.close()
    whole.close()


def test_slab_mode_keeps_the_tie_order_of_the_whole_cloud(pct):
    """A lattice is nothing but ties: slab indices must break them by the WHOLE cloud's index order."""
    from point_cloud_toolbox_b200 import distributed as pdist

    g = np.arange(24, dtype=np.float32)
    pts = np.stack(np.meshgrid(g, g * 0.5, g * 0.25, indexing="ij"), -1).reshape(-1, 3)
    pts = np.ascontiguousarray(pts[np.random.default_rng(2).permutation(len(pts))])
    k = 12
    ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
    cloud = torch.from_numpy(pts).cuda()
    for rank in range(3):
        part = pdist.curvature_knn_slab(cloud, k, rank, 3)
        idx_local, dist_local = part.index.knn(k)
        sel, own = pdist.slab_select(cloud[:, part.axis], part.bounds)
        owned = sel[own].cpu().numpy()
        assert np.array_equal(sel[idx_local.long()].cpu().numpy(), ref_idx[owned])
        assert np.array_equal(dist_local.cpu().numpy(), ref_dist[owned])
        part.index.close()


def test_fewer_neighbours_than_coefficients_through_the_class(pct, bunny):
    """k < 6: lstsq's minimum-norm solution (ref :359), served by the rows path."""
    pts = np.ascontiguousarray(bunny[::5])
    k = 4
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
    pc.plant_kdtree(k)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    rows = np.arange(0, len(pts), 7)
    ref = oracle.knn_curvature(pts, k, rows=rows)
    ok = np.isfinite(ref["K"]) & np.isfinite(K[rows])
    assert ok.mean() > 0.99
    scale = np.abs(ref["coeffs"]).max(axis=1)
    err = np.abs(np.asarray(pc.quadratic_coefficients)[rows] - ref["coeffs"]).max(axis=1) / scale
    assert np.quantile(err[ok], 0.99) < 1e-4
    assert pct._lib.lib.pct_release_scratch() == 0


def test_unresolved_slab_queries_fall_back_to_the_whole_cloud(pct):
    """A cloud with far outliers: their k-th neighbour lies beyond any margin; the slab index says so
    (PCT_STATUS_UNRESOLVED) and the rank answers them from a whole-cloud index."""
    from point_cloud_toolbox_b200 import distributed as pdist

    rng = np.random.default_rng(9)
    base, _, _ = datasets.torus_random(60_000, seed=5)
    # 12 < k + 1 points high above the torus, in the middle of x and y (an interior slab whichever of the
    # two is the longest axis): their k-th neighbour is ~1.7 away, far beyond the slab's margin
    far = (rng.normal(size=(12, 3)) * 0.02 + np.array([0.1, 0.1, 2.0])).astype(np.float32)
    pts = np.concatenate((base, far)).astype(np.float32)
    k = 20
    cloud = torch.from_numpy(pts).cuda()
    whole = pct.GridIndex(cloud, k_hint=k)
    ref = whole.curvature_knn(k, want_coeffs=False).records
    got = torch.full_like(ref, float("nan"))
    unresolved = 0
    for rank in range(4):
        part = pdist.curvature_knn_slab(cloud, k, rank, 4)
        unresolved += part.unresolved
        got[part.ids] = part.records
        part.index.close()
    assert unresolved >= 12
    rel = (got[:, 3:5] - ref[:, 3:5]).abs() / ref[:, 3:5].abs().clamp_min(1e-3)
    assert float(rel.max()) < 1e-4
    whole.close()


def _exchange_rank(rank, world, port, n, k, out_path):
    """One NCCL rank of the slab-exchange path: its rows against the oracle, and a sampled parity block."""
    import torch.distributed as dist

    from oracle import sample_parity
    from point_cloud_toolbox_b200 import distributed as pdist

    torch.cuda.set_device(rank % torch.cuda.device_count())
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        pts, _, _ = datasets.torus_random(n, seed=6)
        pts[7] = (0.0, 0.0, 0.9)                          # isolated: its neighbours lie beyond any margin
        cloud = torch.from_numpy(pts).to(dev)
        b, e = pdist.shard_bounds(n, world, rank)
        part = pdist.curvature_knn_exchange(cloud[b:e].contiguous(), b, n, k, columns=(3, 4, 5, 6))
        rows = np.arange(b, e, 37)
        ref = oracle.knn_curvature(pts, k, rows=rows)
        got_all = part.rows.cpu().numpy()
        sel = rows - b
        rec = part.records
        own_ids = part.own_ids.long()
        # curvature of the rows that came back to this rank (answered by whichever slab owns them)
        dummy_n = ref["normal"]
        got = dict(normal=dummy_n, K=got_all[sel, 0], H=got_all[sel, 1], k1=got_all[sel, 2], k2=got_all[sel, 3])
        rep = compare.curvature_report(got, ref, ref["dist"][:, -1])
        # neighbour rows + records of the slab this rank answered, on Morton runs incl. the cut planes
        c_lo, c_hi, own_lo, own_hi = part.plan.bounds[rank]
        axis = part.plan.axis
        runs = sample_parity.choose_runs(part.index, 4, 2048, seed=rank, planes=(own_lo, own_hi), axis=axis)
        par = sample_parity.check_runs(part.index, cloud, k, runs, local_to_orig=part.local_ids,
                                       owned=lambda c: (c[:, axis] >= own_lo) & (c[:, axis] < own_hi),
                                       records_of=lambda ids: rec[torch.searchsorted(own_ids, ids)])
        # the host-array form: every rank moves its share
        name = f"pct_gpu_test_{port}"
        if rank == 0:
            cin = pdist.SharedHostArray(name + "_in", pts.shape, create=True)
            cout = pdist.SharedHostArray(name + "_out", (2, n), create=True)
            cin.array[:] = pts
            cout.array[:] = np.nan
        dist.barrier()
        if rank != 0:
            cin = pdist.SharedHostArray(name + "_in", pts.shape, create=False)
            cout = pdist.SharedHostArray(name + "_out", (2, n), create=False)
        st = pdist.Stages()
        pdist.curvature_knn_shared(cin, cout, k, stages=st).close()
        host_ok = bool(np.array_equal(cout.array[0, b:e], got_all[:, 0]) and np.array_equal(cout.array[1, b:e], got_all[:, 1]))
        all_written = bool(np.isfinite(cout.array).mean() > 0.999)
        # the host copies in rounds (what CopyRounds picks on boxes whose GPUs share PCIe uplinks): same arrays
        first = cout.array.copy()
        dist.barrier()
        cout.array[:, b:e] = np.nan
        os.environ["PCT_COPY_ROUNDS"] = "2,2" if world % 2 == 0 else "1,1"
        pdist.CopyRounds._cache.clear()
        pdist.curvature_knn_shared(cin, cout, k).close()
        del os.environ["PCT_COPY_ROUNDS"]
        pdist.CopyRounds._cache.clear()
        host_ok = host_ok and bool(np.array_equal(cout.array, first, equal_nan=True))
        stages = st.durations_ms()
        unresolved = torch.tensor([part.unresolved], device=dev)
        dist.all_reduce(unresolved)
        dist.barrier()
        cin.close()
        cout.close()
        part.close()
        # the return fused into the kernel (K, H stored straight into the owner rank's array over peer memory) against
        # the NCCL return, on a cloud without the isolated point (an unresolved query sends everything through NCCL)
        clean = torch.from_numpy(datasets.torus_random(n, seed=6)[0]).to(dev)
        fused = pdist.curvature_knn_exchange(clean[b:e].contiguous(), b, n, k)
        again = pdist.curvature_knn_exchange(clean[b:e].contiguous(), b, n, k)      # the peer arrays are reused
        os.environ["PCT_PEER_RETURN"] = "0"
        pdist.PeerResults._cache.clear()
        plain = pdist.curvature_knn_exchange(clean[b:e].contiguous(), b, n, k)
        del os.environ["PCT_PEER_RETURN"]
        pdist.PeerResults._cache.clear()
        same = bool(torch.equal(fused.rows.view(torch.int32), plain.rows.view(torch.int32)) and
                    torch.equal(again.rows.view(torch.int32), plain.rows.view(torch.int32)))
        modes = (bool(fused.peer_return), bool(plain.peer_return), int(fused.unresolved) + int(plain.unresolved))
        for f in (fused, again, plain):
            f.close()
        pdist.PeerResults.release()
        np.save(out_path + f".{rank}.npy", np.array([rep["violations"], rep["rows"], par["rows_differing"], par["dist_differing"],
                                                    par["violations"], par["rows"], float(host_ok), float(all_written),
                                                    float(unresolved.item()), float(len(stages)), float(same), float(modes[0]),
                                                    float(modes[1]), float(modes[2])]))
    finally:
        dist.destroy_process_group()


def _run_exchange(tmp_path, world, n, k):
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "ex")
    mp.spawn(_exchange_rank, args=(world, port, n, k, out), nprocs=world, join=True)
    for r in range(world):
        res = np.load(out + f".{r}.npy")
        assert res[0] == 0 and res[1] > 100, res          # curvature of the returned rows within tolerance
        assert res[2] == 0 and res[3] == 0 and res[5] > 1000, res   # neighbour rows and distances bit-exact
        assert res[4] == 0, res                           # records of the slab within tolerance
        assert res[6] == 1.0 and res[7] == 1.0, res       # host arrays = device rows, everything written
        assert res[9] >= 6
        # fused peer return = NCCL return bit for bit; the fused mode was really on; nothing unresolved on the clean cloud
        assert res[10] == 1.0 and res[11] == 1.0 and res[12] == 0.0 and res[13] == 0.0, res
    return res


def test_exchange_path_single_rank_group(pct, tmp_path):
    """curvature_knn_exchange / curvature_knn_shared on a one-rank NCCL group (any box)."""
    _run_exchange(tmp_path, 1, 150_000, 20)


@pytest.mark.parametrize("k", [20, 32])
def test_exchange_path_two_ranks_nccl(pct, tmp_path, k):
    """Two real NCCL ranks on two GPUs: bins, all-to-all, slab margins, return order, unresolved redo."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = _run_exchange(tmp_path, 2, 400_000, k)
    assert res[8] >= 1                                    # the isolated point went through the whole-cloud redo


def _fuzz_cloud(rng, case):
    """Random small clouds with the traits that stress the search: scale, offset, anisotropy, clusters, duplicates."""
    n = int(rng.integers(40, 3500))
    kind = case % 6
    if kind == 0:      # volumetric blob
        p = rng.normal(size=(n, 3))
    elif kind == 1:    # noisy sheet
        p = np.stack((rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.02 * rng.normal(size=n)), 1)
    elif kind == 2:    # thin filament: cells mostly empty, long search radii
        t = rng.uniform(0, 6, n)
        p = np.stack((np.cos(t), np.sin(t), 0.3 * t), 1) + 0.01 * rng.normal(size=(n, 3))
    elif kind == 3:    # clusters of very different density
        c = rng.normal(size=(5, 3)) * 3
        s = np.array([0.01, 0.05, 0.2, 0.5, 1.0])
        w = rng.integers(0, 5, n)
        p = c[w] + rng.normal(size=(n, 3)) * s[w, None]
    elif kind == 4:    # surface of a sphere with a few exact duplicates
        v = rng.normal(size=(n, 3))
        p = v / np.linalg.norm(v, axis=1, keepdims=True)
        p[rng.integers(0, n, n // 20)] = p[rng.integers(0, n, n // 20)]
    else:              # jittered lattice: many near-ties
        g = int(round(n ** (1 / 3))) + 1
        m = np.stack(np.meshgrid(*(np.arange(g),) * 3, indexing="ij"), -1).reshape(-1, 3)[:n].astype(np.float64)
        p = m + 1e-4 * rng.normal(size=m.shape)
    scale = 10.0 ** rng.integers(-4, 5)
    offset = rng.normal(size=3) * scale * 10.0 ** rng.integers(0, 3)
    return (p * scale + offset).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_clouds_against_the_oracle(pct, seed):
    """Randomised clouds (sizes 40 .. 3500, k 1 .. 64, scales 1e-4 .. 1e4, offsets up to 1000 extents, sheets, filaments,
    clusters, duplicates, jittered lattices): neighbour rows and distances bit-exact, curvature within the tolerance,
    through both the list kernels and the fused kernel."""
    rng = np.random.default_rng(1000 + seed)
    for case in range(12):
        pts = _fuzz_cloud(rng, case)
        n = len(pts)
        k = int(min(n - 2, rng.choice([1, 3, 7, 12, 20, 33, 64])))
        ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
        pc = pct.PointCloud(points=pts, normals=_empty_normals(n), k_neighbors=k)
        pc.plant_kdtree(k)
        tag = (seed, case, n, k)
        assert compare.neighbor_rows_differing(pc.neighbor_indices, ref_idx) == 0, tag
        assert np.array_equal(pc.dists, ref_dist), tag
        if k >= 6:     # (fewer rows than coefficients: lstsq's minimum-norm answer, covered by the degenerate goldens)
            K, H = pc.compute_pointwise_explicit_quadratic_curvature()
            ref = oracle.knn_curvature(pts, k)
            ok = np.isfinite(ref["K"]) & (np.asarray(pc.fit_status) & 4 == 0)      # rank-deficient rows: see test_tiny_and_degenerate_clouds
            # The reference hands the UNSCALED design [a^2, b^2, ab, a, b, 1] to lstsq (ref :358-359): its condition
            # number grows like sqrt(k) / r_k^2, so for neighbourhood radii below ~1e-4 (absolute units) the reference's
            # own coefficients carry a relative error of ~2e3 * eps / r_k^2 -- 0.1 % at r_k = 7e-7 -- before they drop to
            # zero altogether near 1e-7 (rcond).  The scaled solve here does not; such rows get the reference's noise as
            # extra tolerance (observed on the harness: (|dK| / |K|) * r_k^2 <= 4.8e-13).
            r_k = ref["dist"][:, -1].astype(np.float64)
            tiny = r_k < 1e-4
            rep = compare.curvature_report(_gpu_dict(pc), ref, r_k, rows_ok=ok & ~tiny)
            assert rep["violations"] == 0, (tag, rep)
            if (ok & tiny).any():
                m = ok & tiny
                rel = 1e-3 + 2e-12 / r_k[m] ** 2
                Kr, Hr = ref["K"][m].astype(np.float64), ref["H"][m].astype(np.float64)
                assert np.all(np.abs(np.asarray(pc.K_quadratic, np.float64)[m] - Kr) <= rel * np.abs(Kr) + 1e-5 / r_k[m] ** 2), tag
                assert np.all(np.abs(np.abs(np.asarray(pc.H_quadratic, np.float64)[m]) - np.abs(Hr)) <= rel * np.abs(Hr) + 1e-5 / r_k[m]), tag


@pytest.mark.parametrize("seed", [0, 1])
def test_random_clouds_ball_against_scipy(pct, seed):
    """The same randomised clouds through the epsilon-ball entry: the CSR rows (members ordered by (d, index), the radius
    test inclusive in fp64) equal scipy's query_ball_point bit for bit, empty and huge balls included; the fused ball fit
    stays inside the tolerance wherever a ball holds enough members."""
    rng = np.random.default_rng(2000 + seed)
    for case in range(12):
        pts = _fuzz_cloud(rng, case)
        n = len(pts)
        _, d1, _ = oracle.knn_canonical(pts, 8)
        radius = float(np.quantile(d1[:, -1], rng.choice([0.05, 0.5, 0.95])) * rng.choice([0.5, 1.0, 2.0]))
        if not radius > 0:
            continue
        tag = (seed, case, n, radius)
        pc = pct.PointCloud(points=pts, normals=_empty_normals(n))
        pc.plant_ball(radius)
        off, idx, dist = pc.ball_neighbors()
        roff, ridx, rdist = oracle.ball_canonical(pts, radius)
        assert compare.csr_equal(off, idx, roff, ridx), tag
        assert np.array_equal(dist, rdist), tag
        K, H = pc.compute_pointwise_explicit_quadratic_curvature()
        ref = oracle.curvature_from_csr(pts, roff, ridx)
        enough = (np.diff(roff) >= 10) & (np.asarray(pc.fit_status) & 4 == 0) & np.isfinite(ref["K"])
        if radius >= 1e-4 and enough.any():            # (below: the reference's own lstsq noise, see the kNN test above)
            rep = compare.curvature_report(_gpu_dict(pc), ref, np.full(n, radius), rows_ok=enough)
            assert rep["violations"] == 0, (tag, rep)


def test_tiny_and_degenerate_clouds(pct):
    """Edge cases of the staged kernel: clouds smaller than a chunk, k = 1, k = N - 1, coincident points,
    points on a line and on a plane.  Neighbour rows stay bit-exact; fits of rank-deficient neighbourhoods are lstsq's
    minimum-norm solution (ref :359), finite like the reference's, with the status bit as information."""
    rng = np.random.default_rng(12)
    cases = {
        "seven": (rng.normal(size=(7, 3)).astype(np.float32), [1, 5, 6]),
        "chunk_edge": (rng.normal(size=(257, 3)).astype(np.float32), [1, 20]),
        "line": (np.stack((np.linspace(0, 1, 400), np.zeros(400), np.zeros(400)), 1).astype(np.float32), [8]),
        "plane": (np.concatenate((rng.uniform(size=(3000, 2)), np.zeros((3000, 1))), 1).astype(np.float32), [20]),
        "coincident": (np.repeat(rng.normal(size=(5, 3)).astype(np.float32), 40, axis=0), [10]),
    }
    for name, (pts, ks) in cases.items():
        for k in ks:
            ref_idx, ref_dist, _ = oracle.knn_canonical(pts, k)
            pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
            pc.plant_kdtree(k)
            assert np.array_equal(pc.neighbor_indices, ref_idx), (name, k)
            assert np.array_equal(pc.dists, ref_dist), (name, k)
            pc2 = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
            pc2.plant_kdtree(k)
            K, H = pc2.compute_pointwise_explicit_quadratic_curvature()     # fused kernel, never raises on these
            assert K.shape == (len(pts),) and H.shape == (len(pts),)
            if name == "plane":
                ok = np.isfinite(K)
                assert ok.mean() > 0.99 and np.abs(K[ok]).max() < 1e-3 and np.abs(H[ok]).max() < 1e-2
            if name in ("line", "coincident"):
                # rank-deficient designs with z = 0: the minimum-norm solution is w = 0, like the reference's
                assert np.isfinite(K).all() and np.abs(K).max() < 1e-6 and np.abs(H).max() < 1e-4, (name, np.abs(K).max(), np.abs(H).max())
                assert ((pc2.fit_status & pct._lib.STATUS_RANK_DEFICIENT) != 0).mean() > 0.9
                assert int(pc2.kdtree.index.last_stats().rank_deficient) > 0
    with pytest.raises(IndexError):
        pc = pct.PointCloud(points=cases["seven"][0], normals=_empty_normals(7), k_neighbors=7)
        pc.plant_kdtree(7)


def test_degenerate_goldens_of_the_reference(pct):
    """tests/golden/degenerate.npz: what the UNMODIFIED reference returns for collinear, coincident, planar-lattice and
    two-parallel-lines clouds (oracle/make_golden_degenerate.py) -- finite numbers from lstsq's minimum-norm solution.
    Fused path (own neighbours; the clouds are full of ties, so rows differ from scipy's traversal order and only
    tie-free quantities are compared) and rows path (the reference's own rows: coefficient by coefficient)."""
    g = load_golden("degenerate")
    for name in ("line", "coincident", "plane_grid", "two_lines"):
        pts, k = np.ascontiguousarray(g[name + "_points"]), int(g[name + "_k"])
        pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
        pc.plant_kdtree(k)
        K, H = pc.compute_pointwise_explicit_quadratic_curvature()
        assert np.isfinite(K).all() and np.isfinite(H).all(), name
        if name != "two_lines":   # z = 0 in the tangent frame whatever the tie order: all zeros
            assert np.abs(K).max() < 1e-6 and np.abs(H).max() < 1e-4, (name, np.abs(K).max(), np.abs(H).max())
        # the reference's rows through the rows path
        pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
        pc.neighbor_indices = g[name + "_neighbor_indices"]
        pc.fit_explicit_quadratic_surfaces_to_neighborhoods()
        K, H = pc.calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
        assert np.allclose(K, g[name + "_K"], rtol=1e-4, atol=1e-7), (name, np.abs(K - g[name + "_K"]).max())
        assert np.allclose(H, g[name + "_H"], rtol=1e-4, atol=1e-6), (name, np.abs(H - g[name + "_H"]).max())
        scale = np.maximum(np.abs(g[name + "_coeffs"]).max(axis=1, keepdims=True), 1e-6)
        assert (np.abs(np.asarray(pc.quadratic_coefficients) - g[name + "_coeffs"]) / scale).max() < 1e-3, name


def test_kdtree_query_is_scipy_query(pct, bunny):
    """self.kdtree.query(x, k) (ref :83, :759) for cloud points AND arbitrary coordinates, against scipy itself."""
    from scipy.spatial import cKDTree

    pts = np.ascontiguousarray(bunny[::2])
    tree = cKDTree(pts)
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=20)
    pc.plant_kdtree(20)
    rng = np.random.default_rng(1)
    members = pts[rng.integers(0, len(pts), 200)]
    lo, hi = pts.min(0), pts.max(0)
    inside = (lo + rng.uniform(size=(200, 3)) * (hi - lo)).astype(np.float32)
    outside = (hi + rng.uniform(0.01, 0.3, size=(20, 3)) * (hi - lo)).astype(np.float32)
    for name, x in (("members", members), ("inside", inside), ("outside", outside)):
        for k in (1, 21, 101):
            d_ref, i_ref = tree.query(x, k)
            d, i = pc.kdtree.query(x, k)
            assert d.dtype == np.float64 and i.dtype == np.int64 and d.shape == d_ref.shape and i.shape == i_ref.shape
            assert np.array_equal(i, i_ref), (name, k)               # (bunny has no exact distance ties)
            assert np.allclose(d, d_ref, rtol=1e-12, atol=0), (name, k)
    # one point in, one row out; the point itself comes first (what ref :83-85 drops)
    d, i = pc.kdtree.query(pts[17], 21)
    assert d.shape == (21,) and i[0] == 17 and d[0] == 0.0
    assert np.array_equal(i[1:], pc.neighbor_indices[17])
    d1, i1 = pc.kdtree.query(pts[17])
    assert np.ndim(d1) == 0 and int(i1) == 17
    with pytest.raises(IndexError):
        pc.kdtree.query(pts[0], len(pts) + 1)


def test_assigned_rows_follow_numpy_indexing(pct, bunny):
    """user-assigned neighbor_indices -- negative indices count from the end, rows beyond the cloud are
    ignored, too few rows or an index outside [-N, N) raise IndexError like ref :638-640; nothing reaches the kernel
    unchecked, and the C entry flags bad rows instead of dereferencing them."""
    pts = np.ascontiguousarray(bunny[::7])
    n, k = len(pts), 12
    idx, _, _ = oracle.knn_canonical(pts, k)
    a = pct.PointCloud(points=pts, normals=_empty_normals(n), k_neighbors=k)
    a.neighbor_indices = idx
    a.fit_explicit_quadratic_surfaces_to_neighborhoods()
    want = np.asarray(a.quadratic_coefficients).copy()
    neg = idx.astype(np.int64).copy()
    neg[::2] -= n                                                       # the same points, addressed from the end
    b = pct.PointCloud(points=pts, normals=_empty_normals(n), k_neighbors=k)
    b.neighbor_indices = np.concatenate((neg, neg[:5]))                 # extra rows: the reference never reads them
    b.fit_explicit_quadratic_surfaces_to_neighborhoods()
    assert np.array_equal(np.asarray(b.quadratic_coefficients), want, equal_nan=True)
    for bad in (idx[: n - 3], np.where(idx == idx[0, 0], n, idx), np.where(idx == idx[0, 0], -n - 1, idx)):
        c = pct.PointCloud(points=pts, normals=_empty_normals(n), k_neighbors=k)
        c.neighbor_indices = bad
        with pytest.raises(IndexError):
            c.fit_explicit_quadratic_surfaces_to_neighborhoods()
    # the C ABI itself: a bad row comes back NaN with PCT_STATUS_BAD_INDEX
    from point_cloud_toolbox_b200 import engine

    d_pts = torch.from_numpy(pts).cuda()
    rows = torch.from_numpy(idx[:4].copy()).cuda()
    rows[1, 3] = n + 5
    rows[2, 0] = -n - 9
    fit = engine.fit_from_neighbors(d_pts, rows, query_ids=torch.arange(4, dtype=torch.int32, device="cuda"))
    st = fit.status.cpu().numpy()
    assert (st[[1, 2]] & pct._lib.STATUS_BAD_INDEX).all() and not (st[[0, 3]] & pct._lib.STATUS_BAD_INDEX).any()
    assert torch.isnan(fit.curv[[1, 2]]).all() and torch.isfinite(fit.curv[[0, 3]]).all()
    # in-place edits of the stored coefficients show in K and H (ref :663-672 reads them on every call)
    K0, H0 = a.calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
    a.quadratic_coefficients[5] = 0
    K1, H1 = a.calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
    assert K1[5] == 0 and H1[5] == 0 and np.array_equal(np.delete(K1, 5), np.delete(K0, 5), equal_nan=True)


def test_host_entry_does_not_strand_scratch_memory(pct):
    """pct_curvature_knn_host runs on a private stream; its scratch arena goes with the stream."""
    import ctypes

    pts, _, _ = datasets.torus_random(300_000, seed=2)
    n = len(pts)
    K = np.empty(n, np.float32)
    H = np.empty(n, np.float32)
    lib = pct._lib.lib
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):
        pct._lib.check(lib.pct_curvature_knn_host(pts.ctypes.data_as(ctypes.c_void_p), n, 20, K.ctypes.data_as(ctypes.c_void_p),
                                                  H.ctypes.data_as(ctypes.c_void_p)))
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 << 20, (free0 - free1) >> 20      # (six stranded arenas would be > 200 MB)
    assert np.isfinite(K).mean() > 0.999


def test_implicit_quadric(pct):
    """Implicit 10-coefficient quadric (ref :363-396, :435-480, :617-633, :676-689): the curvature formulas are pinned
    to the unmodified reference (tests/golden/implicit.npz), the fit is the minimiser of the reference's own problem
    (unpinned: SLSQP does not reach it, tests/test_oracle.py) and recovers a sphere exactly."""
    g = load_golden("implicit")
    for c, want in zip(g["coeffs"][:64], g["curv"][:64]):
        got = np.array(pct.PointCloud.calculate_implicit_quadric_curvatures(c))
        assert np.allclose(got, want, rtol=1e-11, atol=0, equal_nan=True), (got, want)
    from point_cloud_toolbox_b200 import engine

    got = engine.implicit_quadric_curvature(torch.from_numpy(g["coeffs"]).cuda()).cpu().numpy()
    assert np.allclose(got, g["curv"], rtol=1e-11, atol=0, equal_nan=True)
    # the fit minimises |A c|^2 on the unit sphere: compare with the spectrum numpy found for the stored neighbourhoods
    for pts, w, obj_slsqp in zip(g["neighbourhoods"], g["eigenvalues"], g["slsqp_objective"]):
        c = pct.PointCloud.fit_implicit_quadric_surface(pts)
        p = pts.astype(np.float32)
        A = np.column_stack((p[:, 0] ** 2, p[:, 1] ** 2, p[:, 2] ** 2, p[:, 0] * p[:, 1], p[:, 0] * p[:, 2], p[:, 1] * p[:, 2],
                             p[:, 0], p[:, 1], p[:, 2], np.ones(len(p))))
        obj = float(np.sum((A @ c) ** 2))
        assert abs(np.linalg.norm(c) - 1) < 1e-12
        assert obj <= 2 * w[0] + 1e-17 and obj < 1e-6 * obj_slsqp, (obj, w[0], obj_slsqp)
    # a sphere is an exact quadric: x^2 + y^2 + z^2 - 2 R n.x = 0 in coordinates centred on a surface point
    R = 2.0
    sph, _, _ = datasets.sphere_fibonacci(20000, radius=R)
    pc = pct.PointCloud(points=sph, normals=_empty_normals(len(sph)), k_neighbors=30)
    pc.plant_kdtree(30)
    pc.K_quadratic, pc.H_quadratic = None, None                  # (ref :633 returns these attributes)
    K, H = pc.compute_pointwise_implicit_quadric_curvature()
    coef = np.asarray(pc.quadric_coefficients)
    assert coef.shape == (len(sph), 10) and np.allclose(np.linalg.norm(coef, axis=1), 1, atol=1e-12)
    # the definition, row by row: |A c| equals the smallest singular value of the neighbourhood's own design matrix
    # (ref :617-633 fits the point and its k - 1 nearest, centred on the point)
    nb = np.concatenate((np.arange(len(sph))[:, None], np.asarray(pc.neighbor_indices)[:, :29]), axis=1)
    P = (sph[nb] - sph[:, None, :]).astype(np.float32)
    x, y, z = P[..., 0], P[..., 1], P[..., 2]
    A = np.stack((x * x, y * y, z * z, x * y, x * z, y * z, x, y, z, np.ones_like(x)), -1).astype(np.float64)
    sv = np.linalg.svd(A, compute_uv=False)
    res = np.linalg.norm(np.einsum("nkj,nj->nk", A, coef), axis=1)
    assert np.all(res <= sv[:, -1] * (1 + 1e-6) + 1e-13 * sv[:, 0]), float((res / sv[:, -1]).max())
    # geometry, as far as fp32 coordinates condition it (the second smallest singular value is only 100 times the
    # smallest: numpy's own SVD has B / A - 1 up to 3.4e-2 on these rows, median 2e-3)
    s = coef[:, 0]
    assert np.median(np.abs(coef[:, 1] / s - 1)) < 5e-3 and np.abs(coef[:, 1] / s - 1).max() < 0.1       # A = B = C
    assert np.median(np.abs(coef[:, 2] / s - 1)) < 5e-3 and np.abs(coef[:, 2] / s - 1).max() < 0.1
    assert np.abs(coef[:, 3:6]).max() < 0.1 * np.abs(s).min()                                     # no mixed terms
    n_out = sph / R                                                                               # outward normal
    assert np.allclose(coef[:, 6:9], 2 * R * s[:, None] * n_out, rtol=0.1, atol=0.05)           # gradient along the normal, away from the neighbours
    assert np.all(np.einsum("ij,ij->i", coef[:, 6:9], n_out) * np.sign(s) > 0)
    assert np.allclose(np.abs(H), 1 / R, rtol=2e-4)                                               # K_h = -+ 1 / R, scale invariant
    assert np.allclose(K, 8 * s ** 3 / (2 * R * np.abs(s)) ** 4, rtol=0.15)                       # the reference's K_g = det(Hess) / |g|^4 (not scale invariant)


def test_estimate_curvature_reference_compatible(pct, bunny):
    """utils.estimate_curvature(reference_compatible=True) computes what utils.py:778-829 computes as written: the
    smallest eigenvalue of the k x k Gram matrix over the eigenvalue sum -- rounding noise around zero."""
    from point_cloud_toolbox_b200 import utils as U

    pts = np.ascontiguousarray(bunny[::12])
    got = U.estimate_curvature(pts, k_fraction=0.004, reference_compatible=True)
    n = len(pts)
    k = min(max(5, int(0.004 * n)), 100)
    from scipy.spatial import cKDTree

    _, idx = cKDTree(pts).query(pts, k)                                  # sklearn's kneighbors(points): self first
    nb = pts[idx]
    c = nb - nb.mean(axis=1, keepdims=True)
    cov = np.einsum("nik,njk->nij", c, c) / (k - 1)                      # utils.py:822
    ev = np.linalg.eigh(cov)[0]
    want = ev[:, 0] / (ev.sum(axis=1) + 1e-10)                           # utils.py:827-828
    assert got.shape == want.shape and np.abs(want).max() < 1e-5
    assert np.abs(got - want).max() < 1e-5 and np.abs(got).max() < 1e-5  # noise on both sides, nothing else
    assert U.estimate_curvature(pts, k_fraction=0.004).min() > 1e-6       # the documented quantity is not noise


def test_errors_mirror_reference(pct):
    with pytest.raises(ValueError, match="Either file_path or points and normals"):
        pct.PointCloud()
    pts = np.random.default_rng(0).normal(size=(50, 3)).astype(np.float32)
    pc = pct.PointCloud(points=pts, normals=_empty_normals(50))
    with pytest.raises(IndexError):
        pc.plant_kdtree(50)
    bad = pts.copy()
    bad[3, 1] = np.nan
    pc = pct.PointCloud(points=bad, normals=_empty_normals(50))
    with pytest.raises(ValueError, match="Non-finite"):
        pc.plant_kdtree(5)


def test_file_loader(pct, tmp_path):
    g = load_golden("loader_case")
    path = tmp_path / "cloud.txt"
    np.savetxt(path, g["table"], fmt="%.5f")
    pc = pct.PointCloud(str(path), k_neighbors=5)
    assert pc.points.dtype == np.float32 and np.array_equal(pc.points, g["points"])
    assert np.array_equal(pc.normals, g["normals"])
    assert np.allclose(pc.x_domain, g["x_domain"]) and np.allclose(pc.z_domain, g["z_domain"])
    assert np.isclose(pc.l1_norm, g["l1_norm"], rtol=1e-5)
    assert np.isclose(pc.l2_norm, g["l2_norm"], rtol=1e-5)
    assert np.isclose(pc.infinity_norm, g["infinity_norm"], rtol=1e-5)
    assert pc.num_points == 50 and pc.num_features == 3


def test_slice_layout_and_ranges(pct, bunny):
    from point_cloud_toolbox_b200._lib import LAYOUT_SLICE

    d = torch.from_numpy(bunny).cuda()
    ix = pct.GridIndex(d, k_hint=20)
    full = ix.curvature_knn(20)
    perm = ix.permutation().long()
    n = len(bunny)
    a, b = n // 3, n // 3 + 5001
    part = ix.curvature_knn(20, a, b, layout=LAYOUT_SLICE)
    assert part.curv.shape == (b - a, 5)
    assert torch.equal(part.curv, full.curv[perm[a:b]])
    assert torch.equal(part.normals, full.normals[perm[a:b]])
    idx_s, dist_s = ix.knn(20, a, b, layout=LAYOUT_SLICE)
    idx_f, dist_f = ix.knn(20)
    assert torch.equal(idx_s, idx_f[perm[a:b]]) and torch.equal(dist_s, dist_f[perm[a:b]])
    assert sorted(perm.cpu().tolist()) == list(range(n))
    info = ix.info()
    assert info.num_points == n and info.cells_level0 > 0


@pytest.mark.parametrize("shape,k", [("sphere", 20), ("sphere", 50), ("torus", 20), ("egg", 50)])
def test_c3_analytic_surfaces_1m(pct, shape, k):
    n = 1_000_000
    if shape == "sphere":
        pts, Kt, Ht = datasets.sphere_fibonacci(n)
        mask = np.ones(n, bool)
    elif shape == "torus":
        pts, Kt, Ht = datasets.torus_random(n, seed=0)
        mask = np.ones(n, bool)
    else:
        pts, Kt, Ht = datasets.egg_carton_random(n, seed=1)
        mask = datasets.interior_mask_xy(pts, 2 * np.pi, 0.2)
    pc = pct.PointCloud(points=pts, normals=_empty_normals(n), k_neighbors=k)
    pc.plant_kdtree(k)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    cf = compare.closed_form_report(K, H, Kt, Ht, mask)
    # discretisation error of the estimator itself: h^2 scaling from 2e-3 at N = 2e4
    lim = {"sphere": (2e-3, 2e-3), "torus": (0.15, 0.08), "egg": (0.05, 0.03)}[shape]
    assert cf["K_abs_p99"] < lim[0] and cf["H_abs_p99"] < lim[1], cf
    rows = np.sort(np.random.default_rng(3).choice(n, 40_000, replace=False))
    ref = oracle.knn_curvature(pts, k, rows=rows)
    idx, dist = pc.kdtree.index.knn(k)
    sel = torch.from_numpy(rows).cuda()
    assert compare.neighbor_rows_differing(idx[sel].cpu().numpy(), ref["idx"]) == 0
    assert np.array_equal(dist[sel].cpu().numpy(), ref["dist"])
    got = dict(normal=pc.normals_quadratic[rows], K=K[rows], H=H[rows], k1=pc.k1_quadratic[rows], k2=pc.k2_quadratic[rows])
    rep = compare.curvature_report(got, ref, ref["dist"][:, -1])
    assert rep["violations"] == 0, rep
    assert rep["tight_fraction"] > 0.995, rep


def test_c4_scanned_sheet_ball(pct):
    pts, _, _ = datasets.scanned_sheet()
    n = len(pts)
    tree_idx, tree_dist, _ = oracle.knn_canonical(pts, 1, rows=np.arange(0, n, 50))
    radius = float(2.5 * np.median(tree_dist[:, 0]))
    pc = pct.PointCloud(points=pts, normals=_empty_normals(n))
    pc.plant_ball(radius)
    counts = pc.kdtree.index.ball_count(radius).cpu().numpy()
    rows = np.sort(np.random.default_rng(4).choice(n, 20_000, replace=False))
    roff, ridx, rdist = oracle.ball_canonical(pts, radius, rows=rows)
    assert np.array_equal(counts[rows], np.diff(roff))
    assert counts.max() > 10 * max(1, np.median(counts))  # the density really varies
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    ref = oracle.curvature_from_csr(pts, roff, ridx, rows=rows)
    got = dict(normal=pc.normals_quadratic[rows], K=K[rows], H=H[rows], k1=pc.k1_quadratic[rows], k2=pc.k2_quadratic[rows])
    rep = compare.curvature_report(got, ref, np.full(len(rows), radius), rows_ok=np.diff(roff) >= 10)
    assert rep["violations"] == 0, rep
    # also k = 100, the profiled configuration of the reference (profile_stats)
    pc.plant_kdtree(100)
    K100, H100 = pc.compute_pointwise_explicit_quadratic_curvature()
    ref100 = oracle.knn_curvature(pts, 100, rows=rows[:5000])
    got = dict(normal=pc.normals_quadratic[rows[:5000]], K=K100[rows[:5000]], H=H100[rows[:5000]],
               k1=pc.k1_quadratic[rows[:5000]], k2=pc.k2_quadratic[rows[:5000]])
    rep = compare.curvature_report(got, ref100, ref100["dist"][:, -1])
    assert rep["violations"] == 0, rep


def test_large_cloud_properties(pct):
    """C5-shaped input at a size the oracle cannot sweep: size-independent properties."""
    n = 8_000_000
    gen = torch.Generator(device="cuda").manual_seed(3)
    u = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * (2 * np.pi)
    v = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * (2 * np.pi)
    R, r = 1.0, 1.0 / 3.0
    pts = torch.stack(((R + r * torch.cos(v)) * torch.cos(u), (R + r * torch.cos(v)) * torch.sin(u), r * torch.sin(v)), 1).float().contiguous()
    ix = pct.GridIndex(pts, k_hint=32)
    out1 = ix.curvature_knn(32)
    out2 = ix.curvature_knn(32)
    assert torch.equal(out1.curv, out2.curv)  # deterministic / idempotent
    K = out1.curv[:, 0].double()
    Kt = torch.cos(v) / (r * (R + r * torch.cos(v)))
    err = (K - Kt).abs()
    assert float(err.median()) < 5e-3 and float(err.quantile(0.99)) < 0.05
    assert int((out1.status != 0).sum()) <= n // 1000
    k1, k2, H = out1.curv[:, 2], out1.curv[:, 3], out1.curv[:, 1]
    assert bool((k1 >= k2).all()) and torch.allclose(0.5 * (k1 + k2), H, rtol=1e-4, atol=1e-4)
    # neighbour rows of a slice: sorted distances, no self, symmetric closest pair
    idx, dist = ix.knn(32, 1000, 201000, layout=1)
    assert bool((dist[:, 1:] >= dist[:, :-1]).all())
    perm = ix.permutation().long()
    assert not bool((idx == perm[1000:201000, None]).any())
    # a sample against the oracle on a spatial crop (everything within the crop + margin)
    host = pts.cpu().numpy()
    crop = np.nonzero((np.abs(host[:, 0] - 1.2) < 0.05) & (np.abs(host[:, 1]) < 0.05))[0]
    inner = np.nonzero((np.abs(host[crop, 0] - 1.2) < 0.03) & (np.abs(host[crop, 1]) < 0.03))[0]
    ref_idx, ref_dist, _ = oracle.knn_canonical(host[crop], 32, rows=inner)
    gi, gd = ix.knn(32)
    sel = torch.from_numpy(crop[inner]).cuda()
    assert np.array_equal(crop[ref_idx], gi[sel].cpu().numpy())
    assert np.array_equal(ref_dist, gd[sel].cpu().numpy())


# ---------------------------------------------------------------------------
# either side of the path: energies, PCA estimators, file round trip
# ---------------------------------------------------------------------------
def test_mesh_energies_against_the_reference_numbers(pct):
    from oracle import around_path as ap
    from point_cloud_toolbox_b200 import utils as U

    g = load_golden("io_energy_pca")
    v, t, K, H = g["energy_vertices"], g["energy_triangles"], g["energy_K"], g["energy_H"]
    got = U.compute_energies(v, t, K, H)
    assert np.allclose(got, g["energy_result"], rtol=1e-12, atol=1e-13)          # fp64 sums in another order
    assert np.allclose(U.compute_energies(v, t), g["energy_result_no_curvature"], rtol=1e-12, atol=0)

    class Mesh:  # pyvista-like
        points = v
        faces = np.concatenate([np.full((len(t), 1), 3, np.int64), t.astype(np.int64)], 1).ravel()
        point_data = {"gaussian_curvature": K, "mean_curvature": H}

    assert np.allclose(U.load_mesh_compute_energies(Mesh), g["energy_result"], rtol=1e-12, atol=1e-13)
    assert U.compute_energies(v, np.zeros((0, 3), np.int32), K, H) == (0, 0, 0)
    assert U.compute_energies(v * 0, t, K, H) == (0, 0, 0)
    with pytest.raises(IndexError):
        U.compute_energies(v, np.array([[0, 1, len(v)]], np.int32), K, H)
    # a mesh of 2 M triangles (more than one wave of the persistent grid) with the path's own K and H
    n = 1001
    verts, tris = datasets.grid_mesh(n, 5)
    pc = pct.PointCloud(points=verts, normals=_empty_normals(len(verts)), k_neighbors=20)
    pc.plant_kdtree(20)
    Kq, Hq = pc.compute_pointwise_explicit_quadratic_curvature()
    got = U.compute_energies(verts, tris, Kq, Hq)
    want = ap.mesh_energies(verts, tris, Kq, Hq)
    assert np.allclose(got, want, rtol=1e-11, atol=0)
    # determinism: the two-stage reduction has a fixed order
    assert got == U.compute_energies(verts, tris, Kq, Hq)


def test_pca_principal_curvatures_against_the_reference(pct):
    from oracle import around_path as ap

    g = load_golden("io_energy_pca")
    pts, k = g["pca_points"], int(g["pca_k"])
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=k)
    pc.principal_curvatures_via_principal_component_analysis(k)
    l1, l2 = pc.pca_principal_curvature_values_1, pc.pca_principal_curvature_values_2
    assert l1.dtype == np.float64 and pc.principal_curvature_directions.shape == (len(pts), 3, 2)
    assert np.all(np.abs(l1 - g["pca_l1"]) <= 1e-12 * g["pca_l1"])
    assert np.all(np.abs(l2 - g["pca_l2"]) <= 1e-12 * g["pca_l1"])
    assert np.allclose(pc.pca_K_values, g["pca_K"], rtol=1e-9, atol=0)
    assert np.allclose(pc.pca_H_values, g["pca_H"], rtol=1e-12, atol=0)
    vals, _ = ap.pca_from_rows(pts, oracle.knn_canonical(pts, k)[0])
    gap = np.minimum(vals[:, 0] - vals[:, 1], vals[:, 1] - vals[:, 2]) / vals[:, 0]
    ok = gap > 1e-3
    dots = np.abs(np.einsum("nij,nij->nj", pc.principal_curvature_directions, g["pca_directions"]))
    assert np.all(dots[ok] > 1 - 1e-9)
    # with a planted tree of the same cloud the index is reused; other k
    pc.plant_kdtree(20)
    pc.principal_curvatures_via_principal_component_analysis(25)
    vals, _ = ap.pca_from_rows(pts, oracle.knn_canonical(pts, 25)[0])
    assert np.all(np.abs(pc.pca_principal_curvature_values_1 - vals[:, 0]) <= 1e-12 * vals[:, 0])
    assert np.allclose(pc.pca_H_values, vals[:, 4], rtol=1e-12, atol=0)


def test_surface_variation_estimate(pct):
    from oracle import around_path as ap
    from point_cloud_toolbox_b200 import utils as U

    pts = _cloud("bunny")[::4].copy()
    got = U.estimate_curvature(pts, k_fraction=0.002)           # k = max(5, 17) = 17, the point itself included
    k = min(max(5, int(0.002 * len(pts))), 100)
    want, _ = ap.pca_from_rows(pts, oracle.knn_canonical(pts, k - 1)[0], include_self=True)
    assert got.shape == (len(pts),) and np.allclose(got, want[:, 5], rtol=1e-7, atol=1e-12)
    assert 0 <= got.min() and got.max() <= 1 / 3 + 1e-12


def test_file_in_file_out_round_trip(pct, tmp_path):
    """Text scan in (library parser), curvature out (library writers), read back with the PLY reader."""
    from oracle import around_path as ap
    from point_cloud_toolbox_b200 import utils as U

    pts = _cloud("bunny")
    src = tmp_path / "scan.txt"
    np.savetxt(src, pts.astype(np.float64), fmt="%.9g")
    pc = pct.PointCloud(str(src), k_neighbors=20)
    shifted = pts.copy()
    shifted[:, 0] -= shifted[:, 0].max()
    shifted[:, 1] -= shifted[:, 1].max()
    assert np.array_equal(pc.points, shifted) and pc.normals.shape == (len(pts), 0)
    pc.plant_kdtree(20)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    out = tmp_path / "output_with_curvatures.ply"
    U.save_curvatures_to_ply(pc.points, K, H, str(out))
    assert out.read_bytes() == ap.curvature_ply_bytes(pc.points, K, H)
    back = U.parse_ply(str(out))
    assert np.array_equal(back, pc.points)
    fg, fm = U.save_curvature_arrays(K, H, "bunny", "scan", 1, str(tmp_path / "curvature_data"))
    assert np.array_equal(np.load(fg), K) and np.array_equal(np.load(fm), H)


def test_energy_and_pca_corner_cases(pct):
    from oracle import around_path as ap
    from point_cloud_toolbox_b200 import engine

    v = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3]], dtype=torch.float32, device="cuda")
    K = torch.tensor([1.0, 2.0, 4.0, float("nan")], device="cuda")
    H = torch.tensor([1.0, -2.0, 3.0, 5.0], device="cuda")
    # no triangles: zeros; one triangle; numpy-style negative indices; an index out of range is counted, not read
    assert engine.mesh_energies(v, torch.zeros((0, 3), dtype=torch.int32, device="cuda"), K, H).tolist() == [0, 0, 0, 0]
    t = torch.tensor([[0, 1, 2]], dtype=torch.int32, device="cuda")
    got = engine.mesh_energies(v, t, K, H).cpu().numpy()
    want = ap.mesh_energies(v.cpu().numpy(), t.cpu().numpy(), K.cpu().numpy(), H.cpu().numpy())
    assert got[3] == 0 and tuple(got[:3]) == tuple(float(x) for x in want)       # a single triangle is exact
    neg = torch.tensor([[0, 1, -2], [0, 1, -1], [0, 4, 1], [0, -5, 1]], dtype=torch.int32, device="cuda")
    got = engine.mesh_energies(v, neg, K, H).cpu().numpy()
    want = ap.mesh_energies(v.cpu().numpy(), np.array([[0, 1, 2], [0, 1, 3]]), K.cpu().numpy(), H.cpu().numpy())
    assert got[3] == 2 and np.allclose(got[:3], want, rtol=1e-15)                # the NaN corner drops out of stretching only
    assert np.isfinite(got[:3]).all()
    # PCA: one neighbour gives np.cov's NaN, n - 1 neighbours use every other point
    pts = _cloud("bunny")[:500].copy()
    pc = pct.PointCloud(points=pts, normals=_empty_normals(len(pts)), k_neighbors=5)
    with np.errstate(all="ignore"):
        pc.principal_curvatures_via_principal_component_analysis(1)
    assert np.isnan(pc.pca_K_values).all()
    small = pts[:60].copy()
    pc = pct.PointCloud(points=small, normals=_empty_normals(60), k_neighbors=5)
    pc.principal_curvatures_via_principal_component_analysis(59)
    idx = oracle.knn_canonical(small, 59)[0]
    vals, _ = ap.pca_from_rows(small, idx)
    assert np.all(np.abs(pc.pca_principal_curvature_values_1 - vals[:, 0]) <= 1e-12 * vals[:, 0])
    assert np.all(np.abs(pc.pca_principal_curvature_values_2 - vals[:, 1]) <= 1e-12 * vals[:, 0])


def test_pageable_upload_is_staged_and_exact(pct):
    from point_cloud_toolbox_b200 import engine

    rng = np.random.default_rng(12)
    a = rng.random((3_000_001, 3), dtype=np.float32)                 # 36 MB, pageable: goes through pct_upload
    d = engine.to_device_points(a)
    assert d.is_cuda and d.dtype == torch.float32 and torch.equal(d.cpu(), torch.from_numpy(a))
    d2 = engine.to_device_points(a)                                   # the staging buffers are reused
    assert torch.equal(d2, d)
    from point_cloud_toolbox_b200 import _lib

    torch.cuda.synchronize()
    assert _lib.lib.pct_release_scratch() == 0                        # frees the staging buffers too
    assert torch.equal(engine.to_device_points(a), d)                 # ... and they come back on demand
    pinned = torch.from_numpy(a).pin_memory()
    assert torch.equal(engine.to_device_points(pinned), d)            # page-locked source: direct copy
    assert torch.equal(engine.to_device_points(a[:1000]), d[:1000])   # small: direct copy
    six = np.concatenate([a[:50000], a[:50000]], 1)                   # (N, 6): the first three columns
    assert torch.equal(engine.to_device_points(six), d[:50000])


def test_host_buffer_entry_of_the_c_abi(pct):
    """pct_curvature_knn_host: plain host pointers in and out, what a binding without a device allocator calls."""
    import ctypes

    from point_cloud_toolbox_b200 import _lib

    for pts in (_cloud("bunny"), np.ascontiguousarray(np.tile(_cloud("bunny"), (80, 1)) +
                                                      np.repeat(np.arange(80, dtype=np.float32), 35947)[:, None])):
        # the second cloud (2.9 M points, 34 MB) is large enough for the staged upload
        n, k = len(pts), 20
        K = np.empty(n, np.float32)
        H = np.empty(n, np.float32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        _lib.check(_lib.lib.pct_curvature_knn_host(P(pts), n, k, P(K), P(H)))
        pc = pct.PointCloud(points=pts, normals=_empty_normals(n), k_neighbors=k)
        pc.plant_kdtree(k)
        K2, H2 = pc.compute_pointwise_explicit_quadratic_curvature()
        assert np.array_equal(K, K2, equal_nan=True) and np.array_equal(H, H2, equal_nan=True)
