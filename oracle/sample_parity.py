"""Parity of a (possibly multi-GPU) run on a cloud too large for an oracle sweep: a seeded sample of queries is
checked against the oracle restricted to padded spatial crops of the WHOLE cloud.

TEST / BENCH INFRASTRUCTURE ONLY (``bench.py``'s parity block after the timed regions, ``tests/``).  The checker
is ``oracle.knn_curvature`` (ref :69-89, :505-509 restated, pinned to the unmodified reference by
``tests/test_oracle.py``); nothing here is on the product path.

Sample = runs of consecutive positions of the index's own Morton order (spatially compact), some of them
placed on the planes where a slab index stops owning points, so the margins of a partitioned run are covered.
For every run the crop holds every cloud point within ``pad`` of the run's bounding box; the oracle's rows are
exact for the run as long as its largest k-th distance stays below ``pad``, which is checked.
"""
from __future__ import annotations

import numpy as np

from . import compare
from .reference_path import knn_curvature


def choose_runs(index, n_runs, run_len, seed, planes=(), axis=0):
    """(begin, end) ranges of sorted positions: ``n_runs`` seeded starts plus one run around the indexed point
    nearest to each plane value in ``planes`` (coordinates on ``axis``)."""
    import torch

    n = int(index.n)
    run_len = int(min(run_len, n))
    rng = np.random.default_rng(seed)
    starts = [int(s) for s in rng.integers(0, max(1, n - run_len + 1), n_runs)]
    if len(planes):
        perm = index.permutation().long()                         # sorted position -> local id
        x_sorted = index.points[:, axis][perm]
        for v in planes:
            if not np.isfinite(v):
                continue
            pos = int(torch.argmin((x_sorted - float(v)).abs()).item())
            starts.append(max(0, min(n - run_len, pos - run_len // 2)))
    return [(s, s + run_len) for s in starts]


def check_runs(index, cloud, k, runs, *, local_to_orig=None, owned=None, records_of=None, host_kh=None, pad_cells=8.0):
    """Compare the index's answers on ``runs`` with the oracle.

    index          GridIndex (whole cloud or one slab)
    cloud          (N, 3) device tensor, the WHOLE cloud in original order (the crops are cut from it)
    local_to_orig  int tensor: original index of every point of the index's own cloud (None: identity)
    owned          callable(coords (m, 3) tensor) -> bool tensor: the queries this index answers (None: all)
    records_of     callable(orig ids int64 tensor) -> (m, 8) records tensor of the device-resident path (optional)
    host_kh        (K, H) host arrays in original order of the end-to-end path (optional)

    Returns a dict of counts: rows, rows_differing (neighbour rows), dist_differing, violations (curvature policy of
    oracle/compare.py on the records), e2e_violations (K, H of the host arrays), tight_fraction, pad_too_small.
    """
    import torch

    from point_cloud_toolbox_b200._lib import LAYOUT_SLICE

    info = index.info()
    pad0 = float(pad_cells) * float(info.cell_size)
    perm = index.permutation().long()
    tot = dict(rows=0, rows_differing=0, dist_differing=0, violations=0, e2e_violations=0, nan_rows=0, tight_rows=0,
               pad_too_small=0, runs=0, crop_points=0)
    for a, b in runs:
        local = perm[a:b]
        coords = index.points[local][:, :3]
        own = owned(coords) if owned is not None else torch.ones(len(local), dtype=torch.bool, device=coords.device)
        if not bool(own.any()):
            continue
        idx_l, dist = index.knn(k, a, b, layout=LAYOUT_SLICE)
        orig = local if local_to_orig is None else local_to_orig[local].long()
        nbr = idx_l[own].long()
        got_idx = (nbr if local_to_orig is None else local_to_orig[nbr].long()).cpu().numpy()
        got_dist = dist[own].cpu().numpy()
        q_orig = orig[own]
        q_xyz = coords[own]
        pad = pad0
        for _ in range(3):
            lo = q_xyz.min(0).values - pad
            hi = q_xyz.max(0).values + pad
            c3 = cloud[:, :3]
            mask = ((c3 >= lo) & (c3 <= hi)).all(1)
            crop_ids = mask.nonzero().squeeze(1)                    # ascending original indices: ties keep the cloud's order
            crop = c3[crop_ids].cpu().numpy()
            inner = torch.searchsorted(crop_ids, q_orig).cpu().numpy()
            ref = knn_curvature(crop, k, rows=inner)
            if float(ref["dist"][:, -1].max()) < pad:
                break
            pad *= 2.0
        else:
            tot["pad_too_small"] += 1
        ids_np = crop_ids.cpu().numpy()
        ref_idx = ids_np[ref["idx"]]
        tot["runs"] += 1
        tot["crop_points"] += int(len(ids_np))
        tot["rows"] += int(len(inner))
        tot["rows_differing"] += compare.neighbor_rows_differing(got_idx, ref_idx)
        tot["dist_differing"] += int(np.count_nonzero((got_dist != ref["dist"]).any(axis=1)))
        r_k = ref["dist"][:, -1]
        if records_of is not None:
            rec = records_of(q_orig).cpu().numpy()
            got = dict(normal=rec[:, 0:3], K=rec[:, 3], H=rec[:, 4], k1=rec[:, 5], k2=rec[:, 6])
            rep = compare.curvature_report(got, ref, r_k)
            tot["violations"] += rep["violations"]
            tot["nan_rows"] += rep["nan_rows"]
            tot["tight_rows"] += int(round(rep["tight_fraction"] * rep["rows"]))
        if host_kh is not None:
            q_np = q_orig.cpu().numpy()
            Kg, Hg = np.asarray(host_kh[0])[q_np].astype(np.float64), np.asarray(host_kh[1])[q_np].astype(np.float64)
            Kr, Hr = ref["K"].astype(np.float64), ref["H"].astype(np.float64)
            ok = np.isfinite(Kr) & np.isfinite(Hr)
            safe = ref["margin"] >= compare.MARGIN
            eK = np.abs(Kg - Kr) > compare.REL * np.abs(Kr) + compare.ABS_FLOOR / r_k ** 2
            eH = np.where(safe, np.abs(Hg - Hr), np.abs(np.abs(Hg) - np.abs(Hr))) > compare.REL * np.abs(Hr) + compare.ABS_FLOOR / r_k
            bad = (eK | eH | ~np.isfinite(Kg) | ~np.isfinite(Hg)) & ok
            tot["e2e_violations"] += int(bad.sum())
    return tot


def merge(parts):
    out = {}
    for p in parts:
        for key, v in p.items():
            out[key] = out.get(key, 0) + v
    if out.get("rows"):
        out["tight_fraction"] = out.pop("tight_rows", 0) / out["rows"]
    else:
        out.pop("tight_rows", None)
    return out
