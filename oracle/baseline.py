"""Timed CPU baseline: the reference's path as it runs, per point, on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (``bench.py``'s ``cpu_baseline`` leg and
``--impl reference`` arm).  The reference is single-threaded Python
(/root/reference/pointCloudToolbox.py:81-85, :638-647, :663-672): one
``kdtree.query(point, k+1)`` and one covariance / SVD / lstsq per point.  This
module runs exactly that loop (through the restated functions of
``oracle/reference_path.py``) and, because points are independent, fans the loop
out over host processes so the number is the best the reference's CPU path can
do on the box, not a single-core strawman.  It is a reported baseline, never the
thing the GPU results are checked with.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np
from scipy.spatial import cKDTree

from .reference_path import neighbourhood_pipeline

_STATE = {}


def _as_is_rows(points, tree, k, lo, hi):
    """ref :81-85 then :640-647 and :668 for rows [lo, hi)."""
    acc = 0.0
    for i in range(lo, hi):
        _, nb = tree.query(points[i], k + 1)          # ref :83
        nb = nb[1:]                                    # ref :85
        _, _, curv = neighbourhood_pipeline(points, i, nb)
        acc += float(curv[0])
    return acc


def _worker(args):
    lo, hi = args
    pts, k = _STATE["pts"], _STATE["k"]
    if "tree" not in _STATE:
        try:  # one BLAS thread per worker process: the workers are the parallelism
            from threadpoolctl import threadpool_limits

            _STATE["limits"] = threadpool_limits(1)
        except Exception:
            pass
        _STATE["tree"] = cKDTree(pts)                  # ref :74
    t = time.perf_counter()
    _as_is_rows(pts, _STATE["tree"], k, lo, hi)
    return time.perf_counter() - t


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def per_point_seconds(points, k, probe=1500):
    tree = cKDTree(points)
    n = min(probe, len(points))
    t = time.perf_counter()
    _as_is_rows(points, tree, k, 0, n)
    return (time.perf_counter() - t) / n


def timed_reference(points, k, seconds=15.0, procs=None, pool=None):
    """Run the as-is loop on a bounded number of rows of ``points`` with ``procs`` processes.

    Returns ``dict(points_per_s, rows, seconds, cores, per_point_us_single_core)``.
    """
    pts = np.ascontiguousarray(points, dtype=np.float32)
    procs = procs or host_cores()
    cost = per_point_seconds(pts, k)
    rows = int(min(len(pts), max(procs * 64, procs * seconds / cost)))
    chunk = max(16, rows // (procs * 4))
    tasks = [(lo, min(rows, lo + chunk)) for lo in range(0, rows, chunk)]
    _STATE.clear()
    _STATE.update(pts=pts, k=k)
    own_pool = pool is None
    if own_pool:
        pool = mp.get_context("fork").Pool(procs)
    try:
        pool.map(_worker, [(0, 8)] * procs)            # builds each worker's tree outside the timed part
        t = time.perf_counter()
        pool.map(_worker, tasks)
        wall = time.perf_counter() - t
    finally:
        if own_pool:
            pool.close()
            pool.join()
    return {
        "points_per_s": rows / wall,
        "rows": rows,
        "seconds": wall,
        "cores": procs,
        "per_point_us_single_core": cost * 1e6,
    }
