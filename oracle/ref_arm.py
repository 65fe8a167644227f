"""The UNMODIFIED reference as a timed CPU arm (``bench.py --impl reference`` and the ``cpu_baseline`` leg).

TEST / BENCH INFRASTRUCTURE ONLY.  ``baseline/_ref/`` holds byte-for-byte copies of
``/root/reference/pointCloudToolbox.py`` and ``sample_scans/{bunny,egg_carton}.txt`` (``install()`` below,
called from ``__graft_entry__.build()`` in the dev container; the directory is git-ignored but travels to the
GPU box with the snapshot).  The file is imported as it is; only its top-level imports that the curvature path
never touches and that are not installed here are satisfied with empty modules (SURVEY.md appendix C:
matplotlib, pymesh, pyvista, memory_profiler -- pointCloudToolbox.py:7, :11, :16-17, :22).

What is timed is the reference's own public call sequence on one thread (the reference has no parallelism):

    pc = PointCloud(points=P, normals=empty, k_neighbors=k)      ref :26-47
    pc.plant_kdtree(k)                                             ref :69-89
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()     ref :505-509

on a self-contained sample cloud drawn from the workload's surface.  ``fan_out`` runs the same unmodified
sequence in one process per host core, each on its own sample cloud, to show what the box could do with the
reference as it is.
"""
from __future__ import annotations

import hashlib
import multiprocessing as mp
import os
import shutil
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
SOURCE = "/root/reference"
FILES = ("pointCloudToolbox.py", os.path.join("sample_scans", "bunny.txt"), os.path.join("sample_scans", "egg_carton.txt"))


def install(verbose=False):
    """Copy the reference files under baseline/_ref (dev container only: /root/reference is absent on the GPU box)."""
    if not os.path.isdir(SOURCE):
        return available()
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(REF_DIR, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
            if verbose:
                print("installed", dst)
    return available()


def available():
    return os.path.exists(os.path.join(REF_DIR, "pointCloudToolbox.py"))


def sha256():
    with open(os.path.join(REF_DIR, "pointCloudToolbox.py"), "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def import_reference():
    """The reference module, unmodified, from baseline/_ref."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "pymesh", "pyvista", "memory_profiler"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.patches"].Patch = object
    sys.modules["memory_profiler"].profile = lambda f: f
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import pointCloudToolbox  # noqa: E402
    assert os.path.dirname(os.path.abspath(pointCloudToolbox.__file__)) == REF_DIR, pointCloudToolbox.__file__
    return pointCloudToolbox


def run_reference(points, k):
    """The reference's public call sequence on ``points``; returns (K, H, seconds)."""
    ref = import_reference()
    pts = np.ascontiguousarray(points, dtype=np.float32)
    t = time.perf_counter()
    pc = ref.PointCloud(points=pts, normals=np.zeros((len(pts), 0), np.float32), k_neighbors=k)
    pc.plant_kdtree(k)
    K, H = pc.compute_pointwise_explicit_quadratic_curvature()
    return K, H, time.perf_counter() - t


def _silenced(fn, *args):
    """The reference prints progress lines; keep them out of bench.py's stdout (one JSON line)."""
    out = sys.stdout
    try:
        sys.stdout = open(os.devnull, "w")
        return fn(*args)
    finally:
        sys.stdout.close()
        sys.stdout = out


def timed_single(make_cloud, k, seconds):
    """One thread: a sample cloud sized so that the run takes about ``seconds``.
    ``make_cloud(n, seed)`` returns an (n, 3) float32 cloud of the workload's surface."""
    try:
        from threadpoolctl import threadpool_limits

        limits = threadpool_limits(1)
    except Exception:
        limits = None
    probe_n = 2000
    _, _, t_probe = _silenced(run_reference, make_cloud(probe_n, 11), k)
    n = int(max(probe_n, min(2_000_000, seconds / (t_probe / probe_n))))
    K, H, t = _silenced(run_reference, make_cloud(n, 3), k)
    del limits
    return {"points_per_s": n / t, "rows": n, "seconds": t, "cores": 1, "nan": int(np.count_nonzero(~np.isfinite(K))),
            "per_point_us": 1e6 * t / n}


def _fan_worker(args):
    make_cloud, n, seed, k = args
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    _, _, t = _silenced(run_reference, make_cloud(n, seed), k)
    return t


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def fan_out(make_cloud, k, seconds, per_point_s):
    """One process per host core, each running the unmodified sequence on its own sample cloud."""
    procs = host_cores()
    n = int(max(2000, seconds / per_point_s))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t = time.perf_counter()
        pool.map(_fan_worker, [(make_cloud, n, 100 + p, k) for p in range(procs)])
        wall = time.perf_counter() - t
    return {"points_per_s": procs * n / wall, "rows": procs * n, "seconds": wall, "cores": procs}
