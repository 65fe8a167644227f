"""Headline benchmark: points/s for kNN + PCA + quadric curvature on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--k 20] [--workload torus|c3_sphere|bunny_ball|sheet_ball]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the UNMODIFIED reference (baseline/_ref) on the host cores

A step is one pass of the hot path over one synthetic cloud:
    value  index build + fused kNN/fit kernel, the cloud already in HBM                 (device timed)
           N > 1: every rank holds 1/N of the cloud in HBM; the slab exchange (one all-to-all each way over
           NVLink) is INSIDE the timed region
    e2e    N = 1: PointCloud(points=host) -> plant_kdtree(k) -> compute_pointwise_explicit_quadratic_curvature()
           N > 1: distributed.curvature_knn_shared (every rank moves its share of the host arrays)
           pinned host input and host K, H output inside the timed region
Multi-GPU is strong scaling on one fixed cloud.  After the timed regions a seeded sample of queries of THIS run is
checked against the oracle (neighbour rows bit-exact, curvature within the stated tolerance) at k = 20 and k = 32:
the "parity" object.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points/sec for kNN+SVD+quadric curvature"
ALG_BYTES_QUERY = 44   # per point: 16 B own record read + 28 B result written; 32 B are actually written
ALG_BYTES_BALL = 48    # + 4 B member count
ALG_BYTES_MEMBER = 16  # ball workloads: one record per ball member has to be seen at least once


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="torus", choices=["torus", "c3_sphere", "bunny_ball", "sheet_ball", "c1_torus", "bunny_knn"])
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--k", type=int, default=None, help="neighbours (default 20; 30 for bunny_knn, BASELINE.json config 2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of each cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-k", default="20,32", help="k values of the parity block (torus workload)")
    ap.add_argument("--parity-rows", type=int, default=100_000, help="queries checked per k, over all ranks")
    ap.add_argument("--clock-interval-ms", type=int, default=200)
    args = ap.parse_args()
    if args.k is None:
        args.k = 30 if args.workload == "bunny_knn" else 20
    return args


# --------------------------------------------------------------------------
# workloads (the config object is identical in both arms)
# --------------------------------------------------------------------------
def workload_config(args):
    n, k = args.points, args.k
    if args.workload == "torus":
        return {"workload": f"synthetic torus surface (R=1, r=1/3, uniform u,v, seed 3), N={n}, k={k} kNN", "k": k, "points": n,
                "l2": "inputs (12 B/point raw + 16 B/point sorted: 2.8 GB at 100M) exceed the 126 MB L2; no flush needed"}
    if args.workload == "c3_sphere":
        return {"workload": f"C3: Fibonacci sphere R=1, N=1000000, k={k} kNN, closed-form K=1, |H|=1", "k": k, "points": 1_000_000,
                "l2": "28 MB of inputs fit the L2: a 256 MB buffer is written between timed steps"}
    if args.workload == "c1_torus":
        return {"workload": f"C1 stand-in: torus R=1, r=1/3 on the reference's 317 x 317 (theta, phi) grid, through %.6f text and the "
                            f"loader's fp32 max-shift, N=100489, k={k} kNN; CPU arm = the unmodified reference on the WHOLE cloud, one thread",
                "k": k, "points": 100_489, "l2": "inputs fit the L2: a 256 MB buffer is written between timed steps"}
    if args.workload == "bunny_knn":
        return {"workload": f"C2: sample_scans/bunny.txt (35947 points), k={k} kNN; CPU arm = the unmodified reference on the WHOLE cloud, one thread",
                "k": k, "points": 35947, "l2": "inputs fit the L2: a 256 MB buffer is written between timed steps"}
    if args.workload == "bunny_ball":
        return {"workload": "C2: sample_scans/bunny.txt (35947 points), epsilon-ball radius 3.8e-3 (mean ~30 members)", "k": 0,
                "points": 35947, "radius": 3.8e-3, "l2": "inputs fit the L2: a 256 MB buffer is written between timed steps"}
    return {"workload": "C4 stand-in: 332757-point scanned sheet z = sin x sin y, 3-component density mixture + N(0, 1e-3) noise, "
                        "epsilon-ball radius 2.5 x median nearest-neighbour spacing", "k": 0, "points": 332_757,
            "l2": "inputs fit the L2: a 256 MB buffer is written between timed steps"}


def host_sample(n_sample, seed=3):
    """Host cloud of the torus workload's distribution (same surface, same sampling law)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 2 * np.pi, n_sample)
    v = rng.uniform(0, 2 * np.pi, n_sample)
    R, r = 1.0, 1.0 / 3.0
    return np.stack(((R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)), 1).astype(np.float32)


def c1_torus_cloud(grid=317, R=1.0, r=1.0 / 3.0):
    """C1 stand-in: the reference's torus generator (utils.py:883-896) on a grid x grid lattice, written
    as 3-column %.6f text and read back the way PointCloud(file_path) conditions a scan (ref :51-57)."""
    import io

    import numpy as np

    t = np.linspace(0, 2 * np.pi, grid, endpoint=False)
    U, V = np.meshgrid(t, t)
    u, v = U.ravel(), V.ravel()
    p64 = np.stack(((R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)), 1)
    buf = io.StringIO()
    np.savetxt(buf, p64, fmt="%.6f")
    buf.seek(0)
    pts = np.loadtxt(buf)[:, 0:3].astype(np.float32)
    pts[:, 0] -= np.max(pts[:, 0])
    pts[:, 1] -= np.max(pts[:, 1])
    return np.ascontiguousarray(pts)


def sphere_sample(n_sample, seed=0):
    import numpy as np

    i = np.arange(0, n_sample, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n_sample)
    theta = np.pi * (1 + np.sqrt(5)) * i
    return np.stack((np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)), 1).astype(np.float32)


def sheet_cloud(n=332_757, seed=2, noise=1e-3, half_width=2 * math.pi):
    """C4 stand-in: non-uniform density so that ball sizes vary by more than 10x."""
    import numpy as np

    rng = np.random.default_rng(seed)
    centres = np.array([[-3.0, -2.0], [2.5, 1.0], [0.0, 4.0]])
    sigmas = np.array([0.8, 2.0, 4.0])
    weights = np.array([0.3, 0.4, 0.3])
    xy = np.empty((0, 2))
    while len(xy) < n:
        m = 2 * (n - len(xy)) + 1024
        comp = rng.choice(3, size=m, p=weights)
        cand = centres[comp] + rng.normal(size=(m, 2)) * sigmas[comp, None]
        xy = np.concatenate((xy, cand[(np.abs(cand) <= half_width).all(axis=1)]))
    xy = xy[:n]
    p = np.stack((xy[:, 0], xy[:, 1], np.sin(xy[:, 0]) * np.sin(xy[:, 1])), 1) + rng.normal(scale=noise, size=(n, 3))
    return p.astype(np.float32)


def small_workload_cloud(args):
    """(points float32 (N, 3), radius or None) of the non-default workloads."""
    import numpy as np

    if args.workload == "c3_sphere":
        return sphere_sample(1_000_000), None
    if args.workload == "c1_torus":
        return c1_torus_cloud(), None
    if args.workload in ("bunny_ball", "bunny_knn"):
        pts = np.load(os.path.join(ROOT, "tests", "golden", "bunny_points.npz"))["points"].astype(np.float32)
        return np.ascontiguousarray(pts), (3.8e-3 if args.workload == "bunny_ball" else None)
    pts = sheet_cloud()
    from scipy.spatial import cKDTree  # input preparation only: the radius is defined from the data's spacing

    sub = pts[:: max(1, len(pts) // 20000)]
    d1 = cKDTree(pts).query(sub, 2)[0][:, 1]
    return pts, float(2.5 * np.median(d1))


# --------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref), imported as it is
# --------------------------------------------------------------------------
def cpu_cloud_maker(args):
    if args.workload == "c3_sphere":
        return sphere_sample
    return host_sample


def reference_leg(args, seconds, with_fan_out=True):
    """cpu_baseline object: the reference's own call sequence on one thread (it has no parallelism), and the same
    unmodified sequence fanned out over every host core, each process on its own sample cloud."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    from oracle import ref_arm

    if args.workload in ("bunny_ball", "sheet_ball"):
        return ball_port_leg(args, seconds)
    if args.workload in ("c1_torus", "bunny_knn"):
        return reference_full_leg(args)[0]
    if not ref_arm.available():
        ref_arm.install()
    if not ref_arm.available():
        return {"unavailable": "baseline/_ref/pointCloudToolbox.py is missing (run __graft_entry__.build() in the dev container)"}
    k = args.k
    make = cpu_cloud_maker(args)
    one = ref_arm.timed_single(make, k, seconds)
    out = {"value": one["points_per_s"], "unit": "points/s", "cores": 1, "kind": "reference",
           "sample": (f"PointCloud(points) -> plant_kdtree({k}) -> compute_pointwise_explicit_quadratic_curvature() of the UNMODIFIED "
                      f"reference (baseline/_ref/pointCloudToolbox.py, sha256 {ref_arm.sha256()[:12]}) on one thread, a {one['rows']}-point "
                      f"cloud of the workload's surface, {one['seconds']:.1f} s"),
           "per_point_us_single_core": one["per_point_us"], "rows": one["rows"], "seconds": one["seconds"]}
    if with_fan_out:
        fan = ref_arm.fan_out(make, k, seconds, one["per_point_us"] * 1e-6)
        out["all_cores"] = {"value": fan["points_per_s"], "unit": "points/s", "cores": fan["cores"], "kind": "reference",
                            "sample": (f"the same unmodified call sequence in {fan['cores']} processes, each on its own "
                                       f"{fan['rows'] // fan['cores']}-point sample cloud, {fan['seconds']:.1f} s wall")}
    return out


def reference_full_leg(args):
    """c1_torus / bunny_knn: the unmodified reference on the WHOLE cloud, one thread (CPU baseline (1)).
    Returns the cpu_baseline object and the reference's own K, H for the parity figure of the run."""
    import numpy as np

    from oracle import ref_arm

    if not ref_arm.available():
        ref_arm.install()
    if not ref_arm.available():
        return {"unavailable": "baseline/_ref/pointCloudToolbox.py is missing (run __graft_entry__.build() in the dev container)"}, None
    try:
        from threadpoolctl import threadpool_limits

        limits = threadpool_limits(1)
    except Exception:
        limits = None
    pts, _ = small_workload_cloud(args)
    K, H, secs = ref_arm._silenced(ref_arm.run_reference, pts, args.k)
    del limits
    n = len(pts)
    return ({"value": n / secs, "unit": "points/s", "cores": 1, "kind": "reference", "rows": n, "seconds": secs,
             "per_point_us_single_core": 1e6 * secs / n,
             "sample": (f"the WHOLE cloud ({n} points): PointCloud(points) -> plant_kdtree({args.k}) -> "
                        f"compute_pointwise_explicit_quadratic_curvature() of the UNMODIFIED reference (baseline/_ref/pointCloudToolbox.py, "
                        f"sha256 {ref_arm.sha256()[:12]}) on one thread, {secs:.1f} s")},
            (np.asarray(K, np.float32), np.asarray(H, np.float32)))


def ball_port_leg(args, seconds):
    """The reference has no epsilon-ball code (README.md:8 only advertises it): the CPU figure is the oracle's
    composition of scipy's query_ball_point with the reference's per-neighbourhood functions, one thread."""
    import numpy as np

    import oracle

    pts, radius = small_workload_cloud(args)
    rows = np.arange(0, len(pts), max(1, len(pts) // 4000))
    t = time.perf_counter()
    off, idx, _ = oracle.ball_canonical(pts, radius, rows=rows)
    oracle.curvature_from_csr(pts, off, idx, rows=rows)
    dt = time.perf_counter() - t
    return {"value": len(rows) / dt, "unit": "points/s", "cores": 1, "kind": "port",
            "sample": f"{len(rows)} strided query points: scipy query_ball_point + the reference's plane / fit / curvature functions (oracle port), {dt:.1f} s"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step_seconds = max(2.0, min(15.0, 90.0 / max(1, args.steps + args.warmup)))
    legs = []
    full = args.workload in ("c1_torus", "bunny_knn")   # a step = the whole cloud (10-30 s): one untimed + two timed runs at most
    plan = ([0] * min(1, args.warmup) + [args.warmup] * min(2, args.steps)) if full else range(args.warmup + args.steps)
    for s in plan:
        leg = reference_leg(args, step_seconds, with_fan_out=False)
        if "unavailable" in leg:
            print(json.dumps({"impl": "reference", "unavailable": leg["unavailable"]}), flush=True)
            return
        if s >= args.warmup:
            legs.append(leg)
    if args.workload in ("bunny_ball", "sheet_ball"):
        value = sum(leg["value"] for leg in legs) / len(legs)
        ms = 0.0
    else:
        rows = sum(leg["rows"] for leg in legs)
        secs = sum(leg["seconds"] for leg in legs)
        value = rows / secs
        ms = 1e3 * secs / len(legs)
    cpu = dict(legs[-1])
    cpu["value"] = value
    if args.workload not in ("bunny_ball", "sheet_ball", "c1_torus", "bunny_knn"):
        from oracle import ref_arm

        fan = ref_arm.fan_out(cpu_cloud_maker(args), args.k, step_seconds, cpu["per_point_us_single_core"] * 1e-6)
        cpu["all_cores"] = {"value": fan["points_per_s"], "unit": "points/s", "cores": fan["cores"], "kind": "reference",
                            "sample": f"the same unmodified call sequence in {fan['cores']} processes, each on its own sample cloud"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 inputs, f64 LAPACK",
        "data": "synthetic" if not args.workload.startswith("bunny") else "sample_scans/bunny.txt (fixture copy)", "config": workload_config(args),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region.

    In-process NVML (nvidia_ml_py) from a thread: an `nvidia-smi -lms` child process was
    measured to slow the synchronisation-heavy index build by 2-4x on this pool (r01 notes
    in profiles/README.md); the NVML calls below do not.
    """

    def __init__(self, gpu_index, interval_ms=200, enabled=True):
        self.interval = interval_ms / 1e3
        self.enabled = enabled
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.error = None

    def prepare(self):
        """NVML initialisation in the calling thread, BEFORE any timed region: nvmlInit takes tens to
        hundreds of milliseconds and stalls concurrent CUDA calls of the process while it runs."""
        if not self.enabled:
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.gpu
            if visible:
                try:
                    index = int(visible.split(",")[self.gpu])
                except (ValueError, IndexError):
                    index = self.gpu
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM)
            self._sample()  # first call of every query outside the timed region too
            self.samples.clear()
        except Exception as exc:  # pragma: no cover
            self.error = repr(exc)
            self._nvml = None

    def _sample(self):
        pynvml, h = self._nvml, self._handle
        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
            else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        try:
            power = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
        except Exception:
            power = None
        self.samples.append((sm, self._max, int(reasons), power))

    def _run(self):
        try:
            while not self.stop_flag.is_set():
                self._sample()
                self.stop_flag.wait(self.interval)
        except Exception as exc:  # pragma: no cover
            self.error = repr(exc)

    def start(self):
        if not self.enabled or getattr(self, "_nvml", None) is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling disabled"]}
        self.stop_flag.set()
        self.thread.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {self.error}"]}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = set()
        for _, _, r, _ in self.samples:
            for bit, name in bits.items():
                if r & bit:
                    reasons.add(name)
        sm = sorted(x[0] for x in self.samples)
        power = [x[3] for x in self.samples if x[3] is not None]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": "NVML in-process"}


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_profile(k):
    """Counters of the committed ncu --set full capture of the dominant kernel (profiles/roofline_traffic.json):
    NOT measured in this run -- the entry names the capture (file, date, commit) it comes from."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
        return t if int(t.get("k", -1)) == k else None
    except Exception:
        return None


def fma_peaks():
    import ctypes

    from point_cloud_toolbox_b200 import _lib

    f32, f64 = ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.lib.pct_measure_fma_peaks(ctypes.byref(f32), ctypes.byref(f64), None))
    return f32.value, f64.value


def make_torus_on_device(n, dev):
    import torch

    gen = torch.Generator(device=dev).manual_seed(3)
    pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
    chunk = 1 << 24
    R, r = 1.0, 1.0 / 3.0
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * (2 * math.pi)
        v = torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * (2 * math.pi)
        w = R + r * torch.cos(v)
        pts[s:s + m, 0] = (w * torch.cos(u)).float()
        pts[s:s + m, 1] = (w * torch.sin(u)).float()
        pts[s:s + m, 2] = (r * torch.sin(v)).float()
        del u, v, w
    return pts


def parity_block(args, world, rank, dev, pts, shared_in, shared_out, host_pts):
    """Seeded sample of THIS run's answers against the oracle, per k: neighbour rows of the index that answered
    (bit-exact), records of the device-resident path and K, H of the end-to-end path (stated tolerance)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from oracle import sample_parity
    from point_cloud_toolbox_b200 import GridIndex, PointCloud
    from point_cloud_toolbox_b200 import distributed as pdist

    n = args.points
    out = {}
    run_len = 4096
    for kk in [int(v) for v in args.parity_k.split(",") if v]:
        n_runs = max(1, args.parity_rows // (run_len * world))
        if world == 1:
            index = GridIndex(pts, k_hint=kk)
            rec = index.curvature_knn(kk, want_coeffs=False).records
            pc = PointCloud(points=host_pts, normals=np.zeros((n, 0), np.float32), k_neighbors=kk)
            pc.plant_kdtree(kk)
            K, H = pc.compute_pointwise_explicit_quadratic_curvature()
            runs = sample_parity.choose_runs(index, n_runs, run_len, seed=1000 + kk)
            res = sample_parity.check_runs(index, pts, kk, runs, records_of=lambda ids: rec[ids], host_kh=(K, H))
            index.close()
            del pc, rec
            parts = [res]
        else:
            begin, end = pdist.shard_bounds(n, world, rank)
            part = pdist.curvature_knn_exchange(pts[begin:end].contiguous(), begin, n, kk)
            if shared_in is not None:
                pdist.curvature_knn_shared(shared_in, shared_out, kk, device=dev)       # collective; out complete on return
                host_kh = (shared_out.array[0], shared_out.array[1])
            else:
                host_kh = None
            res = {}
            if part.index is not None:
                c_lo, c_hi, own_lo, own_hi = part.plan.bounds[rank]
                axis = part.plan.axis
                runs = sample_parity.choose_runs(part.index, n_runs, run_len, seed=1000 + kk + 17 * rank,
                                                 planes=(own_lo, own_hi), axis=axis)
                own_ids = part.own_ids.long()
                res = sample_parity.check_runs(
                    part.index, pts, kk, runs, local_to_orig=part.local_ids,
                    owned=lambda c: (c[:, axis] >= own_lo) & (c[:, axis] < own_hi),
                    records_of=lambda ids: part.records[torch.searchsorted(own_ids, ids)], host_kh=host_kh)
            part.close()
            parts = [None] * world
            dist.all_gather_object(parts, res)
            dist.barrier()
        out[f"k{kk}"] = sample_parity.merge([p for p in parts if p])
    total = {key: sum(v.get(key, 0) for v in out.values()) for key in ("rows", "violations", "rows_differing", "dist_differing",
                                                                        "e2e_violations", "pad_too_small")}
    total["per_k"] = out
    total["checker"] = ("oracle.knn_curvature (ref :69-89, :505-509 restated, pinned to the unmodified reference) on padded spatial crops "
                        "of the whole cloud; sample = Morton runs of the index that answered, incl. runs on the slab cut planes")
    return total


def small_workload(args):
    """c3_sphere / bunny_ball / sheet_ball on one GPU: inputs fit the L2, so it is flushed between timed steps."""
    import numpy as np
    import torch

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    cpu, ref_kh = None, None
    if not args.no_cpu_baseline:
        if args.workload in ("c1_torus", "bunny_knn"):
            cpu, ref_kh = reference_full_leg(args)
        else:
            cpu = reference_leg(args, args.cpu_seconds, with_fan_out=False)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from point_cloud_toolbox_b200 import GridIndex, PointCloud

    host_pts, radius = small_workload_cloud(args)
    n, k = len(host_pts), args.k
    pinned = torch.from_numpy(host_pts).pin_memory()
    host_pts = pinned.numpy()
    pts = pinned.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(0, args.clock_interval_ms, not args.no_clocks)
    sampler.prepare()

    def device_step():
        flush.fill_(1)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if radius is None:
            index = GridIndex(pts, k_hint=k)
            e1.record()
            fit = index.curvature_knn(k, want_coeffs=False)
        else:
            index = GridIndex(pts, cell_hint=radius * 1.001)
            e1.record()
            fit = index.curvature_ball(radius)
        e2.record()
        return index, fit, (e0, e1, e2)

    def e2e_step():
        flush.fill_(1)
        torch.cuda.synchronize()
        t = time.perf_counter()
        pc = PointCloud(points=host_pts, normals=np.zeros((n, 0), np.float32), k_neighbors=max(k, 1))
        if radius is None:
            pc.plant_kdtree(k)
        else:
            pc.plant_ball(radius)
        K, H = pc.compute_pointwise_explicit_quadratic_curvature()
        wall = time.perf_counter() - t
        e2e_step.r_k = np.asarray(pc.dists)[:, -1].astype(np.float64) if (radius is None and ref_kh is not None) else None
        return K, H, wall

    for _ in range(args.warmup):
        index, fit, _ = device_step()
        index.close()
    torch.cuda.synchronize()
    sampler.start()
    steps = []
    launches = 0
    for _ in range(args.steps):
        index, fit, ev = device_step()
        torch.cuda.synchronize()
        steps.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
        launches += int(index.info().build_launches) + (int(index.last_stats().kernel_launches) if radius is None else 2)
        last = (index, fit)
        if _ + 1 < args.steps:
            index.close()
    build_ms = sum(s[0] for s in steps) / len(steps)
    query_ms = sum(s[1] for s in steps) / len(steps)
    index, fit = last
    members = int(fit.counts.sum().item()) if radius is not None else n * k
    nan_rows = int(torch.isnan(fit.column("K")).sum().item())
    index.close()
    for _ in range(args.warmup):
        e2e_step()
    walls = []
    for _ in range(args.steps):
        K, H, w = e2e_step()
        walls.append(w)
    clocks = sampler.stop()
    e2e_ms = 1e3 * sum(walls) / len(walls)
    peak, peak_src = measured_peak_gbs()
    per_q = ALG_BYTES_QUERY if radius is None else ALG_BYTES_BALL
    alg_bytes = n * per_q + (members * ALG_BYTES_MEMBER if radius is not None else 0)
    achieved = alg_bytes / (query_ms * 1e-3) / 1e9
    details = {"build_ms": build_ms, "query_ms": query_ms, "nan_rows": nan_rows, "members": members,
               "members_per_query": members / n, "parallelism": "one GPU"}
    if radius is not None:
        details["radius"] = radius
        details["count_min_max"] = [int(fit.counts.min().item()), int(fit.counts.max().item())]
    elif args.workload == "c3_sphere":
        Kd = np.asarray(K, np.float64)
        details["closed_form"] = {"K_abs_err_median": float(np.median(np.abs(Kd - 1.0))), "H_abs_err_median": float(np.median(np.abs(np.abs(H) - 1.0)))}
    parity = None
    if ref_kh is not None:
        # the K, H this run wrote to the host against the K, H the unmodified reference produced for the same cloud in the
        # same run, with the stated tolerance (oracle/compare.py): |dK| <= 1e-3 |K| + 1e-5 / r_k^2, |dH| <= 1e-3 |H| + 1e-5 / r_k
        Kr, Hr = (a.astype(np.float64) for a in ref_kh)
        Kg, Hg = np.asarray(K, np.float64), np.asarray(H, np.float64)
        r_k = e2e_step.r_k
        ok = np.isfinite(Kr) & np.isfinite(Hr)
        k_bad = (np.abs(Kg - Kr) > 1e-3 * np.abs(Kr) + 1e-5 / r_k ** 2) & ok
        h_bad = (np.abs(Hg - Hr) > 1e-3 * np.abs(Hr) + 1e-5 / r_k) & ok
        parity = {"rows": int(ok.sum()), "K_violations": int(k_bad.sum()), "H_violations": int(h_bad.sum()),
                  "violations": int((k_bad | h_bad).sum()), "nan_rows": int((~np.isfinite(Kg) & ok).sum()),
                  "bit_identical_K": int((np.asarray(K, np.float32).view(np.uint32) == ref_kh[0].view(np.uint32)).sum()),
                  "checker": "K, H of baseline/_ref/pointCloudToolbox.py (unmodified) on the whole cloud in this run; signed, tolerance of oracle/compare.py"
                             + ("; a lattice cloud has exact distance ties at the k-th neighbour, where scipy's traversal order and the (d, index) "
                                "rule pick different, equally valid neighbourhoods: those rows may differ (tests/test_oracle.py pins the rule)"
                                if args.workload == "c1_torus" else "")}
    line = {
        "metric": METRIC, "value": n / ((build_ms + query_ms) * 1e-3), "unit": "points/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": build_ms + query_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 search keys, f64 re-rank and fit, f32 outputs", "data": "synthetic" if not args.workload.startswith("bunny") else "sample_scans/bunny.txt (fixture copy)",
        "config": workload_config(args), "details": details,
        "roofline": {"bound": "hbm", "kernel": "ball_staged_kernel<2, fused>" if radius is not None else "knn_staged_kernel<2, fused>",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "note": (f"algorithmic {per_q} B/query" + (f" + {ALG_BYTES_MEMBER} B per ball member ({members / n:.1f} members/query)" if radius is not None else "")
                              + "; instruction-issue bound, not HBM bound (DESIGN.md 5)")},
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": n * 8,
                "ms_per_step": e2e_ms, "host_io": "one rank"},
        "gpu_launches": launches * 2, "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity"] = parity
    print(json.dumps(line), flush=True)


def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload != "torus":
        if world > 1:
            if rank == 0:
                print(json.dumps({"unavailable": f"workload {args.workload} is a one-GPU case (BASELINE.json configs 1-3)"}), flush=True)
            return
        return small_workload(args)
    n, k = args.points, args.k

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process: the fan-out leg forks worker processes
        cpu = reference_leg(args, args.cpu_seconds)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    # Ranks are spread over the box: neighbouring GPU indices of an HGX board share a PCIe uplink towards the host
    # (profiles/pcie_probe_8gpu_r02n.txt: GPUs 0 and 1 together 72 GB/s D2H, GPUs 0 and 4 together 106 GB/s), so a job
    # of 2 or 4 ranks on 8 visible GPUs takes every 4th / 2nd device.  One rank per GPU either way.
    from point_cloud_toolbox_b200.distributed import device_for_rank

    one_node = int(os.environ.get("LOCAL_WORLD_SIZE", world)) == world
    gpu_of_rank = [device_for_rank(r, world) if (world > 1 and one_node) else r for r in range(max(world, local_rank + 1))]
    gpu_index = gpu_of_rank[local_rank]
    torch.cuda.set_device(gpu_index)
    dev = torch.device("cuda", gpu_index)
    near = None
    if world > 1:
        # host threads and the host pages this rank touches stay on the GPU's socket (no effect on one-socket hosts)
        from point_cloud_toolbox_b200.distributed import bind_near_gpu

        near = bind_near_gpu(gpu_index)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from point_cloud_toolbox_b200 import GridIndex, PointCloud
    from point_cloud_toolbox_b200 import distributed as pdist

    # ---- the cloud: generated on the device (rank 0) before any timing; every rank keeps a replica for the
    # ---- parity crops and the comparison figure, the timed multi-GPU paths only see this rank's SHARE
    pts = make_torus_on_device(n, dev) if rank == 0 else None
    if world > 1:
        pts = pdist.broadcast_cloud(pts, n, None, 0, dev)
    host_pts = None
    if rank == 0:
        host = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
        host.copy_(pts)
        torch.cuda.synchronize()
        host_pts = host.numpy()
    begin, end = pdist.shard_bounds(n, world, rank)
    share = pts[begin:end].clone() if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(record=None):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if world == 1:
            index = GridIndex(pts, k_hint=k)
            e1.record()
            fit = index.curvature_knn(k, want_coeffs=False)
            e2.record()
            handle = index
        else:
            fit = pdist.curvature_knn_exchange(share, begin, n, k)
            e1.record()
            e2.record()
            handle = fit
        if record is not None:
            record.append((e0, e1, e2))
        return handle, fit

    # multi-GPU end to end: the host cloud and the host result live in shared memory mapped by every rank
    # (a one-node job's ranks see the same input), so each rank moves its share over its own PCIe link
    shared_in = shared_out = None
    if world > 1:
        # NUMA placement: every rank touches the rows it will move from a thread bound next to its GPU, page-locks
        # afterwards, and fills its share of the input (every rank holds the cloud for the parity crops)
        tag = f"pct_bench_{os.environ.get('MASTER_PORT', '0')}"
        try:
            shared_in, shared_out = pdist.shared_arrays_placed(tag, [(n, 3), (2, n)], n)
        except Exception as exc:  # pragma: no cover
            raise RuntimeError(f"the multi-GPU end-to-end path needs /dev/shm room for the cloud and the result ({exc!r})")
        shared_in.tensor[begin:end].copy_(pts[begin:end])
        torch.cuda.synchronize()
        dist.barrier()

    def e2e_step(stages=None):
        if world == 1:
            pc = PointCloud(points=host_pts, normals=np.zeros((n, 0), np.float32), k_neighbors=k)
            pc.plant_kdtree(k)
            K, H = pc.compute_pointwise_explicit_quadratic_curvature()
            return K, H
        pdist.curvature_knn_shared(shared_in, shared_out, k, device=dev, stages=stages).close()
        return (shared_out.array[0], shared_out.array[1]) if rank == 0 else (None, None)

    sampler = ClockSampler(gpu_index, args.clock_interval_ms, not args.no_clocks)
    if rank == 0:
        sampler.prepare()

    # Warm-up and timed steps run the SAME loop (same object lifetimes: the previous step's
    # results stay alive until the next step has produced its own, as they would in a caller
    # that keeps its latest result), so pools and caches are in steady state when timing starts.
    def device_loop(steps, events=None):
        last = None
        for _ in range(steps):
            handle, fit = device_step(events)
            if last is not None:
                last[0].close()
            last = (handle, fit)
        return last

    def e2e_loop(steps, walls=None):
        K = H = None
        for _ in range(steps):
            tw = time.perf_counter()
            K, H = e2e_step()
            if walls is not None:
                walls.append(round(1e3 * (time.perf_counter() - tw), 1))
        return K, H

    # ---- device-resident metric ----
    last = device_loop(args.warmup)
    if last is not None:
        last[0].close()
    del last
    barrier()
    if rank == 0:
        sampler.start()
    events = []
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    last = device_loop(args.steps, events)
    t_end.record()
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    launches_step = 0
    peer_return = False
    if world == 1:
        build_steps = [round(a.elapsed_time(b), 2) for a, b, _ in events]
        build_ms = sum(build_steps) / len(events)
        query_ms = sum(b.elapsed_time(c) for _, b, c in events) / len(events)
        index = last[0]
        stats, info = index.last_stats(), index.info()
        status_bad = int((last[1].status != 0).sum().item())
        nan_rows = int(torch.isnan(last[1].column("K")).sum().item())
        pts_per_launch, indexed, unresolved = n, n, 0
        launches_step = int(info.build_launches) + int(stats.kernel_launches)
    else:
        fit = last[1]
        stats, info = fit.index.last_stats(), fit.index.info()
        rec = fit.records
        status_bad = int((rec[:, 7].contiguous().view(torch.int32) != 0).sum().item())
        nan_rows = int(torch.isnan(rec[:, 3]).sum().item())
        pts_per_launch, indexed, unresolved = int(fit.own_ids.numel()), int(fit.indexed), int(fit.unresolved)
        build_steps, build_ms, query_ms = [], None, None
        # bbox/sample torch kernels are not ours; pilot (2), bin count + fill + row starts (3), own flags + rows (2),
        # and with the return fused into the kernel: row ids + peer route (2)
        launches_step = int(info.build_launches) + int(stats.kernel_launches) + 7 + (2 if getattr(fit, "peer_return", False) else 0)
        peer_return = bool(getattr(fit, "peer_return", False))
    last[0].close()
    del last
    torch.cuda.empty_cache()

    # stage times of the multi-GPU device-resident step (separate instrumented pass, after the timed region)
    value_stages = None
    if world > 1:
        st = pdist.Stages()
        fit = pdist.curvature_knn_exchange(share, begin, n, k, stages=st)
        torch.cuda.synchronize()
        value_stages = st.durations_ms()
        fit.close()
        # the kernel's own time: the fused query call alone on this rank's slab
        barrier()

    # ---- end to end through the public API ----
    e2e_loop(args.warmup)
    barrier()
    s0 = torch.cuda.Event(enable_timing=True)
    s1 = torch.cuda.Event(enable_timing=True)
    s0.record()
    wall0 = time.perf_counter()
    step_walls = []
    K, H = e2e_loop(args.steps, step_walls)
    s1.record()
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - wall0)
    e2e_ms = s0.elapsed_time(s1)
    clocks = sampler.stop() if rank == 0 else None
    e2e_stages = None
    if world > 1:
        st = pdist.Stages()
        e2e_step(st)
        torch.cuda.synchronize()
        e2e_stages = st.durations_ms()
        barrier()

    # ---- replicated-cloud comparison figure (round 1's `value`: no collective in the timed region) ----
    replicated_ms = None
    if world > 1:
        for _ in range(2):
            pdist.curvature_knn_slab(pts, k, rank, world).index.close()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            pdist.curvature_knn_slab(pts, k, rank, world).index.close()
        r1.record()
        barrier()
        replicated_ms = r0.elapsed_time(r1) / reps

    # ---- the fused query call alone (roofline), on this rank's slab / the whole index ----
    if world == 1:
        roof_query_ms = query_ms
    else:
        fit = pdist.curvature_knn_exchange(share, begin, n, k)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        q0.record()
        for _ in range(reps):
            fit.index.curvature_knn(k, want_coeffs=False)
        q1.record()
        torch.cuda.synchronize()
        roof_query_ms = q0.elapsed_time(q1) / reps
        fit.close()
        barrier()

    # ---- parity of this run (after every timed region) ----
    parity = None
    if not args.no_parity:
        parity = parity_block(args, world, rank, dev, pts, shared_in, shared_out, host_pts)

    if world > 1:
        pdist.PeerResults.release()
    for sh in (shared_in, shared_out):
        if sh is not None:
            sh.close()

    if world > 1:
        vals = [total_ms, e2e_ms, e2e_wall_ms, roof_query_ms, replicated_ms]
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, e2e_wall_ms, roof_query_ms, replicated_ms = t.tolist()
        for stages in (value_stages, e2e_stages):
            names = sorted(stages)
            t = torch.tensor([stages[s] for s in names], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for s, v in zip(names, t.tolist()):
                stages[s] = round(v, 3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if os.environ.get("PCT_B200_TRACE"):
        from point_cloud_toolbox_b200 import trace

        print("trace (ms, summed over all e2e steps incl. warm-up):", trace.timings, file=sys.stderr)

    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)
    e2e_step_ms = max(e2e_ms, e2e_wall_ms) / args.steps
    peak, peak_src = measured_peak_gbs()
    achieved = pts_per_launch * ALG_BYTES_QUERY / (roof_query_ms * 1e-3) / 1e9
    prof = committed_profile(k)
    details = {
        "parallelism": ("one GPU" if world == 1 else
                        f"{world} slabs across the longest axis; every rank starts with 1/{world} of the cloud, one all-to-all delivers each slab "
                        f"(+ margin) to its rank, "
                        + ("the fused kernel stores K, H straight into the owner rank's array over NVLink peer memory (no return collective)"
                           if peer_return else "one all-to-all returns the rows")
                        + f" (rank 0: {indexed} indexed, {pts_per_launch} answered, {unresolved} redone)"),
        "gpu_of_rank": gpu_of_rank[:world],
        "cell_size": info.cell_size, "cells_level0": info.cells_level0, "index_bytes": info.device_bytes,
        "level1_retries": stats.level1_retries, "exact_path": stats.exact_path, "unstaged": stats.unstaged,
        "build_ms": build_ms, "build_ms_steps": build_steps, "query_ms": roof_query_ms, "status_nonzero": status_bad, "nan_rows": nan_rows,
    }
    if world > 1:
        details["value_stages_ms"] = value_stages
        details["e2e_stages_ms"] = e2e_stages
        details["host_copy_rounds"] = {"h2d_d2h": [list(v) for v in pdist.CopyRounds._cache.values()], "probe_ms": pdist.CopyRounds.timings,
                                       "note": "ranks copy to / from the host in this many rounds (measured on the first call: "
                                               "the GPUs of the box share PCIe uplinks)"}
        details["stage_note"] = "max over ranks of CUDA-event times of one instrumented pass after the timed region; a stage ends where it is named"
    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 search keys, f64 re-rank and fit, f32 outputs", "data": "synthetic",
        "config": workload_config(args), "details": details,
        "roofline": {
            "bound": "hbm", "kernel": "knn_staged_kernel<2, fused> (+ unstaged chunks, level-1 retries and exact tail: one call, CUDA events on its stream)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": float(prof["dram_bytes_per_point"]) * pts_per_launch if prof else None,
            "traffic_source": (prof.get("source") if prof else None),
            "peak_source": peak_src,
            "note": "algorithmic 44 B/point; the kernel is instruction-issue bound, not HBM bound (DESIGN.md 5); `traffic` is the committed "
                    "ncu capture's bytes/point times this run's points, not a measurement of this run",
        },
        "e2e": {"value": n / (e2e_step_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": n * 8,
                "ms_per_step": e2e_step_ms, "step_wall_ms": step_walls,
                "host_io": "one rank" if world == 1 else
                "every rank its share (shared host memory, pages first-touched by the rank that moves them" +
                (f", ranks bound to their GPU's {len(near)} local CPUs)" if near else ", no CPU binding: topology not visible)")},
        "gpu_launches": launches_step * args.steps * 2,
        "gpu_launches_source": "pct_index_info.build_launches + pct_query_stats.kernel_launches of the last timed step (+ 7 slab-exchange kernels at N > 1, + 2 when the return is fused into the kernel), x steps x 2 timed loops",
        "clocks": clocks,
    }
    if world > 1:
        line["value_with_collectives"] = value
        line["value_replicated_no_collective"] = {"value": n / (replicated_ms * 1e-3), "ms_per_step": replicated_ms,
                                                  "note": "round 1's definition: cloud replicated before timing, every rank selects and answers its slab"}
    if parity is not None:
        line["parity"] = parity
    # secondary roofline: algorithmic flops per point -- search 8c + selection 2c with
    # c = 2.9 k candidates, covariance 15 k, rotation 18 k, normal equations 59 k, fixed 600 -- against the FMA
    # rate of the CUDA cores measured here, after the timed regions
    try:
        f32, f64 = fma_peaks()
        flops_pt = 121.0 * k + 600.0
        ach = pts_per_launch * flops_pt / (roof_query_ms * 1e-3) / 1e12
        line["roofline_fp32"] = {"algorithmic_flops_per_point": flops_pt, "achieved": ach, "peak": f32, "unit": "TFLOP/s",
                                 "frac": ach / f32 if f32 > 0 else None, "fp64_peak": f64,
                                 "peak_source": "pct_measure_fma_peaks (FMA chains on this GPU, this run)"}
    except Exception as e:  # diagnostics must not cost the bench line
        line["roofline_fp32"] = {"error": str(e)}
    # the bound that actually holds: warp instructions issued per second against the schedulers' peak.
    # Instructions per point come from the committed ncu capture named in `source`; the time and the clock are this run's.
    if prof and clocks and clocks.get("sm_mhz") and prof.get("warp_inst_per_point"):
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_inst = sms * 4 * float(clocks["sm_mhz"]) * 1e6 / 1e9
        ach_inst = pts_per_launch * float(prof["warp_inst_per_point"]) / (roof_query_ms * 1e-3) / 1e9
        line["roofline_issue"] = {"bound": "instruction issue", "warp_inst_per_point": prof["warp_inst_per_point"],
                                  "achieved": ach_inst, "peak": peak_inst, "unit": "G warp-inst/s", "frac": ach_inst / peak_inst,
                                  "source": prof.get("source"), "note": "instructions per point from the committed capture, not from this run"}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
