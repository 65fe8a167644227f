"""Headline benchmark: points/s for kNN + PCA + quadric curvature on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--k 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path on the host cores

A step is one pass of the hot path over one synthetic cloud:
    value  index build + fused kNN/fit kernel, raw xyz already in HBM        (device timed)
    e2e    PointCloud(points=host) -> plant_kdtree(k) -> compute_pointwise_explicit_quadratic_curvature()
           with pinned host input and host K, H output inside the timed region
Multi-GPU is strong scaling on one fixed cloud: the cloud is replicated (NCCL broadcast), the ranks
agree on cut planes, every rank indexes its own slab (plus a margin) and answers the points in it.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points/sec for kNN+SVD+quadric curvature"
ALG_BYTES_QUERY = 44   # per point: 16 B own record read + 28 B result written; 32 B are actually written
ALG_BYTES_E2E = 76     # + 12 B raw read + 16 B sorted record + 4 B permutation
# our kernels per step (profiles/launches_r01t.txt): bbox, pilot keys, 2 x level hist, Morton keys, gather, table fill,
# staged kNN+fit, L1/L2 kNN+fit x2 (unstaged chunks, level-1 retries), exact tail, stats; CUB's sort kernels not counted
OUR_KERNELS_PER_STEP = 12


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--clock-interval-ms", type=int, default=200)
    return ap.parse_args()


def workload_name(n, k):
    return f"synthetic torus surface (R=1, r=1/3, uniform u,v, seed 3), N={n}, k={k} kNN"


def host_sample(n_sample, seed=3):
    """Host copy of the workload's distribution for the CPU legs (same surface, same sampling law)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 2 * np.pi, n_sample)
    v = rng.uniform(0, 2 * np.pi, n_sample)
    R, r = 1.0, 1.0 / 3.0
    return np.stack(((R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)), 1).astype(np.float32)


def cpu_sample_cloud(n_total, rows_needed):
    """A cloud whose point spacing equals the full workload's would need all N points; the CPU path's
    cost per point does not depend on spacing, so the legs run on a self-contained sample cloud."""
    return host_sample(int(min(n_total, max(20_000, rows_needed))))


# --------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------
def run_cpu_leg(n_total, k, seconds):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    from oracle import baseline

    cores = baseline.host_cores()
    # size the sample cloud so that the leg really spends `seconds` on all cores
    cost = baseline.per_point_seconds(host_sample(20_000), k, probe=1000)
    rows = int(min(n_total, 4_000_000, max(50_000, cores * seconds / cost)))
    pts = cpu_sample_cloud(n_total, rows)
    res = baseline.timed_reference(pts, k, seconds=seconds, procs=cores)
    return res, len(pts)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = args.k
    step_seconds = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    info = None
    for s in range(args.warmup + args.steps):
        res, cloud_n = run_cpu_leg(args.points, k, step_seconds)
        if s >= args.warmup:
            vals.append(res)
        info = (res, cloud_n)
    total_rows = sum(r["rows"] for r in vals)
    total_s = sum(r["seconds"] for r in vals)
    value = total_rows / total_s
    res, cloud_n = info
    sample = (f"{res['rows']} query points per step of a {cloud_n}-point cloud drawn from the workload's surface; the "
              f"reference's per-point loop (cKDTree.query + np.cov + svd + lstsq per point) fanned out over {res['cores']} processes")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(1, len(vals)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 inputs, f64 LAPACK", "data": "synthetic",
        "config": {"workload": workload_name(args.points, k), "k": k, "points": args.points},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": res["cores"], "kind": "port", "sample": sample,
                         "per_point_us_single_core": res["per_point_us_single_core"]},
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


def profiled_traffic(points_per_launch, k):
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if int(t.get("k", -1)) != k:
            return None
        return float(t["dram_bytes_per_point"]) * points_per_launch
    except Exception:
        return None


def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, k = args.points, args.k

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process: the leg forks worker processes
        res, cloud_n = run_cpu_leg(n, k, args.cpu_seconds)
        cpu = {"value": res["points_per_s"], "unit": "points/s", "cores": res["cores"], "kind": "port",
               "sample": (f"{res['rows']} query points of a {cloud_n}-point cloud from the workload's surface, the reference's "
                          f"per-point loop on {res['cores']} processes, {res['seconds']:.1f} s"),
               "per_point_us_single_core": res["per_point_us_single_core"]}

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from point_cloud_toolbox_b200 import GridIndex, PointCloud
    from point_cloud_toolbox_b200 import distributed as pdist
    from point_cloud_toolbox_b200._lib import LAYOUT_SLICE

    # ---- the cloud: generated on the device (rank 0), replicated before any timing ----
    if rank == 0:
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
    else:
        pts = None
    if world > 1:
        pts = pdist.broadcast_cloud(pts, n, None, 0, dev)
    host_pts = None
    if rank == 0:
        host = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
        host.copy_(pts)
        torch.cuda.synchronize()
        host_pts = host.numpy()

    mode = pdist.default_mode(world)
    begin, end = pdist.shard_bounds(n, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(record=None):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e2 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if world == 1:
            index = GridIndex(pts, k_hint=k)
            e1.record()
            fit = index.curvature_knn(k, want_coeffs=False)
        elif mode == "replicated":
            index = GridIndex(pts, k_hint=k)
            e1.record()
            fit = index.curvature_knn(k, begin, end, layout=LAYOUT_SLICE, want_coeffs=False)
        else:
            # cut planes + slab selection + slab index + fused kernel on the points this rank owns
            part = pdist.curvature_knn_slab(pts, k, rank, world, events=(e1,))
            index, fit = part.index, part
        e2.record()
        if record is not None:
            record.append((e0, e1, e2))
        return index, fit

    # multi-GPU end to end: the host cloud and the host result live in shared memory mapped by every rank
    # (a one-node job's ranks see the same input), so each rank moves its share over its own PCIe link;
    # if the node's /dev/shm cannot hold them, rank 0 moves everything (broadcast in, gather out)
    shared_in = shared_out = None
    if world > 1:
        tag = f"pct_bench_{os.environ.get('MASTER_PORT', '0')}"
        ok = torch.ones(1, device=dev)
        try:
            if rank == 0:
                shared_in = pdist.SharedHostArray(tag + "_in", (n, 3), create=True)
                shared_out = pdist.SharedHostArray(tag + "_out", (2, n), create=True)
                shared_in.array[:] = host_pts
        except Exception as exc:  # pragma: no cover
            print(f"shared host memory unavailable ({exc!r}); rank 0 moves all host data", file=sys.stderr)
            ok.zero_()
        dist.broadcast(ok, 0)
        if bool(ok.item()) and rank != 0:
            shared_in = pdist.SharedHostArray(tag + "_in", (n, 3), create=False)
            shared_out = pdist.SharedHostArray(tag + "_out", (2, n), create=False)
        if not bool(ok.item()):
            shared_in = shared_out = None
        dist.barrier()

    def e2e_step():
        if world == 1:
            pc = PointCloud(points=host_pts, normals=np.zeros((n, 0), np.float32), k_neighbors=k)
            pc.plant_kdtree(k)
            K, H = pc.compute_pointwise_explicit_quadratic_curvature()
            return K, H
        if shared_in is not None:
            pdist.curvature_knn_shared(shared_in, shared_out, k, device=dev)
            return (shared_out.array[0], shared_out.array[1]) if rank == 0 else (None, None)
        out = pdist.curvature_knn_sharded(host_pts, n, k, device=dev)
        if rank == 0:
            from point_cloud_toolbox_b200.engine import to_host

            kh = to_host(out.t())
            return kh[0], kh[1]
        return None, None

    sampler = ClockSampler(local_rank, args.clock_interval_ms, not args.no_clocks)
    if rank == 0:
        sampler.prepare()

    # Warm-up and timed steps run the SAME loop (same object lifetimes: the previous step's
    # results stay alive until the next step has produced its own, as they would in a caller
    # that keeps its latest result), so pools and caches are in steady state when timing starts.
    def device_loop(steps, events=None):
        last = None
        for _ in range(steps):
            index, fit = device_step(events)
            if last is not None:
                last[0].close()
            last = (index, fit)
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
    build_steps = [round(a.elapsed_time(b), 2) for a, b, _ in events]
    build_ms = sum(build_steps) / len(events)
    query_ms = sum(b.elapsed_time(c) for _, b, c in events) / len(events)
    stats = last[0].last_stats()
    info = last[0].info()
    if world == 1 or mode == "replicated":
        status_bad = int((last[1].status != 0).sum().item())
        nan_rows = int(torch.isnan(last[1].column("K")).sum().item())
        pts_per_launch, slab_points, unresolved = end - begin, n, 0
    else:
        rec = last[1].records
        status_bad = int((rec[:, 7].contiguous().view(torch.int32) != 0).sum().item())
        nan_rows = int(torch.isnan(rec[:, 3]).sum().item())
        pts_per_launch, slab_points, unresolved = int(last[1].ids.numel()), int(last[0].n), int(last[1].unresolved)
    last[0].close()
    del last
    torch.cuda.empty_cache()

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
    e2e_host_io = "one rank" if world == 1 else ("every rank its share (shared host memory)" if shared_in is not None else "rank 0")
    if rank == 0 and world > 1 and shared_out is not None:
        K = np.array(K[:1024])  # detach from the segment before it is unmapped
    for sh in (shared_in, shared_out):
        if sh is not None:
            sh.close()
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, query_ms, build_ms, e2e_wall_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, query_ms, build_ms, e2e_wall_ms = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if os.environ.get("PCT_B200_TRACE"):
        from point_cloud_toolbox_b200 import trace

        print("trace (ms, summed over all e2e steps incl. warm-up):", trace.timings, file=sys.stderr)

    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)
    e2e_value = n / (max(e2e_ms, e2e_wall_ms) / args.steps * 1e-3)
    peak, peak_src = measured_peak_gbs()
    achieved = pts_per_launch * ALG_BYTES_QUERY / (query_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 search keys, f64 re-rank and fit, f32 outputs", "data": "synthetic",
        "config": {
            "workload": workload_name(n, k), "k": k, "points": n, "parallelism": ("one GPU" if world == 1 else
                                                                    f"cloud replicated, every rank builds the whole index and answers 1/{world} of the Morton-sorted queries"
                                                                    if mode == "replicated" else
                                                                    f"{world} slabs across the longest axis: cloud replicated, each rank indexes and answers its "
                                                                    f"slab (rank 0: {slab_points} indexed, {pts_per_launch} answered, {unresolved} redone on a whole-cloud index)"),
            "l2": "inputs (1.2 GB raw + 1.6 GB sorted at 100M) exceed the 126 MB L2; no flush needed",
            "cell_size": info.cell_size, "cells_level0": info.cells_level0, "index_bytes": info.device_bytes,
            "level1_retries": stats.level1_retries, "exact_path": stats.exact_path,
            "build_ms": build_ms, "build_ms_steps": build_steps, "query_ms": query_ms, "status_nonzero": status_bad, "nan_rows": nan_rows,
        },
        "roofline": {
            "bound": "hbm", "kernel": "knn_staged_kernel<2,true,false> (+ unstaged chunks, level-1 retries and exact tail: one call, CUDA events on its stream)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": profiled_traffic(pts_per_launch, k), "peak_source": peak_src,
            "note": "algorithmic 44 B/point; the kernel is instruction-issue bound (72 % of issue slots, profiles/staged_full_r01t.txt), not HBM bound (DESIGN.md 5)",
        },
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": n * 8,
                "ms_per_step": max(e2e_ms, e2e_wall_ms) / args.steps, "step_wall_ms": step_walls, "host_io": e2e_host_io},
        "gpu_launches": OUR_KERNELS_PER_STEP * args.steps * 2,
        "clocks": clocks,
    }
    # secondary roofline: algorithmic flops per point -- search 8c + selection 2c with
    # c = 2.9 k candidates, covariance 15 k, rotation 18 k, normal equations 59 k, fixed 600 -- against the FMA
    # rate of the CUDA cores measured here, after the timed regions
    try:
        if world > 1:
            raise RuntimeError("measured at N = 1 only (like cpu_baseline)")
        from point_cloud_toolbox_b200 import _lib
        import ctypes

        f32, f64 = ctypes.c_double(), ctypes.c_double()
        _lib.check(_lib.lib.pct_measure_fma_peaks(ctypes.byref(f32), ctypes.byref(f64), None))
        flops_pt = 121.0 * k + 600.0
        ach = pts_per_launch * flops_pt / (query_ms * 1e-3) / 1e12
        line["roofline_fp32"] = {"algorithmic_flops_per_point": flops_pt, "achieved": ach, "peak": f32.value, "unit": "TFLOP/s",
                                 "frac": ach / f32.value if f32.value > 0 else None, "fp64_peak": f64.value,
                                 "peak_source": "pct_measure_fma_peaks (FMA chains on this GPU, this run)"}
    except Exception as e:  # diagnostics must not cost the bench line
        if world == 1:
            line["roofline_fp32"] = {"error": str(e)}
    # the bound that actually holds: warp instructions issued per second against the schedulers' peak.
    # Instructions per point come from the committed ncu capture of the staged kernel, the time is this run's.
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            prof = json.load(f)
        if int(prof.get("k", -1)) == k and clocks and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_inst = sms * 4 * float(clocks["sm_mhz"]) * 1e6 / 1e9
            ach_inst = pts_per_launch * float(prof["warp_inst_per_point"]) / (query_ms * 1e-3) / 1e9
            line["roofline_issue"] = {"bound": "instruction issue", "warp_inst_per_point": prof["warp_inst_per_point"],
                                      "achieved": ach_inst, "peak": peak_inst, "unit": "G warp-inst/s", "frac": ach_inst / peak_inst,
                                      "source": prof.get("source_inst")}
    except Exception:
        pass
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
