"""Timings of the rows of SURVEY section 8(f) either side of the path, on the GPU box (host cores + one B200):

    python scripts/io_bench.py [rows]

text load (library vs np.loadtxt on a bounded sample), curvature PLY write (library vs the reference's per-line
loop on a bounded sample), PLY read, mesh energy kernel (CUDA events; algorithmic 12 B/triangle), PCA rows kernel.
Prints one JSON line.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_cloud_toolbox_b200 import engine, utils as U  # noqa: E402
import point_cloud_toolbox_b200 as pct  # noqa: E402


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t)
    return best, r


def cuda_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")    # 4 x L2: every timed call starts cold
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    res = {"rows": rows, "host_cores": os.cpu_count()}
    rng = np.random.default_rng(0)
    n = int(np.sqrt(rows))
    u, v = np.meshgrid(np.linspace(-6, 6, n), np.linspace(-6, 6, n), indexing="ij")
    pts = np.stack([u, v, np.sin(u) * np.sin(v)], -1).reshape(-1, 3).astype(np.float32)
    pts += rng.uniform(-1e-3, 1e-3, pts.shape).astype(np.float32)
    rows = len(pts)
    with tempfile.TemporaryDirectory() as d:
        pc = pct.PointCloud(points=pts, normals=np.zeros((rows, 0), np.float32), k_neighbors=20)
        pc.plant_kdtree(20)
        K, H = pc.compute_pointwise_explicit_quadratic_curvature()
        # writers
        out = os.path.join(d, "c.ply")
        t, _ = timed(lambda: U.save_curvatures_to_ply(pts, K, H, out), 2)
        res["curvature_ply_write_s"] = t
        res["curvature_ply_bytes"] = os.path.getsize(out)
        res["curvature_ply_rows_per_s"] = rows / t
        sample = min(rows, 200_000)

        def ref_writer():
            with open(os.path.join(d, "ref.ply"), "w") as f:
                for i in range(sample):
                    f.write(f"{pts[i][0]} {pts[i][1]} {pts[i][2]} {K[i]} {H[i]}\n")

        t, _ = timed(ref_writer, 1)
        res["reference_loop_write_rows_per_s"] = sample / t
        # PLY read back
        t, back = timed(lambda: U.parse_ply(out), 2)
        res["ply_read_s"] = t
        res["ply_read_rows_per_s"] = rows / t
        assert np.array_equal(back, pts)
        # text load
        txt = os.path.join(d, "scan.txt")
        pp = os.path.join(d, "p.ply")
        t, _ = timed(lambda: U.save_points_to_ply(pts, pp), 2)
        res["points_ply_write_rows_per_s"] = rows / t
        with open(pp, "rb") as f, open(txt, "wb") as g:
            body = f.read()
            g.write(body[body.index(b"end_header\n") + 11:])
        res["text_bytes"] = os.path.getsize(txt)
        t, tab = timed(lambda: engine.load_text_f32(txt), 2)
        res["text_load_s"] = t
        res["text_load_rows_per_s"] = rows / t
        res["text_load_GBps"] = res["text_bytes"] / t / 1e9
        small = os.path.join(d, "small.txt")
        with open(txt, "rb") as f, open(small, "wb") as g:
            g.write(b"".join(f.readline() for _ in range(min(rows, 2_000_000))))
        t, ref = timed(lambda: np.loadtxt(small), 1)
        res["np_loadtxt_rows_per_s"] = len(ref) / t
        assert np.array_equal(ref.astype(np.float32), tab.numpy()[:len(ref)])
    # the whole caller-side sequence, file in -> file out (validate_shape's order, utils.py:475-551)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "scan.txt")
        U.save_points_to_ply(pts, src)           # rows in the scan text format, after a PLY header
        body = open(src, "rb").read()
        open(src, "wb").write(body[body.index(b"end_header\n") + 11:])
        stages = {}
        t0 = time.perf_counter()
        pc2 = pct.PointCloud(src, k_neighbors=20)
        stages["load_s"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        pc2.plant_kdtree(20)
        K2, H2 = pc2.compute_pointwise_explicit_quadratic_curvature()
        stages["curvature_s"] = time.perf_counter() - t1
        t2 = time.perf_counter()
        U.save_curvatures_to_ply(pc2.points, K2, H2, os.path.join(d, "output_with_curvatures.ply"))
        stages["write_s"] = time.perf_counter() - t2
        stages["total_s"] = time.perf_counter() - t0
        stages["rows_per_s"] = rows / stages["total_s"]
        res["file_to_file"] = stages
    # energy kernel
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = (i * n + j).ravel()
    tris = np.concatenate([np.stack([a, a + 1, a + n], 1), np.stack([a + 1, a + n + 1, a + n], 1)]).astype(np.int32)
    dv, dt = torch.from_numpy(pts).cuda(), torch.from_numpy(tris).cuda()
    dK, dH = torch.from_numpy(np.asarray(K)).cuda(), torch.from_numpy(np.asarray(H)).cuda()
    ms = cuda_ms(lambda: engine.mesh_energies(dv, dt, dK, dH))
    res["energy_triangles"] = len(tris)
    res["energy_ms"] = ms
    res["energy_GBps_algorithmic"] = len(tris) * 12 / ms / 1e6
    res["energy_GBps_with_vertex_arrays"] = (len(tris) * 12 + rows * 20) / ms / 1e6
    res["energies"] = [float(x) for x in engine.mesh_energies(dv, dt, dK, dH).cpu()[:3]]
    # PCA rows
    idx, _ = pc.kdtree.index.knn(20, want_dist=False)
    ms = cuda_ms(lambda: engine.pca_from_neighbors(dv, idx))
    res["pca_rows_ms_k20"] = ms
    res["pca_rows_per_s"] = rows / ms * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
