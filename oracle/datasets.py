"""Synthetic inputs for the BASELINE.json configs, with closed-form curvature.

TEST INFRASTRUCTURE ONLY.  The shape formulas are the ones the reference uses
to validate itself (``/root/reference/utils.py:833-959`` ``generate_pv_shapes``),
re-derived here; sampling follows SURVEY.md section 8(d):

C1  torus R=1, r=1/3 on the reference's 317 x 317 (theta, phi) grid, written
    with 6 decimals and read back like ``PointCloud(file_path)`` would
    (``sample_scans/torus.txt`` itself is absent, .MISSING_LARGE_BLOBS:13)
C3  sphere (Fibonacci lattice, utils.py:858-866), torus (uniform u, v),
    egg carton z = sin x sin y (matches sample_scans/egg_carton.txt)
C4  "scanned sheet": egg carton sampled with a 3-component Gaussian mixture
    density + N(0, 1e-3) noise (mesh_snaps/*.vtk are absent)
"""
from __future__ import annotations

import io

import numpy as np


def sphere_fibonacci(n, radius=1.0):
    """utils.py:858-866. K* = 1/R^2, |H*| = 1/R."""
    i = np.arange(0, n, dtype=np.float64) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + np.sqrt(5)) * i
    p = np.stack((np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)), axis=1) * radius
    K = np.full(n, 1.0 / radius ** 2)
    H = np.full(n, 1.0 / radius)
    return p.astype(np.float32), K, H


def torus_uv(u, v, R=1.0, r=1.0 / 3.0):
    """utils.py:889-891 parametrisation; v is the tube angle."""
    x = (R + r * np.cos(v)) * np.cos(u)
    y = (R + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v)
    K = np.cos(v) / (r * (R + r * np.cos(v)))
    H = (R + 2 * r * np.cos(v)) / (2 * r * (R + r * np.cos(v)))
    return np.stack((x, y, z), axis=1), K, H


def torus_random(n, seed=0, R=1.0, r=1.0 / 3.0):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 2 * np.pi, n)
    v = rng.uniform(0, 2 * np.pi, n)
    p, K, H = torus_uv(u, v, R, r)
    return p.astype(np.float32), K, H


def torus_grid(grid=317, R=1.0, r=1.0 / 3.0):
    """utils.py:883-892: grid x grid (theta, phi) lattice, endpoint excluded."""
    t = np.linspace(0, 2 * np.pi, grid, endpoint=False)
    U, V = np.meshgrid(t, t)
    p, K, H = torus_uv(U.ravel(), V.ravel(), R, r)
    return p, K, H


def as_text_file_cloud(points64, fmt="%.6f"):
    """Round-trip through 3-column text and apply the loader's conditioning.

    Mirrors ``PointCloud(file_path)``: np.loadtxt -> fp32 -> x -= max(x),
    y -= max(y) in fp32 (ref pointCloudToolbox.py:51-57).
    """
    buf = io.StringIO()
    np.savetxt(buf, points64, fmt=fmt)
    buf.seek(0)
    table = np.loadtxt(buf)
    pts = table[:, 0:3].astype(np.float32)
    pts[:, 0] -= np.max(pts[:, 0])
    pts[:, 1] -= np.max(pts[:, 1])
    return pts


def torus_c1():
    """C1 stand-in: 317^2 = 100 489 points (plot_shape_validation_results.py:114-116)."""
    p64, K, H = torus_grid(317)
    return as_text_file_cloud(p64), K, H


def _monge_curvature(zx, zy, zxx, zyy, zxy):
    g = 1 + zx ** 2 + zy ** 2
    K = (zxx * zyy - zxy ** 2) / g ** 2
    H = ((1 + zx ** 2) * zyy - 2 * zx * zy * zxy + (1 + zy ** 2) * zxx) / (2 * g ** 1.5)
    return K, H


def egg_carton_xy(x, y):
    """z = sin x sin y (sample_scans/egg_carton.txt) with Monge-patch K, H."""
    z = np.sin(x) * np.sin(y)
    K, H = _monge_curvature(np.cos(x) * np.sin(y), np.sin(x) * np.cos(y), -z, -z, np.cos(x) * np.cos(y))
    return np.stack((x, y, z), axis=1), K, H


def egg_carton_random(n, seed=1, half_width=2 * np.pi):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-half_width, half_width, n)
    y = rng.uniform(-half_width, half_width, n)
    p, K, H = egg_carton_xy(x, y)
    return p.astype(np.float32), K, H


def egg_carton_generator(n, seed=1):
    """utils.py:905-915 surface z = 0.1 sin(pi x) cos(pi y) on [-1,1]^2, random sampling."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, n)
    y = rng.uniform(-1, 1, n)
    z = 0.1 * np.sin(np.pi * x) * np.cos(np.pi * y)
    zx = 0.1 * np.pi * np.cos(np.pi * x) * np.cos(np.pi * y)
    zy = -0.1 * np.pi * np.sin(np.pi * x) * np.sin(np.pi * y)
    zxx = -np.pi ** 2 * z
    zyy = -np.pi ** 2 * z
    zxy = -0.1 * np.pi ** 2 * np.cos(np.pi * x) * np.sin(np.pi * y)
    K, H = _monge_curvature(zx, zy, zxx, zyy, zxy)
    return np.stack((x, y, z), axis=1).astype(np.float32), K, H


def scanned_sheet(n=332_757, seed=2, noise=1e-3, half_width=2 * np.pi):
    """C4 stand-in: non-uniform density so epsilon-ball counts vary by > 10x."""
    rng = np.random.default_rng(seed)
    centres = np.array([[-3.0, -2.0], [2.5, 1.0], [0.0, 4.0]])
    sigmas = np.array([0.8, 2.0, 4.0])
    weights = np.array([0.3, 0.4, 0.3])
    xy = np.empty((0, 2))
    while len(xy) < n:
        m = 2 * (n - len(xy)) + 1024
        comp = rng.choice(3, size=m, p=weights)
        cand = centres[comp] + rng.normal(size=(m, 2)) * sigmas[comp, None]
        cand = cand[(np.abs(cand) <= half_width).all(axis=1)]
        xy = np.concatenate((xy, cand))
    xy = xy[:n]
    p, K, H = egg_carton_xy(xy[:, 0], xy[:, 1])
    p = p + rng.normal(scale=noise, size=p.shape)
    return p.astype(np.float32), K, H


def interior_mask_xy(points, half_width, margin):
    """Points farther than ``margin`` from the open boundary of an (x, y) patch."""
    return (np.abs(points[:, 0]) < half_width - margin) & (np.abs(points[:, 1]) < half_width - margin)


def grid_mesh(n, seed):
    """Triangulated n x n jittered grid on z = sin x sin y: (vertices (n*n, 3) float32, triangles int32)."""
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.linspace(-2, 2, n), np.linspace(-2, 2, n), indexing="ij")
    u = u + rng.uniform(-0.02, 0.02, u.shape)
    v = v + rng.uniform(-0.02, 0.02, v.shape)
    verts = np.stack([u, v, np.sin(u) * np.sin(v)], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = (i * n + j).ravel()
    tris = np.concatenate([np.stack([a, a + 1, a + n], 1), np.stack([a + 1, a + n + 1, a + n], 1)]).astype(np.int32)
    return verts, tris
