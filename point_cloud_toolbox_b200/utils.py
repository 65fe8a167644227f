"""The callers' I/O either side of the curvature path.

Same names and behaviour as the corresponding pieces of /root/reference/utils.py, the work done by
libpct_b200.so (host threads; nothing here needs the GPU):

* ``parse_ply``            -- utils.py:979-1004
* ``save_points_to_ply``   -- utils.py:963-976
* ``save_curvatures_to_ply`` -- the block validate_shape writes 'output_with_curvatures.ply' with,
  utils.py:538-551 (inline in the reference, a function here)
* ``save_curvature_arrays`` -- the two ``np.save`` calls of utils.py:504-518

Files written are byte-identical to the reference's; arrays read are bit-identical.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._lib import check, lib


def parse_ply(file_path):
    """(N, 3) float32 array of an ASCII PLY's body; ``None`` (after printing why) on any failure, like the reference."""
    try:
        path = os.fspath(file_path)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        rows, off = ctypes.c_int64(), ctypes.c_int64()
        check(lib.pct_ply_shape(path.encode(), ctypes.byref(rows), ctypes.byref(off)))
        print("Removed header from PLY")
        out = np.empty((rows.value, 3), np.float32)
        check(lib.pct_ply_load_f32(path.encode(), off.value, rows.value, out.ctypes.data_as(ctypes.c_void_p), 0))
        if rows.value == 0:
            return np.array([], dtype=np.float32)  # np.array([]) of the reference's empty list: shape (0,)
        return out
    except FileNotFoundError:
        print(f"File not found: {file_path}")
        return None
    except Exception as e:  # the reference catches everything (utils.py:1003)
        print(f"Error parsing PLY file: {e}")
        return None


def save_points_to_ply(points, filename):
    pts = np.asarray(points)
    if pts.dtype != np.float32:
        pts = pts.astype(np.float64, copy=False)
    pts = np.ascontiguousarray(pts)
    if pts.ndim != 2 or pts.shape[1] != 3:
        # np.savetxt's own complaint for a row that does not match '%.6f %.6f %.6f'
        raise ValueError(f"fmt has wrong number of % formats:  %.6f %.6f %.6f")
    check(lib.pct_write_points_ply(os.fspath(filename).encode(), pts.ctypes.data_as(ctypes.c_void_p),
                                   int(pts.dtype == np.float64), len(pts), 0))
    print(f"point cloud saved in ply format as {filename}")


def save_curvatures_to_ply(points, gaussian_curvature, mean_curvature, filename="output_with_curvatures.ply"):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    K = np.ascontiguousarray(gaussian_curvature, dtype=np.float32)
    H = np.ascontiguousarray(mean_curvature, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise IndexError("points must have shape (N, 3)")
    if pts.shape[1] != 3:
        pts = np.ascontiguousarray(pts[:, :3])
    if len(K) < len(pts) or len(H) < len(pts):
        raise IndexError("list index out of range")  # ref utils.py:550 would run off the shorter list
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    check(lib.pct_write_curvature_ply(os.fspath(filename).encode(), P(pts), P(K), P(H), len(pts), 0))
    print("Point cloud with curvatures saved successfully.")


def save_curvature_arrays(gaussian_curvature, mean_curvature, shape_name, variant, radius, output_dir="./curvature_data"):
    """The ``.npy`` pair of utils.py:504-518 (np.save is already one memcpy per array)."""
    os.makedirs(output_dir, exist_ok=True)
    fg = os.path.join(output_dir, f"{shape_name}_{variant}_radius_{radius}_points_{len(gaussian_curvature)}_gaussian.npy")
    fm = os.path.join(output_dir, f"{shape_name}_{variant}_radius_{radius}_points_{len(mean_curvature)}_mean.npy")
    np.save(fg, gaussian_curvature)
    np.save(fm, mean_curvature)
    print(f"Saved curvature data to {output_dir}")
    return fg, fm


# ---------------------------------------------------------------------------------------------------
# consumers of K and H on the device
# ---------------------------------------------------------------------------------------------------
def compute_energies(vertices, triangles, gaussian_curvature=None, mean_curvature=None, device=None):
    """``(bending_energy, stretching_energy, total_area)`` of a triangle mesh with per-vertex K and H.

    The arithmetic of ``load_mesh_compute_energies`` (utils.py:702-765): triangle areas in fp64, the mean
    of H**2 / of K over the three corners in the curvature arrays' fp32, ``nansum`` of the products.  One
    streaming kernel instead of the reference's O(T^2) loop.  Missing curvature = zeros (utils.py:749-753);
    returns ``(0, 0, 0)`` for no triangles or zero total area (utils.py:724-726, :735-737).
    """
    import torch

    from . import engine

    engine.require_cuda()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    tri = np.asarray(triangles)
    if tri.size == 0:
        return 0, 0, 0
    tri = np.ascontiguousarray(tri.reshape(-1, 3), dtype=np.int32)
    v = torch.from_numpy(np.ascontiguousarray(np.asarray(vertices)[:, :3], dtype=np.float32)).to(dev)
    t = torch.from_numpy(tri).to(dev)
    have = gaussian_curvature is not None and mean_curvature is not None
    g = torch.from_numpy(np.ascontiguousarray(gaussian_curvature, dtype=np.float32)).to(dev) if have else None
    m = torch.from_numpy(np.ascontiguousarray(mean_curvature, dtype=np.float32)).to(dev) if have else None
    out = engine.mesh_energies(v, t, g, m).cpu().numpy()
    if out[3] != 0:
        raise IndexError(f"{int(out[3])} triangles index vertices out of bounds for axis 0 with size {len(v)}")
    if out[2] == 0:
        return 0, 0, 0
    return out[0], out[1], out[2]


def load_mesh_compute_energies(mesh, device=None):
    """Same entry as utils.py:702: ``mesh`` is pyvista-like (``points``, ``faces`` as [3, a, b, c, ...],
    ``point_data``) or open3d-like (``vertices``, ``triangles``); pyvista / open3d themselves are not needed."""
    if mesh is None:
        return 0, 0, 0
    if hasattr(mesh, "triangles"):
        verts, tris = np.asarray(mesh.vertices), np.asarray(mesh.triangles)
    else:
        verts = np.asarray(mesh.points)
        faces = np.asarray(mesh.faces)
        tris = faces.reshape(-1, 4)[:, 1:] if faces.ndim == 1 else faces[:, -3:]    # utils.py:685
    data = getattr(mesh, "point_data", {})
    if "gaussian_curvature" in data and "mean_curvature" in data:
        return compute_energies(verts, tris, np.asarray(data["gaussian_curvature"]), np.asarray(data["mean_curvature"]), device)
    return compute_energies(verts, tris, None, None, device)


def estimate_curvature(points, k_fraction=0.025, max_neighbors=100, device=None, reference_compatible=False):
    """Surface variation l3 / (l1 + l2 + l3 + 1e-10) of every point's k-neighbourhood (itself included, as
    sklearn's ``kneighbors(points)`` lists it), float64 -- the quantity utils.py:778-829 documents.

    Deviation, and the switch that undoes it: the reference's ``einsum('nik,njk->nij')`` contracts over the
    COORDINATE axis, so what it diagonalises is the k x k Gram matrix of the centred neighbourhood, whose k - 3
    smallest eigenvalues are exactly zero in exact arithmetic; its return value ``eigenvalues[:, 0] / (sum + 1e-10)``
    is therefore rounding noise around 0 (|value| ~ 1e-17, either sign).  ``reference_compatible=True`` computes
    exactly that -- the Gram matrix, ``eigvalsh`` (batched, on the device), the same ratio -- so a caller gets the
    reference's numbers up to that noise; the default returns the documented ratio of the 3 x 3 covariance.
    Points must be three-dimensional.
    """
    import torch

    from . import engine
    from ._lib import MAX_K

    pts = np.asarray(points)
    n = len(pts)
    k = min(max(5, int(k_fraction * n)), max_neighbors)                           # utils.py:806
    if k > n:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = {n}")
    if k - 1 > MAX_K:
        raise ValueError(f"at most {MAX_K + 1} neighbours")
    d = engine.to_device_points(pts, device)
    index = engine.GridIndex(d, k_hint=k - 1)
    idx, _ = index.knn(k - 1, want_dist=False)
    if reference_compatible:
        out = torch.empty(n, dtype=torch.float64, device=d.device)
        me = torch.arange(n, device=d.device)
        chunk = max(1, (64 << 20) // (k * k * 8))
        src = d.double() if pts.dtype == np.float64 else d                         # the reference computes in the input's dtype
        for b in range(0, n, chunk):
            e = min(n, b + chunk)
            rows = torch.cat((me[b:e, None], idx[b:e].long()), 1)                  # kneighbors(points): the point itself first
            nb = src[rows]                                                         # (m, k, 3)            utils.py:815
            c = nb - nb.mean(1, keepdim=True)                                      # utils.py:818-819
            gram = torch.einsum("nik,njk->nij", c, c) / (k - 1)                    # utils.py:822 (k x k, as written)
            ev = torch.linalg.eigvalsh(gram.double())                              # utils.py:825, ascending
            out[b:e] = ev[:, 0] / (ev.sum(1) + 1e-10)                              # utils.py:827-828
        return out.cpu().numpy()
    values, _ = engine.pca_from_neighbors(d, idx, include_self=True, want_directions=False)
    return values[:, 5].cpu().numpy()
