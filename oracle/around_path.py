"""CPU restatement of the reference code either side of the curvature path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Pinned**: every function here is checked in
tests/test_oracle.py against tests/golden/io_energy_pca.npz, which oracle/make_golden_io.py produced by
running the unmodified reference functions.
"""
from __future__ import annotations

import numpy as np


def parse_ply(text: str):
    """utils.py:979-1004: skip to the line 'end_header', float() of the first three tokens of every later line."""
    lines = text.split("\n")
    if text.endswith("\n"):
        lines = lines[:-1]           # readline() yields no empty last line
    it = iter(lines)
    for line in it:
        if line.strip() == "end_header":
            break
    else:
        raise ValueError("no end_header")
    pts = []
    for line in it:
        x, y, z = map(float, line.split()[:3])
        pts.append([x, y, z])
    return np.array(pts, dtype=np.float32)


def points_ply_bytes(points) -> bytes:
    """utils.py:963-976: header + np.savetxt(fmt='%.6f %.6f %.6f')."""
    points = np.asarray(points)
    head = ("ply\nformat ascii 1.0\n" + f"element vertex {len(points)}\n" +
            "property float x\nproperty float y\nproperty float z\nend_header\n")
    body = "".join("%.6f %.6f %.6f\n" % (float(p[0]), float(p[1]), float(p[2])) for p in points)
    return (head + body).encode()


def curvature_ply_bytes(points, gaussian_curvature, mean_curvature) -> bytes:
    """utils.py:538-551: f-strings of numpy float32 scalars, i.e. repr(float(v))."""
    head = ("ply\nformat ascii 1.0\n" + f"element vertex {len(points)}\n" + "property float x\nproperty float y\n"
            "property float z\nproperty float gaussian_curvature\nproperty float mean_curvature\nend_header\n")
    rows = []
    for i in range(len(points)):
        rows.append(f"{points[i][0]} {points[i][1]} {points[i][2]} {gaussian_curvature[i]} {mean_curvature[i]}\n")
    return (head + "".join(rows)).encode()


def mesh_energies(vertices, triangles, gaussian_curvature=None, mean_curvature=None):
    """utils.py:702-765 without the O(T^2) recomputation: (bending, stretching, total area)."""
    v = np.asarray(vertices, dtype=np.float64)                      # o3d vertices are float64 (:722)
    t = np.asarray(triangles)
    if len(t) == 0:
        return 0, 0, 0                                              # :724-726
    a, b, c = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    areas = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)    # :728-733
    if np.sum(areas) == 0:
        return 0, 0, 0                                              # :735-737
    if gaussian_curvature is None or mean_curvature is None:        # :749-753
        K = np.zeros(len(v)); H = np.zeros(len(v))
    else:
        K = np.asarray(gaussian_curvature); H = np.asarray(mean_curvature)
    H2 = H ** 2                                                     # :747, in the arrays' own dtype
    third = K.dtype.type(3)
    # np.mean of three elements: sequential sum in the array's dtype, divided by 3 (:757-759)
    face_K = (((K[t[:, 0]] + K[t[:, 1]]) + K[t[:, 2]]) / third).astype(np.float64)
    face_H2 = (((H2[t[:, 0]] + H2[t[:, 1]]) + H2[t[:, 2]]) / third).astype(np.float64)
    return np.nansum(face_H2 * areas), np.nansum(face_K * areas), np.sum(areas)   # :762-764


def pca_principal_curvatures(points, k):
    """pointCloudToolbox.py:901-945, per point: (l1, l2, K, H, directions (3, 2))."""
    from scipy.linalg import eigh

    points = np.asarray(points)
    n = len(points)
    l1 = np.zeros(n); l2 = np.zeros(n); dirs = np.zeros((n, 3, 2))
    for i in range(n):
        d = np.linalg.norm(points - points[i], axis=1)              # fp32 for an fp32 cloud (:914)
        nb = points[np.argsort(d)[1:k + 1]]                         # :915-916
        w, vec = eigh(np.cov(nb, rowvar=False))                     # :922-925
        order = np.argsort(w)[::-1]
        vec = vec[:, order]
        l1[i] = w.max(); l2[i] = np.delete(w, np.argmax(w)).max()   # :932-934
        dirs[i] = vec[:, :2]
    return l1, l2, l1 * l2, (l1 + l2) / 2, dirs


def pca_from_rows(points, rows, include_self=False):
    """Same quantities from given neighbour rows (any exact kNN), vectorised: values (n, 6), directions (n, 3, 2).

    values = [l1, l2, l3, l1*l2, (l1+l2)/2, l3/(l1+l2+l3+1e-10)]; the last is the surface variation that
    utils.py:778-829 documents.
    """
    points = np.asarray(points, dtype=np.float64)
    nb = points[rows]
    if include_self:
        nb = np.concatenate([points[:len(rows), None, :], nb], 1)
    c = nb - nb.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / (nb.shape[1] - 1)
    w, vec = np.linalg.eigh(cov)
    w = w[:, ::-1]; vec = vec[:, :, ::-1]
    vals = np.stack([w[:, 0], w[:, 1], w[:, 2], w[:, 0] * w[:, 1], (w[:, 0] + w[:, 1]) / 2,
                     w[:, 2] / (w.sum(1) + 1e-10)], 1)
    return vals, vec[:, :, :2]
