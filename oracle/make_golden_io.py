"""Generate tests/golden/io_energy_pca.npz by running UNMODIFIED reference code for the rows of SURVEY.md
section 8(f) that sit either side of the curvature path.

TEST INFRASTRUCTURE ONLY.  Run in the dev container (``/root/reference`` is absent on the GPU box):

    python -m oracle.make_golden_io

``/root/reference/utils.py`` cannot be imported here (open3d, pyvista are not installed), so the function
definitions are taken out of its syntax tree and compiled as they stand, with the names they use bound to
numpy / logging and -- for ``load_mesh_compute_energies`` only -- a stand-in for ``convert_pv_to_o3d`` that
hands back the vertex and triangle arrays (the conversion is not part of the arithmetic):

* ``save_points_to_ply``            utils.py:963-976
* ``parse_ply``                     utils.py:979-1004
* ``load_mesh_compute_energies``    utils.py:702-765
* the ``with open('output_with_curvatures.ply', 'w')`` block of ``validate_shape``   utils.py:538-551
* ``PointCloud.principal_curvatures_via_principal_component_analysis``   pointCloudToolbox.py:901-945 (imported
  the same way as oracle/make_golden.py does)
"""
from __future__ import annotations

import ast
import contextlib
import io
import logging
import os
import tempfile

import numpy as np

from .datasets import grid_mesh
from .make_golden import GOLDEN_DIR, REFERENCE_ROOT, import_reference


def reference_utils_functions(names):
    src = open(os.path.join(REFERENCE_ROOT, "utils.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "logging": logging, "os": os}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), "/root/reference/utils.py", "exec")
            exec(code, ns)
    return ns, tree


def curvature_ply_block(tree):
    """The `with open('output_with_curvatures.ply', 'w') as ply_file:` statement of validate_shape, as code."""
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == "validate_shape":
            for node in ast.walk(fn):
                if isinstance(node, ast.With):
                    call = node.items[0].context_expr
                    if (isinstance(call, ast.Call) and call.args and isinstance(call.args[0], ast.Constant)
                            and call.args[0].value == "output_with_curvatures.ply"):
                        return compile(ast.Module(body=[node], type_ignores=[]), "/root/reference/utils.py", "exec")
    raise RuntimeError("block not found")


class _FakeO3dMesh:  # what convert_pv_to_o3d returns, reduced to what load_mesh_compute_energies touches
    def __init__(self, vertices, triangles):
        self.vertices = np.asarray(vertices, dtype=np.float64)   # o3d.utility.Vector3dVector is float64
        self.triangles = np.asarray(triangles, dtype=np.int32)   # Vector3iVector is int32

    def has_triangles(self):
        return len(self.triangles) > 0

    def compute_triangle_normals(self):
        pass


class _FakePvMesh:
    def __init__(self, vertices, triangles, point_data):
        self.vertices, self.triangles, self.point_data = vertices, triangles, point_data


def main():
    out = {}
    ns, tree = reference_utils_functions({"save_points_to_ply", "parse_ply", "load_mesh_compute_energies"})
    rng = np.random.default_rng(11)
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
        # --- points PLY writer, float64 and float32 input
        pts64 = rng.standard_normal((300, 3)) * 10.0 ** rng.integers(-8, 6, (300, 3))
        pts64[5] = [0.0, -0.0, 1e-7]
        pts64[6] = [0.5e-6, 1.5e-6, 2.5e-6]                     # ties of '%.6f'
        pts64[7] = [np.nan, np.inf, -np.inf]
        for tag, arr in (("f64", pts64), ("f32", pts64.astype(np.float32))):
            path = os.path.join(d, f"p_{tag}.ply")
            ns["save_points_to_ply"](arr, path)
            out[f"points_ply_in_{tag}"] = arr
            out[f"points_ply_bytes_{tag}"] = np.frombuffer(open(path, "rb").read(), np.uint8)
        # --- PLY parser on that file (extra columns, exponents) and on a file with faces after the vertices
        path = os.path.join(d, "in.ply")
        body = "".join(f"{float(a)!r} {b:.9e} {c:+.4f} 0.5 7\n" for a, b, c in pts64[8:200])
        open(path, "w").write("ply\nformat ascii 1.0\nelement vertex 192\nproperty float x\n  end_header  \n" + body + "3 0 1 2\n")
        out["parse_ply_text"] = np.frombuffer(open(path, "rb").read(), np.uint8)
        out["parse_ply_points"] = ns["parse_ply"](path)
        # --- curvature PLY block
        n = 400
        points = (rng.standard_normal((n, 3)) * 10.0 ** rng.integers(-7, 18, (n, 3))).astype(np.float32)
        gaussian_curvature = (rng.standard_normal(n) * 10.0 ** rng.integers(-9, 9, n)).astype(np.float32)
        mean_curvature = (rng.standard_normal(n) * 10.0 ** rng.integers(-9, 9, n)).astype(np.float32)
        gaussian_curvature[:4] = [np.nan, np.inf, -0.0, 1e-45]
        mean_curvature[:4] = [0.0, -np.inf, 1e16, 9.999e-5]
        points[0] = [0.1, 1e-4, 1e16]
        cwd = os.getcwd()
        os.chdir(d)
        try:
            exec(curvature_ply_block(tree), {"points": points, "gaussian_curvature": gaussian_curvature,
                                             "mean_curvature": mean_curvature})
            out["curv_ply_bytes"] = np.frombuffer(open("output_with_curvatures.ply", "rb").read(), np.uint8)
        finally:
            os.chdir(cwd)
        out["curv_ply_points"], out["curv_ply_K"], out["curv_ply_H"] = points, gaussian_curvature, mean_curvature
        # --- energies
        verts, tris = grid_mesh(28, 3)
        K = (rng.standard_normal(len(verts)) * 3).astype(np.float32)
        H = (rng.standard_normal(len(verts)) * 2).astype(np.float32)
        K[rng.choice(len(verts), 9, replace=False)] = np.nan
        H[rng.choice(len(verts), 7, replace=False)] = np.nan
        ns["convert_pv_to_o3d"] = lambda mesh: _FakeO3dMesh(mesh.vertices, mesh.triangles)
        e = ns["load_mesh_compute_energies"](_FakePvMesh(verts, tris, {"gaussian_curvature": K, "mean_curvature": H}))
        e0 = ns["load_mesh_compute_energies"](_FakePvMesh(verts, tris, {}))
        out.update(energy_vertices=verts, energy_triangles=tris, energy_K=K, energy_H=H,
                   energy_result=np.array(e, np.float64), energy_result_no_curvature=np.array(e0, np.float64))
    # --- PCA estimator of the PointCloud class, on a seeded subsample of bunny (the reference is O(N^2))
    ref = import_reference()
    bunny = np.loadtxt(os.path.join(REFERENCE_ROOT, "sample_scans", "bunny.txt")).astype(np.float32)
    sub = bunny[np.sort(np.random.default_rng(4).choice(len(bunny), 3000, replace=False))]
    with contextlib.redirect_stdout(io.StringIO()):
        pc = ref.PointCloud(points=sub, normals=np.zeros((len(sub), 0), np.float32), k_neighbors=12)
        pc.principal_curvatures_via_principal_component_analysis(12)
    out.update(pca_points=sub, pca_k=np.int64(12), pca_l1=pc.pca_principal_curvature_values_1,
               pca_l2=pc.pca_principal_curvature_values_2, pca_K=pc.pca_K_values, pca_H=pc.pca_H_values,
               pca_directions=pc.principal_curvature_directions)
    path = os.path.join(GOLDEN_DIR, "io_energy_pca.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
