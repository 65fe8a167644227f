"""GPU debug: implicit quadric fit against numpy on the stored neighbourhoods."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from conftest import load_golden
import point_cloud_toolbox_b200 as pct
from point_cloud_toolbox_b200 import engine
g = load_golden("implicit")
for nb in g["neighbourhoods"][:2]:
    p = nb.astype(np.float32)
    A = np.column_stack((p[:, 0] ** 2, p[:, 1] ** 2, p[:, 2] ** 2, p[:, 0] * p[:, 1], p[:, 0] * p[:, 2], p[:, 1] * p[:, 2], p[:, 0], p[:, 1], p[:, 2], np.ones(len(p))))
    w, v = np.linalg.eigh(A.T @ A)
    c = pct.PointCloud.fit_implicit_quadric_surface(p)
    print("gpu c", c)
    print("np  v", v[:, 0])
    print("obj gpu", np.sum((A @ c) ** 2), "lam", w[:3], "overlaps", np.abs(v.T @ c).round(3))
    # same rows through the index form
    pts = np.concatenate((np.zeros((1, 3), np.float32), p[1:]))
    d = torch.from_numpy(pts).cuda()
    rows = torch.arange(len(pts), dtype=torch.int32, device="cuda")[None]
    c2 = engine.implicit_quadric_fit(d, rows, query_ids=torch.zeros(1, dtype=torch.int32, device="cuda")).cpu().numpy()[0]
    print("obj rows", np.sum((A @ c2) ** 2))
