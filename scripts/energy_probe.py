"""Mesh energy kernel alone: `python scripts/energy_probe.py [grid_n] [shuffle]` -> ms per call (cold L2) and GB/s;
`shuffle` renumbers the vertices at random (a mesh without locality)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_cloud_toolbox_b200 import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
shuffle = len(sys.argv) > 2 and sys.argv[2] == "shuffle"
g = torch.Generator(device="cuda").manual_seed(1)
u, v = torch.meshgrid(torch.linspace(-6, 6, n, device="cuda"), torch.linspace(-6, 6, n, device="cuda"), indexing="ij")
verts = torch.stack([u, v, torch.sin(u) * torch.sin(v)], -1).reshape(-1, 3).float().contiguous()
i, j = torch.meshgrid(torch.arange(n - 1, device="cuda"), torch.arange(n - 1, device="cuda"), indexing="ij")
a = (i * n + j).reshape(-1)
tris = torch.cat([torch.stack([a, a + 1, a + n], 1), torch.stack([a + 1, a + n + 1, a + n], 1)]).int().contiguous()
if shuffle:  # a mesh whose vertex numbering has no locality
    perm = torch.randperm(n * n, device="cuda", generator=g)
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n * n, device="cuda")
    verts = verts[perm].contiguous(); tris = inv[tris.long()].int().contiguous()
K = torch.randn(n * n, device="cuda", generator=g)
H = torch.randn(n * n, device="cuda", generator=g)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
engine.mesh_energies(verts, tris, K, H); torch.cuda.synchronize()
ts = []
for _ in range(7):
    flush.fill_(1)
    ev[0].record(); out = engine.mesh_energies(verts, tris, K, H); ev[1].record(); torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]))
ms = float(np.median(ts))
T, V = tris.shape[0], verts.shape[0]
print(json.dumps({"triangles": T, "vertices": V, "shuffled": shuffle,
                  "ms": ms, "GBps_indices": T * 12 / ms / 1e6, "GBps_indices_and_vertex_arrays": (T * 12 + V * 20) / ms / 1e6,
                  "out": [float(x) for x in out.cpu()]}))
