import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from point_cloud_toolbox_b200 import engine
from scripts.tune import torus

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
pts = torus(n)
ix = engine.GridIndex(pts, k_hint=20)
keep = []
for rep in range(6):
    fit = ix.curvature_knn(20, want_coeffs=False)
    if rep % 2 == 0:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    src = fit.curv[:, :2].t().contiguous()
    t1 = time.perf_counter()
    if rep >= 4:
        torch.cuda.synchronize()
    t1b = time.perf_counter()
    h = engine._PINNED.take(tuple(src.shape), src.dtype)
    t2 = time.perf_counter()
    h.copy_(src, non_blocking=True)
    t3 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t4 = time.perf_counter()
    a = h.numpy()
    keep = [a]
    print(rep, f"presync={rep % 2 == 0} contiguous={1e3*(t1-t0):.1f} sync2={1e3*(t1b-t1):.1f} take={1e3*(t2-t1b):.1f} copy_call={1e3*(t3-t2):.1f} wait={1e3*(t4-t3):.1f} pinned={h.is_pinned()} entries={len(engine._PINNED.entries)}", flush=True)
# plain torch D2H for comparison
x = torch.empty(n * 2, dtype=torch.float32, device="cuda")
hp = torch.empty(n * 2, dtype=torch.float32, pin_memory=True)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); hp.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("plain d2h ms", 1e3*(time.perf_counter()-t0))
