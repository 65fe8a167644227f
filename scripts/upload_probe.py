"""Pageable host array -> device: torch's copy against pct_upload (staged by several host threads).
    python scripts/upload_probe.py [points]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from point_cloud_toolbox_b200 import engine  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
a = np.random.default_rng(0).random((n, 3), dtype=np.float32)          # pageable
t = torch.from_numpy(a)
for name, fn in (("torch .to()", lambda: t.to("cuda", non_blocking=True)), ("pct_upload", lambda: engine.to_device_points(a))):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    ok = bool(torch.equal(d.cpu(), t))
    print(f"{name}: {best * 1e3:.1f} ms, {a.nbytes / best / 1e9:.1f} GB/s, equal={ok}", flush=True)
pin = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
pin.copy_(t)
torch.cuda.synchronize(); t0 = time.perf_counter(); d = engine.to_device_points(pin); torch.cuda.synchronize()
print(f"pinned source: {(time.perf_counter() - t0) * 1e3:.1f} ms")
