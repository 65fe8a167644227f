"""Opt-in stage timing (the reference's only profiling hook is cProfile around main(),
/root/reference/main_scans.py:69-73).  Enable with PCT_B200_TRACE=1 or trace.enable():
every stage is bracketed by a device synchronise, so the numbers are attributable but the
pipeline no longer overlaps -- never leave it on for a benchmark."""
from __future__ import annotations

import contextlib
import os
import time

_enabled = os.environ.get("PCT_B200_TRACE", "0") not in ("", "0")
timings: dict[str, float] = {}


def enable(on: bool = True):
    global _enabled
    _enabled = on


def reset():
    timings.clear()


@contextlib.contextmanager
def stage(name: str):
    if not _enabled:
        yield
        return
    import torch

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        yield
    finally:
        torch.cuda.synchronize()
        timings[name] = timings.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
