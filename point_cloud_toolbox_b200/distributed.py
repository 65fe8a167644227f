"""Multi-GPU form of the path: replicate the cloud, shard the queries, gather the results.

One process per GPU (torchrun), ``torch.distributed`` for the plumbing:

  1. rank 0 holds the cloud; ``broadcast`` puts the raw xyz on every GPU (NVLink / NVSwitch)
  2. every rank builds the same index (the build is deterministic)
  3. rank g runs the fused kernel on Morton-sorted positions [g*N/G, (g+1)*N/G) -- a spatially
     coherent slice -- writing slice-local rows (PCT_LAYOUT_SLICE)
  4. ``gather`` to rank 0, which undoes the Morton permutation

There is no exchange step between 2 and 4, so no other collective is involved.
The reference has no distributed code at all; this is new.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Sorted-position range of one rank: contiguous, sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def padded_rows(n: int, world: int) -> int:
    return (n + world - 1) // world


def gather_rows(local: torch.Tensor, n: int, group=None, dst: int = 0):
    """Gather slice-local rows (rank order = sorted order) on ``dst``; returns (n, ...) there, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = padded_rows(n, world)
    if local.shape[0] != rows:
        pad = torch.zeros((rows - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat((local, pad), 0)
    local = local.contiguous()
    if rank == dst:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, parts, dst=dst, group=group)
        keep = []
        for r, p in enumerate(parts):
            b, e = shard_bounds(n, world, r)
            keep.append(p[: e - b])
        return torch.cat(keep, 0)
    dist.gather(local, None, dst=dst, group=group)
    return None


def unpermute(sorted_rows: torch.Tensor, perm: torch.Tensor):
    """rows in sorted order -> rows in original order (perm[sorted position] = original index)."""
    out = torch.empty_like(sorted_rows)
    out[perm.long()] = sorted_rows
    return out


def broadcast_cloud(points_dev, n: int, group=None, src: int = 0, device=None):
    """(N, 3) float32 on every rank; ``points_dev`` is only read on ``src``."""
    if dist.get_rank(group) == src:
        buf = points_dev.contiguous()
    else:
        buf = torch.empty((n, 3), dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src, group=group)
    return buf


def curvature_knn_sharded(points, n: int, k: int, group=None, device=None, columns=(0, 1)):
    """plant_kdtree(k) + compute_pointwise_explicit_quadratic_curvature() over all ranks of ``group``.

    ``points``: host or device (N, 3) float32 on rank 0, ignored elsewhere.
    Returns on rank 0 a (N, len(columns)) device tensor of curvature columns
    (0 = K, 1 = H, 2 = k1, 3 = k2, 4 = H^2) in ORIGINAL point order; None on other ranks.
    """
    from . import engine
    from ._lib import LAYOUT_SLICE

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    d_points = engine.to_device_points(points, device) if rank == 0 else None
    cloud = broadcast_cloud(d_points, n, group, 0, device)
    index = engine.GridIndex(cloud, k_hint=k)
    begin, end = shard_bounds(n, world, rank)
    fit = index.curvature_knn(k, begin, end, layout=LAYOUT_SLICE, want_coeffs=False)
    names = ("K", "H", "k1", "k2", "H2")
    local = torch.stack([fit.column(names[c]) for c in columns], 1)
    gathered = gather_rows(local, n, group, 0)
    if rank != 0:
        return None
    return unpermute(gathered, index.permutation())
