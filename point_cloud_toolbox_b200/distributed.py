"""Multi-GPU form of the path: replicate the cloud, shard the queries, gather the results.

One process per GPU (torchrun), ``torch.distributed`` for the plumbing:

  1. rank 0 holds the cloud; ``broadcast`` puts the raw xyz on every GPU (NVLink / NVSwitch)
  2. the queries are divided
       "slab" (default)   by position: the ranks agree on cut planes across the longest axis of the
                          bounding box (quantiles, so the slabs hold equal numbers of points); rank g
                          builds an index over ITS slab plus a margin of 4.5 cells only and answers the
                          points of its slab.  The index knows where its knowledge ends
                          (pct_index_set_slab): no search radius crosses the margin, and a query whose
                          k-th neighbour could lie beyond it comes back PCT_STATUS_UNRESOLVED and is
                          answered from a whole-cloud index built on demand.  The build, the serial
                          fraction of the replicated form, shrinks with the number of ranks.
       "replicated"       (kept for comparison) every rank builds the same whole-cloud index; rank g answers Morton-sorted
                          positions [g*N/G, (g+1)*N/G) in slice layout (PCT_LAYOUT_SLICE)
  3. ``gather`` to rank 0, which puts the rows in original order

There is no exchange step between 1 and 3, so no other collective is involved.
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


# ---------------------------------------------------------------------------
# slabs
# ---------------------------------------------------------------------------
SLAB_MARGIN_CELLS = 4.5  # level-0 and level-1 searches (radius <= 3 cells) never reach the margin's end


def slab_cuts(axis_coords: torch.Tensor, world: int, sample: int = 1 << 20):
    """world + 1 non-decreasing cut values (first -inf, last +inf): quantiles of a strided sample.
    A pure function of the data, so every rank computes the same cuts from its replica."""
    n = int(axis_coords.numel())
    step = max(1, n // sample)
    s = axis_coords[::step].to(torch.float32).sort().values
    m = int(s.numel())
    inner = [float(s[min(m - 1, (m * g) // world)]) for g in range(1, world)]
    return [float("-inf")] + inner + [float("inf")]


def slab_bounds(cuts, rank: int, margin: float):
    """(complete_lo, complete_hi, own_lo, own_hi) of one rank as float32 values."""
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731
    own_lo, own_hi = f32(cuts[rank]), f32(cuts[rank + 1])
    m = f32(margin)
    return f32(own_lo - m), f32(own_hi + m), own_lo, own_hi


def slab_select(axis_coords: torch.Tensor, bounds):
    """(sel, own): ``sel`` = original indices of the points a slab index is built from, ASCENDING (the
    index breaks distance ties by position in its input, which therefore stays the whole cloud's order);
    ``own`` = boolean mask over ``sel`` of the points the slab owns."""
    c_lo, c_hi, own_lo, own_hi = bounds
    sel = ((axis_coords >= c_lo) & (axis_coords <= c_hi)).nonzero().squeeze(1)
    xs = axis_coords.index_select(0, sel)
    return sel, (xs >= own_lo) & (xs < own_hi)


def gather_scattered(ids: torch.Tensor, rows: torch.Tensor, n: int, group=None, dst: int = 0):
    """Ranks hold rows for disjoint sets of original indices; returns the (n, ...) array on ``dst``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    count = torch.tensor([int(ids.numel())], dtype=torch.int64, device=ids.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c) for c in counts]
    width = max(counts) if counts else 0
    ids_p = torch.zeros((width,), dtype=torch.int64, device=ids.device)
    ids_p[: ids.numel()] = ids.to(torch.int64)
    rows_p = torch.zeros((width,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    rows_p[: rows.shape[0]] = rows
    if rank == dst:
        id_parts = [torch.empty_like(ids_p) for _ in range(world)]
        row_parts = [torch.empty_like(rows_p) for _ in range(world)]
        dist.gather(ids_p, id_parts, dst=dst, group=group)
        dist.gather(rows_p, row_parts, dst=dst, group=group)
        out = torch.empty((n,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        for r in range(world):
            out[id_parts[r][: counts[r]]] = row_parts[r][: counts[r]]
        return out
    dist.gather(ids_p, None, dst=dst, group=group)
    dist.gather(rows_p, None, dst=dst, group=group)
    return None


def default_mode(world: int) -> str:
    """Slabs replace the replicated whole-cloud build (10 ms at 100 M points) by the cut / select / smaller-build
    work of a rank (6.5 ms at 2 ranks, 2.7 ms at 8); measured faster from 2 ranks up (2 GPUs: 36.9 vs 38.1 ms)."""
    import os

    forced = os.environ.get("PCT_MULTI_MODE")  # experiments
    if forced in ("slab", "replicated"):
        return forced
    return "slab" if world >= 2 else "replicated"


class SlabFit:
    """Result of one rank: packed records of the points it owns (``ids`` = their original indices)."""

    def __init__(self, ids, records, unresolved, index, cell_size, bounds, axis):
        self.ids, self.records, self.unresolved = ids, records, unresolved
        self.index, self.cell_size, self.bounds, self.axis = index, cell_size, bounds, axis


def curvature_knn_slab(cloud: torch.Tensor, k: int, rank: int, world: int, events=()) -> SlabFit:
    """The work of one rank on a replicated device cloud: cuts, slab index, fused kernel on its own points.
    ``events``: optional CUDA event recorded once the slab index is built (for stage timing)."""
    from . import engine
    from ._lib import STATUS_UNRESOLVED

    h, lo, hi = engine.estimate_cell_size(cloud, k)
    axis = max(range(3), key=lambda a: hi[a] - lo[a])
    x = cloud[:, axis]
    bounds = slab_bounds(slab_cuts(x, world), rank, SLAB_MARGIN_CELLS * h)
    sel, local, row_map, n_own = engine.slab_select(cloud, axis, bounds)
    if int(sel.numel()) <= k + 1:
        # degenerate slab (tiny cloud): index the whole cloud, still answer only what this rank owns
        bounds = (float("-inf"), float("inf"), bounds[2], bounds[3])
        sel, local, row_map, n_own = engine.slab_select(cloud, axis, bounds)
    if n_own == 0:
        empty = torch.empty((0,), dtype=torch.int64, device=cloud.device)
        return SlabFit(empty, torch.empty((0, 8), dtype=torch.float32, device=cloud.device), 0, None, h, bounds, axis)
    index = engine.GridIndex(local, cell_hint=h, k_hint=k)
    index.set_slab(axis, *bounds, row_map=row_map, mapped_rows=n_own)  # rows of points it does not own are never written
    for ev in events:
        ev.record()
    records = index.curvature_knn(k, want_coeffs=False).records   # (n_own, 8)
    # original indices of the owned points, ascending: row_map steps by one exactly at an owned point
    own = torch.ones_like(row_map, dtype=torch.bool)
    own[1:] = row_map[1:] != row_map[:-1]
    own[0] = bool(row_map[0] == 0)
    ids = sel[own].to(torch.int64)
    n_bad = int(index.last_stats().unresolved)
    if n_bad:
        bad = (records[:, 7].contiguous().view(torch.int32) & STATUS_UNRESOLVED) != 0
        # the k-th neighbour may lie outside the margin: answer these from a whole-cloud index
        whole = engine.GridIndex(cloud, cell_hint=h, k_hint=k)
        redo = whole.curvature_points(ids[bad].to(torch.int32), k)
        records[bad] = redo.records
        whole.close()
    return SlabFit(ids, records, n_bad, index, h, bounds, axis)


# ---------------------------------------------------------------------------
# host buffers shared by the ranks: every rank moves its own share over its own PCIe link
# ---------------------------------------------------------------------------
class SharedHostArray:
    """A float32 array in POSIX shared memory that every rank of the node maps and page-locks.

    The processes of a one-node job usually see the same input anyway (the same file, memory-mapped);
    with the cloud and the result in shared host memory each rank copies only ITS share in and out,
    over its own PCIe link, instead of rank 0 moving everything.  ``create=True`` on exactly one rank
    (which also unlinks the segment on ``close``), ``create=False`` on the others after a barrier.
    """

    def __init__(self, name: str, shape, create: bool):
        import numpy as np
        from multiprocessing import shared_memory

        nbytes = 4
        for d in shape:
            nbytes *= int(d)
        try:
            self._shm = shared_memory.SharedMemory(name=name, create=create, size=max(nbytes, 4))
        except FileExistsError:
            # a segment of that name left behind by a run that did not finish: replace it
            stale = shared_memory.SharedMemory(name=name, create=False)
            stale.close()
            stale.unlink()
            self._shm = shared_memory.SharedMemory(name=name, create=True, size=max(nbytes, 4))
        self._owner = create
        if not create:
            # Python < 3.13 registers attached segments for unlinking at exit as well; only the creator unlinks
            try:
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.array = np.ndarray(tuple(shape), dtype=np.float32, buffer=self._shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self._registered = False
        if torch.cuda.is_available() and nbytes:
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            self._registered = int(rc) == 0

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = None
        self.array = None
        try:
            self._shm.close()
            if self._owner:
                self._shm.unlink()
        except (BufferError, FileNotFoundError):
            pass


def shared_cloud_from_text(path, name: str, group=None) -> SharedHostArray:
    """The scan file of ``PointCloud(file_path)`` as the (N, 3) shared host array ``curvature_knn_shared`` reads.

    Rank 0 parses the text with the library's loader (all host threads) straight into the segment, applies the
    reference's shift of x and y by their maxima in float32 (ref :56-57), and the other ranks map the segment
    afterwards.  Collective over ``group``; the caller closes the array (rank 0 unlinks it)."""
    import ctypes

    import numpy as np

    from ._lib import check, lib

    rank = dist.get_rank(group)
    shape = torch.zeros(2, dtype=torch.int64)
    if rank == 0:
        rows, cols = ctypes.c_int64(), ctypes.c_int64()
        check(lib.pct_text_shape(str(path).encode(), ctypes.byref(rows), ctypes.byref(cols)))
        shape[0], shape[1] = rows.value, cols.value
    backend = dist.get_backend(group)
    if backend == "nccl":
        dev_shape = shape.cuda()
        dist.broadcast(dev_shape, src=0, group=group)
        shape = dev_shape.cpu()
    else:
        dist.broadcast(shape, src=0, group=group)
    n, cols = int(shape[0]), int(shape[1])
    if n < 2 or cols < 3:
        raise ValueError(f"{path}: {n} rows of {cols} columns is not a point cloud")
    shared = None
    if rank == 0:
        shared = SharedHostArray(name, (n, 3), create=True)
        if cols == 3:
            check(lib.pct_text_load_f32(str(path).encode(), n, 3, ctypes.c_void_p(shared.tensor.data_ptr()), 0))
        else:
            table = np.empty((n, cols), np.float32)
            check(lib.pct_text_load_f32(str(path).encode(), n, cols, table.ctypes.data_as(ctypes.c_void_p), 0))
            shared.array[:] = table[:, 0:3]                                  # ref :52
        shared.array[:, 0] -= shared.array[:, 0].max()                       # ref :56 (fp32)
        shared.array[:, 1] -= shared.array[:, 1].max()                       # ref :57
    dist.barrier(group=group)
    if rank != 0:
        shared = SharedHostArray(name, (n, 3), create=False)
    dist.barrier(group=group)
    return shared


def exchange_by_owner(ids: torch.Tensor, rows: torch.Tensor, n: int, group=None):
    """All-to-all of per-point rows to the ranks that own their ORIGINAL index ranges.

    ``ids`` ascending original indices this rank computed, ``rows`` their rows.  Rank r owns original
    indices ``shard_bounds(n, world, r)``; returns that range filled, ``(end - begin, ...)``.
    Ascending ids make every destination a contiguous segment, so no sort is needed."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    edges = torch.tensor([shard_bounds(n, world, r)[0] for r in range(world)] + [n], dtype=ids.dtype, device=ids.device)
    cuts = torch.searchsorted(ids.contiguous(), edges)
    send = (cuts[1:] - cuts[:-1]).to(torch.int64)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    send_l, recv_l = [int(v) for v in send.tolist()], [int(v) for v in recv.tolist()]
    total = sum(recv_l)
    ids_in = torch.empty((total,), dtype=ids.dtype, device=ids.device)
    rows_in = torch.empty((total,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_to_all_single(ids_in, ids.contiguous(), recv_l, send_l, group=group)
    dist.all_to_all_single(rows_in, rows.contiguous(), recv_l, send_l, group=group)
    begin, end = shard_bounds(n, world, rank)
    out = torch.empty((end - begin,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    out[(ids_in - begin).long()] = rows_in
    return out


def curvature_knn_shared(points: SharedHostArray, out: SharedHostArray, k: int, group=None, device=None):
    """plant_kdtree(k) + compute_pointwise_explicit_quadratic_curvature() of a cloud in shared host memory.

    ``points`` (N, 3) and ``out`` (2, N) = [K; H] are mapped by every rank.  Rank r copies rows
    ``shard_bounds(N, world, r)`` of the cloud to its GPU, an all-gather over NVLink replicates the cloud,
    every rank answers its slab, an all-to-all returns the rows to the ranks owning their original index
    ranges, and each rank writes its range of ``out``.  Collective: returns after a barrier, when
    ``out`` is complete on the host."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    n = int(points.tensor.shape[0])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    rows = padded_rows(n, world)
    begin, end = shard_bounds(n, world, rank)
    padded = torch.empty((world * rows, 3), dtype=torch.float32, device=device)
    mine = padded[rank * rows: rank * rows + (end - begin)]
    mine.copy_(points.tensor[begin:end], non_blocking=True)                      # this rank's share, its own PCIe link
    dist.all_gather_into_tensor(padded, padded[rank * rows:(rank + 1) * rows], group=group)
    if n == world * rows:
        cloud = padded
    else:
        cloud = torch.cat([padded[r * rows: r * rows + (shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0])]
                           for r in range(world)], 0)
    part = curvature_knn_slab(cloud, k, rank, world)
    kh = part.records[:, 3:5].contiguous()
    own = exchange_by_owner(part.ids.to(torch.int32), kh, n, group)                              # (end - begin, 2) in original order
    khT = own.t().contiguous()
    out.tensor[0, begin:end].copy_(khT[0], non_blocking=True)
    out.tensor[1, begin:end].copy_(khT[1], non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    if part.index is not None:
        part.index.close()
    dist.barrier(group=group)
    return part


def curvature_knn_sharded(points, n: int, k: int, group=None, device=None, columns=(0, 1), mode="auto"):
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
    names = ("K", "H", "k1", "k2", "H2")
    if mode == "auto":
        mode = default_mode(world)
    if mode == "slab":
        part = curvature_knn_slab(cloud, k, rank, world)
        fit = engine.FitOutputs(records=part.records)
        local = torch.stack([fit.column(names[c]) for c in columns], 1)
        out = gather_scattered(part.ids, local, n, group, 0)
        if part.index is not None:
            part.index.close()
        return out
    index = engine.GridIndex(cloud, k_hint=k)
    begin, end = shard_bounds(n, world, rank)
    fit = index.curvature_knn(k, begin, end, layout=LAYOUT_SLICE, want_coeffs=False)
    local = torch.stack([fit.column(names[c]) for c in columns], 1)
    gathered = gather_rows(local, n, group, 0)
    if rank != 0:
        return None
    return unpermute(gathered, index.permutation())
