"""Multi-GPU forms of the path.  One process per GPU (torchrun), ``torch.distributed`` for the plumbing.

The throughput form is the SLAB EXCHANGE (``curvature_knn_exchange``, ``curvature_knn_shared``): every rank starts with
a contiguous share of the cloud and the cloud is never replicated.

  1. plan      one small all-gather (bounding boxes + a strided sample): every rank derives the same cell edge, slab
               axis and cut planes (quantiles: slabs hold equal numbers of points)
  2. bin       every rank bins its share by destination slab -- the owner, and the neighbours whose margin of 4.5
               cells holds the point
  3. exchange  INSIDE the binning kernel: the records {x, y, z, original index} are stored straight into the slab
               buffer of the destination rank over NVLink peer memory (``PeerResults``, CUDA IPC); the received slab
               is in ascending original index, so distance ties keep the whole cloud's order
  4. answer    slab index + fused kernel.  The index knows where its knowledge ends (pct_index_set_slab): no search
               radius crosses the margin, a query whose k-th neighbour could lie beyond it comes back
               PCT_STATUS_UNRESOLVED and is redone on a whole-cloud index (only then are the shares all-gathered)
  5. return    INSIDE the fused kernel: K, H of every answered query are stored straight into the result array of
               the rank that holds the query's point
  Without peer mapping (other backends, processes that cannot share device memory) steps 3 and 5 are one
  all-to-all each.  ``curvature_knn_shared`` wraps this for host arrays in shared memory: every rank moves its own
  share over its own PCIe link, in as many rounds as ``CopyRounds`` measures to be fastest on the box.

Kept beside it: ``curvature_knn_slab`` / ``curvature_knn_sharded`` -- round 1's forms, which broadcast the whole cloud
to every GPU and either let rank g select and answer its slab ("slab") or build the same whole-cloud index everywhere
and answer Morton slices ("replicated"), then gather to rank 0.  ``bench.py`` reports the former as the comparison
figure ``value_replicated_no_collective``.

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
    picks = torch.tensor([min(m - 1, (m * g) // world) for g in range(1, world)], dtype=torch.int64, device=s.device)
    inner = [float(v) for v in s[picks].tolist()] if world > 1 else []      # one host transfer for all cuts
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

    Page placement: a page of the segment lands on the NUMA node of the thread that first touches it, and page-locking
    touches every page.  With ``register=False`` nothing is touched: every rank then writes (``first_touch``) the rows
    it will move, from a thread bound near its GPU (``bind_near_gpu``), and page-locks afterwards (``register``) --
    ``shared_arrays_placed`` does exactly that.  On a two-socket host the copies of a rank then stay on its socket.
    """

    def __init__(self, name: str, shape, create: bool, register: bool = True):
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
        self._nbytes = nbytes
        if register:
            self.register()

    def first_touch(self, index):
        """Zero ``array[index]`` from the calling thread: its pages are allocated on that thread's NUMA node."""
        self.array[index] = 0.0

    def register(self):
        """Page-lock the whole segment for this process's CUDA context (idempotent)."""
        if not self._registered and torch.cuda.is_available() and self._nbytes:
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self._nbytes, 0)
            self._registered = int(rc) == 0
        return self._registered

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


def device_for_rank(local_rank: int, local_world: int, visible: int = None) -> int:
    """GPU index of a rank of a one-node job: the ranks are spread over the visible devices (rank r takes device
    r * (visible // local_world)) so that a job smaller than the box does not crowd the GPUs of one PCIe uplink --
    on an HGX board neighbouring indices share a switch towards the host (profiles/pcie_probe_8gpu_r02n.txt: two
    neighbours together move 72 GB/s to the host, two GPUs of different halves 106 GB/s).  One rank per GPU either way."""
    if visible is None:
        visible = torch.cuda.device_count()
    stride = visible // local_world if local_world >= 1 and visible >= 2 * local_world else 1
    return int(local_rank) * max(1, stride)


def gpu_local_cpus(device_index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs ``local_cpulist`` of its PCI function), or None."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/local_cpulist"
        with open(path) as f:
            text = f.read().strip()
    except (OSError, AttributeError, RuntimeError):
        return None
    cpus = set()
    for part in text.split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus or None


def bind_near_gpu(device_index: int):
    """Restrict the calling process to the CPUs next to its GPU (intersected with what it may already use), so that
    the host pages it touches and its staging threads sit on the GPU's socket.  Returns the CPU set, or None when the
    topology is not visible (containers without sysfs, one-socket hosts report everything: harmless)."""
    import os

    cpus = gpu_local_cpus(device_index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = os.sched_getaffinity(0) & cpus
    if not allowed:
        return None
    os.sched_setaffinity(0, allowed)
    return allowed


def shared_arrays_placed(name: str, shapes, n: int, group=None, row_axis=None):
    """Collective.  One shared host array per entry of ``shapes``, each with an axis of length ``n`` (``row_axis[i]``,
    default: the first axis of that length) along which the ranks split the cloud: rank r first-touches its
    ``shard_bounds(n, world, r)`` part of every array, then all ranks page-lock.  Returns the list of arrays
    (zero-filled); rank 0 owns the segments."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    arrays = []
    if rank == 0:
        arrays = [SharedHostArray(f"{name}_{i}", shp, create=True, register=False) for i, shp in enumerate(shapes)]
    dist.barrier(group=group)
    if rank != 0:
        arrays = [SharedHostArray(f"{name}_{i}", shp, create=False, register=False) for i, shp in enumerate(shapes)]
    begin, end = shard_bounds(n, world, rank)
    for i, (arr, shp) in enumerate(zip(arrays, shapes)):
        axis = row_axis[i] if row_axis is not None else list(shp).index(n)
        index = [slice(None)] * len(shp)
        index[axis] = slice(begin, end)
        arr.first_touch(tuple(index))
    dist.barrier(group=group)
    for arr in arrays:
        arr.register()
    dist.barrier(group=group)
    return arrays


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


# ---------------------------------------------------------------------------
# slab exchange: the cloud is never replicated
# ---------------------------------------------------------------------------
PLAN_SAMPLE = 1 << 18  # rows of the sample the ranks agree on cell size and cut planes from


class SlabPlan:
    """What every rank derives, identically, from the gathered sample: cell edge, slab axis, bounds of all slabs."""

    def __init__(self, h, axis, cuts, bounds, bbox):
        self.h, self.axis, self.cuts, self.bounds, self.bbox = h, axis, cuts, bounds, bbox


class Stages:
    """Optional stage timer of the exchange path: CUDA events on the current stream between named stages."""

    def __init__(self, enabled=True):
        self.enabled = enabled and torch.cuda.is_available()
        self.marks = []

    def mark(self, name):
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def durations_ms(self):
        """{stage name: ms} -- the time between the previous mark and this one (call after a synchronize)."""
        out = {}
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


def plan_slabs(share: torch.Tensor, n_total: int, k: int, group=None, cell_fn=None) -> SlabPlan:
    """Collective.  Every rank contributes the bounding box of its share and a strided sample of it in ONE
    all-gather; all ranks then compute the same cell edge (density pilot on the sample), the same slab axis
    (longest box axis) and the same cut planes (quantiles of the sample, so slabs hold equal numbers of points)."""
    from . import engine

    world = dist.get_world_size(group)
    per = max(1, min(-(-PLAN_SAMPLE // world), padded_rows(n_total, world)))
    n = int(share.shape[0])
    payload = torch.full((per + 2, 3), float("nan"), dtype=torch.float32, device=share.device)
    if n:
        xyz = share[:, :3]
        payload[0] = xyz.min(0).values               # NaN propagates: the receivers see a non-finite box
        payload[1] = xyz.max(0).values
        step = max(1, n // per)
        smp = xyz[::step][:per]
        payload[2:2 + smp.shape[0]] = smp
    parts = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(parts, payload, group=group)
    allp = torch.stack(parts)                                   # (world, per + 2, 3)
    lo = torch.nan_to_num(allp[:, 0], nan=float("inf")).min(0).values
    hi = torch.nan_to_num(allp[:, 1], nan=float("-inf")).max(0).values
    sample = allp[:, 2:].reshape(-1, 3)
    sample = sample[~torch.isnan(sample[:, 0])].contiguous()
    # (a NaN / Inf coordinate anywhere in a share shows in that share's box; one host transfer for box + flag)
    head = torch.cat((lo, hi, torch.isfinite(allp[:, :2]).all().reshape(1).to(lo.dtype))).tolist()
    bbox = [float(v) for v in head[:6]]
    if not head[6] or not all(abs(v) <= 3.0e38 for v in bbox):
        raise ValueError("Non-finite values in input points")    # ref :273-274
    h = (cell_fn or engine.estimate_cell_size_sample)(sample, n_total, bbox, k)
    axis = max(range(3), key=lambda a: bbox[3 + a] - bbox[a])
    cuts = slab_cuts(sample[:, axis], world, sample=1 << 62)
    bounds = [slab_bounds(cuts, r, SLAB_MARGIN_CELLS * h) for r in range(world)]
    return SlabPlan(h, axis, cuts, bounds, bbox)


def _all_to_all_rows(send: torch.Tensor, send_counts, recv_counts, group=None):
    """all_to_all_single of rows with per-rank row counts (lists of ints)."""
    out = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(out, send.contiguous(), list(recv_counts), list(send_counts), group=group)
    return out


class ExchangeFit:
    """Result of one rank of the exchange path.

    ``rows``: (n_share, C) result columns of this rank's OWN share of the cloud, share order;
    ``index`` / ``local_ids`` / ``records`` / ``own_ids``: the slab this rank answered -- its index, the original index
    of every indexed point, the packed records of the points it owns and their original indices (ascending).
    When the exchanges run over peer memory (``peer_return``), ``slab`` is a view of this rank's persistent slab buffer:
    it (and ``local_ids``) is valid until the next exchange call on the group; ``rows`` is always a private copy."""

    def __init__(self, rows, plan, index, slab, rank, records, unresolved):
        self.rows, self.plan, self.index, self.slab, self.rank = rows, plan, index, slab, rank
        self.records, self.unresolved, self.indexed = records, unresolved, int(slab.shape[0])
        self._own_ids = None

    @property
    def local_ids(self):
        """original index of every point of the slab cloud (int32)"""
        return self.slab[:, 3].contiguous().view(torch.int32)

    @property
    def own_ids(self):
        """original indices of the points this rank answered, ascending (= the rows of ``records``)"""
        if self._own_ids is None:
            self._own_ids = _owned_ids(self.slab, self.plan, self.rank)
        return self._own_ids

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None


class PeerResults:
    """Two arrays per rank of a one-node job, mapped into every rank (CUDA IPC; NVLink peer memory), so that the two
    exchanges of the slab path happen INSIDE kernels instead of after them:

    * ``local`` / ``ptrs``: the K, H array -- ``padded_rows(n_total, world)`` rows of {K, H}; row i belongs to original
      index ``shard_bounds(n_total, world, r)[0] + i``.  The fused kernel of every rank stores results straight into the
      array of the rank that holds the point (``GridIndex.set_peers``): no return all-to-all.
    * ``slab_local`` / ``slab_ptrs``: the slab buffer -- ``slab_cap`` rows of {x, y, z, original index}.  The binning
      kernel of every rank stores its records straight into the buffer of the destination slab
      (``SlabBinCount.fill_peers``): no all-to-all of the points.

    Built once per (group, device, cloud size) and reused: opening IPC handles costs milliseconds.  A job that tears its
    process group down and builds another calls ``PeerResults.release()`` first (collective): the mappings belong to the
    processes of the group they were made in."""

    _cache = {}

    def __init__(self, n_total: int, group, device):
        from torch.multiprocessing.reductions import reduce_tensor

        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        self.rows = padded_rows(n_total, world)
        self.local = torch.zeros((max(self.rows, 1), 2), dtype=torch.float32, device=device)
        # a slab = the points it owns (about 1 / world of the cloud) + its margins; clouds whose slabs outgrow the
        # buffer take the all-to-all instead
        self.slab_cap = int(min(n_total, self.rows + self.rows // 2 + 65536))
        self.slab_local = torch.zeros((max(self.slab_cap, 1), 4), dtype=torch.float32, device=device)
        # every rank takes part in the one collective below whatever happens to it: a rank that cannot export its
        # arrays sends None, a rank that cannot map a peer's remembers why -- ``get`` then lets the ranks agree
        self.error = None
        try:
            # one export per consumer: torch counts the references of an exported block per export, and every peer
            # releases the one it received
            mine = [(reduce_tensor(self.local), reduce_tensor(self.slab_local)) if r != rank else None for r in range(world)]
        except Exception as exc:      # e.g. an allocator configuration whose blocks cannot be exported
            mine, self.error = None, exc
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        here = device.index if device.index is not None else torch.cuda.current_device()
        self.views, self.slab_views = [], []
        try:
            for r, per_consumer in enumerate(handles):
                if r == rank:
                    self.views.append(self.local)
                    self.slab_views.append(self.slab_local)
                    continue
                if per_consumer is None:
                    raise RuntimeError(f"rank {r} could not export its arrays")
                for (fn, args), views in zip(per_consumer[rank], (self.views, self.slab_views)):
                    args = list(args)
                    args[6] = here                                   # storage_device: map the peer's memory into THIS device
                    views.append(fn(*args))
        except Exception as exc:
            self.error = self.error or exc
            self.views, self.slab_views = [], []
        self.ptrs = [int(v.data_ptr()) for v in self.views]
        self.slab_ptrs = [int(v.data_ptr()) for v in self.slab_views]
        self.begins = [shard_bounds(n_total, world, r)[0] for r in range(world)] + [n_total]

    @classmethod
    def release(cls, group=None):
        """Collective: unmap the peers' arrays, then free the own ones (a producer must outlive its consumers' mappings)."""
        for made in cls._cache.values():
            if made is not None:
                made.views, made.slab_views = [], []
                made.ptrs, made.slab_ptrs = [], []
        if dist.is_initialized():
            if torch.cuda.is_available():
                import gc

                torch.cuda.synchronize()
                gc.collect()
                torch.cuda.ipc_collect()      # closes the mappings whose tensors are gone
            dist.barrier(group=group)
        cls._cache.clear()
        if torch.cuda.is_available():
            torch.cuda.ipc_collect()

    @classmethod
    def get(cls, n_total: int, group, device):
        """Collective on first use.  None when the job cannot share device memory (every rank then agrees on that)."""
        import os

        key = (id(group) if group is not None else 0, str(device), int(n_total), dist.get_world_size(group))
        if key in cls._cache:
            return cls._cache[key]
        if cls._cache:
            cls.release(group)        # another cloud size: the old buffers go first (every rank takes this branch together)
        ok, made = 1, None
        if os.environ.get("PCT_PEER_RETURN", "1") == "0" or dist.get_backend(group) != "nccl" or device.type != "cuda":
            ok = 0
        flags = [None] * dist.get_world_size(group)
        dist.all_gather_object(flags, ok, group=group)
        if all(flags):
            made = cls(n_total, group, device)
            if made.error is not None:  # pragma: no cover  (no IPC between these processes: fall back to the all-to-alls)
                import warnings

                warnings.warn(f"peer arrays unavailable ({made.error!r}); the exchanges go through NCCL")
            dist.all_gather_object(flags, 0 if made.error is not None else 1, group=group)
            if not all(flags):
                made.views, made.slab_views = [], []
                made = None
        cls._cache[key] = made
        return made


def answer_slab(recv: torch.Tensor, plan: SlabPlan, rank: int, k: int, n_own: int, peers: PeerResults = None):
    """Index over the received slab cloud (x, y, z, original-index bits) and the fused kernel on the points it owns.
    With ``peers`` the kernel also stores K, H of every answered query into the array of the rank that holds the
    query's point.  Returns (records (n_own, 8), index, unresolved count)."""
    from . import engine

    c_lo, c_hi, own_lo, own_hi = plan.bounds[rank]
    if n_own == 0 or int(recv.shape[0]) == 0:
        return torch.empty((0, 8), dtype=torch.float32, device=recv.device), None, 0
    index = engine.GridIndex(recv, cell_hint=plan.h, k_hint=k)
    row_map = engine.slab_rows(recv, plan.axis, own_lo, own_hi)
    index.set_slab(plan.axis, c_lo, c_hi, own_lo, own_hi, row_map=row_map, mapped_rows=n_own)
    if peers is not None:
        index.set_peers(peers.begins, peers.ptrs, engine.slab_row_ids(recv, plan.axis, own_lo, own_hi, row_map, n_own))
    if int(recv.shape[0]) <= k:
        # fewer points than a neighbourhood needs: nothing can be resolved inside this slab
        rec = torch.full((n_own, 8), float("nan"), dtype=torch.float32, device=recv.device)
        return rec, index, n_own
    rec = index.curvature_knn(k, want_coeffs=False).records
    return rec, index, int(index.last_stats().unresolved)


def curvature_knn_exchange(share: torch.Tensor, id_base: int, n_total: int, k: int, group=None, columns=(3, 4),
                           stages: Stages = None, bin_fn=None, answer_fn=None, cell_fn=None, redo_fn=None) -> ExchangeFit:
    """plant_kdtree(k) + compute_pointwise_explicit_quadratic_curvature() of a cloud DISTRIBUTED over the ranks.

    ``share``: this rank's contiguous rows ``[id_base, id_base + len(share))`` of the cloud, (n, 3) float32 on the
    device (rank order = index order).  Collective over ``group``:

      1. one small all-gather (boxes + sample) -> cell edge, axis, cut planes        (plan_slabs)
      2. every rank bins its share by destination slab (owner + margins)             (pct_slab_bin_*)
      3. the 16-byte point records reach their slab's rank -- stored there by the binning kernel itself over NVLink peer
         memory (``PeerResults``), else one all-to-all of counts and one of the records: each rank now holds its slab + margin,
         in ascending original index (ties keep the whole cloud's order)
      4. slab index + fused kernel on the owned points                               (answer_slab)
      5. the rows return to the ranks whose share the points came from -- stored there by the fused kernel itself
         (``GridIndex.set_peers``), else by one all-to-all in which no ids travel: a slab returns rows in the order it
         received the points, which the sender remembers (``owned_local``)

    Queries a slab cannot resolve inside its margin (isolated points) are redone on a whole-cloud index after an
    all-gather of the shares; that collective only happens when some rank reports such a query.
    ``columns``: record columns returned (3 = K, 4 = H).  The ``*_fn`` hooks exist for the gloo tests, which have no GPU.
    """
    from . import engine

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    st = stages or Stages(False)
    st.mark("start")
    plan = plan_slabs(share, n_total, k, group, cell_fn)
    st.mark("plan")
    # both exchanges happen inside kernels (peer stores over NVLink) when the ranks can map each other's memory
    peers = None
    if bin_fn is None and answer_fn is None and tuple(columns) == (3, 4):
        peers = PeerResults.get(n_total, group, share.device)
    slab = None
    if peers is not None:
        binc = engine.SlabBinCount(share, plan.axis, plan.bounds)
        complete, owned = binc.complete, binc.owned
        mine = torch.tensor(complete + owned, dtype=torch.int64, device=share.device)
        table = torch.empty((world, 2 * world), dtype=torch.int64, device=share.device)
        dist.all_gather_into_tensor(table, mine, group=group)
        table = table.tolist()                                   # the one host sync of the exchange
        recv_complete = [int(table[src][rank]) for src in range(world)]
        recv_owned = [int(table[src][world + rank]) for src in range(world)]
        if all(sum(table[src][d] for src in range(world)) <= peers.slab_cap for d in range(world)):
            # rank d receives the records of the lower ranks first: its slab is in ascending original index
            dest_rows = [sum(table[src][d] for src in range(rank)) for d in range(world)]
            owned_local = binc.fill_peers(id_base, peers.slab_ptrs, dest_rows)
            st.mark("bin")
            landed = torch.zeros(1, dtype=torch.int32, device=share.device)
            dist.all_reduce(landed, group=group)                 # barrier: every rank's records are in place
            slab = peers.slab_local[: sum(recv_complete)]
        else:
            records, owned_local = binc.fill(id_base)
            st.mark("bin")
            slab = _all_to_all_rows(records, complete, recv_complete, group)
    else:
        records, complete, owned, owned_local = (bin_fn or engine.slab_bin)(share, plan.axis, plan.bounds, id_base)
        st.mark("bin")
        send = torch.tensor([complete, owned], dtype=torch.int64, device=share.device).t().contiguous()    # (world, 2)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        recv_l = recv.tolist()                                      # the one host sync of the exchange
        recv_complete, recv_owned = [int(r[0]) for r in recv_l], [int(r[1]) for r in recv_l]
        slab = _all_to_all_rows(records, complete, recv_complete, group)                                # (m, 4)
    st.mark("exchange")
    n_own = sum(recv_owned)
    if peers is not None:
        rec, index, n_bad = answer_slab(slab, plan, rank, k, n_own, peers)
    else:
        rec, index, n_bad = (answer_fn or answer_slab)(slab, plan, rank, k, n_own)
    st.mark("answer")
    bad_total = torch.tensor([n_bad], dtype=torch.int64, device=share.device)
    dist.all_reduce(bad_total, group=group)
    any_bad = int(bad_total.item())
    if any_bad:
        # rare: some k-th neighbour lies beyond a margin -- replicate the cloud and redo those queries
        rows = padded_rows(n_total, world)
        padded = torch.zeros((rows, 3), dtype=torch.float32, device=share.device)
        padded[: share.shape[0]] = share[:, :3]
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        if n_bad:
            cloud = torch.cat([parts[r][: shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0]]
                               for r in range(world)], 0)
            rec = (redo_fn or _redo_unresolved)(cloud, rec, _owned_ids(slab, plan, rank), plan, k)
        del parts
    st.mark("unresolved")
    if peers is not None and not any_bad:
        # every rank's kernel has finished (the all-reduce above is the barrier): my rows are complete in my array.
        # A copy, because the array is written again by the next call.
        out = peers.local[: int(share.shape[0])].clone()
    else:
        cols = rec[:, list(columns)].contiguous()
        back = _all_to_all_rows(cols, recv_owned, owned, group)     # rows of MY share, grouped by the slab that answered
        out = torch.empty((int(share.shape[0]), len(columns)), dtype=rec.dtype, device=share.device)
        out[owned_local] = back          # (int32 indices: no 64-bit copy of the index list)
    st.mark("return")
    fit = ExchangeFit(out, plan, index, slab, rank, rec, n_bad)
    fit.peer_return = peers is not None
    if peers is not None and index is not None and getattr(index, "_peer_keep", None) is not None:
        fit._own_ids = index._peer_keep        # the row ids the kernel routed by = the owned original indices, ascending
    return fit


def _owned_ids(slab, plan, rank):
    c_lo, c_hi, own_lo, own_hi = plan.bounds[rank]
    xs = slab[:, plan.axis]
    return slab[:, 3].contiguous().view(torch.int32)[(xs >= own_lo) & (xs < own_hi)]


def _redo_unresolved(cloud, rec, own_ids, plan, k):
    from . import engine
    from ._lib import STATUS_UNRESOLVED

    bad = (rec[:, 7].contiguous().view(torch.int32) & STATUS_UNRESOLVED) != 0
    whole = engine.GridIndex(cloud, cell_hint=plan.h, k_hint=k)
    redo = whole.curvature_points(own_ids[bad], k)
    rec[bad] = redo.records
    whole.close()
    return rec


class CopyRounds:
    """How many ranks of a one-node job should copy to / from the host AT ONCE.

    The GPUs of a box share PCIe uplinks: on the HGX boards measured (profiles/pcie_probe_8gpu_r02n.txt) eight
    concurrent D2H copies move 96 GB/s in aggregate, four (every second GPU) 108 GB/s, and eight concurrent H2D copies
    187 GB/s against 217 GB/s for four.  So a host copy of all ranks can be faster in ROUNDS -- round p: the ranks with
    rank % R == p copy, a stream-ordered barrier (4-byte all-reduce) separates the rounds, no host synchronisation.
    R is not assumed: the first call on a group times R = 1, 2 and 4 on a 32 MB piece of the caller's own arrays and
    keeps the fastest (a smaller R wins ties within 5 %)."""

    _cache = {}
    timings = []          # [{rounds: ms} of the H2D probe, {rounds: ms} of the D2H probe] of the last calibration
    PROBE_ROWS = 1 << 22

    @staticmethod
    def run(rounds, rank, copy_fn, token, group):
        """The copy of every rank in ``rounds`` rounds; ``token``: a 1-element device tensor for the barriers."""
        for phase in range(rounds):
            if rank % rounds == phase:
                copy_fn()
            if phase + 1 < rounds:
                dist.all_reduce(token, group=group)

    @classmethod
    def get(cls, points, out, begin, end, group, device):
        """(rounds H2D, rounds D2H) of this group; collective on first use."""
        import os

        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        key = (id(group) if group is not None else 0, world, str(device))
        if key in cls._cache:
            return cls._cache[key]
        fixed = os.environ.get("PCT_COPY_ROUNDS")       # "h2d,d2h": skips the probe (experiments)
        if fixed or world < 4 or device.type != "cuda":
            got = tuple(int(v) for v in fixed.split(",")) if fixed else (1, 1)
            cls._cache[key] = got
            return got
        rows = max(1, min(cls.PROBE_ROWS, end - begin))
        dev_in = torch.empty((rows, 3), dtype=torch.float32, device=device)
        dev_out = torch.zeros((rows,), dtype=torch.float32, device=device)
        token = torch.zeros(1, dtype=torch.int32, device=device)
        src = points.tensor[begin:begin + rows]
        dst0, dst1 = out.tensor[0, begin:begin + rows], out.tensor[1, begin:begin + rows]

        def probe_out():
            dst0.copy_(dev_out, non_blocking=True)
            dst1.copy_(dev_out, non_blocking=True)

        best = []
        for copy_fn in (lambda: dev_in.copy_(src, non_blocking=True), probe_out):
            times = {}
            for rounds in (1, 2, 4):
                if world % rounds:
                    continue
                t_best = None
                for rep in range(3):
                    dist.all_reduce(token, group=group)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    cls.run(rounds, rank, copy_fn, token, group)
                    dist.all_reduce(token, group=group)
                    e1.record()
                    e1.synchronize()
                    t = torch.tensor([e0.elapsed_time(e1)], device=device)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
                    if rep and (t_best is None or t.item() < t_best):
                        t_best = t.item()
                times[rounds] = t_best
            pick = 1
            for rounds in sorted(times):
                if times[rounds] < 0.95 * times[pick]:
                    pick = rounds
            best.append(pick)
            cls.timings.append(times)
        cls._cache[key] = tuple(best)
        return cls._cache[key]


def curvature_knn_shared(points: SharedHostArray, out: SharedHostArray, k: int, group=None, device=None, stages: Stages = None):
    """plant_kdtree(k) + compute_pointwise_explicit_quadratic_curvature() of a cloud in shared host memory.

    ``points`` (N, 3) and ``out`` (2, N) = [K; H] are mapped by every rank.  Rank r copies rows
    ``shard_bounds(N, world, r)`` of the cloud to its GPU over its own PCIe link, the ranks trade slabs
    (``curvature_knn_exchange``: the cloud is never replicated), and each rank writes the K and H of its own rows
    to ``out``.  The host copies run in as many rounds as ``CopyRounds`` measured to be fastest on this box.
    Collective: returns after a barrier, when ``out`` is complete on the host."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    n = int(points.tensor.shape[0])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    begin, end = shard_bounds(n, world, rank)
    r_in, r_out = CopyRounds.get(points, out, begin, end, group, device)
    token = torch.zeros(1, dtype=torch.int32, device=device) if max(r_in, r_out) > 1 else None
    st = stages or Stages(False)
    st.mark("begin")
    share = torch.empty((end - begin, 3), dtype=torch.float32, device=device)
    CopyRounds.run(r_in, rank, lambda: share.copy_(points.tensor[begin:end], non_blocking=True), token, group)   # this rank's share
    st.mark("h2d")
    part = curvature_knn_exchange(share, begin, n, k, group, stages=st)
    khT = part.rows.t().contiguous()

    def to_host():
        out.tensor[0, begin:end].copy_(khT[0], non_blocking=True)
        out.tensor[1, begin:end].copy_(khT[1], non_blocking=True)

    CopyRounds.run(r_out, rank, to_host, token, group)
    st.mark("d2h")
    torch.cuda.current_stream(device).synchronize()
    dist.barrier(group=group)
    part.copy_rounds = (r_in, r_out)
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
