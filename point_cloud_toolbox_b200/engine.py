"""Device-side objects behind the PointCloud drop-in: torch owns memory and
streams, libpct_b200.so (through ctypes) does all the work.

``GridIndex`` stands where the reference keeps ``self.kdtree``
(/root/reference/pointCloudToolbox.py:74).
"""
from __future__ import annotations

import ctypes
import torch

from . import _lib
from ._lib import LAYOUT_ORIGINAL, LAYOUT_SLICE, check, lib, ptr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "point_cloud_toolbox_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback"
        )


def to_device_points(points, device=None):
    """(N, 3) float32 contiguous CUDA tensor from numpy / torch input (H2D when on host)."""
    require_cuda()
    if isinstance(points, torch.Tensor):
        t = points
    else:
        import numpy as np

        t = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32))
    if t.ndim != 2 or t.shape[1] < 3:
        raise ValueError("points must have shape (N, 3)")
    if t.shape[1] != 3:
        t = t[:, :3]
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if (not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[1] == 3
            and t.numel() >= (8 << 20) and not t.is_pinned()):
        # a large pageable array: staged by several host threads (pct_upload) instead of the driver's single one
        out = torch.empty(t.shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pct_upload(ptr(out), ctypes.c_void_p(t.data_ptr()), t.numel() * 4, _stream()))
        return out
    return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


class _PinnedPool:
    """Recycles pinned host buffers behind the numpy arrays handed to the caller.

    ``cudaHostAlloc`` costs about as much as copying the data it will hold, so result buffers are reused -- but
    only once the numpy array that was made from a buffer, and every view of it, is gone.  A buffer is LEASED to
    the array ``to_host`` returns: ``weakref.finalize`` on that array gives the lease back.  Views keep their
    parent alive through ``.base`` (numpy stops collapsing base chains at the array whose own base is not an
    ndarray -- the tensor's buffer object here), so the finalizer runs exactly when nobody can see the old data.
    """

    def __init__(self, max_bytes=16 << 30):
        self.entries = []  # [pinned uint8 tensor, leased?]
        self.max_bytes = max_bytes
        self.allocations = 0

    def take(self, shape, dtype):
        """(tensor view of a free buffer, its pool entry); the entry is leased until ``release(entry)``."""
        nbytes = 1
        for d in shape:
            nbytes *= int(d)
        nbytes *= torch.empty((), dtype=dtype).element_size()
        for ent in self.entries:
            buf, leased = ent
            if not leased and nbytes <= buf.numel() <= 2 * nbytes + 4096:
                ent[1] = True
                return buf[:nbytes].view(dtype).view(shape), ent
        held = sum(e[0].numel() for e in self.entries)
        for ent in list(self.entries):
            if held + nbytes <= self.max_bytes:
                break
            if not ent[1]:
                self.entries.remove(ent)
                held -= ent[0].numel()
        self.allocations += 1
        buf = torch.empty((max(nbytes, 1),), dtype=torch.uint8, pin_memory=True)
        ent = [buf, True]
        self.entries.append(ent)
        return buf[:nbytes].view(dtype).view(shape), ent

    @staticmethod
    def release(ent):
        ent[1] = False


_PINNED = _PinnedPool()


def to_host(t: torch.Tensor):
    """Device tensor -> numpy array backed by (recycled) pinned host memory; synchronises the stream."""
    import weakref

    if not t.is_cuda:
        return t.numpy()
    t = t.contiguous()
    h, lease = _PINNED.take(tuple(t.shape), t.dtype)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    arr = h.numpy()
    weakref.finalize(arr, _PinnedPool.release, lease)
    return arr


def load_text_f32(path, threads=0) -> torch.Tensor:
    """Whitespace-separated numeric text file -> (rows, cols) float32 tensor in pinned host memory.

    Values equal ``np.loadtxt(path).astype(np.float32)`` bit for bit (ref :51-53); a file of one row
    comes back one-dimensional and an empty file as shape (0,), like ``np.loadtxt``.  Raises
    ``ValueError`` on a malformed token or a ragged row and ``FileNotFoundError`` like ``np.loadtxt``.
    """
    import os

    path = os.fspath(path)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found.")
    rows, cols = ctypes.c_int64(), ctypes.c_int64()
    check(lib.pct_text_shape(path.encode(), ctypes.byref(rows), ctypes.byref(cols)))
    pin = torch.cuda.is_available()
    out = torch.empty((rows.value, cols.value), dtype=torch.float32, pin_memory=pin)
    check(lib.pct_text_load_f32(path.encode(), rows.value, cols.value, ctypes.c_void_p(out.data_ptr()), int(threads)))
    if rows.value == 0:
        return out.reshape(0)
    if rows.value == 1:
        return out.reshape(cols.value)
    return out


class FitOutputs:
    """Per-point results of a fit, resident on the device.

    Two storage forms: separate arrays (``normals (n,3)``, ``coeffs (n,6)``, ``curv (n,5)``
    = K, H, k1, k2, H^2, ``status (n,) uint8``) or packed 32-byte ``records (n,8)`` =
    nx, ny, nz, K, H, k1, k2, status bits.  The accessors work on either.
    """

    _COL = {"K": 0, "H": 1, "k1": 2, "k2": 3, "H2": 4}

    def __init__(self, normals=None, coeffs=None, curv=None, status=None, counts=None, records=None):
        self._normals, self.coeffs, self._curv, self._status = normals, coeffs, curv, status
        self.counts = counts
        self.records = records

    def column(self, name):
        c = self._COL[name]
        if self.records is None:
            return self._curv[:, c]
        if name == "H2":
            h = self.records[:, 4]
            return h * h  # fp32 product, the same rounding as the kernel's / the reference's K_h**2
        return self.records[:, 3 + c]

    def kh(self):
        """(2, n) contiguous tensor [K; H]."""
        src = self.records[:, 3:5] if self.records is not None else self._curv[:, :2]
        return src.t().contiguous()

    @property
    def normals(self):
        return self.records[:, :3] if self.records is not None else self._normals

    @property
    def status(self):
        if self.records is not None:
            return self.records[:, 7].contiguous().view(torch.int32).to(torch.uint8)
        return self._status

    @property
    def curv(self):
        if self.records is not None:
            return torch.stack([self.column(n) for n in ("K", "H", "k1", "k2", "H2")], 1)
        return self._curv


def _alloc_outputs(rows, device, want_normals=True, want_coeffs=True, want_status=True):
    f32 = dict(dtype=torch.float32, device=device)
    return FitOutputs(
        normals=torch.empty((rows, 3), **f32) if want_normals else None,
        coeffs=torch.empty((rows, 6), **f32) if want_coeffs else None,
        curv=torch.empty((rows, 5), **f32),
        status=torch.empty((rows,), dtype=torch.uint8, device=device) if want_status else None,
    )


class GridIndex:
    """Morton-sorted uniform grid over a cloud resident in HBM."""

    def __init__(self, points_dev: torch.Tensor, k_hint: int = 20, cell_hint: float = 0.0):
        require_cuda()
        if points_dev.dtype != torch.float32 or not points_dev.is_cuda or not points_dev.is_contiguous():
            raise ValueError("GridIndex needs a contiguous float32 CUDA tensor of shape (N, 3|4)")
        self.points = points_dev
        self.n = int(points_dev.shape[0])
        self.device = points_dev.device
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.pct_index_build(ptr(points_dev), self.n, int(points_dev.shape[1]), float(cell_hint), int(k_hint),
                                      _stream(), ctypes.byref(self._handle)))
        self.k_hint = k_hint

    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            lib.pct_index_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection -----------------------------------------------------
    def info(self) -> _lib.IndexInfo:
        out = _lib.IndexInfo()
        check(lib.pct_index_get_info(self._handle, ctypes.byref(out)))
        return out

    def last_stats(self) -> _lib.QueryStats:
        out = _lib.QueryStats()
        with torch.cuda.device(self.device):
            check(lib.pct_index_last_stats(self._handle, _stream(), ctypes.byref(out)))
        return out

    def permutation(self) -> torch.Tensor:
        perm = torch.empty((self.n,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pct_index_permutation(self._handle, ptr(perm), _stream()))
        return perm

    def _range(self, q_begin, q_end, layout):
        q_begin = 0 if q_begin is None else int(q_begin)
        q_end = self.n if q_end is None else int(q_end)
        rows = self.n if layout == LAYOUT_ORIGINAL else q_end - q_begin
        if layout == LAYOUT_ORIGINAL and getattr(self, "_mapped_rows", None) is not None:
            rows = self._mapped_rows
        return q_begin, q_end, rows

    # -- queries -----------------------------------------------------------
    def knn(self, k, q_begin=None, q_end=None, layout=LAYOUT_ORIGINAL, want_dist=True):
        q_begin, q_end, rows = self._range(q_begin, q_end, layout)
        idx = torch.empty((rows, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((rows, k), dtype=torch.float32, device=self.device) if want_dist else None
        with torch.cuda.device(self.device):
            check(lib.pct_knn(self._handle, q_begin, q_end, int(k), ptr(idx), ptr(dist), layout, _stream()))
        return idx, dist

    def set_slab(self, axis, complete_lo, complete_hi, own_lo, own_hi, row_map=None, mapped_rows=None):
        """This index is one slab of a partitioned cloud (see pct_index_set_slab); axis=-1 undoes it.
        ``row_map`` (N int32 on the device): output row of every owned point, outputs of original-layout
        calls then have ``mapped_rows`` rows."""
        self._row_map = row_map  # kept alive
        self._mapped_rows = None if row_map is None or axis < 0 else int(mapped_rows)
        check(lib.pct_index_set_slab(self._handle, int(axis), float(complete_lo), float(complete_hi), float(own_lo),
                                     float(own_hi), ptr(row_map)))

    def set_peers(self, begins, peer_ptrs, row_ids):
        """Multi-GPU return fused into the kernel (pct_index_set_peers): ``begins`` (world + 1 original-index bounds of the
        ranks' shares), ``peer_ptrs`` (device address of every rank's (rows, 2) K, H array as this process maps it),
        ``row_ids`` (device int32: original index of every output row, ``slab_row_ids``).  ``begins=None`` removes it."""
        if begins is None:
            self._peer_keep = None
            check(lib.pct_index_set_peers(self._handle, 0, None, None, None))
            return
        world = len(peer_ptrs)
        b = (ctypes.c_int64 * (world + 1))(*[int(v) for v in begins])
        p = (ctypes.c_void_p * world)(*[int(v) for v in peer_ptrs])
        self._peer_keep = row_ids
        with torch.cuda.device(self.device):
            check(lib.pct_index_set_peers(self._handle, world, b, p, ptr(row_ids)))

    def curvature_points(self, query_ids, k) -> "FitOutputs":
        """Fused search + fit for the cloud points named by original index; packed records, row r = query r."""
        ids = torch.as_tensor(query_ids, device=self.device).to(torch.int32).contiguous()
        nq = int(ids.numel())
        rec = torch.empty((nq, 8), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pct_curvature_points_records(self._handle, ptr(self.points), int(self.points.shape[1]), ptr(ids), nq,
                                                   int(k), ptr(rec), _stream()))
        return FitOutputs(records=rec)

    def knn_points(self, query_ids, k):
        """Ordered kNN rows (original indices, distances) of the cloud points named by original index."""
        ids = torch.as_tensor(query_ids, device=self.device).to(torch.int32).contiguous()
        nq = int(ids.numel())
        if nq and (int(ids.min()) < 0 or int(ids.max()) >= self.n):
            raise IndexError("query ids out of range")
        idx = torch.empty((nq, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pct_knn_points(self._handle, ptr(self.points), int(self.points.shape[1]), ptr(ids), nq, int(k),
                                     ptr(idx), ptr(dist), _stream()))
        return idx, dist

    def query(self, x_dev, k):
        """k nearest cloud points of arbitrary coordinates ``x_dev`` (m, 3) float32 on the device (pct_knn_query):
        ``(dist float64 (m, k), idx int32 (m, k))`` ordered by (distance, index); a query that is a cloud point
        finds itself first."""
        x = x_dev.to(device=self.device, dtype=torch.float32).contiguous()
        m = int(x.shape[0])
        idx = torch.empty((m, k), dtype=torch.int32, device=self.device)
        dist = torch.empty((m, k), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pct_knn_query(self._handle, ptr(x), m, int(k), ptr(idx), ptr(dist), _stream()))
        return dist, idx

    def curvature_knn(self, k, q_begin=None, q_end=None, layout=LAYOUT_ORIGINAL, want_normals=True, want_coeffs=True,
                      want_status=True) -> FitOutputs:
        """Fused search + fit.  Without coefficients the result is written as packed 32-byte records."""
        q_begin, q_end, rows = self._range(q_begin, q_end, layout)
        with torch.cuda.device(self.device):
            if not want_coeffs:
                rec = torch.empty((rows, 8), dtype=torch.float32, device=self.device)
                check(lib.pct_curvature_fused_knn_records(self._handle, q_begin, q_end, int(k), ptr(rec), layout, _stream()))
                return FitOutputs(records=rec)
            out = _alloc_outputs(rows, self.device, want_normals, want_coeffs, want_status)
            check(lib.pct_curvature_fused_knn(self._handle, q_begin, q_end, int(k), ptr(out._normals), ptr(out.coeffs),
                                              ptr(out._curv), ptr(out._status), layout, _stream()))
        return out

    def ball_count(self, radius, q_begin=None, q_end=None, layout=LAYOUT_ORIGINAL):
        q_begin, q_end, rows = self._range(q_begin, q_end, layout)
        counts = torch.empty((rows,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pct_ball_count(self._handle, q_begin, q_end, float(radius), ptr(counts), layout, _stream()))
        return counts

    def ball(self, radius, q_begin=None, q_end=None, layout=LAYOUT_ORIGINAL):
        """CSR (offsets int64, idx int32, dist float32), rows ordered by (d2, index)."""
        counts = self.ball_count(radius, q_begin, q_end, layout)
        offsets = torch.zeros((counts.numel() + 1,), dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        nnz = int(offsets[-1].item())
        idx = torch.empty((max(nnz, 1),), dtype=torch.int32, device=self.device)
        dist = torch.empty((max(nnz, 1),), dtype=torch.float32, device=self.device)
        q_begin, q_end, _ = self._range(q_begin, q_end, layout)
        with torch.cuda.device(self.device):
            check(lib.pct_ball_fill(self._handle, q_begin, q_end, float(radius), ptr(offsets), nnz, ptr(idx), ptr(dist),
                                    layout, _stream()))
        return offsets, idx[:nnz], dist[:nnz]

    def curvature_ball(self, radius, q_begin=None, q_end=None, layout=LAYOUT_ORIGINAL) -> FitOutputs:
        q_begin, q_end, rows = self._range(q_begin, q_end, layout)
        out = _alloc_outputs(rows, self.device)
        out.counts = torch.empty((rows,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.pct_curvature_fused_ball(self._handle, q_begin, q_end, float(radius), ptr(out.counts), ptr(out._normals),
                                               ptr(out.coeffs), ptr(out._curv), ptr(out._status), layout, _stream()))
        return out


# -- list-driven fit and the batched static methods ---------------------------
def estimate_cell_size(points_dev, k_hint=20):
    """(cell edge the index build would choose, bbox min (3,), bbox max (3,)) of a device cloud."""
    h = ctypes.c_float()
    box = (ctypes.c_float * 6)()
    with torch.cuda.device(points_dev.device):
        check(lib.pct_estimate_cell_size(ptr(points_dev), int(points_dev.shape[0]), int(points_dev.shape[1]), int(k_hint),
                                         _stream(), ctypes.byref(h), box))
    return float(h.value), [box[0], box[1], box[2]], [box[3], box[4], box[5]]


def slab_select(points_dev, axis, bounds):
    """Device-side slab selection (pct_slab_select + pct_slab_gather).

    Returns ``(sel int32 (m,), local (m, 3), row_map int32 (m,), n_own)``: ascending original indices of the
    points inside the slab's complete range, those points, the compact output row of every owned point."""
    c_lo, c_hi, own_lo, own_hi = bounds
    n = int(points_dev.shape[0])
    stride = int(points_dev.shape[1])
    sel = torch.empty((n,), dtype=torch.int32, device=points_dev.device)
    m = ctypes.c_int64()
    with torch.cuda.device(points_dev.device):
        check(lib.pct_slab_select(ptr(points_dev), n, stride, int(axis), float(c_lo), float(c_hi), ptr(sel), ctypes.byref(m), _stream()))
        m = int(m.value)
        sel = sel[:m]
        local = torch.empty((m, 3), dtype=torch.float32, device=points_dev.device)
        row_map = torch.empty((m,), dtype=torch.int32, device=points_dev.device)
        n_own = ctypes.c_int64()
        check(lib.pct_slab_gather(ptr(points_dev), stride, int(axis), ptr(sel), m, float(own_lo), float(own_hi), ptr(local),
                                  ptr(row_map), ctypes.byref(n_own), _stream()))
    return sel, local, row_map, int(n_own.value)


def estimate_cell_size_sample(sample_dev, n_total, bbox, k_hint=20):
    """Cell edge the index build would choose for a cloud of ``n_total`` points given a sample of it and the whole
    cloud's bounding box ``bbox`` = [min xyz, max xyz] (pct_estimate_cell_size_sample)."""
    h = ctypes.c_float()
    box = (ctypes.c_float * 6)(*[float(v) for v in bbox])
    with torch.cuda.device(sample_dev.device):
        check(lib.pct_estimate_cell_size_sample(ptr(sample_dev), int(sample_dev.shape[0]), int(sample_dev.shape[1]), int(n_total),
                                                box, int(k_hint), _stream(), ctypes.byref(h)))
    return float(h.value)


class SlabBinCount:
    """First half of the binning (pct_slab_bin_count): how many of the share's points go to every slab."""

    def __init__(self, share_dev, axis, bounds):
        self.share, self.axis, self.bounds = share_dev, int(axis), bounds
        n, world = int(share_dev.shape[0]), len(bounds)
        self.flat = (ctypes.c_float * (4 * world))(*[float(v) for b in bounds for v in b])
        self.complete, self.owned, self.block_pos = [0] * world, [0] * world, None
        if n == 0:
            return
        blocks = int(lib.pct_slab_bin_blocks(n))
        self.block_pos = torch.empty((2 * world * blocks + 1,), dtype=torch.int32, device=share_dev.device)
        counts = (ctypes.c_int64 * (2 * world))()
        with torch.cuda.device(share_dev.device):
            check(lib.pct_slab_bin_count(ptr(share_dev), n, int(share_dev.shape[1]), self.axis, world, self.flat, ptr(self.block_pos),
                                         counts, _stream()))
        self.complete = [int(counts[d]) for d in range(world)]
        self.owned = [int(counts[world + d]) for d in range(world)]

    def fill(self, id_base):
        """Second half into a local array (pct_slab_bin_fill): ``(records (T, 4) grouped by destination, owned_local)``."""
        dev, n = self.share.device, int(self.share.shape[0])
        total = sum(self.complete)
        if n == 0:
            return torch.empty((0, 4), dtype=torch.float32, device=dev), torch.empty((0,), dtype=torch.int32, device=dev)
        records = torch.empty((max(total, 1), 4), dtype=torch.float32, device=dev)
        owned_local = torch.empty((n,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pct_slab_bin_fill(ptr(self.share), n, int(self.share.shape[1]), self.axis, len(self.bounds), self.flat,
                                        ptr(self.block_pos), total, int(id_base), ptr(records), ptr(owned_local), _stream()))
        return records[:total], owned_local[:sum(self.owned)]

    def fill_peers(self, id_base, peer_ptrs, dest_rows):
        """Second half straight into the destination ranks' slab buffers (pct_slab_bin_fill_peers); returns owned_local."""
        dev, n, world = self.share.device, int(self.share.shape[0]), len(self.bounds)
        if n == 0:
            return torch.empty((0,), dtype=torch.int32, device=dev)
        owned_local = torch.empty((n,), dtype=torch.int32, device=dev)
        comp = (ctypes.c_int64 * world)(*self.complete)
        p = (ctypes.c_void_p * world)(*[int(v) for v in peer_ptrs])
        rows = (ctypes.c_int64 * world)(*[int(v) for v in dest_rows])
        with torch.cuda.device(dev):
            check(lib.pct_slab_bin_fill_peers(ptr(self.share), n, int(self.share.shape[1]), self.axis, world, self.flat,
                                              ptr(self.block_pos), comp, int(id_base), p, rows, ptr(owned_local), _stream()))
        return owned_local[:sum(self.owned)]


def slab_bin(share_dev, axis, bounds, id_base):
    """Bins a contiguous share of the cloud by destination slab (pct_slab_bin_count + pct_slab_bin_fill).

    ``bounds``: per slab (complete_lo, complete_hi, own_lo, own_hi).  Returns ``(records (T, 4) float32 {x, y, z,
    original index bits} grouped by destination, complete counts, owned counts, owned_local int32 (n,))``."""
    c = SlabBinCount(share_dev, axis, bounds)
    records, owned_local = c.fill(id_base)
    return records, c.complete, c.owned, owned_local


def slab_rows(cloud_dev, axis, own_lo, own_hi):
    """row_map of GridIndex.set_slab for a slab cloud: rank of every owned point among the owned ones (pct_slab_rows)."""
    m = int(cloud_dev.shape[0])
    row_map = torch.empty((m,), dtype=torch.int32, device=cloud_dev.device)
    with torch.cuda.device(cloud_dev.device):
        check(lib.pct_slab_rows(ptr(cloud_dev), m, int(cloud_dev.shape[1]), int(axis), float(own_lo), float(own_hi), ptr(row_map),
                                _stream()))
    return row_map


def slab_row_ids(cloud_dev, axis, own_lo, own_hi, row_map, n_own):
    """Original (whole-cloud) index of every output row of a slab cloud (x, y, z, id bits): pct_slab_row_ids."""
    if cloud_dev.shape[1] != 4 or not cloud_dev.is_contiguous():
        raise ValueError("slab_row_ids needs the packed (m, 4) slab records")
    row_ids = torch.empty((int(n_own),), dtype=torch.int32, device=cloud_dev.device)
    with torch.cuda.device(cloud_dev.device):
        check(lib.pct_slab_row_ids(ptr(cloud_dev), int(cloud_dev.shape[0]), int(axis), float(own_lo), float(own_hi), ptr(row_map),
                                   ptr(row_ids), _stream()))
    return row_ids


def fit_from_neighbors(points_dev, idx_dev, query_ids=None) -> FitOutputs:
    """Fit rows of original-index neighbour lists (nq, k) on an (N, 3) cloud."""
    if points_dev.shape[1] != 3 or not points_dev.is_contiguous():
        raise ValueError("fit_from_neighbors needs packed (N, 3) points")
    idx_dev = idx_dev.to(torch.int32).contiguous()
    nq, k = idx_dev.shape
    out = _alloc_outputs(nq, points_dev.device)
    with torch.cuda.device(points_dev.device):
        check(lib.pct_fit_from_neighbors(ptr(points_dev), int(points_dev.shape[0]), ptr(idx_dev), nq, k, ptr(query_ids),
                                         ptr(out._normals), ptr(out.coeffs), ptr(out._curv), ptr(out._status), _stream()))
    return out


def fit_from_csr(points_dev, offsets_dev, idx_dev, query_ids=None) -> FitOutputs:
    nq = int(offsets_dev.numel()) - 1
    out = _alloc_outputs(nq, points_dev.device)
    idx_dev = idx_dev.to(torch.int32).contiguous()
    offsets_dev = offsets_dev.to(torch.int64).contiguous()
    with torch.cuda.device(points_dev.device):
        check(lib.pct_fit_from_csr(ptr(points_dev), int(points_dev.shape[0]), ptr(offsets_dev), ptr(idx_dev), nq,
                                   ptr(query_ids), ptr(out._normals), ptr(out.coeffs), ptr(out._curv), ptr(out._status),
                                   _stream()))
    return out


def plane_rotate(centered_dev):
    """(nq, k, 3) float32 centred neighbourhoods -> rotated float64, unit normals float64, status."""
    c = centered_dev.to(torch.float32).contiguous()
    nq, k, _ = c.shape
    rotated = torch.empty((nq, k, 3), dtype=torch.float64, device=c.device)
    normals = torch.empty((nq, 3), dtype=torch.float64, device=c.device)
    status = torch.empty((nq,), dtype=torch.uint8, device=c.device)
    with torch.cuda.device(c.device):
        check(lib.pct_plane_rotate(ptr(c), nq, k, ptr(rotated), ptr(normals), ptr(status), _stream()))
    return rotated, normals, status


def quadric_fit(rotated_dev):
    r = rotated_dev.to(torch.float64).contiguous()
    nq, k, _ = r.shape
    coeffs = torch.empty((nq, 6), dtype=torch.float32, device=r.device)
    status = torch.empty((nq,), dtype=torch.uint8, device=r.device)
    with torch.cuda.device(r.device):
        check(lib.pct_quadric_fit(ptr(r), nq, k, ptr(coeffs), ptr(status), _stream()))
    return coeffs, status


def quadric_curvature(coeffs_dev):
    c = coeffs_dev.to(torch.float32).contiguous()
    curv = torch.empty((c.shape[0], 5), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        check(lib.pct_quadric_curvature(ptr(c), int(c.shape[0]), ptr(curv), _stream()))
    return curv


def implicit_quadric_fit(points_dev=None, idx_dev=None, query_ids=None, centered_dev=None):
    """Unit-norm minimiser of |A c|^2 per neighbourhood (pct_implicit_quadric_fit): rows of original indices on an
    (N, 3) cloud, or ``centered_dev`` (nq, k, 3) float32.  Returns (nq, 10) float64."""
    if centered_dev is not None:
        c = centered_dev.to(torch.float32).contiguous()
        nq, k, _ = c.shape
        out = torch.empty((nq, 10), dtype=torch.float64, device=c.device)
        with torch.cuda.device(c.device):
            check(lib.pct_implicit_quadric_fit(None, 0, None, nq, k, None, ptr(c), ptr(out), _stream()))
        return out
    if points_dev.shape[1] != 3 or not points_dev.is_contiguous():
        raise ValueError("implicit_quadric_fit needs packed (N, 3) points")
    idx_dev = idx_dev.to(torch.int32).contiguous()
    nq, k = idx_dev.shape
    out = torch.empty((nq, 10), dtype=torch.float64, device=points_dev.device)
    with torch.cuda.device(points_dev.device):
        check(lib.pct_implicit_quadric_fit(ptr(points_dev), int(points_dev.shape[0]), ptr(idx_dev), nq, k, ptr(query_ids), None,
                                           ptr(out), _stream()))
    return out


def implicit_quadric_curvature(coeffs_dev):
    """(nq, 10) float64 -> (nq, 4) float64 [K_g, K_h, k1, k2] by the reference's formulas (ref :435-480)."""
    c = coeffs_dev.to(torch.float64).contiguous()
    out = torch.empty((c.shape[0], 4), dtype=torch.float64, device=c.device)
    with torch.cuda.device(c.device):
        check(lib.pct_implicit_quadric_curvature(ptr(c), int(c.shape[0]), ptr(out), _stream()))
    return out


def pca_from_neighbors(points_dev, idx_dev, include_self=False, query_ids=None, want_directions=True):
    """PCA of the neighbourhood rows (nq, k): ``values (nq, 6)`` float64 = [l1, l2, l3, l1*l2, (l1+l2)/2,
    l3/(l1+l2+l3+1e-10)] and ``directions (nq, 3, 2)`` float64 (ref :901-945)."""
    if points_dev.shape[1] != 3 or not points_dev.is_contiguous():
        raise ValueError("pca_from_neighbors needs packed (N, 3) points")
    idx_dev = idx_dev.to(torch.int32).contiguous()
    nq, k = idx_dev.shape
    values = torch.empty((nq, 6), dtype=torch.float64, device=points_dev.device)
    directions = torch.empty((nq, 3, 2), dtype=torch.float64, device=points_dev.device) if want_directions else None
    with torch.cuda.device(points_dev.device):
        check(lib.pct_pca_from_neighbors(ptr(points_dev), int(points_dev.shape[0]), ptr(idx_dev), nq, k, int(bool(include_self)),
                                         ptr(query_ids), ptr(values), ptr(directions), _stream()))
    return values, directions


def mesh_energies(vertices_dev, triangles_dev, gaussian_dev=None, mean_dev=None):
    """Device tensor of 4 float64: bending, stretching, total area, triangles with an index out of range
    (utils.py:702-765)."""
    require_cuda()
    v = vertices_dev.to(torch.float32).contiguous()
    t = triangles_dev.to(torch.int32).contiguous()
    if v.ndim != 2 or v.shape[1] != 3 or t.ndim != 2 or t.shape[1] != 3:
        raise ValueError("vertices must be (V, 3) and triangles (T, 3)")
    g = None if gaussian_dev is None else gaussian_dev.to(torch.float32).contiguous()
    m = None if mean_dev is None else mean_dev.to(torch.float32).contiguous()
    for c in (g, m):
        if c is not None and c.numel() < v.shape[0]:
            raise IndexError(f"index {c.numel()} is out of bounds for axis 0 with size {c.numel()}")
    out = torch.empty(4, dtype=torch.float64, device=v.device)
    with torch.cuda.device(v.device):
        check(lib.pct_mesh_energies(ptr(v), int(v.shape[0]), ptr(t), int(t.shape[0]), ptr(g), ptr(m), ptr(out), _stream()))
    return out
