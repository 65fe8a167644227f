"""Drop-in ``PointCloud`` for the curvature hot path, running on a B200.

Same constructor, method names, attribute names, dtypes and error behaviour as
``PointCloud`` in /root/reference/pointCloudToolbox.py for the path

    PointCloud(...) -> plant_kdtree(k) -> fit_explicit_quadratic_surfaces_to_neighborhoods()
                    -> calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
    (or compute_pointwise_explicit_quadratic_curvature())

plus the three static per-neighbourhood methods.  Everything numerical happens
in libpct_b200.so on the GPU; there is no CPU fallback.  Differences a caller can
observe, all additive or representational:

* ``dists`` / ``neighbor_indices`` are materialised on first access.  When they
  are never read (nor assigned), fit + curvature run as ONE fused kernel and the
  neighbour lists never exist in memory; when they are read or assigned the fit
  uses exactly those rows, like the reference (ref :640).
* ``quadratic_coefficients`` is an (N, 6) float32 array and ``K_quadratic``,
  ``H_quadratic``, ``K_H_sq_quadratic`` are float32 arrays instead of Python
  lists of numpy scalars (same indexing, iteration and ``np.array()`` behaviour).
* equal-distance neighbours are ordered by index (scipy's order is an artefact
  of its tree traversal).
* new: ``plant_ball(radius)`` (the epsilon-ball query README.md:8 advertises),
  ``normals_quadratic``, ``k1_quadratic``, ``k2_quadratic``, ``fit_status``.
* the matrix norms of ``__init__`` (ref :45-47, an SVD of the N x 3 array) are
  computed on first access, on the device.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import engine, trace
from ._lib import STATUS_FEW_NEIGHBORS, STATUS_NONFINITE, MAX_K


class _Tree:
    """What ``self.kdtree`` holds: the device grid index plus a scipy-like ``query``."""

    def __init__(self, cloud, index):
        self._cloud_ref = weakref.ref(cloud)  # no reference cycle: the cloud frees its device memory on refcount
        self.index = index
        self.n = index.n
        self.m = 3

    def query(self, x, k=1):
        """``scipy.spatial.cKDTree.query(x, k)`` for the uses of the reference (ref :83, :759) and beyond: ``x`` is one
        point (3,) or an array (m, 3) of ARBITRARY coordinates; returns ``(dists float64, indices int64)`` of the k
        nearest cloud points, nearest first -- a query that is a cloud point finds itself first, at distance 0.
        Shapes follow scipy: (k,) / (m, k), and the last axis is dropped for ``k == 1``.  Equal distances are ordered
        by index (scipy's order is an artefact of its tree traversal)."""
        x = np.asarray(x, dtype=np.float32)
        single = x.ndim == 1
        pts = np.ascontiguousarray(np.atleast_2d(x))
        if pts.shape[1] != 3:
            raise ValueError(f"x must be of shape (3,) or (m, 3), got {x.shape}")
        kk = int(k)
        if kk < 1:
            raise ValueError("k must be at least 1")
        if kk > self.n:
            raise IndexError(f"index {self.n} is out of bounds for axis 0 with size {self.n}")  # scipy pads with N, the reference trips over it
        d, i = self.index.query(torch.from_numpy(pts).to(self.index.device), kk)
        d = engine.to_host(d)
        i = engine.to_host(i).astype(np.int64)
        if kk == 1:
            d, i = d[:, 0], i[:, 0]
        return (d[0], i[0]) if single else (d, i)


class PointCloud:
    def __init__(self, file_path=None, points=None, normals=None, downsample=False, voxel_size=0, k_neighbors=20,
                 output_path='./output/', max_points_per_voxel=1, device=None):
        self.downsample = downsample
        self.k_neighbors = k_neighbors
        self.voxel_size = voxel_size
        self.max_points_per_voxel = max_points_per_voxel
        self.output_path = output_path
        self.random_indexes = []
        self._device = device
        self._reset_results()

        if file_path:
            self.file_path = file_path
            self.read_from_file()
        elif points is not None and normals is not None:
            self.points = points
            self.normals = normals
        else:
            raise ValueError("Either file_path or points and normals must be provided")  # ref :41

        self.num_points = len(self.points)
        self.num_features = len(self.points[0])

    # ------------------------------------------------------------------
    # state
    # ------------------------------------------------------------------
    def _reset_results(self):
        self.kdtree = None
        self._d_points = None
        self._lists = None          # (idx int32 (N,k), dist float32 (N,k)) numpy, once materialised
        self._lists_dev = None
        self._user_lists = False
        self._ball_radius = None
        self._ball_csr = None
        self._fit = None            # engine.FitOutputs of the last fit
        self._norms = None

    def _host_points(self):
        p = self.points
        if isinstance(p, torch.Tensor):
            p = p.detach().cpu().numpy()
        return np.asarray(p)

    def _upload(self):
        self._d_points = engine.to_device_points(self.points, self._device)
        return self._d_points

    # ------------------------------------------------------------------
    # loading                                                   ref :50-66
    # ------------------------------------------------------------------
    def read_from_file(self):
        engine.require_cuda()
        # np.loadtxt (ref :51) replaced by the library's memory-mapped multi-threaded parser; the table is
        # rounded to float32 (ref :52-53) straight into pinned memory and goes to the device from there
        table = engine.load_text_f32(self.file_path)
        if table.ndim != 2:
            # np.loadtxt squeezes a one-row (or empty) file to one dimension and ref :52 fails on it
            raise IndexError(f"too many indices for array: array is {table.ndim}-dimensional, but 2 were indexed")
        dev = torch.device(self._device) if self._device is not None else torch.device("cuda", torch.cuda.current_device())
        t = table.to(dev, non_blocking=True)
        pts = t[:, 0:3].contiguous()                          # ref :52
        nrm = t[:, 3:6].contiguous()                          # ref :53
        pts[:, 0] -= pts[:, 0].max()                          # ref :56 (fp32)
        pts[:, 1] -= pts[:, 1].max()                          # ref :57
        if self.downsample:
            # ref :59-60 calls a method that only exists as a comment (ref :159-193)
            raise AttributeError("'PointCloud' object has no attribute 'downsample_point_cloud_by_grid'")
        lo = pts.min(0).values.cpu().numpy()
        hi = pts.max(0).values.cpu().numpy()
        # host copies live in page-locked memory, so the upload of plant_kdtree runs at PCIe speed
        self.points = engine.to_host(pts)
        self.normals = engine.to_host(nrm) if nrm.numel() else np.zeros(tuple(nrm.shape), np.float32)
        self.x_domain = [lo[0], hi[0]]                        # ref :64-66
        self.y_domain = [lo[1], hi[1]]
        self.z_domain = [lo[2], hi[2]]

    # matrix norms of the N x 3 array                           ref :45-47
    def _matrix_norms(self):
        if self._norms is None:
            p = engine.to_device_points(self.points, self._device).double()
            l1 = p.abs().sum(0).max()
            linf = p.abs().sum(1).max()
            l2 = torch.linalg.eigvalsh(p.T @ p).max().clamp_min(0).sqrt()
            dt = self._host_points().dtype
            self._norms = tuple(np.asarray(v.item()).astype(dt if dt.kind == "f" else np.float64)[()] for v in (l1, l2, linf))
        return self._norms

    @property
    def l1_norm(self):
        return self._matrix_norms()[0]

    @property
    def l2_norm(self):
        return self._matrix_norms()[1]

    @property
    def infinity_norm(self):
        return self._matrix_norms()[2]

    # ------------------------------------------------------------------
    # neighbour search                                          ref :69-89
    # ------------------------------------------------------------------
    def plant_kdtree(self, k_neighbors):
        k_neighbors = int(k_neighbors)
        if k_neighbors < 1 or k_neighbors > MAX_K:
            raise ValueError(f"k_neighbors must be in [1, {MAX_K}]")
        n = len(self.points)
        if k_neighbors + 1 > n:
            # scipy pads with index N, which the reference trips over at ref :640
            raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
        self.k_neighbors = k_neighbors
        with trace.stage("h2d"):
            d_points = self._upload()
        with trace.stage("index_build"):
            index = engine.GridIndex(d_points, k_hint=k_neighbors)
        self.kdtree = _Tree(self, index)
        self._lists = None
        self._lists_dev = None
        self._user_lists = False
        self._ball_radius = None
        self._ball_csr = None

    def _materialise_lists(self):
        if self._lists is None:
            if self.kdtree is None:
                raise AttributeError("'PointCloud' object has no attribute 'neighbor_indices'")
            idx, dist = self.kdtree.index.knn(self.k_neighbors)
            self._lists_dev = idx
            self._lists = (engine.to_host(idx), engine.to_host(dist))
        return self._lists

    @property
    def neighbor_indices(self):
        return self._materialise_lists()[0]

    @neighbor_indices.setter
    def neighbor_indices(self, value):
        dist = self._lists[1] if self._lists is not None else None
        self._lists = (np.asarray(value), dist)
        self._lists_dev = None
        self._user_lists = True

    @property
    def dists(self):
        return self._materialise_lists()[1]

    @dists.setter
    def dists(self, value):
        idx = self._lists[0] if self._lists is not None else None
        self._lists = (idx, np.asarray(value))

    # epsilon-ball (README.md:8; the reference has no implementation)
    def plant_ball(self, radius):
        radius = float(radius)
        if not radius > 0:
            raise ValueError("radius must be positive")
        d_points = self._upload()
        index = engine.GridIndex(d_points, cell_hint=radius * 1.001)
        self.kdtree = _Tree(self, index)
        self._ball_radius = radius
        self._ball_csr = None
        self._lists = None
        self._lists_dev = None
        self._user_lists = False

    def ball_neighbors(self):
        """CSR ``(offsets int64, indices int32, dists float32)`` of the planted ball query."""
        if self._ball_radius is None:
            raise AttributeError("plant_ball(radius) has not been called")
        if self._ball_csr is None:
            off, idx, dist = self.kdtree.index.ball(self._ball_radius)
            self._ball_csr = (engine.to_host(off), engine.to_host(idx), engine.to_host(dist))
        return self._ball_csr

    # ------------------------------------------------------------------
    # fit                                                       ref :635-647
    # ------------------------------------------------------------------
    def fit_explicit_quadratic_surfaces_to_neighborhoods(self):
        if self.kdtree is None and self._lists is None:
            raise AttributeError("'PointCloud' object has no attribute 'neighbor_indices'")
        if self.kdtree is not None and self._ball_radius is None and self.k_neighbors < 6 and self._lists is None:
            # fewer rows than coefficients: lstsq's minimum-norm solution (ref :359) lives in the rows path
            self._materialise_lists()
        if self._lists is not None and self._lists[0] is not None:
            # rows were read or assigned: fit exactly those rows (ref :640)
            d_points = self._d_points if self._d_points is not None else self._upload()
            idx = self._lists_dev
            n = len(self.points)
            if idx is None:
                rows = np.asarray(self._lists[0])
                if rows.ndim != 2:
                    raise ValueError("neighbor_indices must have shape (num_points, k)")
                if rows.shape[0] < n:
                    # the reference loops over all points and indexes row i (ref :638-640)
                    raise IndexError(f"index {rows.shape[0]} is out of bounds for axis 0 with size {rows.shape[0]}")
                idx = torch.from_numpy(np.ascontiguousarray(rows[:n], dtype=np.int32)).to(d_points.device)
                if idx.numel():
                    lo, hi = int(idx.min()), int(idx.max())
                    if hi >= n or lo < -n:
                        raise IndexError(f"index {hi if hi >= n else lo} is out of bounds for axis 0 with size {n}")
                    if lo < 0:
                        idx = torch.where(idx < 0, idx + n, idx)   # numpy's negative indices (ref :640 is plain fancy indexing)
            fit = engine.fit_from_neighbors(d_points, idx)
        else:
            fit = self._fused_fit(want_coeffs=True)
        self._raise_on_nonfinite(fit)
        self._fit = fit
        self._lazy_fit_to_host = False
        self.quadratic_coefficients = engine.to_host(fit.coeffs)
        self._coeffs_of_fit = self.quadratic_coefficients
        self.normals_quadratic = engine.to_host(fit.normals)
        self.fit_status = engine.to_host(fit.status)

    @staticmethod
    def _raise_on_nonfinite(fit):
        status = fit.status
        if status is not None and bool((status & STATUS_NONFINITE).any()):
            raise ValueError("Non-finite values after rotation")  # ref :318-319 / :356-357

    # ------------------------------------------------------------------
    # curvature                                                 ref :657-674, :505-509
    # ------------------------------------------------------------------
    def calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points(self):
        # always from the STORED coefficients, like the reference (ref :663-672): an in-place edit such as
        # pc.quadratic_coefficients[i] = 0 (the pattern of ref :768-769) must show in K and H
        coeffs = np.ascontiguousarray(np.asarray(self.quadratic_coefficients, dtype=np.float32)).reshape(-1, 6)
        engine.require_cuda()
        dev = self._d_points.device if self._d_points is not None else "cuda"
        curv = engine.quadric_curvature(torch.from_numpy(coeffs).to(dev))
        c = engine.to_host(curv)
        self.K_quadratic = c[:, 0].copy()
        self.H_quadratic = c[:, 1].copy()
        self.k1_quadratic = c[:, 2].copy()
        self.k2_quadratic = c[:, 3].copy()
        self.K_H_sq_quadratic = c[:, 4].copy()
        return self.K_quadratic, self.H_quadratic

    def compute_pointwise_explicit_quadratic_curvature(self):
        """Compute explicit quadratic curvature and return pointwise values (ref :505-509)."""
        if (self._lists is None or self._lists[0] is None) and self.kdtree is not None and \
                (self.k_neighbors >= 6 or self._ball_radius is not None):
            # throughput path: one fused kernel, and only K and H have to come back to the host
            with trace.stage("fused_kernel"):
                fit = self._fused_fit(want_coeffs=False)
            with trace.stage("d2h"):
                kh = engine.to_host(fit.kh())
            for name in ("quadratic_coefficients", "normals_quadratic", "fit_status", "K_H_sq_quadratic",
                         "k1_quadratic", "k2_quadratic"):
                self.__dict__.pop(name, None)
            self._lazy_fit_to_host = True
            self.K_quadratic, self.H_quadratic = kh[0], kh[1]
            return np.asarray(self.K_quadratic), np.asarray(self.H_quadratic)
        self.fit_explicit_quadratic_surfaces_to_neighborhoods()
        K, H = self.calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points()
        return np.array(K), np.array(H)

    def _fused_fit(self, want_coeffs=True):
        index = self.kdtree.index
        if self._ball_radius is not None:
            fit = index.curvature_ball(self._ball_radius)
        else:
            fit = index.curvature_knn(self.k_neighbors, want_coeffs=want_coeffs)
        self._raise_on_nonfinite(fit)
        self._fit = fit
        return fit

    # ------------------------------------------------------------------
    # neighbour study                                           ref :732-800
    # ------------------------------------------------------------------
    def explicit_quadratic_neighbor_study(self, tol=1e-7, sample_size=500, lower_bound=3, upper_bound=99):
        """Average neighbour count at which the quadric Gaussian curvature of sampled points converges.

        Same sampling (``np.random.randint`` on the global generator), same probes and the same binary
        search as the reference (ref :732-800).  Every probe the searches can ask for -- K of the
        neighbourhood {point, its n nearest}, n = lower_bound .. upper_bound + 1 -- is computed in two
        device calls (ordered kNN rows of the sampled points, then one fit per (point, n) prefix row);
        the searches then only read that table.
        """
        if self.kdtree is None:
            raise AttributeError("'PointCloud' object has no attribute 'kdtree'")
        num_total = len(self.points)
        sample_size = min(sample_size, num_total)
        random_indexes = np.random.randint(0, num_total, sample_size)          # ref :751
        if len(random_indexes) == 0:
            return 0                                                             # ref :796-797
        lower, upper = int(lower_bound), int(upper_bound)
        n_max = upper + 1                                                        # the search probes mid and mid + 1
        if lower > upper:
            return int(np.mean(np.full(len(random_indexes), upper))) + 1         # no probe at all: best = upper (ref :791-792)
        if n_max + 1 > num_total:
            raise IndexError(f"index {num_total} is out of bounds for axis 0 with size {num_total}")  # scipy pads with N
        if n_max > MAX_K:
            raise ValueError(f"upper_bound + 1 must not exceed {MAX_K}")
        index = self.kdtree.index
        d_points = index.points
        dev = d_points.device
        uniq, inverse = np.unique(random_indexes, return_inverse=True)
        ids = torch.from_numpy(uniq.astype(np.int32)).to(dev)
        nbr, _ = index.knn_points(ids, n_max)                                    # (S, n_max) nearest OTHER points, ordered
        # one CSR row per (point, n): [point, nbr_1 .. nbr_n]  -- kdtree.query(point, n + 1) of ref :759
        counts = torch.arange(lower, n_max + 1, device=dev)                      # n values
        S, P = len(uniq), int(counts.numel())
        row_len = (counts + 1).repeat(S)
        offsets = torch.zeros(S * P + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(row_len, 0)
        full = torch.cat((ids[:, None], nbr), 1)                                 # (S, n_max + 1)
        col = torch.arange(n_max + 1, device=dev)
        mask = col[None, :] <= counts[:, None]                                   # (P, n_max + 1): first n + 1 entries
        rows = full[:, None, :].expand(S, P, n_max + 1)[mask[None].expand(S, P, n_max + 1)]
        qids = ids.repeat_interleave(P)
        fit = engine.fit_from_csr(d_points if d_points.shape[1] == 3 else d_points[:, :3].contiguous(), offsets, rows, qids)
        K = engine.to_host(fit.curv[:, 0]).reshape(S, P)
        st = engine.to_host(fit.status).reshape(S, P)
        # ref :768-769: a failed fit becomes coefficients (0,...,0), i.e. K = 0
        K = np.where((st & np.uint8(STATUS_NONFINITE | STATUS_FEW_NEIGHBORS)) != 0, np.float32(0), K)
        converged = []
        for s_row in inverse:                                                    # ref :794-795, one search per sampled point
            lo, hi, best = lower, upper, None
            k_of = K[s_row]
            while lo <= hi:                                                      # ref :778-788
                mid = (lo + hi) // 2
                if abs(k_of[mid + 1 - lower] - k_of[mid - lower]) < tol:
                    best = mid
                    hi = mid - 1
                else:
                    lo = mid + 1
            if best is None:
                best = hi
            converged.append(best)
        return int(np.mean(converged)) + 1                                       # ref :800

    # ------------------------------------------------------------------
    # PCA "principal curvatures"                                ref :901-945
    # ------------------------------------------------------------------
    def principal_curvatures_via_principal_component_analysis(self, k_neighbors):
        """Eigenvalues of every neighbourhood's covariance as the reference reports them (ref :901-945).

        The reference ranks all N points by distance for every point (O(N^2)); here the rows come from the grid
        index (exact kNN, the point itself dropped like ``sorted_indices[1:k + 1]``) and one kernel does
        ``np.cov`` + ``eigh`` per row in fp64.  Sets the same five float64 attributes; the sign of each
        direction is arbitrary, as it is in ``eigh``.
        """
        k = int(k_neighbors)
        if k < 1 or k > MAX_K:
            raise ValueError(f"k_neighbors must be in [1, {MAX_K}]")
        n = len(self.points)
        d_points = self._upload()
        if d_points.shape[1] != 3 or not d_points.is_contiguous():
            d_points = d_points[:, :3].contiguous()
        if self.kdtree is not None and self._ball_radius is None and self.kdtree.index.n == n:
            index = self.kdtree.index
        else:
            index = engine.GridIndex(d_points, k_hint=k)
        kk = min(k, n - 1)                                   # a slice past the end just ends (ref :916)
        if kk < 1:
            values = torch.full((n, 6), float("nan"), dtype=torch.float64)
            directions = torch.full((n, 3, 2), float("nan"), dtype=torch.float64)
        else:
            idx, _ = index.knn(kk, want_dist=False)
            values, directions = engine.pca_from_neighbors(d_points, idx)
        v = engine.to_host(values)
        self.pca_principal_curvature_values_1 = v[:, 0].copy()
        self.pca_principal_curvature_values_2 = v[:, 1].copy()
        self.principal_curvature_directions = engine.to_host(directions).copy()
        self.pca_K_values = v[:, 3].copy()
        self.pca_H_values = v[:, 4].copy()

    # ------------------------------------------------------------------
    # implicit 10-coefficient quadric              ref :363-396, :435-480, :617-633, :676-689
    # ------------------------------------------------------------------
    def fit_implicit_quadric_surfaces_all_points(self):
        """``quadric_coefficients`` (N, 10) float64 of every point's neighbourhood -- the point itself and its
        k - 1 nearest, centred on the point (ref :617-633 queries the tree with k, not k + 1).

        PARITY UNPINNED: the reference minimises |A c|^2 on the unit sphere with SLSQP from the all-ones start and
        stops far from the minimiser; this returns the minimiser itself (smallest eigenvector of A^T A), signed so
        that the gradient at the point looks away from the neighbours' centroid.  Returns what the reference
        returns: ``(self.K_quadratic, self.H_quadratic)`` (sic, ref :633)."""
        if self.kdtree is None:
            raise AttributeError("'PointCloud' object has no attribute 'kdtree'")
        k = int(self.k_neighbors)
        index = self.kdtree.index
        d_points = index.points if index.points.shape[1] == 3 else index.points[:, :3].contiguous()
        n = len(self.points)
        if k > n:
            raise IndexError(f"index {n} is out of bounds for axis 0 with size {n}")
        me = torch.arange(n, dtype=torch.int32, device=d_points.device)[:, None]
        if k > 1:
            others, _ = index.knn(k - 1, want_dist=False)
            rows = torch.cat((me, others), 1)
        else:
            rows = me
        coeffs = engine.implicit_quadric_fit(d_points, rows)
        self._implicit_dev = coeffs
        self.quadric_coefficients = engine.to_host(coeffs)
        return self.K_quadratic, self.H_quadratic                       # ref :633

    def calculate_curvatures_of_implicit_quadric_surfaces_for_all_points(self):
        """``K_quadric`` / ``H_quadric`` from the stored ``quadric_coefficients`` by the reference's formulas (ref :676-689)."""
        coeffs = np.ascontiguousarray(np.asarray(self.quadric_coefficients, dtype=np.float64)).reshape(-1, 10)
        engine.require_cuda()
        dev = self._d_points.device if self._d_points is not None else "cuda"
        curv = engine.to_host(engine.implicit_quadric_curvature(torch.from_numpy(coeffs).to(dev)))
        self.K_quadric = curv[:, 0].copy()
        self.H_quadric = curv[:, 1].copy()

    def compute_pointwise_implicit_quadric_curvature(self):
        """ref :511-515"""
        self.fit_implicit_quadric_surfaces_all_points()
        self.calculate_curvatures_of_implicit_quadric_surfaces_for_all_points()
        return np.array(self.K_quadric), np.array(self.H_quadric)

    def __getattr__(self, name):
        # outputs of the fused call are copied to the host only when somebody asks for them
        lazy = ("quadratic_coefficients", "normals_quadratic", "fit_status", "K_H_sq_quadratic", "k1_quadratic", "k2_quadratic")
        if name in lazy and self.__dict__.get("_lazy_fit_to_host") and self.__dict__.get("_fit") is not None:
            fit = self.__dict__["_fit"]
            if name == "quadratic_coefficients":
                if fit.coeffs is None:  # the throughput call skipped them: run the fused kernel once more, with them
                    fit = self._fused_fit(want_coeffs=True)
                val = engine.to_host(fit.coeffs)
                self.__dict__["_coeffs_of_fit"] = val
            elif name == "normals_quadratic":
                val = engine.to_host(fit.normals)
            elif name == "fit_status":
                val = engine.to_host(fit.status)
            else:
                col = {"k1_quadratic": "k1", "k2_quadratic": "k2", "K_H_sq_quadratic": "H2"}[name]
                val = engine.to_host(fit.column(col))
            self.__dict__[name] = val
            return val
        raise AttributeError(f"'PointCloud' object has no attribute '{name}'")

    # ------------------------------------------------------------------
    # per-neighbourhood static methods              ref :270-321, :331-360, :398-431
    # ------------------------------------------------------------------
    @staticmethod
    def get_best_fit_plane_and_rotate(points):
        pts = np.asarray(points)
        if not np.all(np.isfinite(pts)):
            raise ValueError("Non-finite values in input points")
        engine.require_cuda()
        c = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)[None]).cuda()
        rotated, _, status = engine.plane_rotate(c)
        if int(status[0]) & STATUS_NONFINITE:
            raise ValueError("Non-finite values after rotation")
        return rotated[0].cpu().numpy()

    @staticmethod
    def fit_quadratic_surface(points):
        pts = np.array(points, dtype=np.float32)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise ValueError("Input points must have shape (N, 3)")
        if not np.all(np.isfinite(pts)):
            raise ValueError("Input contains non-finite values.")
        engine.require_cuda()
        coeffs, _ = engine.quadric_fit(torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)[None]).cuda())
        return coeffs[0].cpu().numpy()

    @staticmethod
    def fit_implicit_quadric_surface(points):
        """(k, 3) centred points -> 10 coefficients of unit norm minimising |A c|^2 (ref :363-396; see
        fit_implicit_quadric_surfaces_all_points for how this differs from the reference's SLSQP result)."""
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32))
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise ValueError("Input points must have shape (N, 3)")
        engine.require_cuda()
        return engine.implicit_quadric_fit(centered_dev=torch.from_numpy(pts[None]).cuda())[0].cpu().numpy()

    @staticmethod
    def calculate_implicit_quadric_curvatures(coefficients):
        """(K_g, K_h, k1, k2) of the implicit quadric at the origin, exactly the reference's formulas (ref :435-480)."""
        c = np.asarray(coefficients, dtype=np.float64).reshape(1, 10)
        engine.require_cuda()
        out = engine.implicit_quadric_curvature(torch.from_numpy(c).cuda())[0].cpu().numpy()
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def calculate_explicit_quadratic_curvatures(coefficients):
        c = np.asarray(coefficients, dtype=np.float32).reshape(1, 6)
        engine.require_cuda()
        out = engine.quadric_curvature(torch.from_numpy(c).cuda())[0].cpu().numpy()
        return out[0], out[1], out[2], out[3], out[4]
