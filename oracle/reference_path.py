"""numpy/scipy restatement of the reference's per-point curvature pipeline.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
line range of ``/root/reference/pointCloudToolbox.py`` it follows ("ref :a-b").

The restatement is deliberately written the way the reference computes -- same
dtypes at every step, same third-party calls -- because the GPU path is judged
on *numerical* parity, not on the maths alone:

    fp32 points -> (scipy, fp64) kNN -> fp32 centring -> fp64 covariance ->
    fp64 SVD -> orientation by (farthest - nearest) -> fp64 Rodrigues rotation
    -> fp32 quantisation -> fp32 design matrix -> fp64 min-norm least squares
    -> fp32 coefficients -> fp32 curvature formulas.

Two things are added on top of the reference because a comparison needs them
("oracle wrapper duties"):

* canonical tie order: scipy's cKDTree returns equal-distance neighbours in
  tree-traversal order; the contract here is "(squared distance, index)".
* the oriented unit normal, which the reference computes and throws away.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

__all__ = [
    "load_points",
    "squared_distance_key",
    "knn_canonical",
    "ball_canonical",
    "best_fit_plane_and_rotate",
    "fit_quadratic_surface",
    "explicit_quadratic_curvatures",
    "neighbourhood_pipeline",
    "curvature_from_neighbors",
    "curvature_from_neighbors_batched",
    "curvature_from_csr",
    "knn_curvature",
]


# --------------------------------------------------------------------------
# input conditioning                                            ref :50-57
# --------------------------------------------------------------------------
def load_points(file_path):
    """Text file -> (points f32 (N,3), normals f32 (N,0..3)) with the max-shift.

    ref :51-53  columns 0:3 are the points, 3:6 the normals, both cast to fp32.
    ref :56-57  x and y are shifted by their maxima *in fp32, in place*.
    """
    table = np.loadtxt(file_path)
    pts = table[:, 0:3].astype(np.float32)
    nrm = table[:, 3:6].astype(np.float32)
    pts[:, 0] -= np.max(pts[:, 0])
    pts[:, 1] -= np.max(pts[:, 1])
    return pts, nrm


# --------------------------------------------------------------------------
# neighbour search                                               ref :69-85
# --------------------------------------------------------------------------
def squared_distance_key(p64, q64):
    """The ranking key of scipy 1.18's cKDTree for p=2, m=3.

    scipy/spatial/ckdtree/src/distance.h ``sqeuclidean_distance_double`` on a
    3-vector reduces to ``s = 0; s += dx*dx; s += dy*dy; s += dz*dz`` with
    separate multiply and add (x86-64 baseline, no FMA contraction), i.e.
    ``(dx*dx + dy*dy) + dz*dz`` on the fp64 images of the fp32 coordinates.
    SURVEY.md section 7.3(1) verified this bit-for-bit on every bunny pair; the
    golden tests re-verify it.
    """
    d = np.asarray(p64, np.float64) - np.asarray(q64, np.float64)
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def _row_lexsort(d2, idx):
    """Per-row order by (d2, idx)."""
    by_idx = np.argsort(idx, axis=1, kind="stable")
    d2_i = np.take_along_axis(d2, by_idx, 1)
    by_d = np.argsort(d2_i, axis=1, kind="stable")
    return np.take_along_axis(by_idx, by_d, 1)


def knn_canonical(points, k, rows=None, tree=None, workers=-1):
    """k nearest neighbours of cloud points the way ``plant_kdtree`` defines them.

    ref :74     the tree is built on the fp32 cloud (scipy stores fp64 copies).
    ref :83-85  query(point, k+1), the first hit is dropped as "self".
    ref :78-79  distances are stored as fp32, indices as int32.

    Canonicalisation (not reference behaviour): the (k+1)-list is taken as the
    first k+1 entries of *all* cloud points ordered by (d2 fp64, index); scipy
    itself resolves equal keys by traversal order.  Returns
    ``idx (n,k) int32, dist (n,k) float32, d2 (n,k) float64``.
    """
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n_all = len(pts)
    if k + 1 > n_all:
        raise IndexError("k+1 exceeds the number of points (ref :640 IndexError)")
    p64 = pts.astype(np.float64)
    if tree is None:
        tree = cKDTree(pts)
    if rows is None:
        rows = np.arange(n_all)
    rows = np.asarray(rows, dtype=np.int64)
    nq = len(rows)
    out_idx = np.empty((nq, k), np.int32)
    out_d2 = np.empty((nq, k), np.float64)
    todo = np.arange(nq)
    extra = 8
    while len(todo):
        kk = min(k + 1 + extra, n_all)
        _, nb = tree.query(p64[rows[todo]], kk, workers=workers)
        nb = nb.reshape(len(todo), kk)
        d2 = squared_distance_key(p64[nb], p64[rows[todo]][:, None, :])
        order = _row_lexsort(d2, nb)
        nb = np.take_along_axis(nb, order, 1)
        d2 = np.take_along_axis(d2, order, 1)
        # the tie group of the (k+1)-th entry must lie fully inside what we fetched
        complete = (d2[:, -1] > d2[:, k]) | (kk == n_all)
        out_idx[todo[complete]] = nb[complete, 1:k + 1]
        out_d2[todo[complete]] = d2[complete, 1:k + 1]
        todo = todo[~complete]
        extra *= 4
    dist = np.sqrt(out_d2).astype(np.float32)
    return out_idx, dist, out_d2


def ball_canonical(points, radius, rows=None, tree=None, workers=-1):
    """All j != i with ||p_j - p_i|| <= radius, ordered by (d2, index), as CSR.

    Extension (absent in the reference, README.md:8): membership is whatever
    ``cKDTree.query_ball_point(x, r)`` returns -- for p=2 that is
    ``d2 <= r*r`` on the key above (ckdtree/src/query_ball_point.cxx compares
    the squared distance with the squared radius, inclusive).
    Returns ``offsets (n+1) int64, idx (nnz) int32, dist (nnz) float32``.
    """
    pts = np.ascontiguousarray(points, dtype=np.float32)
    p64 = pts.astype(np.float64)
    if tree is None:
        tree = cKDTree(pts)
    if rows is None:
        rows = np.arange(len(pts))
    rows = np.asarray(rows, dtype=np.int64)
    hits = tree.query_ball_point(p64[rows], float(radius), workers=workers, return_sorted=False)
    counts = np.zeros(len(rows) + 1, np.int64)
    idx_parts, dist_parts = [], []
    for r, (i, h) in enumerate(zip(rows, hits)):
        h = np.asarray(h, dtype=np.int64)
        h = h[h != i]
        d2 = squared_distance_key(p64[h], p64[i])
        order = np.lexsort((h, d2))
        idx_parts.append(h[order].astype(np.int32))
        dist_parts.append(np.sqrt(d2[order]).astype(np.float32))
        counts[r + 1] = len(h)
    offsets = np.cumsum(counts)
    idx = np.concatenate(idx_parts) if idx_parts else np.zeros(0, np.int32)
    dist = np.concatenate(dist_parts) if dist_parts else np.zeros(0, np.float32)
    return offsets, idx, dist


# --------------------------------------------------------------------------
# per-neighbourhood functions                     ref :270-321, :331-360, :398-431
# --------------------------------------------------------------------------
def best_fit_plane_and_rotate(centered, return_normal=False):
    """PCA plane of a centred neighbourhood, oriented, rotated so normal -> +z.

    ref :273-274  non-finite input -> ValueError
    ref :277      covariance about the neighbourhood mean, ddof=1, fp64
    ref :280-283  SVD of the 3x3 covariance, normal = last right singular vector
    ref :286-297  reference vector = last point - first point (farthest minus
                  nearest neighbour); flip the normal when the dot product of
                  the two unit vectors is negative (NaN compares False)
    ref :300-312  Rodrigues matrix taking the normal to +z, identity if the
                  cross product has zero length (this includes normal = -z)
    ref :315-319  rotate in fp64; non-finite result -> ValueError
    """
    c = np.asarray(centered)
    if not np.isfinite(c).all():
        raise ValueError("Non-finite values in input points")
    cov = np.cov(c, rowvar=False)
    _, _, vt = np.linalg.svd(cov, full_matrices=True)
    normal = vt[-1]
    ref_vec = c[-1] - c[0]
    with np.errstate(invalid="ignore", divide="ignore"):
        n_hat = normal / np.linalg.norm(normal)
        r_hat = ref_vec / np.linalg.norm(ref_vec)
        if np.dot(n_hat, r_hat) < 0:
            normal = -normal
    a = normal / np.linalg.norm(normal)
    v = np.cross(a, np.array([0, 0, 1]))
    cos_t = a[2]
    sin_t = np.linalg.norm(v)
    if sin_t == 0:
        rot = np.eye(3)
    else:
        vx = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
        rot = np.eye(3) + vx + vx.dot(vx) * ((1 - cos_t) / (sin_t ** 2))
    rotated = (rot @ c.T).T
    if not np.isfinite(rotated).all():
        raise ValueError("Non-finite values after rotation")
    if return_normal:
        return rotated, a
    return rotated


def fit_quadratic_surface(rotated):
    """Least-squares z = A a^2 + B b^2 + C ab + D a + E b + F.

    ref :350      inputs quantised to fp32
    ref :351-357  shape / finiteness checks -> ValueError
    ref :358      design matrix [a^2, b^2, ab, a, b, 1] built in fp32
    ref :359      numpy.linalg.lstsq(rcond=None): LAPACK dgelsd in fp64 with
                  rcond = eps_fp64 * max(M, N), minimum-norm solution, result
                  cast back to fp32 (numpy/linalg/_linalg.py `_commonType`)
    """
    p = np.array(rotated, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("Input points must have shape (N, 3)")
    if not np.isfinite(p).all():
        raise ValueError("Input contains non-finite values.")
    a, b, z = p[:, 0], p[:, 1], p[:, 2]
    design = np.column_stack((a * a, b * b, a * b, a, b, np.ones_like(a))).astype(np.float32)
    coeffs = np.linalg.lstsq(design, z, rcond=None)[0]
    return coeffs


def explicit_quadratic_curvatures(coeffs):
    """Monge-patch curvatures of the fitted quadric at the origin, in fp32.

    ref :403-409  Fx = D, Fy = E, Fxx = 2A, Fyy = 2B, Fxy = C
    ref :412-419  K = (Fxx Fyy - Fxy^2) / g^2,
                  H = ((1+Fx^2) Fyy - 2 Fx Fy Fxy + (1+Fy^2) Fxx) / (2 g^1.5),
                  g = 1 + Fx^2 + Fy^2
    ref :425-429  k1, k2 = H +- sqrt(max(H^2 - K, 0))
    ref :431      returns (K, H, k1, k2, H^2)
    """
    a, b, c, d, e, _ = coeffs
    fx, fy = d, e
    fxx, fyy, fxy = 2 * a, 2 * b, c
    g = 1 + fx ** 2 + fy ** 2
    k_gauss = (fxx * fyy - fxy ** 2) / g ** 2
    k_mean = ((1 + fx ** 2) * fyy - 2 * fx * fy * fxy + (1 + fy ** 2) * fxx) / (2 * g ** 1.5)
    root = np.sqrt(max(k_mean ** 2 - k_gauss, 0))
    return k_gauss, k_mean, k_mean + root, k_mean - root, k_mean ** 2


def neighbourhood_pipeline(points, query_index, neighbor_index_row):
    """One iteration of the reference's fit loop + curvature loop.

    ref :640-641  gather the neighbours, subtract the query point (fp32)
    ref :644-647  rotate, fit
    ref :668      curvatures from the coefficients
    Returns ``(normal f64 (3,), coeffs f32 (6,), (K, H, k1, k2, H2) f32)``.
    """
    centered = points[neighbor_index_row] - points[query_index]
    rotated, normal = best_fit_plane_and_rotate(centered, return_normal=True)
    coeffs = fit_quadratic_surface(rotated)
    return normal, coeffs, explicit_quadratic_curvatures(coeffs)


def _empty_result(n):
    return {
        "normal": np.full((n, 3), np.nan, np.float64),
        "coeffs": np.full((n, 6), np.nan, np.float32),
        "K": np.full(n, np.nan, np.float32),
        "H": np.full(n, np.nan, np.float32),
        "k1": np.full(n, np.nan, np.float32),
        "k2": np.full(n, np.nan, np.float32),
        "H2": np.full(n, np.nan, np.float32),
        "margin": np.full(n, np.nan, np.float64),
    }


def _orientation_margin(centered, normal):
    """|n . r| for unit n and unit r = (last - first): how safe the sign choice is."""
    r = (centered[-1] - centered[0]).astype(np.float64)
    nr = np.linalg.norm(r)
    if nr == 0:
        return 0.0
    return abs(float(np.dot(normal, r / nr)))


def curvature_from_neighbors(points, neighbor_indices, rows=None):
    """Faithful per-point loop (ref :638-647 and :663-672) over ``rows``.

    ``neighbor_indices[r]`` is the neighbour row of cloud point ``rows[r]``.
    """
    pts = np.asarray(points, dtype=np.float32)
    if rows is None:
        rows = np.arange(len(neighbor_indices))
    out = _empty_result(len(rows))
    for r, i in enumerate(rows):
        nb = neighbor_indices[r]
        normal, coeffs, curv = neighbourhood_pipeline(pts, i, nb)
        out["normal"][r] = normal
        out["coeffs"][r] = coeffs
        out["K"][r], out["H"][r], out["k1"][r], out["k2"][r], out["H2"][r] = curv
        out["margin"][r] = _orientation_margin(pts[nb] - pts[i], normal)
    return out


def curvature_from_csr(points, offsets, indices, rows=None):
    """Same loop for variable-size (epsilon-ball) neighbourhoods given as CSR.

    Rows with fewer than 2 neighbours have no covariance; they stay NaN.
    """
    pts = np.asarray(points, dtype=np.float32)
    n = len(offsets) - 1
    if rows is None:
        rows = np.arange(n)
    out = _empty_result(n)
    for r, i in enumerate(rows):
        nb = indices[offsets[r]:offsets[r + 1]]
        if len(nb) < 2:
            continue
        normal, coeffs, curv = neighbourhood_pipeline(pts, i, nb)
        out["normal"][r] = normal
        out["coeffs"][r] = coeffs
        out["K"][r], out["H"][r], out["k1"][r], out["k2"][r], out["H2"][r] = curv
        out["margin"][r] = _orientation_margin(pts[nb] - pts[i], normal)
    return out


# --------------------------------------------------------------------------
# batched form of the same arithmetic (for 1e5..1e6-point checks)
# --------------------------------------------------------------------------
def curvature_from_neighbors_batched(points, neighbor_indices, rows=None, chunk=65536):
    """Vectorised over points; step-for-step the same dtypes as the loop above.

    Differences from the per-point functions are confined to fp64 round-off:
    the covariance is formed by an einsum instead of ``np.cov``'s matmul and the
    minimum-norm least-squares solution comes from ``np.linalg.pinv`` (gesdd,
    same rcond = eps*max(M,N) cut) instead of dgelsd.  ``tests/test_oracle.py``
    checks it against the per-point loop.
    """
    pts = np.asarray(points, dtype=np.float32)
    nbr = np.asarray(neighbor_indices)
    if rows is None:
        rows = np.arange(len(nbr))
    rows = np.asarray(rows)
    n, k = nbr.shape
    out = _empty_result(n)
    ez = np.array([0.0, 0.0, 1.0])
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        c32 = pts[nbr[s:e]] - pts[rows[s:e]][:, None, :]          # ref :640-641 (fp32)
        c = c32.astype(np.float64)
        mu = c.mean(axis=1, keepdims=True)
        x = c - mu
        cov = np.einsum("nki,nkj->nij", x, x) / (k - 1)           # ref :277
        _, _, vt = np.linalg.svd(cov)                             # ref :280
        normal = vt[:, -1, :]                                     # ref :283
        ref32 = c32[:, -1, :] - c32[:, 0, :]                      # ref :286 (fp32)
        with np.errstate(invalid="ignore", divide="ignore"):
            n_hat = normal / np.linalg.norm(normal, axis=1, keepdims=True)
            r_hat = ref32 / np.linalg.norm(ref32, axis=1, keepdims=True)   # fp32
            dot = np.einsum("ni,ni->n", n_hat, r_hat.astype(np.float64))
        normal = np.where((dot < 0)[:, None], -normal, normal)    # ref :293-297
        a = normal / np.linalg.norm(normal, axis=1, keepdims=True)
        v = np.cross(a, ez)                                       # ref :303
        cos_t = a[:, 2]
        sin_t = np.linalg.norm(v, axis=1)
        vx = np.zeros((e - s, 3, 3))
        vx[:, 0, 1], vx[:, 0, 2] = -v[:, 2], v[:, 1]
        vx[:, 1, 0], vx[:, 1, 2] = v[:, 2], -v[:, 0]
        vx[:, 2, 0], vx[:, 2, 1] = -v[:, 1], v[:, 0]
        with np.errstate(invalid="ignore", divide="ignore"):
            fac = (1 - cos_t) / (sin_t ** 2)
        rot = np.eye(3)[None] + vx + (vx @ vx) * fac[:, None, None]   # ref :312
        rot[sin_t == 0] = np.eye(3)                               # ref :308-309
        rotated = np.einsum("nij,nkj->nki", rot, c)               # ref :315
        p = rotated.astype(np.float32)                            # ref :350
        aa, bb, zz = p[..., 0], p[..., 1], p[..., 2]
        design = np.stack((aa * aa, bb * bb, aa * bb, aa, bb, np.ones_like(aa)), axis=-1)  # ref :358 fp32
        d64 = design.astype(np.float64)
        rcond = np.finfo(np.float64).eps * max(k, 6)
        pinv = np.linalg.pinv(d64, rcond=rcond)
        w = np.einsum("nck,nk->nc", pinv, zz.astype(np.float64)).astype(np.float32)  # ref :359
        A, B, C, D, E = (w[:, j] for j in range(5))
        one = np.float32(1)
        two = np.float32(2)
        g = one + D ** 2 + E ** 2                                 # ref :412-413 (fp32)
        fxx, fyy = two * A, two * B
        K = (fxx * fyy - C ** 2) / g ** 2
        H = ((one + D ** 2) * fyy - two * D * E * C + (one + E ** 2) * fxx) / (two * g ** np.float32(1.5))
        root = np.sqrt(np.maximum(H ** 2 - K, np.float32(0)))
        out["normal"][s:e] = a
        out["coeffs"][s:e] = w
        out["K"][s:e], out["H"][s:e] = K, H
        out["k1"][s:e], out["k2"][s:e] = H + root, H - root
        out["H2"][s:e] = H ** 2
        nr = np.linalg.norm(ref32.astype(np.float64), axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            m = np.abs(np.einsum("ni,ni->n", a, ref32.astype(np.float64)) / nr)
        out["margin"][s:e] = np.where(nr == 0, 0.0, m)
    return out


def knn_curvature(points, k, rows=None, batched=True):
    """plant_kdtree(k) + compute_pointwise_explicit_quadratic_curvature() (ref :69, :505)."""
    idx, dist, d2 = knn_canonical(points, k, rows=rows)
    if rows is None:
        rows = np.arange(len(points))
    fn = curvature_from_neighbors_batched if batched else curvature_from_neighbors
    res = fn(points, idx, rows)
    res["idx"], res["dist"], res["d2"] = idx, dist, d2
    return res


def neighbor_study(points, random_indexes, tol=1e-7, lower_bound=3, upper_bound=99, tree=None, return_counts=False):
    """explicit_quadratic_neighbor_study (ref :732-800) for a GIVEN sample of point indices.

    The reference draws the sample with ``np.random.randint(0, N, sample_size)`` (ref :751); everything
    after that is restated here line by line: ``kdtree.query(point, n + 1)`` INCLUDING the point itself
    (ref :759), centring (ref :761), plane + rotation (ref :763), fit with failures mapped to zero
    coefficients (ref :764-767), Gaussian curvature (ref :769), the binary search on
    ``abs(K(n + 1) - K(n)) < tol`` (ref :773-792) and ``int(mean) + 1`` (ref :800).
    """
    pts = np.asarray(points, dtype=np.float32)
    if tree is None:
        tree = cKDTree(pts)

    def k_gauss(i, n):
        nb = tree.query(pts[i], n + 1)[1]
        centered = pts[nb] - pts[i]
        rotated = best_fit_plane_and_rotate(centered)
        try:
            coeffs = fit_quadratic_surface(rotated)
        except Exception:
            coeffs = np.zeros(6, np.float32)
        return explicit_quadratic_curvatures(coeffs)[0]

    counts = []
    for i in np.asarray(random_indexes):
        lower, upper, best = int(lower_bound), int(upper_bound), None
        while lower <= upper:
            mid = (lower + upper) // 2
            if abs(k_gauss(i, mid + 1) - k_gauss(i, mid)) < tol:
                best = mid
                upper = mid - 1
            else:
                lower = mid + 1
        counts.append(upper if best is None else best)
    if not counts:
        return (0, np.zeros(0, np.int64)) if return_counts else 0
    result = int(np.mean(counts)) + 1
    return (result, np.asarray(counts, np.int64)) if return_counts else result
