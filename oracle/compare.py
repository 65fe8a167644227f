"""The parity policy, written once and used by every test.

TEST INFRASTRUCTURE ONLY.

Let r_k be the distance to the last neighbour of a row (the natural length
scale: curvature scales like 1/r_k).  Stated fp32 tolerance, as BASELINE.json's
north_star asks ("relative 1e-3, absolute floor near zero curvature"):

    neighbour indices   exact, in (d2, index) order
    neighbour dists     <= 1 ulp(fp32)
    normal              1 - |n.n_ref| <= 1e-5; sign equal when the orientation
                        margin |n.r| >= MARGIN, exempt below it
    K                   |dK| <= 1e-3 |K_ref| + 1e-5 / r_k^2, signed
    H, k1, k2           |dH| <= 1e-3 |H_ref| + 1e-5 / r_k, signed when the margin
                        >= MARGIN, otherwise |H| and the unordered {|k1|,|k2|}

The GPU path is expected to sit far inside these (it reproduces the reference's
fp64 steps); ``tight_fraction`` reports how many rows agree to 1e-5 relative so
a regression shows up long before the gate does.
"""
from __future__ import annotations

import numpy as np

REL = 1e-3
ABS_FLOOR = 1e-5
NORMAL_TOL = 1e-5
MARGIN = 1e-5


def neighbor_rows_differing(idx_a, idx_b):
    """Number of rows whose ordered index lists differ (target: 0)."""
    a = np.asarray(idx_a)
    b = np.asarray(idx_b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int(np.count_nonzero((a != b).any(axis=1)))


def ulp_distance_f32(a, b):
    """|a - b| in units of fp32 ulps (both finite, same sign expected)."""
    ia = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def csr_equal(off_a, idx_a, off_b, idx_b):
    return bool(np.array_equal(off_a, off_b) and np.array_equal(idx_a, idx_b))


def curvature_report(got, ref, r_k, rows_ok=None):
    """Compare dicts with keys normal, K, H, k1, k2 (+ ref['margin']).

    Returns a dict of violation counts and error statistics.  ``rows_ok`` masks
    rows to be checked (default: every row whose reference is finite).
    """
    r_k = np.asarray(r_k, np.float64)
    Kr, Hr = ref["K"].astype(np.float64), ref["H"].astype(np.float64)
    k1r, k2r = ref["k1"].astype(np.float64), ref["k2"].astype(np.float64)
    Kg, Hg = np.asarray(got["K"], np.float64), np.asarray(got["H"], np.float64)
    k1g, k2g = np.asarray(got["k1"], np.float64), np.asarray(got["k2"], np.float64)
    finite = np.isfinite(Kr) & np.isfinite(Hr)
    if rows_ok is not None:
        finite &= rows_ok
    margin = ref["margin"]
    safe = margin >= MARGIN

    tolK = REL * np.abs(Kr) + ABS_FLOOR / r_k ** 2
    tolH = REL * np.abs(Hr) + ABS_FLOOR / r_k
    errK = np.abs(Kg - Kr)
    errH_signed = np.abs(Hg - Hr)
    errH_abs = np.abs(np.abs(Hg) - np.abs(Hr))
    errH = np.where(safe, errH_signed, errH_abs)

    tolk = REL * np.maximum(np.abs(k1r), np.abs(k2r)) + ABS_FLOOR / r_k
    pair_signed = np.maximum(np.abs(k1g - k1r), np.abs(k2g - k2r))
    sg = np.sort(np.abs(np.stack((k1g, k2g), 1)), axis=1)
    sr = np.sort(np.abs(np.stack((k1r, k2r), 1)), axis=1)
    pair_abs = np.abs(sg - sr).max(axis=1)
    errk = np.where(safe, pair_signed, pair_abs)

    ng = np.asarray(got["normal"], np.float64)
    nr = np.asarray(ref["normal"], np.float64)
    dots = np.einsum("ni,ni->n", ng, nr)
    err_n = 1.0 - np.abs(dots)
    sign_bad = safe & (dots < 0)

    nan_got = ~(np.isfinite(Kg) & np.isfinite(Hg))
    with np.errstate(divide="ignore", invalid="ignore"):
        relK = errK / np.maximum(np.abs(Kr), ABS_FLOOR / r_k ** 2)
        relH = errH / np.maximum(np.abs(Hr), ABS_FLOOR / r_k)
    chk = finite
    rep = {
        "rows": int(chk.sum()),
        "nan_rows": int((nan_got & chk).sum()),
        "K_viol": int(((errK > tolK) & chk).sum()),
        "H_viol": int(((errH > tolH) & chk).sum()),
        "k12_viol": int(((errk > tolk) & chk).sum()),
        "normal_viol": int(((err_n > NORMAL_TOL) & chk).sum()),
        "sign_viol": int((sign_bad & chk).sum()),
        "K_rel_max": float(np.nanmax(np.where(chk, relK, 0))) if chk.any() else 0.0,
        "H_rel_max": float(np.nanmax(np.where(chk, relH, 0))) if chk.any() else 0.0,
        "K_rel_p999": float(np.nanquantile(relK[chk], 0.999)) if chk.any() else 0.0,
        "H_rel_p999": float(np.nanquantile(relH[chk], 0.999)) if chk.any() else 0.0,
        "tight_fraction": float(np.mean((relK[chk] < 1e-5) & (relH[chk] < 1e-5))) if chk.any() else 1.0,
        "unsafe_rows": int((~safe & chk).sum()),
    }
    rep["violations"] = rep["nan_rows"] + rep["K_viol"] + rep["H_viol"] + rep["k12_viol"] + rep["normal_viol"] + rep["sign_viol"]
    return rep


def closed_form_report(K_est, H_est, K_true, H_true, mask=None):
    """K signed, |H| unsigned against analytic curvature (SURVEY.md 8(c) last row)."""
    K_est = np.asarray(K_est, np.float64)
    H_est = np.asarray(H_est, np.float64)
    if mask is None:
        mask = np.ones(len(K_est), bool)
    mask = mask & np.isfinite(K_est) & np.isfinite(H_est)
    dK = np.abs(K_est - K_true)[mask]
    dH = np.abs(np.abs(H_est) - np.abs(H_true))[mask]
    return {
        "rows": int(mask.sum()),
        "K_abs_median": float(np.median(dK)),
        "K_abs_p99": float(np.quantile(dK, 0.99)),
        "H_abs_median": float(np.median(dH)),
        "H_abs_p99": float(np.quantile(dH, 0.99)),
    }
