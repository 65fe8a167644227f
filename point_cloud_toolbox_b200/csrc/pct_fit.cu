// Fit from caller-provided neighbour rows, and the batched static methods.
//
//   fit_rows_kernel / fit_csr_kernel   fit_explicit_quadratic_surfaces_to_neighborhoods +
//                                      calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points
//                                      (/root/reference/pointCloudToolbox.py:635-647, :657-674)
//   plane_rotate_kernel                get_best_fit_plane_and_rotate            (:270-321)
//   quadric_fit_kernel                 fit_quadratic_surface                    (:331-360)
//   quadric_curvature_kernel           calculate_explicit_quadratic_curvatures  (:398-431)
//   pca_rows_kernel                    principal_curvatures_via_principal_component_analysis (:901-945)
//
// One thread per neighbourhood; rows are gathered from the original cloud (L2).
#include <algorithm>

#include "pct_internal.h"

namespace pct {

namespace {

constexpr int kBlock = 128;

__global__ void __launch_bounds__(kBlock)
fit_rows_kernel(const float* __restrict__ xyz, const long long n, const int32_t* __restrict__ idx, long long nq, int k,
                const int32_t* __restrict__ qids, const long long* __restrict__ offsets, FitOutputs out) {
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < nq; r += (long long)gridDim.x * kBlock) {
        long long qi = qids ? (long long)qids[r] : r;
        qi += qi < 0 ? n : 0;
        RowNeighbourhood nb;
        nb.xyz = xyz;
        nb.n = n;
        if (offsets) {
            nb.row = idx + offsets[r];
            nb.count = (int)(offsets[r + 1] - offsets[r]);
        } else {
            nb.row = idx + r * k;
            nb.count = k;
        }
        FitResult res;
        res.status = 0;
        // caller-provided indices: negative ones count from the end like numpy's (ref :640 is fancy indexing), anything
        // else outside the cloud is an error of the row, never an address
        bool valid = qi >= 0 && qi < n;
        for (int m = 0; m < nb.count && valid; ++m) {
            const long long j = nb.row[m];
            valid = j >= -n && j < n;
        }
        if (!valid) {
            fit_fail(res, ST_BAD_INDEX);
            store_fit(out, r, res);
            continue;
        }
        nb.qx = __ldg(xyz + 3 * qi); nb.qy = __ldg(xyz + 3 * qi + 1); nb.qz = __ldg(xyz + 3 * qi + 2);
        fit_neighbourhood<true>(nb, res);
        store_fit(out, r, res);
    }
}

// centred fp32 rows (nq x k x 3) -> rotated fp64 rows + unit normals
struct CentredRow {
    const float* c;
    int count;
    template <class F>
    __device__ __forceinline__ void pass(F& fn) const {
        for (int m = 0; m < count; ++m) fn.add(c[3 * m], c[3 * m + 1], c[3 * m + 2]);
    }
};

__global__ void __launch_bounds__(kBlock)
plane_rotate_kernel(const float* __restrict__ centered, long long nq, int k, double* __restrict__ rotated,
                    double* __restrict__ normals, uint8_t* __restrict__ status) {
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < nq; r += (long long)gridDim.x * kBlock) {
        CentredRow row;
        row.c = centered + r * k * 3;
        row.count = k;
        Moments mom;
        mom.reset();
        row.pass(mom);
        uint32_t st = 0;
        double* o = rotated + r * k * 3;
        if (mom.n < 2 || !mom.finite()) {
            st = mom.finite() ? ST_FEW : ST_NONFINITE;  // ref :273-274 raises on non-finite input
            const double nanv = nan("");
            for (int m = 0; m < 3 * k; ++m) o[m] = nanv;
            if (normals) { normals[3 * r] = nanv; normals[3 * r + 1] = nanv; normals[3 * r + 2] = nanv; }
        } else {
            const float* a = row.c;
            const float* b = row.c + 3 * (k - 1);
            Frame fr;
            plane_frame(mom, fsub_rn(b[0], a[0]), fsub_rn(b[1], a[1]), fsub_rn(b[2], a[2]), fr);
            bool finite = true;
            for (int m = 0; m < k; ++m) {
                double x, y, z;
                rotate_point(fr, row.c[3 * m], row.c[3 * m + 1], row.c[3 * m + 2], x, y, z);
                finite = finite && isfinite(x) && isfinite(y) && isfinite(z);
                o[3 * m] = x; o[3 * m + 1] = y; o[3 * m + 2] = z;
            }
            if (!finite) st = ST_NONFINITE;  // ref :318-319
            if (normals) { normals[3 * r] = fr.nx; normals[3 * r + 1] = fr.ny; normals[3 * r + 2] = fr.nz; }
        }
        if (status) status[r] = (uint8_t)st;
    }
}

__global__ void __launch_bounds__(kBlock)
quadric_fit_kernel(const double* __restrict__ rotated, long long nq, int k, float* __restrict__ coeffs,
                   uint8_t* __restrict__ status) {
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < nq; r += (long long)gridDim.x * kBlock) {
        const double* p = rotated + r * k * 3;
        float max_abs = 0.f;
        for (int m = 0; m < 3 * k; ++m) max_abs = fmaxf(max_abs, fabsf((float)p[m]));
        const float scale = pow2_scale(max_abs);
        Quadric q;
        q.reset();
        for (int m = 0; m < k; ++m) q.add(p[3 * m], p[3 * m + 1], p[3 * m + 2], scale);
        uint32_t st = 0;
        float c[6];
        double w[6];
        if (!q.finite() || !(max_abs <= 3.0e38f)) {
            st = ST_NONFINITE;  // ref :356-357
        } else {
            bool solved = false;
            if (k < 6) {        // underdetermined: lstsq's minimum-norm solution
                FewRows rows;
                rows.reset();
                for (int m = 0; m < k; ++m) rows.add(p[3 * m], p[3 * m + 1], p[3 * m + 2]);
                solved = solve_min_norm(rows, w);
                if (solved)
                    for (int j = 0; j < 6; ++j) c[j] = (float)w[j];
            } else if (solve_normal_equations(q, w)) {
                unscale_coefficients(w, scale, c);
                solved = true;
            }
            if (!solved) {      // dependent rows: lstsq still answers, with the minimum-norm solution (ref :359)
                GivensQR qr;
                qr.reset();
                for (int m = 0; m < k; ++m) qr.add(p[3 * m], p[3 * m + 1], p[3 * m + 2]);
                solve_min_norm_svd(qr, w);
                for (int j = 0; j < 6; ++j) c[j] = (float)w[j];
                st = ST_RANK;   // informational
            }
        }
        if (st & ST_NONFINITE)
            for (int j = 0; j < 6; ++j) c[j] = nanf("");
        for (int j = 0; j < 6; ++j) coeffs[6 * r + j] = c[j];
        if (status) status[r] = (uint8_t)st;
    }
}

__global__ void __launch_bounds__(kBlock)
quadric_curvature_kernel(const float* __restrict__ coeffs, long long nq, float* __restrict__ curv) {
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < nq; r += (long long)gridDim.x * kBlock) {
        float c[6], o[5];
        for (int j = 0; j < 6; ++j) c[j] = coeffs[6 * r + j];
        monge_curvature(c, o);
        for (int j = 0; j < 5; ++j) curv[5 * r + j] = o[j];
    }
}

// PCA of every neighbourhood row: covariance (np.cov, ddof = 1, fp64) of the k listed points (plus the query
// point itself when include_self), eigen-decomposition, and the quantities the reference derives from it.
// values (nq x 6): l1 >= l2 >= l3, K = l1 * l2, H = (l1 + l2) / 2 (ref :935-936), l3 / (l1 + l2 + l3 + 1e-10)
// (the surface variation utils.py:778-829 describes); directions (nq x 3 x 2): eigenvectors of l1 and l2.
__global__ void __launch_bounds__(kBlock)
pca_rows_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ idx, long long nq, int k, int include_self,
                const int32_t* __restrict__ qids, double* __restrict__ values, double* __restrict__ directions) {
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < nq; r += (long long)gridDim.x * kBlock) {
        const long long qi = qids ? (long long)qids[r] : r;
        const double qx = __ldg(xyz + 3 * qi), qy = __ldg(xyz + 3 * qi + 1), qz = __ldg(xyz + 3 * qi + 2);
        PcaMoments m;
        m.reset();
        if (include_self) m.add(0.0, 0.0, 0.0);
        const int32_t* row = idx + r * k;
        for (int j = 0; j < k; ++j) {
            const long long p = row[j];
            m.add((double)__ldg(xyz + 3 * p) - qx, (double)__ldg(xyz + 3 * p + 1) - qy, (double)__ldg(xyz + 3 * p + 2) - qz);
        }
        double c[6], w[3], v[3][3];
        m.covariance(c);
        eig_sym3_descending(c[0], c[1], c[2], c[3], c[4], c[5], w, v);
        double* o = values + 6 * r;
        o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        o[3] = w[0] * w[1];
        o[4] = (w[0] + w[1]) / 2.0;
        o[5] = w[2] / (w[0] + w[1] + w[2] + 1e-10);
        if (directions) {
            double* d = directions + 6 * r;
#pragma unroll
            for (int a = 0; a < 3; ++a) { d[2 * a] = v[a][0]; d[2 * a + 1] = v[a][1]; }
        }
    }
}

// ---- implicit 10-coefficient quadric (ref :363-396, :435-480) ----------------------------------------------
// The reference minimises |A c|^2 on the unit sphere with SLSQP from the all-ones start; the minimiser of that
// problem is the eigenvector of the smallest eigenvalue of A^T A (= the smallest right singular vector of A), which is
// what this kernel returns (SLSQP itself
// stops far from it, tests/test_oracle.py::test_slsqp_does_not_reach_the_minimiser -- parity with the reference's
// COEFFICIENTS is therefore unpinned; the curvature formulas below are pinned).  A = [x^2 y^2 z^2 xy xz yz x y z 1]
// with the monomials formed in fp32 like the reference's (fp32 points, ref :366-377), the rest in fp64.
// Sign: c and -c are both minimisers; the gradient at the origin (G, H, I) is made to point away from the
// neighbours' centroid.
// Unit vector c minimising |A c|: the right singular vector of the smallest singular value of A (k x 10), computed from
// A itself, not from A^T A -- the four nearly flat directions of a smooth neighbourhood have singular values 1e-8 .. 1e-6
// of the largest, whose squares (1e-16 .. 1e-12) are at the resolution limit of a symmetric eigen-solver on A^T A.
// R (10 x 10) of a QR factorisation built row by row with Givens rotations, then one-sided Jacobi rotations on the
// columns of R (the scheme of fit_min_norm, pct_math.cuh).
struct ImplicitQR {
    double r[10][10];
    __device__ void reset() {
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j) r[i][j] = 0.0;
    }
    __device__ void add(const double f[10]) {
        double v[10];
        for (int j = 0; j < 10; ++j) v[j] = f[j];
        for (int j = 0; j < 10; ++j) {
            if (v[j] == 0.0) continue;
            const double p = r[j][j], h = sqrt(p * p + v[j] * v[j]);
            const double cs = p / h, sn = v[j] / h;
            r[j][j] = h;
            for (int l = j + 1; l < 10; ++l) {
                const double t = r[j][l];
                r[j][l] = cs * t + sn * v[l];
                v[l] = cs * v[l] - sn * t;
            }
        }
    }
    __device__ void smallest_right_singular_vector(double c[10]) {
        double v[10][10];
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j) v[i][j] = i == j ? 1.0 : 0.0;
        for (int sweep = 0; sweep < 60; ++sweep) {
            bool rotated = false;
            for (int p = 0; p < 9; ++p)
                for (int q = p + 1; q < 10; ++q) {
                    double alpha = 0.0, beta = 0.0, gamma = 0.0;
                    for (int i = 0; i < 10; ++i) { alpha += r[i][p] * r[i][p]; beta += r[i][q] * r[i][q]; gamma += r[i][p] * r[i][q]; }
                    if (gamma == 0.0 || !(fabs(gamma) > 1e-15 * sqrt(alpha * beta))) continue;
                    rotated = true;
                    const double zeta = (beta - alpha) / (2.0 * gamma);
                    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    for (int i = 0; i < 10; ++i) {
                        const double x = r[i][p], y = r[i][q];
                        r[i][p] = cs * x - sn * y; r[i][q] = sn * x + cs * y;
                        const double vx = v[i][p], vy = v[i][q];
                        v[i][p] = cs * vx - sn * vy; v[i][q] = sn * vx + cs * vy;
                    }
                }
            if (!rotated) break;
        }
        int m = 0;
        double best = 1.7e308;
        for (int j = 0; j < 10; ++j) {
            double s2 = 0.0;
            for (int i = 0; i < 10; ++i) s2 += r[i][j] * r[i][j];
            if (s2 < best) { best = s2; m = j; }
        }
        double nn = 0.0;
        for (int i = 0; i < 10; ++i) nn += v[i][m] * v[i][m];
        nn = 1.0 / sqrt(nn);
        for (int i = 0; i < 10; ++i) c[i] = v[i][m] * nn;
    }
};

// rows: nq x k original indices (the reference's kdtree.query(point, k): the point itself and its k - 1 nearest)
__global__ void __launch_bounds__(64)
implicit_fit_kernel(const float* __restrict__ xyz, const long long n, const int32_t* __restrict__ idx, const long long nq, const int k,
                    const int32_t* __restrict__ qids, const float* __restrict__ centered, double* __restrict__ coeffs) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nq; r += (long long)gridDim.x * blockDim.x) {
        ImplicitQR qr;
        qr.reset();
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (!centered) {
            const long long qi = qids ? (long long)qids[r] : r;
            qx = __ldg(xyz + 3 * qi); qy = __ldg(xyz + 3 * qi + 1); qz = __ldg(xyz + 3 * qi + 2);
        }
        double sx = 0.0, sy = 0.0, sz = 0.0;
        bool ok = true;
        for (int m = 0; m < k; ++m) {
            float x, y, z;
            if (centered) {
                const float* p = centered + (r * k + m) * 3;
                x = p[0]; y = p[1]; z = p[2];
            } else {
                long long j = idx[r * k + m];
                j += j < 0 ? n : 0;
                if (j < 0 || j >= n) { ok = false; break; }
                x = fsub_rn(__ldg(xyz + 3 * j), qx); y = fsub_rn(__ldg(xyz + 3 * j + 1), qy); z = fsub_rn(__ldg(xyz + 3 * j + 2), qz);  // ref :627
            }
            sx += x; sy += y; sz += z;
            const double f[10] = {(double)fmul_rn(x, x), (double)fmul_rn(y, y), (double)fmul_rn(z, z), (double)fmul_rn(x, y),
                                  (double)fmul_rn(x, z), (double)fmul_rn(y, z), (double)x, (double)y, (double)z, 1.0};
            ok = ok && fabs(f[0]) + fabs(f[1]) + fabs(f[2]) <= 1.7e308;  // non-finite coordinates
            if (ok) qr.add(f);
        }
        double c[10];
        if (ok) {
            qr.smallest_right_singular_vector(c);
            if (c[6] * sx + c[7] * sy + c[8] * sz > 0.0)
                for (int i = 0; i < 10; ++i) c[i] = -c[i];
        } else {
            for (int i = 0; i < 10; ++i) c[i] = nan("");
        }
        for (int i = 0; i < 10; ++i) coeffs[10 * r + i] = c[i];
    }
}

// ref :435-480 at the origin of the centred neighbourhood (x = y = z = 0), operation by operation in fp64; out: K_g, K_h, k1, k2
__global__ void __launch_bounds__(kBlock)
implicit_curvature_kernel(const double* __restrict__ coeffs, long long nq, double* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < nq; r += (long long)gridDim.x * kBlock) {
        const double* c = coeffs + 10 * r;
        const double A = c[0], B = c[1], C = c[2], D = c[3], E = c[4], F = c[5], G = c[6], H = c[7], I = c[8];
        const double fx = G, fy = H, fz = I;                                  // ref :450-452 with x = y = z = 0
        const double fxx = __dmul_rn(2.0, A), fyy = __dmul_rn(2.0, B), fzz = __dmul_rn(2.0, C), fxy = D, fxz = E, fyz = F;
        const double g2 = __dadd_rn(__dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy)), __dmul_rn(fz, fz));   // g.dot(g)
        const double mag = sqrt(g2);
        const double trace = __dadd_rn(__dadd_rn(fxx, fyy), fzz);
        const double det = fxx * (fyy * fzz - fyz * fyz) - fxy * (fxy * fzz - fyz * fxz) + fxz * (fxy * fyz - fyy * fxz);  // np.linalg.det (LU there)
        const double hx = fx * fxx + fy * fxy + fz * fxz, hy = fx * fxy + fy * fyy + fz * fyz, hz = fx * fxz + fy * fyz + fz * fzz;
        const double ghg = hx * fx + hy * fy + hz * fz;
        const double Kg = det / (mag * mag * mag * mag);                       // ref :471 (the reference's formula, kept as it is)
        const double Kh = (ghg - (mag * mag) * trace) / (2.0 * (mag * mag * mag));  // ref :472
        const double root = sqrt(Kh * Kh - Kg);                                // ref :475-476 (NaN when negative, like numpy)
        out[4 * r] = Kg; out[4 * r + 1] = Kh; out[4 * r + 2] = Kh + root; out[4 * r + 3] = Kh - root;
    }
}

int grid_for(long long n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::max<long long>(1, std::min<long long>((n + kBlock - 1) / kBlock, (long long)sms * 32));
}

}  // namespace

int launch_fit_rows(const float* xyz, long long n, const int32_t* idx, long long nq, int k, const int32_t* qids,
                    FitOutputs out, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    fit_rows_kernel<<<grid_for(nq), kBlock, 0, s>>>(xyz, n, idx, nq, k, qids, nullptr, out);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_fit_csr(const float* xyz, long long n, const long long* offsets, const int32_t* idx, long long nq,
                   const int32_t* qids, FitOutputs out, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    fit_rows_kernel<<<grid_for(nq), kBlock, 0, s>>>(xyz, n, idx, nq, 0, qids, offsets, out);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_plane_rotate(const float* centered, long long nq, int k, double* rotated, double* normals, uint8_t* status,
                        cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    plane_rotate_kernel<<<grid_for(nq), kBlock, 0, s>>>(centered, nq, k, rotated, normals, status);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_quadric_fit(const double* rotated, long long nq, int k, float* coeffs, uint8_t* status, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    quadric_fit_kernel<<<grid_for(nq), kBlock, 0, s>>>(rotated, nq, k, coeffs, status);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_pca_rows(const float* xyz, const int32_t* idx, long long nq, int k, int include_self, const int32_t* qids,
                    double* values, double* directions, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    pca_rows_kernel<<<grid_for(nq), kBlock, 0, s>>>(xyz, idx, nq, k, include_self, qids, values, directions);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_implicit_fit(const float* xyz, long long n, const int32_t* idx, long long nq, int k, const int32_t* qids,
                        const float* centered, double* coeffs, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    const int grid = (int)std::max<long long>(1, std::min<long long>((nq + 63) / 64, 148ll * 16));
    implicit_fit_kernel<<<grid, 64, 0, s>>>(xyz, n, idx, nq, k, qids, centered, coeffs);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_implicit_curvature(const double* coeffs, long long nq, double* out, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    implicit_curvature_kernel<<<grid_for(nq), kBlock, 0, s>>>(coeffs, nq, out);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_quadric_curvature(const float* coeffs, long long nq, float* curv, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    quadric_curvature_kernel<<<grid_for(nq), kBlock, 0, s>>>(coeffs, nq, curv);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

}  // namespace pct
