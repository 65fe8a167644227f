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

int launch_quadric_curvature(const float* coeffs, long long nq, float* curv, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    quadric_curvature_kernel<<<grid_for(nq), kBlock, 0, s>>>(coeffs, nq, curv);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

}  // namespace pct
