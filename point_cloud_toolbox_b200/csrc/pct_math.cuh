// Per-neighbourhood arithmetic of the curvature path, register resident.
//
// Everything here is what ONE thread does for ONE query point once it can
// enumerate the neighbourhood: moments -> covariance -> smallest eigenvector
// (cyclic Jacobi) -> orientation -> Rodrigues frame -> fp32 quantisation ->
// 6x6 normal equations -> Cholesky -> Monge curvature.  The precision of every
// step is chosen to land on the reference's numbers, not merely near them
// (reference = /root/reference/pointCloudToolbox.py, cited as "ref :line"):
//
//   ref :641      neighbours are centred on the query point in fp32
//   ref :277      covariance in fp64 (np.cov promotes), ddof = 1
//   ref :280-283  normal = singular vector of the smallest singular value
//   ref :286-297  flip so that normal . (last - first) >= 0
//   ref :300-315  rotation I + [v]x + [v]x^2 (1-c)/s^2 in fp64 (identity if s == 0)
//   ref :350,:358 rotated points and the design matrix are quantised to fp32
//   ref :359      least squares solved in fp64, coefficients rounded to fp32
//   ref :403-431  curvature formulas in fp32
//
// The functions are __host__ __device__ so that tests/host_harness can run the
// same code on the CPU against the oracle before any GPU time is spent.  The
// shipped library only ever calls them from kernels.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PCT_HD __host__ __device__ __forceinline__
#define PCT_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define PCT_HD inline
#define PCT_HD_NOINLINE inline
#endif

namespace pct {

// ---- arithmetic with the rounding pinned (no FMA contraction) -------------
// The host build of the harness uses -ffp-contract=off, so plain operators are
// already uncontracted there.
PCT_HD float fsub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
PCT_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
PCT_HD float fmul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
PCT_HD double dadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
PCT_HD double dmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

// fp32 squared distance used for culling.  The SAME function is used by the
// selection pass and the collection pass, so both see identical bits.
PCT_HD float dist2_f32(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = fsub_rn(px, qx), dy = fsub_rn(py, qy), dz = fsub_rn(pz, qz);
    return fmaf(dz, dz, fmaf(dy, dy, fmul_rn(dx, dx)));
}

// The ranking key of scipy's cKDTree (p = 2, three coordinates):
// ((dx*dx + dy*dy) + dz*dz) on the fp64 images of the fp32 coordinates, separate
// multiplies and adds (ckdtree/src/distance.h sqeuclidean_distance_double).
PCT_HD double dist2_f64(float qx, float qy, float qz, float px, float py, float pz) {
    const double dx = (double)px - (double)qx;
    const double dy = (double)py - (double)qy;
    const double dz = (double)pz - (double)qz;
    return dadd_rn(dadd_rn(dmul_rn(dx, dx), dmul_rn(dy, dy)), dmul_rn(dz, dz));
}

// (d2, index) lexicographic order -- the canonical tie rule.
PCT_HD bool key_less(double da, uint32_t ia, double db, uint32_t ib) {
    return (da < db) || (da == db && ia < ib);
}

// ---- first pass over a neighbourhood: raw moments about the query point ----
struct Moments {
    double sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;
    float max_abs;
    int n;
    PCT_HD void reset() {
        sx = sy = sz = sxx = sxy = sxz = syy = syz = szz = 0.0;
        max_abs = 0.f;
        n = 0;
    }
    // ref :273-274: a NaN / Inf among the centred points shows in the sums of squares (which
    // cannot overflow on their own: fp32 squares, fp64 sums)
    PCT_HD bool finite() const { return fabs(sxx + syy + szz) <= 1.7e308; }
    // (cx, cy, cz) = neighbour - query, already rounded to fp32 (ref :641)
    PCT_HD void add(float cx, float cy, float cz) {
        const double x = cx, y = cy, z = cz;
        sx += x; sy += y; sz += z;
        sxx = fma(x, x, sxx); sxy = fma(x, y, sxy); sxz = fma(x, z, sxz);
        syy = fma(y, y, syy); syz = fma(y, z, syz); szz = fma(z, z, szz);
        max_abs = fmaxf(max_abs, fmaxf(fabsf(cx), fmaxf(fabsf(cy), fabsf(cz))));
        ++n;
    }
};

// ---- smallest eigenvector of a symmetric 3x3 (fp64 cyclic Jacobi) ----------
// a = [a00 a01 a02; . a11 a12; . . a22].  The matrix is a covariance, i.e. PSD,
// so the smallest eigenvalue is also the smallest singular value (ref :280-283).
PCT_HD void jacobi_rotate(double& app, double& aqq, double& apq, double& arp, double& arq,
                          double& v0p, double& v0q, double& v1p, double& v1q, double& v2p, double& v2q) {
    if (apq == 0.0) return;
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0);
    const double s = t * c;
    const double tau = s / (1.0 + c);
    app -= t * apq;
    aqq += t * apq;
    apq = 0.0;
    // the remaining row/column r (the third index)
    const double rp = arp, rq = arq;
    arp = rp - s * (rq + tau * rp);
    arq = rq + s * (rp - tau * rq);
    double a, b;
    a = v0p; b = v0q; v0p = a - s * (b + tau * a); v0q = b + s * (a - tau * b);
    a = v1p; b = v1q; v1p = a - s * (b + tau * a); v1q = b + s * (a - tau * b);
    a = v2p; b = v2q; v2p = a - s * (b + tau * a); v2q = b + s * (a - tau * b);
}

PCT_HD_NOINLINE void smallest_eigenvector_sym3_jacobi(double a00, double a01, double a02, double a11, double a12, double a22,
                                      double n[3]) {
    // scale by the trace: eigenvectors are unchanged, fp range issues vanish
    const double tr = a00 + a11 + a22;
    if (!(tr > 0.0)) {  // all neighbours coincide with their mean (or NaN): any unit vector
        n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
        return;
    }
    const double inv = 1.0 / tr;
    a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
#pragma unroll 1
    for (int sweep = 0; sweep < 10; ++sweep) {
        const double off = a01 * a01 + a02 * a02 + a12 * a12;
        if (off < 1e-36) break;  // relative to trace^2 == 1
        jacobi_rotate(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);  // (p,q) = (0,1), r = 2
        jacobi_rotate(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);  // (0,2), r = 1
        jacobi_rotate(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);  // (1,2), r = 0
    }
    if (a00 <= a11 && a00 <= a22) { n[0] = v00; n[1] = v10; n[2] = v20; }
    else if (a11 <= a22)          { n[0] = v01; n[1] = v11; n[2] = v21; }
    else                          { n[0] = v02; n[1] = v12; n[2] = v22; }
}

// All three eigenpairs of a symmetric 3x3 (fp64 cyclic Jacobi), eigenvalues in DESCENDING order, column j of
// v = eigenvector of w[j].  For the PCA estimators (/root/reference/pointCloudToolbox.py:901-945).
PCT_HD_NOINLINE void eig_sym3_descending(double a00, double a01, double a02, double a11, double a12, double a22,
                                         double w[3], double v[3][3]) {
    const double tr = fabs(a00) + fabs(a11) + fabs(a22);
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
    if (tr > 0.0 && tr <= 1.7e308) {
        const double inv = 1.0 / tr;
        a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
#pragma unroll 1
        for (int sweep = 0; sweep < 12; ++sweep) {
            const double off = a01 * a01 + a02 * a02 + a12 * a12;
            if (off < 1e-40) break;
            jacobi_rotate(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);
            jacobi_rotate(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);
            jacobi_rotate(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);
        }
        a00 *= tr; a11 *= tr; a22 *= tr;
    }
    double ev[3] = {a00, a11, a22};
    double vec[3][3] = {{v00, v01, v02}, {v10, v11, v12}, {v20, v21, v22}};
    int o0 = 0, o1 = 1, o2 = 2, t;
    if (ev[o0] < ev[o1]) { t = o0; o0 = o1; o1 = t; }
    if (ev[o1] < ev[o2]) { t = o1; o1 = o2; o2 = t; }
    if (ev[o0] < ev[o1]) { t = o0; o0 = o1; o1 = t; }
    const int ord[3] = {o0, o1, o2};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        w[j] = ev[ord[j]];
#pragma unroll
        for (int r = 0; r < 3; ++r) v[r][j] = vec[r][ord[j]];
    }
}

// Covariance (ddof = 1, np.cov) of a neighbourhood from its raw fp64 moments about ANY origin: np.cov
// subtracts the mean, so the origin drops out; the query point is used because the fp64 differences of
// fp32 coordinates are exact and small.
struct PcaMoments {
    double sx, sy, sz, sxx, sxy, sxz, syy, syz, szz;
    int n;
    PCT_HD void reset() { sx = sy = sz = sxx = sxy = sxz = syy = syz = szz = 0.0; n = 0; }
    PCT_HD void add(double x, double y, double z) {
        sx += x; sy += y; sz += z;
        sxx = fma(x, x, sxx); sxy = fma(x, y, sxy); sxz = fma(x, z, sxz);
        syy = fma(y, y, syy); syz = fma(y, z, syz); szz = fma(z, z, szz);
        ++n;
    }
    PCT_HD void covariance(double c[6]) const {  // c00 c01 c02 c11 c12 c22; n = 1 gives nan like np.cov
        const double inv_n = 1.0 / (double)n, inv = 1.0 / (double)(n - 1);
        c[0] = (sxx - sx * sx * inv_n) * inv; c[1] = (sxy - sx * sy * inv_n) * inv; c[2] = (sxz - sx * sz * inv_n) * inv;
        c[3] = (syy - sy * sy * inv_n) * inv; c[4] = (syz - sy * sz * inv_n) * inv; c[5] = (szz - sz * sz * inv_n) * inv;
    }
};

// Smallest eigenvector in fp64 accuracy for a fraction of the fp64 Jacobi's cost (its rotations are
// serial chains of fp64 divisions and square roots): the closed form of the eigenvalues of a symmetric
// 3x3 (trigonometric solution of the characteristic cubic) in fp32 gives the smallest eigenvalue to
// ~1e-7 of the trace, the cross product of two rows of (A - lambda I) its eigenvector to ~1e-7 / gap,
// and two Rayleigh-quotient iterations in fp64 -- each a multiplication by the adjugate of
// (A - lambda I), no division -- converge cubically from there.  When the two smallest eigenvalues are
// closer than 1e-4 of the trace the start may be poor and the fp64 Jacobi is used instead.
PCT_HD void smallest_eigenvector_sym3(double a00, double a01, double a02, double a11, double a12, double a22,
                                      double n[3]) {
    const double tr = a00 + a11 + a22;
    if (!(tr > 0.0)) {  // all neighbours coincide with their mean (or NaN): any unit vector
        n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
        return;
    }
    const double inv = 1.0 / tr;
    a00 *= inv; a01 *= inv; a02 *= inv; a11 *= inv; a12 *= inv; a22 *= inv;
    // eigenvalues of A (trace 1): q + 2 p cos(phi + 2 pi j / 3), q = 1/3
    const float f01 = (float)a01, f02 = (float)a02, f12 = (float)a12;
    const float q = 1.f / 3.f;
    const float b00 = (float)a00 - q, b11 = (float)a11 - q, b22 = (float)a22 - q;
    const float p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.f * (f01 * f01 + f02 * f02 + f12 * f12);
    const float p = sqrtf(p2 * (1.f / 6.f));
    bool good = p > 1e-6f;
    double x = 0.0, y = 0.0, z = 1.0;
    if (good) {
        const float ip = 1.f / p;
        const float c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = f01 * ip, c02 = f02 * ip, c12 = f12 * ip;
        float r = 0.5f * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
        r = fminf(1.f, fmaxf(-1.f, r));
        const float phi = acosf(r) * (1.f / 3.f);
        const float e1 = q + 2.f * p * cosf(phi);
        const float e3 = q + 2.f * p * cosf(phi + 2.0943951f);  // smallest
        const float e2 = 1.f - e1 - e3;
        good = e2 - e3 > 1e-4f;
        // null vector of A - e3 I: the largest of the three cross products of its rows
        const float m00 = (float)a00 - e3, m11 = (float)a11 - e3, m22 = (float)a22 - e3;
        const float u0 = f01 * f12 - f02 * m11, u1 = f02 * f01 - m00 * f12, u2 = m00 * m11 - f01 * f01;   // row0 x row1
        const float v0 = f01 * m22 - f02 * f12, v1 = f02 * f02 - m00 * m22, v2 = m00 * f12 - f01 * f02;   // row0 x row2
        const float w0 = m11 * m22 - f12 * f12, w1 = f12 * f02 - f01 * m22, w2 = f01 * f12 - m11 * f02;   // row1 x row2
        const float nu = u0 * u0 + u1 * u1 + u2 * u2, nv = v0 * v0 + v1 * v1 + v2 * v2, nw = w0 * w0 + w1 * w1 + w2 * w2;
        if (nu >= nv && nu >= nw) { x = u0; y = u1; z = u2; }
        else if (nv >= nw)        { x = v0; y = v1; z = v2; }
        else                      { x = w0; y = w1; z = w2; }
        good = good && fmaxf(nu, fmaxf(nv, nw)) > 1e-20f;
    }
    if (!good) {
        smallest_eigenvector_sym3_jacobi(a00, a01, a02, a11, a12, a22, n);
        return;
    }
#pragma unroll
    for (int step = 0; step < 2; ++step) {
        const double ax = a00 * x + a01 * y + a02 * z, ay = a01 * x + a11 * y + a12 * z, az = a02 * x + a12 * y + a22 * z;
        const double lam = (x * ax + y * ay + z * az) / (x * x + y * y + z * z);
        const double m00 = a00 - lam, m11 = a11 - lam, m22 = a22 - lam;
        const double c00 = m11 * m22 - a12 * a12, c01 = a02 * a12 - a01 * m22, c02 = a01 * a12 - a02 * m11;
        const double c11 = m00 * m22 - a02 * a02, c12 = a01 * a02 - m00 * a12, c22 = m00 * m11 - a01 * a01;
        const double nx = c00 * x + c01 * y + c02 * z, ny = c01 * x + c11 * y + c12 * z, nz = c02 * x + c12 * y + c22 * z;
        const double nn = nx * nx + ny * ny + nz * nz;
        if (!(nn > 1e-280)) break;
        const double s = 1.0 / sqrt(nn);
        x = nx * s; y = ny * s; z = nz * s;
    }
    n[0] = x; n[1] = y; n[2] = z;
}

// ---- tangent frame ---------------------------------------------------------
struct Frame {
    double r00, r01, r02, r10, r11, r12, r20, r21, r22;  // Rodrigues matrix, row major
    double nx, ny, nz;                                   // oriented unit normal
};

// moments -> covariance -> normal, oriented by ref = c_last - c_first (fp32, ref :286)
PCT_HD void plane_frame(const Moments& m, float rfx, float rfy, float rfz, Frame& f) {
    const double invn = 1.0 / (double)m.n;
    const double denom = 1.0 / (double)(m.n - 1);  // ddof = 1 (ref :277); scale does not move eigenvectors
    const double mx = m.sx * invn, my = m.sy * invn, mz = m.sz * invn;
    const double c00 = (m.sxx - m.sx * mx) * denom, c01 = (m.sxy - m.sx * my) * denom, c02 = (m.sxz - m.sx * mz) * denom;
    const double c11 = (m.syy - m.sy * my) * denom, c12 = (m.syz - m.sy * mz) * denom, c22 = (m.szz - m.sz * mz) * denom;
    double n[3];
    smallest_eigenvector_sym3(c00, c01, c02, c11, c12, c22, n);
    // ref :289-297: only the SIGN of normal.ref matters; a zero-length ref gives NaN -> no flip
    const double dot = n[0] * (double)rfx + n[1] * (double)rfy + n[2] * (double)rfz;
    if (dot < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    const double nn = 1.0 / sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);  // ref :301
    const double ax = n[0] * nn, ay = n[1] * nn, az = n[2] * nn;
    f.nx = ax; f.ny = ay; f.nz = az;
    // v = a x z = (ay, -ax, 0); c = az; s = |v|  (ref :303-305)
    const double s = sqrt(ay * ay + ax * ax);
    if (s == 0.0) {  // ref :308-309, also taken for a == -z
        f.r00 = 1; f.r01 = 0; f.r02 = 0; f.r10 = 0; f.r11 = 1; f.r12 = 0; f.r20 = 0; f.r21 = 0; f.r22 = 1;
        return;
    }
    const double fac = (1.0 - az) / (s * s);  // ref :312
    f.r00 = 1.0 - fac * ax * ax; f.r01 = -fac * ax * ay;      f.r02 = -ax;
    f.r10 = -fac * ax * ay;      f.r11 = 1.0 - fac * ay * ay; f.r12 = -ay;
    f.r20 = ax;                  f.r21 = ay;                  f.r22 = 1.0 - fac * (ax * ax + ay * ay);
}

PCT_HD void rotate_point(const Frame& f, float cx, float cy, float cz, double& x, double& y, double& z) {
    const double a = cx, b = cy, c = cz;
    x = f.r00 * a + f.r01 * b + f.r02 * c;
    y = f.r10 * a + f.r11 * b + f.r12 * c;
    z = f.r20 * a + f.r21 * b + f.r22 * c;
}

// ---- second pass: normal equations of z = A a^2 + B b^2 + C ab + D a + E b + F
// Coordinates are pre-multiplied by a power of two `scale` (about 1/r_k).  This
// is exact in binary floating point, so the fp32 products below are exactly the
// reference's fp32 design-matrix entries times a power of two, while the 6x6
// system becomes well conditioned (unscaled it reaches cond ~ 1e11 on bunny).
struct Quadric {
    double g[21];  // upper triangle of X^T X, row major: (0,0) (0,1) .. (0,5) (1,1) ..
    double r[6];   // X^T z
    PCT_HD void reset() {
#pragma unroll
        for (int i = 0; i < 21; ++i) g[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) r[i] = 0.0;
    }
    // ref :318-319, :356-357: non-finite rotated coordinates show in sum a^2, sum b^2, sum z
    PCT_HD bool finite() const { return fabs(g[15]) <= 1.7e308 && fabs(g[18]) <= 1.7e308 && fabs(r[5]) <= 1.7e308; }
    // rotated coordinates in fp64, quantised here to fp32 like ref :350
    PCT_HD void add(double xr, double yr, double zr, float scale) {
        const float a = (float)xr * scale, b = (float)yr * scale, z = (float)zr * scale;
        double x[6];
        x[0] = (double)fmul_rn(a, a);  // ref :358, fp32 products
        x[1] = (double)fmul_rn(b, b);
        x[2] = (double)fmul_rn(a, b);
        x[3] = (double)a;
        x[4] = (double)b;
        x[5] = 1.0;
        const double zd = (double)z;
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) { g[t] = fma(x[i], x[j], g[t]); ++t; }
            r[i] = fma(x[i], zd, r[i]);
        }
    }
};

// power of two close to 1 / max_abs (exact scaling)
PCT_HD float pow2_scale(float max_abs) {
    if (!(max_abs > 0.f) || !(max_abs < 3.0e38f)) return 1.f;
    int e = ilogbf(max_abs);
    if (e > 100) e = 100;
    if (e < -100) e = -100;
    return ldexpf(1.f, -e);
}

// In-place Cholesky of the 6x6 normal matrix + two triangular solves.
// Returns false when a pivot is not safely positive (rank-deficient design).
PCT_HD bool solve_normal_equations(Quadric& q, double w[6]) {
    // unpack to a full lower triangle with compile-time indices (stays in registers)
    double L[6][6];
    {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) { L[j][i] = q.g[t]; ++t; }
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
        const double d0 = d;
#pragma unroll
        for (int p = 0; p < j; ++p) d -= L[j][p] * L[j][p];
        if (!(d > 1e-13 * d0) || !(d0 > 0.0)) { ok = false; d = 1.0; }
        const double inv = 1.0 / sqrt(d);
        L[j][j] = d * inv;  // sqrt(d)
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = L[i][j];
#pragma unroll
            for (int p = 0; p < j; ++p) s -= L[i][p] * L[j][p];
            L[i][j] = s * inv;
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = q.r[i];
#pragma unroll
        for (int p = 0; p < i; ++p) s -= L[i][p] * y[p];
        y[i] = s / L[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
#pragma unroll
        for (int p = i + 1; p < 6; ++p) s -= L[p][i] * w[p];
        w[i] = s / L[i][i];
    }
    return ok;
}

// Fewer than 6 rows: `lstsq` (ref :359) returns the MINIMUM-NORM solution of the underdetermined
// system, which the neighbour study relies on for its smallest probes (ref :756-770 with n = 3, 4).
// The norm that is minimised is that of the unscaled coefficients, so this path works on the
// unscaled fp32 features.  Rows of X are orthonormalised (modified Gram-Schmidt, twice): X = R^T Q^T,
// w = Q R^-T z -- conditioning of X, not of X X^T.  Rank-deficient rows -> false.
struct FewRows {
    double x[5][6];
    double z[5];
    int m;
    PCT_HD void reset() { m = 0; }
    PCT_HD void add(double xr, double yr, double zr) {
        if (m >= 5) { ++m; return; }
        const float a = (float)xr, b = (float)yr, c = (float)zr;  // ref :350
        x[m][0] = (double)fmul_rn(a, a);
        x[m][1] = (double)fmul_rn(b, b);
        x[m][2] = (double)fmul_rn(a, b);
        x[m][3] = (double)a;
        x[m][4] = (double)b;
        x[m][5] = 1.0;
        z[m] = (double)c;
        ++m;
    }
    PCT_HD bool finite() const {
        for (int i = 0; i < m && i < 5; ++i)
            if (!(fabs(x[i][3]) <= 1.7e308) || !(fabs(x[i][4]) <= 1.7e308) || !(fabs(z[i]) <= 1.7e308)) return false;
        return true;
    }
};

PCT_HD_NOINLINE bool solve_min_norm(FewRows& f, double w[6]) {
    const int m = f.m;
    double q[5][6], r[5][5], y[5];
    for (int i = 0; i < m; ++i) {
        double v[6], norm0 = 0.0;
        for (int c = 0; c < 6; ++c) { v[c] = f.x[i][c]; norm0 += v[c] * v[c]; }
        for (int j = 0; j < 5; ++j) r[j][i] = 0.0;
        for (int rep = 0; rep < 2; ++rep)
            for (int j = 0; j < i; ++j) {
                double dot = 0.0;
                for (int c = 0; c < 6; ++c) dot += q[j][c] * v[c];
                for (int c = 0; c < 6; ++c) v[c] -= dot * q[j][c];
                r[j][i] += dot;
            }
        double norm = 0.0;
        for (int c = 0; c < 6; ++c) norm += v[c] * v[c];
        if (!(norm > 1e-26 * norm0) || !(norm0 > 0.0)) return false;  // this row depends on the previous ones
        norm = sqrt(norm);
        r[i][i] = norm;
        for (int c = 0; c < 6; ++c) q[i][c] = v[c] / norm;
    }
    // X = R^T Q^T  =>  R^T y = z (forward substitution), w = Q y
    for (int i = 0; i < m; ++i) {
        double s = f.z[i];
        for (int j = 0; j < i; ++j) s -= r[j][i] * y[j];
        y[i] = s / r[i][i];
    }
    for (int c = 0; c < 6; ++c) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += q[i][c] * y[i];
        w[c] = s;
    }
    return true;
}

// Rank-deficient designs with any number of rows: `lstsq` (ref :359, LAPACK dgelsd, rcond = eps * max(M, N)) returns
// the MINIMUM-NORM least-squares solution -- finite numbers for collinear scan lines, duplicated points and
// neighbourhoods that lie on a conic.  Restated without storing the rows: a QR factorisation of the UNSCALED design
// (the norm that is minimised is that of the unscaled coefficients) built row by row with Givens rotations, R 6 x 6
// and c = Q^T z; then the SVD of R by one-sided Jacobi rotations (small singular values come out with high relative
// accuracy), singular values up to rcond * s_max dropped like dgelsd does, w = V S^+ U^T c.
struct GivensQR {
    double r[6][6];  // upper triangle
    double c[6];
    int m;
    PCT_HD void reset() {
        for (int i = 0; i < 6; ++i) {
            c[i] = 0.0;
            for (int j = 0; j < 6; ++j) r[i][j] = 0.0;
        }
        m = 0;
    }
    // rotated coordinates in fp64, quantised to fp32 like ref :350; features in fp32 like ref :358
    PCT_HD void add(double xr, double yr, double zr) {
        const float a = (float)xr, b = (float)yr, z = (float)zr;
        double v[6] = {(double)fmul_rn(a, a), (double)fmul_rn(b, b), (double)fmul_rn(a, b), (double)a, (double)b, 1.0};
        double zz = (double)z;
        for (int j = 0; j < 6; ++j) {
            if (v[j] == 0.0) continue;
            const double p = r[j][j], h = sqrt(p * p + v[j] * v[j]);
            const double cs = p / h, sn = v[j] / h;
            r[j][j] = h;
            for (int l = j + 1; l < 6; ++l) {
                const double t = r[j][l];
                r[j][l] = cs * t + sn * v[l];
                v[l] = cs * v[l] - sn * t;
            }
            const double t = c[j];
            c[j] = cs * t + sn * zz;
            zz = cs * zz - sn * t;
        }
        ++m;
    }
    PCT_HD bool finite() const {
        double s = 0.0;
        for (int i = 0; i < 6; ++i) s += fabs(r[i][i]) + fabs(c[i]);
        return s <= 1.7e308;
    }
};

PCT_HD_NOINLINE void solve_min_norm_svd(const GivensQR& f, double w[6]) {
    double a[6][6], v[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { a[i][j] = j >= i ? f.r[i][j] : 0.0; v[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int i = 0; i < 6; ++i) { alpha += a[i][p] * a[i][p]; beta += a[i][q] * a[i][q]; gamma += a[i][p] * a[i][q]; }
                if (gamma == 0.0 || !(fabs(gamma) > 1e-15 * sqrt(alpha * beta))) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < 6; ++i) {
                    const double x = a[i][p], y = a[i][q];
                    a[i][p] = cs * x - sn * y; a[i][q] = sn * x + cs * y;
                    const double vx = v[i][p], vy = v[i][q];
                    v[i][p] = cs * vx - sn * vy; v[i][q] = sn * vx + cs * vy;
                }
            }
        if (!rotated) break;
    }
    double s2[6], s2_max = 0.0;
    for (int j = 0; j < 6; ++j) {
        s2[j] = 0.0;
        for (int i = 0; i < 6; ++i) s2[j] += a[i][j] * a[i][j];
        s2_max = s2[j] > s2_max ? s2[j] : s2_max;
    }
    const double rcond = 2.220446049250313e-16 * (double)(f.m > 6 ? f.m : 6);
    for (int i = 0; i < 6; ++i) w[i] = 0.0;
    for (int j = 0; j < 6; ++j) {
        if (!(s2[j] > rcond * rcond * s2_max) || !(s2[j] > 0.0)) continue;  // s_j <= rcond * s_max: dropped
        double uc = 0.0;
        for (int i = 0; i < 6; ++i) uc += a[i][j] * f.c[i];
        const double g = uc / s2[j];
        for (int i = 0; i < 6; ++i) w[i] += v[i][j] * g;
    }
}

// scaled solution -> reference coefficients [A,B,C,D,E,F] in fp32 (ref :359 result dtype)
PCT_HD void unscale_coefficients(const double w[6], float scale, float c[6]) {
    const double s = (double)scale;
    c[0] = (float)(w[0] * s);
    c[1] = (float)(w[1] * s);
    c[2] = (float)(w[2] * s);
    c[3] = (float)w[3];
    c[4] = (float)w[4];
    c[5] = (float)(w[5] / s);
}

// ref :403-431 in fp32, operation by operation
PCT_HD void monge_curvature(const float c[6], float out[5]) {
    const float A = c[0], B = c[1], C = c[2], D = c[3], E = c[4];
    const float fx2 = fmul_rn(D, D), fy2 = fmul_rn(E, E);
    const float fxx = fmul_rn(2.f, A), fyy = fmul_rn(2.f, B), fxy = C;
    const float g = fadd_rn(fadd_rn(1.f, fx2), fy2);
    const float den_g = fmul_rn(g, g);
    const float den_m = fmul_rn(g, sqrtf(g));  // g ** 1.5
    const float K = fsub_rn(fmul_rn(fxx, fyy), fmul_rn(fxy, fxy)) / den_g;
    const float t1 = fmul_rn(fadd_rn(1.f, fx2), fyy);
    const float t2 = fmul_rn(fmul_rn(fmul_rn(2.f, D), E), fxy);
    const float t3 = fmul_rn(fadd_rn(1.f, fy2), fxx);
    const float H = fadd_rn(fsub_rn(t1, t2), t3) / fmul_rn(2.f, den_m);
    const float H2 = fmul_rn(H, H);
    const float disc = fmaxf(fsub_rn(H2, K), 0.f);
    const float root = sqrtf(disc);
    out[0] = K;
    out[1] = H;
    out[2] = fadd_rn(H, root);
    out[3] = fsub_rn(H, root);
    out[4] = H2;
}

struct FitResult {
    float normal[3];
    float coeffs[6];
    float curv[5];
    uint32_t status;
};

PCT_HD void fit_fail(FitResult& o, uint32_t status) {
    const float nanv = nanf("");
    o.normal[0] = o.normal[1] = o.normal[2] = nanv;
#pragma unroll
    for (int i = 0; i < 6; ++i) o.coeffs[i] = nanv;
#pragma unroll
    for (int i = 0; i < 5; ++i) o.curv[i] = nanv;
    o.status |= status;
}

// status bits (mirrors include/pct_b200.h)
enum : uint32_t { ST_EXACT_PATH = 1u, ST_FEW = 2u, ST_RANK = 4u, ST_NONFINITE = 8u, ST_UNRESOLVED = 16u, ST_BAD_INDEX = 32u };

// minimum-norm solution of a design that has no unique least-squares solution: fewer rows than coefficients, or
// rows that are linearly dependent (kept out of line, it is rare and needs a stack frame)
template <class Nbr>
PCT_HD_NOINLINE void fit_min_norm(Nbr& nb, const Frame& fr, FitResult& out) {
    out.normal[0] = (float)fr.nx; out.normal[1] = (float)fr.ny; out.normal[2] = (float)fr.nz;
    double w[6];
    bool solved = false;
    {
        struct Collect {
            const Frame* f;
            FewRows rows;
            PCT_HD void add(float cx, float cy, float cz) {
                double x, y, z;
                rotate_point(*f, cx, cy, cz, x, y, z);
                rows.add(x, y, z);
            }
        } col;
        col.f = &fr;
        col.rows.reset();
        nb.pass(col);
        if (col.rows.m <= 5) {
            if (!col.rows.finite()) { fit_fail(out, ST_NONFINITE); return; }
            solved = solve_min_norm(col.rows, w);  // independent rows: X = R^T Q^T, w = Q R^-T z
        }
    }
    if (!solved) {
        struct Stream {
            const Frame* f;
            GivensQR qr;
            PCT_HD void add(float cx, float cy, float cz) {
                double x, y, z;
                rotate_point(*f, cx, cy, cz, x, y, z);
                qr.add(x, y, z);
            }
        } st;
        st.f = &fr;
        st.qr.reset();
        nb.pass(st);
        if (!st.qr.finite()) {
            const float nx = out.normal[0], ny = out.normal[1], nz = out.normal[2];
            fit_fail(out, ST_NONFINITE);
            out.normal[0] = nx; out.normal[1] = ny; out.normal[2] = nz;
            return;
        }
        solve_min_norm_svd(st.qr, w);
        out.status |= ST_RANK;  // informational: the design was rank deficient, the coefficients are lstsq's minimum-norm ones
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) out.coeffs[c] = (float)w[c];
    monge_curvature(out.coeffs, out.curv);
}

// Whole per-point pipeline over an abstract neighbourhood.
//   nb.pass(fn)  calls fn(cx, cy, cz) for every neighbour (fp32, centred)
//   nb.reference(rx, ry, rz) gives c_last - c_first in fp32, valid after the first pass
// FEW_ROWS: neighbourhoods of fewer than 6 rows and rank-deficient ones get lstsq's minimum-norm solution
// (fit_min_norm: the fit-from-rows entry points and the kernel that redoes such queries of the search kernels).
// The search kernels instantiate it with false -- the buffers of that path would cost them a 1.2 KB stack frame
// and 8 % of their speed -- and report such neighbourhoods as ST_RANK with NaN outputs, which queues them.
template <bool FEW_ROWS, class Nbr>
PCT_HD void fit_neighbourhood(Nbr& nb, FitResult& out) {
    Moments mom;
    mom.reset();
    nb.pass(mom);
    if (mom.n < 2) { fit_fail(out, ST_FEW); return; }
    if (!mom.finite()) { fit_fail(out, ST_NONFINITE); return; }
    float rx, ry, rz;
    nb.reference(rx, ry, rz);
    Frame fr;
    plane_frame(mom, rx, ry, rz, fr);
    if (FEW_ROWS && mom.n < 6) {
        fit_min_norm(nb, fr, out);
        return;
    }
    struct Second {
        const Frame* f;
        Quadric q;
        float scale;
        PCT_HD void add(float cx, float cy, float cz) {
            double x, y, z;
            rotate_point(*f, cx, cy, cz, x, y, z);
            q.add(x, y, z, scale);
        }
    } sec;
    sec.f = &fr;
    sec.q.reset();
    sec.scale = pow2_scale(mom.max_abs);
    nb.pass(sec);
    out.normal[0] = (float)fr.nx; out.normal[1] = (float)fr.ny; out.normal[2] = (float)fr.nz;
    if (!sec.q.finite()) { fit_fail(out, ST_NONFINITE); return; }
    double w[6];
    const bool ok = solve_normal_equations(sec.q, w);
    if (!ok && FEW_ROWS) {  // rank-deficient rows: lstsq's minimum-norm solution (ref :359)
        fit_min_norm(nb, fr, out);
        return;
    }
    if (!ok) {  // (search kernels: the query is redone by the kernel that carries the minimum-norm solver)
        const float nx = out.normal[0], ny = out.normal[1], nz = out.normal[2];
        fit_fail(out, ST_RANK);
        out.normal[0] = nx; out.normal[1] = ny; out.normal[2] = nz;
        return;
    }
    unscale_coefficients(w, sec.scale, out.coeffs);
    monge_curvature(out.coeffs, out.curv);
}

}  // namespace pct
