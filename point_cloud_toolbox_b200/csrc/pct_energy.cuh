// Per-triangle terms of the energy integration that consumes the path's K and H
// (/root/reference/utils.py:702-765, load_mesh_compute_energies).  __host__ __device__ so that
// tests/host_harness can run the same arithmetic on the CPU.
#pragma once

#include "pct_math.cuh"

namespace pct {

struct TriangleTerms {
    double area;        // 0.5 * |(v1 - v0) x (v2 - v0)|, fp64                         utils.py:726-731
    double bending;     // mean of H^2 over the corners (fp32, like np.mean) * area     utils.py:754, :757
    double stretching;  // mean of K over the corners * area                            utils.py:752, :758
};

// np.mean of three float32: add.reduce in fp32 starting from 0, then a true divide by 3 in fp32
PCT_HD float mean3_f32(float a, float b, float c) {
    return fadd_rn(fadd_rn(fadd_rn(0.0f, a), b), c) / 3.0f;
}

// v*: corner coordinates (the fp32 points of the cloud; the reference holds their fp64 images), k*/h*: K and H there
PCT_HD TriangleTerms triangle_terms(const float* v0, const float* v1, const float* v2, float k0, float k1, float k2,
                                    float h0, float h1, float h2) {
    const double ax = (double)v1[0] - (double)v0[0], ay = (double)v1[1] - (double)v0[1], az = (double)v1[2] - (double)v0[2];
    const double bx = (double)v2[0] - (double)v0[0], by = (double)v2[1] - (double)v0[1], bz = (double)v2[2] - (double)v0[2];
    // np.cross: separate products and one subtraction per component
    const double cx = dadd_rn(dmul_rn(ay, bz), -dmul_rn(az, by));
    const double cy = dadd_rn(dmul_rn(az, bx), -dmul_rn(ax, bz));
    const double cz = dadd_rn(dmul_rn(ax, by), -dmul_rn(ay, bx));
    TriangleTerms t;
    t.area = 0.5 * sqrt(dadd_rn(dadd_rn(dmul_rn(cx, cx), dmul_rn(cy, cy)), dmul_rn(cz, cz)));
    const float hh = mean3_f32(fmul_rn(h0, h0), fmul_rn(h1, h1), fmul_rn(h2, h2));  // mean_curvature ** 2 is fp32
    const float kk = mean3_f32(k0, k1, k2);
    const double b = dmul_rn((double)hh, t.area), s = dmul_rn((double)kk, t.area);
    t.bending = b == b ? b : 0.0;      // np.nansum
    t.stretching = s == s ? s : 0.0;
    return t;
}

}  // namespace pct
