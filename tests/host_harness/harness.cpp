// CPU harness around the __host__ __device__ per-query code of the CUDA library.
//
// TEST INFRASTRUCTURE ONLY: it lets tests/test_host_logic.py run the exact
// search / fit routines of csrc/pct_grid.cuh and csrc/pct_math.cuh against the
// oracle in the dev container, where there is no GPU.  It is never loaded by the
// product (point_cloud_toolbox_b200 refuses to work without its CUDA library).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (see tests/host_harness/build.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pct_grid.cuh"
#include "pct_dispatch.h"

using namespace pct;

struct HostIndex {
    std::vector<Pt> pts;
    std::vector<std::vector<HashSlot>> tables;
    std::vector<uint32_t> pos_of;  // original index -> sorted position
    IndexView view;
};

static void table_set(std::vector<HashSlot>& t, unsigned long long key, uint32_t start, uint32_t end) {
    const uint32_t mask = (uint32_t)t.size() - 1;
    uint32_t slot = hash_key(key) & mask;
    while (t[slot].key != kEmptyKey) slot = (slot + 1) & mask;
    t[slot].key = key; t[slot].start = start; t[slot].end = end;
}

extern "C" {

void* h_build(const float* xyz, long long n, float h) {
    HostIndex* ix = new HostIndex();
    IndexView& v = ix->view;
    std::memset(&v, 0, sizeof(v));
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (long long i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], xyz[3 * i + a]); hi[a] = std::max(hi[a], xyz[3 * i + a]); }
    v.n = n; v.ox = lo[0]; v.oy = lo[1]; v.oz = lo[2];
    v.h = h; v.inv_h = 1.0f / h;
    int maxdim = 1;
    for (int a = 0; a < 3; ++a) {
        v.dims[a] = (int)cell_coord(hi[a], lo[a], v.inv_h) + 1;
        maxdim = std::max(maxdim, v.dims[a]);
    }
    v.bits = 1;
    while ((1 << v.bits) < maxdim) ++v.bits;
    v.num_levels = v.bits + 1;
    v.slack = 4.0f * 1.1920929e-7f * (float)maxdim + 1e-6f;
    v.volumetric = 0;
    v.slab_axis = -1;
    std::vector<std::pair<unsigned long long, uint32_t>> keyed(n);
    for (long long i = 0; i < n; ++i) {
        int cx, cy, cz;
        cell_of(v, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], cx, cy, cz);
        keyed[i] = {morton3(cx, cy, cz), (uint32_t)i};
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](auto& a, auto& b) { return a.first < b.first; });
    ix->pts.resize(n);
    for (long long i = 0; i < n; ++i) {
        const uint32_t o = keyed[i].second;
        ix->pts[i] = Pt{xyz[3 * o], xyz[3 * o + 1], xyz[3 * o + 2], o};
    }
    ix->pos_of.resize(n);
    for (long long i = 0; i < n; ++i) ix->pos_of[ix->pts[i].idx] = (uint32_t)i;
    ix->tables.resize(v.num_levels);
    for (int L = 0; L < v.num_levels; ++L) {
        long long cells = 0;
        for (long long i = 0; i < n; ++i)
            if (i == 0 || (keyed[i].first >> (3 * L)) != (keyed[i - 1].first >> (3 * L))) ++cells;
        size_t cap = 16;
        while (cap < (size_t)(2 * cells)) cap <<= 1;
        ix->tables[L].assign(cap, HashSlot{kEmptyKey, 0, 0});
        long long start = 0;
        for (long long i = 1; i <= n; ++i)
            if (i == n || (keyed[i].first >> (3 * L)) != (keyed[i - 1].first >> (3 * L))) {
                table_set(ix->tables[L], keyed[start].first >> (3 * L), (uint32_t)start, (uint32_t)i);
                start = i;
            }
        v.lvl[L].slots = ix->tables[L].data();
        v.lvl[L].mask = (uint32_t)cap - 1;
    }
    v.pts = ix->pts.data();
    return ix;
}

void h_destroy(void* p) { delete (HostIndex*)p; }
// the index is one slab of a larger cloud (pct_index_set_slab)
void h_set_slab(void* p, int axis, float complete_lo, float complete_hi, float own_lo, float own_hi) {
    IndexView& v = ((HostIndex*)p)->view;
    v.slab_axis = axis; v.complete_lo = complete_lo; v.complete_hi = complete_hi; v.own_lo = own_lo; v.own_hi = own_hi;
}
void h_perm(void* p, int32_t* perm) {
    HostIndex* ix = (HostIndex*)p;
    for (long long i = 0; i < ix->view.n; ++i) perm[i] = (int32_t)ix->pts[i].idx;
}
int h_levels(void* p) { return ((HostIndex*)p)->view.num_levels; }

}  // extern "C"

// Exact brute-force (k+1)-list minus first, sorted positions; stands in for the GPU's exact kernel.
static void exact_rows(const IndexView& v, uint32_t i, int k, uint32_t* out) {
    const Pt q = v.pts[i];
    std::vector<std::pair<std::pair<double, uint32_t>, uint32_t>> all(v.n);
    for (long long j = 0; j < v.n; ++j) {
        const Pt p = v.pts[j];
        all[j] = {{dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z), p.idx}, (uint32_t)j};
    }
    std::partial_sort(all.begin(), all.begin() + k + 1, all.end());
    for (int m = 0; m < k; ++m) out[m] = all[m + 1].second;
}

// Host emulation of the staged kernel's region tables (csrc/pct_knn_fast.cuh, steps A-D) for one
// chunk of PCT_STAGED_BLOCK consecutive sorted queries: same geometry, same table layout, serial code.
template <int U>
struct HostStage {
    static constexpr int S = RegionShape<U>::kSide, C = RegionShape<U>::kCells;
    std::vector<Pt> pts;
    std::vector<uint32_t> off;  // byte offsets into pts (the device keeps shared-window addresses)
    std::vector<int> org;       // 3 per region
    std::vector<int> region_of; // per query of the chunk
    bool ok = true;
    void build(const IndexView& v, long long begin, long long end, int max_regions, size_t cap_pts) {
        pts.clear(); off.clear(); org.clear(); region_of.clear();
        unsigned long long prev = ~0ull;
        for (long long i = begin; i < end; ++i) {
            int cx, cy, cz;
            cell_of(v, v.pts[i].x, v.pts[i].y, v.pts[i].z, cx, cy, cz);
            const unsigned long long parent = (unsigned long long)(cx >> U) | ((unsigned long long)(cy >> U) << 21) |
                                              ((unsigned long long)(cz >> U) << 42);
            if (i == begin || parent != prev) {
                org.push_back(((cx >> U) << U) - 1);
                org.push_back(((cy >> U) << U) - 1);
                org.push_back(((cz >> U) << U) - 1);
            }
            prev = parent;
            region_of.push_back((int)org.size() / 3 - 1);
        }
        const int regions = (int)org.size() / 3;
        ok = regions <= max_regions;
        if (!ok) return;
        for (int r = 0; r < regions; ++r)
            for (int c = 0; c < C; ++c) {
                const int lz = c / (S * S), ly = (c - lz * S * S) / S, lx = c - lz * S * S - ly * S;
                const int gx = org[3 * r] + lx, gy = org[3 * r + 1] + ly, gz = org[3 * r + 2] + lz;
                uint32_t s = 0, e = 0;
                if (gx >= 0 && gx < v.dims[0] && gy >= 0 && gy < v.dims[1] && gz >= 0 && gz < v.dims[2])
                    if (!lookup_cell(v.lvl[0], morton3((uint32_t)gx, (uint32_t)gy, (uint32_t)gz), s, e)) s = e = 0;
                off.push_back((uint32_t)(pts.size() * sizeof(Pt)));
                if (e - s > 0xffffu || pts.size() + (e - s) > cap_pts) { ok = false; return; }
                for (uint32_t j = s; j < e; ++j) pts.push_back(v.pts[j]);
            }
        off.push_back((uint32_t)(pts.size() * sizeof(Pt)));
    }
};

// kNN lists (original indices, sorted by key) + per-query path code:
// 0..levels-1 = level at which the fast path succeeded, 100 = exact fallback, +50 = staged source.
// staged_u: 0 = candidates straight from the sorted cloud, 1 / 2 = staged regions of (1 << U)^3 cells
template <int U>
static void knn_impl(HostIndex* ix, int k, int max_fast_level, int cap_pts, int32_t* idx, float* dist, int32_t* code,
                     float* normals, float* coeffs, float* curv, uint8_t* status) {
    const IndexView& v = ix->view;
    const int cap = k + PCT_TIE_SLACK;
    std::vector<uint32_t> list(cap), runs(54);
    std::vector<uint16_t> list16(2 * cap + 2);
    std::vector<uint32_t> hist(kHistRowBytes / 4);
    SelectScratch<uint32_t> sc{{list.data(), 1, cap - PCT_TIE_SLACK}, hist.data(), 1, cap};
    SelectScratch<uint16_t> sc16{{list16.data(), 2, cap - PCT_TIE_SLACK}, hist.data(), 1, cap};  // 16-bit: two slots per row
    GlobalSource gsrc;
    gsrc.pts = v.pts;
    gsrc.runs.buf = runs.data();
    gsrc.runs.stride = 1;
    HostStage<(U > 0 ? U : 1)> stage;
    for (long long i = 0; i < v.n; ++i) {
        const Pt q = v.pts[i];
        if (U > 0 && i % PCT_STAGED_BLOCK == 0)
            stage.build(v, i, std::min<long long>(i + PCT_STAGED_BLOCK, v.n), U >= 2 ? 2 + PCT_STAGED_BLOCK / 64 : 4 + PCT_STAGED_BLOCK / 16, (size_t)cap_pts);
        if (!query_owned(v, q.x, q.y, q.z)) { code[q.idx] = -2; continue; }  // another slab answers this one
        uint32_t first = 0, last = 0;
        double d2_last = 0;
        int rc = SEL_RETRY_COARSER, level = 0;
        bool staged = false;
        for (; level <= max_fast_level && level < v.num_levels; ++level) {
            Stencil st;
            make_stencil(v, level, q.x, q.y, q.z, st);
            if (U > 0 && level == 0 && stage.ok) {
                constexpr int S = RegionShape<(U > 0 ? U : 1)>::kSide, C = RegionShape<(U > 0 ? U : 1)>::kCells;
                const int r = stage.region_of[i % PCT_STAGED_BLOCK];
                int cx, cy, cz;
                cell_of(v, q.x, q.y, q.z, cx, cy, cz);
                const int lx = cx - stage.org[3 * r], ly = cy - stage.org[3 * r + 1], lz = cz - stage.org[3 * r + 2];
                StagedSource ssrc;
                ssrc.arena = reinterpret_cast<const char*>(stage.pts.data());
                ssrc.tab = stage.off.data() + r * C;
                ssrc.corner = (lx - 1) + S * (ly - 1) + S * S * (lz - 1);
                ssrc.side = S;
                uint16_t f16 = 0, l16 = 0;
                rc = knn_select(v, st, level, ssrc, q, k, sc16, f16, l16, d2_last);
                if (rc == SEL_OK) {
                    // staged slots -> sorted positions, so that the rest of this routine is shared
                    for (int m = 0; m < k; ++m) list[m] = ix->pos_of[stage.pts[sc16.list.lo(m)].idx];
                    first = ix->pos_of[stage.pts[f16].idx];
                    last = ix->pos_of[stage.pts[l16].idx];
                    staged = true;
                }
            } else {
                gsrc.runs.collect(st);
                rc = knn_select(v, st, level, gsrc, q, k, sc, first, last, d2_last);
            }
            if (rc != SEL_RETRY_COARSER) break;
        }
        const long long row = q.idx;
        bool exact = (rc != SEL_OK);
        if (exact) {
            exact_rows(v, (uint32_t)i, k, list.data());
            first = list[0]; last = list[k - 1];
        } else {
            // order by key for list output
            std::sort(list.begin(), list.begin() + k, [&](uint32_t a, uint32_t b) {
                const Pt pa = v.pts[a], pb = v.pts[b];
                return key_less(dist2_f64(q.x, q.y, q.z, pa.x, pa.y, pa.z), pa.idx,
                                dist2_f64(q.x, q.y, q.z, pb.x, pb.y, pb.z), pb.idx);
            });
            if (list[0] != first || list[k - 1] != last) { code[row] = -1; continue; }  // internal inconsistency
        }
        code[row] = exact ? 100 : level + (staged ? 50 : 0);
        for (int m = 0; m < k; ++m) {
            const Pt p = v.pts[list[m]];
            if (idx) idx[row * k + m] = (int32_t)p.idx;
            if (dist) dist[row * k + m] = (float)sqrt(dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z));
        }
        if (curv) {
            ListNeighbourhood<GlobalSource> nb;
            nb.src = &gsrc; nb.list.base = list.data(); nb.list.stride = 1; nb.list.rows = k; nb.count = k; nb.q = q; nb.first = first; nb.last = last;
            FitResult r;
            r.status = exact ? ST_EXACT_PATH : 0;
            fit_neighbourhood<false>(nb, r);
            for (int c = 0; c < 3; ++c) normals[row * 3 + c] = r.normal[c];
            for (int c = 0; c < 6; ++c) coeffs[row * 6 + c] = r.coeffs[c];
            for (int c = 0; c < 5; ++c) curv[row * 5 + c] = r.curv[c];
            status[row] = (uint8_t)r.status;
        }
    }
}

extern "C" {

void h_knn(void* p, int k, int max_fast_level, int32_t* idx, float* dist, int32_t* code,
           float* normals, float* coeffs, float* curv, uint8_t* status) {
    HostIndex* ix = (HostIndex*)p;
    knn_impl<0>(ix, k, max_fast_level, 0, idx, dist, code, normals, coeffs, curv, status);
}

// same through the staged source (level 0), staging buffer of cap_pts points
void h_knn_staged(void* p, int k, int max_fast_level, int staged_u, int cap_pts, int32_t* idx, float* dist, int32_t* code,
                  float* normals, float* coeffs, float* curv, uint8_t* status) {
    HostIndex* ix = (HostIndex*)p;
    if (staged_u == 0) knn_impl<0>(ix, k, max_fast_level, cap_pts, idx, dist, code, normals, coeffs, curv, status);
    else if (staged_u == 1) knn_impl<1>(ix, k, max_fast_level, cap_pts, idx, dist, code, normals, coeffs, curv, status);
    else knn_impl<2>(ix, k, max_fast_level, cap_pts, idx, dist, code, normals, coeffs, curv, status);
}

void h_fit_rows(const float* xyz, const int32_t* idx, long long nq, int k, const int32_t* qids,
                float* normals, float* coeffs, float* curv, uint8_t* status) {
    for (long long r = 0; r < nq; ++r) {
        const long long qi = qids ? qids[r] : r;
        RowNeighbourhood nb;
        nb.xyz = xyz; nb.row = idx + r * k; nb.count = k; nb.n = 1ll << 40;  // (rows of the tests hold valid non-negative indices)
        nb.qx = xyz[3 * qi]; nb.qy = xyz[3 * qi + 1]; nb.qz = xyz[3 * qi + 2];
        FitResult o;
        o.status = 0;
        fit_neighbourhood<true>(nb, o);
        for (int c = 0; c < 3; ++c) normals[r * 3 + c] = o.normal[c];
        for (int c = 0; c < 6; ++c) coeffs[r * 6 + c] = o.coeffs[c];
        for (int c = 0; c < 5; ++c) curv[r * 5 + c] = o.curv[c];
        status[r] = (uint8_t)o.status;
    }
}

// epsilon ball at the coarsest sufficient level: counts + fused fit
void h_ball(void* p, double radius, int32_t* counts, float* normals, float* coeffs, float* curv, uint8_t* status) {
    HostIndex* ix = (HostIndex*)p;
    const IndexView& v = ix->view;
    int level = 0;
    while (level + 1 < v.num_levels && (double)v.h * (double)(1 << level) * (1.0 - 1e-4) - (double)v.slack * v.h < radius) ++level;
    for (long long i = 0; i < v.n; ++i) {
        StencilSource src;
        src.ix = &v;
        BallNeighbourhood<StencilSource> nb;
        nb.src = &src; nb.q = v.pts[i]; nb.tracked = false; nb.count = 0;
        nb.test.set(radius);
        make_stencil(v, level, nb.q.x, nb.q.y, nb.q.z, src.st);
        FitResult o;
        o.status = 0;
        fit_neighbourhood<true>(nb, o);
        const long long row = nb.q.idx;
        counts[row] = nb.count;
        for (int c = 0; c < 3; ++c) normals[row * 3 + c] = o.normal[c];
        for (int c = 0; c < 6; ++c) coeffs[row * 6 + c] = o.coeffs[c];
        for (int c = 0; c < 5; ++c) curv[row * 5 + c] = o.curv[c];
        status[row] = (uint8_t)o.status;
    }
}

}  // extern "C"

// ---- energy integration and PCA rows: the same __host__ __device__ arithmetic the kernels of
// ---- csrc/pct_energy.cu and csrc/pct_fit.cu (pca_rows_kernel) run, serially
#include "pct_energy.cuh"

extern "C" {

void h_mesh_energies(const float* xyz, long long nv, const int32_t* tri, long long nt, const float* K, const float* H,
                     double* out) {
    double acc[4] = {0, 0, 0, 0};
    for (long long t = 0; t < nt; ++t) {
        long long a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        a += a < 0 ? nv : 0; b += b < 0 ? nv : 0; c += c < 0 ? nv : 0;
        if (a < 0 || b < 0 || c < 0 || a >= nv || b >= nv || c >= nv) { acc[3] += 1.0; continue; }
        const TriangleTerms tt = triangle_terms(xyz + 3 * a, xyz + 3 * b, xyz + 3 * c, K ? K[a] : 0.f, K ? K[b] : 0.f,
                                                K ? K[c] : 0.f, H ? H[a] : 0.f, H ? H[b] : 0.f, H ? H[c] : 0.f);
        acc[0] += tt.bending; acc[1] += tt.stretching; acc[2] += tt.area;
    }
    for (int i = 0; i < 4; ++i) out[i] = acc[i];
}

void h_pca_rows(const float* xyz, const int32_t* idx, long long nq, int k, int include_self, double* values,
                double* directions) {
    for (long long r = 0; r < nq; ++r) {
        const double qx = xyz[3 * r], qy = xyz[3 * r + 1], qz = xyz[3 * r + 2];
        PcaMoments m;
        m.reset();
        if (include_self) m.add(0.0, 0.0, 0.0);
        for (int j = 0; j < k; ++j) {
            const long long p = idx[r * k + j];
            m.add((double)xyz[3 * p] - qx, (double)xyz[3 * p + 1] - qy, (double)xyz[3 * p + 2] - qz);
        }
        double c[6], w[3], v[3][3];
        m.covariance(c);
        eig_sym3_descending(c[0], c[1], c[2], c[3], c[4], c[5], w, v);
        double* o = values + 6 * r;
        o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        o[3] = w[0] * w[1];
        o[4] = (w[0] + w[1]) / 2.0;
        o[5] = w[2] / (w[0] + w[1] + w[2] + 1e-10);
        double* d = directions + 6 * r;
        for (int a = 0; a < 3; ++a) { d[2 * a] = v[a][0]; d[2 * a + 1] = v[a][1]; }
    }
}

}  // extern "C"
