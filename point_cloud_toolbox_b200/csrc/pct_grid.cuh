// Morton-sorted uniform grid: the view a query kernel sees, and the per-query
// search routines (one thread = one query point).
//
// Replaces scipy.spatial.cKDTree as used by PointCloud.plant_kdtree
// (/root/reference/pointCloudToolbox.py:74, :83).  Layout in HBM:
//
//   pts[N]           16-byte records {x, y, z, original index}, sorted by the
//                    Morton key of their level-0 cell
//   level tables     for level L = 0 .. bits: open-addressing hash
//                    {key >> 3L  ->  [start, end) in pts}; an aligned 2^L-cube of
//                    level-0 cells is one contiguous run of pts (Morton property),
//                    so coarser levels need no second copy of the cloud
//
// Search = 3x3x3 cells of one level around the query's cell.  Everything closer
// than `safe` (distance to the faces of that block, minus rounding slack) is
// guaranteed to have been seen; a query whose k-th neighbour is not inside
// `safe` is retried one level up (cells twice as large).
//
// Exactness: candidates are culled with an fp32 squared distance, the survivors
// are re-ranked with scipy's fp64 key ((dx*dx + dy*dy) + dz*dz, ties by index).
// |d32 - d64| <= 5 * 2^-24 * d64, so every true neighbour survives a cull at
// tau * (1 + 2.5e-6) where tau is the k-th smallest d32 (proof in DESIGN.md).
#pragma once

#include "pct_dispatch.h"
#include "pct_math.cuh"

// host test harness only: lets a CPU build count which way pass 2 went
#ifndef PCT_SELECT_TRACE
#define PCT_SELECT_TRACE(used_list)
#endif

namespace pct {

struct alignas(16) Pt {
    float x, y, z;
    uint32_t idx;  // original index
};

struct alignas(16) HashSlot {
    unsigned long long key;
    uint32_t start, end;
};

static constexpr unsigned long long kEmptyKey = ~0ull;
static constexpr int kMaxLevels = 22;

struct LevelTable {
    const HashSlot* slots;
    uint32_t mask;  // capacity - 1 (capacity is a power of two)
    uint32_t pad;
};

struct IndexView {
    const Pt* pts;
    long long n;
    float ox, oy, oz;  // grid origin = bounding-box minimum
    float h, inv_h;    // level-0 cell edge
    float slack;       // rounding slack of a cell coordinate, in level-0 cells
    int dims[3];       // level-0 grid size
    int bits;          // bits per axis of the Morton key
    int num_levels;    // tables for levels 0 .. num_levels-1 (last one: a single cell)
    int volumetric;    // density pilot saw a space-filling cloud (intrinsic dimension > 2.5), not a surface
    // Slab of a spatially partitioned cloud (multi-GPU): the index holds every cloud point whose
    // coordinate on `slab_axis` lies in [complete_lo, complete_hi] and nothing is known about the rest,
    // so no search radius may reach beyond those planes; only queries in [own_lo, own_hi) are answered.
    int slab_axis;     // -1: the index holds the whole cloud
    float complete_lo, complete_hi, own_lo, own_hi;
    LevelTable lvl[kMaxLevels];
};

PCT_HD Pt load_pt(const Pt* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    Pt r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.idx = __float_as_uint(v.w);
    return r;
#else
    return *p;
#endif
}

PCT_HD HashSlot load_slot(const HashSlot* p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    HashSlot s;
    s.key = ((unsigned long long)v.y << 32) | v.x;
    s.start = v.z; s.end = v.w;
    return s;
#else
    return *p;
#endif
}

PCT_HD unsigned long long spread3(uint32_t v) {
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
PCT_HD unsigned long long morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
// sum of the four bytes of a word
PCT_HD uint32_t byte_sum4(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __dp4a(w, 0x01010101u, 0u);
#else
    return (w & 255u) + ((w >> 8) & 255u) + ((w >> 16) & 255u) + (w >> 24);
#endif
}
PCT_HD uint32_t hash_key(unsigned long long key) {
    return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32);
}

// continuous cell coordinate of one axis; monotone in x, identical in build and query
PCT_HD float cell_coord(float x, float origin, float inv_h) { return fmul_rn(fsub_rn(x, origin), inv_h); }

PCT_HD bool lookup_cell(const LevelTable& t, unsigned long long key, uint32_t& start, uint32_t& end) {
    uint32_t slot = hash_key(key) & t.mask;
    for (;;) {
        const HashSlot s = load_slot(t.slots + slot);
        if (s.key == key) { start = s.start; end = s.end; return true; }
        if (s.key == kEmptyKey) return false;
        slot = (slot + 1) & t.mask;
    }
}

// The 3x3x3 block of level-L cells around a query, and how far the query is from leaving it.
struct Stencil {
    int lx, ly, lz;      // the query's cell at this level
    int dx, dy, dz;      // grid size at this level
    float safe2;         // squared radius within which every cloud point has been visited
    const LevelTable* table;
};

PCT_HD int imin_(int a, int b) { return a < b ? a : b; }

// level-0 cell of a point (shared by the build and every query)
PCT_HD void cell_of(const IndexView& ix, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = imin_((int)cell_coord(x, ix.ox, ix.inv_h), ix.dims[0] - 1);
    cy = imin_((int)cell_coord(y, ix.oy, ix.inv_h), ix.dims[1] - 1);
    cz = imin_((int)cell_coord(z, ix.oz, ix.inv_h), ix.dims[2] - 1);
}

// One axis of the block [cell-1, cell+1]: cells below it exist iff cell >= 2,
// cells above it iff cell + 2 <= dim - 1.  `u` is the continuous coordinate.
PCT_HD float axis_gap(float u, int cell, int dim, float gap) {
    const float f = u - (float)cell;  // position inside the cell, [0, 1)
    if (cell >= 2) gap = fminf(gap, f + 1.f);
    if (cell + 2 <= dim - 1) gap = fminf(gap, (1.f - f) + 1.f);
    return gap;
}

PCT_HD void make_stencil(const IndexView& ix, int level, float qx, float qy, float qz, Stencil& st) {
    const float sc = ldexpf(1.f, -level);  // exact
    const float ux = cell_coord(qx, ix.ox, ix.inv_h) * sc;
    const float uy = cell_coord(qy, ix.oy, ix.inv_h) * sc;
    const float uz = cell_coord(qz, ix.oz, ix.inv_h) * sc;
    st.dx = ((ix.dims[0] - 1) >> level) + 1;
    st.dy = ((ix.dims[1] - 1) >> level) + 1;
    st.dz = ((ix.dims[2] - 1) >> level) + 1;
    st.lx = imin_((int)ux, st.dx - 1);
    st.ly = imin_((int)uy, st.dy - 1);
    st.lz = imin_((int)uz, st.dz - 1);
    st.table = &ix.lvl[level];
    // distance (in cells of this level) to the nearest block face beyond which cells exist
    float gap = 3.0e38f;
    gap = axis_gap(ux, st.lx, st.dx, gap);
    gap = axis_gap(uy, st.ly, st.dy, gap);
    gap = axis_gap(uz, st.lz, st.dz, gap);
    float g = 3.0e38f;  // the block covers the whole grid on every axis
    if (gap <= 1.0e38f) g = fmaxf(gap - ix.slack, 0.f) * (ix.h * ldexpf(1.f, level));
    if (ix.slab_axis >= 0) {
        // a point closer than d to the query is at most d away on the slab axis, hence inside the slab
        const float qa = ix.slab_axis == 0 ? qx : (ix.slab_axis == 1 ? qy : qz);
        g = fminf(g, fmaxf(fminf(qa - ix.complete_lo, ix.complete_hi - qa), 0.f));
    }
    st.safe2 = g < 1.0e18f ? g * g * 0.99999f : 3.0e38f;
}

// does this index answer the query (always, unless it is one slab of a partitioned cloud)
PCT_HD bool query_owned(const IndexView& ix, float qx, float qy, float qz) {
    if (ix.slab_axis < 0) return true;
    const float qa = ix.slab_axis == 0 ? qx : (ix.slab_axis == 1 ? qy : qz);
    return qa >= ix.own_lo && qa < ix.own_hi;
}

// cell c (0..26) of the block; false when it lies outside the grid or holds no point
PCT_HD bool stencil_cell(const Stencil& st, int c, uint32_t& s, uint32_t& e) {
    const int cz = st.lz + c / 9 - 1, cy = st.ly + (c / 3) % 3 - 1, cx = st.lx + c % 3 - 1;
    if (cx < 0 || cx >= st.dx || cy < 0 || cy >= st.dy || cz < 0 || cz >= st.dz) return false;
    return lookup_cell(*st.table, morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), s, e);
}

// visit every point of the 27 cells: fn(j, pt) with j the sorted position.
// Deliberately NOT unrolled: 27 inlined copies of a selection body blow the
// instruction cache (ncu r01: stall_no_instruction dominated the fused kernel).
template <class F>
PCT_HD void for_each_candidate(const IndexView& ix, const Stencil& st, F& fn) {
#pragma unroll 1
    for (int c = 0; c < 27; ++c) {
        uint32_t s, e;
        if (!stencil_cell(st, c, s, e)) continue;
#pragma unroll 1
        for (uint32_t j = s; j < e; ++j) {
            const Pt p = load_pt(ix.pts + j);
            fn(j, p);
        }
    }
}

// The non-empty cell runs of a block, kept by the thread that owns the query
// (shared memory on the GPU: word w of run r lives at buf[(2r + w) * stride]).
// Adjacent runs are merged; both selection passes and nothing else read them.
struct CellRuns {
    uint32_t* buf;
    int stride;
    int n;

    PCT_HD void push(uint32_t s, uint32_t e, uint32_t& prev_end) {
        if (s == prev_end) {
            buf[(size_t)(2 * n - 1) * stride] = e;  // extends the previous run
        } else {
            buf[(size_t)(2 * n) * stride] = s;
            buf[(size_t)(2 * n + 1) * stride] = e;
            ++n;
        }
        prev_end = e;
    }

    // The 27 hash probes are issued nine at a time (one z-layer) before any of them
    // is consumed, so their latencies overlap; the Morton bits of the three cell
    // coordinates per axis are spread once and OR-ed together per cell.
    PCT_HD void collect(const Stencil& st) {
        n = 0;
        uint32_t prev_end = 0xffffffffu;
        unsigned long long bx[3], by[3], bz[3];
        bool vx[3], vy[3], vz[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int cx = st.lx + t - 1, cy = st.ly + t - 1, cz = st.lz + t - 1;
            vx[t] = cx >= 0 && cx < st.dx;
            vy[t] = cy >= 0 && cy < st.dy;
            vz[t] = cz >= 0 && cz < st.dz;
            bx[t] = spread3((uint32_t)cx);
            by[t] = spread3((uint32_t)cy) << 1;
            bz[t] = spread3((uint32_t)cz) << 2;
        }
        const LevelTable tab = *st.table;
#pragma unroll
        for (int zi = 0; zi < 3; ++zi) {
            if (!vz[zi]) continue;
            HashSlot probe[9];
            uint32_t slot[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const unsigned long long key = bx[t % 3] | by[t / 3] | bz[zi];
                slot[t] = hash_key(key) & tab.mask;
                probe[t].key = kEmptyKey;
                if (vx[t % 3] && vy[t / 3]) probe[t] = load_slot(tab.slots + slot[t]);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const unsigned long long key = bx[t % 3] | by[t / 3] | bz[zi];
                HashSlot hs = probe[t];
                uint32_t sl = slot[t];
                while (hs.key != key && hs.key != kEmptyKey) {  // linear probing, rarely more than one step
                    sl = (sl + 1) & tab.mask;
                    hs = load_slot(tab.slots + sl);
                }
                if (hs.key == key) push(hs.start, hs.end, prev_end);
            }
        }
    }

    // One flat loop over all candidates (the lanes of a warp stay in the same loop body
    // although their runs differ), in batches of kDepth: the positions of the next kDepth
    // candidates are generated first, their records are loaded together (kDepth
    // independent 16-byte loads in flight per thread) and only then processed.  With the
    // shared-memory carve-out this kernel leaves little L1, so most loads are L2 hits
    // (~300 cycles); one-ahead prefetching left the warps stalled on the scoreboard.
    static constexpr int kDepth = 4;
    template <class F>
    PCT_HD void scan(const Pt* pts, F& fn) const {
        int r = 0;
        uint32_t j = 0, e = 0, safe = 0;
        if (n > 0) safe = buf[0];
#pragma unroll 1
        for (;;) {
            uint32_t pos[kDepth];
            bool valid[kDepth];
#pragma unroll
            for (int u = 0; u < kDepth; ++u) {
                if (j == e && r < n) {
                    j = buf[(size_t)(2 * r) * stride];
                    e = buf[(size_t)(2 * r + 1) * stride];
                    ++r;
                }
                valid[u] = j != e;
                pos[u] = valid[u] ? j : safe;
                j += valid[u] ? 1u : 0u;
            }
            if (!valid[0]) break;
            Pt rec[kDepth];
#pragma unroll
            for (int u = 0; u < kDepth; ++u) rec[u] = load_pt(pts + pos[u]);
#pragma unroll
            for (int u = 0; u < kDepth; ++u)
                if (valid[u]) fn(pos[u], rec[u], true);
        }
    }
};

// ---------------------------------------------------------------------------
// kNN selection, thread per query
// ---------------------------------------------------------------------------
enum SelectCode : int { SEL_OK = 0, SEL_RETRY_COARSER = 1, SEL_EXACT = 2 };

// Selection scratch of one query (two-pass selection).  On the GPU all of it lives in shared memory:
//   runs  54 words, word w at runs[w * stride]           (27 cell runs, GlobalSource only)
//   list  cap entries, addressed through ListRef           (neighbour positions)
//   hist  kHistBins byte counters = 16 words, word w at hist[w * hist_stride]   (distance histogram);
//         with hist_stride = threads of the block every thread stays in its own bank
static constexpr int kHistBins = 64;                 // 32 and 128 measured the same or slower (profiles/variants_r02a.txt)
static constexpr int kHistRowBytes = kHistBins + 4;  // + the word of the overflow counter (candidates beyond the range)

// Where the candidates of a query come from.  knn_select() and the fit only need
//   src.scan(fn)   fn(pos, Pt) for every point of the query's 27 cells
//   src.load(pos)  the record at a position handed out by scan
//   src.count()    number of points in the 27 cells
// GlobalSource reads the Morton-sorted cloud through L1/L2 (positions = sorted positions);
// StagedSource reads the copy a CTA has made in shared memory (positions = 16-bit).
struct GlobalSource {
    typedef uint32_t Pos;
    const Pt* pts;
    CellRuns runs;
    template <class F>
    PCT_HD void scan(F& fn) const { runs.scan(pts, fn); }
    PCT_HD Pt load(uint32_t pos) const { return load_pt(pts + pos); }
    PCT_HD uint32_t count() const {
        uint32_t c = 0;
        for (int r = 0; r < runs.n; ++r)
            c += runs.buf[(size_t)(2 * r + 1) * runs.stride] - runs.buf[(size_t)(2 * r) * runs.stride];
        return c;
    }
};

// Staged candidates.  A CTA copies, for every aligned cube of (1 << U)^3 level-0 cells
// ("parent") its queries fall into, that cube plus a one-cell halo into shared memory:
// a REGION of kSide^3 cells.  The points of a region are laid out cell by cell in raster
// order (x fastest), so the three x-neighbours of a stencil row are one contiguous run
// and the 27 cells of a query are 9 runs, found by direct indexing of a table:
//   tab[c]  byte address of the first record of raster cell c; the tables of the CTA's
//           regions are concatenated, one extra entry closes the last cell
// On the GPU the addresses are shared-window addresses and every access is an explicit
// ld.shared (the compiler otherwise rebuilds the window base inside the candidate loop);
// a position is the record's address / 16, which fits 16 bits.  On the host (test harness)
// addresses are byte offsets into `arena`.
template <int U>
struct RegionShape {
    static constexpr int kSide = (1 << U) + 2;
    static constexpr int kCells = kSide * kSide * kSide;
};

struct StagedSource {
    typedef uint16_t Pos;
#if defined(__CUDACC__)
    uint32_t tab;  // shared address of the table of the query's region
    PCT_HD uint32_t cell_begin(int c) const {
        uint32_t v = 0;
#if defined(__CUDA_ARCH__)
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tab + 4u * (uint32_t)c));
#endif
        return v;
    }
    PCT_HD Pt load_at(uint32_t addr) const {
        Pt p = {};
#if defined(__CUDA_ARCH__)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=r"(p.idx)
                     : "r"(addr));
#endif
        return p;
    }
#else
    const char* arena;
    const uint32_t* tab;
    uint32_t cell_begin(int c) const { return tab[c]; }
    Pt load_at(uint32_t addr) const { return *reinterpret_cast<const Pt*>(arena + addr); }
#endif
    int corner;  // raster index of the lowest cell of the query's 3x3x3 block
    int side;    // kSide

    // One flat loop over the 9 runs: the lanes of a warp sit in different cells, so their
    // runs end at different trips; a flat loop makes the warp pay max-of-sums, not sum-of-maxes.
    // kScanWidth candidates per trip: all records are loaded before any is used (loads may run
    // past the run; they stay inside the block's shared memory and are discarded through
    // `valid`), which overlaps their latencies and divides the loop control.
    static constexpr int kScanWidth = 4;   // 3, 6 and 8 measured within 1.5 % of 4 (profiles/variants_r03b.txt)
    template <class F>
    PCT_HD void scan(F& fn) const {
        int r = 0, c0 = corner;
        uint32_t a = 0, e = 0;
#pragma unroll 1
        for (;;) {
            if (a == e) {
                do {
                    if (r == 9) return;
                    a = cell_begin(c0);
                    e = cell_begin(c0 + 3);
                    ++r;
                    c0 += (r == 3 || r == 6) ? side * side - 2 * side : side;
                } while (a == e);
            }
            // kScanWidth records per trip, all loaded before any is used
            Pt p[kScanWidth];
#pragma unroll
            for (int u = 0; u < kScanWidth; ++u) p[u] = load_at(a + 16u * u);
            const uint32_t left = (e - a) >> 4;
#pragma unroll
            for (int u = 0; u < kScanWidth; ++u) fn((uint16_t)((a >> 4) + u), p[u], u == 0 || (uint32_t)u < left);
            a += left < (uint32_t)kScanWidth ? (e - a) : 16u * kScanWidth;
        }
    }
    PCT_HD Pt load(uint16_t pos) const { return load_at((uint32_t)pos << 4); }
    PCT_HD uint32_t count() const {
        uint32_t c = 0;
        int c0 = corner;
#pragma unroll
        for (int r = 1; r <= 9; ++r) {
            c += cell_begin(c0 + 3) - cell_begin(c0);
            c0 += (r == 3 || r == 6) ? side * side - 2 * side : side;
        }
        return c >> 4;
    }
};

// A per-query list of positions inside an array shared by the threads of a block.  Every thread owns
// one 32-bit word of each row of `threads` words, so rows never make two threads share a bank and a
// thread's words can be reused for another per-thread structure (the histogram) without any
// synchronisation.  The list has `rows` LOW slots (slot m < rows) and, behind them, HIGH slots
// (slot rows + z): 32-bit positions take one row per slot; 16-bit positions keep low slot m in the
// lower half of row m and high slot z in the upper half of row z, so the neighbours (always low
// slots) are addressed with one multiply and the high slots cost no extra rows.
template <class PosT>
struct ListRef {
    PosT* base;  // the thread's first element
    int stride;  // elements of PosT between consecutive rows
    int rows;    // number of low slots
    PCT_HD PosT& lo(int m) const { return base[(size_t)m * stride]; }
    PCT_HD PosT& hi(int z) const {
        return sizeof(PosT) == 2 ? base[(size_t)z * stride + 1] : base[(size_t)(rows + z) * stride];
    }
    PCT_HD PosT& at(int m) const { return m < rows ? lo(m) : hi(m - rows); }
    // bytes one thread needs for `slots` slots of which `rows` are low
    PCT_HD static size_t bytes(int rows, int slots) {
        return sizeof(PosT) == 2 ? 4 * (size_t)(rows > slots - rows ? rows : slots - rows) : sizeof(PosT) * (size_t)slots;
    }
};

template <class PosT>
struct SelectScratch {
    typedef ListRef<PosT> List;
    List list;
    uint32_t* hist; // word w of the byte histogram at hist[w * hist_stride]
    int hist_stride;
    int cap;        // list slots = list.rows low slots (neighbours) + PCT_TIE_SLACK high slots (boundary zone)
};

// Two-pass selection: finds the exact k nearest neighbours (scipy order, self excluded) of query `q` inside the
// level-`level` stencil, in O(candidates) work.  Measured against it on B200 and removed (profiles/README.md, r02):
// one-pass variants that list the candidates below a cut from the local density and select from the list (13.0 /
// 19.3 ms against 11.4 / 16.0 ms at 20 M points, k = 20 / 32: the listing pass costs as many instructions per
// candidate as the histogram pass, and the list walks are chains of dependent shared-memory loads).
//
//   pass 1  histogram of the fp32 squared distances over kHistBins equal bins of
//           [0, range2) (squared distance is uniform in area on a surface, so the
//           bins are evenly filled); the bin b that holds the k-th neighbour follows
//           from a prefix sum.  No sorted list, no dependence on k.
//   pass 2  candidates clearly below bin b (d32 < lo) are neighbours and go to the
//           front of the list; candidates in the boundary zone [lo, hi] -- bin b widened
//           by 1e-5 relative on both sides, far more than the 3e-7 fp32 error -- go
//           to the last PCT_TIE_SLACK slots; everything above hi is out.
//   exact   the k - |front| nearest of the boundary zone are chosen with scipy's fp64
//           key (d2, index).  The zone holds one or two points on average.  If the
//           farthest front point and the nearest zone point are closer than 2e-6
//           relative the cut itself is ambiguous and the query goes to the exact kernel.
//
// On SEL_OK, list.lo(m) (m < k) holds the neighbours' positions (unordered),
// `first` / `last` the nearest / farthest by (d2 fp64, original index).
template <class Source, class Scratch>
PCT_HD int knn_select(const IndexView& ix, const Stencil& st, int level, const Source& src, const Pt& q, int k,
                      const Scratch& sc, typename Source::Pos& first,
                      typename Source::Pos& last, double& d2_last) {
    typedef typename Source::Pos Pos;
    const uint32_t self = q.idx;  // original indices are unique: identifies the query among the candidates

    // everything closer than sqrt(range2) is certain to be among the candidates
    const float cell = ix.h * ldexpf(1.f, level);
    const float range2 = (st.safe2 < 1.0e37f ? st.safe2 : 27.f * cell * cell) * 0.999f;
    const float inv_w = (float)kHistBins / range2 * 0.99999f;  // rounded down: bin < kHistBins for d < range2
    if (!(range2 > 1.0e-30f) || !(inv_w < 3.0e38f)) return SEL_EXACT;
    const int zone_slots = PCT_TIE_SLACK;

#pragma unroll
    for (int w = 0; w < (kHistBins + 4) / 4; ++w) sc.hist[(size_t)w * sc.hist_stride] = 0u;

    // The bodies of both passes are executed by the whole warp whenever one lane needs them,
    // so they are kept short and branch-free; everything that can wait is done on the list afterwards.
    // Pass 1 counts the query itself (bin 0): the k-th neighbour is entry k + 1.  Counters are
    // bytes that may wrap; `seen` detects that afterwards.
    struct P1 {
        uint8_t* hist;
        int hist_stride4;  // bytes between consecutive words
        uint32_t seen;
        float qx, qy, qz, inv_w;
        PCT_HD void operator()(Pos, const Pt& p, bool valid) {
            const float d = valid ? dist2_f32(qx, qy, qz, p.x, p.y, p.z) : 3.4e38f;
            // candidates beyond the range land in the overflow counter (bin kHistBins)
            const int b = (int)fminf(d * inv_w, (float)kHistBins);
            uint8_t* const c = hist + (b >> 2) * hist_stride4 + (b & 3);
            *c = (uint8_t)(*c + 1);
            seen += b < kHistBins ? 1u : 0u;  // (a candidate within 1e-5 of range2 may land in the last bin: still inside safe2)
        }
    } p1;
    p1.hist = reinterpret_cast<uint8_t*>(sc.hist); p1.hist_stride4 = 4 * sc.hist_stride; p1.seen = 0;
    p1.qx = q.x; p1.qy = q.y; p1.qz = q.z; p1.inv_w = inv_w;
    src.scan(p1);

    // bin of the k-th neighbour: word-wise byte sums first (one dp4a per four bins), then the
    // four bins of the word in which the running count passes k
    int b = -1;
    uint32_t cum = 0;
    {
        int w_hit = -1;
        uint32_t cum_hit = 0;
#pragma unroll
        for (int w = 0; w < kHistBins / 4; ++w) {
            const uint32_t s4 = byte_sum4(sc.hist[(size_t)w * sc.hist_stride]);
            const bool hit = w_hit < 0 && cum + s4 > (uint32_t)k;
            w_hit = hit ? w : w_hit;
            cum_hit = hit ? cum : cum_hit;
            cum += s4;
        }
        if (w_hit >= 0) {
            const uint32_t word = sc.hist[(size_t)w_hit * sc.hist_stride];
            uint32_t c = cum_hit;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                c += (word >> (8 * t)) & 255u;
                if (b < 0 && c > (uint32_t)k) b = 4 * w_hit + t;
            }
        }
    }
    if (cum != p1.seen) return SEL_EXACT;  // a counter wrapped (> 255 candidates in one bin)
    if (b < 0) return SEL_RETRY_COARSER;   // fewer than k points within the certain radius

    // bin b is [b, b + 1) / inv_w; widened by 1e-5 relative on both sides
    const float bin_w = 1.f / inv_w;
    const float hi = (float)(b + 1) * bin_w * 1.00001f;
    const float lo = b == 0 ? -1.f : (float)b * bin_w * 0.99999f;

    struct P2 {
        typename Scratch::List list;
        int zone_slots;
        uint32_t self, n_front, n_zone;
        float qx, qy, qz, lo, hi;
        PCT_HD void operator()(Pos j, const Pt& p, bool valid) {
            const float d = valid ? dist2_f32(qx, qy, qz, p.x, p.y, p.z) : 3.4e38f;
            const bool in = d <= hi && p.idx != self, front = d < lo;
            const bool is_front = in && front, is_zone = in && !front;
            Pos* const slot = front ? &list.lo((int)n_front) : &list.hi((int)n_zone);  // n_front < k: always room
            if (is_front || (is_zone && (int)n_zone < zone_slots)) *slot = j;
            n_front += is_front ? 1u : 0u;
            n_zone += is_zone ? 1u : 0u;
        }
    } p2;
    p2.list = sc.list; p2.zone_slots = zone_slots;
    p2.self = self; p2.n_front = 0; p2.n_zone = 0;
    p2.qx = q.x; p2.qy = q.y; p2.qz = q.z; p2.lo = lo; p2.hi = hi;
    src.scan(p2);

    const int n_front = (int)p2.n_front;
    int n_zone = (int)p2.n_zone;
    if (n_zone > zone_slots) return SEL_EXACT;   // a large group of (near-)ties in the boundary bin

    // one walk over the k + few listed candidates: the nearest and its runner-up (fp32), the
    // farthest front member and the nearest zone member
    float d_min = 3.4e38f, d_min2 = 3.4e38f, front_max = 0.f, zone_min = 3.4e38f;
    Pos j_min = 0;
    {
        const int n_list = n_front + n_zone;
#pragma unroll 1
        for (int m = 0; m < n_list; ++m) {
            const bool is_front = m < n_front;
            const Pos j = is_front ? sc.list.lo(m) : sc.list.hi(m - n_front);
            const Pt p = src.load(j);
            const float d = dist2_f32(q.x, q.y, q.z, p.x, p.y, p.z);
            front_max = is_front ? fmaxf(front_max, d) : front_max;
            zone_min = is_front ? zone_min : fminf(zone_min, d);
            const bool closer = d < d_min;
            d_min2 = closer ? d_min : fminf(d_min2, d);
            j_min = closer ? j : j_min;
            d_min = closer ? d : d_min;
        }
    }
    if (!(d_min > 1.0e-30f)) return SEL_EXACT;         // duplicates of the query / denormal range: fp64 only
    // `lo` is an arbitrary cut: front and zone must be separated by more than the fp32 error,
    // otherwise a front member could rank behind a zone member in fp64
    if (n_front > 0 && !(zone_min > front_max * 1.000002f)) return SEL_EXACT;
    const int need = k - n_front;                      // 1 <= need <= n_zone by construction
    if (need < 1 || need > n_zone) return SEL_EXACT;   // (a saturated histogram bin can break the invariant)

    // exact choice inside the boundary zone: `need` successive minima of (d2, index)
    int zone0 = 0;  // first occupied zone slot
    for (int t = 0; t < need; ++t) {
        double bd = 0.0;
        uint32_t bi = 0;
        Pos bj = 0;
        int bm = 0;
        for (int m = 0; m < n_zone; ++m) {
            const Pos j = sc.list.hi(zone0 + m);
            const Pt p = src.load(j);
            const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
            if (m == 0 || key_less(d, p.idx, bd, bi)) { bd = d; bi = p.idx; bj = j; bm = m; }
        }
        // remove entry bm from the zone by moving the zone's first entry into its place
        sc.list.hi(zone0 + bm) = sc.list.hi(zone0);
        ++zone0;
        --n_zone;
        sc.list.lo(n_front + t) = bj;
        last = bj;
        d2_last = bd;
    }

    // nearest neighbour: decided in fp32 when the runner-up is clearly farther
    if (d_min2 > d_min * 1.00001f) {
        first = j_min;
    } else {
        double bd = 0.0;
        uint32_t bi = 0;
        for (int m = 0; m < k; ++m) {
            const Pos j = sc.list.lo(m);
            const Pt p = src.load(j);
            const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
            if (m == 0 || key_less(d, p.idx, bd, bi)) { bd = d; bi = p.idx; first = j; }
        }
    }
    return SEL_OK;
}

// Neighbourhood adaptor over a list of candidate positions (fused kNN path: the neighbours sit in
// the low slots; the ball path fills low and high slots, LOW_ONLY = false).
template <class Source, bool LOW_ONLY = true, class List = ListRef<typename Source::Pos>>
struct ListNeighbourhood {
    typedef typename Source::Pos Pos;
    const Source* src;
    List list;
    int count;
    Pt q;
    Pos first, last;
    template <class F>
    PCT_HD void pass(F& fn) const {
        if (count <= 0) return;
        Pt p = src->load(list.lo(0));
#pragma unroll 1
        for (int m = 0; m < count; ++m) {
            const int mn = m + 1 < count ? m + 1 : m;
            const Pt nxt = src->load(LOW_ONLY ? list.lo(mn) : list.at(mn));  // in flight during the fp64 work below
            fn.add(fsub_rn(p.x, q.x), fsub_rn(p.y, q.y), fsub_rn(p.z, q.z));
            p = nxt;
        }
    }
    PCT_HD void reference(float& rx, float& ry, float& rz) const {
        const Pt a = src->load(first), b = src->load(last);
        rx = fsub_rn(fsub_rn(b.x, q.x), fsub_rn(a.x, q.x));  // ref :286 on fp32 centred points
        ry = fsub_rn(fsub_rn(b.y, q.y), fsub_rn(a.y, q.y));
        rz = fsub_rn(fsub_rn(b.z, q.z), fsub_rn(a.z, q.z));
    }
};

// ---------------------------------------------------------------------------
// epsilon-ball, thread per query, streaming (no lists)
// ---------------------------------------------------------------------------
struct BallTest {
    float r2_lo, r2_hi;  // fp32 brackets of r*r
    double r2;           // fl(r*r) in fp64: scipy's inclusive bound
    PCT_HD void set(double radius) {
        r2 = radius * radius;
        r2_lo = (float)(r2 * (1.0 - 2e-6));
        r2_hi = (float)(r2 * (1.0 + 2e-6));
        if (!(r2_lo > 1.0e-30f)) { r2_lo = -1.f; }  // tiny radii: always take the fp64 test
    }
    PCT_HD bool inside(const Pt& q, const Pt& p) const {
        const float d = dist2_f32(q.x, q.y, q.z, p.x, p.y, p.z);
        if (d > r2_hi) return false;
        if (d < r2_lo) return true;
        return dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z) <= r2;
    }
};

// Candidates of a stencil through L1/L2, cell by cell (ball queries at any grid level).
struct StencilSource {
    typedef uint32_t Pos;
    const IndexView* ix;
    Stencil st;
    template <class F>
    PCT_HD void scan(F& fn) const {
        struct Each {
            F* f;
            PCT_HD void operator()(uint32_t j, const Pt& p) { (*f)(j, p, true); }
        } each;
        each.f = &fn;
        for_each_candidate(*ix, st, each);
    }
};

// Neighbourhood adaptor that re-walks the candidates (fused ball path): the members of the ball are
// whatever passes the radius test, in both passes of the fit.  The first pass also finds the
// nearest / farthest member by (d2, index).
template <class Source>
struct BallNeighbourhood {
    const Source* src;
    BallTest test;
    Pt q;
    // tracking
    double dmin, dmax;
    uint32_t imin, imax;
    Pt pmin, pmax;
    int count;
    bool tracked;

    template <class F>
    struct Visit {
        BallNeighbourhood* nb;
        F* fn;
        bool track;
        PCT_HD void operator()(typename Source::Pos, const Pt& p, bool valid) {
            if (!valid || p.idx == nb->q.idx || !nb->test.inside(nb->q, p)) return;
            fn->add(fsub_rn(p.x, nb->q.x), fsub_rn(p.y, nb->q.y), fsub_rn(p.z, nb->q.z));
            if (track) {
                const double d = dist2_f64(nb->q.x, nb->q.y, nb->q.z, p.x, p.y, p.z);
                if (nb->count == 0 || key_less(d, p.idx, nb->dmin, nb->imin)) { nb->dmin = d; nb->imin = p.idx; nb->pmin = p; }
                if (nb->count == 0 || key_less(nb->dmax, nb->imax, d, p.idx)) { nb->dmax = d; nb->imax = p.idx; nb->pmax = p; }
                ++nb->count;
            }
        }
    };
    template <class F>
    PCT_HD void pass(F& fn) {
        Visit<F> v;
        v.nb = this; v.fn = &fn; v.track = !tracked;
        if (!tracked) count = 0;
        src->scan(v);
        tracked = true;
    }
    PCT_HD void reference(float& rx, float& ry, float& rz) const {
        rx = fsub_rn(fsub_rn(pmax.x, q.x), fsub_rn(pmin.x, q.x));
        ry = fsub_rn(fsub_rn(pmax.y, q.y), fsub_rn(pmin.y, q.y));
        rz = fsub_rn(fsub_rn(pmax.z, q.z), fsub_rn(pmin.z, q.z));
    }
};

// Nearest and farthest member of a listed neighbourhood by (d2 fp64, original index) -- ref :286 needs
// them for the orientation of the normal.  Decided in fp32 when the runner-up is more than 1e-5
// (relative) away, which is 30 times the fp32 error of the distance; otherwise in fp64 with the index
// as tie-break.
template <class Source>
PCT_HD void list_extremes(const Source& src, const ListRef<typename Source::Pos>& list, int n, const Pt& q,
                          typename Source::Pos& first, typename Source::Pos& last) {
    typedef typename Source::Pos Pos;
    float d_min = 3.4e38f, d_min2 = 3.4e38f, d_max = -1.f, d_max2 = -1.f;
    Pos j_min = 0, j_max = 0;
#pragma unroll 1
    for (int m = 0; m < n; ++m) {
        const Pos j = list.at(m);
        const Pt p = src.load(j);
        const float d = dist2_f32(q.x, q.y, q.z, p.x, p.y, p.z);
        const bool closer = d < d_min, farther = d > d_max;
        d_min2 = closer ? d_min : fminf(d_min2, d);
        j_min = closer ? j : j_min;
        d_min = closer ? d : d_min;
        d_max2 = farther ? d_max : fmaxf(d_max2, d);
        j_max = farther ? j : j_max;
        d_max = farther ? d : d_max;
    }
    first = j_min;
    last = j_max;
    const bool min_clear = d_min2 > d_min * 1.00001f && d_min > 1.0e-30f;
    const bool max_clear = d_max2 * 1.00001f < d_max;
    if (min_clear && max_clear) return;
    double bd = 0.0, wd = 0.0;
    uint32_t bi = 0, wi = 0;
    for (int m = 0; m < n; ++m) {
        const Pos j = list.at(m);
        const Pt p = src.load(j);
        const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
        if (!min_clear && (m == 0 || key_less(d, p.idx, bd, bi))) { bd = d; bi = p.idx; first = j; }
        if (!max_clear && (m == 0 || key_less(wd, wi, d, p.idx))) { wd = d; wi = p.idx; last = j; }
    }
}

// Neighbourhood adaptor over caller-provided index rows on the ORIGINAL cloud
// (packed xyz, stride 3): the fit of fit_explicit_quadratic_surfaces_to_neighborhoods
// (ref :638-647) -- first / last are the row's first / last entries, as in the reference.
struct RowNeighbourhood {
    const float* xyz;
    const int32_t* row;
    long long n;  // points in the cloud: a negative index counts from the end (the caller has checked the range)
    int count;
    float qx, qy, qz;
    PCT_HD const float* point(int m) const {
        const long long j = row[m];
        return xyz + 3 * (size_t)(j < 0 ? j + n : j);
    }
    template <class F>
    PCT_HD void pass(F& fn) const {
        for (int m = 0; m < count; ++m) {
            const float* p = point(m);
            fn.add(fsub_rn(p[0], qx), fsub_rn(p[1], qy), fsub_rn(p[2], qz));
        }
    }
    PCT_HD void reference(float& rx, float& ry, float& rz) const {
        const float* a = point(0);
        const float* b = point(count - 1);
        rx = fsub_rn(fsub_rn(b[0], qx), fsub_rn(a[0], qx));
        ry = fsub_rn(fsub_rn(b[1], qy), fsub_rn(a[1], qy));
        rz = fsub_rn(fsub_rn(b[2], qz), fsub_rn(a[2], qz));
    }
};

}  // namespace pct
