// Thread-per-query kNN kernels: histogram selection, exact neighbour set in shared
// memory, fused fp64 fit.  One kernel for every k (the selection is O(candidates)).
//
//   knn_staged_kernel<U, FUSED>  the throughput path.  A CTA owns PCT_STAGED_BLOCK (256) consecutive
//       Morton-sorted queries, copies the cells they can reach (their parent cubes plus
//       a one-cell halo) into shared memory with coalesced 16-byte loads, and every
//       thread then selects and fits its own query out of that copy.  The cloud is read
//       from L2/HBM a handful of times per query instead of once per candidate per pass,
//       and no thread probes the hash table for its own 27 cells.
//   knn_fast_kernel<FUSED>       the same selection reading candidates through L1/L2.
//       Runs over queued queries only: chunks whose regions do not fit the staging
//       buffer (level 0) and queries whose k-th neighbour left the level-0 block
//       (level 1).
//
// See pct_query.cu for the overview of the query path.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "pct_internal.h"

namespace pct {

constexpr int kBlock = 128;

struct QueryRange {
    long long q_begin, q_end;
    const uint32_t* list;       // when non-null: sorted positions to process ...
    const unsigned int* count;  // ... and how many of them
    int layout;
    const int32_t* row_map;     // slab index, PCT_LAYOUT_ORIGINAL: output row of an original index (owned points only)
};

struct Queues {
    uint32_t* retry;     // queries to redo one level coarser (may be null: go straight to exact)
    uint32_t* exact;     // queries for the exact kernel
    uint32_t* fallback;  // staged kernel only: queries of chunks that could not be staged
    uint32_t* rank;      // fused fit only: queries whose design matrix is rank deficient (redone with the minimum-norm solver)
    unsigned int* counters;  // [0] = retry count, [1] = exact count, [2] = fallback count, [3] = unresolved, [4] = rank count
};

__device__ __forceinline__ long long out_row(const QueryRange& qr, uint32_t i, uint32_t orig) {
    if (qr.layout != PCT_LAYOUT_ORIGINAL) return (long long)i - qr.q_begin;
    return qr.row_map ? (long long)qr.row_map[orig] : (long long)orig;
}

// What a query does once its k neighbours sit in list[m * stride]: the fused fit, or the
// ordered (index, distance) rows of plant_kdtree (ref :78-85).
template <bool FUSED, class Source, class List>
__device__ __forceinline__ void emit_query(const Source& src, const List& list, int k, const uint32_t i,
                                           const Pt& q, typename Source::Pos first, typename Source::Pos last,
                                           long long row, int32_t* __restrict__ out_idx, float* __restrict__ out_dist,
                                           const FitOutputs& out, const Queues& qu) {
    if (FUSED) {
        ListNeighbourhood<Source, true, List> nb;
        nb.src = &src; nb.list = list; nb.count = k; nb.q = q; nb.first = first; nb.last = last;
        FitResult r;
        r.status = 0;
        fit_neighbourhood<false>(nb, r);
        if ((r.status & ST_RANK) && qu.rank) {
            qu.rank[atomicAdd(&qu.counters[4], 1u)] = i;  // lstsq's minimum-norm answer comes from knn_exact_kernel<true, true>
            return;
        }
        store_fit(out, row, r);
    } else {
        // ordered rows: successive minima of (d2, index) over the k members
        double pd = -1.0;
        uint32_t pi = 0;
        for (int m = 0; m < k; ++m) {
            double bd = 1.0e300;
            uint32_t bi = 0xffffffffu;
            for (int c = 0; c < k; ++c) {
                const Pt p = src.load(list.lo(c));
                const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
                if (key_less(pd, pi, d, p.idx) && key_less(d, p.idx, bd, bi)) { bd = d; bi = p.idx; }
            }
            if (out_idx) out_idx[row * k + m] = (int32_t)bi;
            if (out_dist) out_dist[row * k + m] = (float)sqrt(bd);
            pd = bd; pi = bi;
        }
    }
}

// ---------------------------------------------------------------------------
// candidates through L1/L2
// ---------------------------------------------------------------------------
// shared memory of one block: [54][kBlock] cell runs, [cap][kBlock] neighbour list, [16][kBlock] histogram words
__host__ __device__ inline size_t fast_smem_bytes(int cap) {
    return sizeof(uint32_t) * (size_t)(54 + cap) * kBlock + (size_t)kHistRowBytes * kBlock;
}

template <bool FUSED>
__global__ void __launch_bounds__(kBlock, 4)
knn_fast_kernel(const IndexView ix, const int level, const QueryRange qr, const int k, const int cap,
                int32_t* __restrict__ out_idx, float* __restrict__ out_dist, const FitOutputs out, const Queues qu) {
    extern __shared__ uint32_t smem_words[];
    GlobalSource src;
    src.pts = ix.pts;
    src.runs.buf = smem_words + threadIdx.x;
    src.runs.stride = kBlock;
    SelectScratch<uint32_t> sc;
    sc.list.base = smem_words + 54 * kBlock + threadIdx.x;
    sc.list.stride = kBlock;
    sc.list.rows = cap - PCT_TIE_SLACK;
    sc.hist = smem_words + (54 + cap) * kBlock + threadIdx.x;
    sc.hist_stride = kBlock;
    sc.cap = cap;
    long long total = qr.q_end - qr.q_begin;
    if (qr.list) total = (long long)*qr.count;
    for (long long base = (long long)blockIdx.x * kBlock; base < total; base += (long long)gridDim.x * kBlock) {
        const long long t = base + threadIdx.x;
        if (t >= total) continue;
        const uint32_t i = qr.list ? qr.list[t] : (uint32_t)(qr.q_begin + t);
        const Pt q = load_pt(ix.pts + i);
        if (!query_owned(ix, q.x, q.y, q.z)) continue;
        Stencil st;
        make_stencil(ix, level, q.x, q.y, q.z, st);
        src.runs.collect(st);
        uint32_t first = 0, last = 0;
        double d2_last = 0.0;
        const int rc = knn_select(ix, st, level, src, q, k, sc, first, last, d2_last);
        if (rc != SEL_OK) {
            if (rc == SEL_RETRY_COARSER && qu.retry && level + 1 < ix.num_levels) {
                qu.retry[atomicAdd(&qu.counters[0], 1u)] = i;
            } else {
                qu.exact[atomicAdd(&qu.counters[1], 1u)] = i;
            }
            continue;
        }
        emit_query<FUSED>(src, sc.list, k, i, q, first, last, out_row(qr, i, q.idx), out_idx, out_dist, out, qu);
    }
}

// ---------------------------------------------------------------------------
// candidates staged in shared memory
// ---------------------------------------------------------------------------
constexpr int kStagedBlock = PCT_STAGED_BLOCK;  // queries (= threads) of one CTA
constexpr int kStagedWarps = kStagedBlock / 32;

template <int U>
struct StageShape : RegionShape<U> {
    // a chunk of kStagedBlock Morton-consecutive queries touches this many parent cubes at most
    static constexpr int kMaxRegions = U >= 2 ? 2 + kStagedBlock / 64 : 4 + kStagedBlock / 16;
    static constexpr int kTable = kMaxRegions * RegionShape<U>::kCells;
    static constexpr int kItemsPerThread = (kTable + kStagedBlock - 1) / kStagedBlock;
    // header words: [0, W) and [W, 2W) block-scan partials, then regions-flag-queue base, then the region origins
    static constexpr int kHdrFlag = 2 * kStagedWarps, kHdrQueue = kHdrFlag + 1, kHdrMbar = kHdrFlag + 2 /* even: 8-byte aligned */,
                         kHdrOrg = kHdrFlag + 4;
    static constexpr int kHdrWords = (kHdrOrg + 3 * kMaxRegions + 3) & ~3;
};

struct StagedCell {
    uint32_t first;  // position in the sorted cloud
    uint16_t slot;   // first staged slot
    uint16_t count;
};

// dynamic shared memory of the staged kernel, in this order (every part 16-byte aligned):
//   Pt       pts[cap_pts]
//   uint32   tab[kTable + 4]         shared address of the first record of every region cell
//   int      hdr[kHdrWords]          block-scan partials, flags, region origins
//   scratch  max(per-query scratch, staging temporaries)
//       per query  : uint16 list[cap][kStagedBlock], uint32 hist[17][kStagedBlock]; the list is first written
//                    after the histogram has been read, so the two share their memory
//       temporaries: uint32 first[kTable], StagedCell cells[kTable], uint16 count[kTable]
template <int U>
__host__ __device__ inline size_t staged_smem_bytes(int cap, int cap_pts) {
    const size_t list = ListRef<uint16_t>::bytes(cap - PCT_TIE_SLACK, cap);
    const size_t hist = kHistRowBytes;
    const size_t per_query = (list > hist ? list : hist) * kStagedBlock;
    const size_t temps = (size_t)StageShape<U>::kTable * (sizeof(uint32_t) + sizeof(uint16_t) + sizeof(StagedCell)) + 64;
    const size_t scratch = per_query > temps ? per_query : temps;
    const size_t tab = (size_t)(StageShape<U>::kTable + 4) * sizeof(uint32_t);
    return sizeof(Pt) * (size_t)cap_pts + tab + StageShape<U>::kHdrWords * sizeof(int) + ((scratch + 15) & ~(size_t)15);
}

#if defined(__CUDACC__)
// ---- TMA bulk copies for the staging phase (sm_90+: cp.async.bulk global -> shared, completion on an mbarrier) ----
// One cell run = one contiguous, 16-byte aligned piece of the sorted cloud = one bulk copy; no register round trip.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// bounded: a lost transaction must not hang the GPU (returns false after ~2^22 polls)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) return true;
    }
    return false;
}

// What staging leaves a thread with: its query, the staged source positioned on the query's region,
// and the block's per-query scratch area (the staging temporaries in it are dead).
struct StagedQuery {
    uint32_t i;      // sorted position of the query
    Pt q;
    StagedSource src;
    char* scratch;   // start of the block's scratch area
};

// Steps A-D of the staged kernels: block-cooperative copy of the cells the chunk's queries can reach.
// Returns false for threads that have nothing to do afterwards: threads beyond the range, queries the
// index does not own (slabs), and every thread of a chunk that does not fit the staging buffer (its
// queries are appended to `fallback`).  All threads of the block must call it.
template <int U>
__device__ __forceinline__ bool stage_chunk(const IndexView& ix, const QueryRange& qr, const int cap_pts,
                                            uint32_t* __restrict__ fallback, unsigned int* __restrict__ fallback_count,
                                            StagedQuery& sq) {
    typedef StageShape<U> Shape;
    constexpr int S = Shape::kSide, C = Shape::kCells, B = kStagedBlock, W = kStagedWarps;
    extern __shared__ uint4 smem_u4[];
    Pt* const pts_s = reinterpret_cast<Pt*>(smem_u4);
    uint32_t* const tab = reinterpret_cast<uint32_t*>(pts_s + cap_pts);
    int* const hdr = reinterpret_cast<int*>(tab + Shape::kTable + 4);
    const uint32_t pts_addr = (uint32_t)__cvta_generic_to_shared(pts_s);
    char* const scratch = reinterpret_cast<char*>(hdr + Shape::kHdrWords);
    int* const org = hdr + Shape::kHdrOrg;
    // staging temporaries (dead before the per-query scratch is first written)
    uint32_t* const t_first = reinterpret_cast<uint32_t*>(scratch);
    StagedCell* const t_cells = reinterpret_cast<StagedCell*>(t_first + Shape::kTable);
    uint16_t* const t_count = reinterpret_cast<uint16_t*>(t_cells + Shape::kTable);
    unsigned long long* const t_parent = reinterpret_cast<unsigned long long*>(t_cells);  // [B], before the cells exist

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long total = qr.q_end - qr.q_begin;
    const long long tq = (long long)blockIdx.x * B + t;
    const bool active = tq < total;
    const uint32_t i = (uint32_t)(qr.q_begin + (active ? tq : total - 1));
    const Pt q = load_pt(ix.pts + i);

    // ---- A. the parent cubes of the chunk (Morton order keeps equal parents adjacent)
    int cx, cy, cz;
    cell_of(ix, q.x, q.y, q.z, cx, cy, cz);
    const unsigned long long parent = (unsigned long long)(cx >> U) | ((unsigned long long)(cy >> U) << 21) |
                                      ((unsigned long long)(cz >> U) << 42);
    t_parent[t] = parent;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(hdr + Shape::kHdrMbar);
    if (t == 0) {
        hdr[Shape::kHdrFlag] = 0;
        mbar_init(mbar, 1);
    }
    __syncthreads();
    const bool head = t == 0 || t_parent[t - 1] != parent;
    const unsigned int heads = __ballot_sync(0xffffffffu, head);
    if (lane == 0) hdr[W + warp] = __popc(heads);
    __syncthreads();
    int region = __popc(heads & (0xffffffffu >> (31 - lane))) - 1;
    int regions = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const int c = hdr[W + w];
        region += w < warp ? c : 0;
        regions += c;
    }
    if (regions > Shape::kMaxRegions) {
        if (t == 0) hdr[Shape::kHdrFlag] = 1;
    } else if (head) {
        org[3 * region] = ((cx >> U) << U) - 1;
        org[3 * region + 1] = ((cy >> U) << U) - 1;
        org[3 * region + 2] = ((cz >> U) << U) - 1;
    }
    __syncthreads();  // t_parent is dead from here on

    // ---- B. one hash probe per region cell
    const int n_table = regions > Shape::kMaxRegions ? 0 : regions * C;
    for (int item = t; item < n_table; item += B) {
        const int r = item / C, c = item - r * C;
        const int lz = c / (S * S), ly = (c - lz * S * S) / S, lx = c - lz * S * S - ly * S;
        const int gx = org[3 * r] + lx, gy = org[3 * r + 1] + ly, gz = org[3 * r + 2] + lz;
        uint32_t s = 0, e = 0;
        if (gx >= 0 && gx < ix.dims[0] && gy >= 0 && gy < ix.dims[1] && gz >= 0 && gz < ix.dims[2])
            if (!lookup_cell(ix.lvl[0], morton3((uint32_t)gx, (uint32_t)gy, (uint32_t)gz), s, e)) s = e = 0;
        uint32_t n = e - s;
        if (n > 0xffffu) { n = 0xffffu; hdr[Shape::kHdrFlag] = 1; }
        t_first[item] = s;
        t_count[item] = (uint16_t)n;
    }
    __syncthreads();

    // ---- C. prefix sums over the table: staged slot of every cell, list of non-empty cells
    uint32_t my_pts = 0, my_cells = 0;
#pragma unroll
    for (int u = 0; u < Shape::kItemsPerThread; ++u) {
        const int item = t * Shape::kItemsPerThread + u;
        const uint32_t n = item < n_table ? t_count[item] : 0u;
        my_pts += n;
        my_cells += n ? 1u : 0u;
    }
    uint32_t inc_pts = my_pts, inc_cells = my_cells;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, inc_pts, o), b = __shfl_up_sync(0xffffffffu, inc_cells, o);
        if (lane >= o) { inc_pts += a; inc_cells += b; }
    }
    if (lane == 31) { hdr[warp] = (int)inc_pts; hdr[W + warp] = (int)inc_cells; }
    __syncthreads();
    uint32_t run = inc_pts - my_pts, cell = inc_cells - my_cells;
    uint32_t staged = 0;
    int n_cells = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t a = (uint32_t)hdr[w], b = (uint32_t)hdr[W + w];
        run += w < warp ? a : 0u;
        cell += w < warp ? b : 0u;
        staged += a;
        n_cells += (int)b;
    }
    const bool unstaged = hdr[Shape::kHdrFlag] != 0 || staged > (uint32_t)cap_pts;
    if (unstaged) {
        // the chunk goes to the L1/L2 kernel as a whole; active threads are a prefix of the block
        const long long left = total - (long long)blockIdx.x * B;
        const unsigned int n_active = left < B ? (unsigned int)left : (unsigned int)B;
        if (t == 0) hdr[Shape::kHdrQueue] = (int)atomicAdd(fallback_count, n_active);
        __syncthreads();
        if (active) fallback[(unsigned int)hdr[Shape::kHdrQueue] + (unsigned int)t] = i;
        return false;
    }
#pragma unroll
    for (int u = 0; u < Shape::kItemsPerThread; ++u) {
        const int item = t * Shape::kItemsPerThread + u;
        if (item < n_table) {
            const uint32_t n = t_count[item];
            tab[item] = pts_addr + 16u * run;
            if (n) {
                StagedCell sc;
                sc.first = t_first[item]; sc.slot = (uint16_t)run; sc.count = (uint16_t)n;
                t_cells[cell++] = sc;
            }
            run += n;
        }
    }
    if (t == 0) {
        tab[n_table] = pts_addr + 16u * staged;
#if PCT_TMA_STAGE
        mbar_arrive_expect_tx(mbar, 16u * staged);  // the one arrival of the phase; the copies below bring the bytes
#endif
    }
    __syncthreads();

#if PCT_TMA_STAGE
    // ---- D. copy the non-empty cells: one TMA bulk copy per cell run (global -> shared, no registers in between),
    // all of them in flight at once, completion counted in bytes on the block's mbarrier
    for (int e = t; e < n_cells; e += B) {
        const StagedCell sc = t_cells[e];
        bulk_copy_g2s(pts_addr + 16u * sc.slot, ix.pts + sc.first, 16u * sc.count, mbar);
    }
    const bool landed = staged == 0 || mbar_wait(mbar, 0);
    __syncthreads();  // temporaries are dead, the per-query scratch may be written
    if (!landed) {    // (never observed; a lost transaction sends the chunk to the L1/L2 kernel instead of hanging)
        if (active) fallback[atomicAdd(fallback_count, 1u)] = i;
        return false;
    }
#else
    // ---- D. copy the non-empty cells, eight lanes per cell (through registers)
    for (int e = t >> 3; e < n_cells; e += B / 8) {
        const StagedCell sc = t_cells[e];
        for (uint32_t m = t & 7; m < sc.count; m += 8) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(ix.pts + sc.first + m));
            reinterpret_cast<float4*>(pts_s)[sc.slot + m] = v;
        }
    }
    __syncthreads();  // temporaries are dead, the per-query scratch may be written
#endif
    if (!active || !query_owned(ix, q.x, q.y, q.z)) return false;
    const int lx = cx - org[3 * region], ly = cy - org[3 * region + 1], lz = cz - org[3 * region + 2];
    sq.i = i;
    sq.q = q;
    sq.src.tab = (uint32_t)__cvta_generic_to_shared(tab + region * C);
    sq.src.corner = (lx - 1) + S * (ly - 1) + S * S * (lz - 1);
    sq.src.side = S;
    sq.scratch = scratch;
    return true;
}

template <int U, bool FUSED>
__global__ void __launch_bounds__(kStagedBlock, PCT_STAGED_CTAS)
knn_staged_kernel(const IndexView ix, const QueryRange qr, const int k, const int cap, const int cap_pts,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_dist, const FitOutputs out, const Queues qu) {
    constexpr int B = kStagedBlock;
    StagedQuery sq;
    if (!stage_chunk<U>(ix, qr, cap_pts, qu.fallback, qu.counters + 2, sq)) return;
    const int t = threadIdx.x;
    const uint32_t i = sq.i;
    const Pt q = sq.q;
    const StagedSource& src = sq.src;
    char* const scratch = sq.scratch;

    // ---- E. select out of the staged copy (two passes over the candidates, see knn_select)
    SelectScratch<uint16_t> sel;
    sel.list.base = reinterpret_cast<uint16_t*>(scratch) + 2 * t;
    sel.list.stride = 2 * B;
    sel.list.rows = cap - PCT_TIE_SLACK;
    sel.hist = reinterpret_cast<uint32_t*>(scratch) + t;
    sel.hist_stride = B;
    sel.cap = cap;
    Stencil st;
    make_stencil(ix, 0, q.x, q.y, q.z, st);
    uint16_t first = 0, last = 0;
    double d2_last = 0.0;
    const int rc = knn_select(ix, st, 0, src, q, k, sel, first, last, d2_last);
    if (rc != SEL_OK) {
        if (rc == SEL_RETRY_COARSER && qu.retry && ix.num_levels > 1) {
            qu.retry[atomicAdd(&qu.counters[0], 1u)] = i;
        } else {
            qu.exact[atomicAdd(&qu.counters[1], 1u)] = i;
        }
        return;
    }
    // ---- F. fit (or ordered rows) out of the staged copy
    emit_query<FUSED>(src, sel.list, k, i, q, first, last, out_row(qr, i, q.idx), out_idx, out_dist, out, qu);
}
#endif

struct FastLaunch {
    const pct_index* ix;
    QueryRange qr;
    int k, cap;  // cap: list slots of the L1/L2 kernel (k + PCT_TIE_SLACK)
    bool fused;
    int32_t* idx;
    float* dist;
    FitOutputs out;
    uint32_t* retry1;
    uint32_t* exactq;
    uint32_t* fallback0;
    uint32_t* rankq;
    unsigned int* counters;
    cudaStream_t s;
};

// staged level 0 over the whole range; L1/L2 kernel at level 0 over the chunks that could
// not be staged, then at level 1 over whatever level 0 queued
template <bool FUSED>
static int launch_fast_impl(const FastLaunch& a, unsigned int* launches) {
    const IndexView& v = a.ix->view;
    const long long nq = a.qr.q_end - a.qr.q_begin;
    const size_t smem = fast_smem_bytes(a.cap);
    const int grid_list = (int)std::min<long long>((nq + kBlock - 1) / kBlock, (long long)a.ix->sm_count * 8);
    auto kern = knn_fast_kernel<FUSED>;
    PCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Queues q0{v.num_levels > 1 ? a.retry1 : nullptr, a.exactq, a.fallback0, a.rankq, a.counters};

    // staging buffer: what is left of this CTA's share of the SM's shared memory after the fixed parts.
    // The kernel is compiled for PCT_STAGED_CTAS resident CTAs; when the regions of a chunk are expected
    // to be larger than that share (large k: cells hold 0.4 k points), fewer, larger CTAs are resident.
    constexpr int U = 2;
    auto staged = knn_staged_kernel<U, FUSED>;
    // at least PCT_TIE_SLACK low slots: the zone lives in the upper halves of the first rows
    const int cap_staged = std::max(a.k, PCT_TIE_SLACK) + PCT_TIE_SLACK;
    const size_t fixed = staged_smem_bytes<U>(cap_staged, 0);
    // a chunk covers (points of one parent cube + chunk) * halo growth points on average; the spread is wide:
    // with a buffer of 2.1 times that mean about 4 % of the chunks do not fit, which still beats giving up a
    // third of the resident warps; below that the unstaged share explodes (k = 40: 26 %; scripts/qbench.py)
    const double per_parent = (double)v.n / (double)std::max<long long>(1, a.ix->cells_level[std::min(U, v.num_levels - 1)]);
    double wanted_gain = 2.1;
    if (const char* e = std::getenv("PCT_STAGED_WANTED")) wanted_gain = std::max(0.5, std::atof(e));  // experiments
    const double wanted = wanted_gain * (per_parent + kStagedBlock) * std::pow(1.5, (double)std::min(3.f, std::max(1.f, a.ix->est_dimension)));
    int cap_pts = 0;
    int ctas_max = PCT_STAGED_CTAS;
    if (const char* e = std::getenv("PCT_STAGED_RESIDENT")) ctas_max = std::max(1, std::min(PCT_STAGED_CTAS, std::atoi(e)));  // experiments
    for (int ctas = ctas_max; ctas >= 1; --ctas) {
        const size_t budget = std::min((size_t)a.ix->smem_per_sm / ctas - 1024, (size_t)a.ix->smem_per_block_optin);
        cap_pts = budget > fixed ? (int)((budget - fixed) / sizeof(Pt)) : 0;
        if ((double)cap_pts >= wanted) break;
    }
    if (cap_pts > 0xffff) cap_pts = 0xffff;
    const size_t smem_staged = staged_smem_bytes<U>(cap_staged, cap_pts);
    if (cap_pts >= 512 && smem_staged <= (size_t)a.ix->smem_per_block_optin) {
        PCT_CUDA(cudaFuncSetAttribute(staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_staged));
        const long long chunks = (nq + kStagedBlock - 1) / kStagedBlock;
        staged<<<(unsigned int)chunks, kStagedBlock, smem_staged, a.s>>>(v, a.qr, a.k, cap_staged, cap_pts, a.idx, a.dist, a.out, q0);
        ++*launches;
        QueryRange qf = a.qr;
        qf.list = a.fallback0;
        qf.count = a.counters + 2;
        kern<<<grid_list, kBlock, smem, a.s>>>(v, 0, qf, a.k, a.cap, a.idx, a.dist, a.out, q0);
        ++*launches;
    } else {
        const int grid_all = (int)std::min<long long>((nq + kBlock - 1) / kBlock, (long long)a.ix->sm_count * 64);
        kern<<<grid_all, kBlock, smem, a.s>>>(v, 0, a.qr, a.k, a.cap, a.idx, a.dist, a.out, q0);
        ++*launches;
    }
    if (v.num_levels > 1) {
        QueryRange q1 = a.qr;
        q1.list = a.retry1;
        q1.count = a.counters;
        Queues qq{nullptr, a.exactq, nullptr, a.rankq, a.counters};
        kern<<<grid_list, kBlock, smem, a.s>>>(v, 1, q1, a.k, a.cap, a.idx, a.dist, a.out, qq);
        ++*launches;
    }
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

static int launch_fast(const FastLaunch& a, unsigned int* launches) {
    return a.fused ? launch_fast_impl<true>(a, launches) : launch_fast_impl<false>(a, launches);
}

}  // namespace pct
