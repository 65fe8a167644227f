// Neighbour search fused with the per-point fit.
//
// Replaces the per-point Python loops of the reference:
//   plant_kdtree                                   /root/reference/pointCloudToolbox.py:81-85
//   fit_explicit_quadratic_surfaces_to_neighborhoods                           :638-647
//   calculate_curvatures_of_explicit_quadratic_surfaces_for_all_points         :663-672
//
// Kernels
//   knn_fast_kernel<FUSED>       one thread per query (queries are consecutive
//       Morton-sorted points, so the lanes of a warp walk the same few cells and
//       their loads hit L1).  Histogram selection of the k-th distance, exact
//       neighbour set in shared memory, then either the fp64 fit (FUSED) or the
//       ordered (index, distance) rows.  Neighbourhoods never go to HBM.
//       Queries that cannot be finished at this grid level are queued.
//   knn_exact_kernel<FUSED>      one warp per queued query: successive minima of
//       the fp64 key over a stencil that grows until it provably holds the
//       (k+1)-list.  Handles arbitrary tie groups, duplicates of the query point
//       and isolated points.  Warp shuffles carry the (d2, index) reductions.
//   ball_kernel<MODE>            epsilon-ball count / CSR fill / fused fit.
#include <algorithm>
#include <cmath>
#include <string>

#include "pct_knn_fast.cuh"

namespace pct {

namespace {


// ---- exact path: one warp per query ----------------------------------------
struct Key {
    double d;
    uint32_t idx;  // original index
    uint32_t pos;  // sorted position
};

__device__ __forceinline__ Key warp_min_key(Key k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        Key other;
        other.d = __shfl_xor_sync(0xffffffffu, k.d, o);
        other.idx = __shfl_xor_sync(0xffffffffu, k.idx, o);
        other.pos = __shfl_xor_sync(0xffffffffu, k.pos, o);
        if (key_less(other.d, other.idx, k.d, k.idx)) k = other;
    }
    return k;
}

constexpr int kExactWarps = 4;
constexpr int kLayoutList = 2;  // internal: output row = position in the queue (pct_knn_points)

// sorted position of cloud points given by original index: the point's own cell holds it
__global__ void locate_kernel(const IndexView ix, const float* __restrict__ xyz, const int stride,
                              const int32_t* __restrict__ ids, const long long nq, uint32_t* __restrict__ positions,
                              unsigned int* __restrict__ missing) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= nq) return;
    const long long id = ids[r];
    const float* p = xyz + id * stride;
    int cx, cy, cz;
    cell_of(ix, p[0], p[1], p[2], cx, cy, cz);
    uint32_t s = 0, e = 0, found = 0xffffffffu;
    if (lookup_cell(ix.lvl[0], morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), s, e))
        for (uint32_t j = s; j < e; ++j)
            if (load_pt(ix.pts + j).idx == (uint32_t)id) { found = j; break; }
    if (found == 0xffffffffu) { atomicAdd(missing, 1u); found = 0; }
    positions[r] = found;
}

// MINNORM: the fit carries lstsq's minimum-norm solver for rank-deficient designs (a stack frame the throughput
// kernels do not want); the other instances queue such queries in `rankq` for it.
template <bool FUSED, bool MINNORM = false>
__global__ void __launch_bounds__(kExactWarps * 32)
knn_exact_kernel(const IndexView ix, const QueryRange qr, const int k, int32_t* __restrict__ out_idx,
                 float* __restrict__ out_dist, const FitOutputs out, const uint32_t* __restrict__ queue,
                 const unsigned int* __restrict__ queue_count, unsigned int* __restrict__ unresolved_count,
                 uint32_t* __restrict__ rankq, unsigned int* __restrict__ rank_count) {
    extern __shared__ uint32_t smem_rows[];  // [kExactWarps][k]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* mine = smem_rows + warp * k;
    const unsigned int total = *queue_count;
    for (unsigned int w = blockIdx.x * kExactWarps + warp; w < total; w += gridDim.x * kExactWarps) {
        const uint32_t i = queue[w];
        const Pt q = load_pt(ix.pts + i);
        // 1. smallest level whose block provably contains the (k+1)-list (self included)
        Stencil st;
        uint32_t cs = 0, ce = 0;  // lane c < 27 owns cell c of the block
        int level = 0;
        for (;; ++level) {
            make_stencil(ix, level, q.x, q.y, q.z, st);
            cs = ce = 0;
            if (lane < 27) {
                const int cx = st.lx + lane % 3 - 1, cy = st.ly + (lane / 3) % 3 - 1, cz = st.lz + lane / 9 - 1;
                if (cx >= 0 && cx < st.dx && cy >= 0 && cy < st.dy && cz >= 0 && cz < st.dz)
                    if (!lookup_cell(*st.table, morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), cs, ce)) cs = ce = 0;
            }
            const bool top = level + 1 >= ix.num_levels;  // the block is the whole grid
            if (top && ix.slab_axis < 0) break;           // ... and the grid the whole cloud
            const double safe2 = (double)st.safe2;
            unsigned int inside = 0;
            for (int c = 0; c < 27; ++c) {
                const uint32_t s = __shfl_sync(0xffffffffu, cs, c), e = __shfl_sync(0xffffffffu, ce, c);
                for (uint32_t j = s + lane; j < e; j += 32) {
                    const Pt p = load_pt(ix.pts + j);
                    inside += dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z) < safe2 ? 1u : 0u;
                }
            }
            inside = __reduce_add_sync(0xffffffffu, inside);
            if (inside >= (unsigned int)(k + 1)) break;
            if (top) { level = -1; break; }  // slab: the k-th neighbour may lie outside what this index holds
        }
        if (level < 0) {
            if (lane == 0) {
                if (unresolved_count) atomicAdd(unresolved_count, 1u);
                const long long row = qr.layout == kLayoutList ? (long long)w : out_row(qr, i, q.idx);
                if (FUSED) {
                    FitResult r;
                    r.status = 0;
                    fit_fail(r, ST_UNRESOLVED);
                    store_fit(out, row, r);
                } else {
                    for (int m = 0; m < k; ++m) {
                        if (out_idx) out_idx[row * k + m] = -1;
                        if (out_dist) out_dist[row * k + m] = nanf("");
                    }
                }
            }
            continue;
        }
        const double bound = (level + 1 >= ix.num_levels && ix.slab_axis < 0) ? 1.0e300 : (double)st.safe2;
        // 2. k+1 successive minima of (d2, index); the first one is dropped (ref :84-85)
        Key prev;
        prev.d = -1.0; prev.idx = 0; prev.pos = 0;
        for (int m = 0; m <= k; ++m) {
            Key best;
            best.d = 1.0e301; best.idx = 0xffffffffu; best.pos = 0;
            for (int c = 0; c < 27; ++c) {
                const uint32_t s = __shfl_sync(0xffffffffu, cs, c), e = __shfl_sync(0xffffffffu, ce, c);
                for (uint32_t j = s + lane; j < e; j += 32) {
                    const Pt p = load_pt(ix.pts + j);
                    const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
                    if (d < bound && key_less(prev.d, prev.idx, d, p.idx) && key_less(d, p.idx, best.d, best.idx)) {
                        best.d = d; best.idx = p.idx; best.pos = j;
                    }
                }
            }
            best = warp_min_key(best);
            if (m >= 1 && lane == 0) {
                mine[m - 1] = best.pos;
                const long long row = qr.layout == kLayoutList ? (long long)w : out_row(qr, i, q.idx);
                if (!FUSED) {
                    if (out_idx) out_idx[row * k + (m - 1)] = (int32_t)best.idx;
                    if (out_dist) out_dist[row * k + (m - 1)] = (float)sqrt(best.d);
                }
            }
            prev = best;
        }
        __syncwarp();
        if (FUSED && lane == 0) {
            GlobalSource src;
            src.pts = ix.pts;
            ListNeighbourhood<GlobalSource> nb;
            nb.src = &src; nb.list.base = mine; nb.list.stride = 1; nb.list.rows = k; nb.count = k; nb.q = q; nb.first = mine[0]; nb.last = mine[k - 1];
            FitResult r;
            r.status = ST_EXACT_PATH;
            fit_neighbourhood<MINNORM>(nb, r);
            if (!MINNORM && (r.status & ST_RANK) && rankq)
                rankq[atomicAdd(rank_count, 1u)] = i;
            else
                store_fit(out, qr.layout == kLayoutList ? (long long)w : out_row(qr, i, q.idx), r);
        }
        __syncwarp();
    }
}

// ---- kdtree.query(x, k) for ARBITRARY coordinates (ref :83, :759 only ever pass cloud points, scipy takes any x) ----
// One warp per query point: smallest level whose 3x3x3 block provably holds the k nearest cloud points, then k
// successive minima of (d2 fp64, original index).  Nothing is dropped: a query that is a cloud point finds itself
// first, at distance 0, like scipy.  A query outside the grid's box is answered at the top level (the block is
// the whole cloud): correct, and slow for large clouds.
__global__ void __launch_bounds__(kExactWarps * 32)
knn_query_kernel(const IndexView ix, const float* __restrict__ queries, const long long nq, const int k,
                 int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long w = (long long)blockIdx.x * kExactWarps + warp; w < nq; w += (long long)gridDim.x * kExactWarps) {
        const float qx = __ldg(queries + 3 * w), qy = __ldg(queries + 3 * w + 1), qz = __ldg(queries + 3 * w + 2);
        const float ux = cell_coord(qx, ix.ox, ix.inv_h), uy = cell_coord(qy, ix.oy, ix.inv_h), uz = cell_coord(qz, ix.oz, ix.inv_h);
        const bool inside = ux >= 0.f && uy >= 0.f && uz >= 0.f && ux < (float)ix.dims[0] && uy < (float)ix.dims[1] && uz < (float)ix.dims[2];
        Stencil st;
        uint32_t cs = 0, ce = 0;
        int level = inside ? 0 : ix.num_levels - 1;
        for (;; ++level) {
            make_stencil(ix, level, qx, qy, qz, st);
            if (!inside) { st.lx = st.ly = st.lz = 0; }   // top level: one cell
            cs = ce = 0;
            if (lane < 27) {
                const int cx = st.lx + lane % 3 - 1, cy = st.ly + (lane / 3) % 3 - 1, cz = st.lz + lane / 9 - 1;
                if (cx >= 0 && cx < st.dx && cy >= 0 && cy < st.dy && cz >= 0 && cz < st.dz)
                    if (!lookup_cell(*st.table, morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz), cs, ce)) cs = ce = 0;
            }
            if (level + 1 >= ix.num_levels) break;  // the block is the whole grid
            const double safe2 = (double)st.safe2;
            unsigned int in_safe = 0;
            for (int c = 0; c < 27; ++c) {
                const uint32_t s = __shfl_sync(0xffffffffu, cs, c), e = __shfl_sync(0xffffffffu, ce, c);
                for (uint32_t j = s + lane; j < e; j += 32) {
                    const Pt p = load_pt(ix.pts + j);
                    in_safe += dist2_f64(qx, qy, qz, p.x, p.y, p.z) < safe2 ? 1u : 0u;
                }
            }
            in_safe = __reduce_add_sync(0xffffffffu, in_safe);
            if (in_safe >= (unsigned int)k) break;
        }
        const double bound = level + 1 >= ix.num_levels ? 1.0e300 : (double)st.safe2;
        Key prev;
        prev.d = -1.0; prev.idx = 0; prev.pos = 0;
        for (int m = 0; m < k; ++m) {
            Key best;
            best.d = 1.0e301; best.idx = 0xffffffffu; best.pos = 0;
            for (int c = 0; c < 27; ++c) {
                const uint32_t s = __shfl_sync(0xffffffffu, cs, c), e = __shfl_sync(0xffffffffu, ce, c);
                for (uint32_t j = s + lane; j < e; j += 32) {
                    const Pt p = load_pt(ix.pts + j);
                    const double d = dist2_f64(qx, qy, qz, p.x, p.y, p.z);
                    if (d < bound && key_less(prev.d, prev.idx, d, p.idx) && key_less(d, p.idx, best.d, best.idx)) {
                        best.d = d; best.idx = p.idx; best.pos = j;
                    }
                }
            }
            best = warp_min_key(best);
            if (lane == 0) {
                out_idx[w * k + m] = (int32_t)best.idx;
                out_dist[w * k + m] = sqrt(best.d);
            }
            prev = best;
        }
    }
}

__global__ void publish_stats_kernel(const unsigned int* counters, unsigned int* stats, unsigned int queries, unsigned int launches) {
    stats[0] = counters[0];
    stats[1] = counters[1];
    stats[2] = launches;
    stats[3] = queries;
    stats[4] = counters[2];
    stats[5] = counters[3];
    stats[6] = counters[4];
}

// ---- epsilon ball ----------------------------------------------------------
enum BallMode : int { BALL_COUNT = 0, BALL_FILL = 1, BALL_FUSED = 2 };

struct CountOnly {
    int n;
    __device__ __forceinline__ void add(float, float, float) { ++n; }
};

// MINNORM (fused mode): the instance that redoes the rank-deficient balls the others queue in `rankq`
template <int MODE, bool MINNORM = false>
__global__ void __launch_bounds__(kBlock)
ball_kernel(const IndexView ix, const int level, const QueryRange qr, const double radius, int32_t* __restrict__ counts,
            const long long* __restrict__ offsets, int32_t* __restrict__ out_idx, float* __restrict__ out_dist,
            double* __restrict__ scratch_d2, const FitOutputs out, uint32_t* __restrict__ rankq = nullptr,
            unsigned int* __restrict__ rank_count = nullptr) {
    const long long total = qr.list ? (long long)*qr.count : qr.q_end - qr.q_begin;
    for (long long t = (long long)blockIdx.x * kBlock + threadIdx.x; t < total; t += (long long)gridDim.x * kBlock) {
        const uint32_t i = qr.list ? qr.list[t] : (uint32_t)(qr.q_begin + t);
        StencilSource src;
        src.ix = &ix;
        BallNeighbourhood<StencilSource> nb;
        nb.src = &src; nb.q = load_pt(ix.pts + i); nb.tracked = false; nb.count = 0;
        nb.test.set(radius);
        make_stencil(ix, level, nb.q.x, nb.q.y, nb.q.z, src.st);
        const long long row = out_row(qr, i, nb.q.idx);
        if (MODE == BALL_COUNT) {
            CountOnly c;
            c.n = 0;
            nb.tracked = true;  // no first/last bookkeeping needed
            nb.pass(c);
            counts[row] = c.n;
        } else if (MODE == BALL_FUSED) {
            FitResult r;
            r.status = 0;
            fit_neighbourhood<MINNORM>(nb, r);
            if (counts) counts[row] = nb.count;
            if (!MINNORM && (r.status & ST_RANK) && rankq) {
                rankq[atomicAdd(rank_count, 1u)] = i;
                continue;
            }
            store_fit(out, row, r);
        } else {
            // CSR fill: members in walk order, then insertion sort of the row by (d2, index)
            const long long o = offsets[row];
            int n = 0;
            struct Walk {
                BallNeighbourhood<StencilSource>* nb;
                long long o;
                int* n;
                int32_t* idx;
                double* d2;
                __device__ __forceinline__ void operator()(uint32_t, const Pt& p, bool) {
                    if (p.idx == nb->q.idx || !nb->test.inside(nb->q, p)) return;
                    const double d = dist2_f64(nb->q.x, nb->q.y, nb->q.z, p.x, p.y, p.z);
                    int m = (*n)++;
                    // insertion from the back keeps the row ordered as it grows
                    while (m > 0 && key_less(d, p.idx, d2[o + m - 1], (uint32_t)idx[o + m - 1])) {
                        d2[o + m] = d2[o + m - 1];
                        idx[o + m] = idx[o + m - 1];
                        --m;
                    }
                    d2[o + m] = d;
                    idx[o + m] = (int32_t)p.idx;
                }
            } wk;
            wk.nb = &nb; wk.o = o; wk.n = &n; wk.idx = out_idx; wk.d2 = scratch_d2;
            src.scan(wk);
            if (out_dist)
                for (int m = 0; m < n; ++m) out_dist[o + m] = (float)sqrt(scratch_d2[o + m]);
        }
    }
}

// Fused ball fit out of the staged copy (level 0: the cell edge is at least the radius, so the 3x3x3 block
// holds the ball).  Same staging as the kNN kernel; one pass over the candidates lists the members of the
// ball in shared memory, then the fit runs over that list exactly like the kNN fit.  Balls with more
// members than the list holds and chunks that do not fit the staging buffer go to ball_kernel (which
// streams the candidates twice) through a queue.
constexpr int kBallListSlots = 64;

template <int U, int MODE>
__global__ void __launch_bounds__(kStagedBlock, PCT_STAGED_CTAS)
ball_staged_kernel(const IndexView ix, const QueryRange qr, const double radius, const int cap_pts,
                   int32_t* __restrict__ counts, const FitOutputs out, const long long* __restrict__ offsets,
                   int32_t* __restrict__ out_idx, float* __restrict__ out_dist, uint32_t* __restrict__ fallback,
                   unsigned int* __restrict__ fallback_count, uint32_t* __restrict__ rankq) {
    StagedQuery sq;
    if (!stage_chunk<U>(ix, qr, cap_pts, fallback, fallback_count, sq)) return;
    const Pt q = sq.q;
    const uint32_t qi = sq.i;
    const StagedSource& src = sq.src;
    ListRef<uint16_t> list;
    list.base = reinterpret_cast<uint16_t*>(sq.scratch) + 2 * threadIdx.x;
    list.stride = 2 * kStagedBlock;
    list.rows = kBallListSlots / 2;  // both halves of every row are used
    struct Collect {
        ListRef<uint16_t> list;
        BallTest test;
        Pt q;
        int n;
        __device__ __forceinline__ void operator()(uint16_t j, const Pt& p, bool valid) {
            if (!valid || p.idx == q.idx || !test.inside(q, p)) return;
            if (n < kBallListSlots) list.at(n) = j;
            ++n;
        }
    } col;
    col.list = list; col.q = q; col.n = 0;
    col.test.set(radius);
    src.scan(col);
    if (col.n > kBallListSlots) {  // a ball larger than the list: streamed by ball_kernel
        fallback[atomicAdd(fallback_count, 1u)] = qi;
        return;
    }
    const long long row = out_row(qr, qi, q.idx);
    if (MODE == BALL_FILL) {
        // CSR row ordered by (d2, index): successive minima over the listed members
        const long long o = offsets[row];
        double pd = -1.0;
        uint32_t pi = 0;
        for (int m = 0; m < col.n; ++m) {
            double bd = 1.0e300;
            uint32_t bi = 0xffffffffu;
            for (int c = 0; c < col.n; ++c) {
                const Pt p = src.load(list.at(c));
                const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
                if (key_less(pd, pi, d, p.idx) && key_less(d, p.idx, bd, bi)) { bd = d; bi = p.idx; }
            }
            out_idx[o + m] = (int32_t)bi;
            if (out_dist) out_dist[o + m] = (float)sqrt(bd);
            pd = bd; pi = bi;
        }
        return;
    }
    FitResult r;
    r.status = 0;
    ListNeighbourhood<StagedSource, false> nb;
    nb.src = &src; nb.list = list; nb.count = col.n; nb.q = q; nb.first = 0; nb.last = 0;
    if (col.n >= 2) list_extremes(src, list, col.n, q, nb.first, nb.last);
    fit_neighbourhood<false>(nb, r);
    if (counts) counts[row] = col.n;
    if ((r.status & ST_RANK) && rankq) {
        rankq[atomicAdd(fallback_count + 1, 1u)] = qi;  // redone by ball_kernel<BALL_FUSED, true>
        return;
    }
    store_fit(out, row, r);
}

int ball_level(const IndexView& v, double radius) {
    // smallest level whose cell edge (less rounding slack) is >= radius: then the 3x3x3 block holds the ball
    int level = 0;
    while (level + 1 < v.num_levels) {
        const double edge = (double)v.h * std::ldexp(1.0, level);
        const double safe = (1.0 - (double)v.slack) * edge * 0.99999;
        if (safe * safe * 0.99999 > radius * radius) break;
        ++level;
    }
    return level;
}

}  // namespace

int launch_knn(const pct_index* ix, long long q_begin, long long q_end, int k, bool fused, int32_t* idx, float* dist,
               FitOutputs out, int layout, cudaStream_t s) {
    const IndexView& v = ix->view;
    const long long nq = q_end - q_begin;
    if (nq == 0) return PCT_OK;
    const int cap = k + PCT_TIE_SLACK;
    // four work queues of up to nq entries each, from the stream's scratch arena
    ScratchSession scratch(s, sizeof(uint32_t) * 4 * (size_t)nq + 4096);
    uint32_t* queues = static_cast<uint32_t*>(scratch.take(sizeof(uint32_t) * 4 * (size_t)nq));
    unsigned int* counters = static_cast<unsigned int*>(scratch.take(sizeof(unsigned int) * 8));
    const bool pooled = !queues || !counters;
    if (pooled) {
        PCT_CUDA(cudaMallocAsync(&queues, sizeof(uint32_t) * 4 * (size_t)nq, s));
        PCT_CUDA(cudaMallocAsync(&counters, sizeof(unsigned int) * 8, s));
    }
    PCT_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned int) * 8, s));
    uint32_t* retry1 = queues;
    uint32_t* exactq = queues + nq;
    uint32_t* fallback0 = queues + 2 * nq;
    uint32_t* rankq = fused ? queues + 3 * nq : nullptr;
    QueryRange qr{q_begin, q_end, nullptr, nullptr, layout, ix->row_map};
    unsigned int launches = 0;
    FastLaunch fl{ix, qr, k, cap, fused, idx, dist, out, retry1, exactq, fallback0, rankq, counters, s};
    int rc = PCT_OK;
    rc = launch_fast(fl, &launches);
    if (rc != PCT_OK) return rc;

    const size_t smem_exact = sizeof(uint32_t) * (size_t)k * kExactWarps;
    const int grid_exact = ix->sm_count * 4;
    if (fused) {
        knn_exact_kernel<true><<<grid_exact, kExactWarps * 32, smem_exact, s>>>(v, qr, k, idx, dist, out, exactq, counters + 1, counters + 3,
                                                                               rankq, counters + 4);
        // rank-deficient neighbourhoods (collinear / duplicated / lattice points): lstsq's minimum-norm solution
        knn_exact_kernel<true, true><<<grid_exact, kExactWarps * 32, smem_exact, s>>>(v, qr, k, idx, dist, out, rankq, counters + 4, counters + 3,
                                                                                     nullptr, nullptr);
        ++launches;
    } else {
        knn_exact_kernel<false><<<grid_exact, kExactWarps * 32, smem_exact, s>>>(v, qr, k, idx, dist, out, exactq, counters + 1, counters + 3,
                                                                                nullptr, nullptr);
    }
    ++launches;
    publish_stats_kernel<<<1, 1, 0, s>>>(counters, ix->stats, (unsigned int)nq, launches);
    PCT_CUDA(cudaGetLastError());
    if (pooled) {
        PCT_CUDA(cudaFreeAsync(queues, s));
        PCT_CUDA(cudaFreeAsync(counters, s));
    }
    return PCT_OK;
}

int launch_knn_points(const pct_index* ix, const float* xyz, int stride, const int32_t* ids, long long nq, int k,
                      int32_t* idx, float* dist, float* records, cudaStream_t s) {
    const IndexView& v = ix->view;
    if (nq == 0) return PCT_OK;
    uint32_t* positions = nullptr;
    unsigned int* counters = nullptr;  // [0] = queue length, [1] = ids that are not cloud points
    PCT_CUDA(cudaMallocAsync(&positions, sizeof(uint32_t) * (size_t)nq, s));
    PCT_CUDA(cudaMallocAsync(&counters, sizeof(unsigned int) * 2, s));
    const unsigned int h_init[2] = {(unsigned int)nq, 0u};
    PCT_CUDA(cudaMemcpyAsync(counters, h_init, sizeof(h_init), cudaMemcpyHostToDevice, s));
    locate_kernel<<<(int)((nq + 127) / 128), 128, 0, s>>>(v, xyz, stride, ids, nq, positions, counters + 1);
    QueryRange qr{0, v.n, nullptr, nullptr, kLayoutList, nullptr};
    FitOutputs outs{nullptr, nullptr, nullptr, nullptr, records};
    const size_t smem_exact = sizeof(uint32_t) * (size_t)k * kExactWarps;
    const int grid = (int)std::min<long long>((nq + kExactWarps - 1) / kExactWarps, (long long)ix->sm_count * 8);
    if (records)
        knn_exact_kernel<true, true><<<grid, kExactWarps * 32, smem_exact, s>>>(v, qr, k, nullptr, nullptr, outs, positions, counters, nullptr,
                                                                               nullptr, nullptr);
    else
        knn_exact_kernel<false><<<grid, kExactWarps * 32, smem_exact, s>>>(v, qr, k, idx, dist, outs, positions, counters, nullptr,
                                                                          nullptr, nullptr);
    PCT_CUDA(cudaGetLastError());
    unsigned int h_missing = 0;
    PCT_CUDA(cudaMemcpyAsync(&h_missing, counters + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));
    PCT_CUDA(cudaFreeAsync(positions, s));
    PCT_CUDA(cudaFreeAsync(counters, s));
    if (h_missing) {
        set_error("pct_knn_points: " + std::to_string(h_missing) + " query ids do not name points of the indexed cloud");
        return PCT_ERR_INVALID_ARGUMENT;
    }
    return PCT_OK;
}

int launch_knn_query(const pct_index* ix, const float* queries, long long nq, int k, int32_t* idx, double* dist, cudaStream_t s) {
    if (nq == 0) return PCT_OK;
    const int grid = (int)std::min<long long>((nq + kExactWarps - 1) / kExactWarps, (long long)ix->sm_count * 8);
    knn_query_kernel<<<grid, kExactWarps * 32, 0, s>>>(ix->view, queries, nq, k, idx, dist);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int launch_ball(const pct_index* ix, long long q_begin, long long q_end, double radius, int mode, int32_t* counts,
                const long long* offsets, long long nnz, int32_t* idx, float* dist, FitOutputs out, int layout,
                cudaStream_t s) {
    const IndexView& v = ix->view;
    const long long nq = q_end - q_begin;
    if (nq == 0) return PCT_OK;
    const int level = ball_level(v, radius);
    QueryRange qr{q_begin, q_end, nullptr, nullptr, layout, ix->row_map};
    const int grid = (int)std::min<long long>((nq + kBlock - 1) / kBlock, (long long)ix->sm_count * 64);
    if (mode == BALL_COUNT) {
        ball_kernel<BALL_COUNT><<<grid, kBlock, 0, s>>>(v, level, qr, radius, counts, nullptr, nullptr, nullptr, nullptr, out);
    } else {
        constexpr int U = 2;
        const size_t fixed = staged_smem_bytes<U>(kBallListSlots / 2 + PCT_TIE_SLACK, 0);
        const size_t budget = std::min((size_t)ix->smem_per_sm / PCT_STAGED_CTAS - 1024, (size_t)ix->smem_per_block_optin);
        const int cap_pts = (int)std::min<size_t>(budget > fixed ? (budget - fixed) / sizeof(Pt) : 0, 0xffff);
        const size_t d2_bytes = mode == BALL_FILL ? sizeof(double) * (size_t)std::max<long long>(nnz, 1) : 0;
        ScratchSession scratch(s, sizeof(uint32_t) * 2 * (size_t)nq + 8192 + d2_bytes);
        uint32_t* fallback = static_cast<uint32_t*>(scratch.take(sizeof(uint32_t) * (size_t)nq));
        uint32_t* rankq = mode == BALL_FUSED ? static_cast<uint32_t*>(scratch.take(sizeof(uint32_t) * (size_t)nq)) : nullptr;
        unsigned int* fb_count = static_cast<unsigned int*>(scratch.take(sizeof(unsigned int) * 4));
        // squared distances of the rows ball_kernel<fill> sorts by insertion
        double* scratch_d2 = d2_bytes ? static_cast<double*>(scratch.take(d2_bytes)) : nullptr;
        bool d2_pooled = false;
        if (d2_bytes && !scratch_d2) {
            PCT_CUDA(cudaMallocAsync(&scratch_d2, d2_bytes, s));
            d2_pooled = true;
        }
        if (level == 0 && cap_pts >= 512 && fallback && fb_count) {
            // staged kernel over the whole range, L1/L2 kernel over the chunks and balls that did not fit
            const size_t smem = staged_smem_bytes<U>(kBallListSlots / 2 + PCT_TIE_SLACK, cap_pts);
            PCT_CUDA(cudaMemsetAsync(fb_count, 0, sizeof(unsigned int) * 4, s));
            const long long chunks = (nq + kStagedBlock - 1) / kStagedBlock;
            QueryRange ql = qr;
            ql.list = fallback;
            ql.count = fb_count;
            const int grid_list = (int)std::min<long long>((nq + kBlock - 1) / kBlock, (long long)ix->sm_count * 8);
            if (mode == BALL_FUSED) {
                PCT_CUDA(cudaFuncSetAttribute(ball_staged_kernel<U, BALL_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ball_staged_kernel<U, BALL_FUSED><<<(unsigned int)chunks, kStagedBlock, smem, s>>>(
                    v, qr, radius, cap_pts, counts, out, nullptr, nullptr, nullptr, fallback, fb_count, rankq);
                ball_kernel<BALL_FUSED><<<grid_list, kBlock, 0, s>>>(v, level, ql, radius, counts, nullptr, nullptr, nullptr, nullptr, out,
                                                                     rankq, fb_count + 1);
                if (rankq) {  // rank-deficient balls: lstsq's minimum-norm solution
                    QueryRange qk = qr;
                    qk.list = rankq;
                    qk.count = fb_count + 1;
                    ball_kernel<BALL_FUSED, true><<<grid_list, kBlock, 0, s>>>(v, level, qk, radius, counts, nullptr, nullptr, nullptr, nullptr, out);
                }
            } else {
                PCT_CUDA(cudaFuncSetAttribute(ball_staged_kernel<U, BALL_FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ball_staged_kernel<U, BALL_FILL><<<(unsigned int)chunks, kStagedBlock, smem, s>>>(
                    v, qr, radius, cap_pts, nullptr, out, offsets, idx, dist, fallback, fb_count, nullptr);
                ball_kernel<BALL_FILL><<<grid_list, kBlock, 0, s>>>(v, level, ql, radius, nullptr, offsets, idx, dist, scratch_d2, out);
            }
        } else if (mode == BALL_FUSED) {
            // (coarser levels / no staging: one kernel that carries the minimum-norm solver itself)
            ball_kernel<BALL_FUSED, true><<<grid, kBlock, 0, s>>>(v, level, qr, radius, counts, nullptr, nullptr, nullptr, nullptr, out);
        } else {
            ball_kernel<BALL_FILL><<<grid, kBlock, 0, s>>>(v, level, qr, radius, nullptr, offsets, idx, dist, scratch_d2, out);
        }
        if (d2_pooled) PCT_CUDA(cudaFreeAsync(scratch_d2, s));
    }
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

}  // namespace pct
