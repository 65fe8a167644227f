// Thread-per-query kNN kernel: histogram selection, exact neighbour set in shared
// memory, fused fp64 fit.  One kernel for every k (the selection is O(candidates)).
// See pct_query.cu for the overview of the query path.
#pragma once

#include <algorithm>

#include "pct_internal.h"

namespace pct {

constexpr int kBlock = 128;

struct QueryRange {
    long long q_begin, q_end;
    const uint32_t* list;       // when non-null: sorted positions to process ...
    const unsigned int* count;  // ... and how many of them
    int layout;
};

struct Queues {
    uint32_t* retry;   // queries to redo one level coarser (may be null: go straight to exact)
    uint32_t* exact;   // queries for the exact kernel
    unsigned int* counters;  // [0] = retry count, [1] = exact count
};

__device__ __forceinline__ long long out_row(const QueryRange& qr, uint32_t i, uint32_t orig) {
    return qr.layout == PCT_LAYOUT_ORIGINAL ? (long long)orig : (long long)i - qr.q_begin;
}

// shared memory of one block: [54][kBlock] cell runs, [cap][kBlock] neighbour list, [kBlock][68 B] histograms
__host__ __device__ inline size_t fast_smem_bytes(int cap) {
    return sizeof(uint32_t) * (size_t)(54 + cap) * kBlock + (size_t)kHistRowBytes * kBlock;
}

template <bool FUSED>
__global__ void __launch_bounds__(kBlock, 4)
knn_fast_kernel(const IndexView ix, const int level, const QueryRange qr, const int k, const int cap,
                int32_t* __restrict__ out_idx, float* __restrict__ out_dist, const FitOutputs out, const Queues qu) {
    extern __shared__ uint32_t smem_words[];
    SelectScratch sc;
    sc.runs = smem_words + threadIdx.x;
    sc.list = smem_words + 54 * kBlock + threadIdx.x;
    sc.hist = reinterpret_cast<uint8_t*>(smem_words + (54 + cap) * kBlock) + kHistRowBytes * threadIdx.x;
    sc.stride = kBlock;
    sc.cap = cap;
    long long total = qr.q_end - qr.q_begin;
    if (qr.list) total = (long long)*qr.count;
    for (long long base = (long long)blockIdx.x * kBlock; base < total; base += (long long)gridDim.x * kBlock) {
        const long long t = base + threadIdx.x;
        if (t >= total) continue;
        const uint32_t i = qr.list ? qr.list[t] : (uint32_t)(qr.q_begin + t);
        const Pt q = load_pt(ix.pts + i);
        uint32_t* list = sc.list;
        uint32_t first = 0, last = 0;
        double d2_last = 0.0;
        const int rc = knn_select(ix, level, i, q, k, sc, first, last, d2_last);
        if (rc != SEL_OK) {
            if (rc == SEL_RETRY_COARSER && qu.retry && level + 1 < ix.num_levels) {
                qu.retry[atomicAdd(&qu.counters[0], 1u)] = i;
            } else {
                qu.exact[atomicAdd(&qu.counters[1], 1u)] = i;
            }
            continue;
        }
        const long long row = out_row(qr, i, q.idx);
        if (FUSED) {
            ListNeighbourhood nb;
            nb.ix = &ix; nb.list = list; nb.stride = kBlock; nb.count = k; nb.q = q; nb.first = first; nb.last = last;
            FitResult r;
            r.status = 0;
            fit_neighbourhood(nb, r);
            store_fit(out, row, r);
        } else {
            // ordered rows: successive minima of (d2, index) over the k members
            double pd = -1.0;
            uint32_t pi = 0;
            for (int m = 0; m < k; ++m) {
                double bd = 1.0e300;
                uint32_t bi = 0xffffffffu;
                for (int c = 0; c < k; ++c) {
                    const Pt p = load_pt(ix.pts + list[c * kBlock]);
                    const double d = dist2_f64(q.x, q.y, q.z, p.x, p.y, p.z);
                    if (key_less(pd, pi, d, p.idx) && key_less(d, p.idx, bd, bi)) { bd = d; bi = p.idx; }
                }
                if (out_idx) out_idx[row * k + m] = (int32_t)bi;
                if (out_dist) out_dist[row * k + m] = (float)sqrt(bd);
                pd = bd; pi = bi;
            }
        }
    }
}

struct FastLaunch {
    const pct_index* ix;
    QueryRange qr;
    int k, cap;
    bool fused;
    int32_t* idx;
    float* dist;
    FitOutputs out;
    uint32_t* retry1;
    uint32_t* exactq;
    unsigned int* counters;
    cudaStream_t s;
};

// level 0 over the whole range, then level 1 over whatever level 0 queued
template <bool FUSED>
static int launch_fast_impl(const FastLaunch& a, unsigned int* launches) {
    const IndexView& v = a.ix->view;
    const long long nq = a.qr.q_end - a.qr.q_begin;
    const size_t smem = fast_smem_bytes(a.cap);
    const int grid_all = (int)std::min<long long>((nq + kBlock - 1) / kBlock, (long long)a.ix->sm_count * 64);
    const int grid_retry = (int)std::min<long long>((nq + kBlock - 1) / kBlock, (long long)a.ix->sm_count * 8);
    auto kern = knn_fast_kernel<FUSED>;
    PCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Queues q0{v.num_levels > 1 ? a.retry1 : nullptr, a.exactq, a.counters};
    kern<<<grid_all, kBlock, smem, a.s>>>(v, 0, a.qr, a.k, a.cap, a.idx, a.dist, a.out, q0);
    ++*launches;
    if (v.num_levels > 1) {
        QueryRange q1 = a.qr;
        q1.list = a.retry1;
        q1.count = a.counters;
        Queues qq{nullptr, a.exactq, a.counters};
        kern<<<grid_retry, kBlock, smem, a.s>>>(v, 1, q1, a.k, a.cap, a.idx, a.dist, a.out, qq);
        ++*launches;
    }
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

static int launch_fast(const FastLaunch& a, unsigned int* launches) {
    return a.fused ? launch_fast_impl<true>(a, launches) : launch_fast_impl<false>(a, launches);
}

}  // namespace pct
