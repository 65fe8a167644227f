// Host-side internals shared by the translation units of libpct_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/pct_b200.h"
#include "pct_dispatch.h"
#include "pct_grid.cuh"

namespace pct { struct PeerRoute; }

struct pct_index {
    pct::IndexView view;       // device pointers inside
    pct::Pt* pts = nullptr;    // N sorted records
    pct::HashSlot* table_mem = nullptr;  // all level tables, one allocation
    const int32_t* row_map = nullptr;    // caller-owned (pct_index_set_slab): output row per original index
    pct::PeerRoute* peers = nullptr;  // device copy of the peer routing (pct_index_set_peers), owned by the index
    unsigned int* stats = nullptr;       // device: [retries, exact, launches, queries, unstaged, -, -, -]
    pct_index_info info{};
    int device = 0;
    long long cells_level[pct::kMaxLevels] = {};  // occupied cells per level
    float est_dimension = 2.f;                    // intrinsic dimension seen by the density pilot
    int sm_count = 148;
    int smem_per_sm = 233472;           // shared memory of one SM
    int smem_per_block_optin = 232448;  // largest dynamic allocation of one block
    cudaStream_t stream = nullptr;  // stream the index was built on; its memory is freed in that stream's order
};

namespace pct {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PCT_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return pct::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define PCT_REQUIRE(cond, msg)                 \
    do {                                       \
        if (!(cond)) {                         \
            pct::set_error(msg);               \
            return PCT_ERR_INVALID_ARGUMENT;   \
        }                                      \
    } while (0)

// Bump allocation out of the calling stream's scratch arena for the duration of one call
// (pct_scratch.cu).  take() returns nullptr when the arena is exhausted or could not be allocated;
// callers then fall back to cudaMallocAsync.
class ScratchSession {
 public:
    ScratchSession(cudaStream_t s, size_t bytes);
    void* take(size_t bytes);

 private:
    char* p_ = nullptr;
    size_t left_ = 0;
};
void release_scratch();
void release_scratch_of(cudaStream_t s);  // one stream of the current device; call before destroying a library-owned stream
void release_upload_stage();  // pct_transfer.cu

// Multi-GPU return fused into the fit's store (pct_index_set_peers): the K, H of an answered query go straight to the
// rank that holds the query's share of the cloud -- an 8-byte store into that rank's result array through NVLink peer
// memory -- instead of a local write, a column extraction, an all-to-all and a scatter after the kernel.
constexpr int kPeerMax = 32;
struct PeerRoute {
    float* base[kPeerMax];            // (rows of rank r) x {K, H}: rank r's result array, mapped into this process
    long long begin[kPeerMax + 1];    // rank r holds the original indices begin[r] .. begin[r + 1] - 1
    const int32_t* row_ids;           // original (whole-cloud) index of output row r of this index (owned points, row order)
    int world;
};

// per-row output pointers of the fit (any may be null)
struct FitOutputs {
    float* normals;
    float* coeffs;
    float* curv;
    uint8_t* status;
    float* records;  // rows x 8 floats {nx, ny, nz, K, H, k1, k2, status bits}: one aligned 32-byte sector per point
    const PeerRoute* peers = nullptr;  // device pointer; when set, K and H of a row also go to the rank that owns its point
};

__device__ __forceinline__ void store_fit(const FitOutputs& o, long long row, const FitResult& r) {
    if (o.normals) {
        float* p = o.normals + 3 * row;
        p[0] = r.normal[0]; p[1] = r.normal[1]; p[2] = r.normal[2];
    }
    if (o.coeffs) {
        float* p = o.coeffs + 6 * row;
#pragma unroll
        for (int c = 0; c < 6; ++c) p[c] = r.coeffs[c];
    }
    if (o.curv) {
        float* p = o.curv + 5 * row;
#pragma unroll
        for (int c = 0; c < 5; ++c) p[c] = r.curv[c];
    }
    if (o.status) o.status[row] = (uint8_t)r.status;
    if (o.records) {
        // two 16-byte stores that together fill exactly one 32-byte sector: a scattered
        // (original-order) row costs 32 B of DRAM traffic instead of three partial sectors
        float4* p = reinterpret_cast<float4*>(o.records + 8 * row);
        p[0] = make_float4(r.normal[0], r.normal[1], r.normal[2], r.curv[0]);
        p[1] = make_float4(r.curv[1], r.curv[2], r.curv[3], __uint_as_float(r.status));
    }
    if (o.peers) {
        const PeerRoute& pr = *o.peers;
        const long long g = pr.row_ids[row];
        int owner = 0;
        for (int t = 1; t < pr.world; ++t) owner += g >= pr.begin[t] ? 1 : 0;
        reinterpret_cast<float2*>(pr.base[owner])[g - pr.begin[owner]] = make_float2(r.curv[0], r.curv[1]);
    }
}

// query-side launchers (pct_query.cu)
int launch_knn(const pct_index* ix, long long q_begin, long long q_end, int k, bool fused,
               int32_t* idx, float* dist, FitOutputs out, int layout, cudaStream_t s);
int launch_knn_points(const pct_index* ix, const float* xyz, int stride, const int32_t* ids, long long nq, int k,
                      int32_t* idx, float* dist, float* records, cudaStream_t s);
int launch_knn_query(const pct_index* ix, const float* queries, long long nq, int k, int32_t* idx, double* dist, cudaStream_t s);
int launch_ball(const pct_index* ix, long long q_begin, long long q_end, double radius, int mode,
                int32_t* counts, const long long* offsets, long long nnz, int32_t* idx, float* dist,
                FitOutputs out, int layout, cudaStream_t s);

// fit-side launchers (pct_fit.cu)
int launch_fit_rows(const float* xyz, long long n, const int32_t* idx, long long nq, int k,
                    const int32_t* qids, FitOutputs out, cudaStream_t s);
int launch_fit_csr(const float* xyz, long long n, const long long* offsets, const int32_t* idx, long long nq,
                   const int32_t* qids, FitOutputs out, cudaStream_t s);
int launch_plane_rotate(const float* centered, long long nq, int k, double* rotated, double* normals,
                        uint8_t* status, cudaStream_t s);
int launch_quadric_fit(const double* rotated, long long nq, int k, float* coeffs, uint8_t* status, cudaStream_t s);
int launch_quadric_curvature(const float* coeffs, long long nq, float* curv, cudaStream_t s);
int launch_implicit_fit(const float* xyz, long long n, const int32_t* idx, long long nq, int k, const int32_t* qids,
                        const float* centered, double* coeffs, cudaStream_t s);
int launch_implicit_curvature(const double* coeffs, long long nq, double* out, cudaStream_t s);
int launch_pca_rows(const float* xyz, const int32_t* idx, long long nq, int k, int include_self, const int32_t* qids,
                    double* values, double* directions, cudaStream_t s);

}  // namespace pct
