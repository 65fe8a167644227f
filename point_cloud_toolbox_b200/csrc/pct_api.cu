// extern "C" surface of libpct_b200.so (include/pct_b200.h): argument checks,
// error strings, launches.  No C++ exception crosses this file.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "pct_internal.h"

namespace pct {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    std::snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return PCT_ERR_NO_DEVICE;
    return PCT_ERR_CUDA;
}

__global__ void permutation_kernel(const Pt* __restrict__ pts, long long n, int32_t* __restrict__ perm) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (int32_t)pts[i].idx;
}

static int check_range(const pct_index* ix, int64_t q_begin, int64_t q_end, int layout, const char* who) {
    if (!ix) { set_error(std::string(who) + ": index is NULL"); return PCT_ERR_INVALID_ARGUMENT; }
    if (q_begin < 0 || q_end < q_begin || q_end > ix->view.n) { set_error(std::string(who) + ": bad query range"); return PCT_ERR_INVALID_ARGUMENT; }
    if (layout != PCT_LAYOUT_ORIGINAL && layout != PCT_LAYOUT_SLICE) { set_error(std::string(who) + ": bad layout"); return PCT_ERR_INVALID_ARGUMENT; }
    return PCT_OK;
}

static int check_k(const pct_index* ix, int k, const char* who) {
    if (k < 1 || k > PCT_MAX_K) { set_error(std::string(who) + ": k must be in [1, 128]"); return PCT_ERR_INVALID_ARGUMENT; }
    if ((long long)k + 1 > ix->view.n) { set_error(std::string(who) + ": k + 1 exceeds the number of points"); return PCT_ERR_K_TOO_LARGE; }
    return PCT_OK;
}

}  // namespace pct

using namespace pct;

extern "C" {

int pct_version(void) { return PCT_VERSION; }

const char* pct_last_error(void) { return g_last_error.c_str(); }

int pct_index_permutation(const pct_index* ix, int32_t* perm, void* stream) {
    PCT_REQUIRE(ix && perm, "pct_index_permutation: NULL argument");
    const long long n = ix->view.n;
    permutation_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ix->pts, n, perm);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int pct_index_last_stats(const pct_index* ix, void* stream, pct_query_stats* stats) {
    PCT_REQUIRE(ix && stats, "pct_index_last_stats: NULL argument");
    unsigned int h[7];
    PCT_CUDA(cudaMemcpyAsync(h, ix->stats, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PCT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    stats->level1_retries = h[0];
    stats->exact_path = h[1];
    stats->kernel_launches = h[2];
    stats->queries = h[3];
    stats->unstaged = h[4];
    stats->unresolved = h[5];
    stats->rank_deficient = h[6];
    return PCT_OK;
}

int pct_knn(const pct_index* ix, int64_t q_begin, int64_t q_end, int k, int32_t* idx, float* dist, int layout, void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_knn");
    if (rc) return rc;
    rc = check_k(ix, k, "pct_knn");
    if (rc) return rc;
    FitOutputs none{nullptr, nullptr, nullptr, nullptr, nullptr};
    return launch_knn(ix, q_begin, q_end, k, false, idx, dist, none, layout, (cudaStream_t)stream);
}

int pct_knn_points(const pct_index* ix, const float* xyz, int stride, const int32_t* query_ids, int64_t nq, int k,
                   int32_t* idx, float* dist, void* stream) {
    PCT_REQUIRE(ix && xyz && (stride == 3 || stride == 4) && nq >= 0 && (query_ids || nq == 0), "pct_knn_points: bad argument");
    const int rc = check_k(ix, k, "pct_knn_points");
    if (rc) return rc;
    return launch_knn_points(ix, xyz, stride, query_ids, nq, k, idx, dist, nullptr, (cudaStream_t)stream);
}

int pct_knn_query(const pct_index* ix, const float* queries, int64_t nq, int k, int32_t* idx, double* dist, void* stream) {
    PCT_REQUIRE(ix && nq >= 0 && (queries || nq == 0) && (idx || nq == 0) && (dist || nq == 0), "pct_knn_query: bad argument");
    PCT_REQUIRE(ix->view.slab_axis < 0, "pct_knn_query: not available on a slab index");
    PCT_REQUIRE(k >= 1 && k <= PCT_MAX_K, "pct_knn_query: k must be in [1, 128]");
    if ((long long)k > ix->view.n) { set_error("pct_knn_query: k exceeds the number of points"); return PCT_ERR_K_TOO_LARGE; }
    return launch_knn_query(ix, queries, nq, k, idx, dist, (cudaStream_t)stream);
}

int pct_curvature_points_records(const pct_index* ix, const float* xyz, int stride, const int32_t* query_ids, int64_t nq,
                                 int k, float* records, void* stream) {
    PCT_REQUIRE(ix && xyz && (stride == 3 || stride == 4) && nq >= 0 && (query_ids || nq == 0) &&
                    (records || nq == 0) && (reinterpret_cast<uintptr_t>(records) & 31) == 0,
                "pct_curvature_points_records: bad argument (records must be 32-byte aligned)");
    const int rc = check_k(ix, k, "pct_curvature_points_records");
    if (rc) return rc;
    return launch_knn_points(ix, xyz, stride, query_ids, nq, k, nullptr, nullptr, records, (cudaStream_t)stream);
}

int pct_index_set_slab(pct_index* ix, int axis, float complete_lo, float complete_hi, float own_lo, float own_hi,
                       const int32_t* row_map) {
    PCT_REQUIRE(ix && axis >= -1 && axis <= 2, "pct_index_set_slab: bad argument");
    PCT_REQUIRE(axis < 0 || (complete_lo <= own_lo && own_lo <= own_hi && own_hi <= complete_hi),
                "pct_index_set_slab: need complete_lo <= own_lo <= own_hi <= complete_hi");
    ix->view.slab_axis = axis;
    ix->view.complete_lo = complete_lo; ix->view.complete_hi = complete_hi;
    ix->view.own_lo = own_lo; ix->view.own_hi = own_hi;
    ix->row_map = axis >= 0 ? row_map : nullptr;
    return PCT_OK;
}

}  // extern "C"
namespace {
__global__ void store_peers_kernel(const pct::PeerRoute route, pct::PeerRoute* dst) { *dst = route; }
}  // namespace
extern "C" {

int pct_index_set_peers(pct_index* ix, int world, const int64_t* begins, void* const* peer_rows, const int32_t* row_ids) {
    PCT_REQUIRE(ix != nullptr, "pct_index_set_peers: index is NULL");
    cudaStream_t s = ix->stream;
    if (world <= 0) {  // removes the routing
        if (ix->peers) PCT_CUDA(cudaFreeAsync(ix->peers, s));
        ix->peers = nullptr;
        return PCT_OK;
    }
    PCT_REQUIRE(world <= pct::kPeerMax && begins && peer_rows && row_ids, "pct_index_set_peers: bad argument");
    PCT_REQUIRE(ix->view.slab_axis >= 0 && ix->row_map, "pct_index_set_peers: the index must be a slab with a row map (pct_index_set_slab)");
    pct::PeerRoute h;
    std::memset(&h, 0, sizeof(h));
    for (int r = 0; r < world; ++r) {
        PCT_REQUIRE(begins[r] <= begins[r + 1], "pct_index_set_peers: share bounds must ascend");
        PCT_REQUIRE(peer_rows[r] != nullptr && (reinterpret_cast<uintptr_t>(peer_rows[r]) & 7) == 0 || begins[r] == begins[r + 1],
                    "pct_index_set_peers: peer arrays must be non-NULL and 8-byte aligned");
        h.base[r] = static_cast<float*>(peer_rows[r]);
        h.begin[r] = begins[r];
    }
    h.begin[world] = begins[world];
    h.row_ids = row_ids;
    h.world = world;
    if (!ix->peers) PCT_CUDA(cudaMallocAsync(&ix->peers, sizeof(pct::PeerRoute), s));
    store_peers_kernel<<<1, 1, 0, s>>>(h, ix->peers);  // by value through the parameter space: no host synchronisation
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int pct_curvature_fused_knn(const pct_index* ix, int64_t q_begin, int64_t q_end, int k, float* normals, float* coeffs,
                            float* curv, uint8_t* status, int layout, void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_curvature_fused_knn");
    if (rc) return rc;
    rc = check_k(ix, k, "pct_curvature_fused_knn");
    if (rc) return rc;
    FitOutputs out{normals, coeffs, curv, status, nullptr};
    return launch_knn(ix, q_begin, q_end, k, true, nullptr, nullptr, out, layout, (cudaStream_t)stream);
}

int pct_curvature_fused_knn_records(const pct_index* ix, int64_t q_begin, int64_t q_end, int k, float* records, int layout,
                                    void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_curvature_fused_knn_records");
    if (rc) return rc;
    rc = check_k(ix, k, "pct_curvature_fused_knn_records");
    if (rc) return rc;
    PCT_REQUIRE(records != nullptr && (reinterpret_cast<uintptr_t>(records) & 31) == 0,
                "pct_curvature_fused_knn_records: records must be non-NULL and 32-byte aligned");
    FitOutputs out{nullptr, nullptr, nullptr, nullptr, records};
    if (ix->peers) {
        PCT_REQUIRE(layout == PCT_LAYOUT_ORIGINAL, "pct_curvature_fused_knn_records: peer routing needs PCT_LAYOUT_ORIGINAL");
        out.peers = ix->peers;
    }
    return launch_knn(ix, q_begin, q_end, k, true, nullptr, nullptr, out, layout, (cudaStream_t)stream);
}

int pct_curvature_fused_ball_records(const pct_index* ix, int64_t q_begin, int64_t q_end, double radius, int32_t* counts,
                                     float* records, int layout, void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_curvature_fused_ball_records");
    if (rc) return rc;
    PCT_REQUIRE(radius >= 0.0 && records != nullptr && (reinterpret_cast<uintptr_t>(records) & 31) == 0,
                "pct_curvature_fused_ball_records: bad radius or records not 32-byte aligned");
    FitOutputs out{nullptr, nullptr, nullptr, nullptr, records};
    return launch_ball(ix, q_begin, q_end, radius, 2, counts, nullptr, 0, nullptr, nullptr, out, layout, (cudaStream_t)stream);
}

int pct_ball_count(const pct_index* ix, int64_t q_begin, int64_t q_end, double radius, int32_t* counts, int layout, void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_ball_count");
    if (rc) return rc;
    PCT_REQUIRE(radius >= 0.0 && counts, "pct_ball_count: radius must be >= 0 and counts non-NULL");
    FitOutputs none{nullptr, nullptr, nullptr, nullptr, nullptr};
    return launch_ball(ix, q_begin, q_end, radius, 0, counts, nullptr, 0, nullptr, nullptr, none, layout, (cudaStream_t)stream);
}

int pct_ball_fill(const pct_index* ix, int64_t q_begin, int64_t q_end, double radius, const int64_t* offsets, int64_t nnz,
                  int32_t* idx, float* dist, int layout, void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_ball_fill");
    if (rc) return rc;
    PCT_REQUIRE(radius >= 0.0 && offsets && idx && nnz >= 0, "pct_ball_fill: bad argument");
    FitOutputs none{nullptr, nullptr, nullptr, nullptr, nullptr};
    return launch_ball(ix, q_begin, q_end, radius, 1, nullptr, (const long long*)offsets, nnz, idx, dist, none, layout,
                       (cudaStream_t)stream);
}

int pct_curvature_fused_ball(const pct_index* ix, int64_t q_begin, int64_t q_end, double radius, int32_t* counts,
                             float* normals, float* coeffs, float* curv, uint8_t* status, int layout, void* stream) {
    int rc = check_range(ix, q_begin, q_end, layout, "pct_curvature_fused_ball");
    if (rc) return rc;
    PCT_REQUIRE(radius >= 0.0, "pct_curvature_fused_ball: radius must be >= 0");
    FitOutputs out{normals, coeffs, curv, status, nullptr};
    return launch_ball(ix, q_begin, q_end, radius, 2, counts, nullptr, 0, nullptr, nullptr, out, layout, (cudaStream_t)stream);
}

int pct_fit_from_neighbors(const float* xyz, int64_t n, const int32_t* idx, int64_t nq, int k, const int32_t* query_ids,
                           float* normals, float* coeffs, float* curv, uint8_t* status, void* stream) {
    PCT_REQUIRE(xyz && idx && n >= 1 && nq >= 0 && k >= 1, "pct_fit_from_neighbors: bad argument");
    FitOutputs out{normals, coeffs, curv, status, nullptr};
    return launch_fit_rows(xyz, n, idx, nq, k, query_ids, out, (cudaStream_t)stream);
}

int pct_fit_from_csr(const float* xyz, int64_t n, const int64_t* offsets, const int32_t* idx, int64_t nq,
                     const int32_t* query_ids, float* normals, float* coeffs, float* curv, uint8_t* status, void* stream) {
    PCT_REQUIRE(xyz && offsets && n >= 1 && nq >= 0, "pct_fit_from_csr: bad argument");
    FitOutputs out{normals, coeffs, curv, status, nullptr};
    return launch_fit_csr(xyz, n, (const long long*)offsets, idx, nq, query_ids, out, (cudaStream_t)stream);
}

int pct_plane_rotate(const float* centered, int64_t nq, int k, double* rotated, double* normals, uint8_t* status, void* stream) {
    PCT_REQUIRE(centered && rotated && nq >= 0 && k >= 1, "pct_plane_rotate: bad argument");
    return launch_plane_rotate(centered, nq, k, rotated, normals, status, (cudaStream_t)stream);
}

int pct_quadric_fit(const double* rotated, int64_t nq, int k, float* coeffs, uint8_t* status, void* stream) {
    PCT_REQUIRE(rotated && coeffs && nq >= 0 && k >= 1, "pct_quadric_fit: bad argument");
    return launch_quadric_fit(rotated, nq, k, coeffs, status, (cudaStream_t)stream);
}

int pct_quadric_curvature(const float* coeffs, int64_t nq, float* curv, void* stream) {
    PCT_REQUIRE(coeffs && curv && nq >= 0, "pct_quadric_curvature: bad argument");
    return launch_quadric_curvature(coeffs, nq, curv, (cudaStream_t)stream);
}

int pct_implicit_quadric_fit(const float* xyz, int64_t n, const int32_t* idx, int64_t nq, int k, const int32_t* query_ids,
                              const float* centered, double* coeffs, void* stream) {
    PCT_REQUIRE(coeffs && nq >= 0 && k >= 1 && ((xyz && idx && n >= 1) || centered), "pct_implicit_quadric_fit: bad argument");
    return launch_implicit_fit(xyz, n, idx, nq, k, query_ids, centered, coeffs, (cudaStream_t)stream);
}

int pct_implicit_quadric_curvature(const double* coeffs, int64_t nq, double* curv, void* stream) {
    PCT_REQUIRE(coeffs && curv && nq >= 0, "pct_implicit_quadric_curvature: bad argument");
    return launch_implicit_curvature(coeffs, nq, curv, (cudaStream_t)stream);
}

int pct_pca_from_neighbors(const float* xyz, int64_t n, const int32_t* idx, int64_t nq, int k, int include_self,
                           const int32_t* query_ids, double* values, double* directions, void* stream) {
    PCT_REQUIRE(xyz && idx && values && n >= 1 && nq >= 0 && k >= 1, "pct_pca_from_neighbors: bad argument");
    return launch_pca_rows(xyz, idx, nq, k, include_self, query_ids, values, directions, (cudaStream_t)stream);
}

int pct_curvature_knn_host(const float* xyz_host, int64_t n, int k, float* K_host, float* H_host) {
    PCT_REQUIRE(xyz_host && K_host && H_host && n >= 1, "pct_curvature_knn_host: bad argument");
    cudaStream_t s;
    PCT_CUDA(cudaStreamCreate(&s));
    float* d_xyz = nullptr;
    float* d_curv = nullptr;
    float* h_curv = nullptr;
    pct_index* ix = nullptr;
    int rc = PCT_OK;
    auto cleanup = [&]() {
        if (ix) pct_index_destroy(ix);
        if (d_xyz) cudaFree(d_xyz);
        if (d_curv) cudaFree(d_curv);
        if (h_curv) cudaFreeHost(h_curv);
        release_scratch_of(s);  // the build's and the query's temporaries live in an arena keyed by this stream
        cudaStreamDestroy(s);
    };
#define PCT_TRY(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call, __FILE__, __LINE__); cleanup(); return rc; } \
    } while (0)
    PCT_TRY(cudaMalloc(&d_xyz, sizeof(float) * 3 * (size_t)n));
    PCT_TRY(cudaMalloc(&d_curv, sizeof(float) * 5 * (size_t)n));
    PCT_TRY(cudaMallocHost(&h_curv, sizeof(float) * 5 * (size_t)n));
    rc = pct_upload(d_xyz, xyz_host, (int64_t)(sizeof(float) * 3 * (size_t)n), s);  // staged when the source is pageable
    if (rc != PCT_OK) { cleanup(); return rc; }
    rc = pct_index_build(d_xyz, n, 3, 0.f, k, s, &ix);
    if (rc == PCT_OK) rc = pct_curvature_fused_knn(ix, 0, n, k, nullptr, nullptr, d_curv, nullptr, PCT_LAYOUT_ORIGINAL, s);
    if (rc != PCT_OK) { cleanup(); return rc; }
    PCT_TRY(cudaMemcpyAsync(h_curv, d_curv, sizeof(float) * 5 * (size_t)n, cudaMemcpyDeviceToHost, s));
    PCT_TRY(cudaStreamSynchronize(s));
#undef PCT_TRY
    for (int64_t i = 0; i < n; ++i) { K_host[i] = h_curv[5 * i]; H_host[i] = h_curv[5 * i + 1]; }
    cleanup();
    return PCT_OK;
}

}  // extern "C"
