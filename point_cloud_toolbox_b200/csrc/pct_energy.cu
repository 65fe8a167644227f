// Energy integration over a triangle mesh whose vertices carry the path's K and H: the consumer
// right after the curvature path (/root/reference/utils.py:702-765).  The reference recomputes the
// three sums inside its per-triangle loop (O(T^2), 870 s of its 930 s profile); here it is one
// streaming pass: thread per triangle, fp64 partial sums per block, a second one-block kernel adds the
// partials in a fixed order, so the result does not depend on scheduling.
// Algorithmic bytes per triangle: 12 (indices); the nine coordinate and six curvature gathers hit L2.
#include <cstdlib>

#include "pct_energy.cuh"
#include "pct_internal.h"

namespace pct {
namespace {

constexpr int kEnergyBlock = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// sums[0..3] of the block's threads -> out[4] by thread 0
__device__ __forceinline__ void block_sum4(double v[4], double* out) {
    __shared__ double part[4][kEnergyBlock / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double s = warp_sum(v[c]);
        if (lane == 0) part[c][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kEnergyBlock / 32; ++i) s += part[threadIdx.x][i];
        out[threadIdx.x] = s;
    }
}

// one triangle: gathers first (all 15 loads of a triangle are independent), then the arithmetic
struct Corner {
    float x, y, z, k, h;
};

__device__ __forceinline__ bool load_corner(const float* __restrict__ xyz, const float* __restrict__ K,
                                            const float* __restrict__ H, long long n_vertices, long long i, Corner& c) {
    i += i < 0 ? n_vertices : 0;  // numpy indexing: negative indices count from the end
    const bool ok = i >= 0 && i < n_vertices;
    const long long j = ok ? i : 0;
    c.x = __ldg(xyz + 3 * j); c.y = __ldg(xyz + 3 * j + 1); c.z = __ldg(xyz + 3 * j + 2);
    c.k = K ? __ldg(K + j) : 0.f;
    c.h = H ? __ldg(H + j) : 0.f;
    return ok;
}

// kTrisPerThread triangles per thread and trip: their index loads, then all their gathers, are independent
template <int kTrisPerThread, int kMinBlocks>
__global__ void __launch_bounds__(kEnergyBlock, kMinBlocks)
energy_partials_kernel(const float* __restrict__ xyz, long long n_vertices, const int32_t* __restrict__ tri,
                       long long n_triangles, const float* __restrict__ K, const float* __restrict__ H,
                       double* __restrict__ partials) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // bending, stretching, area, triangles with an index out of range
    const long long groups = (n_triangles + kTrisPerThread - 1) / kTrisPerThread;
    const bool aligned = (reinterpret_cast<uintptr_t>(tri) & 15) == 0;
    for (long long g = (long long)blockIdx.x * kEnergyBlock + threadIdx.x; g < groups;
         g += (long long)gridDim.x * kEnergyBlock) {
        const long long t0 = g * kTrisPerThread;
        int32_t id[3 * kTrisPerThread];
        if (kTrisPerThread % 4 == 0 && aligned && t0 + kTrisPerThread <= n_triangles) {
            const int4* p = reinterpret_cast<const int4*>(tri + 3 * t0);
#pragma unroll
            for (int q = 0; q < 3 * kTrisPerThread / 4; ++q) {
                const int4 v = __ldg(p + q);
                id[4 * q] = v.x; id[4 * q + 1] = v.y; id[4 * q + 2] = v.z; id[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 3 * kTrisPerThread; ++q)
                id[q] = 3 * t0 + q < 3 * n_triangles ? __ldg(tri + 3 * t0 + q) : 0;
        }
        Corner c[3 * kTrisPerThread];
        bool ok[kTrisPerThread];
#pragma unroll
        for (int u = 0; u < kTrisPerThread; ++u) {
            const bool a = load_corner(xyz, K, H, n_vertices, id[3 * u], c[3 * u]);
            const bool b = load_corner(xyz, K, H, n_vertices, id[3 * u + 1], c[3 * u + 1]);
            const bool d = load_corner(xyz, K, H, n_vertices, id[3 * u + 2], c[3 * u + 2]);
            ok[u] = a && b && d;
        }
#pragma unroll
        for (int u = 0; u < kTrisPerThread; ++u) {
            if (t0 + u >= n_triangles) break;
            if (!ok[u]) { acc[3] += 1.0; continue; }
            const Corner &a = c[3 * u], &b = c[3 * u + 1], &d = c[3 * u + 2];
            const TriangleTerms tt = triangle_terms(&a.x, &b.x, &d.x, a.k, b.k, d.k, a.h, b.h, d.h);
            acc[0] += tt.bending; acc[1] += tt.stretching; acc[2] += tt.area;
        }
    }
    block_sum4(acc, partials + 4 * (long long)blockIdx.x);
}

__global__ void __launch_bounds__(kEnergyBlock)
energy_final_kernel(const double* __restrict__ partials, int blocks, double* __restrict__ out) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < blocks; b += kEnergyBlock)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += partials[4 * b + c];
    block_sum4(acc, out);
}

}  // namespace

int launch_mesh_energies(const float* xyz, long long n_vertices, const int32_t* tri, long long n_triangles,
                         const float* K, const float* H, double* out, cudaStream_t s) {
    if (n_triangles == 0) {
        PCT_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(double), s));
        return PCT_OK;
    }
    int dev = 0, sms = 148;
    PCT_CUDA(cudaGetDevice(&dev));
    PCT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // persistent grid: up to 8 blocks of 256 threads per SM, four triangles per thread and trip
    static const int variant = [] { const char* e = getenv("PCT_ENERGY_VARIANT"); return e ? atoi(e) : 2; }();
    const int tpt = variant == 1 ? 1 : variant == 2 ? 2 : 4;
    const int resident = variant == 1 ? 8 : variant == 2 ? 4 : variant == 3 ? 3 : 2;
    const long long per_block = (long long)kEnergyBlock * tpt;
    const long long want = (n_triangles + per_block - 1) / per_block;
    const int blocks = (int)(want < (long long)sms * resident ? want : (long long)sms * resident);
    ScratchSession scratch(s, (size_t)blocks * 4 * sizeof(double) + 256);
    double* partials = static_cast<double*>(scratch.take((size_t)blocks * 4 * sizeof(double)));
    const bool own = partials == nullptr;
    if (own) PCT_CUDA(cudaMallocAsync(&partials, (size_t)blocks * 4 * sizeof(double), s));
    if (variant == 1) energy_partials_kernel<1, 8><<<blocks, kEnergyBlock, 0, s>>>(xyz, n_vertices, tri, n_triangles, K, H, partials);
    else if (variant == 2) energy_partials_kernel<2, 4><<<blocks, kEnergyBlock, 0, s>>>(xyz, n_vertices, tri, n_triangles, K, H, partials);
    else if (variant == 3) energy_partials_kernel<4, 3><<<blocks, kEnergyBlock, 0, s>>>(xyz, n_vertices, tri, n_triangles, K, H, partials);
    else energy_partials_kernel<4, 2><<<blocks, kEnergyBlock, 0, s>>>(xyz, n_vertices, tri, n_triangles, K, H, partials);
    energy_final_kernel<<<1, kEnergyBlock, 0, s>>>(partials, blocks, out);
    const cudaError_t e = cudaGetLastError();
    if (own) cudaFreeAsync(partials, s);
    PCT_CUDA(e);
    return PCT_OK;
}

}  // namespace pct

extern "C" int pct_mesh_energies(const float* vertices, int64_t n_vertices, const int32_t* triangles, int64_t n_triangles,
                                 const float* gaussian, const float* mean, double* out, void* stream) {
    PCT_REQUIRE(out && n_vertices >= 0 && n_triangles >= 0 && ((vertices && triangles) || n_triangles == 0),
                "pct_mesh_energies: bad argument");
    return pct::launch_mesh_energies(vertices, n_vertices, triangles, n_triangles, gaussian, mean, out, (cudaStream_t)stream);
}
