// Energy integration over a triangle mesh whose vertices carry the path's K and H: the consumer
// right after the curvature path (/root/reference/utils.py:702-765).  The reference recomputes the
// three sums inside its per-triangle loop (O(T^2), 870 s of its 930 s profile); here it is one
// streaming pass: thread per triangle, fp64 partial sums per block, a second one-block kernel adds the
// partials in a fixed order, so the result does not depend on scheduling.
// Algorithmic bytes per triangle: 12 (indices); the nine coordinate and six curvature gathers hit L2.
#include "pct_energy.cuh"
#include "pct_internal.h"

namespace pct {
namespace {

constexpr int kEnergyBlock = 256;
constexpr int kTrisPerThread = 2;   // triangles per thread and trip
constexpr int kEnergyResident = 4;  // blocks per SM the kernel is compiled for (64 registers)

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

// 32-bit index arithmetic (n_vertices < 2^31 because the indices are int32): one wide multiply-add per address
__device__ __forceinline__ bool load_corner(const float* __restrict__ xyz, const float* __restrict__ K,
                                            const float* __restrict__ H, uint32_t n_vertices, int32_t i, Corner& c) {
    const uint32_t u = (uint32_t)i + (i < 0 ? n_vertices : 0u);  // numpy indexing: negative indices count from the end
    const bool ok = u < n_vertices;                              // (a negative index below -n wraps to a large value)
    const uint32_t j = ok ? u : 0u;
    const float* p = xyz + (size_t)j * 3;
    c.x = __ldg(p); c.y = __ldg(p + 1); c.z = __ldg(p + 2);
    c.k = K ? __ldg(K + j) : 0.f;
    c.h = H ? __ldg(H + j) : 0.f;
    return ok;
}

// kTrisPerThread triangles per thread and trip: their index loads, then all their gathers, are independent
__global__ void __launch_bounds__(kEnergyBlock, kEnergyResident)
energy_partials_kernel(const float* __restrict__ xyz, long long n_vertices, const int32_t* __restrict__ tri,
                       long long n_triangles, const float* __restrict__ K, const float* __restrict__ H,
                       double* __restrict__ partials) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // bending, stretching, area, triangles with an index out of range
    const long long groups = (n_triangles + kTrisPerThread - 1) / kTrisPerThread;
    const uint32_t nv = (uint32_t)n_vertices;
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
            const bool a = load_corner(xyz, K, H, nv, id[3 * u], c[3 * u]);
            const bool b = load_corner(xyz, K, H, nv, id[3 * u + 1], c[3 * u + 1]);
            const bool d = load_corner(xyz, K, H, nv, id[3 * u + 2], c[3 * u + 2]);
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
    // Two triangles per thread, four blocks per SM.  Measured on a B200 (scripts/energy_probe.py, cold L2, 32 M
    // triangles of a grid mesh / 18 M triangles with shuffled vertex numbering): 1 x 8 blocks 0.35 / 1.11 ms,
    // 2 x 4 blocks 0.30 / 0.79 ms, 4 x 3 blocks 0.38 / 0.98 ms, 4 x 2 blocks 0.37 / 0.74 ms.
    const long long per_block = (long long)kEnergyBlock * kTrisPerThread;
    const long long want = (n_triangles + per_block - 1) / per_block;
    const int blocks = (int)(want < (long long)sms * kEnergyResident ? want : (long long)sms * kEnergyResident);
    ScratchSession scratch(s, (size_t)blocks * 4 * sizeof(double) + 256);
    double* partials = static_cast<double*>(scratch.take((size_t)blocks * 4 * sizeof(double)));
    const bool own = partials == nullptr;
    if (own) PCT_CUDA(cudaMallocAsync(&partials, (size_t)blocks * 4 * sizeof(double), s));
    energy_partials_kernel<<<blocks, kEnergyBlock, 0, s>>>(xyz, n_vertices, tri, n_triangles, K, H, partials);
    energy_final_kernel<<<1, kEnergyBlock, 0, s>>>(partials, blocks, out);
    const cudaError_t e = cudaGetLastError();
    if (own) cudaFreeAsync(partials, s);
    PCT_CUDA(e);
    return PCT_OK;
}

}  // namespace pct

extern "C" int pct_mesh_energies(const float* vertices, int64_t n_vertices, const int32_t* triangles, int64_t n_triangles,
                                 const float* gaussian, const float* mean, double* out, void* stream) {
    PCT_REQUIRE(out && n_vertices >= 0 && n_vertices < (1ll << 31) && n_triangles >= 0 &&
                    ((vertices && triangles) || n_triangles == 0),
                "pct_mesh_energies: bad argument");
    return pct::launch_mesh_energies(vertices, n_vertices, triangles, n_triangles, gaussian, mean, out, (cudaStream_t)stream);
}
