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

__global__ void __launch_bounds__(kEnergyBlock)
energy_partials_kernel(const float* __restrict__ xyz, long long n_vertices, const int32_t* __restrict__ tri,
                       long long n_triangles, const float* __restrict__ K, const float* __restrict__ H,
                       double* __restrict__ partials) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // bending, stretching, area, triangles with an index out of range
    for (long long t = (long long)blockIdx.x * kEnergyBlock + threadIdx.x; t < n_triangles;
         t += (long long)gridDim.x * kEnergyBlock) {
        long long a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        // numpy indexing: negative indices count from the end
        a += a < 0 ? n_vertices : 0; b += b < 0 ? n_vertices : 0; c += c < 0 ? n_vertices : 0;
        if (a < 0 || b < 0 || c < 0 || a >= n_vertices || b >= n_vertices || c >= n_vertices) { acc[3] += 1.0; continue; }
        const TriangleTerms tt = triangle_terms(xyz + 3 * a, xyz + 3 * b, xyz + 3 * c, K ? K[a] : 0.f, K ? K[b] : 0.f,
                                                K ? K[c] : 0.f, H ? H[a] : 0.f, H ? H[b] : 0.f, H ? H[c] : 0.f);
        acc[0] += tt.bending; acc[1] += tt.stretching; acc[2] += tt.area;
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
    // persistent grid: 8 blocks of 256 threads per SM (2048 resident threads), fewer for small meshes
    const long long want = (n_triangles + kEnergyBlock - 1) / kEnergyBlock;
    const int blocks = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
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
    PCT_REQUIRE(out && n_vertices >= 0 && n_triangles >= 0 && ((vertices && triangles) || n_triangles == 0),
                "pct_mesh_energies: bad argument");
    return pct::launch_mesh_energies(vertices, n_vertices, triangles, n_triangles, gaussian, mean, out, (cudaStream_t)stream);
}
