// Spatial index build: bounding box -> density pilot -> Morton keys -> radix sort
// -> sorted 16-byte records -> per-level hash tables of cell ranges.
//
// Replaces `sp.spatial.cKDTree(np.array(self.points, dtype=np.float32))`
// (/root/reference/pointCloudToolbox.py:74).  All kernels here are HBM-bound
// streaming passes; the sort uses CUB's radix sort (library code, 8-bit digits
// over 3*bits key bits).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pct_internal.h"

namespace pct {

namespace {

constexpr int kThreads = 256;
constexpr int kPilotSample = 1 << 18;
constexpr int kPilotBits = 10;  // 1024^3 virtual grid

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// order-preserving float <-> uint map so that atomicMin / atomicMax work on floats
__host__ __device__ __forceinline__ unsigned int ordered_bits(float f) {
#if defined(__CUDA_ARCH__)
    const unsigned int u = __float_as_uint(f);
#else
    unsigned int u;
    std::memcpy(&u, &f, sizeof(u));
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float from_ordered_bits(unsigned int o) {
    const unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
    std::memcpy(&f, &u, sizeof(f));
    return f;
}

// box[0..2] = min xyz, box[3..5] = max xyz (ordered bits); box[6] |= 1 when a coordinate is NaN/Inf
__global__ void __launch_bounds__(kThreads) bbox_kernel(const float* __restrict__ xyz, long long n, int stride,
                                                         unsigned int* __restrict__ box) {
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* p = xyz + i * stride;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(p + a);
            bad = bad || !(fabsf(v) <= 3.0e38f);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
    __shared__ float sm[kThreads / 32][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min(lo[a]);
        hi[a] = warp_max(hi[a]);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(box + 6, 1u);
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { sm[warp][a] = lo[a]; sm[warp][3 + a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = sm[0][threadIdx.x];
        for (int w = 1; w < kThreads / 32; ++w) v = threadIdx.x < 3 ? fminf(v, sm[w][threadIdx.x]) : fmaxf(v, sm[w][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(box + threadIdx.x, ordered_bits(v));
        else atomicMax(box + threadIdx.x, ordered_bits(v));
    }
}

// ---- density pilot: 30-bit Morton keys of a strided sample on a 1024^3 grid ----
__global__ void pilot_keys_kernel(const float* __restrict__ xyz, long long n, int stride, long long step, int samples,
                                  float ox, float oy, float oz, float inv_cell, uint32_t* __restrict__ keys) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= samples) return;
    const long long i = min((long long)t * step, n - 1);
    const float* p = xyz + i * stride;
    const int lim = (1 << kPilotBits) - 1;
    const int cx = min((int)((__ldg(p) - ox) * inv_cell), lim);
    const int cy = min((int)((__ldg(p + 1) - oy) * inv_cell), lim);
    const int cz = min((int)((__ldg(p + 2) - oz) * inv_cell), lim);
    keys[t] = (uint32_t)morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
}

// hist[l] += #adjacent pairs of sorted keys whose highest differing bit lies in level l
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) level_hist_kernel(const KeyT* __restrict__ keys, long long n,
                                                              unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[kMaxLevels];
    if (threadIdx.x < kMaxLevels) sh[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x + 1; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long x = (unsigned long long)(keys[i] ^ keys[i - 1]);
        if (x) atomicAdd(&sh[(63 - __clzll((long long)x)) / 3], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kMaxLevels && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void __launch_bounds__(kThreads) morton_keys_kernel(const float* __restrict__ xyz, long long n, int stride,
                                                                IndexView ix, unsigned long long* __restrict__ keys,
                                                                uint32_t* __restrict__ vals) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = xyz + i * stride;
    int cx, cy, cz;
    cell_of(ix, __ldg(p), __ldg(p + 1), __ldg(p + 2), cx, cy, cz);
    keys[i] = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kThreads) gather_kernel(const float* __restrict__ xyz, long long n, int stride,
                                                           const uint32_t* __restrict__ vals, Pt* __restrict__ pts) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t v = vals[i];
    const float* p = xyz + (long long)v * stride;
    float4 r;
    r.x = __ldg(p); r.y = __ldg(p + 1); r.z = __ldg(p + 2); r.w = __uint_as_float(v);
    reinterpret_cast<float4*>(pts)[i] = r;
}

struct BuildTables {
    HashSlot* slots[kMaxLevels];
    uint32_t mask[kMaxLevels];
    int num_levels;
};

__device__ __forceinline__ HashSlot* claim_slot(HashSlot* slots, uint32_t mask, unsigned long long key) {
    uint32_t s = hash_key(key) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&slots[s].key, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) return slots + s;
        s = (s + 1) & mask;
    }
}

// position i in [0, n] is a boundary between sorted points i-1 and i: it ends the
// cells of point i-1 and starts the cells of point i at every level where they differ
__global__ void __launch_bounds__(kThreads) fill_tables_kernel(const unsigned long long* __restrict__ keys, long long n,
                                                                BuildTables t) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i > n) return;
    int top;
    if (i == 0 || i == n) {
        top = t.num_levels - 1;
    } else {
        const unsigned long long x = keys[i] ^ keys[i - 1];
        if (!x) return;
        top = (63 - __clzll((long long)x)) / 3;
    }
    for (int L = 0; L <= top; ++L) {
        if (i > 0) claim_slot(t.slots[L], t.mask[L], keys[i - 1] >> (3 * L))->end = (uint32_t)i;
        if (i < n) claim_slot(t.slots[L], t.mask[L], keys[i] >> (3 * L))->start = (uint32_t)i;
    }
}

// a temporary of the build: from the call's scratch session, else from the stream-ordered pool
struct DeviceTemp {
    void* p = nullptr;
    cudaStream_t s;
    ScratchSession* scratch;
    bool pooled = false;
    DeviceTemp(cudaStream_t st, ScratchSession* ss) : s(st), scratch(ss) {}
    cudaError_t alloc(size_t bytes) {
        if (scratch && (p = scratch->take(bytes ? bytes : 16))) return cudaSuccess;
        pooled = true;
        return cudaMallocAsync(&p, bytes ? bytes : 16, s);
    }
    ~DeviceTemp() { if (p && pooled) cudaFreeAsync(p, s); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

// `n_total`: size of the cloud the n rows of `xyz` were drawn from (n itself when xyz is the whole cloud)
static int choose_cell_size(const float* xyz, long long n, int stride, long long n_total, const float lo[3], float extent_max,
                            int k_hint, cudaStream_t s, ScratchSession* ss, float* h_out, float* dim_out) {
    *dim_out = 2.f;
    if (!(extent_max > 0.f)) { *h_out = 1.f; return PCT_OK; }
    const int samples = (int)std::min<long long>(n, kPilotSample);
    const long long step = std::max<long long>(1, n / samples);
    const float cell = extent_max * (1.0f + 1e-6f) / (float)(1 << kPilotBits);
    DeviceTemp keys_a(s, ss), keys_b(s, ss), tmp(s, ss), hist(s, ss);
    PCT_CUDA(keys_a.alloc(sizeof(uint32_t) * samples));
    PCT_CUDA(keys_b.alloc(sizeof(uint32_t) * samples));
    PCT_CUDA(hist.alloc(sizeof(unsigned long long) * kMaxLevels));
    PCT_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned long long) * kMaxLevels, s));
    pilot_keys_kernel<<<(samples + kThreads - 1) / kThreads, kThreads, 0, s>>>(xyz, n, stride, step, samples, lo[0], lo[1], lo[2],
                                                                                 1.0f / cell, keys_a.as<uint32_t>());
    size_t tmp_bytes = 0;
    PCT_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys_a.as<uint32_t>(), keys_b.as<uint32_t>(), samples, 0,
                                            3 * kPilotBits, s));
    PCT_CUDA(tmp.alloc(tmp_bytes));
    PCT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys_a.as<uint32_t>(), keys_b.as<uint32_t>(), samples, 0,
                                            3 * kPilotBits, s));
    level_hist_kernel<uint32_t><<<std::min(1024, (samples + kThreads - 1) / kThreads), kThreads, 0, s>>>(
        keys_b.as<uint32_t>(), samples, hist.as<unsigned long long>());
    unsigned long long h_hist[kMaxLevels];
    PCT_CUDA(cudaMemcpyAsync(h_hist, hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));

    // occupied cells of the sample at level L (cell edge = cell * 2^L)
    double cells[kPilotBits + 1];
    double acc = 1.0;
    for (int L = kPilotBits; L >= 0; --L) {
        if (L < kPilotBits) acc += (double)h_hist[L];
        cells[L] = acc;
    }
    // finest level at which the sample still puts >= 8 points in a cell
    int Ls = kPilotBits;
    for (int L = 0; L <= kPilotBits; ++L)
        if ((double)samples / cells[L] >= 8.0) { Ls = L; break; }
    double dim = 2.0;
    if (Ls + 1 <= kPilotBits && cells[Ls + 1] >= 8.0) dim = std::log2(cells[Ls] / cells[Ls + 1]);
    dim = std::min(3.0, std::max(1.0, dim));
    const double ppc_full = (double)n_total / cells[Ls];
    // points per cell, tuned on B200 with the staged kernel (scripts/ppc_sweep.py): small k wants few
    // level-1 retries, large k cells small enough for the staging buffer
    const double kh = (double)(k_hint > 0 ? k_hint : 20);
    const double target = kh <= 24.0 ? 0.46 * kh : (kh <= 40.0 ? 0.42 * kh : 0.37 * kh);
    double h = (double)cell * std::ldexp(1.0, Ls) * std::pow(target / ppc_full, 1.0 / dim);
    h = std::min(h, 2.0 * (double)extent_max);
    *h_out = (float)h;
    *dim_out = (float)dim;
    return PCT_OK;
}

// min xyz, max xyz of the cloud; PCT_ERR_NONFINITE if any coordinate is NaN / Inf.  Synchronises `s`.
static int bounding_box(const float* xyz, long long n, int stride, int sm_count, cudaStream_t s, ScratchSession* ss,
                        float h_bbox[6]) {
    const int bb_blocks = std::max(1, std::min<int>(sm_count * 8, (int)((n + kThreads - 1) / kThreads)));
    DeviceTemp bbox(s, ss);
    PCT_CUDA(bbox.alloc(sizeof(unsigned int) * 8));
    PCT_CUDA(cudaMemsetAsync(bbox.p, 0xFF, sizeof(unsigned int) * 3, s));
    PCT_CUDA(cudaMemsetAsync(bbox.as<unsigned int>() + 3, 0, sizeof(unsigned int) * 5, s));
    bbox_kernel<<<bb_blocks, kThreads, 0, s>>>(xyz, n, stride, bbox.as<unsigned int>());
    unsigned int h_box[8];
    PCT_CUDA(cudaMemcpyAsync(h_box, bbox.p, sizeof(h_box), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));
    if (h_box[6]) {
        set_error("Non-finite values in input points");
        return PCT_ERR_NONFINITE;
    }
    for (int a = 0; a < 6; ++a) h_bbox[a] = from_ordered_bits(h_box[a]);
    return PCT_OK;
}

static int build_impl(const float* xyz, long long n, int stride, float cell_hint, int k_hint, cudaStream_t s,
                      pct_index* ix) {
    PCT_CUDA(cudaGetDevice(&ix->device));
    ix->stream = s;
    PCT_CUDA(cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, ix->device));
    PCT_CUDA(cudaDeviceGetAttribute(&ix->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ix->device));
    PCT_CUDA(cudaDeviceGetAttribute(&ix->smem_per_block_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ix->device));

    // temporaries come from the stream-ordered pool; keep freed blocks cached across builds
    {
        cudaMemPool_t pool;
        PCT_CUDA(cudaDeviceGetDefaultMemPool(&pool, ix->device));
        unsigned long long keep = ~0ull;
        PCT_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }

    // keys and values twice, radix-sort workspace (about its input again), pilot and small buffers
    ScratchSession scratch(s, (size_t)n * 24 + (size_t)n * 13 + ((size_t)32 << 20));

    // 1. bounding box + finiteness
    float h_bbox[6];
    {
        const int rc = bounding_box(xyz, n, stride, ix->sm_count, s, &scratch, h_bbox);
        if (rc != PCT_OK) return rc;
    }
    const float lo[3] = {h_bbox[0], h_bbox[1], h_bbox[2]};
    const float ext[3] = {h_bbox[3] - h_bbox[0], h_bbox[4] - h_bbox[1], h_bbox[5] - h_bbox[2]};
    const float extent_max = std::max(ext[0], std::max(ext[1], ext[2]));

    // 2. cell size
    float h = cell_hint, est_dim = 2.f;
    if (!(h > 0.f)) {
        const int rc = choose_cell_size(xyz, n, stride, n, lo, extent_max, k_hint, s, &scratch, &h, &est_dim);
        if (rc != PCT_OK) return rc;
    }
    if (!(h > 0.f) || !std::isfinite(h)) h = 1.f;
    const float h_min = extent_max / (float)((1 << 20) - 2);  // at most 2^20 cells per axis
    if (h < h_min) h = h_min;

    IndexView& v = ix->view;
    std::memset(&v, 0, sizeof(v));
    v.n = n;
    v.ox = lo[0]; v.oy = lo[1]; v.oz = lo[2];
    v.h = h;
    v.inv_h = 1.0f / h;
    int maxdim = 1;
    for (int a = 0; a < 3; ++a) {
        v.dims[a] = (int)((h_bbox[3 + a] - lo[a]) * v.inv_h) + 1;  // same arithmetic as cell_coord()
        maxdim = std::max(maxdim, v.dims[a]);
    }
    v.bits = 1;
    while ((1 << v.bits) < maxdim) ++v.bits;
    v.num_levels = v.bits + 1;
    v.slack = 4.0f * 1.1920929e-7f * (float)maxdim + 1e-6f;
    v.slab_axis = -1;
    v.volumetric = est_dim > 2.5f ? 1 : 0;

    // 3. keys, sort, gather
    DeviceTemp keys_a(s, &scratch), keys_b(s, &scratch), vals_a(s, &scratch), vals_b(s, &scratch), sort_tmp(s, &scratch), hist(s, &scratch);
    PCT_CUDA(keys_a.alloc(sizeof(unsigned long long) * n));
    PCT_CUDA(keys_b.alloc(sizeof(unsigned long long) * n));
    PCT_CUDA(vals_a.alloc(sizeof(uint32_t) * n));
    PCT_CUDA(vals_b.alloc(sizeof(uint32_t) * n));
    const int blocks_n = (int)((n + kThreads - 1) / kThreads);
    morton_keys_kernel<<<blocks_n, kThreads, 0, s>>>(xyz, n, stride, v, keys_a.as<unsigned long long>(), vals_a.as<uint32_t>());
    cub::DoubleBuffer<unsigned long long> dk(keys_a.as<unsigned long long>(), keys_b.as<unsigned long long>());
    cub::DoubleBuffer<uint32_t> dv(vals_a.as<uint32_t>(), vals_b.as<uint32_t>());
    size_t tmp_bytes = 0;
    PCT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (long long)n, 0, 3 * v.bits, s));
    PCT_CUDA(sort_tmp.alloc(tmp_bytes));
    PCT_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp.p, tmp_bytes, dk, dv, (long long)n, 0, 3 * v.bits, s));
    const unsigned long long* keys = dk.Current();
    PCT_CUDA(cudaMallocAsync(&ix->pts, sizeof(Pt) * (size_t)n, s));
    gather_kernel<<<blocks_n, kThreads, 0, s>>>(xyz, n, stride, dv.Current(), ix->pts);
    v.pts = ix->pts;

    // 4. occupied cells per level
    PCT_CUDA(hist.alloc(sizeof(unsigned long long) * kMaxLevels));
    PCT_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned long long) * kMaxLevels, s));
    level_hist_kernel<unsigned long long><<<std::min(ix->sm_count * 8, blocks_n), kThreads, 0, s>>>(keys, n, hist.as<unsigned long long>());
    unsigned long long h_hist[kMaxLevels];
    PCT_CUDA(cudaMemcpyAsync(h_hist, hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));
    long long cells[kMaxLevels];
    long long acc = 1;
    for (int L = v.num_levels - 1; L >= 0; --L) {
        if (L < v.num_levels - 1) acc += (long long)h_hist[L];
        cells[L] = acc;
    }

    // 5. hash tables
    BuildTables bt;
    size_t total_slots = 0;
    size_t offset[kMaxLevels];
    for (int L = 0; L < v.num_levels; ++L) {
        size_t cap = 16;
        while (cap < (size_t)(2 * cells[L])) cap <<= 1;
        offset[L] = total_slots;
        total_slots += cap;
        bt.mask[L] = (uint32_t)(cap - 1);
    }
    PCT_CUDA(cudaMallocAsync(&ix->table_mem, sizeof(HashSlot) * total_slots, s));
    PCT_CUDA(cudaMemsetAsync(ix->table_mem, 0xFF, sizeof(HashSlot) * total_slots, s));
    bt.num_levels = v.num_levels;
    for (int L = 0; L < v.num_levels; ++L) {
        bt.slots[L] = ix->table_mem + offset[L];
        v.lvl[L].slots = bt.slots[L];
        v.lvl[L].mask = bt.mask[L];
    }
    fill_tables_kernel<<<(int)((n + 1 + kThreads - 1) / kThreads), kThreads, 0, s>>>(keys, n, bt);

    PCT_CUDA(cudaMallocAsync(&ix->stats, sizeof(unsigned int) * 8, s));
    PCT_CUDA(cudaMemsetAsync(ix->stats, 0, sizeof(unsigned int) * 8, s));
    PCT_CUDA(cudaGetLastError());

    pct_index_info& info = ix->info;
    info.num_points = n;
    info.cell_size = h;
    for (int a = 0; a < 3; ++a) { info.origin[a] = lo[a]; info.extent[a] = ext[a]; info.dims[a] = v.dims[a]; }
    info.bits_per_axis = v.bits;
    info.num_levels = v.num_levels;
    info.cells_level0 = cells[0];
    for (int L = 0; L < kMaxLevels; ++L) ix->cells_level[L] = L < v.num_levels ? cells[L] : 1;
    ix->est_dimension = est_dim;
    info.device_bytes = (int64_t)(sizeof(Pt) * (size_t)n + sizeof(HashSlot) * total_slots);
    info.est_dimension = est_dim;
    // bbox, Morton keys, record gather, level histogram, table fill (+ pilot keys and its level histogram)
    info.build_launches = cell_hint > 0.f ? 5 : 7;
    return PCT_OK;
}

}  // namespace pct

extern "C" {

int pct_index_build(const float* xyz, int64_t n, int stride, float cell_hint, int k_hint, void* stream, pct_index** out) {
    PCT_REQUIRE(out != nullptr, "pct_index_build: out is NULL");
    *out = nullptr;
    PCT_REQUIRE(xyz != nullptr, "pct_index_build: xyz is NULL");
    PCT_REQUIRE(n >= 1 && n < (1ll << 31), "pct_index_build: N must be in [1, 2^31)");
    PCT_REQUIRE(stride == 3 || stride == 4, "pct_index_build: stride must be 3 or 4 floats");
    pct_index* ix = new pct_index();
    const int rc = pct::build_impl(xyz, n, stride, cell_hint, k_hint, (cudaStream_t)stream, ix);
    if (rc != PCT_OK) {
        pct_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return PCT_OK;
}

int pct_estimate_cell_size(const float* xyz, int64_t n, int stride, int k_hint, void* stream, float* cell_size,
                           float* bbox_min_max) {
    PCT_REQUIRE(xyz && cell_size && n >= 1 && (stride == 3 || stride == 4), "pct_estimate_cell_size: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    int device = 0, sm_count = 148;
    PCT_CUDA(cudaGetDevice(&device));
    PCT_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    float h_bbox[6];
    pct::ScratchSession scratch(s, (size_t)32 << 20);
    int rc = pct::bounding_box(xyz, n, stride, sm_count, s, &scratch, h_bbox);
    if (rc != PCT_OK) return rc;
    const float lo[3] = {h_bbox[0], h_bbox[1], h_bbox[2]};
    const float extent_max = std::max(h_bbox[3] - h_bbox[0], std::max(h_bbox[4] - h_bbox[1], h_bbox[5] - h_bbox[2]));
    float h = 1.f, dim = 2.f;
    rc = pct::choose_cell_size(xyz, n, stride, n, lo, extent_max, k_hint, s, &scratch, &h, &dim);
    if (rc != PCT_OK) return rc;
    *cell_size = h;
    if (bbox_min_max)
        for (int a = 0; a < 6; ++a) bbox_min_max[a] = h_bbox[a];
    return PCT_OK;
}

int pct_index_destroy(pct_index* ix) {
    if (!ix) return PCT_OK;
    // stream-ordered free: safe against work already queued on the build stream, returns the
    // blocks to the pool for the next build without a device synchronisation
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != ix->device) cudaSetDevice(ix->device);
    if (ix->pts) cudaFreeAsync(ix->pts, ix->stream);
    if (ix->table_mem) cudaFreeAsync(ix->table_mem, ix->stream);
    if (ix->stats) cudaFreeAsync(ix->stats, ix->stream);
    if (ix->peers) cudaFreeAsync(ix->peers, ix->stream);
    if (cur != ix->device) cudaSetDevice(cur);
    delete ix;
    return PCT_OK;
}

int pct_index_get_info(const pct_index* ix, pct_index_info* info) {
    PCT_REQUIRE(ix && info, "pct_index_get_info: NULL argument");
    *info = ix->info;
    return PCT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// slab selection (multi-GPU): which points of the replicated cloud does one rank index, which does it own
// ---------------------------------------------------------------------------
namespace pct {
namespace {

struct InComplete {
    const float* xyz;
    int stride, axis;
    float c_lo, c_hi;
    __device__ __forceinline__ bool operator()(const int i) const {
        const float v = xyz[(size_t)i * stride + axis];
        return v >= c_lo && v <= c_hi;
    }
};

// local cloud (packed xyz) of the selected points and the flag "owned" per local point
__global__ void slab_gather_kernel(const float* __restrict__ xyz, int stride, int axis, const int32_t* __restrict__ sel,
                                   long long m, float own_lo, float own_hi, float* __restrict__ local,
                                   int32_t* __restrict__ own_flag) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= m) return;
    const float* p = xyz + (size_t)sel[r] * stride;
    const float x = p[0], y = p[1], z = p[2];
    local[3 * r] = x; local[3 * r + 1] = y; local[3 * r + 2] = z;
    const float v = axis == 0 ? x : (axis == 1 ? y : z);
    own_flag[r] = v >= own_lo && v < own_hi ? 1 : 0;
}

// original (whole-cloud) index of every output row: the id column of the owned points, in row order
__global__ void slab_row_ids_kernel(const float* __restrict__ xyz4, long long m, int axis, float own_lo, float own_hi,
                                    const int32_t* __restrict__ row_map, int32_t* __restrict__ row_ids) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= m) return;
    const float4 p = __ldg(reinterpret_cast<const float4*>(xyz4) + r);
    const float v = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    if (v >= own_lo && v < own_hi) row_ids[row_map[r]] = __float_as_int(p.w);
}

__global__ void slab_rows_kernel(const int32_t* __restrict__ incl, long long m, int32_t* __restrict__ row_map) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r < m) row_map[r] = incl[r] - 1;
}

}  // namespace
}  // namespace pct

extern "C" {

int pct_slab_select(const float* xyz, int64_t n, int stride, int axis, float complete_lo, float complete_hi,
                    int32_t* sel, int64_t* num_selected, void* stream) {
    PCT_REQUIRE(xyz && sel && num_selected && n >= 1 && n < (1ll << 31) && (stride == 3 || stride == 4) && axis >= 0 && axis <= 2,
                "pct_slab_select: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    pct::InComplete pred{xyz, stride, axis, complete_lo, complete_hi};
    cub::CountingInputIterator<int> ids(0);
    size_t tmp_bytes = 0;
    PCT_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, ids, sel, (int*)nullptr, (int)n, pred, s));
    pct::ScratchSession scratch(s, tmp_bytes + 4096);
    void* tmp = scratch.take(tmp_bytes);
    int* d_num = static_cast<int*>(scratch.take(sizeof(int)));
    const bool pooled = !tmp || !d_num;
    if (pooled) {
        PCT_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
        PCT_CUDA(cudaMallocAsync(&d_num, sizeof(int), s));
    }
    PCT_CUDA(cub::DeviceSelect::If(tmp, tmp_bytes, ids, sel, d_num, (int)n, pred, s));
    int h_num = 0;
    PCT_CUDA(cudaMemcpyAsync(&h_num, d_num, sizeof(int), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));
    if (pooled) {
        PCT_CUDA(cudaFreeAsync(tmp, s));
        PCT_CUDA(cudaFreeAsync(d_num, s));
    }
    *num_selected = h_num;
    return PCT_OK;
}

int pct_slab_gather(const float* xyz, int stride, int axis, const int32_t* sel, int64_t m, float own_lo, float own_hi,
                    float* local_xyz, int32_t* row_map, int64_t* num_owned, void* stream) {
    PCT_REQUIRE(xyz && sel && local_xyz && row_map && num_owned && m >= 0 && (stride == 3 || stride == 4), "pct_slab_gather: bad argument");
    *num_owned = 0;
    if (m == 0) return PCT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    size_t tmp_bytes = 0;
    PCT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)m, s));
    pct::ScratchSession scratch(s, tmp_bytes + 2 * sizeof(int32_t) * (size_t)m + 8192);
    void* tmp = scratch.take(tmp_bytes);
    int32_t* flag = static_cast<int32_t*>(scratch.take(sizeof(int32_t) * (size_t)m));
    int32_t* incl = static_cast<int32_t*>(scratch.take(sizeof(int32_t) * (size_t)m));
    const bool pooled = !tmp || !flag || !incl;
    if (pooled) {
        PCT_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
        PCT_CUDA(cudaMallocAsync(&flag, sizeof(int32_t) * (size_t)m, s));
        PCT_CUDA(cudaMallocAsync(&incl, sizeof(int32_t) * (size_t)m, s));
    }
    const int blocks = (int)((m + 255) / 256);
    pct::slab_gather_kernel<<<blocks, 256, 0, s>>>(xyz, stride, axis, sel, m, own_lo, own_hi, local_xyz, flag);
    PCT_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, flag, incl, (int)m, s));
    pct::slab_rows_kernel<<<blocks, 256, 0, s>>>(incl, m, row_map);
    int32_t h_last = 0;
    PCT_CUDA(cudaMemcpyAsync(&h_last, incl + (m - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));
    PCT_CUDA(cudaGetLastError());
    if (pooled) {
        PCT_CUDA(cudaFreeAsync(tmp, s));
        PCT_CUDA(cudaFreeAsync(flag, s));
        PCT_CUDA(cudaFreeAsync(incl, s));
    }
    *num_owned = h_last;
    return PCT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// slab exchange (multi-GPU): every rank holds a contiguous share of the cloud (original indices id_base ..
// id_base + n), bins it by destination slab and the ranks trade the bins with ONE all-to-all.  A point goes to
// every slab whose complete range [complete_lo, complete_hi] holds its coordinate (its owner, and the neighbours
// whose margin it lies in).  Bins keep ascending original index, so the concatenation a slab receives in rank
// order is ascending too and distance ties keep the whole cloud's order.
// ---------------------------------------------------------------------------
namespace pct {
namespace {

constexpr int kBinThreads = 256, kBinItems = 4, kBinWarps = kBinThreads / 32;
constexpr int kBinBlockPoints = kBinThreads * kBinItems;
constexpr int kBinStage = kBinBlockPoints + kBinBlockPoints / 2;  // records of one block staged in shared memory (24 KB)
constexpr int kMaxWorld = 32;

struct SlabBounds {
    float c_lo[kMaxWorld], c_hi[kMaxWorld], own_lo[kMaxWorld], own_hi[kMaxWorld];
    int world;
};

// Where the records of destination d go: dest.base[d] + (position in the concatenated local order), or the local
// `records` array when dest.base[d] is null.  A peer pointer (the slab buffer of rank d, mapped over NVLink) arrives
// already shifted by (first row this rank may write there) - (first position of bin d).
struct SlabDest {
    float4* base[kMaxWorld];
};

// bit d of the result: the point belongs to slab d's complete range; `owner` = the slab that answers it (-1: none)
__device__ __forceinline__ unsigned int slab_mask(const SlabBounds& b, float x, int& owner) {
    unsigned int m = 0;
    owner = -1;
    for (int d = 0; d < b.world; ++d) {
        if (x >= b.c_lo[d] && x <= b.c_hi[d]) m |= 1u << d;
        if (x >= b.own_lo[d] && x < b.own_hi[d]) owner = d;
    }
    return m;
}

// counts[row][block], rows 0 .. world-1 = complete members per destination, rows world .. 2 world-1 = owned members.
// FILL: `pos` = exclusive scan of the flattened counts, i.e. the output position of the block's first member of a row.
template <bool FILL>
__global__ void __launch_bounds__(kBinThreads)
slab_bin_kernel(const float* __restrict__ xyz, const long long n, const int stride, const int axis, const SlabBounds b,
                const long long blocks, int32_t* __restrict__ counts, const int32_t* __restrict__ pos,
                const long long total_complete, const long long id_base, float4* __restrict__ records,
                int32_t* __restrict__ owned_local, const SlabDest dest) {
    __shared__ int cnt[2 * kMaxWorld][kBinItems * kBinWarps + 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long base = (long long)blockIdx.x * kBinBlockPoints;
    unsigned int mask[kBinItems];
    int owner[kBinItems];
#pragma unroll
    for (int j = 0; j < kBinItems; ++j) {
        const long long i = base + j * kBinThreads + t;
        mask[j] = 0;
        owner[j] = -1;
        if (i < n) mask[j] = slab_mask(b, __ldg(xyz + i * stride + axis), owner[j]);
    }
    // per (row, item, warp) member counts; item-major then warp = ascending point index
    for (int d = 0; d < b.world; ++d) {
#pragma unroll
        for (int j = 0; j < kBinItems; ++j) {
            const unsigned int bc = __ballot_sync(0xffffffffu, (mask[j] >> d) & 1u);
            const unsigned int bo = __ballot_sync(0xffffffffu, owner[j] == d);
            if (lane == 0) {
                cnt[d][j * kBinWarps + warp] = __popc(bc);
                cnt[b.world + d][j * kBinWarps + warp] = __popc(bo);
            }
        }
    }
    __syncthreads();
    if (t < 2 * b.world) {  // exclusive scan of the row's 32 entries; the total goes to the extra slot
        int run = 0;
        for (int e = 0; e < kBinItems * kBinWarps; ++e) {
            const int c = cnt[t][e];
            cnt[t][e] = run;
            run += c;
        }
        cnt[t][kBinItems * kBinWarps] = run;
        if (!FILL) counts[(long long)t * blocks + blockIdx.x] = run;
    }
    if (!FILL) return;
    __syncthreads();
    const unsigned int lt = (1u << lane) - 1u;
    // The block's records are first grouped by destination in shared memory and then copied out destination by
    // destination, so that a destination receives the block's share as ONE contiguous piece (about 2 KB at 8 slabs)
    // instead of a few 16-byte records per warp: what a peer buffer behind NVLink needs, and kinder to local DRAM too.
    // A block whose records do not fit (margins that cover most of the cloud: tiny clouds) writes them directly.
    __shared__ float4 staged[kBinStage];
    __shared__ int dest_off[kMaxWorld + 1];
    if (t == 0) {
        int run = 0;
        for (int d = 0; d < b.world; ++d) { dest_off[d] = run; run += cnt[d][kBinItems * kBinWarps]; }
        dest_off[b.world] = run;
    }
    __syncthreads();
    const bool stage = dest_off[b.world] <= kBinStage;
    for (int d = 0; d < b.world; ++d) {
        const long long p_c = pos[(long long)d * blocks + blockIdx.x];
        float4* const out_d = dest.base[d] ? dest.base[d] : records;
        const long long p_o = (long long)pos[(long long)(b.world + d) * blocks + blockIdx.x] - total_complete;
#pragma unroll
        for (int j = 0; j < kBinItems; ++j) {
            const bool fc = (mask[j] >> d) & 1u, fo = owner[j] == d;
            const unsigned int bc = __ballot_sync(0xffffffffu, fc);
            const unsigned int bo = __ballot_sync(0xffffffffu, fo);
            const long long i = base + j * kBinThreads + t;
            if (fc) {
                const float* p = xyz + i * stride;
                float4 r;
                r.x = __ldg(p); r.y = __ldg(p + 1); r.z = __ldg(p + 2);
                r.w = __int_as_float((int)(id_base + i));
                const int within = cnt[d][j * kBinWarps + warp] + __popc(bc & lt);
                if (stage) staged[dest_off[d] + within] = r;
                else out_d[p_c + within] = r;
            }
            if (fo) owned_local[p_o + cnt[b.world + d][j * kBinWarps + warp] + __popc(bo & lt)] = (int32_t)i;
        }
    }
    if (!stage) return;
    __syncthreads();
    for (int d = 0; d < b.world; ++d) {
        const long long p_c = pos[(long long)d * blocks + blockIdx.x];
        float4* const out_d = dest.base[d] ? dest.base[d] : records;
        const int first = dest_off[d], count = dest_off[d + 1] - first;
        for (int e = t; e < count; e += kBinThreads) out_d[p_c + e] = staged[first + e];
    }
}

__global__ void slab_row_starts_kernel(const int32_t* __restrict__ pos, long long blocks, int rows, int32_t* __restrict__ out) {
    const int r = threadIdx.x;
    if (r < rows) out[r] = pos[(long long)r * blocks];
}

__global__ void slab_own_flag_kernel(const float* __restrict__ xyz, long long m, int stride, int axis, float own_lo,
                                     float own_hi, int32_t* __restrict__ flag) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= m) return;
    const float v = __ldg(xyz + r * stride + axis);
    flag[r] = v >= own_lo && v < own_hi ? 1 : 0;
}

int fill_bounds(SlabBounds& b, int world, const float* bounds) {
    if (world < 1 || world > kMaxWorld || !bounds) return PCT_ERR_INVALID_ARGUMENT;
    b.world = world;
    for (int d = 0; d < world; ++d) {
        b.c_lo[d] = bounds[4 * d]; b.c_hi[d] = bounds[4 * d + 1];
        b.own_lo[d] = bounds[4 * d + 2]; b.own_hi[d] = bounds[4 * d + 3];
    }
    return PCT_OK;
}

}  // namespace
}  // namespace pct

extern "C" {

int pct_estimate_cell_size_sample(const float* sample, int64_t n_sample, int stride, int64_t n_total,
                                  const float* bbox_min_max, int k_hint, void* stream, float* cell_size) {
    PCT_REQUIRE(sample && cell_size && bbox_min_max && n_sample >= 1 && n_total >= n_sample && (stride == 3 || stride == 4),
                "pct_estimate_cell_size_sample: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    pct::ScratchSession scratch(s, (size_t)32 << 20);
    const float lo[3] = {bbox_min_max[0], bbox_min_max[1], bbox_min_max[2]};
    const float extent_max = std::max(bbox_min_max[3] - lo[0], std::max(bbox_min_max[4] - lo[1], bbox_min_max[5] - lo[2]));
    float h = 1.f, dim = 2.f;
    const int rc = pct::choose_cell_size(sample, n_sample, stride, n_total, lo, extent_max, k_hint, s, &scratch, &h, &dim);
    if (rc != PCT_OK) return rc;
    *cell_size = h;
    return PCT_OK;
}

int64_t pct_slab_bin_blocks(int64_t n) { return (n + pct::kBinBlockPoints - 1) / pct::kBinBlockPoints; }

int pct_slab_bin_count(const float* xyz, int64_t n, int stride, int axis, int world, const float* bounds,
                       int32_t* block_pos, int64_t* counts, void* stream) {
    PCT_REQUIRE(xyz && block_pos && counts && n >= 1 && n < (1ll << 31) && (stride == 3 || stride == 4) && axis >= 0 && axis <= 2,
                "pct_slab_bin_count: bad argument");
    pct::SlabBounds b;
    PCT_REQUIRE(pct::fill_bounds(b, world, bounds) == PCT_OK, "pct_slab_bin_count: world must be in [1, 32]");
    cudaStream_t s = (cudaStream_t)stream;
    const long long blocks = pct_slab_bin_blocks(n);
    const long long cells = 2ll * world * blocks;
    PCT_REQUIRE(cells + 1 < (1ll << 31), "pct_slab_bin_count: share too large");
    pct::slab_bin_kernel<false><<<(unsigned int)blocks, pct::kBinThreads, 0, s>>>(xyz, n, stride, axis, b, blocks, block_pos, nullptr, 0, 0,
                                                                                  nullptr, nullptr, pct::SlabDest{});
    // exclusive scan of the flattened counts in place (+ one closing entry = grand total)
    size_t tmp_bytes = 0;
    PCT_CUDA(cudaMemsetAsync(block_pos + cells, 0, sizeof(int32_t), s));
    PCT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, block_pos, block_pos, (int)(cells + 1), s));
    pct::ScratchSession scratch(s, tmp_bytes + 4096);
    void* tmp = scratch.take(tmp_bytes);
    int32_t* starts = static_cast<int32_t*>(scratch.take(sizeof(int32_t) * (2 * pct::kMaxWorld + 1)));
    const bool pooled = !tmp || !starts;
    if (pooled) {
        PCT_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
        PCT_CUDA(cudaMallocAsync(&starts, sizeof(int32_t) * (2 * pct::kMaxWorld + 1), s));
    }
    PCT_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, block_pos, block_pos, (int)(cells + 1), s));
    pct::slab_row_starts_kernel<<<1, 2 * pct::kMaxWorld + 1, 0, s>>>(block_pos, blocks, 2 * world + 1, starts);
    int32_t h_starts[2 * pct::kMaxWorld + 1];
    PCT_CUDA(cudaMemcpyAsync(h_starts, starts, sizeof(int32_t) * (2 * world + 1), cudaMemcpyDeviceToHost, s));
    PCT_CUDA(cudaStreamSynchronize(s));
    PCT_CUDA(cudaGetLastError());
    if (pooled) {
        PCT_CUDA(cudaFreeAsync(tmp, s));
        PCT_CUDA(cudaFreeAsync(starts, s));
    }
    for (int r = 0; r < 2 * world; ++r) counts[r] = (int64_t)h_starts[r + 1] - (int64_t)h_starts[r];
    return PCT_OK;
}

int pct_slab_bin_fill(const float* xyz, int64_t n, int stride, int axis, int world, const float* bounds,
                      const int32_t* block_pos, int64_t total_complete, int64_t id_base, float* records,
                      int32_t* owned_local, void* stream) {
    PCT_REQUIRE(xyz && block_pos && records && owned_local && n >= 1 && (stride == 3 || stride == 4) && axis >= 0 && axis <= 2 &&
                    id_base >= 0 && id_base + n <= (1ll << 31),
                "pct_slab_bin_fill: bad argument");
    pct::SlabBounds b;
    PCT_REQUIRE(pct::fill_bounds(b, world, bounds) == PCT_OK, "pct_slab_bin_fill: world must be in [1, 32]");
    cudaStream_t s = (cudaStream_t)stream;
    const long long blocks = pct_slab_bin_blocks(n);
    pct::slab_bin_kernel<true><<<(unsigned int)blocks, pct::kBinThreads, 0, s>>>(xyz, n, stride, axis, b, blocks, nullptr, block_pos,
                                                                                 total_complete, id_base,
                                                                                 reinterpret_cast<float4*>(records), owned_local, pct::SlabDest{});
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int pct_slab_bin_fill_peers(const float* xyz, int64_t n, int stride, int axis, int world, const float* bounds,
                            const int32_t* block_pos, const int64_t* complete_counts, int64_t id_base,
                            void* const* peer_slabs, const int64_t* dest_rows, int32_t* owned_local, void* stream) {
    PCT_REQUIRE(xyz && block_pos && complete_counts && peer_slabs && dest_rows && owned_local && n >= 1 && (stride == 3 || stride == 4) &&
                    axis >= 0 && axis <= 2 && id_base >= 0 && id_base + n <= (1ll << 31),
                "pct_slab_bin_fill_peers: bad argument");
    pct::SlabBounds b;
    PCT_REQUIRE(pct::fill_bounds(b, world, bounds) == PCT_OK, "pct_slab_bin_fill_peers: world must be in [1, 32]");
    pct::SlabDest dest{};
    long long bin_start = 0;
    for (int d = 0; d < world; ++d) {
        PCT_REQUIRE(complete_counts[d] == 0 || (peer_slabs[d] && (reinterpret_cast<uintptr_t>(peer_slabs[d]) & 15) == 0 && dest_rows[d] >= 0),
                    "pct_slab_bin_fill_peers: every destination needs a 16-byte aligned slab buffer");
        dest.base[d] = peer_slabs[d] ? static_cast<float4*>(peer_slabs[d]) + (dest_rows[d] - bin_start) : nullptr;
        bin_start += complete_counts[d];
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long blocks = pct_slab_bin_blocks(n);
    // (a destination without a buffer holds no record of this share: its null base is never dereferenced)
    pct::slab_bin_kernel<true><<<(unsigned int)blocks, pct::kBinThreads, 0, s>>>(xyz, n, stride, axis, b, blocks, nullptr, block_pos,
                                                                                 bin_start, id_base, nullptr, owned_local, dest);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int pct_slab_row_ids(const float* xyz4, int64_t m, int axis, float own_lo, float own_hi, const int32_t* row_map,
                     int32_t* row_ids, void* stream) {
    PCT_REQUIRE(xyz4 && row_map && row_ids && m >= 1 && axis >= 0 && axis <= 2 && (reinterpret_cast<uintptr_t>(xyz4) & 15) == 0,
                "pct_slab_row_ids: bad argument");
    pct::slab_row_ids_kernel<<<(int)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz4, m, axis, own_lo, own_hi, row_map, row_ids);
    PCT_CUDA(cudaGetLastError());
    return PCT_OK;
}

int pct_slab_rows(const float* xyz, int64_t m, int stride, int axis, float own_lo, float own_hi, int32_t* row_map,
                  void* stream) {
    PCT_REQUIRE(xyz && row_map && m >= 1 && m < (1ll << 31) && (stride == 3 || stride == 4) && axis >= 0 && axis <= 2,
                "pct_slab_rows: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    size_t tmp_bytes = 0;
    PCT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)m, s));
    pct::ScratchSession scratch(s, tmp_bytes + 2 * sizeof(int32_t) * (size_t)m + 8192);
    void* tmp = scratch.take(tmp_bytes);
    int32_t* flag = static_cast<int32_t*>(scratch.take(sizeof(int32_t) * (size_t)m));
    int32_t* incl = static_cast<int32_t*>(scratch.take(sizeof(int32_t) * (size_t)m));
    const bool pooled = !tmp || !flag || !incl;
    if (pooled) {
        PCT_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
        PCT_CUDA(cudaMallocAsync(&flag, sizeof(int32_t) * (size_t)m, s));
        PCT_CUDA(cudaMallocAsync(&incl, sizeof(int32_t) * (size_t)m, s));
    }
    const int blocks = (int)((m + 255) / 256);
    pct::slab_own_flag_kernel<<<blocks, 256, 0, s>>>(xyz, m, stride, axis, own_lo, own_hi, flag);
    PCT_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, flag, incl, (int)m, s));
    pct::slab_rows_kernel<<<blocks, 256, 0, s>>>(incl, m, row_map);
    PCT_CUDA(cudaGetLastError());
    if (pooled) {
        PCT_CUDA(cudaFreeAsync(tmp, s));
        PCT_CUDA(cudaFreeAsync(flag, s));
        PCT_CUDA(cudaFreeAsync(incl, s));
    }
    return PCT_OK;
}

}  // extern "C"
