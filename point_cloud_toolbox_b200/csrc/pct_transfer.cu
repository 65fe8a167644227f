// Host -> device copy of a caller's array that is NOT page-locked (a numpy array handed to PointCloud(points=...),
// ref :26-47): cudaMemcpyAsync stages such a source through one driver-owned buffer with a single host thread,
// about a fifth of the PCIe rate.  Here several host threads copy 8 MB pieces into their own page-locked
// double buffers and queue each piece's DMA on the caller's stream as soon as it is staged, so the memcpy of one
// piece overlaps the DMA of others and the link is the limit.  Page-locked sources take the direct path.
#include <atomic>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "pct_internal.h"

namespace pct {
namespace {

constexpr size_t kPiece = (size_t)8 << 20;
constexpr int kWorkers = 4;
constexpr int kSlots = 2;

struct Stage {
    char* buf[kWorkers][kSlots] = {};
    cudaEvent_t done[kWorkers][kSlots] = {};
    int device = -1;
    std::mutex mutex;  // one staged upload at a time per process (the buffers are shared)
};
Stage g_stage;

int ensure_stage(int device) {
    if (g_stage.device == device) return PCT_OK;
    if (g_stage.device >= 0) {  // the process moved to another device: events belong to a device, rebuild them
        for (int w = 0; w < kWorkers; ++w)
            for (int b = 0; b < kSlots; ++b) {
                if (g_stage.done[w][b]) { cudaEventSynchronize(g_stage.done[w][b]); cudaEventDestroy(g_stage.done[w][b]); g_stage.done[w][b] = nullptr; }
            }
    }
    for (int w = 0; w < kWorkers; ++w)
        for (int b = 0; b < kSlots; ++b) {
            if (!g_stage.buf[w][b]) PCT_CUDA(cudaMallocHost(&g_stage.buf[w][b], kPiece));
            PCT_CUDA(cudaEventCreateWithFlags(&g_stage.done[w][b], cudaEventDisableTiming));
        }
    g_stage.device = device;
    return PCT_OK;
}

}  // namespace

// page-locked staging buffers of pct_upload (64 MB) go with the scratch arenas (pct_release_scratch)
void release_upload_stage() {
    std::lock_guard<std::mutex> lock(g_stage.mutex);
    for (int w = 0; w < kWorkers; ++w)
        for (int b = 0; b < kSlots; ++b) {
            if (g_stage.done[w][b]) { cudaEventSynchronize(g_stage.done[w][b]); cudaEventDestroy(g_stage.done[w][b]); g_stage.done[w][b] = nullptr; }
            if (g_stage.buf[w][b]) { cudaFreeHost(g_stage.buf[w][b]); g_stage.buf[w][b] = nullptr; }
        }
    g_stage.device = -1;
}

}  // namespace pct

extern "C" int pct_upload(void* dst_device, const void* src_host, int64_t bytes, void* stream) {
    using namespace pct;
    PCT_REQUIRE(bytes >= 0 && ((dst_device && src_host) || bytes == 0), "pct_upload: bad argument");
    if (bytes == 0) return PCT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, src_host) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();  // an unregistered pointer is not an error here
    if (pinned || (size_t)bytes < 4 * kPiece) {
        PCT_CUDA(cudaMemcpyAsync(dst_device, src_host, (size_t)bytes, cudaMemcpyHostToDevice, s));
        return PCT_OK;
    }
    int device = 0;
    PCT_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(g_stage.mutex);
    {
        const int rc = ensure_stage(device);
        if (rc != PCT_OK) return rc;
    }
    const int64_t pieces = (bytes + (int64_t)kPiece - 1) / (int64_t)kPiece;
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&](int w) {
        if (cudaSetDevice(device) != cudaSuccess) { failed.store((int)cudaErrorInvalidDevice); return; }
        int slot = 0;
        for (;;) {
            const int64_t p = next.fetch_add(1, std::memory_order_relaxed);
            if (p >= pieces || failed.load(std::memory_order_relaxed)) return;
            const size_t off = (size_t)p * kPiece, len = std::min(kPiece, (size_t)bytes - off);
            cudaError_t e = cudaEventSynchronize(g_stage.done[w][slot]);  // the DMA that last read this buffer
            if (e == cudaSuccess) {
                std::memcpy(g_stage.buf[w][slot], static_cast<const char*>(src_host) + off, len);
                e = cudaMemcpyAsync(static_cast<char*>(dst_device) + off, g_stage.buf[w][slot], len, cudaMemcpyHostToDevice, s);
            }
            if (e == cudaSuccess) e = cudaEventRecord(g_stage.done[w][slot], s);
            if (e != cudaSuccess) { failed.store((int)e); return; }
            slot ^= 1;
        }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < kWorkers; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& th : pool) th.join();
    if (failed.load()) return cuda_fail((cudaError_t)failed.load(), "pct_upload (staged copy)", __FILE__, __LINE__);
    return PCT_OK;
}
