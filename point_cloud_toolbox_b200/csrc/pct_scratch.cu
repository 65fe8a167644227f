// Scratch arenas for the large temporaries of one library call (sort buffers of the index build,
// work queues of a query): one growing cudaMalloc block per (device, stream), handed out by bump
// allocation for the duration of a call and reused by the next call on that stream.
//
// Why not the stream-ordered pool (cudaMallocAsync) for these: a caller that builds the next index
// while the previous one is still alive and never synchronises makes the pool re-map gigabytes of
// physical memory inside cudaMallocAsync -- index builds of 10 ms were measured to take up to 1.5 s
// (scripts/build_probe.py).  Calls on one stream execute in order, so a later call may overwrite the
// arena of an earlier one as soon as it is enqueued behind it; different streams get different arenas.
#include <map>
#include <mutex>
#include <utility>

#include "pct_internal.h"

namespace pct {

namespace {

struct Arena {
    void* base = nullptr;
    size_t cap = 0;
};

std::mutex g_mutex;
std::map<std::pair<int, cudaStream_t>, Arena> g_arenas;

}  // namespace

ScratchSession::ScratchSession(cudaStream_t s, size_t bytes) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return;
    bytes = (bytes + 4095) & ~(size_t)4095;
    std::lock_guard<std::mutex> lock(g_mutex);
    Arena& a = g_arenas[std::make_pair(device, s)];
    if (a.cap < bytes) {
        // grow: everything that may still read the old block is queued on this stream
        if (a.base) {
            cudaStreamSynchronize(s);
            cudaFree(a.base);
            a.base = nullptr;
            a.cap = 0;
        }
        const size_t want = bytes + bytes / 8;
        if (cudaMalloc(&a.base, want) == cudaSuccess) {
            a.cap = want;
        } else {
            cudaGetLastError();  // callers fall back to the stream-ordered pool
            a.base = nullptr;
        }
    }
    p_ = static_cast<char*>(a.base);
    left_ = a.cap;
}

void* ScratchSession::take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (!p_ || bytes > left_) return nullptr;
    void* r = p_;
    p_ += bytes;
    left_ -= bytes;
    return r;
}

void release_scratch() {
    std::lock_guard<std::mutex> lock(g_mutex);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_arenas) {
        if (!kv.second.base) continue;
        cudaSetDevice(kv.first.first);
        // the stream belongs to the caller and may be gone by now (an invalid handle is an error code, not a fault;
        // whoever destroys a stream should call release_scratch_of first): fall back to a device synchronisation
        if (cudaStreamSynchronize(kv.first.second) != cudaSuccess) {
            cudaGetLastError();
            cudaDeviceSynchronize();
        }
        cudaFree(kv.second.base);
    }
    g_arenas.clear();
    cudaSetDevice(cur);
}

// frees the arena of ONE stream of the current device (synchronises that stream): for library-owned streams that
// are about to be destroyed -- an arena keyed by a dead stream handle would never be reused or freed
void release_scratch_of(cudaStream_t s) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return;
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_arenas.find(std::make_pair(device, s));
    if (it == g_arenas.end()) return;
    if (it->second.base) {
        cudaStreamSynchronize(s);
        cudaFree(it->second.base);
    }
    g_arenas.erase(it);
}

}  // namespace pct

extern "C" int pct_release_scratch(void) {
    pct::release_scratch();
    pct::release_upload_stage();
    return PCT_OK;
}
