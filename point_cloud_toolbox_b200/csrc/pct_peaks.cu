// Measured arithmetic peaks of the device the library runs on: MEASURED_PEAKS.json holds the HBM copy
// bandwidth and the bf16 tensor throughput, but the kernels of this path run on the CUDA cores, and
// SURVEY.md section 8(d) asks for the FP32 (and the FP64) FMA rate as the secondary roofline.
// Every thread runs eight independent FMA chains; 2048 threads per SM.  Diagnostics only.
#include <algorithm>

#include "pct_internal.h"

namespace pct {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T* __restrict__ out, int iters, T b, T c) {
    T a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (T)(threadIdx.x + j) * (T)1e-3;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == (T)123456789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chains alive
}

template <typename T>
int measure(int sms, int iters, cudaStream_t s, T* scratch, double* tflops) {
    cudaEvent_t e0, e1;
    PCT_CUDA(cudaEventCreate(&e0));
    PCT_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {  // the first repetition warms up
        cudaEventRecord(e0, s);
        fma_chain_kernel<T><<<blocks, 256, 0, s>>>(scratch, iters, (T)0.999999, (T)1e-6);
        cudaEventRecord(e1, s);
        const cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); PCT_CUDA(err); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * 4.0 * (double)iters * (double)blocks * 256.0;
        if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return PCT_OK;
}

}  // namespace
}  // namespace pct

extern "C" int pct_measure_fma_peaks(double* fp32_tflops, double* fp64_tflops, void* stream) {
    PCT_REQUIRE(fp32_tflops && fp64_tflops, "pct_measure_fma_peaks: NULL argument");
    int dev = 0, sms = 0;
    PCT_CUDA(cudaGetDevice(&dev));
    PCT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s = (cudaStream_t)stream;
    void* scratch = nullptr;
    PCT_CUDA(cudaMalloc(&scratch, (size_t)sms * 8 * 256 * sizeof(double)));
    int rc = pct::measure<float>(sms, 1 << 15, s, static_cast<float*>(scratch), fp32_tflops);     // ~ 60 ms
    if (rc == PCT_OK) rc = pct::measure<double>(sms, 1 << 14, s, static_cast<double*>(scratch), fp64_tflops);
    cudaFree(scratch);
    return rc;
}
