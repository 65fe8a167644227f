// knn_fast_kernel<32, *>: see pct_knn_fast.cuh
#include "pct_knn_fast.cuh"

namespace pct {
PCT_INSTANTIATE_FAST(32)
}  // namespace pct
