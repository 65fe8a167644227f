// knn_fast_kernel<64, *>: see pct_knn_fast.cuh
#include "pct_knn_fast.cuh"

namespace pct {
PCT_INSTANTIATE_FAST(64)
}  // namespace pct
