// knn_fast_kernel<100, *>: see pct_knn_fast.cuh
#include "pct_knn_fast.cuh"

namespace pct {
PCT_INSTANTIATE_FAST(100)
}  // namespace pct
