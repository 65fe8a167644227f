// knn_fast_kernel<50, *>: see pct_knn_fast.cuh
#include "pct_knn_fast.cuh"

namespace pct {
PCT_INSTANTIATE_FAST(50)
}  // namespace pct
