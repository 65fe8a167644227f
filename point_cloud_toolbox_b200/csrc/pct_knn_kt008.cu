// knn_fast_kernel<8, *>: see pct_knn_fast.cuh
#include "pct_knn_fast.cuh"

namespace pct {
PCT_INSTANTIATE_FAST(8)
}  // namespace pct
