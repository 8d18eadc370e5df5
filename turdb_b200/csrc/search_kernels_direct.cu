// search_kernels_direct.cu — instantiations of the direct (one warp per query) traversal kernel for ONE metric
// (-DTURDB_TU_METRIC=0|1|2).
#include "search_kernels.h"

#ifndef TURDB_TU_METRIC
#error "compile with -DTURDB_TU_METRIC=0|1|2"
#endif

namespace turdb {

#if TURDB_TU_METRIC == 0
#define TURDB_TU_GETTER get_direct_kernel_l2
#elif TURDB_TU_METRIC == 1
#define TURDB_TU_GETTER get_direct_kernel_cosine
#else
#define TURDB_TU_GETTER get_direct_kernel_ip
#endif

SearchKernelFn TURDB_TU_GETTER(bool gv, bool filt) {
  constexpr int M = TURDB_TU_METRIC;
  if (gv) return filt ? hnsw_search_warp_kernel<M, true, true> : hnsw_search_warp_kernel<M, true, false>;
  return filt ? hnsw_search_warp_kernel<M, false, true> : hnsw_search_warp_kernel<M, false, false>;
}

#if TURDB_TU_METRIC == 0
SearchKernelFn get_insert_kernel_staged(bool gv);
SearchKernelFn get_insert_kernel(bool gv, bool direct) {
  if (!direct) return get_insert_kernel_staged(gv);
  return gv ? hnsw_insert_search_warp_kernel<true> : hnsw_insert_search_warp_kernel<false>;
}
#endif

}  // namespace turdb
