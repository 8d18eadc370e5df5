// search_kernels_staged.cu — instantiations of the staged traversal kernel for ONE metric (-DTURDB_TU_METRIC=0|1|2).
#include "search_kernels.h"

#ifndef TURDB_TU_METRIC
#error "compile with -DTURDB_TU_METRIC=0|1|2"
#endif

namespace turdb {

#if TURDB_TU_METRIC == 0
#define TURDB_TU_GETTER get_staged_kernel_l2
#elif TURDB_TU_METRIC == 1
#define TURDB_TU_GETTER get_staged_kernel_cosine
#else
#define TURDB_TU_GETTER get_staged_kernel_ip
#endif

SearchKernelFn TURDB_TU_GETTER(bool gv, bool filt, bool sq8) {
  constexpr int M = TURDB_TU_METRIC;
  if (sq8) {
    if (gv) return filt ? hnsw_search_kernel<M, true, true, true> : hnsw_search_kernel<M, true, false, true>;
    return filt ? hnsw_search_kernel<M, false, true, true> : hnsw_search_kernel<M, false, false, true>;
  }
  if (gv) return filt ? hnsw_search_kernel<M, true, true, false> : hnsw_search_kernel<M, true, false, false>;
  return filt ? hnsw_search_kernel<M, false, true, false> : hnsw_search_kernel<M, false, false, false>;
}

#if TURDB_TU_METRIC == 0
SearchKernelFn get_insert_kernel_staged(bool gv) { return gv ? hnsw_insert_search_kernel<true> : hnsw_insert_search_kernel<false>; }
#endif

}  // namespace turdb
