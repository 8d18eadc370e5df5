// search_kernels.h — the traversal kernel's instantiations live in their own translation units (one per metric and
// form, compiled in parallel); the launch code picks one through these getters.
#pragma once

#include "hnsw_search.cuh"

namespace turdb {

using SearchKernelFn = void (*)(const SearchArgs);

// staged form (hnsw_search_kernel): team of 2-4 warps per query, rows through shared-memory staging
SearchKernelFn get_staged_kernel_l2(bool global_visited, bool filtered, bool sq8);
SearchKernelFn get_staged_kernel_cosine(bool global_visited, bool filtered, bool sq8);
SearchKernelFn get_staged_kernel_ip(bool global_visited, bool filtered, bool sq8);
// direct form (hnsw_search_warp_kernel): one warp per query, rows straight into registers (FP32 rows only)
SearchKernelFn get_direct_kernel_l2(bool global_visited, bool filtered);
SearchKernelFn get_direct_kernel_cosine(bool global_visited, bool filtered);
SearchKernelFn get_direct_kernel_ip(bool global_visited, bool filtered);

// the insert path's searches (squared L2 only): compiled with the L2 units
SearchKernelFn get_insert_kernel(bool global_visited, bool direct);

inline SearchKernelFn get_search_kernel(int metric, bool global_visited, bool filtered, bool sq8, bool direct) {
  if (direct) {
    switch (metric) {
      case kCosine: return get_direct_kernel_cosine(global_visited, filtered);
      case kIP: return get_direct_kernel_ip(global_visited, filtered);
      default: return get_direct_kernel_l2(global_visited, filtered);
    }
  }
  switch (metric) {
    case kCosine: return get_staged_kernel_cosine(global_visited, filtered, sq8);
    case kIP: return get_staged_kernel_ip(global_visited, filtered, sq8);
    default: return get_staged_kernel_l2(global_visited, filtered, sq8);
  }
}

}  // namespace turdb
