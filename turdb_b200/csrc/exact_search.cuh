// exact_search.cuh — exact (brute-force) top-k path and the multi-shard top-k merge.
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace turdb {

// One warp per query: k-way merge of n_shards ascending lists, ties by (distance, row_id).
// gathered_* are [n_shards][nq][k]; lane s walks shard s (n_shards <= 32).
__global__ void merge_topk_kernel(const uint64_t* __restrict__ g_rows, const float* __restrict__ g_dist,
                                  const uint32_t* __restrict__ g_counts, uint32_t n_shards, uint32_t nq,
                                  uint32_t k, uint64_t* __restrict__ out_rows, float* __restrict__ out_dist,
                                  uint32_t* __restrict__ out_counts) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  uint32_t head = 0, cnt = 0;
  size_t base = 0;
  if (lane < n_shards) {
    cnt = min(g_counts[(size_t)lane * nq + q], k);
    base = ((size_t)lane * nq + q) * k;
  }
  uint32_t produced = 0;
  for (; produced < k; ++produced) {
    float d = INFINITY;
    uint64_t r = 0xFFFFFFFFFFFFFFFFull;
    bool have = head < cnt;
    if (have) {
      d = g_dist[base + head];
      r = g_rows[base + head];
    }
    if (!__any_sync(kFullMask, have)) break;
    float bd = d;
    uint64_t br = r;
    uint32_t bl = have ? lane : 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t off = 16; off >= 1; off >>= 1) {
      float od = __shfl_xor_sync(kFullMask, bd, off);
      uint64_t orow = __shfl_xor_sync(kFullMask, br, off);
      uint32_t ol = __shfl_xor_sync(kFullMask, bl, off);
      bool take = (ol != 0xFFFFFFFFu) && (bl == 0xFFFFFFFFu || od < bd || (od == bd && orow < br) ||
                                          (od == bd && orow == br && ol < bl));
      if (take) {
        bd = od;
        br = orow;
        bl = ol;
      }
    }
    if (lane == bl) head += 1;
    if (lane == 0) {
      out_rows[(size_t)q * k + produced] = br;
      out_dist[(size_t)q * k + produced] = bd;
    }
  }
  for (uint32_t i = produced + lane; i < k; i += 32) {
    out_rows[(size_t)q * k + i] = 0xFFFFFFFFFFFFFFFFull;
    out_dist[(size_t)q * k + i] = INFINITY;
  }
  if (lane == 0) out_counts[q] = produced;
}

}  // namespace turdb
