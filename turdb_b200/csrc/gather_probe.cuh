// gather_probe.cuh — diagnostics: the ceiling of the traversal kernel's access pattern.
//
// The layer-0 traversal gathers whole arena rows (dim*4 bytes, 512 B .. 3 KB) from uniformly random node ids
// with one cp.async.bulk each.  This kernel issues the SAME copies (same row size, same staging stride, same
// number of resident CTAs and staging slots per CTA) with NO dependency between them and no arithmetic:
// every staging group is re-armed the moment it lands.  Its GB/s is what the memory system gives this
// pattern at this footprint — the practical ceiling next to the streaming-copy peak of MEASURED_PEAKS.json.
#pragma once

#include "common.cuh"

namespace turdb {

struct GatherProbeArgs {
  const float* arena;
  uint64_t n;
  uint32_t ds;         // floats per arena row
  uint32_t vec_bytes;  // bytes copied per row
  uint32_t stride;     // bytes between staging slots
  uint32_t n_groups;   // groups of 8 slots, one mbarrier each
  uint32_t rounds;     // copies per slot
  uint32_t off_stage;
  uint32_t* sink;
};

__device__ __forceinline__ uint32_t probe_hash(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(128) gather_probe_kernel(const GatherProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const uint32_t bar0 = smem_u32(smem);
  const uint32_t stage = smem_u32(smem + a.off_stage);
  if (threadIdx.x == 0) {
    for (uint32_t g = 0; g < a.n_groups; ++g) mbar_init(bar0 + 8 * g, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](uint32_t g, uint32_t r) {
    if (lane == 0) mbar_expect_tx(bar0 + 8 * g, 8 * a.vec_bytes);
    __syncwarp();
    if (lane < 8) {
      const uint32_t h = probe_hash((blockIdx.x * 64u + g * 8u + lane) * 0x9E3779B1u + r * 0x85EBCA6Bu);
      const uint64_t id = ((uint64_t)h * a.n) >> 32;
      bulk_g2s(stage + (g * 8 + lane) * a.stride, a.arena + id * a.ds, a.vec_bytes, bar0 + 8 * g);
    }
  };
  for (uint32_t g = warp; g < a.n_groups; g += W) issue(g, 0);
  uint32_t acc = 0;
  for (uint32_t r = 0; r < a.rounds; ++r) {
    for (uint32_t g = warp; g < a.n_groups; g += W) {
      mbar_wait(bar0 + 8 * g, r & 1u);
      // one word per slot keeps the copies observable
      if (lane < 8) acc += *reinterpret_cast<const uint32_t*>(smem + a.off_stage + (g * 8 + lane) * a.stride);
      __syncwarp();
      if (r + 1 < a.rounds) issue(g, r + 1);
    }
  }
  if (acc == 0x12345678u) a.sink[0] = acc;
}

}  // namespace turdb
