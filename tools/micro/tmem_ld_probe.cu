// Micro-probe: does reading an accumulator out of TMEM (tcgen05.ld) cost tensor-pipe time on sm_100a, and does the
// load shape matter?  One CTA per SM (grid = #SMs so the clocks/power are the bench's), 18 warps like the exact kernel:
//   warp 1 lane 0 issues `n_mma` tcgen05.mma (M128 N256 K16, SS, zero operands) into TMEM columns 256..511;
//   warps 2..2+W-1 each read `n_ld` times 64 columns x 32 lanes (8 KB) of columns 0..255 with the given shape.
// mode 1 = MMAs only, 2 = loads only, 3 = both.  Prints cycles of the MMA stream and of the slowest load warp.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
#define REGS32(v)                                                                                                       \
  "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), \
      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), \
      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define OUT32                                                                                                          \
  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, " \
  "%25, %26, %27, %28, %29, %30, %31}, [%32];"
template <int SHAPE>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  if (SHAPE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " OUT32 : REGS32(v) : "r"(taddr) : "memory");
  if (SHAPE == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " OUT32 : REGS32(v) : "r"(taddr) : "memory");
  if (SHAPE == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " OUT32 : REGS32(v) : "r"(taddr) : "memory");
  if (SHAPE == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " OUT32 : REGS32(v) : "r"(taddr) : "memory");
}

template <int SHAPE>
__global__ void __launch_bounds__(576, 1) probe(int mode, int n_ld_warps, int n_mma, int n_ld, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  unsigned long long t0 = clock64();
  if (warp == 1 && lane == 0 && (mode & 1)) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 16384));
    for (int i = 0; i < n_mma; ++i) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base + 256),
          "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(idesc), "r"(1u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
  if (warp >= 2 && warp < 2 + (uint32_t)n_ld_warps && (mode & 2)) {
    const uint32_t quarter = warp & 3, cslice = ((warp - 2) >> 2) * 64;
    uint32_t acc = 0;
    for (int i = 0; i < n_ld; ++i) {
      uint32_t v[32], w[32];
      if (SHAPE == 0) {
        const uint32_t t = tmem_base + ((quarter * 32) << 16) + cslice;
        ld32<SHAPE>(t, v);
        ld32<SHAPE>(t + 32, w);
      } else {  // 16-lane shapes: 64 columns x 16 lanes per instruction, two lane halves
        const uint32_t t = tmem_base + ((quarter * 32) << 16) + cslice;
        ld32<SHAPE>(t, v);
        ld32<SHAPE>(t + (16u << 16), w);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
    }
    if (acc == 0x12345678u) out[15] = acc;  // keep the loads alive
    __syncwarp();
    if (blockIdx.x == 0 && lane == 0) atomicMax(out + 1, clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int SHAPE>
static void run(const char* name, int sms, unsigned long long* d_out) {
  cudaFuncSetAttribute(probe<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 24 * 64, n_ld = 64;  // 64 "tiles" of K = 384; one tile = 16 warps x 8 KB
  for (int warps : {4, 8, 16}) {
    unsigned long long h[3][2] = {};
    for (int mode = 1; mode <= 3; ++mode) {
      // loads per warp scaled so that every configuration reads 64 x 128 KB in total
      const int per_warp = n_ld * 16 / warps;
      cudaMemset(d_out, 0, 16 * 8);
      probe<SHAPE><<<sms, 576, 64 * 1024>>>(mode, warps, n_mma, per_warp, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
      cudaMemcpy(h[mode - 1], d_out, 16, cudaMemcpyDeviceToHost);
    }
    const double bytes = 64.0 * 128 * 1024;
    printf("%-10s %2d load warps | mma alone %7llu cyc (%.0f/MMA) | loads alone %7llu cyc (%.1f B/clk) | together: mma %7llu (%.0f/MMA), loads %7llu (%.1f B/clk)\n",
           name, warps, h[0][0], (double)h[0][0] / n_mma, h[1][1], bytes / (double)h[1][1], h[2][0], (double)h[2][0] / n_mma, h[2][1],
           bytes / (double)h[2][1]);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_out;
  cudaMalloc(&d_out, 16 * 8);
  run<0>("32x32b.x32", sms, d_out);
  run<1>("16x256b.x8", sms, d_out);
  run<2>("16x128b.x16", sms, d_out);
  run<3>("16x64b.x32", sms, d_out);
  return 0;
}
