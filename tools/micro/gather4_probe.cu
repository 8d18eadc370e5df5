// Micro-probe: semantics of TMA tile::gather4 (cp.async.bulk.tensor.2d ... tile::gather4) on sm_100a.
// Matrix M[rows=64][cols=48] with M[r][c] = r * 1000 + c.  Tensor maps with box {W, 1} and {W, 4} are tried; four rows
// {5, 17, 3, 60} are gathered at column 8 and the shared-memory image is printed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, uint32_t bytes, int c0, int r0, int r1, int r2, int r3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2048; ++i) reinterpret_cast<float*>(smem)[i] = -1.f;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(&map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a) : "memory");
    uint32_t ok = 0;
    for (int spin = 0; spin < 2000000 && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a) : "memory");
    out[2048] = (float)ok;
    for (int i = 0; i < 2048; ++i) out[i] = reinterpret_cast<float*>(smem)[i];
  }
}

int main() {
  const int R = 64, C = 48;
  std::vector<float> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 1000.f + c;
  float *d, *d_out;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, 2049 * 4);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const int W = 24;  // box width in floats (96 B per row)
  for (int box_rows : {1}) {  // {W, 4} is an illegal instruction (probed)
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {%d,%d}: encode rc=%d\n", W, box_rows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    probe<<<1, 32, 16384>>>(map, d_out, 4 * W * 4, 8, 5, 17, 3, 60);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(2049);
    cudaMemcpy(o.data(), d_out, 2049 * 4, cudaMemcpyDeviceToHost);
    printf("  barrier completed: %d\n", (int)o[2048]);
    for (int row = 0; row < 5; ++row) {
      printf("  smem[%3d..]:", row * W);
      for (int i = 0; i < 6; ++i) printf(" %8.0f", o[row * W + i]);
      printf(" ... %8.0f\n", o[row * W + W - 1]);
    }
  }
  // out-of-bounds columns and rows: box reaching past the last column (c0 = 40: columns 48..63 do not exist) and a row
  // index == number of rows; does the barrier still receive the full box byte count, and what lands?
  {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {(cuuint32_t)W, 1};
    cuuint32_t estr[2] = {1, 1};
    enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    probe<<<1, 32, 16384>>>(map, d_out, 4 * W * 4, 40, 5, 64, 3, 60);
    cudaError_t e = cudaDeviceSynchronize();
    printf("OOB probe launch: %s\n", cudaGetErrorString(e));
    std::vector<float> o(2049);
    cudaMemcpy(o.data(), d_out, 2049 * 4, cudaMemcpyDeviceToHost);
    printf("  barrier completed with the full box byte count: %d\n", (int)o[2048]);
    for (int row = 0; row < 4; ++row) {
      printf("  smem[%3d..]:", row * W);
      for (int i = 0; i < 10; ++i) printf(" %6.0f", o[row * W + i]);
      printf("\n");
    }
  }
  return 0;
}
