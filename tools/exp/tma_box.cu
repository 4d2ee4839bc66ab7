// Experiment (not product code): layout + L2->SM throughput of two ways of staging a conv input slab with TMA.
//   (a) [rows x 32 floats] box, SWIZZLE_128B (what gemm_tma_kernel uses today: one box per tap)
//   (b) 3-D map over the same memory with dims (4 floats, rows, 16-byte chunks): box [4][rows][8] lands in shared memory as
//       [chunk][row][16 B] = the interleaved no-swizzle K-major layout in which a conv tap is a +16 B*shift descriptor offset
// Reports: whether (b) encodes, whether its shared-memory image is the expected one, and GB/s of each with 148 CTAs x 4-deep ring.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)); }
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(smem_u32(dst)),
               "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

constexpr int ROWS_BOX = 144, STAGES = 4, BOX_BYTES = ROWS_BOX * 128;

// mode 0: map (a): coords (k0 floats, row, 0); mode 1: map (b): coords (0, row, chunk0)
__global__ void __launch_bounds__(64, 1) tma_bw_kernel(const __grid_constant__ CUtensorMap map, int mode, int iters, int rows_total, int kblocks,
                                                       float* dump) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[STAGES];
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tiles = rows_total / 128 - 1;
    int issued = 0, done = 0;
    auto issue = [&](int it) {
      const int s = it % STAGES;
      const int t = (blockIdx.x * 7 + it / kblocks) % tiles, kb = it % kblocks;
      mbar_expect_tx(&full[s], BOX_BYTES);
      if (mode == 0) tma_load_3d(sm + s * BOX_BYTES, &map, kb * 32, t * 128, 0, &full[s]);
      else tma_load_3d(sm + s * BOX_BYTES, &map, 0, t * 128, kb * 8, &full[s]);
    };
    for (; issued < STAGES && issued < iters; ++issued) issue(issued);
    for (; done < iters; ++done) {
      mbar_wait(&full[done % STAGES], (done / STAGES) & 1);
      if (issued < iters) { issue(issued); ++issued; }
    }
  }
  __syncthreads();
  if (dump && blockIdx.x == 0)          // image of the LAST box that landed in stage (iters-1) % STAGES
    for (int i = threadIdx.x; i < BOX_BYTES / 4; i += blockDim.x) dump[i] = ((float*)(sm + ((iters - 1) % STAGES) * BOX_BYTES))[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)f;
  const int C = 128, ROWS = 16 * 2560;                    // 16 x 2560 positions x 128 channels fp32 = 21 MB (L2 resident)
  const size_t n = (size_t)ROWS * C;
  float* hX = (float*)malloc(n * 4);
  for (size_t i = 0; i < n; ++i) hX[i] = (float)(i % 65521);
  float *dX, *dDump;
  CK(cudaMalloc(&dX, n * 4)); CK(cudaMalloc(&dDump, BOX_BYTES));
  CK(cudaMemcpy(dX, hX, n * 4, cudaMemcpyHostToDevice));
  CUtensorMap ma, mb;
  {
    cuuint64_t gd[3] = {(cuuint64_t)C, (cuuint64_t)ROWS, 1}, gs[2] = {(cuuint64_t)C * 4, (cuuint64_t)ROWS * C * 4};
    cuuint32_t bx[3] = {32, ROWS_BOX, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dX, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (a) swizzle-128B [32 fl][%d rows]: %d\n", ROWS_BOX, (int)r);
  }
  int ok_b;
  {
    cuuint64_t gd[3] = {4, (cuuint64_t)ROWS, (cuuint64_t)C / 4}, gs[2] = {(cuuint64_t)C * 4, 16};
    cuuint32_t bx[3] = {4, ROWS_BOX, 8}, es[3] = {1, 1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dX, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (b) no-swizzle [4 fl][%d rows][8 chunks], strides (C*4, 16): %d\n", ROWS_BOX, (int)r);
    ok_b = (r == CUDA_SUCCESS);
  }
  const int SMEM = STAGES * BOX_BYTES + 1024;
  CK(cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  float* hD = (float*)malloc(BOX_BYTES);
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 1 && !ok_b) break;
    // layout check: one box at tile 0 (block 0, it = 0 -> t = 0, kb = 0)
    tma_bw_kernel<<<1, 64, SMEM>>>(mode ? mb : ma, mode, 1, ROWS, C / 32, dDump);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD, dDump, BOX_BYTES, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < ROWS_BOX; ++r)
      for (int c = 0; c < 32; ++c) {
        float want = hX[(size_t)r * C + c], got;
        if (mode == 0) { const int chunk = (c / 4) ^ (r & 7); got = hD[r * 32 + chunk * 4 + (c & 3)]; }
        else got = hD[((c / 4) * ROWS_BOX + r) * 4 + (c & 3)];
        bad += (got != want);
      }
    printf("mode %d layout mismatches: %d\n", mode, bad);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 4000;
    tma_bw_kernel<<<148, 64, SMEM>>>(mode ? mb : ma, mode, iters, ROWS, C / 32, nullptr);
    CK(cudaEventRecord(e0));
    tma_bw_kernel<<<148, 64, SMEM>>>(mode ? mb : ma, mode, iters, ROWS, C / 32, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = 148.0 * iters * BOX_BYTES;
    printf("mode %d: %.3f ms, %.1f GB/s total, %.1f B/clk/SM at 1.965 GHz\n", mode, ms, bytes / ms / 1e6, bytes / ms / 1e6 * 1e9 / 148 / 1.965e9);
  }
  return 0;
}
