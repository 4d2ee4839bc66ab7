// Experiment (not product code): can a tcgen05 SWIZZLE_128B K-major A descriptor start at an arbitrary ROW of a TMA-loaded
// slab (start address not 1024-byte aligned), and which `base_offset` value (descriptor bits 49..51) makes it correct?
// One CTA: TMA-load a 160-row x 32-float slab (128-byte swizzle), then for every (row shift s, base offset bo) run
//   D[128 x 64] = slab[s .. s+127][0..31] * W[64][32]^T   (4 x tcgen05.mma.kind::tf32, K = 8 each)
// and count mismatches against a CUDA-core evaluation (small-integer data: exact in TF32).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o slab_desc slab_desc.cu ; run: ./slab_desc
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(dst)),
               "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"(__cvta_generic_to_shared(bar))); }

constexpr int SLAB_ROWS = 160, BN = 64, NSHIFT = 24;

__global__ void __launch_bounds__(128, 1) slab_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                                                      const float* X, const float* W, int* mism /*[NSHIFT][8]*/, int kadv_mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* slab = sm;                                  // 160 x 128 B = 20480 B (1024-aligned)
  uint8_t* wt = sm + 20480;                            // 64 x 128 B = 8192 B (1024-aligned: 20480 = 20 x 1024)
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_tma, SLAB_ROWS * 128 + BN * 128);
    tma_load_2d(slab, &mapX, 0, 0, &bar_tma);
    tma_load_2d(wt, &mapW, 0, 0, &bar_tma);
  }
  mbar_wait(&bar_tma, 0);
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int phase = 0;
  for (int s = 0; s < NSHIFT; ++s)
    for (int bo = 0; bo < 8; ++bo) {
      if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const uint32_t a0 = smem_u32(slab) + (uint32_t)s * 128u;
        const uint64_t da = sw128_desc(a0, (uint32_t)bo), db = sw128_desc(smem_u32(wt), 0);
        for (int k = 0; k < 4; ++k) umma_tf32(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, k > 0 ? 1u : 0u);
        umma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, phase & 1);
      ++phase;
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      // each warp reads its lane quadrant: row = warp*32 + lane, 64 columns
      int bad = 0;
      for (int c0 = 0; c0 < BN; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int row = warp * 32 + lane;
        for (int e = 0; e < 8; ++e) {
          float ref = 0.f;
          for (int k = 0; k < 32; ++k) ref += X[(size_t)(row + s) * 32 + k] * W[(size_t)(c0 + e) * 32 + k];
          if (__uint_as_float(r[e]) != ref) ++bad;
        }
      }
      if (bad) atomicAdd(&mism[s * 8 + bo], bad);
      asm volatile("tcgen05.fence::before_thread_sync;\n");
      __syncthreads();
    }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(64));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)f;
  const int ROWS = 256;
  float *hX = (float*)malloc(ROWS * 32 * 4), *hW = (float*)malloc(BN * 32 * 4);
  srand(7);
  for (int i = 0; i < ROWS * 32; ++i) hX[i] = (float)(rand() % 9 - 4);
  for (int i = 0; i < BN * 32; ++i) hW[i] = (float)(rand() % 7 - 3);
  float *dX, *dW; int* dM;
  CK(cudaMalloc(&dX, ROWS * 32 * 4)); CK(cudaMalloc(&dW, BN * 32 * 4)); CK(cudaMalloc(&dM, NSHIFT * 8 * 4));
  CK(cudaMemcpy(dX, hX, ROWS * 32 * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, hW, BN * 32 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dM, 0, NSHIFT * 8 * 4));
  CUtensorMap mx, mw;
  cuuint64_t gd[2] = {32, (cuuint64_t)ROWS}, gs[1] = {128};
  cuuint32_t bx[2] = {32, SLAB_ROWS}, es[2] = {1, 1};
  if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode X failed\n"); return 1; }
  cuuint64_t gdw[2] = {32, BN};
  cuuint32_t bw[2] = {32, BN};
  if (enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dW, gdw, gs, bw, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode W failed\n"); return 1; }
  const int SMEM = 20480 + 8192 + 1024;
  CK(cudaFuncSetAttribute(slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  slab_kernel<<<1, 128, SMEM>>>(mx, mw, dX, dW, dM, 0);
  CK(cudaDeviceSynchronize());
  int hM[NSHIFT * 8];
  CK(cudaMemcpy(hM, dM, sizeof(hM), cudaMemcpyDeviceToHost));
  printf("mismatching elements of the 128x64 tile (0 = correct); rows = shift s, cols = base_offset 0..7; (s & 7) shown\n");
  for (int s = 0; s < NSHIFT; ++s) {
    printf("s=%2d (s&7=%d):", s, s & 7);
    for (int bo = 0; bo < 8; ++bo) printf(" %5d", hM[s * 8 + bo]);
    printf("\n");
  }
  return 0;
}
