// Experiment (not product code): tcgen05.mma.kind::tf32 with MN-MAJOR operands fed by TMA (SWIZZLE_128B).
// Weight gradients are dW[n][c] = sum_pos dY[pos][n] * X[pos][c]: in channels-last memory both operands have the contraction
// index (pos) as the SLOW dimension, i.e. they are MN-major.  Round 1 transposed both with extra kernels (6 % of the step).
// This probes which (instruction-descriptor major bits, LBO, SBO) combination reads such tiles correctly:
//   A^T in global: [K = 32 rows][M = 128 contiguous], B^T in global: [K = 32][N = 64 contiguous], small-integer data.
//   smem image by TMA: 3-D box [32 fl][K rows][M/32 groups] -> group g at g*Kbox*128 B, row k at k*128 B, 128-byte swizzle.
// For every variant the 128 x 64 result is compared with a CUDA-core evaluation; 0 mismatches = that encoding is right.
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
// generic descriptor: start | LBO | SBO | version 1 | swizzle mode 2 (128B)
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"(__cvta_generic_to_shared(bar))); }

constexpr int M = 128, N = 64, K = 32;
constexpr int NVAR = 8;

// variant v: bit0: (LBO, SBO) assignment for MN-major operands: 0 -> LBO = group stride (K*128), SBO = 1024 (8 k-rows)
//                                                                  1 -> LBO = 1024,               SBO = group stride
//            bit1: K advance per MMA (8 k-rows): 0 -> start += 1024 B;  1 -> start += 8*128 via SBO units (same) / alt: += 32 B (wrong on purpose, control)
//            bit2: which operands are MN-major: 0 -> A only (B K-major from a K-major copy), 1 -> both
__global__ void __launch_bounds__(128, 1) mn_kernel(const __grid_constant__ CUtensorMap mapAt, const __grid_constant__ CUtensorMap mapBt,
                                                    const __grid_constant__ CUtensorMap mapBk, const float* At, const float* Bt, int* mism) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sAt = sm;                   // 4 groups x 32 k x 128 B = 16 KB
  uint8_t* sBt = sm + 16384;           // 2 groups x 32 k x 128 B = 8 KB
  uint8_t* sBk = sm + 16384 + 8192;    // K-major B: 64 rows x 128 B = 8 KB
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
    mbar_expect_tx(&bar_tma, 16384 + 8192 + 8192);
    tma_load_3d(sAt, &mapAt, 0, 0, 0, &bar_tma);
    tma_load_3d(sBt, &mapBt, 0, 0, 0, &bar_tma);
    tma_load_3d(sBk, &mapBk, 0, 0, 0, &bar_tma);
  }
  mbar_wait(&bar_tma, 0);
  int phase = 0;
  for (int v = 0; v < NVAR; ++v) {
    const bool swapLS = v & 1, ctrl = v & 2, both = v & 4;
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      const uint32_t grp = K * 128;        // bytes between 32-element groups along M/N
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | (1u << 15);   // A MN-major
      if (both) idesc |= (1u << 16);                                                                                               // B MN-major
      for (int k = 0; k < K / 8; ++k) {
        const uint32_t kadv = ctrl ? (uint32_t)k * 32u : (uint32_t)k * 1024u;
        const uint64_t da = swapLS ? mk_desc(smem_u32(sAt) + kadv, 1024, grp) : mk_desc(smem_u32(sAt) + kadv, grp, 1024);
        uint64_t db;
        if (both) db = swapLS ? mk_desc(smem_u32(sBt) + kadv, 1024, grp) : mk_desc(smem_u32(sBt) + kadv, grp, 1024);
        else db = mk_desc(smem_u32(sBk) + (uint32_t)k * 32u, 16, 1024);          // known-good K-major encoding
        umma_tf32(tmem, da, db, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, phase & 1);
    ++phase;
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    int bad = 0;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      const int m = warp * 32 + lane;
      for (int e = 0; e < 8; ++e) {
        float ref = 0.f;
        for (int k = 0; k < K; ++k) ref += At[k * M + m] * Bt[k * N + c0 + e];
        if (__uint_as_float(r[e]) != ref) ++bad;
      }
    }
    if (bad) atomicAdd(&mism[v], bad);
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
  float *hAt = (float*)malloc(K * M * 4), *hBt = (float*)malloc(K * N * 4), *hBk = (float*)malloc(N * K * 4);
  srand(11);
  for (int i = 0; i < K * M; ++i) hAt[i] = (float)(rand() % 9 - 4);
  for (int i = 0; i < K * N; ++i) hBt[i] = (float)(rand() % 7 - 3);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hBk[n * K + k] = hBt[k * N + n];
  float *dAt, *dBt, *dBk; int* dM;
  CK(cudaMalloc(&dAt, K * M * 4)); CK(cudaMalloc(&dBt, K * N * 4)); CK(cudaMalloc(&dBk, N * K * 4)); CK(cudaMalloc(&dM, NVAR * 4));
  CK(cudaMemcpy(dAt, hAt, K * M * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBt, hBt, K * N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBk, hBk, N * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(dM, 0, NVAR * 4));
  CUtensorMap mAt, mBt, mBk;
  cuuint32_t es[3] = {1, 1, 1};
  {   // A^T [K][M]: dims (32 fl inner, K rows, M/32 groups): strides (M*4, 128)
    cuuint64_t gd[3] = {32, (cuuint64_t)K, (cuuint64_t)M / 32}, gs[2] = {(cuuint64_t)M * 4, 128};
    cuuint32_t bx[3] = {32, (cuuint32_t)K, (cuuint32_t)M / 32};
    printf("encode A^T: %d\n", (int)enc(&mAt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dAt, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  }
  {
    cuuint64_t gd[3] = {32, (cuuint64_t)K, (cuuint64_t)N / 32}, gs[2] = {(cuuint64_t)N * 4, 128};
    cuuint32_t bx[3] = {32, (cuuint32_t)K, (cuuint32_t)N / 32};
    printf("encode B^T: %d\n", (int)enc(&mBt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dBt, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  }
  {   // K-major B [N][K = 32 fl]
    cuuint64_t gd[3] = {32, (cuuint64_t)N, 1}, gs[2] = {128, (cuuint64_t)N * 128};
    cuuint32_t bx[3] = {32, (cuuint32_t)N, 1};
    printf("encode B K-major: %d\n", (int)enc(&mBk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dBk, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  }
  const int SMEM = 16384 + 8192 + 8192 + 1024;
  CK(cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  mn_kernel<<<1, 128, SMEM>>>(mAt, mBt, mBk, dAt, dBt, dM);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  int hM[NVAR];
  CK(cudaMemcpy(hM, dM, sizeof(hM), cudaMemcpyDeviceToHost));
  printf("mismatches out of 8192 (0 = correct):\n");
  for (int v = 0; v < NVAR; ++v)
    printf("  variant %d: %-5d  [%s | K-advance %s | MN-major: %s]\n", v, hM[v], (v & 1) ? "LBO=1024 SBO=group" : "LBO=group SBO=1024",
           (v & 2) ? "+32 B (control)" : "+1024 B", (v & 4) ? "A and B" : "A only");
  return 0;
}
