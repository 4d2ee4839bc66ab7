// TMA-fed, warp-specialised, persistent tcgen05 GEMM (sm_100a):   D[M][N] (+)= epi( A[M][K] * B[N][K]^T )
// fp32 storage, TF32 tensor-core math, fp32 accumulation in TMEM.  Serves every dense contraction without taps:
// nn.Linear / 1x1 conv forward (A = activations rows, B = packed weight [N][C]), their data gradients
// (A = dY rows, B = packed transposed weight [C][N]) and -- through evk_gemm_tf32 with split-K -- weight gradients
// on pre-transposed operands.
//
//   warp 0 (one lane)  : TMA producer.  cp.async.bulk.tensor 2-D boxes [128 x 32 fl] (A) and [BN x 32 fl] (B), 128-byte
//                        swizzle, into a STAGES-deep shared-memory ring; full[s] mbarrier with expect_tx.
//   warp 1 (one lane)  : MMA issuer.  4 x tcgen05.mma.kind::tf32 (K = 8) per stage, descriptors advance 32 B inside the
//                        128-byte swizzle atom; tcgen05.commit -> empty[s] releases the stage, -> acc_full[a] after
//                        the last K block of a tile.
//   warps 2..5         : epilogue.  tcgen05.ld 32x32b.x32 from the accumulator (each warp owns its TMEM lane quadrant),
//                        bias / residual / activation, 128-byte row segments to global (or fp32 atomics for split-K),
//                        then arrive on acc_empty[a].
// Two TMEM accumulators (2 x BN columns) let the epilogue of tile i overlap the main loop of tile i+1; CTAs are
// persistent (one per SM) and walk the tile list n-fastest so that concurrent CTAs share the A row block in L2.
// Out-of-range rows / K tails are zero-filled by TMA, so no shape needs padding.
#include <cuda.h>

#include "evk_common.cuh"

namespace evk {
namespace {

constexpr int BM = 128, BK = 32;                    // 32 floats = one 128-byte swizzle row
constexpr int GT_THREADS = 192;

struct GemmP {
  float* d; int ldd;
  const float* bias; const float* res; int ldr;
  int M, N, K, act; float slope;
  int tiles_m, tiles_n, splits, kb_per_split;        // kb = K blocks of 32
  int atomic;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
// K-major, SWIZZLE_128B: start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"(__cvta_generic_to_shared(bar)));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ GemmP p) {
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TCOLS = 2 * BN;
  static_assert(TCOLS <= 512 && (TCOLS & (TCOLS - 1)) == 0, "TMEM columns");
  extern __shared__ uint8_t gsm_raw[];
  uint8_t* gsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float epi_s[4 * 32 * 33];                          // per-epilogue-warp transpose tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&mapA)));
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&mapB)));
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem_base = tmem_base_s;

  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total = tiles_mn * p.splits;
  const int kb_total = (p.K + BK - 1) / BK;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int sp = tile / tiles_mn, mn = tile - sp * tiles_mn;
        const int tm = mn / p.tiles_n, tn = mn - tm * p.tiles_n;
        const int kb0 = sp * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* sa = gsm + (size_t)s * STAGE_BYTES;
          tma_load_2d(sa, &mapA, kb * BK, tm * BM, &full[s]);
          tma_load_2d(sa + A_BYTES, &mapB, kb * BK, tn * BN, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = tf32, both K-major, N = BN, M = 128
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
        const int sp = tile / tiles_mn;
        const int kb0 = sp * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        const int a = tcount & 1;
        mbar_wait(&acc_empty[a], ((tcount >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n");
          const uint32_t sa = smem_u32(gsm + (size_t)s * STAGE_BYTES);
          const uint64_t da = sw128_desc(sa), db = sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)
            umma_tf32(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else {
    const int lq = warp & 3;                                    // TMEM lane quadrant this warp may read
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int sp = tile / tiles_mn, mn = tile - sp * tiles_mn;
      const int tm = mn / p.tiles_n, tn = mn - tm * p.tiles_n;
      const int a = tcount & 1;
      mbar_wait(&acc_full[a], (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n");
      // TMEM gives each lane one ROW (32 consecutive columns); a direct store would touch 32 different rows per
      // instruction.  Transpose through a padded per-warp smem tile so that every store instruction writes four full
      // 128-byte row segments (and bias / residual loads are coalesced the same way).
      float* tr = epi_s + (warp - 2) * (32 * 33);
      const int r_sub = lane >> 3, c4 = (lane & 7) * 4;
      const int row_base = tm * BM + lq * 32;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(a * BN + c0), v);
        const int n = tn * BN + c0;
        if (n >= p.N) continue;                                  // warp-uniform
#pragma unroll
        for (int e = 0; e < 32; ++e) tr[lane * 33 + e] = v[e];
        __syncwarp();
        const int nn = n + c4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool full4 = nn + 4 <= p.N;
        if (p.bias && !p.atomic) {
          if (full4 && ((reinterpret_cast<uintptr_t>(p.bias + nn) & 15) == 0)) bv = *reinterpret_cast<const float4*>(p.bias + nn);
          else { if (nn < p.N) bv.x = p.bias[nn]; if (nn + 1 < p.N) bv.y = p.bias[nn + 1]; if (nn + 2 < p.N) bv.z = p.bias[nn + 2]; if (nn + 3 < p.N) bv.w = p.bias[nn + 3]; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + r_sub, row = row_base + rl;
          if (row >= p.M || nn >= p.N) continue;
          float t[4] = {tr[rl * 33 + c4], tr[rl * 33 + c4 + 1], tr[rl * 33 + c4 + 2], tr[rl * 33 + c4 + 3]};
          float* dp = p.d + (size_t)row * p.ldd + nn;
          if (p.atomic) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (nn + e < p.N) atomicAdd(dp + e, t[e]);
            continue;
          }
          t[0] += bv.x; t[1] += bv.y; t[2] += bv.z; t[3] += bv.w;
          if (p.res) {
            const float* rp = p.res + (size_t)row * p.ldr + nn;
            if (full4 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
              const float4 rv = *reinterpret_cast<const float4*>(rp);
              t[0] += rv.x; t[1] += rv.y; t[2] += rv.z; t[3] += rv.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (nn + e < p.N) t[e] += rp[e];
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (p.act == EVK_ACT_LRELU) t[e] = t[e] > 0.f ? t[e] : t[e] * p.slope;
            else if (p.act == EVK_ACT_RELU) t[e] = fmaxf(t[e], 0.f);
            else if (p.act == EVK_ACT_TANH) t[e] = tanhf(t[e]);
          }
          if (full4 && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
            *reinterpret_cast<float4*>(dp) = make_float4(t[0], t[1], t[2], t[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (nn + e < p.N) dp[e] = t[e];
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TCOLS));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(f);
  }
  return fn;
}

// row-major [rows][K] fp32 with pitch ld (floats): box = [box_rows][32 floats], 128-byte swizzle
bool make_map(CUtensorMap* m, const float* base, long long rows, long long K, long long ld, int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_sm_count = 0;

template <int BN, int STAGES>
int launch_gemm(const float* A, int lda, const float* B, int ldb, GemmP& p, int splits, cudaStream_t st) {
  CUtensorMap ma, mb;
  if (!make_map(&ma, A, p.M, p.K, lda, BM) || !make_map(&mb, B, p.N, p.K, ldb, BN)) return 1;
  constexpr int STAGE_BYTES = (BM + BN) * BK * 4;
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
  auto kern = gemm_tma_kernel<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM) != cudaSuccess) return 1;
    attr_set = true;
  }
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  p.tiles_m = cdiv(p.M, BM);
  p.tiles_n = cdiv(p.N, BN);
  const int kb_total = cdiv(p.K, BK);
  splits = max(1, min(splits, kb_total));
  p.kb_per_split = cdiv(kb_total, splits);
  p.splits = cdiv(kb_total, p.kb_per_split);
  const long long total = (long long)p.tiles_m * p.tiles_n * p.splits;
  if (total <= 0) return EVK_OK;
  const int grid = (int)(total < (long long)g_sm_count ? total : (long long)g_sm_count);
  kern<<<grid, GT_THREADS, SMEM, st>>>(ma, mb, p);
  return check_launch("gemm_tma_kernel");
}

int run_gemm(const float* A, int lda, const float* B, int ldb, GemmP& p, int splits, cudaStream_t st) {
  if (p.N > 128) return launch_gemm<256, 4>(A, lda, B, ldb, p, splits, st);
  return launch_gemm<128, 6>(A, lda, B, ldb, p, splits, st);
}

}  // namespace

int g_backend_tma = 1;

// returns 0 on success, < 0 on error, 1 if this launch is not eligible (caller falls through to gconv_tc / mma.sync)
int gemm_tma_try(const evk_gconv_desc* d, cudaStream_t st) {
  if (!g_backend_tma) return 1;
  if (d->Q != 1 || d->is != 1 || d->os != 1 || d->o0 != 0 || d->P != 1 || d->H != 1 || d->off[0] != 0) return 1;
  if (d->in_len || d->out_len || d->J != d->Tin) return 1;
  if ((d->C % 4) || (d->ldx % 4) || (d->ldw % 4) || (d->ldy % 4) || (d->res && (d->ldr % 4))) return 1;
  if (((uintptr_t)d->x | (uintptr_t)d->w | (uintptr_t)d->y) & 15) return 1;
  const long long rows = (long long)d->Z * d->J;
  if (d->Z > 1 && (d->x_sb != (long long)d->Tin * d->ldx || d->y_sb != (long long)d->J * d->ldy || d->w_sb != 0 ||
                   (d->res && d->r_sb != (long long)d->J * d->ldr)))
    return 1;
  if (rows < 512 || d->C < 64 || d->N < 64 || rows > 0x7fffffff) return 1;       // small problems: the tap kernel's finer tiles win
  GemmP p{};
  p.d = d->y; p.ldd = d->ldy; p.bias = d->bias; p.res = d->res; p.ldr = d->ldr;
  p.M = (int)rows; p.N = d->N; p.K = d->C; p.act = d->act; p.slope = d->slope; p.atomic = 0;
  return run_gemm(d->x, d->ldx, d->w, d->ldw, p, 1, st);
}

}  // namespace evk

using namespace evk;

extern "C" int evk_set_backend_tma(int32_t on) { g_backend_tma = on ? 1 : 0; return EVK_OK; }

extern "C" int evk_gemm_tf32(const float* A, int32_t lda, const float* B, int32_t ldb, float* D, int32_t ldd, int32_t M, int32_t N,
                             int32_t K, const float* bias, const float* res, int32_t ldr, int32_t act, float slope, int32_t splits,
                             cudaStream_t st) {
  EVK_REQUIRE(M > 0 && N > 0 && K > 0, EVK_ERR_ARG, "gemm_tf32: empty problem");
  EVK_REQUIRE((lda % 4) == 0 && (ldb % 4) == 0 && ((((uintptr_t)A) | ((uintptr_t)B)) & 15) == 0, EVK_ERR_ARG,
              "gemm_tf32: operands must be 16-byte aligned with pitches that are multiples of 4 floats");
  EVK_REQUIRE(splits >= 1, EVK_ERR_ARG, "gemm_tf32: splits");
  GemmP p{};
  p.d = D; p.ldd = ldd; p.bias = bias; p.res = res; p.ldr = ldr; p.M = M; p.N = N; p.K = K; p.act = act; p.slope = slope;
  p.atomic = splits > 1 ? 1 : 0;
  int rc = run_gemm(A, lda, B, ldb, p, splits, st);
  EVK_REQUIRE(rc != 1, EVK_ERR_UNSUPPORTED, "gemm_tf32: cuTensorMapEncodeTiled unavailable or rejected the operand");
  return rc;
}
