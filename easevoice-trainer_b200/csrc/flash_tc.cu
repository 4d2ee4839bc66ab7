// Prefix-LM attention of the stage-1 AR GPT on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM):
// the same three kernels, arguments, mask and dropout stream as flash.cu (t2s_model.py:456-487, patched_mha_with_cache.py SDPA
// call), but every contraction -- S = Q K^T, O = P V, dP = dO V^T, dQ = dS K, dV = P^T dO, dK = dS^T Q -- is a
// tcgen05.mma.kind::tf32 issued by one thread, and the softmax warps read scores with tcgen05.ld (one TMEM lane = one row).
//
//   flash_tc_fwd : CTA = 128 queries (UMMA M) of one (b, h); streams 64-key tiles.   S [128 x 64] -> TMEM, P -> shared memory
//                  (A operand of the second product), O_tile = P V [128 x 32] -> TMEM, rescaled running sum in registers.
//   flash_tc_dq  : same tiling; S and dP in TMEM, dS -> shared memory, dQ accumulates in TMEM over the key tiles.
//   flash_tc_dkv : CTA = 128 keys (UMMA M); streams 32-query tiles.  S^T and dP^T in TMEM, P^T / dS^T -> shared memory,
//                  dV and dK accumulate in TMEM over the query tiles.
//
// Operand layout: every operand is K-major in the canonical interleaved (SWIZZLE_NONE) core-matrix layout gconv_tc.cu uses --
// for each 16-byte K chunk a panel [rows][16 B] (pitch == 4 mod 8 rows: conflict-free 16-byte writes).  TF32 operands must be
// K-major (an MN-major TF32 descriptor is not usable, profiles/r2_experiments.md section 3), so the operands whose contraction
// index is the token axis (V in P V, K in dS K, dO in P^T dO, Q in dS^T Q) are staged TRANSPOSED: global rows -> registers ->
// 4-byte shared-memory stores into [token chunk][d][16 B] panels (a per-thread rotation of the four components makes the
// stores bank-conflict free).  Operands whose contraction index is d are staged with cp.async straight into their panels.
// Raw fp32 operands are truncated to TF32 by the tensor core (-3.5e-4 relative per operand on average); every product here
// has two raw operands and is compensated by (1 + g_trunc_comp)^2 where it is read (as gemm_tma.cu does).
// Single-stage pipeline per CTA (stage -> MMA -> softmax -> MMA); two or three CTAs per SM overlap each other's phases.
#include "flash_common.cuh"

namespace evk {
namespace {

constexpr int QP = 132;      // panel pitch (rows) of 128-row operands
constexpr int KP = 68;       // panel pitch of 64-row operands
constexpr int TP = 36;       // panel pitch of 32-row operands (transposed operands: rows = d; 32-query tiles)

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait: a tensor-core commit arrives within microseconds; ~2 s of polling means a lost arrival (a bug), and a trap
// (reported as a launch failure by the next runtime call) is better than a kernel that spins forever.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
#ifdef FT_TESTWAIT
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#endif
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 26)) __trap();
  }
}
// SWIZZLE_NONE, K-major: [0,14) addr>>4 | [16,30) LBO>>4 (next 16-byte K chunk) | [32,46) SBO>>4 (next 8 rows) | version 1
__device__ __forceinline__ uint64_t panel_desc(uint32_t saddr, uint32_t pitch_rows) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(((pitch_rows * 16u) >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];\n" ::"l"(__cvta_generic_to_shared(bar)) : "memory");
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// K = 8 * steps: the A / B descriptors advance two 16-byte chunks (= 2 panels) per step
__device__ __forceinline__ void mma_panels(uint32_t tacc, uint32_t a_addr, uint32_t a_pitch, uint32_t b_addr, uint32_t b_pitch, int steps,
                                           uint32_t idesc, bool accumulate) {
  for (int k2 = 0; k2 < steps; ++k2)
    umma_tf32(tacc, panel_desc(a_addr + (uint32_t)(2 * k2) * a_pitch * 16u, a_pitch), panel_desc(b_addr + (uint32_t)(2 * k2) * b_pitch * 16u, b_pitch),
              idesc, (accumulate || k2 > 0) ? 1u : 0u);
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

#ifdef FT_PROFILE
// developer build only (tools/exp/ft_profile.py): cycle counters of the forward kernel's phases as thread 0 sees them, summed
// over all CTAs.  slots: 0 stage+sync, 1 issue S, 2 wait S, 3 prefetch issue, 4 tcgen05.ld S, 5 softmax + P stores, 6 fence+sync,
// 7 issue PV, 8 wait PV, 9 ld O + accumulate, 10 tiles, 11 kernel
__device__ unsigned long long g_ft_prof[16];
#define FT_DECL() unsigned long long ft_acc[12] = {0ull}; long long ft_t = clock64(); const long long ft_t0 = ft_t
#define FT_MARK(i) do { if (threadIdx.x == 0) { const long long ft_n = clock64(); ft_acc[i] += (unsigned long long)(ft_n - ft_t); ft_t = ft_n; } } while (0)
#define FT_INC(i) ft_acc[i] += 1ull
#define FT_FLUSH() do { if (threadIdx.x == 0) { ft_acc[11] = (unsigned long long)(clock64() - ft_t0); for (int ft_i = 0; ft_i < 12; ++ft_i) atomicAdd(&g_ft_prof[ft_i], ft_acc[ft_i]); } } while (0)
// v2 roles: 0 softmax waits S, 1 ld S + max + exp (+ dropout), 2 waits O_tile, 3 ld O + accumulate, 4 P stores + fence + arrive,
// 5 MMA waits kv_full, 6 issue S, 7 waits p_full, 8 issue P V, 9 producer waits kv_empty, 10 issue loads, 11 rotate + V^T stores (load
// latency), 12 cp.async wait + fence + arrive, 13 softmax tiles, 14 kernel, 15 producer tiles
#define FT2_DECL(cond) const bool ft2_on = (cond); long long ft2_t = clock64(); const long long ft2_t0 = ft2_t; unsigned long long ft2_acc[16] = {0ull}
#define FT2_MARK(i) do { if (ft2_on) { const long long ft2_n = clock64(); ft2_acc[i] += (unsigned long long)(ft2_n - ft2_t); ft2_t = ft2_n; } } while (0)
#define FT2_INC(i) ft2_acc[i] += 1ull
#define FT2_FLUSH(lo, hi, tot) do { if (ft2_on) { if ((tot) >= 0) ft2_acc[(tot) & 15] = (unsigned long long)(clock64() - ft2_t0); for (int ft_i = (lo); ft_i <= (hi); ++ft_i) atomicAdd(&g_ft_prof[ft_i], ft2_acc[ft_i]); if ((tot) >= 0) atomicAdd(&g_ft_prof[(tot) & 15], ft2_acc[(tot) & 15]); } } while (0)
#else
#define FT_DECL()
#define FT_MARK(i)
#define FT_INC(i)
#define FT_FLUSH()
#define FT2_DECL(cond)
#define FT2_MARK(i)
#define FT2_INC(i)
#define FT2_FLUSH(lo, hi, tot)
#endif

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "n"(COLS) : "memory");
}

// rows [r0, r0 + ROWS) x 32 floats of a [.., ld] matrix (zero past L) -> K-major panels (contraction index = d): cp.async
template <int ROWS, int PITCH>
__device__ __forceinline__ void stage_rows(uint8_t* s, const float* g, int ld, int r0, int L) {
  for (int c = threadIdx.x; c < ROWS * 8; c += 128) {
    const int r = c >> 3, j = c & 7, row = r0 + r;
    cp_async16(s + ((size_t)j * PITCH + r) * 16, g + (size_t)(row < L ? row : 0) * ld + j * 4, row < L ? 16 : 0);
  }
}
// The same rows as a TRANSPOSED operand (contraction index = the row / token axis): panel = token chunk (4 tokens), panel row =
// d, pitch TP.  Thread (kk = idx & 3, dq = (idx >> 2) & 7, chunk = idx >> 5) loads the float4 d = 4 dq .. 4 dq + 3 of token
// 4 chunk + kk and stores its four components to four panel rows; component order rotated by dq >> 1 so that the 32 lanes of
// a store hit 32 different banks (bank = 16 (dq & 1) + 4 ((e + (dq >> 1)) & 3) + kk).
template <int ROWS>
__device__ __forceinline__ void load_t(float4 (&v)[ROWS / 16], const float* g, int ld, int r0, int L) {
#pragma unroll
  for (int it = 0; it < ROWS / 16; ++it) {
    const int idx = threadIdx.x + 128 * it;
    const int kk = idx & 3, dq = (idx >> 2) & 7, row = r0 + 4 * (idx >> 5) + kk;
    v[it] = row < L ? __ldg(reinterpret_cast<const float4*>(g + (size_t)row * ld + 4 * dq)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int ROWS>
__device__ __forceinline__ void store_t(uint8_t* s, const float4 (&v)[ROWS / 16]) {
  const int rot = (threadIdx.x >> 3) & 3;                           // dq >> 1 (idx and threadIdx.x agree modulo 32)
#pragma unroll
  for (int it = 0; it < ROWS / 16; ++it) {
    const int idx = threadIdx.x + 128 * it;
    const int kk = idx & 3, dq = (idx >> 2) & 7, ch = idx >> 5;
    const float a0 = v[it].x, a1 = v[it].y, a2 = v[it].z, a3 = v[it].w;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ee = (e + rot) & 3;
      const float val = ee == 0 ? a0 : (ee == 1 ? a1 : (ee == 2 ? a2 : a3));
      *reinterpret_cast<float*>(s + ((size_t)ch * TP + 4 * dq + ee) * 16 + kk * 4) = val;
    }
  }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float selp_f(float a, float b, bool c) {   // c ? a : b as ONE select (a nested ?: chain became divergent branches)
  float r;
  asm("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\nselp.f32 %0, %1, %2, p;\n}\n" : "=f"(r) : "f"(a), "f"(b), "r"((int)c));
  return r;
}
// rotate (a0..a3) left by rot (0..3) with two select stages
__device__ __forceinline__ void rot4(float (&a)[4], int rot) {
  const bool r0 = rot & 1, r1 = rot & 2;
  const float b0 = selp_f(a[1], a[0], r0), b1 = selp_f(a[2], a[1], r0), b2 = selp_f(a[3], a[2], r0), b3 = selp_f(a[0], a[3], r0);
  a[0] = selp_f(b2, b0, r1); a[1] = selp_f(b3, b1, r1); a[2] = selp_f(b0, b2, r1); a[3] = selp_f(b1, b3, r1);
}

// every (query i0 .. i0 + NQ - 1, key j0 .. j0 + NK - 1) pair visible?  (tile-level mask skip, as flash.cu)
template <int NK>
__device__ __forceinline__ bool tile_full_tc(int i0, int j0, int X, int xl, int yl) {
  const int j1 = j0 + NK - 1;
  if (j1 < X) return j1 < xl;
  if (j0 >= X) return ((j1 - X) < yl) & (j1 <= i0);
  return false;
}

// ---- shared-memory map of the three kernels (bytes; every panel block starts 128-byte aligned) ----------------------------------
constexpr int P128 = 8 * QP * 16;      // [128 rows][32 d] panels                       16 896
constexpr int P64 = 8 * KP * 16;       // [64 rows][32 d]                                8 704
constexpr int P32 = 8 * TP * 16;       // [32 rows][32 d] or transposed [32 d][32 tokens] 4 608
constexpr int T64 = 16 * TP * 16;      // transposed [32 d][64 tokens]                   9 216
constexpr int A64 = 16 * QP * 16;      // [128 rows][64 tokens] (P, dS)                 33 792
constexpr int A32 = 8 * QP * 16;       // [128 rows][32 tokens] (P^T, dS^T)             16 896
constexpr int FWD_SMEM = P128 + P64 + T64 + A64;                       // 68 608
constexpr int DQ_SMEM = 2 * P128 + 2 * P64 + T64 + A64;                // 94 208
constexpr int DKV_SMEM = 2 * P128 + 4 * P32 + 2 * A32;                 // 86 016

// ------------------------------------------------------------------------------------------------------------------------------
#ifdef FT_PROFILE
#define FT_MINB 2
#else
#define FT_MINB 3
#endif
__global__ void __launch_bounds__(128, FT_MINB) flash_tc_fwd_kernel(FlashArgs a, float comp2) {
  extern __shared__ __align__(128) uint8_t fsm[];
  uint8_t *sQ = fsm, *sK = sQ + P128, *sV = sK + P64, *sP = sV + T64;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, i0 = (int)(gridDim.x - 1 - blockIdx.x) * 128;      // long (late) query tiles first
  const int L = a.L, X = a.X;
  const int xl = (int)a.xlen[b], yl = (int)a.ylen[b];
  const size_t boff = (size_t)b * L * a.ld + (size_t)h * DK;
  const float *Q = a.q + boff, *K = a.k + boff, *V = a.v + boff;
  const DropKey dkey = drop_key(a);
  const float sl2 = a.scale * LOG2E * comp2;

  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  if (tid == 32) {
    mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // key tiles this query tile can see (flash.cu key_tiles with 128 query rows)
  int jend = max(X, min(i0 + 128, L));
  jend = min(jend, X + yl);
  jend = max(jend, min(X, L));
  const int nt = (jend + 63) >> 6;
  float4 vt[4];
  stage_rows<128, QP>(sQ, Q, a.ld, i0, L);
  if (nt > 0) {
    stage_rows<64, KP>(sK, K, a.ld, 0, L);
    load_t<64>(vt, V, a.ld, 0, L);
  }
  cp_async_commit();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem + ((uint32_t)(warp * 32) << 16), tO = tS + 64;
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);

  const int i = i0 + tid;                                            // this thread's query row = TMEM lane
  const uint32_t z = (uint32_t)(b * a.H + h);
  const uint32_t rowh = drop_row(dkey, z * (uint32_t)L + (uint32_t)i);
  float m = -INFINITY, l = 0.f;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;

  FT_DECL();
  for (int kt = 0; kt < nt; ++kt) {
    const int j0 = kt * 64;
    FT_MARK(11);
    FT_INC(10);
    store_t<64>(sV, vt);                                             // P V of the previous tile has completed (bar_o below)
    cp_async_wait<0>();
    fence_async_smem();
    fence_before();
    __syncthreads();
    FT_MARK(0);
    if (tid == 0) {
      fence_after();
      mma_panels(tmem, aQ, QP, aK, KP, 4, idesc_n(64), false);       // S = Q K^T
      umma_commit(&bar_s);
    }
    FT_MARK(1);
    mbar_wait(&bar_s, (uint32_t)(kt & 1));
    fence_after();
    FT_MARK(2);
    if (kt + 1 < nt) {                                               // the K panels are free again: next tile in flight under the softmax
      stage_rows<64, KP>(sK, K, a.ld, j0 + 64, L);
      load_t<64>(vt, V, a.ld, j0 + 64, L);
    }
    cp_async_commit();
    FT_MARK(3);
    float s[64];
    tmem_ld32(tS, s);
    tmem_ld32(tS + 32, s + 32);
    FT_MARK(4);
    float rmax = -INFINITY;
    if (tile_full_tc<64>(i0, j0, X, xl, yl)) {                       // CTA-uniform
#pragma unroll
      for (int c = 0; c < 64; ++c) rmax = fmaxf(rmax, s[c]);
    } else {
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        s[c] = allowed(i, j0 + c, X, xl, yl) ? s[c] : -INFINITY;
        rmax = fmaxf(rmax, s[c]);
      }
    }
    const float mx = fmaxf(m, rmax * sl2);
    const float e = (mx == -INFINITY) ? 0.f : mx;
    const float corr = ex2(m - e);                                   // m == -inf -> 0
    m = mx;
    float rs = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      float p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        p[u] = ex2(fmaf(s[4 * c4 + u], sl2, -e));
        rs += p[u];
      }
      if (dkey.thr) {
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
          bool k0, k1;
          drop_pair(dkey, rowh, j0 + 4 * c4 + u, k0, k1);
          p[u] = k0 ? p[u] * dkey.inv : 0.f;
          p[u + 1] = k1 ? p[u + 1] * dkey.inv : 0.f;
        }
      }
      *reinterpret_cast<float4*>(sP + ((size_t)c4 * QP + tid) * 16) = make_float4(p[0], p[1], p[2], p[3]);
    }
    l = l * corr + rs;
    FT_MARK(5);
    fence_async_smem();
    fence_before();
    __syncthreads();
    FT_MARK(6);
    if (tid == 0) {
      fence_after();
      mma_panels(tmem + 64, aP, QP, aV, TP, 8, idesc_n(32), false);  // O_tile = P V
      umma_commit(&bar_o);
    }
    FT_MARK(7);
    mbar_wait(&bar_o, (uint32_t)(kt & 1));
    fence_after();
    FT_MARK(8);
    float o[32];
    tmem_ld32(tO, o);
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = fmaf(acc[c], corr, o[c]);
    FT_MARK(9);
  }
  FT_FLUSH();
  if (i < L) {
    const float inv = l > 0.f ? comp2 / l : 0.f;
    float* O = a.o + ((size_t)b * L + i) * a.ldo + (size_t)h * DK;
#pragma unroll
    for (int c = 0; c < 32; c += 4)
      *reinterpret_cast<float4*>(O + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
    a.lse[(size_t)z * L + i] = m + log2f(l);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) {
    fence_after();
    tmem_free<128>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Forward, warp-specialised (v2).  The single-stage kernel above spends 40 % of a tile waiting for its own operand staging and
// the rest in one serial chain (tools/exp/ft_profile.py, profiles/r2_flash_tc.md); here the roles run concurrently:
//   warps 0-3 / 4-7 : softmax groups A / B, one query row per thread (CTA = 256 queries = two 128-row UMMA tiles sharing every
//                     K / V stage).  S(kt) arrives in one of two TMEM buffers, so the tensor core computes S(kt+1) during softmax(kt).
//   warps 8-11      : producers; warp w fills ring stage w with tiles kt = w, w+4, ...: K rows by cp.async, V transposed through
//                     registers, 16 + 16 loads in flight per lane.
//   warps 12 / 13   : MMA issue for group A / B (warp-uniform loops, one elected lane issues -- as gemm_tma.cu).
// mbarriers: kv_full[s] (32 producer lanes) / kv_empty[s] (2 commits), s_full[g][2] (commit), p_full[g] (128 softmax threads: P
// written, S and O_tile of the previous tile consumed), o_full[g] (commit: O_tile ready, P buffer free).
constexpr int NS2 = 4;
constexpr int STAGE2 = P64 + T64;                                   // 17 920
constexpr int FWD2_SMEM = 2 * P128 + NS2 * STAGE2 + 2 * A64;        // 173 056
constexpr int FWD2_THREADS = 448;

__global__ void __launch_bounds__(FWD2_THREADS, 1) flash_tc_fwd2_kernel(FlashArgs a, float comp2) {
  extern __shared__ __align__(128) uint8_t fsm[];
  uint8_t* sQ = fsm;                                                // [2][P128]
  uint8_t* ring = sQ + 2 * P128;                                    // [NS2][K panels P64 | V^T panels T64]
  uint8_t* sP = ring + NS2 * STAGE2;                                // [2][A64]
  __shared__ __align__(8) uint64_t q_full, kv_full[NS2], kv_empty[NS2], s_full[2][2], p_full[2], o_full[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, i0 = (int)(gridDim.x - 1 - blockIdx.x) * 256;
  const int L = a.L, X = a.X;
  const int xl = (int)a.xlen[b], yl = (int)a.ylen[b];
  const size_t boff = (size_t)b * L * a.ld + (size_t)h * DK;
  const float *Q = a.q + boff, *K = a.k + boff, *V = a.v + boff;

  if (tid == 0) {
    mbar_init(&q_full, 128);
    for (int s = 0; s < NS2; ++s) { mbar_init(&kv_full[s], 32); mbar_init(&kv_empty[s], 2); }
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g][0], 1); mbar_init(&s_full[g][1], 1); mbar_init(&p_full[g], 128); mbar_init(&o_full[g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 12) tmem_alloc<512>(&tmem_slot);
  int jend = max(X, min(i0 + 256, L));
  jend = min(jend, X + yl);
  jend = max(jend, min(X, L));
  const int nt = (jend + 63) >> 6;
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ---------------- softmax groups ----------------
    const int g = warp >> 2, r = (warp & 3) * 32 + lane;             // row of the group's 128-row tile = TMEM lane
    const int i = i0 + g * 128 + r;
    const uint32_t tg = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 192);     // S0 +0, S1 +64, O_tile +128
    uint8_t* myP = sP + g * A64;
    const DropKey dkey = drop_key(a);
    const float sl2 = a.scale * LOG2E * comp2;
    const uint32_t z = (uint32_t)(b * a.H + h);
    const uint32_t rowh = drop_row(dkey, z * (uint32_t)L + (uint32_t)i);
    const int ig0 = i0 + g * 128;                                    // first row of the group: tile-level mask test
    float m = -INFINITY, l = 0.f, corr_prev = 0.f;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    FT2_DECL(tid == 0);
    for (int kt = 0; kt < nt; ++kt) {
      const int j0 = kt * 64;
      FT2_MARK(14);
      FT2_INC(13);
      mbar_wait(&s_full[g][kt & 1], (uint32_t)((kt >> 1) & 1));
      fence_after();
      FT2_MARK(0);
      float s[64];
      tmem_ld32(tg + (uint32_t)((kt & 1) * 64), s);
      tmem_ld32(tg + (uint32_t)((kt & 1) * 64 + 32), s + 32);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (tile_full_tc<64>(ig0, j0, X, xl, yl)) {                    // warp-group uniform
#pragma unroll
        for (int c = 0; c < 64; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], s[c]);
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          s[c] = allowed(i, j0 + c, X, xl, yl) ? s[c] : -INFINITY;
          mx4[c & 3] = fmaxf(mx4[c & 3], s[c]);
        }
      }
      const float rmax = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float mx = fmaxf(m, rmax * sl2);
      const float e = (mx == -INFINITY) ? 0.f : mx;
      const float corr = ex2(m - e);
      m = mx;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        s[c] = ex2(fmaf(s[c], sl2, -e));
        rs4[c & 3] += s[c];
      }
      l = l * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      if (dkey.thr) {
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          bool k0, k1;
          drop_pair(dkey, rowh, j0 + c, k0, k1);
          s[c] = k0 ? s[c] * dkey.inv : 0.f;
          s[c + 1] = k1 ? s[c + 1] * dkey.inv : 0.f;
        }
      }
      FT2_MARK(1);
      if (kt > 0) {                                                  // O_tile(kt-1) ready, P buffer free
        mbar_wait(&o_full[g], (uint32_t)((kt - 1) & 1));
        fence_after();
        FT2_MARK(2);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float o[16];
          {
            uint32_t rr[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
                           "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
                         : "r"(tg + 128u + (uint32_t)(16 * hf)));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = __uint_as_float(rr[c]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[16 * hf + c] = fmaf(acc[16 * hf + c], corr_prev, o[c]);
        }
      }
      FT2_MARK(3);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4)
        *reinterpret_cast<float4*>(myP + ((size_t)c4 * QP + r) * 16) = make_float4(s[4 * c4], s[4 * c4 + 1], s[4 * c4 + 2], s[4 * c4 + 3]);
      corr_prev = corr;
      fence_async_smem();
      fence_before();
      mbar_arrive(&p_full[g]);
      FT2_MARK(4);
    }
    FT2_FLUSH(0, 4, 14);
#ifdef FT_PROFILE
    if (ft2_on) atomicAdd(&g_ft_prof[13], ft2_acc[13]);
#endif
    if (nt > 0) {
      mbar_wait(&o_full[g], (uint32_t)((nt - 1) & 1));
      fence_after();
      float o[32];
      tmem_ld32(tg + 128u, o);
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = fmaf(acc[c], corr_prev, o[c]);
    }
    if (i < L) {
      const float inv = l > 0.f ? comp2 / l : 0.f;
      float* O = a.o + ((size_t)b * L + i) * a.ldo + (size_t)h * DK;
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(O + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
      a.lse[(size_t)z * L + i] = m + log2f(l);
    }
  } else if (warp < 12) {
    // ---------------- producers ----------------
    const int pw = warp - 8, pt = tid - 256;
    for (int c = pt; c < 256 * 8; c += 128) {                        // both Q tiles
      const int rr = c >> 3, j = c & 7, row = i0 + rr;
      cp_async16(sQ + (size_t)(rr >> 7) * P128 + ((size_t)j * QP + (rr & 127)) * 16, Q + (size_t)(row < L ? row : 0) * a.ld + j * 4, row < L ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_async_smem();
    mbar_arrive(&q_full);
    const int rot = (lane >> 3) & 3;
    uint8_t* sKs = ring + (size_t)pw * STAGE2;
    uint8_t* sVs = sKs + P64;
    FT2_DECL(tid == 256);
    for (int kt = pw, n = 0; kt < nt; kt += NS2, ++n) {
      const int j0 = kt * 64;
      FT2_MARK(14);
      FT2_INC(15);
      if (n > 0) mbar_wait(&kv_empty[pw], (uint32_t)((n - 1) & 1));
      FT2_MARK(9);
      float4 v[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int c = lane + 32 * it, rr = c >> 3, j = c & 7, row = j0 + rr;
        cp_async16(sKs + ((size_t)j * KP + rr) * 16, K + (size_t)(row < L ? row : 0) * a.ld + j * 4, row < L ? 16 : 0);
      }
      cp_async_commit();
#pragma unroll
      for (int it = 0; it < 16; ++it) {                              // token chunk it: lane = (kk = lane & 3, dq = lane >> 2)
        const int row = j0 + 4 * it + (lane & 3);
        v[it] = row < L ? __ldg(reinterpret_cast<const float4*>(V + (size_t)row * a.ld + 4 * (lane >> 2))) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      FT2_MARK(10);
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        float x[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
        rot4(x, rot);                                                // x[e] = component (e + rot) & 3
#pragma unroll
        for (int e = 0; e < 4; ++e)
          *reinterpret_cast<float*>(sVs + ((size_t)it * TP + 4 * (lane >> 2) + ((e + rot) & 3)) * 16 + (lane & 3) * 4) = x[e];
      }
      FT2_MARK(11);
      cp_async_wait<0>();
      fence_async_smem();
      mbar_arrive(&kv_full[pw]);
      FT2_MARK(12);
    }
    FT2_FLUSH(9, 12, -1);
#ifdef FT_PROFILE
    if (ft2_on) atomicAdd(&g_ft_prof[15], ft2_acc[15]);
#endif
  } else {
    // ---------------- MMA issue, one warp per group ----------------
    const int g = warp - 12;
    const bool leader = elect_one();
    const uint32_t aQ = smem_u32(sQ + (size_t)g * P128), aP = smem_u32(sP + (size_t)g * A64), aR = smem_u32(ring);
    const uint32_t tS = tmem + (uint32_t)(g * 192), tO = tS + 128;
    mbar_wait(&q_full, 0);
    fence_after();
    FT2_DECL(warp == 12 && leader);
    for (int kt = 0; kt <= nt; ++kt) {
      FT2_MARK(14);
      if (kt < nt) {
        const int st = kt % NS2;
        mbar_wait(&kv_full[st], (uint32_t)((kt / NS2) & 1));
        fence_after();
        FT2_MARK(5);
        if (leader) {
          mma_panels(tS + (uint32_t)((kt & 1) * 64), aQ, QP, aR + (uint32_t)(st * STAGE2), KP, 4, idesc_n(64), false);      // S = Q K^T
          umma_commit(&s_full[g][kt & 1]);
        }
        __syncwarp();
        FT2_MARK(6);
      }
      if (kt > 0) {
        const int t = kt - 1, st = t % NS2;
        mbar_wait(&p_full[g], (uint32_t)(t & 1));
        fence_after();
        FT2_MARK(7);
        if (leader) {
          mma_panels(tO, aP, QP, aR + (uint32_t)(st * STAGE2 + P64), TP, 8, idesc_n(32), false);                            // O_tile = P V
          umma_commit(&o_full[g]);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
        FT2_MARK(8);
      }
    }
    FT2_FLUSH(5, 8, -1);
  }
  fence_before();
  __syncthreads();
  if (warp == 12) {
    fence_after();
    tmem_free<512>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Forward, warp-specialised with TWO threads per query row (v3).  The cycle counters of fwd2 (profiles/r2_flash_tc.md) left the
// softmax thread as the critical path with two softmax warps per scheduler; here a row's 64 scores of a tile are split between
// two threads (warps w and w + 4 read the same TMEM lane quadrant, different columns; they exchange the row maximum through
// shared memory and a 64-thread named barrier, keep partial row sums and each accumulate 16 of the 32 output dimensions), the CTA
// is one 128-row UMMA tile again and two CTAs share an SM, so that 16 softmax warps are resident per SM and one CTA's prologue /
// drain overlaps the other's steady state.   warps 0-7 softmax, 8-10 producers (one ring stage each), 11 MMA issue.
constexpr int NS3 = 3;
constexpr int FWD3_SMEM = P128 + NS3 * STAGE2 + A64;                // 104 448
constexpr int FWD3_THREADS = 384;

__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(FWD3_THREADS, 2) flash_tc_fwd3_kernel(FlashArgs a, float comp2) {
  extern __shared__ __align__(128) uint8_t fsm[];
  uint8_t* sQ = fsm;
  uint8_t* ring = sQ + P128;                                        // [NS3][K panels P64 | V^T panels T64]
  uint8_t* sP = ring + NS3 * STAGE2;
  __shared__ __align__(8) uint64_t q_full, kv_full[NS3], kv_empty[NS3], s_full[2], p_full, o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float xch[2][2][128], xl[2][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, i0 = (int)(gridDim.x - 1 - blockIdx.x) * 128;
  const int L = a.L, X = a.X;
  const int xl_ = (int)a.xlen[b], yl = (int)a.ylen[b];
  const size_t boff = (size_t)b * L * a.ld + (size_t)h * DK;
  const float *Q = a.q + boff, *K = a.k + boff, *V = a.v + boff;

  if (tid == 0) {
    mbar_init(&q_full, 96);
    for (int s = 0; s < NS3; ++s) { mbar_init(&kv_full[s], 32); mbar_init(&kv_empty[s], 1); }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(&p_full, 256); mbar_init(&o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 11) tmem_alloc<256>(&tmem_slot);
  int jend = max(X, min(i0 + 128, L));
  jend = min(jend, X + yl);
  jend = max(jend, min(X, L));
  const int nt = (jend + 63) >> 6;
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 8) {
    // ---------------- softmax: row r, column half hc ----------------
    const int wq = warp & 3, hc = warp >> 2, r = wq * 32 + lane;
    const int i = i0 + r;
    const uint32_t tg = tmem + ((uint32_t)(wq * 32) << 16);         // S0 +0, S1 +64, O_tile +128
    const DropKey dkey = drop_key(a);
    const float sl2 = a.scale * LOG2E * comp2;
    const uint32_t z = (uint32_t)(b * a.H + h);
    const uint32_t rowh = drop_row(dkey, z * (uint32_t)L + (uint32_t)i);
    float m = -INFINITY, l = 0.f, corr_prev = 0.f;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
    FT2_DECL(tid == 0);
    for (int kt = 0; kt < nt; ++kt) {
      const int j0 = kt * 64 + hc * 32;                              // first key of this thread's half tile
      FT2_MARK(14);
      FT2_INC(13);
      mbar_wait(&s_full[kt & 1], (uint32_t)((kt >> 1) & 1));
      fence_after();
      FT2_MARK(0);
      float s[32];
      tmem_ld32(tg + (uint32_t)((kt & 1) * 64 + hc * 32), s);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (tile_full_tc<64>(i0, kt * 64, X, xl_, yl)) {                // CTA-uniform
#pragma unroll
        for (int c = 0; c < 32; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], s[c]);
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          s[c] = allowed(i, j0 + c, X, xl_, yl) ? s[c] : -INFINITY;
          mx4[c & 3] = fmaxf(mx4[c & 3], s[c]);
        }
      }
      float rmax = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      xch[kt & 1][hc][r] = rmax;
      pair_sync(1 + wq);
      rmax = fmaxf(rmax, xch[kt & 1][hc ^ 1][r]);
      const float mx = fmaxf(m, rmax * sl2);
      const float e = (mx == -INFINITY) ? 0.f : mx;
      const float corr = ex2(m - e);
      m = mx;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        s[c] = ex2(fmaf(s[c], sl2, -e));
        rs4[c & 3] += s[c];
      }
      l = l * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      if (dkey.thr) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          bool k0, k1;
          drop_pair(dkey, rowh, j0 + c, k0, k1);
          s[c] = k0 ? s[c] * dkey.inv : 0.f;
          s[c + 1] = k1 ? s[c + 1] * dkey.inv : 0.f;
        }
      }
      FT2_MARK(1);
      if (kt > 0) {                                                  // O_tile(kt-1) ready, P buffer free
        mbar_wait(&o_full, (uint32_t)((kt - 1) & 1));
        fence_after();
        FT2_MARK(2);
        uint32_t rr[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
                       "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
                     : "r"(tg + 128u + (uint32_t)(16 * hc)));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(acc[c], corr_prev, __uint_as_float(rr[c]));
      }
      FT2_MARK(3);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(sP + ((size_t)(hc * 8 + c4) * QP + r) * 16) = make_float4(s[4 * c4], s[4 * c4 + 1], s[4 * c4 + 2], s[4 * c4 + 3]);
      corr_prev = corr;
      fence_async_smem();
      fence_before();
      mbar_arrive(&p_full);
      FT2_MARK(4);
    }
    FT2_FLUSH(0, 4, 14);
#ifdef FT_PROFILE
    if (ft2_on) atomicAdd(&g_ft_prof[13], ft2_acc[13]);
#endif
    if (nt > 0) {
      mbar_wait(&o_full, (uint32_t)((nt - 1) & 1));
      fence_after();
      uint32_t rr[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                   : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
                     "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
                   : "r"(tg + 128u + (uint32_t)(16 * hc)));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = fmaf(acc[c], corr_prev, __uint_as_float(rr[c]));
    }
    xl[hc][r] = l;                                                   // the two partial row sums share every running maximum
    pair_sync(1 + wq);
    l += xl[hc ^ 1][r];
    if (i < L) {
      const float inv = l > 0.f ? comp2 / l : 0.f;
      float* O = a.o + ((size_t)b * L + i) * a.ldo + (size_t)h * DK + 16 * hc;
#pragma unroll
      for (int c = 0; c < 16; c += 4)
        *reinterpret_cast<float4*>(O + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
      if (hc == 0) a.lse[(size_t)z * L + i] = m + log2f(l);
    }
  } else if (warp < 11) {
    // ---------------- producers: warp pw owns ring stage pw ----------------
    const int pw = warp - 8, pt = tid - 256;
    for (int c = pt; c < 128 * 8; c += 96) {
      const int rr = c >> 3, j = c & 7, row = i0 + rr;
      cp_async16(sQ + ((size_t)j * QP + rr) * 16, Q + (size_t)(row < L ? row : 0) * a.ld + j * 4, row < L ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_async_smem();
    mbar_arrive(&q_full);
    const int rot = (lane >> 3) & 3;
    uint8_t* sKs = ring + (size_t)pw * STAGE2;
    uint8_t* sVs = sKs + P64;
    FT2_DECL(tid == 256);
    for (int kt = pw, n = 0; kt < nt; kt += NS3, ++n) {
      const int j0 = kt * 64;
      FT2_MARK(14);
      FT2_INC(15);
      if (n > 0) mbar_wait(&kv_empty[pw], (uint32_t)((n - 1) & 1));
      FT2_MARK(9);
      float4 v[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int c = lane + 32 * it, rr = c >> 3, j = c & 7, row = j0 + rr;
        cp_async16(sKs + ((size_t)j * KP + rr) * 16, K + (size_t)(row < L ? row : 0) * a.ld + j * 4, row < L ? 16 : 0);
      }
      cp_async_commit();
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int row = j0 + 4 * it + (lane & 3);
        v[it] = row < L ? __ldg(reinterpret_cast<const float4*>(V + (size_t)row * a.ld + 4 * (lane >> 2))) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      FT2_MARK(10);
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        float x[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
        rot4(x, rot);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          *reinterpret_cast<float*>(sVs + ((size_t)it * TP + 4 * (lane >> 2) + ((e + rot) & 3)) * 16 + (lane & 3) * 4) = x[e];
      }
      FT2_MARK(11);
      cp_async_wait<0>();
      fence_async_smem();
      mbar_arrive(&kv_full[pw]);
      FT2_MARK(12);
    }
    FT2_FLUSH(9, 12, -1);
#ifdef FT_PROFILE
    if (ft2_on) atomicAdd(&g_ft_prof[15], ft2_acc[15]);
#endif
  } else {
    // ---------------- MMA issue ----------------
    const bool leader = elect_one();
    const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP), aR = smem_u32(ring);
    mbar_wait(&q_full, 0);
    fence_after();
    FT2_DECL(leader);
    for (int kt = 0; kt <= nt; ++kt) {
      FT2_MARK(14);
      if (kt < nt) {
        const int st = kt % NS3;
        mbar_wait(&kv_full[st], (uint32_t)((kt / NS3) & 1));
        fence_after();
        FT2_MARK(5);
        if (leader) {
          mma_panels(tmem + (uint32_t)((kt & 1) * 64), aQ, QP, aR + (uint32_t)(st * STAGE2), KP, 4, idesc_n(64), false);      // S = Q K^T
          umma_commit(&s_full[kt & 1]);
        }
        __syncwarp();
        FT2_MARK(6);
      }
      if (kt > 0) {
        const int t = kt - 1, st = t % NS3;
        mbar_wait(&p_full, (uint32_t)(t & 1));
        fence_after();
        FT2_MARK(7);
        if (leader) {
          mma_panels(tmem + 128u, aP, QP, aR + (uint32_t)(st * STAGE2 + P64), TP, 8, idesc_n(32), false);                     // O_tile = P V
          umma_commit(&o_full);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
        FT2_MARK(8);
      }
    }
    FT2_FLUSH(5, 8, -1);
  }
  fence_before();
  __syncthreads();
  if (warp == 11) {
    fence_after();
    tmem_free<256>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 2) flash_tc_dq_kernel(FlashArgs a, float comp2) {
  extern __shared__ __align__(128) uint8_t fsm[];
  uint8_t *sQ = fsm, *sD = sQ + P128, *sK = sD + P128, *sV = sK + P64, *sKT = sV + P64, *sS = sKT + T64;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, i0 = (int)(gridDim.x - 1 - blockIdx.x) * 128;
  const int L = a.L, X = a.X;
  const int xl = (int)a.xlen[b], yl = (int)a.ylen[b];
  const size_t boff = (size_t)b * L * a.ld + (size_t)h * DK;
  const float *Q = a.q + boff, *K = a.k + boff, *V = a.v + boff;
  const float* dO = a.dout + (size_t)b * L * a.lddo + (size_t)h * DK;
  const DropKey dkey = drop_key(a);
  const float sl2 = a.scale * LOG2E * comp2;

  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  if (tid == 32) {
    mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  int jend = max(X, min(i0 + 128, L));
  jend = min(jend, X + yl);
  jend = max(jend, min(X, L));
  const int nt = (jend + 63) >> 6;
  float4 kt4[4];
  stage_rows<128, QP>(sQ, Q, a.ld, i0, L);
  stage_rows<128, QP>(sD, dO, a.lddo, i0, L);
  if (nt > 0) {
    stage_rows<64, KP>(sK, K, a.ld, 0, L);
    stage_rows<64, KP>(sV, V, a.ld, 0, L);
    load_t<64>(kt4, K, a.ld, 0, L);
  }
  cp_async_commit();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem + ((uint32_t)(warp * 32) << 16), tP = tS + 64, tQ = tS + 128;
  const uint32_t aQ = smem_u32(sQ), aD = smem_u32(sD), aK = smem_u32(sK), aV = smem_u32(sV), aKT = smem_u32(sKT), aS = smem_u32(sS);

  const int i = i0 + tid;
  const uint32_t z = (uint32_t)(b * a.H + h);
  const uint32_t rowh = drop_row(dkey, z * (uint32_t)L + (uint32_t)i);
  const float lse_i = i < L ? a.lse[(size_t)z * L + i] : 0.f;
  const float del_i = i < L ? a.delta[(size_t)z * L + i] : 0.f;
  const float ginv = dkey.inv * comp2;                               // dP = (dO V^T) * comp2, then the dropout scale

  for (int kt = 0; kt < nt; ++kt) {
    const int j0 = kt * 64;
    if (kt > 0) {                                                    // dQ += dS K of the previous tile done: sS and sKT are free
      mbar_wait(&bar_o, (uint32_t)((kt - 1) & 1));
      fence_after();
    }
    store_t<64>(sKT, kt4);
    cp_async_wait<0>();
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (tid == 0) {
      fence_after();
      mma_panels(tmem, aQ, QP, aK, KP, 4, idesc_n(64), false);       // S = Q K^T
      mma_panels(tmem + 64, aD, QP, aV, KP, 4, idesc_n(64), false);  // dP = dO V^T
      umma_commit(&bar_s);
    }
    mbar_wait(&bar_s, (uint32_t)(kt & 1));
    fence_after();
    if (kt + 1 < nt) {
      stage_rows<64, KP>(sK, K, a.ld, j0 + 64, L);
      stage_rows<64, KP>(sV, V, a.ld, j0 + 64, L);
      load_t<64>(kt4, K, a.ld, j0 + 64, L);
    }
    cp_async_commit();
    const bool full = tile_full_tc<64>(i0, j0, X, xl, yl);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float s[32], dp[32];
      tmem_ld32(tS + 32 * hf, s);
      tmem_ld32(tP + 32 * hf, dp);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float ds[4];
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
          const int c = 4 * c4 + u, j = j0 + 32 * hf + c;
          float p0 = ex2(fmaf(s[c], sl2, -lse_i)), p1 = ex2(fmaf(s[c + 1], sl2, -lse_i));
          if (!full) {
            p0 = allowed(i, j, X, xl, yl) ? p0 : 0.f;
            p1 = allowed(i, j + 1, X, xl, yl) ? p1 : 0.f;
          }
          float g0 = dp[c] * ginv, g1 = dp[c + 1] * ginv;
          if (dkey.thr) {
            bool k0, k1;
            drop_pair(dkey, rowh, j, k0, k1);
            g0 = k0 ? g0 : 0.f;
            g1 = k1 ? g1 : 0.f;
          }
          ds[u] = p0 * (g0 - del_i);
          ds[u + 1] = p1 * (g1 - del_i);
        }
        *reinterpret_cast<float4*>(sS + ((size_t)(8 * hf + c4) * QP + tid) * 16) = make_float4(ds[0], ds[1], ds[2], ds[3]);
      }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (tid == 0) {
      fence_after();
      mma_panels(tmem + 128, aS, QP, aKT, TP, 8, idesc_n(32), kt > 0);   // dQ += dS K
      umma_commit(&bar_o);
    }
  }
  float dq[32];
  if (nt > 0) {
    mbar_wait(&bar_o, (uint32_t)((nt - 1) & 1));
    fence_after();
    tmem_ld32(tQ, dq);
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) dq[c] = 0.f;
  }
  if (i < L) {
    const float sc = a.scale * comp2;
    float* DQ = a.dq + ((size_t)b * L + i) * a.lddq + (size_t)h * DK;
#pragma unroll
    for (int c = 0; c < 32; c += 4)
      *reinterpret_cast<float4*>(DQ + c) = make_float4(dq[c] * sc, dq[c + 1] * sc, dq[c + 2] * sc, dq[c + 3] * sc);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) {
    fence_after();
    tmem_free<256>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 2) flash_tc_dkv_kernel(FlashArgs a, float comp2) {
  extern __shared__ __align__(128) uint8_t fsm[];
  uint8_t *sK = fsm, *sV = sK + P128, *sQ = sV + P128, *sD = sQ + P32, *sQT = sD + P32, *sDT = sQT + P32, *sPT = sDT + P32, *sST = sPT + A32;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float sL[32], sDl[32];
  __shared__ uint32_t sRh[32];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * 128;   // early key tiles (seen by the most queries) first
  const int L = a.L, X = a.X;
  const int xl = (int)a.xlen[b], yl = (int)a.ylen[b];
  const size_t boff = (size_t)b * L * a.ld + (size_t)h * DK;
  const float *Q = a.q + boff, *K = a.k + boff, *V = a.v + boff;
  const float* dO = a.dout + (size_t)b * L * a.lddo + (size_t)h * DK;
  const DropKey dkey = drop_key(a);
  const float sl2 = a.scale * LOG2E * comp2;
  const uint32_t z = (uint32_t)(b * a.H + h);
  const float* lse = a.lse + (size_t)z * L;
  const float* dl = a.delta + (size_t)z * L;

  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  if (tid == 32) {
    mbar_init(&bar_s, 1); mbar_init(&bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // 32-query tiles that can see this key tile: text keys are seen by every query, audio key j only by queries >= j
  const bool dead = (j0 >= X + yl) && (j0 >= X);                     // every key of the tile is padding: gradients are zero
  const int qt0 = (j0 < X) ? 0 : (j0 >> 5);
  const int qt1 = dead ? qt0 : (L + 31) >> 5;
  const int nt = qt1 - qt0;
  float4 qt4[2], dt4[2];
  if (nt > 0) {
    stage_rows<128, QP>(sK, K, a.ld, j0, L);
    stage_rows<128, QP>(sV, V, a.ld, j0, L);
    stage_rows<32, TP>(sQ, Q, a.ld, qt0 * 32, L);
    stage_rows<32, TP>(sD, dO, a.lddo, qt0 * 32, L);
    load_t<32>(qt4, Q, a.ld, qt0 * 32, L);
    load_t<32>(dt4, dO, a.lddo, qt0 * 32, L);
  }
  cp_async_commit();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem + ((uint32_t)(warp * 32) << 16), tP = tS + 32, tV = tS + 64, tK = tS + 96;
  const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aD = smem_u32(sD), aQT = smem_u32(sQT), aDT = smem_u32(sDT),
                 aPT = smem_u32(sPT), aST = smem_u32(sST);
  const int j = j0 + tid;                                            // this thread's key row = TMEM lane
  const uint32_t colterm = drop_col(dkey, j), colmul = (j & 1) ? DROP_M2 : DROP_M1;
  const float ginv = dkey.inv * comp2;

  for (int t = 0; t < nt; ++t) {
    const int i0 = (qt0 + t) * 32;
    if (t > 0) {                                                     // dV / dK products of the previous tile done: sPT, sST, sQT, sDT free
      mbar_wait(&bar_o, (uint32_t)((t - 1) & 1));
      fence_after();
    }
    store_t<32>(sQT, qt4);
    store_t<32>(sDT, dt4);
    if (tid < 32) {
      const int i = i0 + tid;
      sL[tid] = i < L ? lse[i] : 0.f;
      sDl[tid] = i < L ? dl[i] : 0.f;
    } else if (tid < 64) {
      sRh[tid - 32] = drop_row(dkey, z * (uint32_t)L + (uint32_t)(i0 + tid - 32));
    }
    cp_async_wait<0>();
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (tid == 0) {
      fence_after();
      mma_panels(tmem, aK, QP, aQ, TP, 4, idesc_n(32), false);       // S^T = K Q^T
      mma_panels(tmem + 32, aV, QP, aD, TP, 4, idesc_n(32), false);  // dP^T = V dO^T
      umma_commit(&bar_s);
    }
    mbar_wait(&bar_s, (uint32_t)(t & 1));
    fence_after();
    if (t + 1 < nt) {
      stage_rows<32, TP>(sQ, Q, a.ld, i0 + 32, L);
      stage_rows<32, TP>(sD, dO, a.lddo, i0 + 32, L);
      load_t<32>(qt4, Q, a.ld, i0 + 32, L);
      load_t<32>(dt4, dO, a.lddo, i0 + 32, L);
    }
    cp_async_commit();
    float s[32], dp[32];
    tmem_ld32(tS, s);
    tmem_ld32(tP, dp);
    const bool full = (i0 + 32 <= L) && tile_full_tc<128>(i0, j0, X, xl, yl);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      float pd[4], ds[4];
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        const int c = 4 * c4 + u, i = i0 + c;                         // queries (i, i + 1): one dropout hash
        float p0 = ex2(fmaf(s[c], sl2, -sL[c])), p1 = ex2(fmaf(s[c + 1], sl2, -sL[c + 1]));
        if (!full) {
          p0 = (i < L && allowed(i, j, X, xl, yl)) ? p0 : 0.f;
          p1 = (i + 1 < L && allowed(i + 1, j, X, xl, yl)) ? p1 : 0.f;
        }
        float g0 = dp[c] * ginv, g1 = dp[c + 1] * ginv;
        float q0 = p0 * dkey.inv, q1 = p1 * dkey.inv;
        if (dkey.thr) {
          const bool k0 = drop_one(dkey, sRh[c], colterm, colmul), k1 = drop_one(dkey, sRh[c + 1], colterm, colmul);
          g0 = k0 ? g0 : 0.f; q0 = k0 ? q0 : 0.f;
          g1 = k1 ? g1 : 0.f; q1 = k1 ? q1 : 0.f;
        }
        pd[u] = q0; pd[u + 1] = q1;
        ds[u] = p0 * (g0 - sDl[c]);
        ds[u + 1] = p1 * (g1 - sDl[c + 1]);
      }
      *reinterpret_cast<float4*>(sPT + ((size_t)c4 * QP + tid) * 16) = make_float4(pd[0], pd[1], pd[2], pd[3]);
      *reinterpret_cast<float4*>(sST + ((size_t)c4 * QP + tid) * 16) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (tid == 0) {
      fence_after();
      mma_panels(tmem + 64, aPT, QP, aDT, TP, 4, idesc_n(32), t > 0);    // dV += Pd^T dO
      mma_panels(tmem + 96, aST, QP, aQT, TP, 4, idesc_n(32), t > 0);    // dK += dS^T Q
      umma_commit(&bar_o);
    }
  }
  float dv[32], dk[32];
  if (nt > 0) {
    mbar_wait(&bar_o, (uint32_t)((nt - 1) & 1));
    fence_after();
    tmem_ld32(tV, dv);
    tmem_ld32(tK, dk);
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) dv[c] = dk[c] = 0.f;
  }
  if (j < L) {
    const float sk = a.scale * comp2;
    float* DK_ = a.dk + ((size_t)b * L + j) * a.lddq + (size_t)h * DK;
    float* DV_ = a.dv + ((size_t)b * L + j) * a.lddq + (size_t)h * DK;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      *reinterpret_cast<float4*>(DK_ + c) = make_float4(dk[c] * sk, dk[c + 1] * sk, dk[c + 2] * sk, dk[c + 3] * sk);
      *reinterpret_cast<float4*>(DV_ + c) = make_float4(dv[c] * comp2, dv[c + 1] * comp2, dv[c + 2] * comp2, dv[c + 3] * comp2);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) {
    fence_after();
    tmem_free<128>(tmem);
  }
}

template <typename Kern>
int set_smem(Kern kern, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  EVK_REQUIRE(e == cudaSuccess, EVK_ERR_CUDA, "flash_tc: cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e));
  return 0;
}

}  // namespace

int g_flash_tc = 0;            // 1: the tcgen05 kernels serve evk_flash_attn_fwd / _bwd (not in 3xTF32 test mode); 0: mma.sync kernels
float g_flash_comp = 3.52e-4f; // relative compensation per raw (truncated) TF32 operand, see gemm_tma.cu g_trunc_comp

static float comp2_now() { const float c = 1.f + g_flash_comp; return c * c; }

// -> 0 launched, < 0 error, 1 not eligible (caller runs the mma.sync kernel)
int flash_tc_fwd_try(const FlashArgs& a, cudaStream_t st) {
  if (!g_flash_tc || (a.ld % 4) || (a.ldo % 4) || ((uintptr_t)a.o % 16)) return 1;
  static bool attr = false;
  if (!attr) {
    if (int rc = set_smem(flash_tc_fwd_kernel, FWD_SMEM)) return rc;
    if (int rc = set_smem(flash_tc_fwd2_kernel, FWD2_SMEM)) return rc;
    if (int rc = set_smem(flash_tc_fwd3_kernel, FWD3_SMEM)) return rc;
    attr = true;
  }
  if (g_flash_tc >= 3) {
    dim3 grid3(cdiv(a.L, 128), a.H, a.B);
    flash_tc_fwd3_kernel<<<grid3, FWD3_THREADS, FWD3_SMEM, st>>>(a, comp2_now());
    return check_launch("flash_tc_fwd3");
  }
  if (g_flash_tc >= 2) {
    dim3 grid2(cdiv(a.L, 256), a.H, a.B);
    flash_tc_fwd2_kernel<<<grid2, FWD2_THREADS, FWD2_SMEM, st>>>(a, comp2_now());
    return check_launch("flash_tc_fwd2");
  }
  dim3 grid(cdiv(a.L, 128), a.H, a.B);
  flash_tc_fwd_kernel<<<grid, 128, FWD_SMEM, st>>>(a, comp2_now());
  return check_launch("flash_tc_fwd");
}

int flash_tc_bwd_try(const FlashArgs& a, cudaStream_t st) {
  if (!g_flash_tc || (a.ld % 4) || (a.lddo % 4) || (a.lddq % 4) || (((uintptr_t)a.dq | (uintptr_t)a.dk | (uintptr_t)a.dv) % 16)) return 1;
  static bool attr = false;
  if (!attr) {
    if (int rc = set_smem(flash_tc_dq_kernel, DQ_SMEM)) return rc;
    if (int rc = set_smem(flash_tc_dkv_kernel, DKV_SMEM)) return rc;
    attr = true;
  }
  dim3 grid(cdiv(a.L, 128), a.H, a.B);
  flash_tc_dq_kernel<<<grid, 128, DQ_SMEM, st>>>(a, comp2_now());
  if (int rc = check_launch("flash_tc_dq")) return rc;
  flash_tc_dkv_kernel<<<grid, 128, DKV_SMEM, st>>>(a, comp2_now());
  return check_launch("flash_tc_dkv");
}

}  // namespace evk

extern "C" int evk_set_flash_tc(int32_t on, float trunc_comp) {
  evk::g_flash_tc = on < 0 ? 0 : on;         // 0 off, 1 single-stage kernels, 2 warp-specialised forward
  if (trunc_comp >= 0.f) evk::g_flash_comp = trunc_comp;
  return 0;
}
extern "C" int evk_get_flash_tc(void) { return evk::g_flash_tc; }

#ifdef FT_PROFILE
extern "C" int evk_ft_prof_read(unsigned long long* host, int reset) {
  if (host) cudaMemcpyFromSymbol(host, evk::g_ft_prof, sizeof(evk::g_ft_prof));
  if (reset) { static unsigned long long z[16]; cudaMemcpyToSymbol(evk::g_ft_prof, z, sizeof(z)); }
  return 0;
}
#endif
