// Fused mel front end: reflect-pad -> Hann -> 2048-point real FFT -> sqrt(re^2+im^2+1e-6)
// -> sparse Slaney filterbank -> log(clamp(., 1e-5)), one CTA per frame, everything in shared memory.
// fp32 Stockham radix-4 (5 passes over a 1024-point complex FFT of the even/odd-packed frame) so the
// result is fp32-FFT accurate (the parity oracle is torch.stft in fp32).
// Reference: src/easevoice/module/mel_processing.py:40-142.
#include "evk_common.cuh"
#include <math.h>

namespace evk {

constexpr int NFFT = 2048;
constexpr int NH = NFFT / 2;      // complex FFT length
constexpr int NBIN = NH + 1;      // 1025

__device__ float2 g_tw[NFFT];     // e^{-2 pi i k / 2048}
__device__ float g_hann[NFFT];    // periodic Hann
__device__ float2 g_tw32[32 * 32]; // [k1][t] = e^{-2 pi i t k1 / 1024}: inter-stage twiddles of the 32 x 32 warp FFT
static bool g_tables_ready = false;

int mel_init_tables() {
  if (g_tables_ready) return EVK_OK;
  static float2 tw[NFFT];
  static float hw[NFFT];
  for (int i = 0; i < NFFT; ++i) {
    double a = 2.0 * M_PI * (double)i / (double)NFFT;
    tw[i] = make_float2((float)cos(a), (float)(-sin(a)));
    hw[i] = (float)(0.5 - 0.5 * cos(a));
  }
  if (cudaMemcpyToSymbol(g_tw, tw, sizeof(tw)) != cudaSuccess) return EVK_ERR_CUDA;
  if (cudaMemcpyToSymbol(g_hann, hw, sizeof(hw)) != cudaSuccess) return EVK_ERR_CUDA;
  static float2 t32[32 * 32];
  for (int k1 = 0; k1 < 32; ++k1)
    for (int t = 0; t < 32; ++t) {
      double a = 2.0 * M_PI * (double)((t * k1) % 1024) / 1024.0;
      t32[k1 * 32 + t] = make_float2((float)cos(a), (float)(-sin(a)));
    }
  if (cudaMemcpyToSymbol(g_tw32, t32, sizeof(t32)) != cudaSuccess) return EVK_ERR_CUDA;
  g_tables_ready = true;
  return EVK_OK;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ int reflect_idx(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

// spectrum of one frame into smem: X[k], k = 0..1024 written to sx (float2[1025]); uses d0,d1 as scratch
__device__ void frame_rfft(const float* __restrict__ wav, int L, int s0, float2* d0, float2* d1, float2* sx) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = tid + 256 * i;
    const int n0 = 2 * m, n1 = 2 * m + 1;
    const float a = wav[reflect_idx(s0 + n0, L)] * g_hann[n0];
    const float b = wav[reflect_idx(s0 + n1, L)] * g_hann[n1];
    d0[m] = make_float2(a, b);
  }
  __syncthreads();
  float2* src = d0;
  float2* dst = d1;
#pragma unroll
  for (int Ns = 1; Ns < NH; Ns *= 4) {
    const int j = tid;
    const int k = j & (Ns - 1);
    const int tstep = (NH / (4 * Ns)) * 2;          // index step in the 2048-entry table
    float2 v0 = src[j], v1 = src[j + NH / 4], v2 = src[j + NH / 2], v3 = src[j + 3 * NH / 4];
    if (Ns > 1) {
      v1 = cmul(v1, g_tw[k * tstep]);
      v2 = cmul(v2, g_tw[2 * k * tstep]);
      v3 = cmul(v3, g_tw[3 * k * tstep]);
    }
    const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 t = make_float2(v1.x - v3.x, v1.y - v3.y);
    const float2 a3 = make_float2(t.y, -t.x);        // (v1 - v3) * (-i)
    const int idx = (j / Ns) * Ns * 4 + k;
    dst[idx] = make_float2(a0.x + a2.x, a0.y + a2.y);
    dst[idx + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
    dst[idx + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
    dst[idx + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
    __syncthreads();
    float2* tmp = src; src = dst; dst = tmp;
  }
  // real-FFT untangle: X[k] = (Z[k] + conj(Z[N/2-k]))/2 - (i/2) e^{-2 pi i k/N} (Z[k] - conj(Z[N/2-k]))
  for (int k = tid; k <= NH; k += 256) {
    const float2 zk = src[k & (NH - 1)];
    float2 zc = src[(NH - k) & (NH - 1)];
    zc.y = -zc.y;
    const float2 s = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 dd = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
    const float2 w = g_tw[k];
    const float2 t = cmul(w, dd);                     // e^{-i th} * d
    sx[k] = make_float2(s.x + t.y, s.y - t.x);        // s - i*t
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) mel_fwd_kernel(const float* __restrict__ wav, int L, int ldw, int T, int hop,
                                                      int n_mels, const int* __restrict__ fb_ptr,
                                                      const int* __restrict__ fb_idx, const float* __restrict__ fb_val,
                                                      float* __restrict__ spec, int ld_spec, float* __restrict__ mel,
                                                      int ld_mel, float* __restrict__ cplx) {
  __shared__ float2 d0[NH], d1[NH];
  __shared__ float2 sx[NBIN + 3];
  __shared__ float mag[NBIN + 3];
  const int frame = blockIdx.x;
  const int b = frame / T, f = frame - b * T;
  const int pad = (NFFT - hop) / 2;
  frame_rfft(wav + (long long)b * ldw, L, f * hop - pad, d0, d1, sx);
  for (int k = threadIdx.x; k < NBIN; k += 256) {
    const float2 x = sx[k];
    const float m = sqrtf(x.x * x.x + x.y * x.y + 1e-6f);
    mag[k] = m;
    if (spec) spec[(long long)frame * ld_spec + k] = m;
    if (cplx) reinterpret_cast<float2*>(cplx)[(long long)frame * NBIN + k] = x;
  }
  __syncthreads();
  if (mel) {
    for (int m = threadIdx.x; m < n_mels; m += 256) {
      float acc = 0.f;
      for (int e = fb_ptr[m]; e < fb_ptr[m + 1]; ++e) acc = fmaf(fb_val[e], mag[fb_idx[e]], acc);
      mel[(long long)frame * ld_mel + m] = logf(fmaxf(acc, 1e-5f));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Warp-per-frame variant (the fast path): the 1024-point complex FFT of the even/odd packed frame is computed as
// 32 x 32 with all data in registers -- each lane runs a 32-point radix-2 FFT on its own registers (compile-time
// indices, compile-time twiddles), the warp transposes through shared memory once, each lane runs a second
// 32-point FFT.  No block-wide barrier, 8 frames per CTA, every global access coalesced.
// ------------------------------------------------------------------------------------------------------------
constexpr int MEL_WPB = 6;
constexpr int MEL_WARP_SMEM = 32 * 33 * 8 + (NBIN + 3) * 4;   // 12 560 B per warp
constexpr int MEL_FB_MAX = 2064;                                          // filterbank non-zeros staged in smem (2 016 for 128 Slaney mels)
constexpr int MEL_TAB_SMEM = 32 * 32 * 8 + (NH + 8) * 8 + MEL_FB_MAX * 8 + 144 * 4;   // twiddles (2 tables) + CSR val/idx/ptr

__host__ __device__ constexpr int brev5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// cos(2 pi k / 32), sin(2 pi k / 32), k = 0..15
__device__ constexpr float C32[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                      0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                                      0.19509032201612825f, 0.f, -0.19509032201612825f, -0.38268343236508977f,
                                      -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                                      -0.92387953251128674f, -0.98078528040323043f};
__device__ constexpr float S32[16] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                                      0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f,
                                      0.98078528040323043f, 1.f, 0.98078528040323043f, 0.92387953251128674f,
                                      0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                                      0.38268343236508977f, 0.19509032201612825f};

// in-place 32-point DIF FFT on registers; X[k] ends up in element brev5(k)
__device__ __forceinline__ void fft32_regs(float (&xr)[32], float (&xi)[32]) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
    for (int blk = 0; blk < 32; blk += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const int i0 = blk + j, i1 = i0 + half;
        const int tw = j * (16 / half);                       // exponent of W_32 (compile-time after unrolling)
        const float ar = xr[i0], ai = xi[i0], br = xr[i1], bi = xi[i1];
        xr[i0] = ar + br; xi[i0] = ai + bi;
        const float dr = ar - br, di = ai - bi;
        if (tw == 0) { xr[i1] = dr; xi[i1] = di; }
        else if (tw == 8) { xr[i1] = di; xi[i1] = -dr; }     // * (-i)
        else {                                                 // (dr + i di)(c - i s)
          xr[i1] = fmaf(dr, C32[tw], di * S32[tw]);
          xi[i1] = fmaf(di, C32[tw], -dr * S32[tw]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(MEL_WPB * 32, 2) mel_fwd_warp_kernel(
    const float* __restrict__ wav, int L, int ldw, int T, long long nframes, int hop, int n_mels,
    const int* __restrict__ fb_ptr, const int* __restrict__ fb_idx, const float* __restrict__ fb_val,
    float* __restrict__ spec, int ld_spec, float* __restrict__ mel, int ld_mel, float* __restrict__ cplx) {
  // persistent CTA: the three lookup tables are staged in shared memory once (the 200 KB of per-warp buffers leave
  // almost no L1, so table reads from global memory would otherwise go to L2 on every frame)
  extern __shared__ __align__(16) uint8_t msm[];
  float2* s_tw32 = reinterpret_cast<float2*>(msm);
  float2* s_tw = reinterpret_cast<float2*>(msm + 32 * 32 * 8);
  float* s_fval = reinterpret_cast<float*>(msm + 32 * 32 * 8 + (NH + 8) * 8);
  int* s_fidx = reinterpret_cast<int*>(s_fval + MEL_FB_MAX);
  int* s_fptr = s_fidx + MEL_FB_MAX;
  uint8_t* wsm = msm + MEL_TAB_SMEM;                           // per warp: float2[32*33] transpose buffer + float[NBIN+3]
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) s_tw32[i] = g_tw32[i];
  for (int i = threadIdx.x; i <= NH; i += blockDim.x) s_tw[i] = g_tw[i];
  if (mel) {
    const int nnz = fb_ptr[n_mels];
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) { s_fval[i] = fb_val[i]; s_fidx[i] = fb_idx[i]; }
    for (int i = threadIdx.x; i <= n_mels; i += blockDim.x) s_fptr[i] = fb_ptr[i];
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
  for (long long frame = (long long)blockIdx.x * MEL_WPB + w; frame < nframes; frame += (long long)gridDim.x * MEL_WPB) {
  const int b = (int)(frame / T), f = (int)(frame - (long long)b * T);
  const int pad = (NFFT - hop) / 2, s0 = f * hop - pad;
  const float* wv = wav + (long long)b * ldw;
  {                                                            // pull the next frame of this warp towards L2 while this one computes
    const long long nf = frame + (long long)gridDim.x * MEL_WPB;
    if (nf < nframes) {
      const int nb = (int)(nf / T), nfi = (int)(nf - (long long)nb * T);
      const int ns0 = max(0, min(nfi * hop - pad, L - NFFT));
      const float* np_ = wav + (long long)nb * ldw + ns0 + t * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(np_));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + 32));
    }
  }
  float xr[32], xi[32];
  const bool interior = (s0 >= 0) && (s0 + NFFT <= L) && (((s0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(wv) & 7) == 0));
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) {                            // z[m] = (x[2m] w[2m], x[2m+1] w[2m+1]), m = 32 n1 + t
    const int m = 32 * n1 + t;
    const float2 hw = *reinterpret_cast<const float2*>(&g_hann[2 * m]);
    float2 v;
    if (interior) v = *reinterpret_cast<const float2*>(wv + s0 + 2 * m);
    else v = make_float2(wv[reflect_idx(s0 + 2 * m, L)], wv[reflect_idx(s0 + 2 * m + 1, L)]);
    xr[n1] = v.x * hw.x;
    xi[n1] = v.y * hw.y;
  }
  fft32_regs(xr, xi);                                          // A[k1][n2 = t] in element brev5(k1)
  float2* z = reinterpret_cast<float2*>(wsm + (size_t)w * MEL_WARP_SMEM);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {                            // inter-stage twiddle, then transpose through smem
    const float2 tw = s_tw32[k1 * 32 + t];
    const float ar = xr[brev5(k1)], ai = xi[brev5(k1)];
    z[k1 * 33 + t] = make_float2(ar * tw.x - ai * tw.y, ar * tw.y + ai * tw.x);
  }
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {                            // lane t now owns k1 = t
    const float2 v = z[t * 33 + n2];
    xr[n2] = v.x; xi[n2] = v.y;
  }
  __syncwarp();
  fft32_regs(xr, xi);                                          // Z[t + 32 k2] in element brev5(k2)
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) z[t + 32 * k2] = make_float2(xr[brev5(k2)], xi[brev5(k2)]);
  __syncwarp();
  float* mg = reinterpret_cast<float*>(wsm + (size_t)w * MEL_WARP_SMEM + 32 * 33 * sizeof(float2));
  float2* cp = cplx ? reinterpret_cast<float2*>(cplx) + frame * NBIN : nullptr;
#pragma unroll
  for (int j = 0; j <= 32; ++j) {                              // real-FFT untangle: bins k = t + 32 j (+ bin 1024)
    const int k = t + 32 * j;
    if (j == 32 && t != 0) break;
    const float2 zk = (j < 32) ? make_float2(xr[brev5(j & 31)], xi[brev5(j & 31)]) : z[0];
    float2 zc = z[(NH - k) & (NH - 1)];
    zc.y = -zc.y;
    const float2 sm = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 dd = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
    const float2 tw = s_tw[k];
    const float2 tt = cmul(tw, dd);
    const float2 X = make_float2(sm.x + tt.y, sm.y - tt.x);
    const float m = sqrtf(X.x * X.x + X.y * X.y + 1e-6f);
    mg[k] = m;
    if (spec) spec[frame * ld_spec + k] = m;
    if (cp) cp[k] = X;
  }
  __syncwarp();
  if (mel) {                                                   // sparse filterbank, CSR by mel row staged in shared memory
    for (int mm = t; mm < n_mels; mm += 32) {
      float acc = 0.f;
      const int e1 = s_fptr[mm + 1];
      for (int e = s_fptr[mm]; e < e1; ++e) acc = fmaf(s_fval[e], mg[s_fidx[e]], acc);
      mel[frame * ld_mel + mm] = logf(fmaxf(acc, 1e-5f));
    }
  }
  __syncwarp();                                                // the per-warp buffers are reused by the next frame
  }
}

// adjoint: d wav from d logmel (per frame direct inverse transform; only B*32 frames per train step)
__global__ void __launch_bounds__(256) mel_bwd_kernel(const float* __restrict__ dmel, int ld_dmel,
                                                      const float* __restrict__ cplx, const float* __restrict__ mel,
                                                      int ld_mel, int L, int ldw, int T, int hop, int n_mels,
                                                      const int* __restrict__ fb_ptr, const int* __restrict__ fb_idx,
                                                      const float* __restrict__ fb_val, float* __restrict__ dwav) {
  __shared__ float2 tw[NFFT];
  __shared__ float dre[NBIN + 3], dim_[NBIN + 3], dmag[NBIN + 3];
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int b = frame / T, f = frame - b * T;
  for (int i = tid; i < NFFT; i += 256) tw[i] = g_tw[i];
  for (int k = tid; k < NBIN; k += 256) dmag[k] = 0.f;
  __syncthreads();
  for (int m = tid; m < n_mels; m += 256) {
    const float lm = mel[(long long)frame * ld_mel + m];
    // d log(clamp(s,1e-5))/ds = 1/s where s >= 1e-5 (torch.clamp passes the gradient at s >= min)
    const float g = (lm > logf(1e-5f)) ? dmel[(long long)frame * ld_dmel + m] * expf(-lm) : 0.f;
    if (g != 0.f)
      for (int e = fb_ptr[m]; e < fb_ptr[m + 1]; ++e) atomicAdd(&dmag[fb_idx[e]], fb_val[e] * g);
  }
  __syncthreads();
  for (int k = tid; k < NBIN; k += 256) {
    const float2 x = reinterpret_cast<const float2*>(cplx)[(long long)frame * NBIN + k];
    const float m = sqrtf(x.x * x.x + x.y * x.y + 1e-6f);
    const float s = dmag[k] / m;
    dre[k] = s * x.x;
    dim_[k] = s * x.y;
  }
  __syncthreads();
  const int pad = (NFFT - hop) / 2, s0 = f * hop - pad;
  float* dw = dwav + (long long)b * ldw;
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int n = tid + 256 * i;
    float acc = 0.f;
    int ph = 0;                                   // (k * n) mod 2048
    for (int k = 0; k < NBIN; ++k) {
      const float2 t = tw[ph];
      acc = fmaf(dre[k], t.x, acc);
      acc = fmaf(dim_[k], t.y, acc);
      ph = (ph + n) & (NFFT - 1);
    }
    atomicAdd(&dw[reflect_idx(s0 + n, L)], acc * g_hann[n]);
  }
}

__global__ void spec_to_mel_kernel(const float* __restrict__ spec, long long rows, int ld_spec, int n_mels,
                                   const int* __restrict__ fb_ptr, const int* __restrict__ fb_idx,
                                   const float* __restrict__ fb_val, float* __restrict__ mel, int ld_mel) {
  const long long row = blockIdx.x;
  const float* s = spec + row * ld_spec;
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    float acc = 0.f;
    for (int e = fb_ptr[m]; e < fb_ptr[m + 1]; ++e) acc = fmaf(fb_val[e], s[fb_idx[e]], acc);
    mel[row * ld_mel + m] = logf(fmaxf(acc, 1e-5f));
  }
}

}  // namespace evk
using namespace evk;

static int g_mel_variant = 1;      // 1: warp-per-frame register FFT (default); 0: CTA-per-frame shared-memory FFT
extern "C" int evk_set_mel_variant(int32_t v) { g_mel_variant = v ? 1 : 0; return EVK_OK; }

static int frames_of(int L, int hop) { return (L + 2 * ((NFFT - hop) / 2) - NFFT) / hop + 1; }

extern "C" int evk_mel_fwd(const float* wav, int32_t B, int32_t L, int32_t ldw, int32_t hop, int32_t n_mels,
                           const int32_t* fb_ptr, const int32_t* fb_idx, const float* fb_val, float* spec,
                           int32_t ld_spec, float* mel, int32_t ld_mel, float* cplx, evk_stream_t stream) {
  EVK_REQUIRE(wav && B >= 1 && hop >= 1 && hop <= NFFT, EVK_ERR_ARG, "mel_fwd: bad arguments");
  EVK_REQUIRE(L > (NFFT - hop) / 2, EVK_ERR_ARG, "mel_fwd: L=%d too short for reflect padding %d", L, (NFFT - hop) / 2);
  EVK_REQUIRE(!mel || (fb_ptr && fb_idx && fb_val), EVK_ERR_ARG, "mel_fwd: filterbank required");
  int rc = mel_init_tables();
  if (rc) { set_error("mel_fwd: table init failed"); return rc; }
  const int T = frames_of(L, hop);
  if (T <= 0) return EVK_OK;
  const long long nframes = (long long)B * T;
  if (g_mel_variant == 0 || n_mels > 143) {                                      // reference variant: one 256-thread CTA per frame
    mel_fwd_kernel<<<B * T, 256, 0, (cudaStream_t)stream>>>(wav, L, ldw, T, hop, n_mels, fb_ptr, fb_idx, fb_val, spec,
                                                            ld_spec, mel, ld_mel, cplx);
    return check_launch("mel_fwd_kernel");
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mel_fwd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEL_TAB_SMEM + MEL_WPB * MEL_WARP_SMEM);
    attr_set = true;
  }
  const long long need = (nframes + MEL_WPB - 1) / MEL_WPB;
  const unsigned grid = (unsigned)(need < 2 * 148 ? need : 2 * 148);            // persistent: two CTAs per SM
  mel_fwd_warp_kernel<<<grid, MEL_WPB * 32, MEL_TAB_SMEM + MEL_WPB * MEL_WARP_SMEM, (cudaStream_t)stream>>>(
      wav, L, ldw, T, nframes, hop, n_mels, fb_ptr, fb_idx, fb_val, spec, ld_spec, mel, ld_mel, cplx);
  return check_launch("mel_fwd_warp_kernel");
}

extern "C" int evk_mel_bwd(const float* dmel, int32_t ld_dmel, const float* cplx, const float* mel, int32_t ld_mel,
                           int32_t B, int32_t L, int32_t ldw, int32_t hop, int32_t n_mels, const int32_t* fb_ptr,
                           const int32_t* fb_idx, const float* fb_val, float* dwav, evk_stream_t stream) {
  EVK_REQUIRE(dmel && cplx && mel && dwav && fb_ptr && fb_idx && fb_val, EVK_ERR_ARG, "mel_bwd: null tensor");
  int rc = mel_init_tables();
  if (rc) { set_error("mel_bwd: table init failed"); return rc; }
  const int T = frames_of(L, hop);
  if (T <= 0) return EVK_OK;
  mel_bwd_kernel<<<B * T, 256, 0, (cudaStream_t)stream>>>(dmel, ld_dmel, cplx, mel, ld_mel, L, ldw, T, hop, n_mels,
                                                          fb_ptr, fb_idx, fb_val, dwav);
  return check_launch("mel_bwd_kernel");
}

extern "C" int evk_spec_to_mel(const float* spec, int64_t rows, int32_t ld_spec, int32_t n_mels, const int32_t* fb_ptr,
                               const int32_t* fb_idx, const float* fb_val, float* mel, int32_t ld_mel,
                               evk_stream_t stream) {
  EVK_REQUIRE(spec && mel && fb_ptr && fb_idx && fb_val, EVK_ERR_ARG, "spec_to_mel: null tensor");
  if (rows <= 0) return EVK_OK;
  spec_to_mel_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(spec, rows, ld_spec, n_mels, fb_ptr, fb_idx,
                                                                       fb_val, mel, ld_mel);
  return check_launch("spec_to_mel_kernel");
}
