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
  mel_fwd_kernel<<<B * T, 256, 0, (cudaStream_t)stream>>>(wav, L, ldw, T, hop, n_mels, fb_ptr, fb_idx, fb_val, spec,
                                                          ld_spec, mel, ld_mel, cplx);
  return check_launch("mel_fwd_kernel");
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
