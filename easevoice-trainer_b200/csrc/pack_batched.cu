// Whole-network weight-norm + operand packing in ONE launch (and its backward in one launch).
// Round 1 issued one weight_pack launch per layer and step (342 + 304 launches of 2-20 us in the stage-2 step, plus as many
// zero-fills of the per-layer gradient buffers): the step's launch-bound tail.  Here the host builds, once per network, a
// table of jobs (pointers into the flat parameter arena, the packed arena and the packed-gradient arena are all static), and
// every step runs pack_batched over all (job, output-channel) rows.
//   forward : pa[q][d0][d1] = pb[q][d1][d0] = tf32_rn( v[d0][d1][q] * (g ? g[d0] / |v[d0]| : 1) )
//   backward: dv, dg from dpa (the tensor-core weight-gradient kernels accumulate into the dpa arena)
#include "evk_common.cuh"

namespace evk {

struct PackJob {                      // mirrored by ops.PackJobC (ctypes)
  const float* v; const float* g;     // parameters (g null: plain weight)
  float* pa; float* pb;               // packed operands (pb null: not needed)
  const float* dpa;                   // packed gradient (backward)
  float* dv; float* dg;               // parameter gradients (backward)
  int D0, D1, Q, lda, D0p, ldb, D1p, row0;   // row0: first global row (block) of this job
};

extern int g_precise;

__global__ void __launch_bounds__(256) pack_batched_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ job_of_row, int round_tf32) {
  __shared__ float red[33];
  const PackJob j = jobs[job_of_row[blockIdx.x]];
  const int d0 = blockIdx.x - j.row0;
  const long long n = (long long)j.D1 * j.Q;
  const float* vr = j.v + (long long)d0 * n;
  float scale = 1.f;
  if (j.g) {
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += vr[i] * vr[i];
    s = block_sum(s, red);
    scale = j.g[d0] / sqrtf(s);
  }
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int d1 = (int)(i / j.Q), q = (int)(i - (long long)d1 * j.Q);
    float w = vr[i] * scale;
    if (round_tf32) w = __uint_as_float(f2tf32(w));
    j.pa[((long long)q * j.D0p + d0) * j.lda + d1] = w;
    if (j.pb) j.pb[((long long)q * j.D1p + d1) * j.ldb + d0] = w;
  }
}

__global__ void __launch_bounds__(256) pack_bwd_batched_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ job_of_row) {
  __shared__ float red[33];
  const PackJob j = jobs[job_of_row[blockIdx.x]];
  const int d0 = blockIdx.x - j.row0;
  const long long n = (long long)j.D1 * j.Q;
  const float* vr = j.v + (long long)d0 * n;
  float* dvr = j.dv + (long long)d0 * n;
  if (!j.g) {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const int d1 = (int)(i / j.Q), q = (int)(i - (long long)d1 * j.Q);
      dvr[i] = j.dpa[((long long)q * j.D0p + d0) * j.lda + d1];
    }
    return;
  }
  float ss = 0.f, dot = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int d1 = (int)(i / j.Q), q = (int)(i - (long long)d1 * j.Q);
    const float vv = vr[i];
    ss += vv * vv;
    dot += j.dpa[((long long)q * j.D0p + d0) * j.lda + d1] * vv;
  }
  ss = block_sum(ss, red);
  dot = block_sum(dot, red);
  const float nrm = sqrtf(ss), gg = j.g[d0];
  if (threadIdx.x == 0) j.dg[d0] = dot / nrm;
  const float sc = gg / nrm, k = dot / ss;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int d1 = (int)(i / j.Q), q = (int)(i - (long long)d1 * j.Q);
    dvr[i] = sc * (j.dpa[((long long)q * j.D0p + d0) * j.lda + d1] - vr[i] * k);
  }
}

}  // namespace evk
using namespace evk;

// jobs: device array of PackJob (88 bytes each, layout above); job_of_row: device int32 [nrows], one entry per (job, d0) row.
extern "C" int evk_weight_pack_batched(const void* jobs, const int32_t* job_of_row, int32_t nrows, evk_stream_t stream) {
  EVK_REQUIRE(jobs && job_of_row && nrows >= 1, EVK_ERR_ARG, "weight_pack_batched: bad arguments");
  pack_batched_kernel<<<nrows, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs), job_of_row, !g_precise);
  return check_launch("pack_batched_kernel");
}

extern "C" int evk_weight_pack_bwd_batched(const void* jobs, const int32_t* job_of_row, int32_t nrows, evk_stream_t stream) {
  EVK_REQUIRE(jobs && job_of_row && nrows >= 1, EVK_ERR_ARG, "weight_pack_bwd_batched: bad arguments");
  pack_bwd_batched_kernel<<<nrows, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PackJob*>(jobs), job_of_row);
  return check_launch("pack_bwd_batched_kernel");
}
