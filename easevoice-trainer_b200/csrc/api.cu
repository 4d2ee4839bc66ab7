// Library bookkeeping: init / error reporting.
#include "evk_common.cuh"
#include <stdarg.h>

namespace evk {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return EVK_ERR_CUDA;
  }
  return EVK_OK;
}
int mel_init_tables();
double g_disp_flops[EVK_DISPATCH_SLOTS] = {0};
}  // namespace evk
using namespace evk;

extern "C" const char* evk_last_error(void) { return g_err; }
extern "C" int evk_version(void) { return 100; }
extern "C" int evk_gconv_desc_size(void) { return (int)sizeof(evk_gconv_desc); }

extern "C" int evk_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("evk_init: no CUDA device (%s) -- there is no CPU fallback", cudaGetErrorString(e)); return EVK_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { set_error("evk_init: %s", cudaGetErrorString(e)); return EVK_ERR_CUDA; }
  if (prop.major != 10) {
    set_error("evk_init: device '%s' is sm_%d%d; libevk_sm100 is built for sm_100a (B200) only", prop.name, prop.major, prop.minor);
    return EVK_ERR_ARCH;
  }
  int rc = mel_init_tables();
  if (rc) { set_error("evk_init: twiddle/window table upload failed"); return rc; }
  return EVK_OK;
}

extern "C" int evk_sync_check(evk_stream_t stream) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("evk_sync_check: %s", cudaGetErrorString(e)); return EVK_ERR_CUDA; }
  return EVK_OK;
}

// Host-side dispatch accounting (parity tests / bench.py): algorithmic flops of every contraction launch since the last
// reset, by the kernel family that served it.  Counted when the launch is enqueued -- a CUDA-graph replay adds nothing,
// so callers account one eager step of the shape they replay.
extern "C" int evk_dispatch_stats(double* out, int32_t n) {
  EVK_REQUIRE(out && n >= 1, EVK_ERR_ARG, "dispatch_stats: null output");
  for (int i = 0; i < n && i < EVK_DISPATCH_SLOTS; ++i) out[i] = g_disp_flops[i];
  return EVK_OK;
}
extern "C" int evk_dispatch_stats_reset(void) {
  for (int i = 0; i < EVK_DISPATCH_SLOTS; ++i) g_disp_flops[i] = 0.0;
  return EVK_OK;
}
