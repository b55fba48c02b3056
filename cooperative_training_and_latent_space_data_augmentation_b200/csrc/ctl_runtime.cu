// Library-wide plumbing of the C ABI: version, thread-local error string, device properties.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "ctl_common.cuh"

namespace ctl {
namespace {
thread_local char g_error[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return CTL_ERR_CUDA;
}

int sm_count() {
  // per-device cache; device count is small
  static int cache[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice (no CUDA device? this library has no CPU fallback)"); return -1; }
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaDeviceGetAttribute(MultiProcessorCount)"); return -1; }
  if (dev >= 0 && dev < 64) cache[dev] = sms;
  return sms;
}

bool pdl_chain_enabled() {
  static const bool on = [] { const char* e = getenv("CTL_PDL_ALL"); return !(e && e[0] == '0'); }();
  return on;
}

// Profiling by elimination (tools/diag_conv.py): stages of the tcgen05 kernels can be switched off to see which one
// bounds the pipeline.  Results are garbage with any bit set; never set outside the diagnostic tool.  Compiled in only with -DCTL_DIAG
// (make DIAG=1); the default build ignores the variable and the kernels carry no skip branches.
int diag_flags() {
#ifdef CTL_DIAG
  const char* e = getenv("CTL_DIAG_SKIP");
  return e ? atoi(e) : 0;
#else
  return 0;
#endif
}
}  // namespace ctl

extern "C" int ctl_version(void) { return CTL_B200_VERSION; }
extern "C" const char* ctl_last_error(void) { return ctl::g_error; }
extern "C" int ctl_device_sm_count(void) { return ctl::sm_count(); }
