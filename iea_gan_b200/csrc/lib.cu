// lib.cu -- library bookkeeping: version, thread-local error string, device check.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace iea {
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
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -4;
  }
  return 0;
}
}  // namespace iea

extern "C" {
int iea_version(void) { return 100; }
const char* iea_last_error(void) { return iea::g_err; }
int iea_require_sm100(int device) {
  cudaDeviceProp p;
  IEA_CUDA(cudaGetDeviceProperties(&p, device));
  IEA_CHECK_ARG(p.major == 10, "device %d is sm_%d%d; libiea_sm100 is built for sm_100a only (no fallback)",
                device, p.major, p.minor);
  return 0;
}
int iea_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -3;
  return n;
}
}
