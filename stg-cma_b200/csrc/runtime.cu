// Error plumbing and device queries for libstgcma.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace stgcma {

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
    return STGCMA_ECUDA;
  }
  return STGCMA_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
    if (cached <= 0) cached = 148;
  }
  return cached;
}

}  // namespace stgcma

extern "C" const char* stgcma_last_error(void) { return stgcma::g_err; }

extern "C" int stgcma_version(void) { return 100; }

extern "C" int stgcma_check_device(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    stgcma::set_error("check_device: %s", cudaGetErrorString(e));
    return STGCMA_ECUDA;
  }
  if (p.major != 10 || p.minor != 0) {   // the cubins are sm_100a: arch-specific, not forward compatible to other 10.x parts
    stgcma::set_error("check_device: device %d is sm_%d%d; libstgcma is built for sm_100a only (no fallback)", dev,
                      p.major, p.minor);
    return STGCMA_EARCH;
  }
  return STGCMA_OK;
}
