// Library-level entry points of libsegma_b200: error text, version, device check.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace segma {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SEGMA_OK;
  set_last_error("%s: %s", what, cudaGetErrorString(e));
  return SEGMA_ERR_CUDA;
}

int device_sm_count() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached[dev] = n;
    else
      return 148;
  }
  return cached[dev];
}

}  // namespace segma

extern "C" {

const char* segma_last_error(void) { return segma::g_last_error; }

int segma_version(void) { return 100; }

int segma_sm_count(void) { return segma::device_sm_count(); }

int segma_device_check(void) {
  int dev = 0, major = 0;
  SEGMA_CUDA_OK(cudaGetDevice(&dev));
  SEGMA_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    segma::set_last_error("libsegma_b200 is built for sm_100a only; device has compute capability %d.x", major);
    return SEGMA_ERR_UNSUPPORTED;
  }
  return SEGMA_OK;
}

}  // extern "C"
