// semk_api.cu -- library-level entry points: version, error string, device probe.
#include "semk_common.cuh"

static thread_local std::string g_last_error;

void semk_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" int semk_version(void) { return SEMK_VERSION; }

extern "C" const char *semk_last_error(void) { return g_last_error.c_str(); }

extern "C" int semk_device_available(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();  // clear the sticky "no device" status
    return 0;
  }
  return n > 0 ? 1 : 0;
}
