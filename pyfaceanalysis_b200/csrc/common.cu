// libhgsfa: error channel and version entry points.
#include "common.cuh"

namespace hgsfa {

std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return 1;
}

}  // namespace hgsfa

extern "C" const char* hgsfa_last_error(void) { return hgsfa::last_error_ref().c_str(); }
extern "C" int hgsfa_version(void) { return HGSFA_VERSION; }
extern "C" int hgsfa_device_count(int* count) {
  HG_CHECK(count, "hgsfa_device_count: null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return hgsfa::fail("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  *count = n;
  return 0;
}
