// capi.cu -- error string, version, device probing (no compute here).
#include <stdarg.h>

#include "common.cuh"

namespace pomp {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device visible (%s); libpomp_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return POMP_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return POMP_ERR_INVALID;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
    return POMP_ERR_CUDA;
  }
  return POMP_OK;
}

int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int v = 148;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) v = 148;
  if (device >= 0 && device < 64) cached[device] = v;
  return v;
}

}  // namespace pomp

POMP_API const char* pomp_last_error(void) { return pomp::g_err; }
POMP_API int pomp_version(void) { return 100; }
POMP_API int pomp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
POMP_API int64_t pomp_launch_count(void) { return pomp::g_launches.load(); }
