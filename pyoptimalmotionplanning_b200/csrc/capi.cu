// capi.cu -- error string, version, device probing (no compute here).
#include <stdarg.h>

#include "common.cuh"

namespace pomp {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<uint64_t> g_realloc_gen{0};

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

// ---- diagnostics: un-fused FP64 throughput ---------------------------------------------------
// The brute-force kernels of the planner regime (N <= ~1e5) are bound by the FP64 pipe WITHOUT
// fused multiply-add (the reference rounds after every operation, so the library is compiled with
// -fmad=false): Q*N*(3D+2) separate DADD/DMUL per batch.  This measures what the device sustains on
// exactly that instruction mix -- the denominator bench.py reports the brute-force regime against
// (SURVEY 8d), instead of the DFMA peak or HBM.
namespace pomp {
__global__ void __launch_bounds__(256) fp64_unfused_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.125, a2 = a0 + 0.25, a3 = a0 + 0.375;
  double b0 = 1.0000001, b1 = 0.9999999, b2 = 1.0000002, b3 = 0.9999998;
  const double c = 1e-9;
  for (int i = 0; i < iters; i++) {  // 8 independent chains x (1 DMUL + 1 DADD): 16 FP64 ops per iteration
    a0 = a0 * b0;
    a1 = a1 * b1;
    a2 = a2 * b2;
    a3 = a3 * b3;
    b0 = b0 + c;
    b1 = b1 - c;
    b2 = b2 + c;
    b3 = b3 - c;
    a0 = a0 + c;
    a1 = a1 - c;
    a2 = a2 + c;
    a3 = a3 - c;
    b0 = b0 * 0.99999999;
    b1 = b1 * 1.00000001;
    b2 = b2 * 0.99999999;
    b3 = b3 * 1.00000001;
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((b0 + b1) + (b2 + b3));
}
}  // namespace pomp

POMP_API int pomp_fp64_unfused_peak(int device, double* flops_out) {
  POMP_REQUIRE(flops_out != nullptr, "flops_out is NULL");
  pomp::DeviceGuard g(device);
  if (!g.ok) return POMP_ERR_CUDA;
  const int blocks = pomp::sm_count(device) * 8, threads = 256, iters = 20000;
  double* d = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  LAUNCH(pomp::fp64_unfused_kernel, blocks, threads, 0, 0, d, 1000);  // warm-up
  CUDA_TRY(cudaEventRecord(e0, 0));
  LAUNCH(pomp::fp64_unfused_kernel, blocks, threads, 0, 0, d, iters);
  CUDA_TRY(cudaEventRecord(e1, 0));
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *flops_out = 16.0 * iters * (double)blocks * threads / (ms * 1e-3);
  return POMP_OK;
}
