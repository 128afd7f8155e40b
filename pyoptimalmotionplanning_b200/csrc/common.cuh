// common.cuh -- shared host/device helpers of libpomp_b200 (sm_100a only).
//
// Numerics contract: IEEE float64, one rounding per operation, the
// reference's operation order.  The library is compiled with -fmad=false so nvcc never contracts
// a*b+c into an FMA; sqrt and division on doubles are IEEE-correct in CUDA.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/pomp_b200.h"

#define POMP_API extern "C" __attribute__((visibility("default")))

namespace pomp {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<uint64_t> g_realloc_gen;  // bumped by every DevBuf reallocation (captured CUDA graphs bake scratch pointers in)

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pomp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return POMP_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define POMP_TRY(expr)       \
  do {                       \
    int _s = (expr);         \
    if (_s != POMP_OK) return _s; \
  } while (0)

#define POMP_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      pomp::set_error(__VA_ARGS__);  \
      return POMP_ERR_INVALID;       \
    }                                \
  } while (0)

// every kernel launch of the library goes through this so gpu_launches can be reported
#define LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                 \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
    pomp::g_launches.fetch_add(1, std::memory_order_relaxed);          \
  } while (0)

// Programmatic dependent launch: the kernel may be scheduled (and run up to its griddepcontrol.wait) while the previous
// kernel of the stream is still draining -- hides the launch gap between the kernels of one step.  The kernel MUST
// execute pomp::grid_dependency_wait() before it touches anything the previous kernel wrote.  Falls back to a plain
// launch while the stream is being captured into a CUDA graph.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cs == cudaStreamCaptureStatusNone ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                         \
  do {                                                                             \
    pomp::launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
    pomp::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
  } while (0)

#define CHECK_LAUNCH() CUDA_TRY(cudaGetLastError())

int validate_metric(const pomp_metric_desc* m);
int validate_dyn(const pomp_dyn_desc* d);
int ensure_device(int device);  // cudaSetDevice + "no CPU fallback" error
int sm_count(int device);

// RAII device guard for entry points
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = (ensure_device(device) == POMP_OK);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// growable device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {  // contents NOT preserved
    if (n <= cap) return POMP_OK;
    g_realloc_gen.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(p, 0, want * sizeof(T));  // scratch counters start from a defined state
    // the memset runs on the legacy default stream, which the library's NON-BLOCKING streams do not wait for:
    // without this it can land after the first kernel has already written into the new buffer
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? POMP_ERR_NOMEM : POMP_ERR_CUDA;
    }
    cap = want;
    return POMP_OK;
  }
  void release() {
    if (p) g_realloc_gen.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
};

#ifdef __CUDACC__
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- device math ------------------------------------------------------------------------
#define POMP_PI 3.141592653589793
#define POMP_TWOPI 6.283185307179586
#define POMP_INF (__longlong_as_double(0x7ff0000000000000LL))

// CPython float % (floatobject.c float_rem) as used by so2.normalize (pomp/klampt/so2.py:23-28).
// CUDA's fmod() is exact, like libm's.
__device__ __forceinline__ double py_fmod(double a, double m) {
  double r = fmod(a, m);
  if (r != 0.0) {
    if ((m < 0.0) != (r < 0.0)) r += m;
  } else {
    r = copysign(0.0, m);
  }
  return r;
}

__device__ __forceinline__ double so2_normalize(double a) {
  // already in (0, 2 pi): fmod returns its argument unchanged (exactly), skip the ~100-instruction
  // software fmod -- the common case for planner states
  if (a > 0.0 && a < POMP_TWOPI) return a;
  double res = py_fmod(a, POMP_TWOPI);
  if (res < 0) return res + POMP_TWOPI;
  return res;
}

// pomp/klampt/so2.py:30-37
__device__ __forceinline__ double so2_diff(double a, double b) {
  double d = so2_normalize(a) - so2_normalize(b);
  if (d < -POMP_PI) return d + POMP_TWOPI;
  if (d > POMP_PI) return d - POMP_TWOPI;
  return d;
}

// CPython >= 3.12 builtin sum() over floats: Neumaier compensation (bltinmodule.c).  Reached by
// L1Metric and WeightedEuclideanMetric (pomp/spaces/metric.py:10-11,19-21).
struct PySum {
  double s = 0.0, c = 0.0;
  __device__ __forceinline__ void add(double x) {
    double t = s + x;
    if (fabs(s) >= fabs(x)) c += (s - t) + x;
    else c += (x - t) + s;
    s = t;
  }
  __device__ __forceinline__ double result() const {
    if (c != 0.0 && isfinite(c)) return s + c;
    return s;
  }
};

// ---- metrics ------------------------------------------------------------------------------
// Every metric of the path depends on (a,b) only through per-coordinate differences e_i
// (a_i-b_i, or the SO(2) difference), plus CostMetric's asymmetric cost term.  Splitting
// "differences" from "combine" lets the grid search evaluate the SAME combine() on per-axis gaps
// to get a lower bound that is rigorous in floating point (every operation is monotone).

// e[0..d): differences of the base coordinates
__device__ __forceinline__ void metric_diffs(const pomp_metric_desc& m, const double* a, const double* b,
                                             int d, double* e) {
  if (m.kind == POMP_METRIC_GEODESIC_PRODUCT) {
    int i = 0;
    for (int c = 0; c < m.n_comp; c++) {
      if (m.comp_kind[c] == POMP_COMP_SO2) {
        e[i] = so2_diff(a[i], b[i]);  // so2space.py:9-10
        i++;
      } else {
        for (int j = 0; j < m.comp_dim[c]; j++, i++) e[i] = a[i] - b[i];
      }
    }
  } else {
    for (int i = 0; i < d; i++) e[i] = a[i] - b[i];
  }
}

// combine per-coordinate differences into the base distance (no cost term)
__device__ __forceinline__ double metric_combine(const pomp_metric_desc& m, const double* e, int d) {
  switch (m.kind) {
    case POMP_METRIC_EUCLIDEAN: {  // vectorops.py:94-102
      double s = 0.0;
      for (int i = 0; i < d; i++) s = s + e[i] * e[i];
      return sqrt(s);
    }
    case POMP_METRIC_WEIGHTED_EUCLIDEAN: {  // metric.py:19-21 (x**2 evaluated as x*x)
      PySum s;
      for (int i = 0; i < d; i++) s.add(m.dim_weight[i] * (e[i] * e[i]));
      return sqrt(s.result());
    }
    case POMP_METRIC_SCALED_EUCLIDEAN: {  // metric.py:22-23
      double s = 0.0;
      for (int i = 0; i < d; i++) s = s + e[i] * e[i];
      return m.scale * sqrt(s);
    }
    case POMP_METRIC_L1: {  // metric.py:10-11
      PySum s;
      for (int i = 0; i < d; i++) s.add(fabs(e[i]));
      return s.result();
    }
    case POMP_METRIC_LINF: {  // metric.py:13-14
      double s = fabs(e[0]);
      for (int i = 1; i < d; i++) {
        double v = fabs(e[i]);
        if (v > s) s = v;
      }
      return s;
    }
    case POMP_METRIC_GEODESIC_PRODUCT: {  // geodesicspace.py:45-52
      double res = 0.0;
      int i = 0;
      for (int c = 0; c < m.n_comp; c++) {
        double dc;
        if (m.comp_kind[c] == POMP_COMP_SO2) {
          dc = fabs(e[i]);
          i++;
        } else {
          double s = 0.0;
          for (int j = 0; j < m.comp_dim[c]; j++, i++) s = s + e[i] * e[i];
          dc = sqrt(s);
        }
        res += (dc * dc) * m.comp_weight[c];  // (dist**2)*w : rooted value squared again, then weighted
      }
      return sqrt(res);
    }
  }
  return nan("");
}

// metric(a, b): a = stored node, b = query (CostMetric is asymmetric: kinodynamicplanner.py:905-913)
__device__ __forceinline__ double metric_eval(const pomp_metric_desc& m, const double* a, const double* b) {
  double e[POMP_MAX_DIM];
  int d = m.has_cost ? m.dim - 1 : m.dim;
  metric_diffs(m, a, b, d, e);
  double dist = metric_combine(m, e, d);
  if (m.has_cost && m.cost_weight_set) {
    if (m.cost_max_set && (a[d] > m.cost_max || b[d] > m.cost_max)) return POMP_INF;
    double dc = a[d] - b[d];
    dist += (dc > 0.0 ? dc : 0.0) * m.cost_weight;
  }
  return dist;
}

// Euclidean fast path with compile-time dimension
template <int D>
__device__ __forceinline__ double euclid_fixed(const double* a, const double* b) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    double t = a[i] - b[i];
    s = s + t * t;
  }
  return sqrt(s);
}

// (distance, index) lexicographic "less": the library-wide tie rule (lowest insertion index wins)
__device__ __forceinline__ bool key_less(double d1, int i1, double d2, int i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

}  // namespace pomp
