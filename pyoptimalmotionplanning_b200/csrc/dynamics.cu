// dynamics.cu -- batched ControlSpace.nextState and the fused RandomControlSelector.select
// (pomp/planners/kinodynamicplanner.py:66-84): nextState + metric-to-target + strict-< arg-min.
#include "common.cuh"
#include "dyn.cuh"
#include "topk.cuh"

using namespace pomp;

__global__ void next_state_kernel(pomp_dyn_desc dy, const double* __restrict__ xs, const double* __restrict__ us,
                                  long long n, double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int XF = dy.x_dim + (dy.cost_kind != POMP_COST_NONE), U = dy.u_dim;
  double x[POMP_MAX_DIM], u[POMP_MAX_DIM], o[POMP_MAX_DIM];
  for (int j = 0; j < XF; j++) x[j] = xs[i * XF + j];
  for (int j = 0; j < U; j++) u[j] = us[i * U + j];
  next_state_full(dy, x, u, o);
  for (int j = 0; j < XF; j++) out[i * XF + j] = o[j];
}

// one warp per parent state; lanes stride over its S candidate controls.  The reference keeps the
// first candidate with the strictly smallest distance == lexicographic min of (distance, sample#).
__global__ void select_control_kernel(pomp_dyn_desc dy, pomp_metric_desc m, const double* __restrict__ xs,
                                      const double* __restrict__ xdes, const double* __restrict__ us,
                                      const uint8_t* __restrict__ valid, long long B, int S,
                                      int* __restrict__ best, double* __restrict__ xbest,
                                      double* __restrict__ dbest) {
  long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int XF = dy.x_dim + (dy.cost_kind != POMP_COST_NONE), U = dy.u_dim;
  double x[POMP_MAX_DIM], xd[POMP_MAX_DIM], u[POMP_MAX_DIM], o[POMP_MAX_DIM], ob[POMP_MAX_DIM];
  for (int j = 0; j < XF; j++) {
    x[j] = xs[b * XF + j];
    xd[j] = xdes[b * XF + j];
    ob[j] = nan("");
  }
  double bd = POMP_INF;
  int bi = IDX_NONE;
  for (int s = lane; s < S; s += 32) {
    if (valid && !valid[b * S + s]) continue;
    for (int j = 0; j < U; j++) u[j] = us[(b * S + s) * U + j];
    next_state_full(dy, x, u, o);
    double d = metric_eval(m, o, xd);  // metric(xnext, xdesired)
    if (d < bd) {                      // ascending s per lane: strict '<' keeps the first
      bd = d;
      bi = s;
      for (int j = 0; j < XF; j++) ob[j] = o[j];
    }
  }
  // warp arg-min on (distance, sample index); +inf / NaN never win
  double wd = bd;
  int wi = bi;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double od = __shfl_xor_sync(FULL_MASK, wd, off);
    int oi = __shfl_xor_sync(FULL_MASK, wi, off);
    if (key_less(od, oi, wd, wi)) {
      wd = od;
      wi = oi;
    }
  }
  bool found = wd < POMP_INF;
  if (found && bi == wi) {  // exactly one lane owns the winning sample
    for (int j = 0; j < XF; j++) xbest[b * XF + j] = ob[j];
  }
  if (lane == 0) {
    best[b] = found ? wi : -1;
    dbest[b] = found ? wd : POMP_INF;
    if (!found)
      for (int j = 0; j < XF; j++) xbest[b * XF + j] = nan("");
  }
}

POMP_API int pomp_next_state(const pomp_dyn_desc* dyn, const double* x, const double* u, int64_t n, double* xnext,
                             void* stream) {
  POMP_TRY(validate_dyn(dyn));
  POMP_REQUIRE((n == 0 || (x && u && xnext)) && n >= 0, "bad arguments to pomp_next_state");
  if (n == 0) return POMP_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  POMP_TRY(ensure_device(dev));
  LAUNCH(next_state_kernel, (int)((n + 127) / 128), 128, 0, (cudaStream_t)stream, *dyn, x, u, (long long)n, xnext);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_select_control(const pomp_dyn_desc* dyn, const pomp_metric_desc* metric, const double* x,
                                 const double* xdes, const double* u, const uint8_t* valid, int64_t B, int32_t S,
                                 int32_t* best, double* xnext_best, double* dist_best, void* stream) {
  POMP_TRY(validate_dyn(dyn));
  POMP_TRY(validate_metric(metric));
  POMP_REQUIRE(metric->dim == dyn->x_dim + (dyn->cost_kind != POMP_COST_NONE),
               "metric dim %d does not match the state dimension", metric->dim);
  POMP_REQUIRE((B == 0 || (x && xdes && u && best && xnext_best && dist_best)) && B >= 0 && S >= 1,
               "bad arguments to pomp_select_control");
  if (B == 0) return POMP_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  POMP_TRY(ensure_device(dev));
  LAUNCH(select_control_kernel, (int)((B * 32 + 127) / 128), 128, 0, (cudaStream_t)stream, *dyn, *metric, x, xdes, u,
         valid, (long long)B, S, best, xnext_best, dist_best);
  CHECK_LAUNCH();
  return POMP_OK;
}

// ---- host-buffer variants (planners call these once per iteration) ---------------------------------------------------
// One call = pack the inputs into a pinned buffer, ONE async copy in, the kernel, ONE async copy of the packed outputs,
// one stream synchronisation.  The staging (device + pinned, grow-only) and the stream persist per host thread and
// device.  (First version: cudaMalloc + four blocking pageable copies in + three out + cudaFree per call, ~0.1 ms.)
namespace {
struct HostStage {
  int device = -1;
  unsigned char* dev = nullptr;
  unsigned char* pin = nullptr;
  size_t cap = 0;
  cudaStream_t st = nullptr;
  ~HostStage() {  // process exit: the context may already be gone, errors are ignored
    if (dev) cudaFree(dev);
    if (pin) cudaFreeHost(pin);
    if (st) cudaStreamDestroy(st);
  }
  int reserve(int device_now, size_t bytes) {
    if (device_now != device) {
      if (dev) cudaFree(dev);
      if (pin) cudaFreeHost(pin);
      if (st) cudaStreamDestroy(st);
      dev = pin = nullptr;
      st = nullptr;
      cap = 0;
      device = device_now;
    }
    if (!st) CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (bytes <= cap) return POMP_OK;
    if (dev) cudaFree(dev);
    if (pin) cudaFreeHost(pin);
    dev = pin = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    if (cudaMalloc((void**)&dev, want) != cudaSuccess || cudaMallocHost((void**)&pin, want) != cudaSuccess) {
      set_error("staging allocation of %zu bytes failed: %s", want, cudaGetErrorString(cudaGetLastError()));
      return POMP_ERR_NOMEM;
    }
    cap = want;
    return POMP_OK;
  }
};
thread_local HostStage g_stage;
inline size_t up8(size_t n) { return (n + 7) / 8 * 8; }
}  // namespace

POMP_API int pomp_next_state_h(const pomp_dyn_desc* dyn, const double* x_h, const double* u_h, int64_t n,
                               double* xnext_h) {
  POMP_TRY(validate_dyn(dyn));
  POMP_REQUIRE((n == 0 || (x_h && u_h && xnext_h)) && n >= 0, "bad arguments to pomp_next_state_h");
  if (n == 0) return POMP_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  POMP_TRY(ensure_device(dev));
  const int XF = dyn->x_dim + (dyn->cost_kind != POMP_COST_NONE);
  const size_t bx = sizeof(double) * (size_t)n * XF, bu = sizeof(double) * (size_t)n * dyn->u_dim;
  HostStage& S = g_stage;
  POMP_TRY(S.reserve(dev, 2 * bx + bu));
  // layout (device and pinned alike): x | u | xnext
  memcpy(S.pin, x_h, bx);
  memcpy(S.pin + bx, u_h, bu);
  CUDA_TRY(cudaMemcpyAsync(S.dev, S.pin, bx + bu, cudaMemcpyHostToDevice, S.st));
  POMP_TRY(pomp_next_state(dyn, (const double*)S.dev, (const double*)(S.dev + bx), n, (double*)(S.dev + bx + bu), S.st));
  CUDA_TRY(cudaMemcpyAsync(S.pin + bx + bu, S.dev + bx + bu, bx, cudaMemcpyDeviceToHost, S.st));
  CUDA_TRY(cudaStreamSynchronize(S.st));
  memcpy(xnext_h, S.pin + bx + bu, bx);
  return POMP_OK;
}

POMP_API int pomp_select_control_h(const pomp_dyn_desc* dyn, const pomp_metric_desc* metric, const double* x_h,
                                   const double* xdes_h, const double* u_h, const uint8_t* valid_h, int64_t B,
                                   int32_t S, int32_t* best_h, double* xnext_best_h, double* dist_best_h) {
  POMP_TRY(validate_dyn(dyn));
  POMP_REQUIRE((B == 0 || (x_h && xdes_h && u_h && best_h && xnext_best_h && dist_best_h)) && B >= 0 && S >= 1,
               "bad arguments to pomp_select_control_h");
  if (B == 0) return POMP_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  POMP_TRY(ensure_device(dev));
  const int XF = dyn->x_dim + (dyn->cost_kind != POMP_COST_NONE);
  const size_t bx = sizeof(double) * (size_t)B * XF, bu = sizeof(double) * (size_t)B * S * dyn->u_dim;
  const size_t bv = up8((size_t)B * S), bb = up8(sizeof(int32_t) * (size_t)B), bd = sizeof(double) * (size_t)B;
  // layout (device and pinned alike): inputs x | xdes | u | valid, outputs xbest | dist | best
  const size_t o_x = 0, o_xd = bx, o_u = 2 * bx, o_v = 2 * bx + bu, in_bytes = o_v + bv;
  const size_t o_xb = in_bytes, o_d = o_xb + bx, o_b = o_d + bd, out_bytes = bx + bd + bb;
  HostStage& G = g_stage;
  POMP_TRY(G.reserve(dev, in_bytes + out_bytes));
  memcpy(G.pin + o_x, x_h, bx);
  memcpy(G.pin + o_xd, xdes_h, bx);
  memcpy(G.pin + o_u, u_h, bu);
  if (valid_h) memcpy(G.pin + o_v, valid_h, (size_t)B * S);
  CUDA_TRY(cudaMemcpyAsync(G.dev, G.pin, valid_h ? in_bytes : o_v, cudaMemcpyHostToDevice, G.st));
  POMP_TRY(pomp_select_control(dyn, metric, (const double*)(G.dev + o_x), (const double*)(G.dev + o_xd),
                               (const double*)(G.dev + o_u), valid_h ? (const uint8_t*)(G.dev + o_v) : nullptr, B, S,
                               (int32_t*)(G.dev + o_b), (double*)(G.dev + o_xb), (double*)(G.dev + o_d), G.st));
  CUDA_TRY(cudaMemcpyAsync(G.pin + o_xb, G.dev + o_xb, out_bytes, cudaMemcpyDeviceToHost, G.st));
  CUDA_TRY(cudaStreamSynchronize(G.st));
  memcpy(xnext_best_h, G.pin + o_xb, bx);
  memcpy(dist_best_h, G.pin + o_d, bd);
  memcpy(best_h, G.pin + o_b, sizeof(int32_t) * (size_t)B);
  return POMP_OK;
}
