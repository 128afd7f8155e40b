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

// ---- host-buffer variants (one-shot staging; planners call these once per iteration) --------
namespace {
struct Staging {
  void* p = nullptr;
  ~Staging() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? POMP_ERR_NOMEM : POMP_ERR_CUDA;
    }
    return POMP_OK;
  }
};
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
  size_t bx = sizeof(double) * (size_t)n * XF, bu = sizeof(double) * (size_t)n * dyn->u_dim;
  Staging s;
  POMP_TRY(s.alloc(2 * bx + bu));
  double* dx = (double*)s.p;
  double* dout = dx + (size_t)n * XF;
  double* du = dout + (size_t)n * XF;
  CUDA_TRY(cudaMemcpy(dx, x_h, bx, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(du, u_h, bu, cudaMemcpyHostToDevice));
  POMP_TRY(pomp_next_state(dyn, dx, du, n, dout, nullptr));
  CUDA_TRY(cudaMemcpy(xnext_h, dout, bx, cudaMemcpyDeviceToHost));
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
  size_t bx = sizeof(double) * (size_t)B * XF, bu = sizeof(double) * (size_t)B * S * dyn->u_dim;
  size_t bv = (size_t)B * S, bb = sizeof(int32_t) * (size_t)B, bd = sizeof(double) * (size_t)B;
  Staging s;
  POMP_TRY(s.alloc(3 * bx + bu + bd + bb + bv + 64));
  double* dx = (double*)s.p;
  double* dxd = dx + (size_t)B * XF;
  double* dxb = dxd + (size_t)B * XF;
  double* du = dxb + (size_t)B * XF;
  double* ddist = du + (size_t)B * S * dyn->u_dim;
  int32_t* dbest = (int32_t*)(ddist + B);
  uint8_t* dvalid = (uint8_t*)(dbest + B);
  CUDA_TRY(cudaMemcpy(dx, x_h, bx, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dxd, xdes_h, bx, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(du, u_h, bu, cudaMemcpyHostToDevice));
  if (valid_h) CUDA_TRY(cudaMemcpy(dvalid, valid_h, bv, cudaMemcpyHostToDevice));
  POMP_TRY(pomp_select_control(dyn, metric, dx, dxd, du, valid_h ? dvalid : nullptr, B, S, dbest, dxb, ddist, nullptr));
  CUDA_TRY(cudaMemcpy(best_h, dbest, bb, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(xnext_best_h, dxb, bx, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(dist_best_h, ddist, bd, cudaMemcpyDeviceToHost));
  return POMP_OK;
}
