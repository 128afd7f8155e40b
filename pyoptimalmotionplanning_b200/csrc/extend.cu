// extend.cu -- one kinematic RRT.expand step in a single launch
// (pomp/planners/kinodynamicplanner.py:368-411): pickNode -> NearestNeighbors.nearest (:451-461),
// KinematicControlSelector.select (:95-100), LinearInterpolator edge, EpsilonEdgeChecker.feasible
// (edgechecker.py:20-30), and on success the append NearestNeighbors.add (:407) -- the new node is
// written into the device point store by the kernel itself.  The planner loop is latency bound
// (tiny trees, one query per iteration); fusing the four hot calls leaves one launch and one
// stream synchronisation per iteration.
#include "common.cuh"
#include "nn_handle.cuh"
#include "space.cuh"
#include "topk.cuh"

using namespace pomp;

__global__ void rrt_extend_kernel(double* __restrict__ pts, long long n, int D, const uint8_t* __restrict__ skip,
                                  SpaceDev s, QueryArg xr, int check_sample, double max_step, double res, double* part_d, int* part_i,
                                  unsigned int* ticket, Mailbox* out) {
  __shared__ int sample_ok;
  if (threadIdx.x == 0) sample_ok = check_sample ? (feasible_pt(s, xr.v) ? 1 : 0) : 1;
  __syncthreads();
  if (!sample_ok) {  // RRT.expand: "if not self.cspace.feasible(xrand): return None" (:362-363)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      out->idx = -2;
      out->dist = POMP_INF;
      out->aux[0] = 0;
      out->len = 0.0;
      __threadfence_system();
    }
    return;
  }
  double p[POMP_MAX_DIM];
  double bd = POMP_INF;
  int bi = IDX_NONE;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (skip && skip[i]) continue;
    double sm = 0.0;
    for (int j = 0; j < D; j++) {
      p[j] = pts[i * D + j];
      double t = p[j] - xr.v[j];
      sm = sm + t * t;
    }
    double d = sqrt(sm);
    if (key_less(d, (int)i, bd, bi)) {
      bd = d;
      bi = (int)i;
    }
  }
  __shared__ double sd[32];
  __shared__ int si[32];
  __shared__ bool last;
  __shared__ double xn[POMP_MAX_DIM], xnew[POMP_MAX_DIM];
  __shared__ long long ks;
  __shared__ double elen;
  __shared__ int parent;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double od = __shfl_xor_sync(FULL_MASK, bd, o);
    int oi = __shfl_xor_sync(FULL_MASK, bi, o);
    if (key_less(od, oi, bd, bi)) {
      bd = od;
      bi = oi;
    }
  }
  if (lane == 0) {
    sd[w] = bd;
    si[w] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int x = 1; x < nw; x++)
      if (key_less(sd[x], si[x], bd, bi)) {
        bd = sd[x];
        bi = si[x];
      }
    part_d[blockIdx.x] = bd;
    part_i[blockIdx.x] = bi;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    bd = POMP_INF;
    bi = IDX_NONE;
    for (int b = 0; b < (int)gridDim.x; b++) {
      double d = ((volatile double*)part_d)[b];
      int i = ((volatile int*)part_i)[b];
      if (key_less(d, i, bd, bi)) {
        bd = d;
        bi = i;
      }
    }
    *ticket = 0;
    bool found = bd < POMP_INF;
    parent = found ? bi : -1;
    ks = -1;
    elen = 0.0;
    if (found) {
      // KinematicControlSelector.select: d = cspace.distance(x, xdesired); steer to maxDistance
      double sm = 0.0;
      for (int j = 0; j < D; j++) {
        xn[j] = pts[(long long)bi * D + j];
        double t = xn[j] - xr.v[j];
        sm = sm + t * t;
      }
      double d = sqrt(sm);
      if (d > max_step) {
        double u = max_step / d;  // e.eval(maxDistance/d) = x + u*(xdesired - x)
        for (int j = 0; j < D; j++) xnew[j] = xn[j] + u * (xr.v[j] - xn[j]);
      } else {
        for (int j = 0; j < D; j++) xnew[j] = xr.v[j];
      }
      // EpsilonEdgeChecker: length of the LinearInterpolator xn -> xnew
      double l = 0.0;
      for (int j = 0; j < D; j++) {
        double t = xn[j] - xnew[j];
        l = l + t * t;
      }
      elen = sqrt(l);
      ks = edge_samples(elen, res);
    }
  }
  __syncthreads();
  int ok = 1;
  if (parent < 0) {
    ok = 0;
  } else {
    // endpoints + k interior samples spread over the block
    long long total = ks + 2;
    const double kk2 = (double)(ks + 2);
    for (long long t = threadIdx.x; t < total && ok; t += blockDim.x) {
      double x[POMP_MAX_DIM];
      if (t == 0) {
        for (int j = 0; j < D; j++) x[j] = xn[j];
      } else if (t == 1) {
        for (int j = 0; j < D; j++) x[j] = xnew[j];
      } else {
        double u = (double)(t - 1) / kk2;  // i = t-2, u = (i+1)/(k+2)
        for (int j = 0; j < D; j++) x[j] = xn[j] + u * (xnew[j] - xn[j]);
      }
      if (!feasible_pt(s, x)) ok = 0;
    }
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0) {
    out->idx = parent;
    out->dist = bd;
    out->aux[0] = ok;
    out->len = elen;
    for (int j = 0; j < D; j++) out->x[j] = parent >= 0 ? xnew[j] : nan("");
    if (ok)
      for (int j = 0; j < D; j++) pts[n * D + j] = xnew[j];  // NearestNeighbors.add
    __threadfence_system();
  }
}

POMP_API int pomp_rrt_extend_h(pomp_nn* h, pomp_space* sp, const double* xrand_h, int check_sample, double max_step,
                               double resolution, int64_t* parent_h, double* xnew_h, double* edge_len_h,
                               int32_t* added_h) {
  POMP_REQUIRE(h && sp && xrand_h && parent_h && xnew_h && added_h, "NULL argument to pomp_rrt_extend_h");
  POMP_REQUIRE(h->metric.kind == POMP_METRIC_EUCLIDEAN && !h->metric.has_cost, "rrt_extend needs the Euclidean metric");
  POMP_REQUIRE(sp->dev.dim == h->D, "space dim %d != point dim %d", sp->dev.dim, h->D);
  POMP_REQUIRE(resolution > 0.0 && max_step >= 0.0, "bad step/resolution");
  POMP_REQUIRE(sp->device == h->device, "space and index live on different devices");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = h->stream;
  POMP_TRY(nn_flush(h, st));
  POMP_TRY(nn_grow(h, h->n_dev + 1, st));
  const int64_t n = h->n_dev;
  QueryArg qa;
  for (int j = 0; j < POMP_MAX_DIM; j++) qa.v[j] = j < h->D ? xrand_h[j] : 0.0;
  int blocks = (int)std::min<int64_t>(std::max<int64_t>((n + 2047) / 2048, 1), (int64_t)sm_count(h->device));
  POMP_TRY(h->part_d.reserve((size_t)blocks));
  POMP_TRY(h->part_i.reserve((size_t)blocks));
  LAUNCH(rrt_extend_kernel, blocks, 256, 0, st, h->d_pts, (long long)n, h->D, h->skip_ptr(), sp->dev, qa, check_sample,
         max_step, resolution, h->part_d.p, h->part_i.p, h->ticket, (Mailbox*)h->mailbox_d);
  CHECK_LAUNCH();
  CUDA_TRY(cudaStreamSynchronize(st));
  const Mailbox* mb = (const Mailbox*)h->mailbox_h;
  *parent_h = mb->idx;
  *added_h = (int32_t)mb->aux[0];
  if (edge_len_h) *edge_len_h = mb->len;
  for (int j = 0; j < h->D; j++) xnew_h[j] = mb->idx >= 0 ? mb->x[j] : NAN;
  if (mb->aux[0]) h->n_dev += 1;
  return POMP_OK;
}
