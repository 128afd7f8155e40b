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
      ks = edge_samples(s, elen, res);
    }
  }
  __syncthreads();
  int ok = 1;
  if (parent < 0 || ks < 0) {  // no parent, or the edge asks for more than POMP_MAX_EDGE_SAMPLES samples
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
  POMP_TRY(space_overflow_check(sp));
  const Mailbox* mb = (const Mailbox*)h->mailbox_h;
  *parent_h = mb->idx;
  *added_h = (int32_t)mb->aux[0];
  if (edge_len_h) *edge_len_h = mb->len;
  for (int j = 0; j < h->D; j++) xnew_h[j] = mb->idx >= 0 ? mb->x[j] : NAN;
  if (mb->aux[0]) h->n_dev += 1;
  return POMP_OK;
}

// ============================================================================ RRT* rewire batch
// RRTStar.rewire (pomp/planners/rrtstarplanner.py:122-187) asks for the k nearest nodes of the new
// node and then checks, one call each, the straight edges neighbour -> new (incoming rewiring, in
// order of accumulated cost, :147-161) and new -> neighbour (outgoing, :172-183).  Edge
// feasibility is a pure function of the endpoints, so all 2k checks can be done up front in one
// call; the planner's own ordering / early exits then run on the cached answers.
__global__ void rewire_gather_kernel(const double* __restrict__ pts, int D, const int64_t* __restrict__ idx, int k,
                                     QueryArg xnew, double* __restrict__ a, double* __restrict__ b) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2 * k) return;
  const int jj = j % k, dir = j / k;  // dir 0: neighbour -> new, 1: new -> neighbour
  const int64_t i = idx[jj];
  for (int d = 0; d < D; d++) {
    const double pn = i >= 0 ? pts[i * D + d] : xnew.v[d];
    a[(long long)j * D + d] = dir == 0 ? pn : xnew.v[d];
    b[(long long)j * D + d] = dir == 0 ? xnew.v[d] : pn;
  }
}

POMP_API int pomp_rewire_candidates_h(pomp_nn* h, pomp_space* sp, const double* x_h, int32_t k, double resolution,
                                      int64_t* idx_h, double* dist_h, int32_t* count_h, uint8_t* ok_in_h,
                                      uint8_t* ok_out_h) {
  POMP_REQUIRE(h && sp && x_h && idx_h && dist_h && count_h && ok_in_h && ok_out_h,
               "NULL argument to pomp_rewire_candidates_h");
  POMP_REQUIRE(k >= 1 && k <= 128, "k=%d outside 1..128", k);
  POMP_REQUIRE(sp->dev.dim == h->D, "space dim %d != point dim %d", sp->dev.dim, h->D);
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  POMP_REQUIRE(sp->device == h->device, "space and index live on different devices");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = h->stream;
  const int D = h->D;
  // one packed result record [idx k][dist k][count, pad][ok 2k]: a single device-to-host copy into pinned memory
  const size_t off_dist = sizeof(int64_t) * (size_t)k, off_cnt = off_dist + sizeof(double) * (size_t)k;
  const size_t off_ok = off_cnt + 8, pack_bytes = off_ok + 2 * (size_t)k;
  POMP_TRY(h->q_dev.reserve((size_t)D));
  POMP_TRY(h->rw_ab.reserve((size_t)4 * k * D));
  POMP_TRY(h->rw_ok.reserve(pack_bytes + 16));
  if (h->rw_pin_cap < pack_bytes) {
    if (h->rw_pin) cudaFreeHost(h->rw_pin);
    h->rw_pin = nullptr;
    CUDA_TRY(cudaHostAlloc(&h->rw_pin, pack_bytes + 256, cudaHostAllocDefault));
    h->rw_pin_cap = pack_bytes + 256;
  }
  unsigned char* pk = h->rw_ok.p;
  int64_t* d_idx = reinterpret_cast<int64_t*>(pk);
  double* d_dist = reinterpret_cast<double*>(pk + off_dist);
  int32_t* d_cnt = reinterpret_cast<int32_t*>(pk + off_cnt);
  uint8_t* d_ok = pk + off_ok;
  CUDA_TRY(cudaMemcpyAsync(h->q_dev.p, x_h, sizeof(double) * D, cudaMemcpyHostToDevice, st));
  POMP_TRY(pomp_nn_knearest(h, h->q_dev.p, 1, k, d_idx, d_dist, d_cnt, st));
  QueryArg qa;
  for (int j = 0; j < POMP_MAX_DIM; j++) qa.v[j] = j < D ? x_h[j] : 0.0;
  double* a = h->rw_ab.p;
  double* b = h->rw_ab.p + (size_t)2 * k * D;
  LAUNCH(rewire_gather_kernel, (2 * k + 63) / 64, 64, 0, st, h->d_pts, D, d_idx, k, qa, a, b);
  CHECK_LAUNCH();
  POMP_TRY(pomp_edge_check(sp, a, b, 2 * k, resolution, d_ok, st));
  CUDA_TRY(cudaMemcpyAsync(h->rw_pin, pk, pack_bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  POMP_TRY(space_overflow_check(sp));
  const unsigned char* hp = static_cast<const unsigned char*>(h->rw_pin);
  memcpy(idx_h, hp, sizeof(int64_t) * (size_t)k);
  memcpy(dist_h, hp + off_dist, sizeof(double) * (size_t)k);
  memcpy(count_h, hp + off_cnt, sizeof(int32_t));
  memcpy(ok_in_h, hp + off_ok, (size_t)k);
  memcpy(ok_out_h, hp + off_ok + k, (size_t)k);
  return POMP_OK;
}

// ============================================================================ forest: lock-step trials
// T independent kinematic RRTs advanced by one RRT.expand each in ONE launch (one block per
// trial).  The planner loop is latency bound (one launch + one sync per iteration, tiny trees);
// running independent seeded trials in lock step -- main.py runs 10 per configuration,
// BASELINE config 5 asks for 4096 -- amortises that latency over all of them.  Every trial keeps
// its own host RNG, so each tree is identical to the one the single-trial driver grows.
struct pomp_forest {
  int device = 0, D = 0, T = 0;
  int64_t cap = 0;
  double* d_pts = nullptr;  // [T][cap][D]
  int* d_count = nullptr;   // [T] nodes per trial (device-side truth; host mirror in h_count)
  std::vector<int64_t> h_count;
  // zero-copy mailboxes (mapped pinned memory)
  struct In {
    double xrand[POMP_MAX_DIM];
    double root[POMP_MAX_DIM];
    int32_t active, reset;
  };
  struct Out {
    int64_t parent;
    double xnew[POMP_MAX_DIM];
    double len;
    int32_t added, overflow;
  };
  In* in_h = nullptr;
  In* in_d = nullptr;
  Out* out_h = nullptr;
  Out* out_d = nullptr;
  // multi-iteration launches (pomp_forest_extend_multi_h): up to MAXM expansions per trial per launch
  static constexpr int MAXM = 32;
  double* xr_h = nullptr;   // [T][MAXM][D] samples
  double* xr_d = nullptr;
  Out* outm_h = nullptr;    // [T][MAXM]
  Out* outm_d = nullptr;
  int32_t* ndone_h = nullptr;  // [T] expansions executed
  int32_t* ndone_d = nullptr;
  cudaStream_t stream = nullptr;
};

__global__ void __launch_bounds__(256)
forest_extend_kernel(double* __restrict__ pts, int* __restrict__ count, long long cap, int D, SpaceDev s,
                     const pomp_forest::In* __restrict__ in, pomp_forest::Out* __restrict__ out, int check_sample,
                     double max_step, double res) {
  const int t = blockIdx.x;
  __shared__ pomp_forest::In job;
  __shared__ int n_sh, sample_ok;
  if (threadIdx.x == 0) {
    job = in[t];  // one read of the mapped host record
    int n = count[t];
    if (job.reset) {  // RRT.reset (:312-320): the tree restarts from its root
      for (int j = 0; j < D; j++) pts[(long long)t * cap * D + j] = job.root[j];
      n = 1;
      count[t] = 1;
    }
    n_sh = n;
    sample_ok = (!check_sample || feasible_pt(s, job.xrand)) ? 1 : 0;
  }
  __syncthreads();
  pomp_forest::Out* o = out + t;
  if (!job.active) return;
  const int n = n_sh;
  double* P = pts + (long long)t * cap * D;
  if (!sample_ok || n >= cap) {
    if (threadIdx.x == 0) {
      o->parent = sample_ok ? -1 : -2;
      o->added = 0;
      o->overflow = (sample_ok && n >= cap) ? 1 : 0;
      o->len = 0.0;
      __threadfence_system();
    }
    return;
  }
  // nearest node of this trial's tree (strict '<' in index order: nearestneighbors.py:88-96)
  double bd = POMP_INF;
  int bi = IDX_NONE;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double sm = 0.0;
    for (int j = 0; j < D; j++) {
      double tt = P[(long long)i * D + j] - job.xrand[j];
      sm = sm + tt * tt;
    }
    double d = sqrt(sm);
    if (key_less(d, i, bd, bi)) {
      bd = d;
      bi = i;
    }
  }
  __shared__ double sd[8];
  __shared__ int si[8];
  __shared__ double xn[POMP_MAX_DIM], xnew[POMP_MAX_DIM];
  __shared__ long long ks;
  __shared__ double elen;
  __shared__ int parent;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double od = __shfl_xor_sync(FULL_MASK, bd, off);
    int oi = __shfl_xor_sync(FULL_MASK, bi, off);
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
    bool found = bd < POMP_INF;
    parent = found ? bi : -1;
    ks = -1;
    elen = 0.0;
    if (found) {
      double sm = 0.0;
      for (int j = 0; j < D; j++) {
        xn[j] = P[(long long)bi * D + j];
        double tt = xn[j] - job.xrand[j];
        sm = sm + tt * tt;
      }
      double d = sqrt(sm);
      if (d > max_step) {
        double u = max_step / d;
        for (int j = 0; j < D; j++) xnew[j] = xn[j] + u * (job.xrand[j] - xn[j]);
      } else {
        for (int j = 0; j < D; j++) xnew[j] = job.xrand[j];
      }
      double l = 0.0;
      for (int j = 0; j < D; j++) {
        double tt = xn[j] - xnew[j];
        l = l + tt * tt;
      }
      elen = sqrt(l);
      ks = edge_samples(s, elen, res);
    }
  }
  __syncthreads();
  int ok = 1;
  if (parent < 0 || ks < 0) {  // no parent, or the edge asks for more than POMP_MAX_EDGE_SAMPLES samples
    ok = 0;
  } else {
    const long long total = ks + 2;
    const double kk2 = (double)(ks + 2);
    for (long long q = threadIdx.x; q < total && ok; q += blockDim.x) {
      double x[POMP_MAX_DIM];
      if (q == 0) {
        for (int j = 0; j < D; j++) x[j] = xn[j];
      } else if (q == 1) {
        for (int j = 0; j < D; j++) x[j] = xnew[j];
      } else {
        double u = (double)(q - 1) / kk2;
        for (int j = 0; j < D; j++) x[j] = xn[j] + u * (xnew[j] - xn[j]);
      }
      if (!feasible_pt(s, x)) ok = 0;
    }
  }
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0) {
    o->parent = parent;
    o->added = ok;
    o->overflow = 0;
    o->len = elen;
    for (int j = 0; j < D; j++) o->xnew[j] = parent >= 0 ? xnew[j] : nan("");
    if (ok) {
      for (int j = 0; j < D; j++) P[(long long)n * D + j] = xnew[j];
      count[t] = n + 1;
    }
    __threadfence_system();
  }
}

// M expansions per trial in ONE launch.  The planner loop is bound by launch + synchronisation latency
// (~19 us per iteration for a 5 us kernel); RRT's samples do not depend on the tree, so the host draws the
// samples of the next M iterations up front and the block applies them one after the other to its tree,
// exactly as M successive forest_extend_kernel launches would.  A trial stops after the expansion whose
// new node lies in the goal ball (RepeatedRRT.planMore returns there, kinodynamicplanner.py:1291-1312);
// the host puts the unused random numbers back.
__global__ void __launch_bounds__(256)
forest_extend_multi_kernel(double* __restrict__ pts, int* __restrict__ count, long long cap, int D, SpaceDev s,
                           const pomp_forest::In* __restrict__ in, const double* __restrict__ xr, int M,
                           pomp_forest::Out* __restrict__ outm, int32_t* __restrict__ ndone, int check_sample,
                           double max_step, double res, QueryArg goal, int has_goal, double goal_radius) {
  const int t = blockIdx.x;
  __shared__ pomp_forest::In job;
  __shared__ int n_sh, sample_ok, stop_sh;
  __shared__ double xrand[POMP_MAX_DIM];
  __shared__ double sd[8];
  __shared__ int si[8];
  __shared__ double xn[POMP_MAX_DIM], xnew[POMP_MAX_DIM];
  __shared__ long long ks;
  __shared__ double elen;
  __shared__ int parent;
  if (threadIdx.x == 0) {
    job = in[t];
    int n = count[t];
    if (job.reset) {
      for (int j = 0; j < D; j++) pts[(long long)t * cap * D + j] = job.root[j];
      n = 1;
      count[t] = 1;
    }
    n_sh = n;
    stop_sh = 0;
  }
  __syncthreads();
  if (!job.active) {
    if (threadIdx.x == 0) ndone[t] = 0;
    return;
  }
  double* P = pts + (long long)t * cap * D;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int done = 0;
  for (int it = 0; it < M; it++) {
    pomp_forest::Out* o = outm + (long long)t * pomp_forest::MAXM + it;
    if (threadIdx.x == 0) {
      for (int j = 0; j < D; j++) xrand[j] = xr[((long long)t * pomp_forest::MAXM + it) * D + j];
      sample_ok = (!check_sample || feasible_pt(s, xrand)) ? 1 : 0;
    }
    __syncthreads();
    const int n = n_sh;
    done = it + 1;
    if (!sample_ok || n >= cap) {
      if (threadIdx.x == 0) {
        o->parent = sample_ok ? -1 : -2;
        o->added = 0;
        o->overflow = (sample_ok && n >= cap) ? 1 : 0;
        o->len = 0.0;
      }
      const bool full = sample_ok && n >= cap;
      __syncthreads();
      if (full) break;   // the host reports the capacity error
      continue;
    }
    double bd = POMP_INF;
    int bi = IDX_NONE;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double sm = 0.0;
      for (int j = 0; j < D; j++) {
        double tt = P[(long long)i * D + j] - xrand[j];
        sm = sm + tt * tt;
      }
      double d = sqrt(sm);
      if (key_less(d, i, bd, bi)) {
        bd = d;
        bi = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double od = __shfl_xor_sync(FULL_MASK, bd, off);
      int oi = __shfl_xor_sync(FULL_MASK, bi, off);
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
      bool found = bd < POMP_INF;
      parent = found ? bi : -1;
      ks = -1;
      elen = 0.0;
      if (found) {
        double sm = 0.0;
        for (int j = 0; j < D; j++) {
          xn[j] = P[(long long)bi * D + j];
          double tt = xn[j] - xrand[j];
          sm = sm + tt * tt;
        }
        double d = sqrt(sm);
        if (d > max_step) {
          double u = max_step / d;
          for (int j = 0; j < D; j++) xnew[j] = xn[j] + u * (xrand[j] - xn[j]);
        } else {
          for (int j = 0; j < D; j++) xnew[j] = xrand[j];
        }
        double l = 0.0;
        for (int j = 0; j < D; j++) {
          double tt = xn[j] - xnew[j];
          l = l + tt * tt;
        }
        elen = sqrt(l);
        ks = edge_samples(s, elen, res);
      }
    }
    __syncthreads();
    int ok = 1;
    if (parent < 0 || ks < 0) {
      ok = 0;
    } else {
      const long long total = ks + 2;
      const double kk2 = (double)(ks + 2);
      for (long long q = threadIdx.x; q < total && ok; q += blockDim.x) {
        double x[POMP_MAX_DIM];
        if (q == 0) {
          for (int j = 0; j < D; j++) x[j] = xn[j];
        } else if (q == 1) {
          for (int j = 0; j < D; j++) x[j] = xnew[j];
        } else {
          double u = (double)(q - 1) / kk2;
          for (int j = 0; j < D; j++) x[j] = xn[j] + u * (xnew[j] - xn[j]);
        }
        if (!feasible_pt(s, x)) ok = 0;
      }
    }
    ok = __syncthreads_and(ok);
    if (threadIdx.x == 0) {
      o->parent = parent;
      o->added = ok;
      o->overflow = 0;
      o->len = elen;
      for (int j = 0; j < D; j++) o->xnew[j] = parent >= 0 ? xnew[j] : nan("");
      if (ok) {
        for (int j = 0; j < D; j++) P[(long long)n * D + j] = xnew[j];
        n_sh = n + 1;
        if (has_goal) {  // goal ball test as the host does it: sqrt(sum (x - goal)^2) <= radius
          double sm = 0.0;
          for (int j = 0; j < D; j++) {
            double tt = xnew[j] - goal.v[j];
            sm = sm + tt * tt;
          }
          if (sqrt(sm) <= goal_radius) stop_sh = 1;
        }
      }
    }
    __syncthreads();
    if (stop_sh) break;
  }
  if (threadIdx.x == 0) {
    count[t] = n_sh;
    ndone[t] = done;
    __threadfence_system();
  }
}

POMP_API int pomp_forest_create(int dim, int n_trials, int64_t capacity_per_trial, int device, pomp_forest** out) {
  POMP_REQUIRE(out, "out is NULL");
  *out = nullptr;
  POMP_REQUIRE(dim >= 1 && dim <= POMP_MAX_DIM && n_trials >= 1 && capacity_per_trial >= 2, "bad forest shape");
  DeviceGuard g(device);
  if (!g.ok) return POMP_ERR_CUDA;
  pomp_forest* f = new pomp_forest();
  f->device = device;
  f->D = dim;
  f->T = n_trials;
  f->cap = capacity_per_trial;
  f->h_count.assign((size_t)n_trials, 0);
  cudaError_t e = cudaMalloc((void**)&f->d_pts, sizeof(double) * (size_t)n_trials * capacity_per_trial * dim);
  if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_count, sizeof(int) * (size_t)n_trials);
  if (e == cudaSuccess) e = cudaMemset(f->d_count, 0, sizeof(int) * (size_t)n_trials);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->in_h, sizeof(pomp_forest::In) * (size_t)n_trials, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->out_h, sizeof(pomp_forest::Out) * (size_t)n_trials, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&f->in_d, f->in_h, 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&f->out_d, f->out_h, 0);
  const size_t TM = (size_t)n_trials * pomp_forest::MAXM;
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->xr_h, sizeof(double) * TM * dim, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->outm_h, sizeof(pomp_forest::Out) * TM, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&f->ndone_h, sizeof(int32_t) * (size_t)n_trials, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&f->xr_d, f->xr_h, 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&f->outm_d, f->outm_h, 0);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&f->ndone_d, f->ndone_h, 0);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("pomp_forest_create: %s", cudaGetErrorString(e));
    pomp_forest_destroy(f);
    return e == cudaErrorMemoryAllocation ? POMP_ERR_NOMEM : POMP_ERR_CUDA;
  }
  memset(f->in_h, 0, sizeof(pomp_forest::In) * (size_t)n_trials);
  CUDA_TRY(cudaDeviceSynchronize());  // the default-stream memset above vs the non-blocking stream of the launches
  *out = f;
  return POMP_OK;
}

POMP_API int pomp_forest_destroy(pomp_forest* f) {
  if (!f) return POMP_OK;
  DeviceGuard g(f->device);
  if (f->d_pts) cudaFree(f->d_pts);
  if (f->d_count) cudaFree(f->d_count);
  if (f->in_h) cudaFreeHost(f->in_h);
  if (f->out_h) cudaFreeHost(f->out_h);
  if (f->xr_h) cudaFreeHost(f->xr_h);
  if (f->outm_h) cudaFreeHost(f->outm_h);
  if (f->ndone_h) cudaFreeHost(f->ndone_h);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return POMP_OK;
}

POMP_API int64_t pomp_forest_size(const pomp_forest* f, int trial) {
  return (f && trial >= 0 && trial < f->T) ? f->h_count[(size_t)trial] : -1;
}

POMP_API int pomp_forest_extend_h(pomp_forest* f, pomp_space* sp, const double* xrand_h, const double* root_h,
                                  const uint8_t* active_h, const uint8_t* reset_h, int check_sample, double max_step,
                                  double resolution, int64_t* parent_h, double* xnew_h, double* edge_len_h,
                                  int32_t* added_h) {
  POMP_REQUIRE(f && sp && xrand_h && root_h && parent_h && xnew_h && added_h, "NULL argument to pomp_forest_extend_h");
  POMP_REQUIRE(sp->dev.dim == f->D && sp->device == f->device, "space does not match the forest");
  POMP_REQUIRE(resolution > 0.0 && max_step >= 0.0, "bad step/resolution");
  DeviceGuard g(f->device);
  if (!g.ok) return POMP_ERR_CUDA;
  const int D = f->D, T = f->T;
  for (int t = 0; t < T; t++) {
    pomp_forest::In& in = f->in_h[t];
    for (int j = 0; j < D; j++) {
      in.xrand[j] = xrand_h[(size_t)t * D + j];
      in.root[j] = root_h[j];
    }
    in.active = active_h ? (active_h[t] != 0) : 1;
    in.reset = (reset_h ? (reset_h[t] != 0) : 0) || f->h_count[(size_t)t] == 0;
    if (in.reset) f->h_count[(size_t)t] = 1;
  }
  LAUNCH(forest_extend_kernel, T, 256, 0, f->stream, f->d_pts, f->d_count, (long long)f->cap, D, sp->dev, f->in_d,
         f->out_d, check_sample, max_step, resolution);
  CHECK_LAUNCH();
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  int overflow = 0;
  for (int t = 0; t < T; t++) {
    if (active_h && !active_h[t]) {
      parent_h[t] = -3;
      added_h[t] = 0;
      continue;
    }
    const pomp_forest::Out& o = f->out_h[t];
    parent_h[t] = o.parent;
    added_h[t] = o.added;
    if (edge_len_h) edge_len_h[t] = o.len;
    for (int j = 0; j < D; j++) xnew_h[(size_t)t * D + j] = o.xnew[j];
    if (o.added) f->h_count[(size_t)t] += 1;
    overflow |= o.overflow;
  }
  if (overflow) {
    set_error("a trial's tree reached the forest capacity of %lld nodes", (long long)f->cap);
    return POMP_ERR_CAPACITY;
  }
  return POMP_OK;
}

POMP_API int pomp_forest_extend_multi_h(pomp_forest* f, pomp_space* sp, int32_t n_iters, const double* xrand_h,
                                        const double* root_h, const uint8_t* active_h, const uint8_t* reset_h,
                                        int check_sample, double max_step, double resolution, const double* goal_h,
                                        double goal_radius, int32_t* n_done_h, int64_t* parent_h, double* xnew_h,
                                        double* edge_len_h, int32_t* added_h) {
  POMP_REQUIRE(f && sp && xrand_h && root_h && n_done_h && parent_h && xnew_h && added_h,
               "NULL argument to pomp_forest_extend_multi_h");
  POMP_REQUIRE(n_iters >= 1 && n_iters <= pomp_forest::MAXM, "n_iters=%d outside 1..%d", n_iters, pomp_forest::MAXM);
  POMP_REQUIRE(sp->dev.dim == f->D && sp->device == f->device, "space does not match the forest");
  POMP_REQUIRE(resolution > 0.0 && max_step >= 0.0, "bad step/resolution");
  DeviceGuard g(f->device);
  if (!g.ok) return POMP_ERR_CUDA;
  const int D = f->D, T = f->T, M = n_iters, MM = pomp_forest::MAXM;
  for (int t = 0; t < T; t++) {
    pomp_forest::In& in = f->in_h[t];
    for (int j = 0; j < D; j++) in.root[j] = root_h[j];
    in.active = active_h ? (active_h[t] != 0) : 1;
    in.reset = (reset_h ? (reset_h[t] != 0) : 0) || f->h_count[(size_t)t] == 0;
    if (in.reset && in.active) f->h_count[(size_t)t] = 1;
    if (!in.active) in.reset = 0;  // an inactive trial keeps its pending reset for its next active launch
    if (in.active)
      for (int i = 0; i < M; i++)
        for (int j = 0; j < D; j++) f->xr_h[((size_t)t * MM + i) * D + j] = xrand_h[((size_t)t * M + i) * D + j];
  }
  QueryArg ga;
  for (int j = 0; j < POMP_MAX_DIM; j++) ga.v[j] = (goal_h && j < D) ? goal_h[j] : 0.0;
  LAUNCH(forest_extend_multi_kernel, T, 256, 0, f->stream, f->d_pts, f->d_count, (long long)f->cap, D, sp->dev, f->in_d,
         f->xr_d, M, f->outm_d, f->ndone_d, check_sample, max_step, resolution, ga, goal_h ? 1 : 0, goal_radius);
  CHECK_LAUNCH();
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  int overflow = 0;
  for (int t = 0; t < T; t++) {
    if (active_h && !active_h[t]) {
      n_done_h[t] = 0;
      continue;
    }
    const int nd = f->ndone_h[t];
    n_done_h[t] = nd;
    for (int i = 0; i < nd; i++) {
      const pomp_forest::Out& o = f->outm_h[(size_t)t * MM + i];
      parent_h[(size_t)t * M + i] = o.parent;
      added_h[(size_t)t * M + i] = o.added;
      if (edge_len_h) edge_len_h[(size_t)t * M + i] = o.len;
      for (int j = 0; j < D; j++) xnew_h[((size_t)t * M + i) * D + j] = o.xnew[j];
      if (o.added) f->h_count[(size_t)t] += 1;
      overflow |= o.overflow;
    }
  }
  if (overflow) {
    set_error("a trial's tree reached the forest capacity of %lld nodes", (long long)f->cap);
    return POMP_ERR_CAPACITY;
  }
  return POMP_OK;
}
