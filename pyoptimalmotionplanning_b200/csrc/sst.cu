// sst.cu -- lock-step Stable-Sparse-RRT* forest (BASELINE config 5): T independent sst* trials, ONE BLOCK PER
// TRIAL, several planner iterations per launch.  Host side: sst.py (all random draws stay on the host, in the
// reference's order; the restart schedule of sst* is iteration-count driven and arrives as per-iteration radii /
// restart flags).
//
// Per trial and iteration (pomp/planners/rrtstarplanner.py):
//   pickNode        :390-410  radius scan over the trial's LIVE nodes: cheapest ACTIVE node within selectionRadius
//                             (d <= r, strict '<' on the cost: the first = lowest index wins ties), else the nearest
//                             active node
//   edge            :313-320  stepwise trajectory of the control space (CVControlSpace.trajectory,
//                             spaces/statespace.py:47-61, through euler_step) as a PiecewiseLinearInterpolator
//                             (interpolators.py:73-102), checked by EpsilonEdgeChecker's schedule (edgechecker.py:20-30);
//                             the block evaluates the samples in parallel
//   cost            :321      parent cost + objective.incremental = u[0] (TimeObjectiveFunction)
//   nodeLocallyBest :345-366  nearest witness; farther than witnessRadius -> new witness; else the node must beat the
//                             representative of SOME witness within witnessRadius
//   doPruning       :368-388  the NEAREST witness takes the node as representative (whatever `better` held: the loop of
//                             :339-340 re-queries the nearest witness every time); the old representative turns
//                             inactive and the chain of inactive leaves above it leaves the tree and the NN set
//   goal            :288-297  goal.contains(x) = cspace.distance(x, c) <= r; best cost = min over goal nodes
// Distances between states are the planner's metric (plain Euclidean for the `euclidean=True` problems: sequential
// sum of squares, one rounding per operation); the C-space geodesic (product metric) measures path lengths and the
// goal test, as in the reference.
#include "common.cuh"
#include "dyn.cuh"
#include "space.cuh"
#include <vector>

using namespace pomp;

namespace {

constexpr int SST_THREADS = 128;
constexpr int SST_MAXWP = 128;  // waypoints of one stepwise trajectory kept in shared memory

struct SstDev {
  int X, U, T, ncap, wcap;
  double* nx;       // [T][ncap][X]
  double* nc;       // [T][ncap]
  int* nparent;     // [T][ncap]
  int* nch;         // [T][ncap] live children
  uint8_t* nalive;  // in the tree and in the NN set
  uint8_t* nactive;
  double* wx;       // [T][wcap][X]
  int* wrep;        // [T][wcap]
  int* ncount;      // [T] nodes ever created since the last restart
  int* wcount;
  double* best;     // [T]
  int* none;        // [T] iterations in which pickNode found nothing (must stay 0)
  int* overflow;    // [T] arena / waypoint capacity exceeded
  const double* root;
};

__device__ __forceinline__ double euclid_rt(const double* a, const double* b, int X) {
  double s = 0.0;
  for (int j = 0; j < X; j++) {
    const double t = a[j] - b[j];
    s = s + t * t;
  }
  return sqrt(s);
}

__device__ void sst_reset_trial(const SstDev& d, int t) {
  for (int j = 0; j < d.X; j++) {
    d.nx[((long long)t * d.ncap) * d.X + j] = d.root[j];
    d.wx[((long long)t * d.wcap) * d.X + j] = d.root[j];
  }
  const long long n0 = (long long)t * d.ncap, w0 = (long long)t * d.wcap;
  d.nc[n0] = 0.0;
  d.nparent[n0] = -1;
  d.nch[n0] = 0;
  d.nalive[n0] = 1;
  d.nactive[n0] = 1;
  d.wrep[w0] = 0;
  d.ncount[t] = 1;
  d.wcount[t] = 1;
}

struct GoalArg {
  double c[POMP_MAX_DIM];
  double r;  // < 0: no goal
};

// block-wide arg-min on (key, index): smallest key, lowest index on ties; index -1 = nothing
__device__ __forceinline__ void block_argmin(double& key, int& idx, double* sk, int* si) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ok = __shfl_xor_sync(0xffffffffu, key, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (oi >= 0 && (idx < 0 || ok < key || (ok == key && oi < idx))) {
      key = ok;
      idx = oi;
    }
  }
  if (lane == 0) {
    sk[w] = key;
    si[w] = idx;
  }
  __syncthreads();
  key = sk[0];
  idx = si[0];
  for (int x = 1; x < SST_THREADS / 32; x++)
    if (si[x] >= 0 && (idx < 0 || sk[x] < key || (sk[x] == key && si[x] < idx))) {
      key = sk[x];
      idx = si[x];
    }
  __syncthreads();
}

__global__ void __launch_bounds__(SST_THREADS)
sst_step_kernel(SstDev d, SpaceDev s, pomp_dyn_desc dy, pomp_metric_desc geo, int M, const double* __restrict__ xr,
                const double* __restrict__ uu, const uint8_t* __restrict__ valid, const double* __restrict__ sel,
                const double* __restrict__ wit, const uint8_t* __restrict__ restart, int restart_now, GoalArg goal,
                double res, const double* __restrict__ half_sq_last, double half_sq_full) {
  const int t = blockIdx.x, tid = threadIdx.x;
  const int X = d.X, U = d.U;
  double* nx = d.nx + (long long)t * d.ncap * X;
  double* nc = d.nc + (long long)t * d.ncap;
  int* nparent = d.nparent + (long long)t * d.ncap;
  int* nch = d.nch + (long long)t * d.ncap;
  uint8_t* nalive = d.nalive + (long long)t * d.ncap;
  uint8_t* nactive = d.nactive + (long long)t * d.ncap;
  double* wx = d.wx + (long long)t * d.wcap * X;
  int* wrep = d.wrep + (long long)t * d.wcap;
  __shared__ double sk[SST_THREADS / 32];
  __shared__ int si[SST_THREADS / 32];
  __shared__ double path[SST_MAXWP][POMP_MAX_DIM];
  __shared__ double s_xrand[POMP_MAX_DIM], s_u[POMP_MAX_DIM];
  __shared__ int s_np, s_go;
  __shared__ long long s_k;
  if (restart_now) {
    if (tid == 0) sst_reset_trial(d, t);
    __syncthreads();
  }
  for (int m = 0; m < M; m++) {
    if (restart[m]) {
      if (tid == 0) sst_reset_trial(d, t);
    }
    if (tid < X) s_xrand[tid] = xr[((long long)t * M + m) * X + tid];
    if (tid < U) s_u[tid] = uu[((long long)t * M + m) * U + tid];
    __syncthreads();
    const int n = d.ncount[t], nw = d.wcount[t];
    const double rsel = sel[m], rwit = wit[m];
    // ---- pickNode: cheapest active node within the selection radius, else the nearest active node
    double ck = POMP_INF, dk = POMP_INF;
    int ci = -1, di = -1;
    for (int i = tid; i < n; i += SST_THREADS) {
      if (!nalive[i] || !nactive[i]) continue;
      const double dd = euclid_rt(nx + (long long)i * X, s_xrand, X);
      if (dd <= rsel && (ci < 0 || nc[i] < ck)) {
        ck = nc[i];
        ci = i;
      }
      if (dd < dk) {  // +inf / NaN distances are never admitted (d < dbest in the reference)
        dk = dd;
        di = i;
      }
    }
    block_argmin(ck, ci, sk, si);
    block_argmin(dk, di, sk, si);
    const int pick = ci >= 0 ? ci : di;
    if (pick < 0) {
      if (tid == 0) d.none[t] += 1;
      continue;  // (uniform: every thread sees the same pick)
    }
    if (!valid[(long long)t * M + m]) continue;
    // ---- stepwise trajectory + geodesic length (thread 0), then the sampled edge check by the whole block
    if (tid == 0) {
      const double duration = s_u[0];
      const double dt0 = euler_dt0(dy, duration);
      int np = 1;
      double l = 0.0, tt = 0.0;
      for (int j = 0; j < X; j++) path[0][j] = nx[(long long)pick * X + j];
      bool over = false;
      while (tt < duration) {
        if (np >= SST_MAXWP) {
          over = true;
          break;
        }
        const double dt = fmin(dt0, duration - tt);
        if (dy.kind == POMP_DYN_CV_EULER && half_sq_last) {
          // statespace.py:55-59 with the caller's 0.5*dt**2 (Python's ** is libm pow, which is not always dt*dt:
          // one ulp in a node coordinate is a different tree)
          const double hh = dt == dt0 ? half_sq_full : half_sq_last[(long long)t * M + m];
          const int nh = X / 2;
          for (int i = 0; i < nh; i++) {
            const double vnew = path[np - 1][nh + i] + dt * s_u[1 + i];
            const double m1 = path[np - 1][i] + hh * s_u[1 + i];
            const double m2 = path[np - 1][nh + i] * dt;
            path[np][i] = m1 + m2;
            path[np][nh + i] = vnew;
          }
        } else {
          euler_step(dy, path[np - 1], s_u, dt, path[np]);
        }
        l += metric_eval(geo, path[np - 1], path[np]);
        np++;
        tt = fmin(tt + dt0, duration);
      }
      long long k = edge_samples(s, l, res);
      int go = 1;
      if (over || k < 0) {
        d.overflow[t] = 1;
        go = 0;
      } else if (!feasible_pt(s, path[0]) || !feasible_pt(s, path[np - 1])) {
        go = 0;
      }
      s_np = np;
      s_k = k;
      s_go = go;
    }
    __syncthreads();
    if (!s_go) continue;
    const int np = s_np;
    const long long k = s_k;
    int ok = 1;
    {
      const double kk2 = (double)(k + 2);
      for (long long i = tid; i < k && ok; i += SST_THREADS) {
        const double uq = (double)(i + 1) / kk2;
        double pt[POMP_MAX_DIM];
        if (np == 1) {
          for (int j = 0; j < X; j++) pt[j] = path[0][j];
        } else {
          const double kk = uq * (double)(np - 1);
          const double fl = floor(kk);
          const int seg = (int)fl;
          geodesic_interp(geo, path[seg], path[seg + 1], kk - fl, X, pt);
        }
        if (!feasible_pt(s, pt)) ok = 0;
      }
    }
    ok = __syncthreads_and(ok);
    if (!ok) continue;
    const double* xnew = path[np - 1];
    const double newcost = nc[pick] + s_u[0];
    // ---- nodeLocallyBest: nearest witness, and whether the node beats some representative within the radius
    double wk = POMP_INF;
    int wi = -1, better = 0;
    for (int j = tid; j < nw; j += SST_THREADS) {
      const double dd = euclid_rt(wx + (long long)j * X, xnew, X);
      if (dd < wk) {
        wk = dd;
        wi = j;
      }
      if (dd <= rwit) {
        const int r = wrep[j];
        if (r < 0 || newcost < nc[r]) better = 1;
      }
    }
    block_argmin(wk, wi, sk, si);
    better = __syncthreads_or(better);
    if (tid == 0) {
      int go = 1;
      int w = wi;
      if (wi < 0 || wk > rwit) {  // no witness nearby: the node founds one
        if (nw >= d.wcap) {
          d.overflow[t] = 1;
          go = 0;
        } else {
          for (int j = 0; j < X; j++) wx[(long long)nw * X + j] = xnew[j];
          wrep[nw] = -1;
          w = nw;
          d.wcount[t] = nw + 1;
        }
      } else if (!better) {
        go = 0;  // dominated: the node is dropped
      }
      if (go && n >= d.ncap) {
        d.overflow[t] = 1;
        go = 0;
      }
      if (go) {
        for (int j = 0; j < X; j++) nx[(long long)n * X + j] = xnew[j];
        nc[n] = newcost;
        nparent[n] = pick;
        nch[n] = 0;
        nalive[n] = 1;
        nactive[n] = 1;
        nch[pick] += 1;
        d.ncount[t] = n + 1;
        // doPruning: the nearest witness takes the node; its old representative and the inactive leaf chain go
        int npeer = wrep[w];
        wrep[w] = n;
        if (npeer >= 0 && npeer != n) {
          nactive[npeer] = 0;
          while (npeer >= 0 && nch[npeer] == 0 && !nactive[npeer]) {
            const int p = nparent[npeer];
            nalive[npeer] = 0;
            if (p >= 0) nch[p] -= 1;
            nparent[npeer] = -1;
            npeer = p;
          }
        }
        if (goal.r >= 0.0 && metric_eval(geo, xnew, goal.c) <= goal.r && newcost < d.best[t]) d.best[t] = newcost;
      }
    }
    __syncthreads();
  }
}

}  // namespace

// ============================================================================ host
struct pomp_sst {
  int device = 0;
  SstDev dev;
  pomp_dyn_desc dyn;
  pomp_metric_desc geo;
  DevBuf<double> nx, nc, wx, best, root, in_x, in_u, in_sel, in_wit, in_hh;
  DevBuf<int> nparent, nch, wrep, ncount, wcount, none, overflow;
  DevBuf<uint8_t> nalive, nactive, in_valid, in_restart;
  cudaStream_t stream = nullptr;
  std::vector<double> h_best;
  std::vector<int> h_cnt;
};

POMP_API int pomp_sst_create(int32_t x_dim, int32_t u_dim, int32_t n_trials, int32_t node_capacity,
                             int32_t witness_capacity, const pomp_dyn_desc* dyn, const pomp_metric_desc* geodesic,
                             int device, pomp_sst** out) {
  POMP_REQUIRE(out && dyn && geodesic, "NULL argument to pomp_sst_create");
  POMP_REQUIRE(x_dim >= 1 && x_dim <= POMP_MAX_DIM && u_dim >= 1 && u_dim <= POMP_MAX_DIM && n_trials >= 1 &&
                   node_capacity >= 2 && witness_capacity >= 2, "bad sizes for pomp_sst_create");
  POMP_REQUIRE(dyn->x_dim == x_dim && dyn->u_dim == u_dim && geodesic->dim == x_dim, "descriptor dimensions disagree");
  POMP_TRY(validate_dyn(dyn));
  POMP_TRY(validate_metric(geodesic));
  DeviceGuard g(device);
  if (!g.ok) return POMP_ERR_CUDA;
  pomp_sst* f = new pomp_sst();
  f->device = device;
  f->dyn = *dyn;
  f->geo = *geodesic;
  const size_t T = (size_t)n_trials, nc = (size_t)node_capacity, wc = (size_t)witness_capacity, X = (size_t)x_dim;
  int rc = POMP_OK;
  auto R = [&](int r) { if (rc == POMP_OK) rc = r; };
  R(f->nx.reserve(T * nc * X));
  R(f->nc.reserve(T * nc));
  R(f->nparent.reserve(T * nc));
  R(f->nch.reserve(T * nc));
  R(f->nalive.reserve(T * nc));
  R(f->nactive.reserve(T * nc));
  R(f->wx.reserve(T * wc * X));
  R(f->wrep.reserve(T * wc));
  R(f->ncount.reserve(T));
  R(f->wcount.reserve(T));
  R(f->best.reserve(T));
  R(f->none.reserve(T));
  R(f->overflow.reserve(T));
  R(f->root.reserve(POMP_MAX_DIM));
  if (rc == POMP_OK && cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess) rc = POMP_ERR_CUDA;
  if (rc != POMP_OK) {
    delete f;
    return rc;
  }
  SstDev& d = f->dev;
  d.X = x_dim;
  d.U = u_dim;
  d.T = n_trials;
  d.ncap = node_capacity;
  d.wcap = witness_capacity;
  d.nx = f->nx.p;
  d.nc = f->nc.p;
  d.nparent = f->nparent.p;
  d.nch = f->nch.p;
  d.nalive = f->nalive.p;
  d.nactive = f->nactive.p;
  d.wx = f->wx.p;
  d.wrep = f->wrep.p;
  d.ncount = f->ncount.p;
  d.wcount = f->wcount.p;
  d.best = f->best.p;
  d.none = f->none.p;
  d.overflow = f->overflow.p;
  d.root = f->root.p;
  f->h_best.assign(T, INFINITY);
  f->h_cnt.assign(3 * T, 0);
  *out = f;
  return POMP_OK;
}

POMP_API int pomp_sst_destroy(pomp_sst* f) {
  if (!f) return POMP_OK;
  DeviceGuard g(f->device);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return POMP_OK;
}

static __global__ void sst_reset_kernel(SstDev d) {  // (after the anonymous namespace: uses its device helpers)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.T) return;
  sst_reset_trial(d, t);
  d.best[t] = POMP_INF;
  d.none[t] = 0;
  d.overflow[t] = 0;
}

POMP_API int pomp_sst_reset_h(pomp_sst* f, const double* root_h) {
  POMP_REQUIRE(f && root_h, "NULL argument to pomp_sst_reset_h");
  DeviceGuard g(f->device);
  if (!g.ok) return POMP_ERR_CUDA;
  CUDA_TRY(cudaMemcpyAsync(f->root.p, root_h, sizeof(double) * f->dev.X, cudaMemcpyHostToDevice, f->stream));
  LAUNCH(sst_reset_kernel, (f->dev.T + 127) / 128, 128, 0, f->stream, f->dev);
  CHECK_LAUNCH();
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return POMP_OK;
}

POMP_API int pomp_sst_step_h(pomp_sst* f, pomp_space* sp, int32_t n_iters, const double* xrand_h, const double* u_h,
                             const uint8_t* valid_h, const double* sel_radius_h, const double* wit_radius_h,
                             const uint8_t* restart_h, int32_t restart_now, const double* goal_center_h,
                             double goal_radius, double resolution, const double* half_sq_last_h, double half_sq_full,
                             double* best_cost_h, int32_t* n_nodes_h, int32_t* n_none_h) {
  POMP_REQUIRE(f && sp && best_cost_h && n_nodes_h && n_none_h && n_iters >= 0, "bad arguments to pomp_sst_step_h");
  POMP_REQUIRE(n_iters == 0 || (xrand_h && u_h && valid_h && sel_radius_h && wit_radius_h && restart_h),
               "NULL sample arrays for pomp_sst_step_h");
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  POMP_REQUIRE(sp->device == f->device && sp->dev.dim == f->dev.X, "space does not match the forest");
  DeviceGuard g(f->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = f->stream;
  const size_t T = (size_t)f->dev.T, M = (size_t)n_iters, X = (size_t)f->dev.X, U = (size_t)f->dev.U;
  if (M) {
    POMP_TRY(f->in_x.reserve(T * M * X));
    POMP_TRY(f->in_u.reserve(T * M * U));
    POMP_TRY(f->in_valid.reserve(T * M));
    POMP_TRY(f->in_sel.reserve(M));
    POMP_TRY(f->in_wit.reserve(M));
    POMP_TRY(f->in_restart.reserve(M));
    CUDA_TRY(cudaMemcpyAsync(f->in_x.p, xrand_h, sizeof(double) * T * M * X, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(f->in_u.p, u_h, sizeof(double) * T * M * U, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(f->in_valid.p, valid_h, T * M, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(f->in_sel.p, sel_radius_h, sizeof(double) * M, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(f->in_wit.p, wit_radius_h, sizeof(double) * M, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(f->in_restart.p, restart_h, M, cudaMemcpyHostToDevice, st));
    if (half_sq_last_h) {
      POMP_TRY(f->in_hh.reserve(T * M));
      CUDA_TRY(cudaMemcpyAsync(f->in_hh.p, half_sq_last_h, sizeof(double) * T * M, cudaMemcpyHostToDevice, st));
    }
  }
  GoalArg ga;
  for (int j = 0; j < POMP_MAX_DIM; j++) ga.c[j] = (goal_center_h && j < (int)X) ? goal_center_h[j] : 0.0;
  ga.r = goal_center_h ? goal_radius : -1.0;
  LAUNCH(sst_step_kernel, (int)T, SST_THREADS, 0, st, f->dev, sp->dev, f->dyn, f->geo, (int)M, f->in_x.p, f->in_u.p,
         f->in_valid.p, f->in_sel.p, f->in_wit.p, f->in_restart.p, restart_now, ga, resolution,
         (const double*)((M && half_sq_last_h) ? f->in_hh.p : nullptr), half_sq_full);
  CHECK_LAUNCH();
  CUDA_TRY(cudaMemcpyAsync(best_cost_h, f->best.p, sizeof(double) * T, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(n_none_h, f->none.p, sizeof(int) * T, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(f->h_cnt.data(), f->overflow.p, sizeof(int) * T, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(n_nodes_h, f->ncount.p, sizeof(int) * T, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  POMP_TRY(space_overflow_check(sp));
  for (size_t t = 0; t < T; t++)
    if (f->h_cnt[t]) {
      set_error("sst forest: trial %zu ran out of node / witness / waypoint capacity", t);
      return POMP_ERR_CAPACITY;
    }
  return POMP_OK;
}

POMP_API int pomp_sst_get_trial_h(pomp_sst* f, int32_t trial, double* x_h, int32_t* parent_h, uint8_t* alive_h,
                                  double* cost_h, uint8_t* active_h, int32_t* count_h) {
  POMP_REQUIRE(f && count_h && trial >= 0 && trial < f->dev.T, "bad arguments to pomp_sst_get_trial_h");
  DeviceGuard g(f->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = f->stream;
  int n = 0;
  CUDA_TRY(cudaMemcpyAsync(&n, f->ncount.p + trial, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *count_h = n;
  const size_t o = (size_t)trial * f->dev.ncap;
  if (x_h) CUDA_TRY(cudaMemcpyAsync(x_h, f->nx.p + o * f->dev.X, sizeof(double) * n * f->dev.X, cudaMemcpyDeviceToHost, st));
  if (parent_h) CUDA_TRY(cudaMemcpyAsync(parent_h, f->nparent.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  if (alive_h) CUDA_TRY(cudaMemcpyAsync(alive_h, f->nalive.p + o, n, cudaMemcpyDeviceToHost, st));
  if (cost_h) CUDA_TRY(cudaMemcpyAsync(cost_h, f->nc.p + o, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  if (active_h) CUDA_TRY(cudaMemcpyAsync(active_h, f->nactive.p + o, n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return POMP_OK;
}
