// space.cu -- ConfigurationSpace.feasible and EpsilonEdgeChecker.feasible on the device.
//
// EpsilonEdgeChecker (pomp/spaces/edgechecker.py:20-30) is a SAMPLED check, not a swept one:
// l = length(); k = ceil(l/res); endpoints first; then eval((i+1)/(k+2)) for i in [0,k).  The
// kernels reproduce that schedule exactly (one edge per thread, early exit on the first
// infeasible sample), so booleans match the reference even where a swept test would differ.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "dyn.cuh"
#include "space.cuh"

using namespace pomp;

// ============================================================================ kernels
__global__ void feasible_kernel(SpaceDev s, const double* __restrict__ x, long long n, uint8_t* __restrict__ ok) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p[POMP_MAX_DIM];
  for (int j = 0; j < s.dim; j++) p[j] = x[i * s.dim + j];
  ok[i] = feasible_pt(s, p) ? 1 : 0;
}

// LinearInterpolator edge a->b (interpolators.py:28-33): length = vectorops.distance,
// eval(u) = a + u*(b-a) (vectorops.py:14-18,115-117)
template <int D>
__device__ __forceinline__ bool edge_linear(const SpaceDev& s, const double* a, const double* b, double res) {
  constexpr int DA = D ? D : POMP_MAX_DIM;
  const int dim = D ? D : s.dim;
  double l = 0.0;
#pragma unroll
  for (int j = 0; j < DA; j++)
    if (j < dim) {
      double t = a[j] - b[j];
      l = l + t * t;
    }
  l = sqrt(l);
  long long k = edge_samples(s, l, res);
  if (k < 0) return false;
  if (!feasible_pt(s, a) || !feasible_pt(s, b)) return false;
  double dlt[DA];
#pragma unroll
  for (int j = 0; j < DA; j++) dlt[j] = j < dim ? b[j] - a[j] : 0.0;
  const double kk2 = (double)(k + 2);
  for (long long i = 0; i < k; i++) {
    double u = (double)(i + 1) / kk2;
    double x[DA];
#pragma unroll
    for (int j = 0; j < DA; j++) x[j] = j < dim ? a[j] + u * dlt[j] : 0.0;
    if (!feasible_pt(s, x)) return false;
  }
  return true;
}

// ---- 2-D linear edges, warp-cooperative --------------------------------------------------------
// EpsilonEdgeChecker's result is a pure conjunction over the endpoints and the k interior samples,
// so the samples may be evaluated in any order.  Thread-per-edge evaluation leaves ~3 of 32 lanes
// busy on obstacle-dense roadmaps (edges die at different samples); here a warp takes 32 edges,
// checks all endpoints at once, then lays the interior samples of the surviving edges out as one
// item queue (edge-major) and evaluates 32 items per round, one per lane.
__device__ __forceinline__ int warp_incl_scan_i(int v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// EpsilonEdgeChecker's result is a pure conjunction over the endpoints and the k interior samples, so they may be
// evaluated in any order -- used three times here:
//   * the samples are laid out SAMPLE-MAJOR over the edges that are still alive: a round takes the next S samples of
//     each of the A alive edges (A * S <= 64 items, two per lane), then drops the edges that failed or are complete
//     (edge-major order evaluated all k samples of an edge whose second sample already hits an obstacle: at res 0.001,
//     150 samples per edge, 23 rounds per warp instead of 2);
//   * a round only CLASSIFIES its points with the fine occupancy map (one load): empty cells pass, full cells fail;
//   * the points of partially covered cells are parked in a per-warp queue and tested against their buckets 64 at a
//     time (two dependent loads, all lanes busy, two points per lane), while their edges go on sampling.
constexpr int EQ_CAP = 160;  // queue entries per warp; popped when fewer than 64 slots (one round's worth) are free
struct EdgeQueue {
  double x[EQ_CAP], y[EQ_CAP];
  int owner[EQ_CAP];
  unsigned char rank2lane[32];
};

__device__ __forceinline__ void edge_queue_push(EdgeQueue& Q, int& qn, bool push, double px, double py, int owner) {
  const unsigned m = __ballot_sync(0xffffffffu, push);
  if (push) {
    const int at = qn + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
    Q.x[at] = px;
    Q.y[at] = py;
    Q.owner[at] = owner;
  }
  qn += __popc(m);
}

// tests the top min(qn, 64) queued points against their buckets, two per lane with their range loads in flight
// together; points of edges that failed meanwhile are skipped
__device__ __forceinline__ void edge_queue_pop64(const SpaceDev& s, EdgeQueue& Q, int& qn, unsigned* failbits) {
  __syncwarp();
  const int lane = threadIdx.x & 31;
  const int take = min(qn, 64);
  const unsigned fb = *(volatile unsigned*)failbits;
  bool act[2];
  double px[2], py[2];
  int ow[2], b[2], e[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int t = h * 32 + lane;
    act[h] = t < take;
    ow[h] = 0;
    px[h] = py[h] = 0.0;
    b[h] = e[h] = 0;
    if (act[h]) {
      const int at = qn - take + t;
      ow[h] = Q.owner[at];
      px[h] = Q.x[at];
      py[h] = Q.y[at];
      act[h] = !((fb >> ow[h]) & 1u);
      if (act[h]) bucket_range(s, px[h], py[h], b[h], e[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < 2; h++)
    if (act[h] && bucket_scan(s, b[h], e[h], px[h], py[h])) atomicOr(failbits, 1u << ow[h]);
  qn -= take;
  __syncwarp();
}

// ep_known: -1 = test the endpoints here; 0 / 1 = ConfigurationSpace.feasible of BOTH endpoints is already known
// (roadmap edges: every node is an endpoint of ~2k edges, its test is done once per node instead)
__device__ __forceinline__ bool edge_batch_warp2d(const SpaceDev& s, double2 a, double2 b, bool valid, double res,
                                                  unsigned* failbits /* one word per warp, shared */, EdgeQueue& Q,
                                                  int ep_known = -1) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double pa[2] = {a.x, a.y}, pb[2] = {b.x, b.y};
  double t0 = pa[0] - pb[0], t1 = pa[1] - pb[1];
  double l = 0.0;
  l = l + t0 * t0;
  l = l + t1 * t1;
  l = sqrt(l);
  long long k = valid ? edge_samples(s, l, res) : 0;
  const double dx = pb[0] - pa[0], dy = pb[1] - pa[1];
  const bool big = k > 4096;  // pathological resolution: this lane walks its own edge afterwards
  // without an obstacle grid (a handful of obstacles) or with NaN coordinates the plain loop decides at once
  const bool mapped = s.use_grid != 0;
  // BoxSet.contains on both coordinates (sets.py:177-182); NaN passes, as in Python
  auto inb = [&](double x0, double x1) { return !(x0 < s.lo[0] || x0 > s.hi[0] || x1 < s.lo[1] || x1 > s.hi[1]); };
  bool alive = valid && k >= 0 && (ep_known >= 0 ? ep_known != 0 : (inb(pa[0], pa[1]) && inb(pb[0], pb[1])));
  if (__all_sync(FULL, !valid || (k >= 0 && k <= 2))) {
    // a warp of short edges (roadmap edges to near neighbours: one or two samples each): the queue machinery costs
    // more than it saves, every lane walks its own edge
    bool okv = alive;
    if (okv && ep_known < 0) okv = !hits_obstacle(s, s.od0 ? pa[1] : pa[0], s.od1 ? pa[1] : pa[0]);
    if (okv && ep_known < 0) okv = !hits_obstacle(s, s.od0 ? pb[1] : pb[0], s.od1 ? pb[1] : pb[0]);
    for (int i = 0; okv && i < (int)k; i++) {
      const double u = (double)(i + 1) / (double)(k + 2);
      const double x0 = pa[0] + u * dx, x1 = pa[1] + u * dy;
      okv = inb(x0, x1) && !hits_obstacle(s, s.od0 ? x1 : x0, s.od1 ? x1 : x0);
    }
    return okv;
  }
  if (lane == 0) *failbits = 0u;
  __syncwarp();
  int qn = 0;
  unsigned char* rank2lane = Q.rank2lane;
  // ---- endpoints: two more members of the conjunction
  {
    bool push[2] = {false, false};
    double epx[2], epy[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const double e0 = h ? pb[0] : pa[0], e1 = h ? pb[1] : pa[1];
      const double px = s.od0 ? e1 : e0, py = s.od1 ? e1 : e0;
      epx[h] = px;
      epy[h] = py;
      if (alive && ep_known < 0) {
        if (!mapped || isnan(px) || isnan(py)) {
          if (hits_obstacle(s, px, py)) alive = false;
        } else {
          const int cls = obs_class(s, px, py);
          if (cls == OBS_FULL) alive = false;
          push[h] = cls == OBS_PART;
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
      edge_queue_push(Q, qn, push[h] && alive, epx[h], epy[h], lane);
  }
  const int cnt = (alive && !big) ? (int)k : 0;
  unsigned live = __ballot_sync(FULL, cnt > 0) & ~(*(volatile unsigned*)failbits);  // edges with samples left
  int sbase = 0;                                                                    // samples of every live edge done
  while (live) {
    const int A = __popc(live);
    const int S = 64 / A;  // >= 2
    const float invA = __frcp_rn((float)A);
    // lane of the r-th live edge (a table instead of __fns, which is a software loop)
    if ((live >> lane) & 1u) rank2lane[__popc(live & ((1u << lane) - 1u))] = (unsigned char)lane;
    __syncwarp();
    bool push[2];
    double qx[2], qy[2];
    int own[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int j = h * 32 + lane;
      int sj = (int)(__int2float_rz(j) * invA);  // j / A, off by at most one
      int r = j - sj * A;
      if (r < 0) {
        r += A;
        sj--;
      } else if (r >= A) {
        r -= A;
        sj++;
      }
      own[h] = rank2lane[r];
      const int smp = sbase + sj;
      const int ok_ = __shfl_sync(FULL, cnt, own[h]);
      const double ax = __shfl_sync(FULL, pa[0], own[h]), ay = __shfl_sync(FULL, pa[1], own[h]);
      const double ddx = __shfl_sync(FULL, dx, own[h]), ddy = __shfl_sync(FULL, dy, own[h]);
      push[h] = false;
      qx[h] = qy[h] = 0.0;
      if (j < A * S && smp < ok_) {
        const double u = (double)(smp + 1) / (double)(ok_ + 2);
        const double x0 = ax + u * ddx, x1 = ay + u * ddy;
        const double px = s.od0 ? x1 : x0, py = s.od1 ? x1 : x0;
        bool hit;
        if (!inb(x0, x1)) {
          hit = true;
        } else if (!mapped || isnan(px) || isnan(py)) {
          hit = hits_obstacle(s, px, py);
        } else {
          const int cls = obs_class(s, px, py);
          hit = cls == OBS_FULL;
          push[h] = cls == OBS_PART;
          qx[h] = px;
          qy[h] = py;
        }
        if (hit) atomicOr(failbits, 1u << own[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) edge_queue_push(Q, qn, push[h], qx[h], qy[h], own[h]);
    if (qn > EQ_CAP - 64) edge_queue_pop64(s, Q, qn, failbits);
    __syncwarp();
    sbase += S;
    live &= ~(*(volatile unsigned*)failbits);
    live &= __ballot_sync(FULL, cnt > sbase);
  }
  while (qn > 0) edge_queue_pop64(s, Q, qn, failbits);
  __syncwarp();
  const bool fail = ((*(volatile unsigned*)failbits) >> lane) & 1u;
  __syncwarp();
  bool okv = alive && !fail;
  if (okv && big) {
    const double kk2 = (double)(k + 2);
    for (long long i = 0; i < k; i++) {
      double u = (double)(i + 1) / kk2;
      double x[2] = {pa[0] + u * dx, pa[1] + u * dy};
      if (!feasible_pt(s, x)) {
        okv = false;
        break;
      }
    }
  }
  return okv;
}

__global__ void __launch_bounds__(128)
edge_check_warp2d_kernel(SpaceDev s, const double* __restrict__ a, const double* __restrict__ b, long long n,
                         double res, uint8_t* __restrict__ ok) {
  __shared__ unsigned failbits[4];
  __shared__ EdgeQueue queues[4];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = i < n;
  double2 va = make_double2(0, 0), vb = make_double2(0, 0);
  if (valid) {
    va = __ldg(reinterpret_cast<const double2*>(a) + i);
    vb = __ldg(reinterpret_cast<const double2*>(b) + i);
  }
  bool r = edge_batch_warp2d(s, va, vb, valid, res, &failbits[threadIdx.x >> 5], queues[threadIdx.x >> 5]);
  if (valid) ok[i] = r ? 1 : 0;
}

__global__ void __launch_bounds__(128)
edge_check_indexed_warp2d_kernel(SpaceDev s, const double* __restrict__ nodes, const int* __restrict__ ij,
                                 long long n, double res, uint8_t* __restrict__ ok,
                                 const uint8_t* __restrict__ node_ok) {
  __shared__ unsigned failbits[4];
  __shared__ EdgeQueue queues[4];
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = e < n;
  int epk = -1;
  double2 va = make_double2(0, 0), vb = make_double2(0, 0);
  if (valid) {
    int2 pr = __ldg(reinterpret_cast<const int2*>(ij) + e);
    va = __ldg(reinterpret_cast<const double2*>(nodes) + pr.x);
    vb = __ldg(reinterpret_cast<const double2*>(nodes) + pr.y);
    if (node_ok) epk = (__ldg(node_ok + pr.x) && __ldg(node_ok + pr.y)) ? 1 : 0;
  }
  bool r = edge_batch_warp2d(s, va, vb, valid, res, &failbits[threadIdx.x >> 5], queues[threadIdx.x >> 5], epk);
  if (valid) ok[e] = r ? 1 : 0;
}

template <int D>
__global__ void edge_check_kernel(SpaceDev s, const double* __restrict__ a, const double* __restrict__ b,
                                  long long n, double res, uint8_t* __restrict__ ok) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int DA = D ? D : POMP_MAX_DIM;
  const int dim = D ? D : s.dim;
  double pa[DA], pb[DA];
  if (D == 2) {
    double2 va = __ldg(reinterpret_cast<const double2*>(a) + i);
    double2 vb = __ldg(reinterpret_cast<const double2*>(b) + i);
    pa[0] = va.x;
    pa[1] = va.y;
    pb[0] = vb.x;
    pb[1] = vb.y;
  } else {
    for (int j = 0; j < dim; j++) {
      pa[j] = a[i * dim + j];
      pb[j] = b[i * dim + j];
    }
  }
  ok[i] = edge_linear<D>(s, pa, pb, res) ? 1 : 0;
}

template <int D>
__global__ void edge_check_indexed_kernel(SpaceDev s, const double* __restrict__ nodes, const int* __restrict__ ij,
                                          long long n, double res, uint8_t* __restrict__ ok) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  constexpr int DA = D ? D : POMP_MAX_DIM;
  const int dim = D ? D : s.dim;
  int2 pr = __ldg(reinterpret_cast<const int2*>(ij) + e);
  double pa[DA], pb[DA];
  for (int j = 0; j < dim; j++) {
    pa[j] = __ldg(nodes + (long long)pr.x * dim + j);
    pb[j] = __ldg(nodes + (long long)pr.y * dim + j);
  }
  ok[e] = edge_linear<D>(s, pa, pb, res) ? 1 : 0;
}

// PiecewiseLinearInterpolator (interpolators.py:73-102) over explicit waypoints
__global__ void edge_check_path_kernel(SpaceDev s, pomp_metric_desc g, const double* __restrict__ wp,
                                       const int* __restrict__ off, long long n_paths, double res,
                                       uint8_t* __restrict__ ok) {
  long long pi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= n_paths) return;
  const int D = s.dim;
  const double* P = wp + (long long)off[pi] * D;
  int np = off[pi + 1] - off[pi];
  if (np <= 0) {
    ok[pi] = 0;
    return;
  }
  double a[POMP_MAX_DIM], b[POMP_MAX_DIM], x[POMP_MAX_DIM];
  double l = 0.0;  // l = 0; l += distance(path[i], path[i+1])  (:88-92)
  for (int j = 0; j < D; j++) a[j] = P[j];
  for (int i = 0; i + 1 < np; i++) {
    for (int j = 0; j < D; j++) b[j] = P[(long long)(i + 1) * D + j];
    l += metric_eval(g, a, b);
    for (int j = 0; j < D; j++) a[j] = b[j];
  }
  long long k = edge_samples(s, l, res);
  for (int j = 0; j < D; j++) {
    a[j] = P[j];
    b[j] = P[(long long)(np - 1) * D + j];
  }
  if (k < 0 || !feasible_pt(s, a) || !feasible_pt(s, b)) {
    ok[pi] = 0;
    return;
  }
  const double kk2 = (double)(k + 2);
  for (long long i = 0; i < k; i++) {
    double u = (double)(i + 1) / kk2;
    // eval(u) (:93-102); u is strictly inside (0,1) here
    double kk = u * (double)(np - 1);
    double fl = floor(kk);
    double sfrac = kk - fl;
    long long seg = (long long)fl;
    for (int j = 0; j < D; j++) {
      a[j] = P[seg * D + j];
      b[j] = P[(seg + 1) * D + j];
    }
    geodesic_interp(g, a, b, sfrac, D, x);
    if (!feasible_pt(s, x)) {
      ok[pi] = 0;
      return;
    }
  }
  ok[pi] = 1;
}

// controlSpace.interpolator(x,u) + EpsilonEdgeChecker without materialising the curve
__global__ void edge_check_control_kernel(SpaceDev s, pomp_dyn_desc dy, pomp_metric_desc g,
                                          const double* __restrict__ xs, const double* __restrict__ us, long long n,
                                          double res, uint8_t* __restrict__ ok) {
  long long ei = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ei >= n) return;
  const int X = dy.x_dim, U = dy.u_dim;
  double x[POMP_MAX_DIM], u[POMP_MAX_DIM], cur[POMP_MAX_DIM], nxt[POMP_MAX_DIM], pt[POMP_MAX_DIM];
  for (int j = 0; j < X; j++) x[j] = xs[ei * X + j];
  for (int j = 0; j < U; j++) u[j] = us[ei * U + j];
  if (dy.kind == POMP_DYN_ADAPTOR) {
    ok[ei] = edge_linear<0>(s, x, u, res) ? 1 : 0;
    return;
  }
  if (dy.kind == POMP_DYN_DUBINS) {  // DubinsCarInterpolator (dubins.py:62-73)
    dubins_next(x, u[0], u[1], nxt);
    long long k = edge_samples(s, fabs(u[0]), res);
    bool good = k >= 0 && feasible_pt(s, x) && feasible_pt(s, nxt);
    const double kk2 = (double)(k + 2);
    for (long long i = 0; good && i < k; i++) {
      double uu = (double)(i + 1) / kk2;
      dubins_next(x, u[0] * uu, u[1], pt);
      good = feasible_pt(s, pt);
    }
    ok[ei] = good ? 1 : 0;
    return;
  }
  // integrated kinds: pass 1 = waypoint count + geodesic length
  double duration = u[0];
  double dt0 = euler_dt0(dy, duration);
  long long np = 1;
  double l = 0.0, t = 0.0;
  for (int j = 0; j < X; j++) cur[j] = x[j];
  while (t < duration) {
    double dt = fmin(dt0, duration - t);
    euler_step(dy, cur, u, dt, nxt);
    l += metric_eval(g, cur, nxt);
    for (int j = 0; j < X; j++) cur[j] = nxt[j];
    np++;
    t = fmin(t + dt0, duration);
  }
  long long k = edge_samples(s, l, res);
  if (k < 0 || !feasible_pt(s, x) || !feasible_pt(s, cur)) {
    ok[ei] = 0;
    return;
  }
  // pass 2: samples are increasing in u, so one more forward sweep serves them all
  long long seg_cur = 0;
  t = 0.0;
  for (int j = 0; j < X; j++) cur[j] = x[j];
  if (np > 1) {
    double dt = fmin(dt0, duration - t);
    euler_step(dy, cur, u, dt, nxt);
    t = fmin(t + dt0, duration);
  }
  const double kk2 = (double)(k + 2);
  for (long long i = 0; i < k; i++) {
    double uu = (double)(i + 1) / kk2;
    double kk = uu * (double)(np - 1);
    double fl = floor(kk);
    double sfrac = kk - fl;
    long long seg = (long long)fl;
    while (seg_cur < seg) {  // advance [cur,nxt] to waypoints [seg, seg+1]
      for (int j = 0; j < X; j++) cur[j] = nxt[j];
      double dt = fmin(dt0, duration - t);
      euler_step(dy, cur, u, dt, nxt);
      t = fmin(t + dt0, duration);
      seg_cur++;
    }
    geodesic_interp(g, cur, nxt, sfrac, X, pt);
    if (!feasible_pt(s, pt)) {
      ok[ei] = 0;
      return;
    }
  }
  ok[ei] = 1;
}

// ONE straight edge of a planner loop (EpsilonEdgeChecker.feasible(LinearInterpolator), edgechecker.py:20-30): the
// endpoints travel in the kernel arguments, one block evaluates the reference's samples in parallel (a conjunction:
// order is free), the answer lands in mapped host memory -- one launch + one synchronisation per call.  (The batch
// path for n = 1: two pageable copies in, the kernel, one copy out.)
struct EdgeArg {
  double a[POMP_MAX_DIM], b[POMP_MAX_DIM];
};
__global__ void __launch_bounds__(128)
edge_one_kernel(SpaceDev s, EdgeArg e, double res, int* __restrict__ out) {
  __shared__ int fail;
  const int dim = s.dim;
  if (threadIdx.x == 0) fail = 0;
  __syncthreads();
  double l = 0.0;
  for (int j = 0; j < dim; j++) {
    const double t = e.a[j] - e.b[j];
    l = l + t * t;
  }
  l = sqrt(l);
  const long long k = edge_samples(s, l, res);
  if (k < 0) {
    if (threadIdx.x == 0) {
      *out = 0;
      __threadfence_system();
    }
    return;
  }
  if (threadIdx.x == 0 && !feasible_pt(s, e.a)) fail = 1;
  if (threadIdx.x == 32 && !feasible_pt(s, e.b)) fail = 1;
  __syncthreads();
  if (!fail) {
    const double kk2 = (double)(k + 2);
    for (long long i = threadIdx.x; i < k; i += blockDim.x) {
      if (*(volatile int*)&fail) break;
      const double u = (double)(i + 1) / kk2;
      double x[POMP_MAX_DIM];
      for (int j = 0; j < dim; j++) x[j] = e.a[j] + u * (e.b[j] - e.a[j]);
      if (!feasible_pt(s, x)) fail = 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *out = fail ? 0 : 1;
    __threadfence_system();
  }
}

// ============================================================================ host: space handle
static int build_obstacle_grid(pomp_space* sp, const pomp_space_desc* d, cudaStream_t st) {
  SpaceDev& s = sp->dev;
  int nb = d->n_boxes, nc = d->n_circles, no = nb + nc;
  s.use_grid = 0;
  if (no <= 16) return POMP_OK;  // a handful of obstacles: the plain loop is faster
  std::vector<double> ox0(no), oy0(no), ox1(no), oy1(no);
  for (int o = 0; o < nb; o++) {
    const double* b = d->boxes_h + 4 * o;
    ox0[o] = b[0];
    oy0[o] = b[1];
    ox1[o] = b[2];
    oy1[o] = b[3];
  }
  for (int o = 0; o < nc; o++) {
    const double* c = d->circles_h + 3 * o;
    double sl = 1e-9 * (fabs(c[0]) + fabs(c[1]) + fabs(c[2])) + 1e-300;
    ox0[nb + o] = c[0] - c[2] - sl;
    ox1[nb + o] = c[0] + c[2] + sl;
    oy0[nb + o] = c[1] - c[2] - sl;
    oy1[nb + o] = c[1] + c[2] + sl;
  }
  double glo0 = INFINITY, glo1 = INFINITY, ghi0 = -INFINITY, ghi1 = -INFINITY;
  for (int o = 0; o < no; o++) {
    if (!(ox0[o] <= ox1[o]) || !(oy0[o] <= oy1[o])) continue;  // empty / NaN obstacle: never contains a non-NaN point
    glo0 = fmin(glo0, ox0[o]);
    ghi0 = fmax(ghi0, ox1[o]);
    glo1 = fmin(glo1, oy0[o]);
    ghi1 = fmax(ghi1, oy1[o]);
  }
  if (!(isfinite(glo0) && isfinite(ghi0) && isfinite(glo1) && isfinite(ghi1))) return POMP_OK;
  double e0 = ghi0 - glo0, e1 = ghi1 - glo1;
  if (!(e0 > 0.0) || !(e1 > 0.0)) return POMP_OK;
  const char* gf = getenv("POMP_OBS_GRID_FACTOR");  // tuning aid: cells per obstacle
  // 10 K boxes, 4 Mi edges (profiles/edge_probe.py): 4 -> 0.49 ms, 16 -> 0.41 ms, 32 -> 0.39 ms; 16 keeps the
  // bucket table at ~4 MB (L2-resident) when a few obstacles are large
  const double factor = gf ? atof(gf) : 16.0;
  int res = (int)ceil(sqrt(factor * no));
  res = std::max(4, std::min(res, 4096));
  std::vector<int> start, items;
  int gnx = 0, gny = 0;
  double inv0 = 0, inv1 = 0;
  for (int attempt = 0; attempt < 8; attempt++) {
    gnx = gny = res;
    inv0 = (double)gnx / e0;
    inv1 = (double)gny / e1;
    std::vector<int> cnt((size_t)gnx * gny + 1, 0);
    long long total = 0;
    for (int o = 0; o < no; o++) {
      if (!(ox0[o] <= ox1[o]) || !(oy0[o] <= oy1[o])) continue;
      int cx0 = obs_cell(ox0[o], glo0, inv0, gnx), cx1 = obs_cell(ox1[o], glo0, inv0, gnx);
      int cy0 = obs_cell(oy0[o], glo1, inv1, gny), cy1 = obs_cell(oy1[o], glo1, inv1, gny);
      for (int cy = cy0; cy <= cy1; cy++)
        for (int cx = cx0; cx <= cx1; cx++) cnt[(size_t)cy * gnx + cx]++;
      total += (long long)(cx1 - cx0 + 1) * (cy1 - cy0 + 1);
    }
    if (total > 64LL * no && res > 4) {  // obstacles much larger than cells: coarsen
      res = std::max(4, res / 2);
      continue;
    }
    start.assign((size_t)gnx * gny + 1, 0);
    for (size_t c = 0; c < (size_t)gnx * gny; c++) start[c + 1] = start[c] + cnt[c];
    items.assign((size_t)std::max<long long>(total, 1), 0);
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int o = 0; o < no; o++) {
      if (!(ox0[o] <= ox1[o]) || !(oy0[o] <= oy1[o])) continue;
      int cx0 = obs_cell(ox0[o], glo0, inv0, gnx), cx1 = obs_cell(ox1[o], glo0, inv0, gnx);
      int cy0 = obs_cell(oy0[o], glo1, inv1, gny), cy1 = obs_cell(oy1[o], glo1, inv1, gny);
      for (int cy = cy0; cy <= cy1; cy++)
        for (int cx = cx0; cx <= cx1; cx++) items[(size_t)fill[(size_t)cy * gnx + cx]++] = o;
    }
    break;
  }
  if (start.empty()) return POMP_OK;
  CUDA_TRY(cudaMalloc((void**)&sp->d_cell_start, sizeof(int) * start.size()));
  CUDA_TRY(cudaMalloc((void**)&sp->d_cell_items, sizeof(int) * items.size()));
  CUDA_TRY(cudaMemcpyAsync(sp->d_cell_start, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(sp->d_cell_items, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice, st));
  std::vector<double> geom(items.size() * 4, 0.0);
  for (size_t i = 0; i < items.size(); i++) {
    const int o = items[i];
    if (o < nb) {
      for (int j = 0; j < 4; j++) geom[4 * i + j] = d->boxes_h[4 * (size_t)o + j];
    } else {
      for (int j = 0; j < 3; j++) geom[4 * i + j] = d->circles_h[3 * (size_t)(o - nb) + j];
      geom[4 * i + 3] = NAN;
    }
  }
  CUDA_TRY(cudaMalloc((void**)&sp->d_cell_geom, sizeof(double) * geom.size()));
  CUDA_TRY(cudaMemcpyAsync(sp->d_cell_geom, geom.data(), sizeof(double) * geom.size(), cudaMemcpyHostToDevice, st));
  // fine occupancy map (space.cuh).  PART: every cell between the cells of an obstacle's corners (obs_cell is monotone,
  // so no point of the obstacle maps outside that range).  FULL: the cells STRICTLY between the corner cells of a box
  // on both axes -- every x of such a cell satisfies x0 < x < x1 (x <= x0 would put it into a cell <= cell(x0)),
  // likewise y, so the box test returns true for every point of the cell.
  s.fine = nullptr;
  s.fnx = s.fny = 0;
  s.finv0 = s.finv1 = 0.0;
  {
    const char* ff = getenv("POMP_OBS_FINE_FACTOR");  // fine cells per obstacle (0: no map)
    const double ffac = ff ? atof(ff) : 256.0;
    int fres = (int)ceil(sqrt(ffac * no));
    fres = std::min(fres, 8192);
    if (ffac > 0.0 && fres >= 16) {
      fres = (fres + 15) / 16 * 16;
      const double fi0 = (double)fres / e0, fi1 = (double)fres / e1;
      std::vector<unsigned> fine(((size_t)fres * fres + 15) / 16, 0u);
      auto put = [&](int cx, int cy, unsigned code) {
        const size_t c = (size_t)cy * fres + cx;
        unsigned& w = fine[c >> 4];
        const int sh = (int)(c & 15) * 2;
        const unsigned cur = (w >> sh) & 3u;
        if (cur == (unsigned)OBS_FULL) return;  // FULL wins
        w = (w & ~(3u << sh)) | (code << sh);
      };
      for (int o = 0; o < no; o++) {
        if (!(ox0[o] <= ox1[o]) || !(oy0[o] <= oy1[o])) continue;
        const int cx0 = obs_cell(ox0[o], glo0, fi0, fres), cx1 = obs_cell(ox1[o], glo0, fi0, fres);
        const int cy0 = obs_cell(oy0[o], glo1, fi1, fres), cy1 = obs_cell(oy1[o], glo1, fi1, fres);
        for (int cy = cy0; cy <= cy1; cy++)
          for (int cx = cx0; cx <= cx1; cx++) {
            const bool inner = o < nb && cx > cx0 && cx < cx1 && cy > cy0 && cy < cy1;
            put(cx, cy, inner ? (unsigned)OBS_FULL : (unsigned)OBS_PART);
          }
      }
      CUDA_TRY(cudaMalloc((void**)&sp->d_fine, sizeof(unsigned) * fine.size()));
      CUDA_TRY(cudaMemcpyAsync(sp->d_fine, fine.data(), sizeof(unsigned) * fine.size(), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      s.fine = sp->d_fine;
      s.fnx = s.fny = fres;
      s.finv0 = fi0;
      s.finv1 = fi1;
    }
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  s.use_grid = 1;
  s.gnx = gnx;
  s.gny = gny;
  s.glo0 = glo0;
  s.glo1 = glo1;
  s.ghi0 = ghi0;
  s.ghi1 = ghi1;
  s.ginv0 = inv0;
  s.ginv1 = inv1;
  s.cell_start = sp->d_cell_start;
  s.cell_items = sp->d_cell_items;
  s.cell_geom = sp->d_cell_geom;
  return POMP_OK;
}

POMP_API int pomp_space_create(const pomp_space_desc* d, int device, pomp_space** out) {
  POMP_REQUIRE(out != nullptr && d != nullptr, "NULL argument to pomp_space_create");
  *out = nullptr;
  POMP_REQUIRE(d->dim >= 1 && d->dim <= POMP_MAX_DIM, "space dim %d outside 1..%d", d->dim, POMP_MAX_DIM);
  POMP_REQUIRE(d->n_boxes >= 0 && d->n_circles >= 0, "negative obstacle count");
  POMP_REQUIRE(d->n_boxes == 0 || d->boxes_h, "boxes_h is NULL");
  POMP_REQUIRE(d->n_circles == 0 || d->circles_h, "circles_h is NULL");
  if (d->n_boxes + d->n_circles > 0)
    POMP_REQUIRE(d->obstacle_dim0 >= 0 && d->obstacle_dim0 < d->dim && d->obstacle_dim1 >= 0 &&
                     d->obstacle_dim1 < d->dim,
                 "obstacle dims (%d,%d) outside the space", d->obstacle_dim0, d->obstacle_dim1);
  DeviceGuard g(device);
  if (!g.ok) return POMP_ERR_CUDA;
  pomp_space* sp = new pomp_space();
  sp->device = device;
  SpaceDev& s = sp->dev;
  memset(&s, 0, sizeof(s));
  s.dim = d->dim;
  s.od0 = d->n_boxes + d->n_circles > 0 ? d->obstacle_dim0 : 0;
  s.od1 = d->n_boxes + d->n_circles > 0 ? d->obstacle_dim1 : 0;
  s.n_boxes = d->n_boxes;
  s.n_circles = d->n_circles;
  for (int i = 0; i < d->dim; i++) {
    s.lo[i] = d->lo[i];
    s.hi[i] = d->hi[i];
  }
  auto fail = [&](int rc) {
    pomp_space_destroy(sp);
    return rc;
  };
  cudaError_t e = cudaStreamCreateWithFlags(&sp->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
    return fail(POMP_ERR_CUDA);
  }
  if (d->n_boxes) {
    if (cudaMalloc((void**)&sp->d_boxes, sizeof(double) * 4 * d->n_boxes) != cudaSuccess ||
        cudaMemcpy(sp->d_boxes, d->boxes_h, sizeof(double) * 4 * d->n_boxes, cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("uploading boxes failed: %s", cudaGetErrorString(cudaGetLastError()));
      return fail(POMP_ERR_CUDA);
    }
  }
  if (d->n_circles) {
    if (cudaMalloc((void**)&sp->d_circles, sizeof(double) * 3 * d->n_circles) != cudaSuccess ||
        cudaMemcpy(sp->d_circles, d->circles_h, sizeof(double) * 3 * d->n_circles, cudaMemcpyHostToDevice) !=
            cudaSuccess) {
      set_error("uploading circles failed: %s", cudaGetErrorString(cudaGetLastError()));
      return fail(POMP_ERR_CUDA);
    }
  }
  s.boxes = sp->d_boxes;
  s.circles = sp->d_circles;
  // mapped pinned words: [0] the overflow flag, [4] the answer of the single-edge fast path
  if (cudaHostAlloc((void**)&sp->overflow_h, 16 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&s.overflow, sp->overflow_h, 0) != cudaSuccess) {
    set_error("mapped flag allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(POMP_ERR_CUDA);
  }
  for (int j = 0; j < 16; j++) sp->overflow_h[j] = 0;
  // mapped scratch of the single-path / single-control-edge calls of planner loops (inputs read and the answer
  // written by the kernel over PCIe: no copies around a 1-thread kernel)
  if (cudaHostAlloc((void**)&sp->map_h, POMP_MAP_BYTES, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&sp->map_d, sp->map_h, 0) != cudaSuccess) {
    set_error("mapped scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(POMP_ERR_CUDA);
  }
  int rc = build_obstacle_grid(sp, d, sp->stream);
  if (rc != POMP_OK) return fail(rc);
  // obstacle tables were copied on the default stream; queries run on non-blocking streams
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("pomp_space_create: device synchronisation failed");
    pomp_space_destroy(sp);
    return POMP_ERR_CUDA;
  }
  *out = sp;
  return POMP_OK;
}

POMP_API int pomp_space_destroy(pomp_space* sp) {
  if (!sp) return POMP_OK;
  DeviceGuard g(sp->device);
  if (sp->d_boxes) cudaFree(sp->d_boxes);
  if (sp->d_circles) cudaFree(sp->d_circles);
  if (sp->d_cell_start) cudaFree(sp->d_cell_start);
  if (sp->d_cell_items) cudaFree(sp->d_cell_items);
  if (sp->d_cell_geom) cudaFree(sp->d_cell_geom);
  if (sp->d_fine) cudaFree(sp->d_fine);
  if (sp->stream) cudaStreamDestroy(sp->stream);
  if (sp->overflow_h) cudaFreeHost(sp->overflow_h);
  if (sp->map_h) cudaFreeHost(sp->map_h);
  delete sp;
  return POMP_OK;
}

POMP_API int pomp_space_get_stat(pomp_space* sp, const char* name, double* out) {
  POMP_REQUIRE(sp && name && out, "NULL argument to pomp_space_get_stat");
  if (!strcmp(name, "edge_overflow")) {  // read-and-clear; valid once the stream of the checked call has drained
    *out = sp->overflow_h ? (double)*sp->overflow_h : 0.0;
    if (sp->overflow_h) *sp->overflow_h = 0;
    return POMP_OK;
  }
  if (!strcmp(name, "n_boxes")) { *out = sp->dev.n_boxes; return POMP_OK; }
  if (!strcmp(name, "n_circles")) { *out = sp->dev.n_circles; return POMP_OK; }
  if (!strcmp(name, "obstacle_grid_cells")) { *out = sp->dev.use_grid ? (double)sp->dev.gnx * sp->dev.gny : 0.0; return POMP_OK; }
  set_error("unknown stat '%s'", name);
  return POMP_ERR_INVALID;
}

POMP_API int pomp_space_set_bounds(pomp_space* sp, const double* lo_h, const double* hi_h) {
  POMP_REQUIRE(sp && lo_h && hi_h, "NULL argument to pomp_space_set_bounds");
  for (int i = 0; i < sp->dev.dim; i++) {
    sp->dev.lo[i] = lo_h[i];
    sp->dev.hi[i] = hi_h[i];
  }
  return POMP_OK;
}

// ---- device entry points -------------------------------------------------------------------
POMP_API int pomp_feasible(pomp_space* sp, const double* x, int64_t n, uint8_t* ok, void* stream) {
  POMP_REQUIRE(sp && (n == 0 || (x && ok)) && n >= 0, "bad arguments to pomp_feasible");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n == 0) return POMP_OK;
  LAUNCH(feasible_kernel, (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream, sp->dev, x, (long long)n, ok);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_edge_check(pomp_space* sp, const double* a, const double* b, int64_t n, double resolution,
                             uint8_t* ok, void* stream) {
  POMP_REQUIRE(sp && (n == 0 || (a && b && ok)) && n >= 0, "bad arguments to pomp_edge_check");
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n == 0) return POMP_OK;
  int blocks = (int)((n + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (sp->dev.dim == 2) LAUNCH(edge_check_warp2d_kernel, blocks, 128, 0, st, sp->dev, a, b, (long long)n, resolution, ok);
  else LAUNCH((edge_check_kernel<0>), blocks, 128, 0, st, sp->dev, a, b, (long long)n, resolution, ok);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_edge_check_indexed(pomp_space* sp, const double* nodes, const int32_t* ij, int64_t n_edges,
                                     double resolution, uint8_t* ok, void* stream) {
  POMP_REQUIRE(sp && (n_edges == 0 || (nodes && ij && ok)) && n_edges >= 0, "bad arguments to pomp_edge_check_indexed");
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n_edges == 0) return POMP_OK;
  int blocks = (int)((n_edges + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (sp->dev.dim == 2)
    LAUNCH(edge_check_indexed_warp2d_kernel, blocks, 128, 0, st, sp->dev, nodes, ij, (long long)n_edges, resolution, ok,
           (const uint8_t*)nullptr);
  else
    LAUNCH((edge_check_indexed_kernel<0>), blocks, 128, 0, st, sp->dev, nodes, ij, (long long)n_edges, resolution, ok);
  CHECK_LAUNCH();
  return POMP_OK;
}

// The same with the node count known: ConfigurationSpace.feasible is evaluated ONCE per node (node_ok[n_nodes], device,
// caller-provided scratch that also returns the flags) and the edges only gather the two flags and test their interior
// samples.  In a k-NN roadmap every node is an endpoint of ~2k edges.
POMP_API int pomp_edge_check_indexed_nodes(pomp_space* sp, const double* nodes, int64_t n_nodes, const int32_t* ij,
                                           int64_t n_edges, double resolution, uint8_t* node_ok, uint8_t* ok,
                                           void* stream) {
  POMP_REQUIRE(sp && n_nodes >= 0 && n_edges >= 0 && (n_nodes == 0 || (nodes && node_ok)) && (n_edges == 0 || (ij && ok)),
               "bad arguments to pomp_edge_check_indexed_nodes");
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n_nodes > 0) POMP_TRY(pomp_feasible(sp, nodes, n_nodes, node_ok, stream));
  if (n_edges == 0) return POMP_OK;
  if (sp->dev.dim != 2) return pomp_edge_check_indexed(sp, nodes, ij, n_edges, resolution, ok, stream);
  LAUNCH(edge_check_indexed_warp2d_kernel, (int)((n_edges + 127) / 128), 128, 0, (cudaStream_t)stream, sp->dev, nodes,
         ij, (long long)n_edges, resolution, ok, (const uint8_t*)node_ok);
  CHECK_LAUNCH();
  return POMP_OK;
}

static int validate_geodesic(const pomp_metric_desc* g, int dim) {
  POMP_REQUIRE(g != nullptr, "geodesic descriptor is NULL");
  POMP_REQUIRE(!g->has_cost, "geodesic descriptor must not carry a cost coordinate");
  POMP_REQUIRE(g->dim == dim, "geodesic dim %d != space dim %d", g->dim, dim);
  POMP_REQUIRE(g->kind == POMP_METRIC_EUCLIDEAN || g->kind == POMP_METRIC_GEODESIC_PRODUCT,
               "geodesic must be EUCLIDEAN or GEODESIC_PRODUCT");
  return POMP_OK;
}

POMP_API int pomp_edge_check_path(pomp_space* sp, const pomp_metric_desc* geodesic, const double* waypoints,
                                  const int32_t* path_offsets, int64_t n_paths, double resolution, uint8_t* ok,
                                  void* stream) {
  POMP_REQUIRE(sp && (n_paths == 0 || (waypoints && path_offsets && ok)) && n_paths >= 0,
               "bad arguments to pomp_edge_check_path");
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  POMP_TRY(validate_geodesic(geodesic, sp->dev.dim));
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n_paths == 0) return POMP_OK;
  LAUNCH(edge_check_path_kernel, (int)((n_paths + 63) / 64), 64, 0, (cudaStream_t)stream, sp->dev, *geodesic,
         waypoints, path_offsets, (long long)n_paths, resolution, ok);
  CHECK_LAUNCH();
  return POMP_OK;
}

int pomp::validate_dyn(const pomp_dyn_desc* d) {
  POMP_REQUIRE(d != nullptr, "dynamics descriptor is NULL");
  POMP_REQUIRE(d->kind >= 0 && d->kind <= POMP_DYN_FLAPPY, "unknown dynamics kind %d", d->kind);
  POMP_REQUIRE(d->x_dim >= 1 && d->x_dim + (d->cost_kind != POMP_COST_NONE) <= POMP_MAX_DIM, "bad x_dim %d", d->x_dim);
  POMP_REQUIRE(d->u_dim >= 1 && d->u_dim <= POMP_MAX_DIM, "bad u_dim %d", d->u_dim);
  POMP_REQUIRE(d->cost_kind >= 0 && d->cost_kind <= POMP_COST_PATH_LENGTH, "unknown cost kind %d", d->cost_kind);
  if (d->kind == POMP_DYN_CV || d->kind == POMP_DYN_CV_EULER)
    POMP_REQUIRE(d->x_dim % 2 == 0 && d->u_dim == d->x_dim / 2 + 1, "CV dynamics need x=(q,v), u=(dt,a)");
  if (d->kind == POMP_DYN_CV_EULER || d->kind == POMP_DYN_PENDULUM)
    POMP_REQUIRE(d->dt > 0.0, "integration step dt must be positive");
  if (d->kind == POMP_DYN_DUBINS) POMP_REQUIRE(d->x_dim == 3 && d->u_dim == 2, "Dubins needs x_dim=3, u_dim=2");
  if (d->kind == POMP_DYN_PENDULUM) POMP_REQUIRE(d->x_dim == 2 && d->u_dim == 2, "Pendulum needs x_dim=2, u_dim=2");
  if (d->kind == POMP_DYN_FLAPPY) POMP_REQUIRE(d->x_dim == 3 && d->u_dim == 2, "Flappy needs x_dim=3, u_dim=2");
  if (d->kind == POMP_DYN_ADAPTOR) POMP_REQUIRE(d->u_dim == d->x_dim, "adaptor needs u_dim == x_dim");
  if (d->cost_kind == POMP_COST_PATH_LENGTH) POMP_REQUIRE(d->kind == POMP_DYN_ADAPTOR, "path-length cost needs an adaptor");
  return POMP_OK;
}

POMP_API int pomp_edge_check_control(pomp_space* sp, const pomp_dyn_desc* dyn, const pomp_metric_desc* geodesic,
                                     const double* x, const double* u, int64_t n, double resolution, uint8_t* ok,
                                     void* stream) {
  POMP_REQUIRE(sp && (n == 0 || (x && u && ok)) && n >= 0, "bad arguments to pomp_edge_check_control");
  POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
  POMP_TRY(validate_dyn(dyn));
  POMP_REQUIRE(dyn->x_dim == sp->dev.dim, "dynamics x_dim %d != space dim %d", dyn->x_dim, sp->dev.dim);
  POMP_REQUIRE(dyn->kind == POMP_DYN_ADAPTOR || dyn->kind == POMP_DYN_DUBINS || dyn->kind == POMP_DYN_PENDULUM ||
                   dyn->kind == POMP_DYN_CV_EULER,
               "edge_check_control supports ADAPTOR, DUBINS, PENDULUM, CV_EULER");
  pomp_metric_desc gd;
  memset(&gd, 0, sizeof(gd));
  gd.kind = POMP_METRIC_EUCLIDEAN;
  gd.dim = sp->dev.dim;
  if (dyn->kind == POMP_DYN_PENDULUM || dyn->kind == POMP_DYN_CV_EULER) {
    POMP_TRY(validate_geodesic(geodesic, sp->dev.dim));
    gd = *geodesic;
  }
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n == 0) return POMP_OK;
  LAUNCH(edge_check_control_kernel, (int)((n + 63) / 64), 64, 0, (cudaStream_t)stream, sp->dev, *dyn, gd, x, u,
         (long long)n, resolution, ok);
  CHECK_LAUNCH();
  return POMP_OK;
}

// ---- host-buffer variants ------------------------------------------------------------------
#define H2D(buf, src, count, T)                                                                       \
  do {                                                                                                \
    POMP_TRY(buf.reserve((size_t)(count)));                                                           \
    CUDA_TRY(cudaMemcpyAsync(buf.p, src, sizeof(T) * (size_t)(count), cudaMemcpyHostToDevice, st));   \
  } while (0)

static int finish_ok(pomp_space* sp, uint8_t* ok_h, int64_t n, cudaStream_t st) {
  CUDA_TRY(cudaMemcpyAsync(ok_h, sp->out_ok.p, (size_t)n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return space_overflow_check(sp);
}

POMP_API int pomp_feasible_h(pomp_space* sp, const double* x_h, int64_t n, uint8_t* ok_h) {
  POMP_REQUIRE(sp && (n == 0 || (x_h && ok_h)) && n >= 0, "bad arguments to pomp_feasible_h");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n == 0) return POMP_OK;
  cudaStream_t st = sp->stream;
  H2D(sp->in_a, x_h, n * sp->dev.dim, double);
  POMP_TRY(sp->out_ok.reserve((size_t)n));
  POMP_TRY(pomp_feasible(sp, sp->in_a.p, n, sp->out_ok.p, st));
  return finish_ok(sp, ok_h, n, st);
}

POMP_API int pomp_edge_check_h(pomp_space* sp, const double* a_h, const double* b_h, int64_t n, double resolution,
                               uint8_t* ok_h) {
  POMP_REQUIRE(sp && (n == 0 || (a_h && b_h && ok_h)) && n >= 0, "bad arguments to pomp_edge_check_h");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n == 0) return POMP_OK;
  cudaStream_t st = sp->stream;
  if (n == 1) {
    POMP_REQUIRE(resolution > 0.0, "resolution must be positive");
    EdgeArg e;
    for (int j = 0; j < POMP_MAX_DIM; j++) {
      e.a[j] = j < sp->dev.dim ? a_h[j] : 0.0;
      e.b[j] = j < sp->dev.dim ? b_h[j] : 0.0;
    }
    LAUNCH(edge_one_kernel, 1, 128, 0, st, sp->dev, e, resolution, sp->dev.overflow + 4);
    CHECK_LAUNCH();
    CUDA_TRY(cudaStreamSynchronize(st));
    ok_h[0] = sp->overflow_h[4] ? 1 : 0;
    return space_overflow_check(sp);
  }
  H2D(sp->in_a, a_h, n * sp->dev.dim, double);
  H2D(sp->in_b, b_h, n * sp->dev.dim, double);
  POMP_TRY(sp->out_ok.reserve((size_t)n));
  POMP_TRY(pomp_edge_check(sp, sp->in_a.p, sp->in_b.p, n, resolution, sp->out_ok.p, st));
  return finish_ok(sp, ok_h, n, st);
}

POMP_API int pomp_edge_check_path_h(pomp_space* sp, const pomp_metric_desc* geodesic, const double* waypoints_h,
                                    const int32_t* path_offsets_h, int64_t n_paths, double resolution,
                                    uint8_t* ok_h) {
  POMP_REQUIRE(sp && (n_paths == 0 || (waypoints_h && path_offsets_h && ok_h)) && n_paths >= 0,
               "bad arguments to pomp_edge_check_path_h");
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n_paths == 0) return POMP_OK;
  cudaStream_t st = sp->stream;
  int64_t nw = path_offsets_h[n_paths];
  POMP_REQUIRE(nw >= 0, "negative waypoint count");
  H2D(sp->in_a, waypoints_h, std::max<int64_t>(nw, 1) * sp->dev.dim, double);
  H2D(sp->in_i, path_offsets_h, n_paths + 1, int32_t);
  POMP_TRY(sp->out_ok.reserve((size_t)n_paths));
  POMP_TRY(pomp_edge_check_path(sp, geodesic, sp->in_a.p, sp->in_i.p, n_paths, resolution, sp->out_ok.p, st));
  return finish_ok(sp, ok_h, n_paths, st);
}

POMP_API int pomp_edge_check_control_h(pomp_space* sp, const pomp_dyn_desc* dyn, const pomp_metric_desc* geodesic,
                                       const double* x_h, const double* u_h, int64_t n, double resolution,
                                       uint8_t* ok_h) {
  POMP_REQUIRE(sp && dyn && (n == 0 || (x_h && u_h && ok_h)) && n >= 0, "bad arguments to pomp_edge_check_control_h");
  POMP_TRY(validate_dyn(dyn));
  DeviceGuard g(sp->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (n == 0) return POMP_OK;
  cudaStream_t st = sp->stream;
  if (n == 1 && sp->map_h) {
    // one control edge of a planner loop: state, control and the answer in mapped host memory
    unsigned char* mh = sp->map_h;
    unsigned char* md = sp->map_d;
    memcpy(mh + 256, x_h, sizeof(double) * dyn->x_dim);
    memcpy(mh + 512, u_h, sizeof(double) * dyn->u_dim);
    POMP_TRY(pomp_edge_check_control(sp, dyn, geodesic, (const double*)(md + 256), (const double*)(md + 512), 1,
                                     resolution, (uint8_t*)md, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ok_h[0] = mh[0];
    return space_overflow_check(sp);
  }
  H2D(sp->in_a, x_h, n * dyn->x_dim, double);
  H2D(sp->in_b, u_h, n * dyn->u_dim, double);
  POMP_TRY(sp->out_ok.reserve((size_t)n));
  POMP_TRY(pomp_edge_check_control(sp, dyn, geodesic, sp->in_a.p, sp->in_b.p, n, resolution, sp->out_ok.p, st));
  return finish_ok(sp, ok_h, n, st);
}
