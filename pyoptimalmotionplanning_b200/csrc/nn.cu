// nn.cu -- NearestNeighbors on the device (pomp/structures/nearestneighbors.py:13-141).
//
// Two exact search regimes over the same insertion-ordered fp64 AoS point store:
//   * brute force, warp per (query, point-chunk): any metric, any k; used while the node set or
//     the query batch is small (planner-in-the-loop calls, N <= ~1e5) -- FP64-ALU bound;
//   * uniform-grid bucket index (grid.cuh), thread per query: large node sets and query batches
//     (the 16 Mi point / 1 Mi query configuration) -- HBM/L2 bound.
// Both order candidates by (distance, insertion index), so they agree with each other and with
// the reference's brute-force scan bit for bit.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "grid.cuh"
#include "grid_fast.cuh"
#include "grid_se2.cuh"
#include "nn_handle.cuh"
#include "scan.cuh"
#include "topk.cuh"

using namespace pomp;

static int base_dim(const pomp_metric_desc& m) { return m.has_cost ? m.dim - 1 : m.dim; }

int pomp::validate_metric(const pomp_metric_desc* m) {
  POMP_REQUIRE(m != nullptr, "metric descriptor is NULL");
  POMP_REQUIRE(m->dim >= 1 && m->dim <= POMP_MAX_DIM, "metric dim %d outside 1..%d", m->dim, POMP_MAX_DIM);
  POMP_REQUIRE(m->kind >= 0 && m->kind <= POMP_METRIC_GEODESIC_PRODUCT, "unknown metric kind %d", m->kind);
  POMP_REQUIRE(!m->has_cost || m->dim >= 2, "cost metric needs at least one base coordinate");
  if (m->kind == POMP_METRIC_GEODESIC_PRODUCT) {
    POMP_REQUIRE(m->n_comp >= 1 && m->n_comp <= POMP_MAX_COMP, "bad component count %d", m->n_comp);
    int s = 0;
    for (int c = 0; c < m->n_comp; c++) {
      POMP_REQUIRE(m->comp_dim[c] >= 1, "component %d has dimension %d", c, m->comp_dim[c]);
      POMP_REQUIRE(m->comp_kind[c] != POMP_COMP_SO2 || m->comp_dim[c] == 1, "SO2 component must have dim 1");
      s += m->comp_dim[c];
    }
    POMP_REQUIRE(s == base_dim(*m), "component dims sum to %d, expected %d", s, base_dim(*m));
  }
  return POMP_OK;
}

int pomp::nn_grow(pomp_nn* h, int64_t need, cudaStream_t st) {
  if (need <= h->cap) return POMP_OK;
  int64_t ncap = std::max<int64_t>(need, std::max<int64_t>(1024, h->cap * 2));
  double* np = nullptr;
  uint8_t* ns = nullptr;
  cudaError_t e = cudaMalloc((void**)&np, sizeof(double) * (size_t)ncap * h->D);
  if (e == cudaSuccess) e = cudaMalloc((void**)&ns, (size_t)ncap);
  if (e != cudaSuccess) {
    if (np) cudaFree(np);
    set_error("cudaMalloc for %lld points failed: %s", (long long)ncap, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? POMP_ERR_NOMEM : POMP_ERR_CUDA;
  }
  CUDA_TRY(cudaMemsetAsync(ns, 0, (size_t)ncap, st));
  if (h->n_dev > 0) {
    CUDA_TRY(cudaMemcpyAsync(np, h->d_pts, sizeof(double) * (size_t)h->n_dev * h->D, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ns, h->d_skip, (size_t)h->n_dev, cudaMemcpyDeviceToDevice, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  if (h->d_pts) cudaFree(h->d_pts);
  if (h->d_skip) cudaFree(h->d_skip);
  h->d_pts = np;
  h->d_skip = ns;
  h->cap = ncap;
  return POMP_OK;
}

// upload host-side pending adds
int pomp::nn_flush(pomp_nn* h, cudaStream_t st) {
  if (h->pending.empty()) return POMP_OK;
  int64_t m = (int64_t)h->pending.size() / h->D;
  POMP_TRY(nn_grow(h, h->n_dev + m, st));
  CUDA_TRY(cudaMemcpyAsync(h->d_pts + h->n_dev * h->D, h->pending.data(), sizeof(double) * h->pending.size(),
                           cudaMemcpyHostToDevice, st));
  h->n_dev += m;
  h->pending.clear();
  return POMP_OK;
}

// ============================================================================ brute force kernels
// One warp per (query, split).  Lanes stride over the split's points.

__device__ __forceinline__ void load_q(const double* q, int D, double* out) {
  for (int j = 0; j < D; j++) out[j] = q[j];
}

struct SplitGeom {
  int nsplit;
  long long chunk;
};

__global__ void bf_nearest_kernel(const double* __restrict__ pts, long long n, pomp_metric_desc m,
                                  const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq,
                                  SplitGeom sg, double* __restrict__ part_d, int* __restrict__ part_i) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= nq * sg.nsplit) return;
  long long qi = warp / sg.nsplit;
  int sp = (int)(warp % sg.nsplit);
  const int D = m.dim;
  double qq[POMP_MAX_DIM], p[POMP_MAX_DIM];
  load_q(q + qi * D, D, qq);
  long long b = (long long)sp * sg.chunk, e = min(n, b + sg.chunk);
  double bd = POMP_INF;
  int bi = IDX_NONE;
  const bool euclid = (m.kind == POMP_METRIC_EUCLIDEAN && !m.has_cost);
  for (long long i = b + lane; i < e; i += 32) {
    if (skip && skip[i]) continue;
    for (int j = 0; j < D; j++) p[j] = __ldg(pts + i * D + j);
    double d;
    if (euclid) {
      double s = 0.0;
      for (int j = 0; j < D; j++) {
        double t = p[j] - qq[j];
        s = s + t * t;
      }
      d = sqrt(s);
    } else {
      d = metric_eval(m, p, qq);
    }
    if (d < bd) {  // strict '<' in ascending index order keeps the first minimum (nearestneighbors.py:93)
      bd = d;
      bi = (int)i;
    }
  }
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
    part_d[warp] = bd;
    part_i[warp] = bi;
  }
}

// Plain Euclidean nearest neighbour for query BATCHES against a small node set (the planner regime,
// N below the grid threshold).  N-body style tiling: a block stages tiles of points in shared memory
// and every thread keeps QT queries in registers, so one broadcast shared-memory read feeds 32*QT
// pairs (the warp kernel above streams every point from L2 once per query and takes a square root
// per pair).  Per pair: D subtractions, D multiplications, D-1 additions and one compare on the
// SQUARED distance against fl(T^2)(1+2^-50); only candidates that pass take the square root and are
// compared on the rooted value like the reference (strict '<' in ascending index order).
template <int D, int QT>
__global__ void __launch_bounds__(256)
bf_nearest_tiled_kernel(const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                        const double* __restrict__ q, long long nq, SplitGeom sg, double* __restrict__ part_d,
                        int* __restrict__ part_i) {
  constexpr int TILE = 512;
  __shared__ double sp[TILE * D];
  __shared__ uint8_t ssk[TILE];
  const long long qbase = (long long)blockIdx.x * (256 * QT);
  const int split = blockIdx.y;
  double qq[QT][D], bd[QT], tsq[QT];
  int bi[QT];
#pragma unroll
  for (int j = 0; j < QT; j++) {
    long long qi = qbase + threadIdx.x + (long long)j * 256;
    if (qi >= nq) qi = nq - 1;
#pragma unroll
    for (int a = 0; a < D; a++) qq[j][a] = q[qi * D + a];
    bd[j] = POMP_INF;
    tsq[j] = POMP_INF;
    bi[j] = IDX_NONE;
  }
  const long long b = (long long)split * sg.chunk, e = min(n, b + sg.chunk);
  for (long long t0 = b; t0 < e; t0 += TILE) {
    const int cnt = (int)min((long long)TILE, e - t0);
    for (int x = threadIdx.x; x < cnt * D; x += 256) sp[x] = __ldg(pts + t0 * D + x);
    if (skip)
      for (int x = threadIdx.x; x < cnt; x += 256) ssk[x] = skip[t0 + x];
    __syncthreads();
    for (int p = 0; p < cnt; p++) {
      if (skip && ssk[p]) continue;
      double pc[D];
#pragma unroll
      for (int a = 0; a < D; a++) pc[a] = sp[p * D + a];
#pragma unroll
      for (int j = 0; j < QT; j++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; a++) {
          double t = pc[a] - qq[j][a];
          s = s + t * t;
        }
        if (s <= tsq[j]) {
          double d = sqrt(s);
          if (d < bd[j]) {
            bd[j] = d;
            bi[j] = (int)(t0 + p);
            tsq[j] = (d * d) * (1.0 + 0x1p-50);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < QT; j++) {
    long long qi = qbase + threadIdx.x + (long long)j * 256;
    if (qi < nq) {
      part_d[qi * sg.nsplit + split] = bd[j];
      part_i[qi * sg.nsplit + split] = bi[j];
    }
  }
}

__global__ void bf_nearest_reduce_kernel(const double* __restrict__ part_d, const int* __restrict__ part_i,
                                         long long nq, int nsplit, int64_t* __restrict__ idx,
                                         double* __restrict__ dist) {
  long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  double bd = POMP_INF;
  int bi = IDX_NONE;
  for (int s = 0; s < nsplit; s++) {
    double d = part_d[qi * nsplit + s];
    int i = part_i[qi * nsplit + s];
    if (key_less(d, i, bd, bi)) {
      bd = d;
      bi = i;
    }
  }
  bool found = bd < POMP_INF;
  idx[qi] = found ? (int64_t)bi : -1;
  dist[qi] = found ? bd : POMP_INF;
}

// Single query passed by value, result written straight into a mapped host mailbox by the last
// block to finish: one launch + one stream sync per planner-side nearest() call.
__global__ void bf_single_nearest_kernel(const double* __restrict__ pts, long long n, pomp_metric_desc m,
                                         const uint8_t* __restrict__ skip, QueryArg qa, double* part_d,
                                         int* part_i, unsigned int* ticket, Mailbox* out) {
  const int D = m.dim;
  double p[POMP_MAX_DIM];
  double bd = POMP_INF;
  int bi = IDX_NONE;
  const bool euclid = (m.kind == POMP_METRIC_EUCLIDEAN && !m.has_cost);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (skip && skip[i]) continue;
    for (int j = 0; j < D; j++) p[j] = __ldg(pts + i * D + j);
    double d;
    if (euclid) {
      double s = 0.0;
      for (int j = 0; j < D; j++) {
        double t = p[j] - qa.v[j];
        s = s + t * t;
      }
      d = sqrt(s);
    } else {
      d = metric_eval(m, p, qa.v);
    }
    if (key_less(d, (int)i, bd, bi)) {
      bd = d;
      bi = (int)i;
    }
  }
  __shared__ double sd[32];
  __shared__ int si[32];
  __shared__ bool last;
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
  if (w == 0) {
    bd = lane < nw ? sd[lane] : POMP_INF;
    bi = lane < nw ? si[lane] : IDX_NONE;
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
      part_d[blockIdx.x] = bd;
      part_i[blockIdx.x] = bi;
      __threadfence();
      unsigned int t = atomicAdd(ticket, 1u);
      last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  bd = POMP_INF;
  bi = IDX_NONE;
  for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
    double d = ((volatile double*)part_d)[b];
    int i = ((volatile int*)part_i)[b];
    if (key_less(d, i, bd, bi)) {
      bd = d;
      bi = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double od = __shfl_xor_sync(FULL_MASK, bd, o);
    int oi = __shfl_xor_sync(FULL_MASK, bi, o);
    if (key_less(od, oi, bd, bi)) {
      bd = od;
      bi = oi;
    }
  }
  __syncthreads();
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
    out->idx = found ? (int64_t)bi : -1;
    out->dist = found ? bd : POMP_INF;
    *ticket = 0;  // re-arm for the next call
    __threadfence_system();
  }
}

template <int KL>
__global__ void bf_knn_kernel(const double* __restrict__ pts, long long n, pomp_metric_desc m,
                              const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq,
                              int k, SplitGeom sg, double* __restrict__ part_d, int* __restrict__ part_i) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= nq * sg.nsplit) return;
  long long qi = warp / sg.nsplit;
  int sp = (int)(warp % sg.nsplit);
  const int D = m.dim;
  double qq[POMP_MAX_DIM], p[POMP_MAX_DIM];
  load_q(q + qi * D, D, qq);
  long long b = (long long)sp * sg.chunk, e = min(n, b + sg.chunk);
  WarpTopK<KL, int> top;
  top.init(k);
  for (long long i0 = b; i0 < e; i0 += 32) {
    long long i = i0 + lane;
    double d = POMP_INF;
    if (i < e && !(skip && skip[i])) {
      for (int j = 0; j < D; j++) p[j] = __ldg(pts + i * D + j);
      d = metric_eval(m, p, qq);
    }
    top.offer(d, (int)i);
  }
  top.store(part_d + warp * k, part_i + warp * k);
}

// merge `nlists` sorted-or-not candidate lists of length k per query into the k best
template <int KL, typename InI>
__global__ void merge_lists_kernel(const double* __restrict__ in_d, const InI* __restrict__ in_i, int nlists,
                                   long long nq, int k, long long list_stride_q, long long list_stride_l,
                                   int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                   int32_t* __restrict__ out_cnt) {
  long long qi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (qi >= nq) return;
  WarpTopK<KL, int64_t> top;
  top.init(k);
  for (int l = 0; l < nlists; l++) {
    const double* ld = in_d + qi * list_stride_q + (long long)l * list_stride_l;
    const InI* li = in_i + qi * list_stride_q + (long long)l * list_stride_l;
    for (int j0 = 0; j0 < k; j0 += 32) {
      int j = j0 + lane;
      double d = POMP_INF;
      int64_t id = 0;
      if (j < k) {
        id = (int64_t)li[j];
        d = id >= 0 ? ld[j] : POMP_INF;
      }
      top.offer(d, id);
    }
  }
  int cnt = top.store(out_dist + qi * k, out_idx + qi * k);
  if (lane == 0 && out_cnt) out_cnt[qi] = cnt;
}

// radius: per (query, split) hit counts, then fill in index order
template <bool FILL>
__global__ void bf_radius_kernel(const double* __restrict__ pts, long long n, pomp_metric_desc m,
                                 const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq,
                                 double radius, int inclusive, SplitGeom sg, int* __restrict__ counts,
                                 const int64_t* __restrict__ offsets, int64_t* __restrict__ out) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= nq * sg.nsplit) return;
  long long qi = warp / sg.nsplit;
  int sp = (int)(warp % sg.nsplit);
  const int D = m.dim;
  double qq[POMP_MAX_DIM], p[POMP_MAX_DIM];
  load_q(q + qi * D, D, qq);
  long long b = (long long)sp * sg.chunk, e = min(n, b + sg.chunk);
  long long cnt = 0;
  int64_t base = FILL ? offsets[warp] : 0;
  for (long long i0 = b; i0 < e; i0 += 32) {
    long long i = i0 + lane;
    bool hit = false;
    if (i < e && !(skip && skip[i])) {
      for (int j = 0; j < D; j++) p[j] = __ldg(pts + i * D + j);
      double d = metric_eval(m, p, qq);
      hit = inclusive ? (d <= radius) : (d < radius);
    }
    unsigned mask = __ballot_sync(FULL_MASK, hit);
    if (FILL && hit) out[base + cnt + __popc(mask & ((1u << lane) - 1))] = i;
    cnt += __popc(mask);
  }
  if (!FILL && lane == 0) counts[warp] = (int)cnt;
}

// Single radius query of the planner loops (SST's pickNode / nodeLocallyBest, EST): the query travels in the kernel
// arguments, every block scans its share of the nodes, hits are appended (warp-aggregated) to a list in MAPPED HOST
// memory and the last block publishes the count: one launch + one stream synchronisation per neighbors() call, the
// host sorts the handful of hits into insertion order.  (The general path: count kernel, 3-kernel scan, two copies, a
// synchronisation for the total, fill kernel, two more copies.)
constexpr int RADBOX_CAP = 8192;
struct RadBox {
  unsigned int count;
  int hits[RADBOX_CAP];
};
__global__ void bf_single_radius_kernel(const double* __restrict__ pts, long long n, pomp_metric_desc m,
                                        const uint8_t* __restrict__ skip, QueryArg qa, double radius, int inclusive,
                                        unsigned int* counters /* [0] hits, [1] ticket */, RadBox* out) {
  const int D = m.dim;
  const int lane = threadIdx.x & 31;
  double p[POMP_MAX_DIM];
  // whole warps stay in the loop together (ballot): the bound is rounded up to a multiple of 32 per warp
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
    const long long i = i0 + lane;
    bool hit = false;
    if (i < n && !(skip && skip[i])) {
      for (int j = 0; j < D; j++) p[j] = __ldg(pts + i * D + j);
      const double d = metric_eval(m, p, qa.v);
      hit = inclusive ? (d <= radius) : (d < radius);
    }
    const unsigned mask = __ballot_sync(FULL_MASK, hit);
    if (mask) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(counters, (unsigned)__popc(mask));
      base = __shfl_sync(FULL_MASK, base, 0);
      const unsigned int at = base + (unsigned)__popc(mask & ((1u << lane) - 1u));
      if (hit && at < (unsigned)RADBOX_CAP) out->hits[at] = (int)i;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int t = atomicAdd(counters + 1, 1u);
    if (t == gridDim.x - 1) {
      out->count = *(volatile unsigned int*)counters;
      counters[0] = 0;  // re-arm for the next call
      counters[1] = 0;
      __threadfence_system();
    }
  }
}

// ============================================================================ grid build kernels
// DC > 0: compile-time dimension (bounds in registers, compare + select: sm_100 has no FP64 min / max instruction and
// fmin / fmax expand to a NaN-aware sequence; `v < mn ? v : mn` ignores NaN coordinates the same way); DC = 0: any D
template <int DC>
__global__ void __launch_bounds__(256)
bbox_kernel(const double* __restrict__ pts, long long n, int D, double* __restrict__ part) {
  constexpr int DA = DC ? DC : POMP_MAX_DIM;
  const int dim = DC ? DC : D;
  double mn[DA], mx[DA];
#pragma unroll
  for (int j = 0; j < DA; j++) {
    mn[j] = POMP_INF;
    mx[j] = -POMP_INF;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double p[DA];
    if (DC) load_point<DA>(pts, i, p);
    else
      for (int j = 0; j < dim; j++) p[j] = pts[i * dim + j];
#pragma unroll
    for (int j = 0; j < DA; j++)
      if (j < dim) {
        mn[j] = p[j] < mn[j] ? p[j] : mn[j];
        mx[j] = p[j] > mx[j] ? p[j] : mx[j];
      }
  }
  __shared__ double smn[32][POMP_MAX_DIM], smx[32][POMP_MAX_DIM];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < DA; j++) {
    if (j >= dim) break;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double a = __shfl_xor_sync(FULL_MASK, mn[j], o), b2 = __shfl_xor_sync(FULL_MASK, mx[j], o);
      mn[j] = a < mn[j] ? a : mn[j];
      mx[j] = b2 > mx[j] ? b2 : mx[j];
    }
    if (lane == 0) {
      smn[w][j] = mn[j];
      smx[w][j] = mx[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < dim) {
    int j = threadIdx.x;
    double a = POMP_INF, b = -POMP_INF;
    for (int x = 0; x < nw; x++) {
      a = smn[x][j] < a ? smn[x][j] : a;
      b = smx[x][j] > b ? smx[x][j] : b;
    }
    part[((long long)blockIdx.x * 2 + 0) * POMP_MAX_DIM + j] = a;
    part[((long long)blockIdx.x * 2 + 1) * POMP_MAX_DIM + j] = b;
  }
}

__device__ __forceinline__ long long grid_cell_id(const GridDev& g, const double* p) {
  long long id = 0;
  for (int a = 0; a < g.G; a++) id += (long long)grid_cell_p(g, a, p[g.gdim[a]]) * g.stride[a];
  return id;
}

// Index build = counting sort of the points by cell id.  Scattering the POINTS (16-48 B each) to random
// positions costs a read-modify-write of a DRAM sector per point (1.59 ms for 2^24 2-D points); instead
// only the 4-byte insertion index is scattered, and the points are then GATHERED in sorted order
// (random reads, coalesced writes).
__global__ void cell_count_kernel(GridDev g, const double* __restrict__ pts, long long n, int D,
                                  int* __restrict__ counts, int* __restrict__ slot, int* __restrict__ cellid) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p[POMP_MAX_DIM];
  for (int j = 0; j < D; j++) p[j] = pts[i * D + j];
  long long c = grid_cell_id(g, p);
  cellid[i] = (int)c;
  slot[i] = atomicAdd(counts + c, 1);
}
__global__ void cell_scatter_kernel(long long n, const int* __restrict__ cell_start, const int* __restrict__ slot,
                                    const int* __restrict__ cellid, int* __restrict__ sidx) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sidx[(long long)cell_start[cellid[i]] + slot[i]] = (int)i;
}
template <int D>
__global__ void cell_gather_kernel(const double* __restrict__ pts, long long n, int Drt,
                                   const int* __restrict__ sidx, double* __restrict__ spts) {
  long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const long long i = sidx[pos];
  if (D == 2) {
    reinterpret_cast<double2*>(spts)[pos] = __ldg(reinterpret_cast<const double2*>(pts) + i);
  } else {
    const int dim = D ? D : Drt;
    for (int j = 0; j < dim; j++) spts[pos * dim + j] = __ldg(pts + i * dim + j);
  }
}
__global__ void query_cell_kernel(GridDev g, const double* __restrict__ q, long long nq, int D, int shift, int nbx,
                                  int row_div,
                                  int* __restrict__ counts, int* __restrict__ qcell, int* __restrict__ slot,
                                  int* __restrict__ zero_me) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && zero_me) *zero_me = 0;  // overflow counter of the search that follows
  if (i >= nq) return;
  double p[POMP_MAX_DIM];
  for (int j = 0; j < D; j++) p[j] = q[i * D + j];
  int c0 = grid_cell(g, 0, p[g.gdim[0]]);
  long long row = 0;
  for (int a = 1; a < g.G; a++) row += (long long)grid_cell_p(g, a, p[g.gdim[a]]) * (g.stride[a] / g.ncell[0]);
  long long b = (row / row_div) * nbx + (c0 >> shift);  // row_div > 1: bins are 2-D tiles (2-axis grids only)
  qcell[i] = (int)b;
  slot[i] = atomicAdd(counts + b, 1);
}

__global__ void query_perm_kernel(const int* __restrict__ qcell, const int* __restrict__ slot,
                                  const int* __restrict__ start, const double* __restrict__ q, long long nq, int D,
                                  int* __restrict__ perm, double* __restrict__ qsorted) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int t = start[qcell[i]] + slot[i];
  perm[t] = (int)i;
  if (qsorted)
    for (int j = 0; j < D; j++) qsorted[(long long)t * D + j] = q[i * D + j];
}

// ============================================================================ grid query kernels
template <int K>
struct KnnColl {
  ThreadTopK<K>& top;
  const int* __restrict__ sidx;
  __device__ __forceinline__ double threshold() const { return top.threshold(); }
  __device__ __forceinline__ void offer(double d, long long pos) {
    if (top.maybe(d)) top.insert(d, __ldg(sidx + pos));
  }
};

// D = 0: runtime dimension (generic metrics)
template <int D, int K, bool EUCLID>
__global__ void __launch_bounds__(128)
grid_knn_kernel(GridDev g, pomp_metric_desc m, const double* __restrict__ pts, long long n,
                const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq,
                const int* __restrict__ qperm, int k, int64_t* __restrict__ out_idx,
                double* __restrict__ out_dist, int32_t* __restrict__ out_cnt) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  long long qi = qperm ? (long long)qperm[t] : t;
  constexpr int DA = D ? D : POMP_MAX_DIM;
  const int dim = D ? D : m.dim;
  double qq[DA];
#pragma unroll
  for (int j = 0; j < DA; j++) qq[j] = j < dim ? q[qi * dim + j] : 0.0;
  ThreadTopK<K> top;
  top.init(k);
  KnnColl<K> coll{top, g.sidx};
  // phase A: any k real points bound the k-th distance from above; take them next to the query's
  // home cell in the sorted order (spatially close), growing the window while masked points hide it
  long long home = 0;
  for (int a = 0; a < g.G; a++) home += (long long)grid_cell_p(g, a, qq[g.gdim[a]]) * g.stride[a];
  int hs = __ldg(g.cell_start + home), he = __ldg(g.cell_start + home + 1);
  int center = (hs + he) >> 1;
  int w0 = max(0, center - K), w1 = min(g.n_indexed, center + K);
  {
    int a0 = w0, a1 = w1;
    while (true) {
      for (int pos = a0; pos < a1; pos++) {
        if (skip && skip[__ldg(g.sidx + pos)]) continue;
        double p[DA];
        if (D) load_point<DA>(g.spts, pos, p);
        else
          for (int j = 0; j < dim; j++) p[j] = __ldg(g.spts + (long long)pos * dim + j);
        coll.offer(D ? point_dist<DA, EUCLID>(m, p, qq) : metric_eval(m, p, qq), pos);
      }
      if (top.full() || (w0 == 0 && w1 == g.n_indexed)) break;
      // grow to the left, then to the right
      int len = w1 - w0;
      if (w0 > 0) {
        a1 = w0;
        w0 = max(0, w0 - len);
        a0 = w0;
      } else {
        a0 = w1;
        w1 = min(g.n_indexed, w1 + len);
        a1 = w1;
      }
    }
  }
  grid_visit<D, EUCLID>(g, m, skip, qq, coll, w0, w1);
  // points appended after the last index build
  for (long long i = g.n_indexed; i < n; i++) {
    if (skip && skip[i]) continue;
    double p[DA];
    for (int j = 0; j < dim; j++) p[j] = __ldg(pts + i * dim + j);
    double d = (D && EUCLID) ? euclid_fixed<DA>(p, qq) : metric_eval(m, p, qq);
    if (top.maybe(d)) top.insert(d, (int)i);
  }
  int cnt = 0;
#pragma unroll(K <= 32 ? K : 1)
  for (int j = 0; j < K; j++)
    if (j < k) {
      bool fin = top.d[j] < POMP_INF;
      out_idx[qi * k + j] = fin ? (int64_t)top.i[j] : -1;
      out_dist[qi * k + j] = top.d[j];
      cnt += fin;
    }
  if (out_cnt) out_cnt[qi] = cnt;
}

template <bool FILL>
struct RadiusColl {
  double radius;
  int inclusive;
  long long cnt;
  const int* __restrict__ sidx;
  int64_t* out;
  __device__ __forceinline__ double threshold() const { return radius; }
  __device__ __forceinline__ void offer_idx(double d, long long idx) {
    bool hit = inclusive ? (d <= radius) : (d < radius);
    if (hit) {
      if (FILL) out[cnt] = idx;
      cnt++;
    }
  }
  __device__ __forceinline__ void offer(double d, long long pos) {
    bool hit = inclusive ? (d <= radius) : (d < radius);
    if (hit) {
      if (FILL) out[cnt] = __ldg(sidx + pos);
      cnt++;
    }
  }
};

template <int D, bool EUCLID, bool FILL>
__global__ void __launch_bounds__(128)
grid_radius_kernel(GridDev g, pomp_metric_desc m, const double* __restrict__ pts, long long n,
                   const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq, double radius,
                   int inclusive, int* __restrict__ counts, const int64_t* __restrict__ offsets,
                   int64_t* __restrict__ out) {
  long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  constexpr int DA = D ? D : POMP_MAX_DIM;
  const int dim = D ? D : m.dim;
  double qq[DA];
#pragma unroll
  for (int j = 0; j < DA; j++) qq[j] = j < dim ? q[qi * dim + j] : 0.0;
  RadiusColl<FILL> coll{radius, inclusive, 0, g.sidx, FILL ? out + offsets[qi] : nullptr};
  grid_visit<D, EUCLID>(g, m, skip, qq, coll, 0, 0);
  for (long long i = g.n_indexed; i < n; i++) {
    if (skip && skip[i]) continue;
    double p[DA];
    for (int j = 0; j < dim; j++) p[j] = __ldg(pts + i * dim + j);
    coll.offer_idx((D && EUCLID) ? euclid_fixed<DA>(p, qq) : metric_eval(m, p, qq), i);
  }
  if (FILL) {
    if (coll.cnt > 32) heapsort_i64(out + offsets[qi], coll.cnt);  // short segments: segment_sort32_kernel
  } else {
    counts[qi] = (int)coll.cnt;
  }
}

// ---- Gaussian kernel density (EST): sum over the points within `radius` of exp(-(d/sigma)^2) ------
// pomp/planners/kinodynamicplanner.py:688-693 (EST.density) and :676-686 (onAddNode): a radius query
// whose hits are folded into one number, so nothing is materialised.
struct DensityColl {
  double radius, inv_sigma;
  int inclusive;
  double sum;
  int cnt;
  __device__ __forceinline__ double threshold() const { return radius; }
  __device__ __forceinline__ void add(double d) {
    bool hit = inclusive ? (d <= radius) : (d < radius);
    if (hit) {
      double t = d * inv_sigma;
      sum += exp(-(t * t));
      cnt++;
    }
  }
  __device__ __forceinline__ void offer(double d, long long) { add(d); }
};

__global__ void __launch_bounds__(128)
grid_density_kernel(GridDev g, pomp_metric_desc m, const double* __restrict__ pts, long long n,
                    const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq, double radius,
                    double sigma, int inclusive, double* __restrict__ dens, int32_t* __restrict__ cnt) {
  long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const int dim = m.dim;
  double qq[POMP_MAX_DIM];
#pragma unroll
  for (int j = 0; j < POMP_MAX_DIM; j++) qq[j] = j < dim ? q[qi * dim + j] : 0.0;
  DensityColl coll{radius, 1.0 / sigma, inclusive, 0.0, 0};
  grid_visit<0, false>(g, m, skip, qq, coll, 0, 0);
  for (long long i = g.n_indexed; i < n; i++) {
    if (skip && skip[i]) continue;
    double p[POMP_MAX_DIM];
    for (int j = 0; j < dim; j++) p[j] = __ldg(pts + i * dim + j);
    coll.add(metric_eval(m, p, qq));
  }
  dens[qi] = coll.sum;
  if (cnt) cnt[qi] = coll.cnt;
}

// small point sets: warp per query, lanes stride over the points
__global__ void bf_density_kernel(const double* __restrict__ pts, long long n, pomp_metric_desc m,
                                  const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq,
                                  double radius, double sigma, int inclusive, double* __restrict__ dens,
                                  int32_t* __restrict__ cnt) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= nq) return;
  const int D = m.dim;
  double qq[POMP_MAX_DIM], p[POMP_MAX_DIM];
  load_q(q + warp * D, D, qq);
  DensityColl coll{radius, 1.0 / sigma, inclusive, 0.0, 0};
  for (long long i = lane; i < n; i += 32) {
    if (skip && skip[i]) continue;
    for (int j = 0; j < D; j++) p[j] = __ldg(pts + i * D + j);
    coll.add(metric_eval(m, p, qq));
  }
  double s = coll.sum;
  int c = coll.cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(FULL_MASK, s, o);
    c += __shfl_xor_sync(FULL_MASK, c, o);
  }
  if (lane == 0) {
    dens[warp] = s;
    if (cnt) cnt[warp] = c;
  }
}

// ============================================================================ host: grid build
static void axis_scales(const pomp_metric_desc& m, double* s /*[base_dim]*/) {
  int d = base_dim(m);
  for (int i = 0; i < d; i++) s[i] = 0.0;
  if (m.has_cost && m.cost_weight_set && !(m.cost_weight >= 0.0)) return;  // bound needs cost term >= 0
  switch (m.kind) {
    case POMP_METRIC_EUCLIDEAN:
    case POMP_METRIC_L1:
    case POMP_METRIC_LINF:
      for (int i = 0; i < d; i++) s[i] = 1.0;
      break;
    case POMP_METRIC_SCALED_EUCLIDEAN:
      if (m.scale > 0.0)
        for (int i = 0; i < d; i++) s[i] = m.scale;
      break;
    case POMP_METRIC_WEIGHTED_EUCLIDEAN: {
      bool ok = true;
      for (int i = 0; i < d; i++) ok = ok && (m.dim_weight[i] >= 0.0);
      if (ok)
        for (int i = 0; i < d; i++) s[i] = m.dim_weight[i] > 0.0 ? sqrt(m.dim_weight[i]) : 0.0;
      break;
    }
    case POMP_METRIC_GEODESIC_PRODUCT: {
      bool ok = true;
      for (int c = 0; c < m.n_comp; c++) ok = ok && (m.comp_weight[c] >= 0.0);
      if (!ok) break;
      int i = 0;
      for (int c = 0; c < m.n_comp; c++)
        for (int j = 0; j < m.comp_dim[c]; j++, i++)
          s[i] = (m.comp_kind[c] == POMP_COMP_CARTESIAN && m.comp_weight[c] > 0.0) ? sqrt(m.comp_weight[c]) : 0.0;
      break;
    }
  }
}

static bool is_plain_euclid(const pomp_metric_desc& m);
static int build_grid(pomp_nn* h, cudaStream_t st) {
  POMP_TRY(nn_flush(h, st));
  h->grid_valid = false;
  const int D = h->D;
  const int64_t n = h->n_dev;
  POMP_REQUIRE(n > 0, "cannot index an empty point set");
  POMP_REQUIRE(n < (int64_t)0x7fffffff, "more than 2^31-1 points per shard are not supported");
  double s[POMP_MAX_DIM];
  axis_scales(h->metric, s);
  // bounding box
  int nb = std::min<int64_t>((n + 255) / 256, (int64_t)sm_count(h->device) * 8);
  POMP_TRY(h->part_d.reserve((size_t)nb * 2 * POMP_MAX_DIM));
  if (D == 2) LAUNCH(bbox_kernel<2>, nb, 256, 0, st, h->d_pts, (long long)n, D, h->part_d.p);
  else if (D == 3) LAUNCH(bbox_kernel<3>, nb, 256, 0, st, h->d_pts, (long long)n, D, h->part_d.p);
  else if (D == 4) LAUNCH(bbox_kernel<4>, nb, 256, 0, st, h->d_pts, (long long)n, D, h->part_d.p);
  else if (D == 6) LAUNCH(bbox_kernel<6>, nb, 256, 0, st, h->d_pts, (long long)n, D, h->part_d.p);
  else LAUNCH(bbox_kernel<0>, nb, 256, 0, st, h->d_pts, (long long)n, D, h->part_d.p);
  CHECK_LAUNCH();
  std::vector<double> part((size_t)nb * 2 * POMP_MAX_DIM);
  CUDA_TRY(cudaMemcpyAsync(part.data(), h->part_d.p, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  double lo[POMP_MAX_DIM], hi[POMP_MAX_DIM];
  for (int j = 0; j < D; j++) {
    lo[j] = INFINITY;
    hi[j] = -INFINITY;
    for (int b = 0; b < nb; b++) {
      lo[j] = fmin(lo[j], part[((size_t)b * 2 + 0) * POMP_MAX_DIM + j]);
      hi[j] = fmax(hi[j], part[((size_t)b * 2 + 1) * POMP_MAX_DIM + j]);
    }
  }
  GridDev g;
  memset(&g, 0, sizeof(g));
  int G = 0;
  double wext[POMP_MAX_DIM];
  for (int j = 0; j < base_dim(h->metric); j++) {
    double ext = hi[j] - lo[j];
    if (s[j] > 0.0 && isfinite(lo[j]) && isfinite(hi[j]) && ext > 0.0 && isfinite(ext)) {
      g.gdim[G] = j;
      g.lo[G] = lo[j];
      g.hi[G] = hi[j];
      wext[G] = ext * s[j];
      g.rscale[G] = (1.0 + 1e-9) / s[j];
      g.slack[G] = 1e-9 * fmax(fmax(fabs(lo[j]), fabs(hi[j])), ext) + 1e-300;
      G++;
    }
  }
  // SO(2) components of a product metric become PERIODIC axes (after the Cartesian ones: axis 0,
  // the contiguous run, is never periodic): angles are bucketed by so2.normalize over [0, 2 pi)
  if (G > 0 && h->metric.kind == POMP_METRIC_GEODESIC_PRODUCT && h->periodic_axes != 0) {
    int j = 0;
    for (int c = 0; c < h->metric.n_comp; c++) {
      if (h->metric.comp_kind[c] == POMP_COMP_SO2 && h->metric.comp_weight[c] > 0.0 && G < POMP_MAX_DIM &&
          isfinite(lo[j]) && isfinite(hi[j])) {
        g.gdim[G] = j;
        g.lo[G] = 0.0;
        g.hi[G] = 6.283185307179586;
        g.periodic[G] = 1;
        const double sc = sqrt(h->metric.comp_weight[c]);
        wext[G] = 6.283185307179586 * sc;
        g.rscale[G] = (1.0 + 1e-9) / sc;
        g.slack[G] = 1e-8;
        G++;
      }
      j += h->metric.comp_kind[c] == POMP_COMP_SO2 ? 1 : h->metric.comp_dim[c];
    }
  }
  if (G == 0) {
    h->grid_unusable = true;
    set_error("metric/point set gives no usable grid axis");
    return POMP_ERR_INVALID;
  }
  // measured on B200 (bench.py, 2^24 points): 2-D nearest 0.097 ms at 1.5 vs 0.106 ms at 2.0 points per cell
  // (k = 16 prefers 2.0: 1.17 vs 1.44 ms -- set the points_per_cell knob for k-NN-heavy callers)
  // generic metrics (SE(2) states, 2^24 points): 3.1 ms at 2, 2.8 ms at 4 points per cell (profiles/se2_probe.py)
  const bool generic3 = G == 3 && !is_plain_euclid(h->metric);
  // 4-D: 4 beats 2, 8, 16 for both kernel families.  6-D Euclidean (warp-per-query shell kernel, whose cost is the row
  // enumeration): 8 beats 4 -- k = 1 / 4 / 8 / 16: 6.9 / 10.8 / 14.2 / 19.4 ms vs 7.1 / 11.7 / 15.7 / 21.7 ms
  // (profiles/ppc6_probe.py); the thread-per-query kernels preferred 4 there (k = 8: 14.9 vs 20.3 ms at 8)
  const bool warp6 = G >= 6 && is_plain_euclid(h->metric) && h->warp_search != 0;
  double ppc = h->points_per_cell > 0 ? h->points_per_cell
                                      : (G == 2 ? 1.5 : (G == 3 ? (generic3 ? 4.0 : 2.0) : (warp6 ? 8.0 : 4.0)));
  double target = fmax(1.0, (double)n / ppc);
  target = fmin(target, 268435456.0);
  // cubic cells in metric space: ncell_a proportional to the weighted extent
  double logvol = 0;
  for (int a = 0; a < G; a++) logvol += log(wext[a]);
  double cell = exp((logvol - log(target)) / G);
  long long total = 1;
  for (int a = 0; a < G; a++) {
    long long nc = (long long)floor(wext[a] / cell + 0.5);
    nc = std::max<long long>(1, std::min<long long>(nc, 1 << 20));
    if (a == 0 && G == 2 && nc >= 64) nc = (nc + 3) / 4 * 4;  // 16-byte aligned cell_start slices for the TMA path
    g.ncell[a] = (int)nc;
    total *= nc;
  }
  while (total > 536870912LL) {  // safety cap: halve the finest axis
    int am = 0;
    for (int a = 1; a < G; a++)
      if (g.ncell[a] > g.ncell[am]) am = a;
    total /= g.ncell[am];
    g.ncell[am] = std::max(1, g.ncell[am] / 2);
    total *= g.ncell[am];
  }
  long long stride = 1;
  for (int a = 0; a < G; a++) {
    double ext = g.hi[a] - g.lo[a];
    g.h[a] = ext / (double)g.ncell[a];
    g.inv_h[a] = (double)g.ncell[a] / ext;
    g.stride[a] = stride;
    stride *= g.ncell[a];
  }
  g.G = G;
  g.n_cells = total;
  g.n_indexed = (int)n;
  // counting sort of the points by cell id
  POMP_TRY(h->cell_start.reserve((size_t)total + 1 + 512));  // slack: the TMA path copies fixed-width slices
  POMP_TRY(h->tmp_i32.reserve((size_t)2 * n + (size_t)total));
  POMP_TRY(h->scan_tmp.reserve(scan_scratch_elems(total)));
  POMP_TRY(h->spts.reserve((size_t)n * D));
  POMP_TRY(h->sidx.reserve((size_t)n));
  int* slot = h->tmp_i32.p;
  int* cellid = h->tmp_i32.p + n;
  int* counts = h->tmp_i32.p + 2 * n;
  CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)total, st));
  int blocks = (int)((n + 255) / 256);
  LAUNCH(cell_count_kernel, blocks, 256, 0, st, g, h->d_pts, (long long)n, D, counts, slot, cellid);
  POMP_TRY(exclusive_scan<int>(counts, total, h->cell_start.p, h->scan_tmp.p, st));
  LAUNCH(cell_scatter_kernel, blocks, 256, 0, st, (long long)n, h->cell_start.p, slot, cellid, h->sidx.p);
  if (D == 2) LAUNCH((cell_gather_kernel<2>), blocks, 256, 0, st, h->d_pts, (long long)n, D, h->sidx.p, h->spts.p);
  else LAUNCH((cell_gather_kernel<0>), blocks, 256, 0, st, h->d_pts, (long long)n, D, h->sidx.p, h->spts.p);
  CHECK_LAUNCH();
  g.cell_start = h->cell_start.p;
  g.spts = h->spts.p;
  g.sidx = h->sidx.p;
  h->grid = g;
  h->grid_valid = true;
  h->grid_unusable = false;
  h->grid_builds++;
  return POMP_OK;
}

// decide the regime for a batch of nq queries; builds the index when it pays
static int want_grid(pomp_nn* h, int64_t nq, bool* use, cudaStream_t st) {
  *use = false;
  int64_t n = h->size();
  if (h->grid_unusable || n < (int64_t)h->grid_min_points || nq < (int64_t)h->grid_min_queries) return POMP_OK;
  POMP_TRY(nn_flush(h, st));
  if (!h->grid_valid || (h->n_dev - h->grid.n_indexed) > (int64_t)h->max_tail) {
    int rc = build_grid(h, st);
    if (rc == POMP_ERR_INVALID && h->grid_unusable) return POMP_OK;  // fall back to brute force kernels
    POMP_TRY(rc);
  }
  *use = true;
  return POMP_OK;
}

static SplitGeom split_geom(pomp_nn* h, int64_t nq, int64_t n) {
  SplitGeom sg;
  long long want_warps = (long long)sm_count(h->device) * 32;
  long long ns = nq >= want_warps ? 1 : (want_warps + nq - 1) / std::max<int64_t>(nq, 1);
  long long max_ns = std::max<long long>(1, (n + 255) / 256);  // at least 256 points per warp
  ns = std::max<long long>(1, std::min(ns, max_ns));
  sg.nsplit = (int)ns;
  sg.chunk = std::max<long long>(1, (n + ns - 1) / ns);
  return sg;
}

static bool is_plain_euclid(const pomp_metric_desc& m) { return m.kind == POMP_METRIC_EUCLIDEAN && !m.has_cost; }

// ---- k-NN dispatch ---------------------------------------------------------------------------
static int launch_grid_knn_generic(pomp_nn* h, const double* q, int64_t nq, const int* qperm, int k, int64_t* idx,
                                   double* dist, int32_t* cnt, cudaStream_t st) {
  int blocks = (int)((nq + 127) / 128);
  const uint8_t* skip = h->skip_ptr();
#define GK(KK)                                                                                              \
  LAUNCH((grid_knn_kernel<0, KK, false>), blocks, 128, 0, st, h->grid, h->metric, h->d_pts, (long long)h->n_dev, \
         skip, q, (long long)nq, qperm, k, idx, dist, cnt)
  prof_begin(h, st);
  if (k == 1) GK(1);
  else if (k <= 4) GK(4);
  else if (k <= 8) GK(8);
  else if (k <= 16) GK(16);
  else if (k <= 32) GK(32);
  else if (k <= 64) GK(64);
  else GK(128);
#undef GK
  prof_end(h, st);
  CHECK_LAUNCH();
  return POMP_OK;
}

// plain Euclidean metric, every coordinate a grid axis: register-resident kernels (grid_fast.cuh)
template <int D>
static int launch_grid_knn_fast(pomp_nn* h, const double* q, int64_t nq, const int* qperm, int k, int64_t* idx,
                                double* dist, int32_t* cnt, cudaStream_t st) {
  int blocks = (int)((nq + 127) / 128);
  const uint8_t* skip = h->skip_ptr();
#define GK(KK)                                                                                                 \
  LAUNCH((grid_knn_fast_kernel<D, KK>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q, \
         (long long)nq, qperm, k, idx, dist, cnt, (const int*)nullptr, (const int*)nullptr)
  prof_begin(h, st);
  if (k == 1) GK(1);
  else if (k <= 4) GK(4);
  else if (k <= 8) GK(8);
  else if (k <= 16) GK(16);
  else GK(32);
#undef GK
  prof_end(h, st);
  CHECK_LAUNCH();
  return POMP_OK;
}

// 2-D / 3-D: fixed block of cells + certificate, overflow redone by the pruned traversal
template <int D>
static int launch_grid_knn_block(pomp_nn* h, const double* q, int64_t nq, const int* qperm, int k, int64_t* idx,
                                 double* dist, int32_t* cnt, cudaStream_t st) {
  const int KK = k == 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32)));
  // block radius: expect >= ~2.5 K points inside the certified ball of radius R cells
  double ppc = (double)h->grid.n_indexed / (double)h->grid.n_cells;
  double vol = D == 2 ? 3.141592653589793 : 4.1887902047863905;
  int R = 1;
  while (R < 6 && ppc * vol * pow((double)R, D) < 2.5 * KK + 3.0) R++;
  POMP_TRY(h->ss[h->cur].ovf.reserve((size_t)nq + 1));
  int* ovf_count = h->ss[h->cur].ovf.p;
  int* ovf_list = h->ss[h->cur].ovf.p + 1;
  if (!h->ss[h->cur].ovf_zeroed) CUDA_TRY(cudaMemsetAsync(ovf_count, 0, sizeof(int), st));
  int blocks = (int)((nq + 127) / 128);
  int oblocks = std::min(blocks, sm_count(h->device) * 4);
  const uint8_t* skip = h->skip_ptr();
#define GK(KC)                                                                                                   \
  do {                                                                                                           \
    LAUNCH((grid_knn_block_kernel<D, KC>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q, \
           (long long)nq, qperm, k, R, idx, dist, cnt, ovf_list, ovf_count);                                     \
    prof_end(h, st);                                                                                             \
    LAUNCH((grid_knn_fast_kernel<D, KC>), oblocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q, \
           (long long)nq, (const int*)nullptr, k, idx, dist, cnt, (const int*)ovf_list, (const int*)ovf_count);  \
  } while (0)
  prof_begin(h, st);
  if (KK == 1) GK(1);
  else if (KK == 4) GK(4);
  else if (KK == 8) GK(8);
  else if (KK == 16) GK(16);
  else GK(32);
#undef GK
  CHECK_LAUNCH();
  return POMP_OK;
}

struct SortInfo {
  const int* qperm = nullptr;
  const double* qsorted = nullptr;
  const int* qstart = nullptr;
  int shift = 0, nbx = 0;
};

static bool knn_is_fast(const pomp_nn* h, int k) { return is_plain_euclid(h->metric) && h->grid.G == h->D && k <= 32; }

// counting sort of one query batch by coarse home bin into scratch set S (stream st)
static int grid_knn_sort(pomp_nn* h, SortScratch& S, const double* q, int64_t nq, int k, SortInfo* info,
                         cudaStream_t st) {
  const bool fast = knn_is_fast(h, k);
  info->shift = 0;
  info->nbx = h->grid.ncell[0];
  S.ovf_zeroed = false;
  if (!(h->sort_queries != 0 && nq >= (int64_t)h->sort_min_queries && h->grid.n_cells <= 0x7fffffff)) return POMP_OK;
  int shift = 0;
  // 2^shift-cell bins along axis 0
  if (h->grid.ncell[0] >= 256) shift = (int)h->sort_shift;
  int nbx = (h->grid.ncell[0] + (1 << shift) - 1) >> shift;
  const int row_div = 1;
  long long nrow = h->grid.n_cells / h->grid.ncell[0];
  long long nb = ((nrow + row_div - 1) / row_div) * nbx;
  POMP_TRY(S.counts.reserve((size_t)nb + 1 + (size_t)nb + 1));
  POMP_TRY(S.qcell.reserve((size_t)nq * 2));
  POMP_TRY(S.qperm.reserve((size_t)nq));
  POMP_TRY(S.ovf.reserve(2 * (size_t)nq + 2));  // two overflow lists; sized HERE: the launchers must not reallocate
                                                 // after query_cell_kernel zeroed the first counter
  if (fast) POMP_TRY(S.qsorted.reserve((size_t)nq * h->D));
  POMP_TRY(S.scan_tmp.reserve(scan_scratch_elems(nb)));
  int* counts = S.counts.p;
  int* start = S.counts.p + nb + 1;
  CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)nb, st));
  int blocks = (int)((nq + 255) / 256);
  LAUNCH(query_cell_kernel, blocks, 256, 0, st, h->grid, q, (long long)nq, h->D, shift, nbx, row_div, counts, S.qcell.p,
         S.qcell.p + nq, S.ovf.p);
  S.ovf_zeroed = true;
  if (nb <= 8192) LAUNCH((scan_small_kernel<int>), 1, 1024, 0, st, counts, (int)nb, start);
  else POMP_TRY(exclusive_scan<int>(counts, nb, start, S.scan_tmp.p, st));
  LAUNCH(query_perm_kernel, blocks, 256, 0, st, S.qcell.p, S.qcell.p + nq, start, q, (long long)nq, h->D, S.qperm.p,
         fast ? S.qsorted.p : (double*)nullptr);
  CHECK_LAUNCH();
  info->qperm = S.qperm.p;
  info->qstart = start;
  if (fast) info->qsorted = S.qsorted.p;
  info->shift = shift;
  info->nbx = nbx;
  return POMP_OK;
}

// search of one (sorted or unsorted) batch; uses the scratch set h->cur for its overflow list
static int grid_knn_search(pomp_nn* h, const SortInfo& si, const double* q, int64_t nq, int k, int64_t* idx,
                           double* dist, int32_t* cnt, cudaStream_t st) {
  const int* qperm = si.qperm;
  const double* qsorted = si.qsorted;
  const bool fast = knn_is_fast(h, k);
  if (fast && warp_knn_eligible(h, k)) return launch_grid_knn_warp(h, q, qsorted, nq, qperm, k, idx, dist, cnt, st);
  if (fast) {
    switch (h->D) {
      case 2: {
        if (h->block_search != 0) {
          const int KK = k == 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32)));
          double ppc = (double)h->grid.n_indexed / (double)h->grid.n_cells;
          // expected points inside the certified ball of radius R cells: generous for tiny k (a miss
          // costs an in-thread retry), tight for k >= 8 (misses are compacted into a second pass)
          double need = KK >= 8 ? 1.4 * KK + 2.0 : (KK == 1 ? 4.0 : 2.5 * KK + 3.0);  // k = 1: misses are cheap (warp retry)
          if (KK == 1 && h->block_need > 0) need = h->block_need;
          int R = 1;
          while (R < 6 && ppc * 3.141592653589793 * R * R < need) R++;
          if (R <= 3) return launch_grid_knn_flat2d(h, q, qsorted, nq, qperm, k, R, idx, dist, cnt, st);
          return launch_grid_knn_block<2>(h, q, nq, qperm, k, idx, dist, cnt, st);
        }
        return launch_grid_knn_fast<2>(h, q, nq, qperm, k, idx, dist, cnt, st);
      }
      case 3:
        // 3-D: the pruned traversal beats the fixed 5x5x5 block (k=1: 0.49 vs 0.61 ms, k=16: 5.1 vs 6.9 ms at
        // 2^24 points); block_search = 2 still selects the block kernel
        return h->block_search >= 2 ? launch_grid_knn_block<3>(h, q, nq, qperm, k, idx, dist, cnt, st)
                                    : launch_grid_knn_fast<3>(h, q, nq, qperm, k, idx, dist, cnt, st);
      case 4: return launch_grid_knn_fast<4>(h, q, nq, qperm, k, idx, dist, cnt, st);
      case 6: return launch_grid_knn_fast<6>(h, q, nq, qperm, k, idx, dist, cnt, st);
      default: break;
    }
  }
  {
    // SE(2) product metric (R^2 x SO(2), no cost lift) on its 3-axis grid: compile-time kernel (grid_se2.cuh)
    const pomp_metric_desc& m = h->metric;
    const GridDev& g = h->grid;
    const bool se2 = h->se2_fast != 0 && m.kind == POMP_METRIC_GEODESIC_PRODUCT && !m.has_cost && m.n_comp == 2 &&
                     m.comp_kind[0] == POMP_COMP_CARTESIAN && m.comp_dim[0] == 2 && m.comp_kind[1] == POMP_COMP_SO2 &&
                     m.comp_weight[0] > 0.0 && m.comp_weight[1] > 0.0 && h->D == 3 && g.G == 3 && g.gdim[0] == 0 &&
                     g.gdim[1] == 1 && g.gdim[2] == 2 && g.periodic[2] == 1 && !g.periodic[0] && !g.periodic[1] &&
                     k <= 32;
    if (se2) {
      const int blocks = (int)((nq + 127) / 128);
      const uint8_t* skip = h->skip_ptr();
#define SK(KK)                                                                                                    \
  LAUNCH((grid_knn_se2_kernel<KK>), blocks, 128, 0, st, g, m.comp_weight[0], m.comp_weight[1], h->d_pts,            \
         (long long)h->n_dev, skip, q, (long long)nq, qperm, k, idx, dist, cnt)
      prof_begin(h, st);
      if (k == 1) SK(1);
      else if (k <= 4) SK(4);
      else if (k <= 8) SK(8);
      else if (k <= 16) SK(16);
      else SK(32);
#undef SK
      prof_end(h, st);
      CHECK_LAUNCH();
      return POMP_OK;
    }
  }
  return launch_grid_knn_generic(h, q, nq, qperm, k, idx, dist, cnt, st);
}

// Large batches are cut into chunks: the bucket sort of chunk c+1 runs on an auxiliary stream
// while chunk c is searched on the caller's stream (two scratch sets, events for the hand-over),
// so the sort's latency-bound little kernels hide behind the search kernel.
static int grid_knn(pomp_nn* h, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* cnt,
                    cudaStream_t st) {
  const bool sorting = h->sort_queries != 0 && nq >= (int64_t)h->sort_min_queries && h->grid.n_cells <= 0x7fffffff;
  const int C = (int)h->overlap_chunks;
  // dense 2-D nearest() batches: streaming kernel over the strip index (strip2d.cu), tile-major query sort
  if (strip_eligible(h, nq, k)) return strip_nearest(h, q, nq, k, idx, dist, cnt, st);
  if (!(sorting && C > 1 && nq >= (int64_t)h->overlap_min_queries)) {
    h->cur = 0;
    SortInfo si;
    POMP_TRY(grid_knn_sort(h, h->ss[0], q, nq, k, &si, st));
    return grid_knn_search(h, si, q, nq, k, idx, dist, cnt, st);
  }
  if (!h->aux_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
  while ((int)h->ovl_ev.size() < 1 + 2 * C) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->ovl_ev.push_back(e);
  }
  cudaStream_t aux = h->aux_stream;
  CUDA_TRY(cudaEventRecord(h->ovl_ev[0], st));
  CUDA_TRY(cudaStreamWaitEvent(aux, h->ovl_ev[0], 0));
  int64_t per = ((nq + C - 1) / C + 127) / 128 * 128;
  for (int c = 0; c < C; c++) {
    int64_t b = (int64_t)c * per, e = std::min<int64_t>(nq, b + per);
    if (b >= e) break;
    int64_t m = e - b;
    cudaEvent_t ev_sorted = h->ovl_ev[1 + 2 * c], ev_done = h->ovl_ev[2 + 2 * c];
    h->cur = c & 1;
    if (c >= 2) CUDA_TRY(cudaStreamWaitEvent(aux, h->ovl_ev[2 + 2 * (c - 2)], 0));  // scratch set free again
    SortInfo si;
    POMP_TRY(grid_knn_sort(h, h->ss[h->cur], q + b * h->D, m, k, &si, aux));
    CUDA_TRY(cudaEventRecord(ev_sorted, aux));
    CUDA_TRY(cudaStreamWaitEvent(st, ev_sorted, 0));
    POMP_TRY(grid_knn_search(h, si, q + b * h->D, m, k, idx + b * k, dist + b * k, cnt ? cnt + b : nullptr, st));
    CUDA_TRY(cudaEventRecord(ev_done, st));
  }
  return POMP_OK;
}

static int bf_knn(pomp_nn* h, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* cnt,
                  cudaStream_t st) {
  const int64_t n = h->n_dev;
  SplitGeom sg = split_geom(h, nq, n);
  long long warps = (long long)nq * sg.nsplit;
  int blocks = (int)((warps * 32 + 255) / 256);
  const uint8_t* skip = h->skip_ptr();
  if (k == 1) {
    const bool tiled = is_plain_euclid(h->metric) && nq >= 1024 && h->bf_tiled != 0 &&
                       (h->D == 2 || h->D == 3 || h->D == 4 || h->D == 6);
    if (tiled) {
      constexpr int QT = 4;  // 8 queries per thread (124 registers) measured slower: 0.75 vs 0.45 ms
      const int xb = (int)((nq + 256 * QT - 1) / (256 * QT));
      long long ns = std::max<long long>(1, (2LL * sm_count(h->device) + xb - 1) / xb);
      ns = std::min<long long>(ns, std::max<long long>(1, (n + 2047) / 2048));
      ns = std::min<long long>(ns, 65535);
      sg.nsplit = (int)ns;
      sg.chunk = std::max<long long>(1, (n + ns - 1) / ns);
      POMP_TRY(h->part_d.reserve((size_t)nq * ns));
      POMP_TRY(h->part_i.reserve((size_t)nq * ns));
      dim3 grid((unsigned)xb, (unsigned)ns);
#define BT(DD, QQ)                                                                                              \
  LAUNCH((bf_nearest_tiled_kernel<DD, QQ>), grid, 256, 0, st, h->d_pts, (long long)n, skip, q, (long long)nq, sg, \
         h->part_d.p, h->part_i.p)
      if (h->D == 2) BT(2, 4);
      else if (h->D == 3) BT(3, 4);
      else if (h->D == 4) BT(4, 4);
      else BT(6, 4);
#undef BT
    } else {
      POMP_TRY(h->part_d.reserve((size_t)warps));
      POMP_TRY(h->part_i.reserve((size_t)warps));
      LAUNCH(bf_nearest_kernel, blocks, 256, 0, st, h->d_pts, (long long)n, h->metric, skip, q, (long long)nq, sg,
             h->part_d.p, h->part_i.p);
    }
    LAUNCH(bf_nearest_reduce_kernel, (int)((nq + 255) / 256), 256, 0, st, h->part_d.p, h->part_i.p, (long long)nq,
           sg.nsplit, idx, dist);
    CHECK_LAUNCH();
    if (cnt) {
      // count = (idx >= 0); done by the merge kernel path below for k>1, here by a tiny memset-free trick:
      // reuse merge kernel with one list of length 1
      LAUNCH((merge_lists_kernel<1, int64_t>), (int)((nq * 32 + 255) / 256), 256, 0, st, dist, idx, 1, (long long)nq, 1,
             1LL, 1LL, idx, dist, cnt);
      CHECK_LAUNCH();
    }
    return POMP_OK;
  }
  POMP_TRY(h->part_d.reserve((size_t)warps * k));
  POMP_TRY(h->part_i.reserve((size_t)warps * k));
  int mblocks = (int)((nq * 32 + 255) / 256);
#define BK(KL)                                                                                                     \
  do {                                                                                                             \
    LAUNCH((bf_knn_kernel<KL>), blocks, 256, 0, st, h->d_pts, (long long)n, h->metric, skip, q, (long long)nq, k, sg, \
           h->part_d.p, h->part_i.p);                                                                              \
    LAUNCH((merge_lists_kernel<KL, int>), mblocks, 256, 0, st, h->part_d.p, h->part_i.p, sg.nsplit, (long long)nq, k, \
           (long long)sg.nsplit * k, (long long)k, idx, dist, cnt);                                                \
  } while (0)
  if (k <= 32) BK(1);
  else if (k <= 64) BK(2);
  else BK(4);
#undef BK
  CHECK_LAUNCH();
  return POMP_OK;
}

static int knn_device(pomp_nn* h, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* cnt,
                      cudaStream_t st) {
  POMP_REQUIRE(k >= 1 && k <= 128, "k=%d outside 1..128", k);
  if (nq == 0) return POMP_OK;
  POMP_TRY(nn_flush(h, st));
  bool use_grid = false;
  POMP_TRY(want_grid(h, nq, &use_grid, st));
  if (use_grid) return grid_knn(h, q, nq, k, idx, dist, cnt, st);
  return bf_knn(h, q, nq, k, idx, dist, cnt, st);
}

// ---- radius dispatch -------------------------------------------------------------------------
template <int D>
static int launch_grid_radius_fast(pomp_nn* h, bool fill, const double* q, int64_t nq, double radius, int inclusive,
                                   int* counts, const int64_t* offsets, int64_t* out, cudaStream_t st) {
  int blocks = (int)((nq + 127) / 128);
  const uint8_t* skip = h->skip_ptr();
  if (fill)
    LAUNCH((grid_radius_fast_kernel<D, true>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,
           (long long)nq, radius, inclusive, counts, offsets, out);
  else
    LAUNCH((grid_radius_fast_kernel<D, false>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,
           (long long)nq, radius, inclusive, counts, offsets, out);
  CHECK_LAUNCH();
  return POMP_OK;
}

static int grid_radius(pomp_nn* h, bool fill, const double* q, int64_t nq, double radius, int inclusive,
                       int* counts, const int64_t* offsets, int64_t* out, cudaStream_t st) {
  if (is_plain_euclid(h->metric) && h->grid.G == h->D) {
    switch (h->D) {
      case 2: return launch_grid_radius_fast<2>(h, fill, q, nq, radius, inclusive, counts, offsets, out, st);
      case 3: return launch_grid_radius_fast<3>(h, fill, q, nq, radius, inclusive, counts, offsets, out, st);
      case 4: return launch_grid_radius_fast<4>(h, fill, q, nq, radius, inclusive, counts, offsets, out, st);
      case 6: return launch_grid_radius_fast<6>(h, fill, q, nq, radius, inclusive, counts, offsets, out, st);
      default: break;
    }
  }
  int blocks = (int)((nq + 127) / 128);
  const uint8_t* skip = h->skip_ptr();
  if (fill)
    LAUNCH((grid_radius_kernel<0, false, true>), blocks, 128, 0, st, h->grid, h->metric, h->d_pts,
           (long long)h->n_dev, skip, q, (long long)nq, radius, inclusive, counts, offsets, out);
  else
    LAUNCH((grid_radius_kernel<0, false, false>), blocks, 128, 0, st, h->grid, h->metric, h->d_pts,
           (long long)h->n_dev, skip, q, (long long)nq, radius, inclusive, counts, offsets, out);
  CHECK_LAUNCH();
  return POMP_OK;
}

// offsets: device int64 [nq+1]; returns total through *needed_h (synchronises)
static int radius_device(pomp_nn* h, const double* q, int64_t nq, double radius, int inclusive, int64_t* offsets,
                         int64_t* idx, int64_t idx_capacity, int64_t* needed_h, cudaStream_t st) {
  if (needed_h) *needed_h = 0;
  if (nq == 0) {
    CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(int64_t), st));
    return POMP_OK;
  }
  POMP_TRY(nn_flush(h, st));
  bool use_grid = false;
  POMP_TRY(want_grid(h, nq, &use_grid, st));
  const int64_t n = h->n_dev;
  const uint8_t* skip = h->skip_ptr();
  int64_t total = 0;
  if (use_grid && is_plain_euclid(h->metric) && h->grid.G == h->D && nq <= (int64_t)4 << 20 &&
      (h->D == 2 || h->D == 3 || h->D == 4 || h->D == 6)) {
    // one traversal: count + stage, scan, warp-per-query finalize; a second traversal only for the
    // queries with more than RADIUS_STAGE_CAP hits
    POMP_TRY(h->counts.reserve((size_t)nq + 2));
    POMP_TRY(h->scan_tmp.reserve(scan_scratch_elems(nq)));
    POMP_TRY(h->rstage.reserve((size_t)nq * RADIUS_STAGE_CAP));
    int* n_long = h->counts.p + nq + 1;
    CUDA_TRY(cudaMemsetAsync(n_long, 0, sizeof(int), st));
    const int sb = (int)((nq + 127) / 128);
    // large batches are bucket-sorted by coarse home bin first (the same sort as the k-NN path): neighbouring threads
    // then walk neighbouring cells
    SortInfo rsi;
    h->cur = 0;
    POMP_TRY(grid_knn_sort(h, h->ss[0], q, nq, 1, &rsi, st));
#define RS(DD)                                                                                                       \
  LAUNCH((grid_radius_stage_kernel<DD>), sb, 128, 0, st, h->grid, h->d_pts, (long long)n, skip, q, rsi.qsorted,      \
         rsi.qperm, (long long)nq, radius, inclusive, h->counts.p, h->rstage.p, n_long)
    if (h->D == 2) RS(2);
    else if (h->D == 3) RS(3);
    else if (h->D == 4) RS(4);
    else RS(6);
#undef RS
    POMP_TRY(exclusive_scan<int64_t>(h->counts.p, nq, offsets, h->scan_tmp.p, st));
    int nl = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, offsets + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&nl, n_long, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (needed_h) *needed_h = total;
    if (total > idx_capacity) {
      set_error("radius query produced %lld hits, capacity %lld", (long long)total, (long long)idx_capacity);
      return POMP_ERR_CAPACITY;
    }
    if (total > 0) {
      LAUNCH(radius_finalize_kernel, (int)((nq * 32 + 255) / 256), 256, 0, st, h->counts.p, offsets, h->rstage.p,
             (long long)nq, idx);
      if (nl > 0) POMP_TRY(grid_radius(h, true, q, nq, radius, inclusive, h->counts.p, offsets, idx, st));
      CHECK_LAUNCH();
    }
    return POMP_OK;
  }
  if (use_grid) {
    POMP_TRY(h->counts.reserve((size_t)nq + 1));
    POMP_TRY(h->scan_tmp.reserve(scan_scratch_elems(nq)));
    POMP_TRY(grid_radius(h, false, q, nq, radius, inclusive, h->counts.p, nullptr, nullptr, st));
    POMP_TRY(exclusive_scan<int64_t>(h->counts.p, nq, offsets, h->scan_tmp.p, st));
    CUDA_TRY(cudaMemcpyAsync(&total, offsets + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (needed_h) *needed_h = total;
    if (total > idx_capacity) {
      set_error("radius query produced %lld hits, capacity %lld", (long long)total, (long long)idx_capacity);
      return POMP_ERR_CAPACITY;
    }
    if (total > 0) {
      POMP_TRY(grid_radius(h, true, q, nq, radius, inclusive, nullptr, offsets, idx, st));
      LAUNCH(segment_sort32_kernel, (int)((nq * 32 + 255) / 256), 256, 0, st, offsets, idx, (long long)nq);
      CHECK_LAUNCH();
    }
    return POMP_OK;
  }
  SplitGeom sg = split_geom(h, nq, n);
  long long warps = (long long)nq * sg.nsplit;
  int blocks = (int)((warps * 32 + 255) / 256);
  POMP_TRY(h->counts.reserve((size_t)warps + 1));
  POMP_TRY(h->out_off.reserve((size_t)warps + 1));
  POMP_TRY(h->scan_tmp.reserve(scan_scratch_elems(warps)));
  LAUNCH((bf_radius_kernel<false>), blocks, 256, 0, st, h->d_pts, (long long)n, h->metric, skip, q, (long long)nq,
         radius, inclusive, sg, h->counts.p, nullptr, nullptr);
  POMP_TRY(exclusive_scan<int64_t>(h->counts.p, warps, h->out_off.p, h->scan_tmp.p, st));
  // per-query offsets are every nsplit-th entry of the per-warp offsets
  CUDA_TRY(cudaMemcpy2DAsync(offsets, sizeof(int64_t), h->out_off.p, sizeof(int64_t) * sg.nsplit, sizeof(int64_t),
                             (size_t)nq, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(offsets + nq, h->out_off.p + warps, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(&total, h->out_off.p + warps, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (needed_h) *needed_h = total;
  if (total > idx_capacity) {
    set_error("radius query produced %lld hits, capacity %lld", (long long)total, (long long)idx_capacity);
    return POMP_ERR_CAPACITY;
  }
  if (total > 0) {
    LAUNCH((bf_radius_kernel<true>), blocks, 256, 0, st, h->d_pts, (long long)n, h->metric, skip, q, (long long)nq,
           radius, inclusive, sg, nullptr, h->out_off.p, idx);
    CHECK_LAUNCH();
  }
  return POMP_OK;
}

static int density_device(pomp_nn* h, const double* q, int64_t nq, double radius, double sigma, int inclusive,
                          double* dens, int32_t* cnt, cudaStream_t st) {
  if (nq == 0) return POMP_OK;
  POMP_TRY(nn_flush(h, st));
  bool use_grid = false;
  POMP_TRY(want_grid(h, nq, &use_grid, st));
  const uint8_t* skip = h->skip_ptr();
  const bool fastm = is_plain_euclid(h->metric) && h->grid.G == h->D;
  const int db = (int)((nq + 127) / 128);
#define DF(DD)                                                                                                   \
  LAUNCH((grid_density_fast_kernel<DD>), db, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q, (long long)nq, \
         radius, sigma, inclusive, dens, cnt)
  if (use_grid && fastm && h->D == 2) DF(2);
  else if (use_grid && fastm && h->D == 3) DF(3);
  else if (use_grid && fastm && h->D == 4) DF(4);
  else if (use_grid && fastm && h->D == 6) DF(6);
  else if (use_grid) {
    LAUNCH(grid_density_kernel, db, 128, 0, st, h->grid, h->metric, h->d_pts, (long long)h->n_dev,
           skip, q, (long long)nq, radius, sigma, inclusive, dens, cnt);
  } else {
    LAUNCH(bf_density_kernel, (int)((nq * 32 + 255) / 256), 256, 0, st, h->d_pts, (long long)h->n_dev, h->metric, skip, q,
           (long long)nq, radius, sigma, inclusive, dens, cnt);
  }
#undef DF
  CHECK_LAUNCH();
  return POMP_OK;
}

// ============================================================================ C ABI
POMP_API int pomp_nn_create(const pomp_metric_desc* metric, int64_t capacity_hint, int device, pomp_nn** out) {
  POMP_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  POMP_TRY(validate_metric(metric));
  DeviceGuard g(device);
  if (!g.ok) return POMP_ERR_CUDA;
  pomp_nn* h = new pomp_nn();
  h->device = device;
  h->metric = *metric;
  h->D = metric->dim;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaHostAlloc(&h->mailbox_h, sizeof(Mailbox), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->mailbox_d, h->mailbox_h, 0);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->ticket, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->ticket, 0, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaHostAlloc(&h->radbox_h, sizeof(RadBox), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&h->radbox_d, h->radbox_h, 0);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->rad_counters, 2 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->rad_counters, 0, 2 * sizeof(unsigned int));
  if (e != cudaSuccess) {
    set_error("pomp_nn_create: %s", cudaGetErrorString(e));
    delete h;
    return POMP_ERR_CUDA;
  }
  cudaDeviceSynchronize();  // default-stream memset above vs the handle's non-blocking stream
  int rc = nn_grow(h, std::max<int64_t>(capacity_hint, 1024), h->stream);
  if (rc != POMP_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return POMP_OK;
}

POMP_API int pomp_nn_destroy(pomp_nn* h) {
  if (!h) return POMP_OK;
  DeviceGuard g(h->device);
  if (h->d_pts) cudaFree(h->d_pts);
  if (h->d_skip) cudaFree(h->d_skip);
  if (h->mailbox_h) cudaFreeHost(h->mailbox_h);
  if (h->ticket) cudaFree(h->ticket);
  if (h->radbox_h) cudaFreeHost(h->radbox_h);
  if (h->rad_counters) cudaFree(h->rad_counters);
  if (h->stream) cudaStreamDestroy(h->stream);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  if (h->rw_pin) cudaFreeHost(h->rw_pin);
  for (cudaEvent_t e : h->pipe_ev) cudaEventDestroy(e);
  for (auto& cg : h->chunk_graphs)
    if (cg.exec) cudaGraphExecDestroy(cg.exec);
  for (cudaEvent_t e : h->ovl_ev) cudaEventDestroy(e);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->copy_in) cudaStreamDestroy(h->copy_in);
  if (h->copy_out) cudaStreamDestroy(h->copy_out);
  delete h;
  return POMP_OK;
}

POMP_API int pomp_nn_reset(pomp_nn* h) {
  POMP_REQUIRE(h, "handle is NULL");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (h->any_dead || h->has_reject) CUDA_TRY(cudaMemsetAsync(h->d_skip, 0, (size_t)h->n_dev, h->stream));
  h->n_dev = 0;
  h->pending.clear();
  h->h_dead.clear();
  h->h_reject.clear();
  h->any_dead = h->has_reject = false;
  h->grid_valid = false;
  h->grid_unusable = false;
  return POMP_OK;
}

POMP_API int64_t pomp_nn_size(const pomp_nn* h) { return h ? h->size() : 0; }

POMP_API int pomp_nn_add_h(pomp_nn* h, const double* pts_h, int64_t n, int64_t* first_index_out_h) {
  POMP_REQUIRE(h && (pts_h || n == 0) && n >= 0, "bad arguments to pomp_nn_add_h");
  if (first_index_out_h) *first_index_out_h = h->size();
  h->pending.insert(h->pending.end(), pts_h, pts_h + n * h->D);
  return POMP_OK;
}

POMP_API int pomp_nn_add(pomp_nn* h, const double* pts, int64_t n, int64_t* first_index_out_h, void* stream) {
  POMP_REQUIRE(h && (pts || n == 0) && n >= 0, "bad arguments to pomp_nn_add");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  POMP_TRY(nn_flush(h, st));
  if (first_index_out_h) *first_index_out_h = h->n_dev;
  POMP_TRY(nn_grow(h, h->n_dev + n, st));
  CUDA_TRY(cudaMemcpyAsync(h->d_pts + h->n_dev * h->D, pts, sizeof(double) * (size_t)n * h->D,
                           cudaMemcpyDeviceToDevice, st));
  h->n_dev += n;
  return POMP_OK;
}

POMP_API int pomp_nn_set(pomp_nn* h, const double* pts, int64_t n, void* stream) {
  POMP_TRY(pomp_nn_reset(h));
  POMP_TRY(pomp_nn_add(h, pts, n, nullptr, stream));
  return POMP_OK;
}

POMP_API int pomp_nn_set_h(pomp_nn* h, const double* pts_h, int64_t n) {
  POMP_TRY(pomp_nn_reset(h));
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  POMP_REQUIRE((pts_h || n == 0) && n >= 0, "bad arguments to pomp_nn_set_h");
  POMP_TRY(nn_grow(h, n, h->stream));
  if (n > 0) {
    CUDA_TRY(cudaMemcpyAsync(h->d_pts, pts_h, sizeof(double) * (size_t)n * h->D, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  h->n_dev = n;
  return POMP_OK;
}

static int upload_skip(pomp_nn* h) {
  int64_t n = h->size();
  POMP_TRY(nn_flush(h, h->stream));
  std::vector<uint8_t> comb((size_t)n, 0);
  for (int64_t i = 0; i < n; i++) {
    uint8_t v = 0;
    if ((size_t)i < h->h_dead.size()) v |= h->h_dead[i];
    if (h->has_reject && (size_t)i < h->h_reject.size()) v |= h->h_reject[i];
    comb[i] = v ? 1 : 0;
  }
  if (n > 0) {
    CUDA_TRY(cudaMemcpyAsync(h->d_skip, comb.data(), (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return POMP_OK;
}

POMP_API int pomp_nn_remove(pomp_nn* h, int64_t index) {
  POMP_REQUIRE(h, "handle is NULL");
  POMP_REQUIRE(index >= 0 && index < h->size(), "remove index %lld out of range", (long long)index);
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (h->h_dead.size() < (size_t)h->size()) h->h_dead.resize((size_t)h->size(), 0);
  h->h_dead[index] = 1;
  h->any_dead = true;
  POMP_TRY(nn_flush(h, h->stream));
  uint8_t one = 1;
  CUDA_TRY(cudaMemcpyAsync(h->d_skip + index, &one, 1, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return POMP_OK;
}

POMP_API int pomp_nn_set_mask_h(pomp_nn* h, const uint8_t* reject_mask_h, int64_t n) {
  POMP_REQUIRE(h, "handle is NULL");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (!reject_mask_h) {
    if (!h->has_reject) return POMP_OK;
    h->has_reject = false;
    h->h_reject.clear();
    return upload_skip(h);
  }
  POMP_REQUIRE(n == h->size(), "mask length %lld != point count %lld", (long long)n, (long long)h->size());
  h->h_reject.assign(reject_mask_h, reject_mask_h + n);
  h->has_reject = true;
  return upload_skip(h);
}

POMP_API int pomp_nn_update_metric(pomp_nn* h, const pomp_metric_desc* metric) {
  POMP_REQUIRE(h, "handle is NULL");
  POMP_TRY(validate_metric(metric));
  POMP_REQUIRE(metric->dim == h->D, "metric dim %d != point dim %d", metric->dim, h->D);
  // the grid layout depends on the metric's per-axis scales; cost bounds/weights do not touch it
  double s0[POMP_MAX_DIM], s1[POMP_MAX_DIM];
  axis_scales(h->metric, s0);
  axis_scales(*metric, s1);
  bool same = (metric->kind == h->metric.kind) && (metric->has_cost == h->metric.has_cost);
  for (int i = 0; i < base_dim(*metric) && same; i++) same = (s0[i] == s1[i]);
  h->metric = *metric;
  h->mut_count++;
  if (!same) {
    h->grid_valid = false;
    h->grid_unusable = false;
  }
  return POMP_OK;
}

POMP_API int pomp_nn_set_param(pomp_nn* h, const char* name, double value) {
  if (h) h->mut_count++;
  POMP_REQUIRE(h && name, "bad arguments to pomp_nn_set_param");
  std::string s(name);
  if (s == "grid_min_points") h->grid_min_points = value;
  else if (s == "points_per_cell") {
    h->points_per_cell = value;
    h->grid_valid = false;
  } else if (s == "sort_queries") h->sort_queries = value;
  else if (s == "grid_min_queries") h->grid_min_queries = value;
  else if (s == "max_tail") h->max_tail = value;
  else if (s == "block_search") h->block_search = value;
  else if (s == "block_need") h->block_need = value;
  else if (s == "single_radius") h->single_radius = value;
  else if (s == "warp_search") h->warp_search = value;
  else if (s == "warp_need") h->warp_need = value;
  else if (s == "warp_lanes") h->warp_lanes = value;
  else if (s == "strip_search") h->strip_search = value;
  else if (s == "strip_spin_ns") h->strip_spin_ns = value;
  else if (s == "strip_probe") h->strip_probe = value;
  else if (s == "strip_min_per_tile") h->strip_min_per_tile = value;
  else if (s == "se2_fast") h->se2_fast = value;
  else if (s == "bf_tiled") h->bf_tiled = value;
  else if (s == "use_graphs") h->use_graphs = value;
  else if (s == "pipeline_shape") h->pipeline_shape = value;
  else if (s == "periodic_axes") {
    h->periodic_axes = value;
    h->grid_valid = false;
  }

  else if (s == "sort_min_queries") h->sort_min_queries = value;
  else if (s == "sort_shift") h->sort_shift = value;
  else if (s == "pipeline_chunks") h->pipeline_chunks = value;
  else if (s == "overlap_chunks") h->overlap_chunks = value;
  else if (s == "overlap_min_queries") h->overlap_min_queries = value;
  else if (s == "profile") {
    h->profile = value != 0;
    h->prof_used = 0;
  }
  else {
    set_error("unknown parameter '%s'", name);
    return POMP_ERR_INVALID;
  }
  return POMP_OK;
}

POMP_API int pomp_nn_get_stat(pomp_nn* h, const char* name, double* out) {
  POMP_REQUIRE(h && name && out, "bad arguments to pomp_nn_get_stat");
  std::string s(name);
  *out = 0.0;
  if (s == "search_kernel_ms" || s == "search_kernel_launches") {
    // call after synchronising the stream the queries ran on
    double ms = 0.0;
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
      float t = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
      ms += t;
    }
    *out = (s == "search_kernel_ms") ? ms : (double)(h->prof_used / 2);
  } else if (s == "overflow_queries") {  // of the last large 2-D batch (first-stage deferrals)
    int c = 0;
    if (h->ovf_count_dev) CUDA_TRY(cudaMemcpy(&c, h->ovf_count_dev, sizeof(int), cudaMemcpyDeviceToHost));
    *out = (double)c;
  } else if (s == "strip_launches") *out = (double)h->strip_launches;
  else if (s == "strip_slots") *out = h->strip_build == h->grid_builds && h->grid_valid ? (double)h->strip.nslots : 0.0;
  else if (s == "n_cells") *out = h->grid_valid ? (double)h->grid.n_cells : 0.0;
  else if (s == "n_indexed") *out = h->grid_valid ? (double)h->grid.n_indexed : 0.0;
  else if (s == "grid_axes") *out = h->grid_valid ? (double)h->grid.G : 0.0;
  else {
    set_error("unknown stat '%s'", name);
    return POMP_ERR_INVALID;
  }
  return POMP_OK;
}

POMP_API int pomp_nn_build_index(pomp_nn* h, void* stream) {
  POMP_REQUIRE(h, "handle is NULL");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  return build_grid(h, (cudaStream_t)stream);
}

POMP_API int pomp_nn_knearest(pomp_nn* h, const double* q, int64_t nq, int32_t k, int64_t* idx, double* dist,
                              int32_t* count, void* stream) {
  POMP_REQUIRE(h && (nq == 0 || (q && idx && dist)) && nq >= 0, "bad arguments to pomp_nn_knearest");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  return knn_device(h, q, nq, k, idx, dist, count, (cudaStream_t)stream);
}

POMP_API int pomp_nn_nearest(pomp_nn* h, const double* q, int64_t nq, int64_t* idx, double* dist, void* stream) {
  return pomp_nn_knearest(h, q, nq, 1, idx, dist, nullptr, stream);
}

// everything a captured chunk search bakes into its kernel arguments
static uint64_t launch_signature(const pomp_nn* h) {
  uint64_t x = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { x = (x ^ v) * 1099511628211ull; };
  mix((uint64_t)h->n_dev);
  mix((uint64_t)(uintptr_t)h->d_pts);
  mix((uint64_t)(uintptr_t)h->skip_ptr());
  mix(h->grid_builds);
  mix(h->mut_count);
  // any scratch buffer that was freed / regrown since the capture (a batch of another size, a radius or density
  // call, another handle): the graph's baked-in pointers may dangle -- recapture
  mix(pomp::g_realloc_gen.load(std::memory_order_relaxed));
  return x | 1ull;
}

// One chunk of the pipelined *_h path.  The second time the same chunk shape comes by under the same
// launch signature its launches are captured into a CUDA graph; from then on the chunk is ONE graph
// launch (the first, direct run has sized every scratch buffer, so nothing allocates while capturing).
static int knn_chunk(pomp_nn* h, int c, const double* q, int64_t m, int k, int64_t* idx, double* dist, int32_t* cnt,
                     cudaStream_t st) {
  if (h->use_graphs == 0 || h->profile) return knn_device(h, q, m, k, idx, dist, cnt, st);
  if ((int)h->chunk_graphs.size() <= c) h->chunk_graphs.resize((size_t)c + 1);
  pomp_nn::ChunkGraph& G = h->chunk_graphs[c];
  POMP_TRY(nn_flush(h, st));
  bool use_grid = false;
  POMP_TRY(want_grid(h, m, &use_grid, st));  // builds / refreshes the index outside any capture
  const uint64_t sig = launch_signature(h);
  const bool same_shape = G.q == q && G.m == m && G.k == k && G.idx == idx && G.dist == dist && G.cnt == cnt;
  if (G.exec && same_shape && G.sig == sig) {
    CUDA_TRY(cudaGraphLaunch(G.exec, st));
    pomp::g_launches.fetch_add(G.launches, std::memory_order_relaxed);
    return POMP_OK;
  }
  if (G.exec) {
    cudaGraphExecDestroy(G.exec);
    G.exec = nullptr;
  }
  if (same_shape && G.seen_sig == sig) {
    const int64_t l0 = pomp::g_launches.load();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
    int rc = POMP_OK;
    if (e == cudaSuccess) {
      rc = knn_device(h, q, m, k, idx, dist, cnt, st);
      e = cudaStreamEndCapture(st, &graph);
    }
    if (e == cudaSuccess && rc == POMP_OK && graph) e = cudaGraphInstantiate(&G.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc != POMP_OK || !G.exec) {
      // capture refused: nothing ran; run the chunk directly and stop trying
      (void)cudaGetLastError();
      G.exec = nullptr;
      h->use_graphs = 0;
      return knn_device(h, q, m, k, idx, dist, cnt, st);
    }
    G.sig = sig;
    G.launches = pomp::g_launches.load() - l0;
    CUDA_TRY(cudaGraphLaunch(G.exec, st));
    return POMP_OK;
  }
  G.q = q;
  G.m = m;
  G.k = k;
  G.idx = idx;
  G.dist = dist;
  G.cnt = cnt;
  G.seen_sig = sig;
  return knn_device(h, q, m, k, idx, dist, cnt, st);
}

POMP_API int pomp_nn_knearest_h(pomp_nn* h, const double* q_h, int64_t nq, int32_t k, int64_t* idx_h,
                                double* dist_h, int32_t* count_h) {
  POMP_REQUIRE(h && (nq == 0 || (q_h && idx_h)) && nq >= 0, "bad arguments to pomp_nn_knearest_h");  // dist_h may be NULL
  POMP_REQUIRE(k >= 1 && k <= 128, "k=%d outside 1..128", k);
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (nq == 0) return POMP_OK;
  cudaStream_t st = h->stream;
  POMP_TRY(h->q_dev.reserve((size_t)nq * h->D));
  POMP_TRY(h->out_idx.reserve((size_t)nq * k));
  POMP_TRY(h->out_dist.reserve((size_t)nq * k));
  POMP_TRY(h->out_cnt.reserve((size_t)nq));
  // Large batches: software pipeline over chunks -- H2D of chunk c+1 (copy-in stream), search of
  // chunk c (compute stream, scratch reused in order) and D2H of chunk c-1 (copy-out stream) overlap,
  // so PCIe runs in both directions while the SMs work.  Small batches: one pass on one stream.
  const int64_t kMinPipelined = 131072;
  int nchunk = (nq >= kMinPipelined && h->pipeline_chunks > 1) ? (int)h->pipeline_chunks : 1;
  if (nchunk > 1) {
    POMP_TRY(nn_flush(h, st));
    bool use_grid = false;
    POMP_TRY(want_grid(h, nq, &use_grid, st));  // build / refresh the index once, before the pipeline
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!h->copy_in) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
    }
    while ((int)h->pipe_ev.size() < 2 * nchunk) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->pipe_ev.push_back(e);
    }
    // chunk sizes shrink towards the end (weights nchunk+1, nchunk, ..., 2): the search and copy-out of
    // the LAST chunk cannot hide behind anything, so it is the smallest (host-to-device is the slower
    // direction on this box: 36 vs 56 GB/s, profiles/pcie_probe.py)
    // pipeline_shape 1: triangle (small first chunk so that copy-out starts early, small last chunk)
    std::vector<double> wts((size_t)nchunk);
    double wtot = 0.0;
    for (int c = 0; c < nchunk; c++) {
      wts[c] = h->pipeline_shape != 0 ? (double)std::min(c + 1, nchunk - c) + 0.5 : (double)(nchunk + 1 - c);
      wtot += wts[c];
    }
    int64_t b = 0;
    // POMP_PIPE_TRACE=1: per-chunk completion times of the three stages on stderr (debug aid)
    static const bool trace = getenv("POMP_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) {
      tev.resize(1 + 3 * (size_t)nchunk);
      for (auto& e : tev) CUDA_TRY(cudaEventCreate(&e));
      CUDA_TRY(cudaEventRecord(tev[0], st));
      CUDA_TRY(cudaStreamWaitEvent(h->copy_in, tev[0], 0));
    }
    int nc_done = 0;
    for (int c = 0; c < nchunk && b < nq; c++) {
      int64_t m = ((int64_t)((double)nq * wts[c] / wtot) + 127) / 128 * 128;
      if (c == nchunk - 1 || b + m > nq) m = nq - b;
      const int64_t e = b + m;
      CUDA_TRY(cudaMemcpyAsync(h->q_dev.p + b * h->D, q_h + b * h->D, sizeof(double) * (size_t)m * h->D,
                               cudaMemcpyHostToDevice, h->copy_in));
      CUDA_TRY(cudaEventRecord(h->pipe_ev[2 * c], h->copy_in));
      if (trace) CUDA_TRY(cudaEventRecord(tev[1 + 3 * c], h->copy_in));
      CUDA_TRY(cudaStreamWaitEvent(st, h->pipe_ev[2 * c], 0));
      h->pipe_chunks_active = nchunk;  // strip_eligible: a chunk is a fraction of the batch (gather kernels below 64 / tile)
      const int chunk_status = knn_chunk(h, c, h->q_dev.p + b * h->D, m, k, h->out_idx.p + b * k, h->out_dist.p + b * k,
                                         count_h ? h->out_cnt.p + b : nullptr, st);
      h->pipe_chunks_active = 0;
      POMP_TRY(chunk_status);
      CUDA_TRY(cudaEventRecord(h->pipe_ev[2 * c + 1], st));
      if (trace) CUDA_TRY(cudaEventRecord(tev[2 + 3 * c], st));
      CUDA_TRY(cudaStreamWaitEvent(h->copy_out, h->pipe_ev[2 * c + 1], 0));
      CUDA_TRY(cudaMemcpyAsync(idx_h + b * k, h->out_idx.p + b * k, sizeof(int64_t) * (size_t)m * k,
                               cudaMemcpyDeviceToHost, h->copy_out));
      if (dist_h)
        CUDA_TRY(cudaMemcpyAsync(dist_h + b * k, h->out_dist.p + b * k, sizeof(double) * (size_t)m * k,
                                 cudaMemcpyDeviceToHost, h->copy_out));
      if (count_h)
        CUDA_TRY(cudaMemcpyAsync(count_h + b, h->out_cnt.p + b, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost,
                                 h->copy_out));
      if (trace) CUDA_TRY(cudaEventRecord(tev[3 + 3 * c], h->copy_out));
      nc_done = c + 1;
      b = e;
    }
    CUDA_TRY(cudaStreamSynchronize(h->copy_out));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (trace) {
      for (int c = 0; c < nc_done; c++) {
        float t1 = 0, t2 = 0, t3 = 0;
        cudaEventElapsedTime(&t1, tev[0], tev[1 + 3 * c]);
        cudaEventElapsedTime(&t2, tev[0], tev[2 + 3 * c]);
        cudaEventElapsedTime(&t3, tev[0], tev[3 + 3 * c]);
        fprintf(stderr, "pomp pipe chunk %d: h2d done %.3f  search done %.3f  d2h done %.3f ms\n", c, t1, t2, t3);
      }
      for (auto& e : tev) cudaEventDestroy(e);
    }
    return POMP_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(h->q_dev.p, q_h, sizeof(double) * (size_t)nq * h->D, cudaMemcpyHostToDevice, st));
  POMP_TRY(knn_device(h, h->q_dev.p, nq, k, h->out_idx.p, h->out_dist.p, count_h ? h->out_cnt.p : nullptr, st));
  CUDA_TRY(cudaMemcpyAsync(idx_h, h->out_idx.p, sizeof(int64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
  if (dist_h)
    CUDA_TRY(cudaMemcpyAsync(dist_h, h->out_dist.p, sizeof(double) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
  if (count_h) CUDA_TRY(cudaMemcpyAsync(count_h, h->out_cnt.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return POMP_OK;
}

POMP_API int pomp_nn_nearest_h(pomp_nn* h, const double* q_h, int64_t nq, int64_t* idx_h, double* dist_h) {
  POMP_REQUIRE(h && (nq == 0 || (q_h && idx_h)) && nq >= 0, "bad arguments to pomp_nn_nearest_h");  // dist_h may be NULL
  if (nq != 1) return pomp_nn_knearest_h(h, q_h, nq, 1, idx_h, dist_h, nullptr);
  // planner fast path: query travels in the kernel arguments, the answer lands in mapped host memory
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = h->stream;
  POMP_TRY(nn_flush(h, st));
  const int64_t n = h->n_dev;
  if (n == 0) {
    *idx_h = -1;
    if (dist_h) *dist_h = INFINITY;
    return POMP_OK;
  }
  QueryArg qa;
  for (int j = 0; j < POMP_MAX_DIM; j++) qa.v[j] = j < h->D ? q_h[j] : 0.0;
  int blocks = (int)std::min<int64_t>((n + 1023) / 1024, (int64_t)sm_count(h->device) * 2);
  blocks = std::max(blocks, 1);
  int threads = n <= 256 ? 64 : 256;
  POMP_TRY(h->part_d.reserve((size_t)blocks));
  POMP_TRY(h->part_i.reserve((size_t)blocks));
  LAUNCH(bf_single_nearest_kernel, blocks, threads, 0, st, h->d_pts, (long long)n, h->metric, h->skip_ptr(), qa,
         h->part_d.p, h->part_i.p, h->ticket, (Mailbox*)h->mailbox_d);
  CHECK_LAUNCH();
  CUDA_TRY(cudaStreamSynchronize(st));
  const Mailbox* mb = (const Mailbox*)h->mailbox_h;
  *idx_h = mb->idx;
  if (dist_h) *dist_h = mb->dist;
  return POMP_OK;
}

POMP_API int pomp_nn_neighbors(pomp_nn* h, const double* q, int64_t nq, double radius, int inclusive,
                               int64_t* offsets, int64_t* idx, int64_t idx_capacity, int64_t* needed_h,
                               void* stream) {
  POMP_REQUIRE(h && offsets && (nq == 0 || q) && nq >= 0 && idx_capacity >= 0, "bad arguments to pomp_nn_neighbors");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  return radius_device(h, q, nq, radius, inclusive, offsets, idx, idx_capacity, needed_h, (cudaStream_t)stream);
}

POMP_API int pomp_nn_density(pomp_nn* h, const double* q, int64_t nq, double radius, double sigma, int inclusive,
                             double* density, int32_t* count, void* stream) {
  POMP_REQUIRE(h && (nq == 0 || (q && density)) && nq >= 0, "bad arguments to pomp_nn_density");
  POMP_REQUIRE(radius >= 0.0 && sigma > 0.0, "density needs radius >= 0 and sigma > 0");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  return density_device(h, q, nq, radius, sigma, inclusive, density, count, (cudaStream_t)stream);
}

POMP_API int pomp_nn_density_h(pomp_nn* h, const double* q_h, int64_t nq, double radius, double sigma, int inclusive,
                               double* density_h, int32_t* count_h) {
  POMP_REQUIRE(h && (nq == 0 || (q_h && density_h)) && nq >= 0, "bad arguments to pomp_nn_density_h");
  POMP_REQUIRE(radius >= 0.0 && sigma > 0.0, "density needs radius >= 0 and sigma > 0");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (nq == 0) return POMP_OK;
  cudaStream_t st = h->stream;
  POMP_TRY(h->q_dev.reserve((size_t)nq * h->D));
  POMP_TRY(h->out_dist.reserve((size_t)nq));
  POMP_TRY(h->out_cnt.reserve((size_t)nq));
  CUDA_TRY(cudaMemcpyAsync(h->q_dev.p, q_h, sizeof(double) * (size_t)nq * h->D, cudaMemcpyHostToDevice, st));
  POMP_TRY(density_device(h, h->q_dev.p, nq, radius, sigma, inclusive, h->out_dist.p, h->out_cnt.p, st));
  CUDA_TRY(cudaMemcpyAsync(density_h, h->out_dist.p, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, st));
  if (count_h)
    CUDA_TRY(cudaMemcpyAsync(count_h, h->out_cnt.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return POMP_OK;
}

POMP_API int pomp_nn_neighbors_h(pomp_nn* h, const double* q_h, int64_t nq, double radius, int inclusive,
                                 int64_t* offsets_h, int64_t* idx_h, int64_t idx_capacity, int64_t* needed_h) {
  POMP_REQUIRE(h && offsets_h && (nq == 0 || q_h) && nq >= 0 && idx_capacity >= 0,
               "bad arguments to pomp_nn_neighbors_h");
  DeviceGuard g(h->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = h->stream;
  if (nq == 0) {
    offsets_h[0] = 0;
    if (needed_h) *needed_h = 0;
    return POMP_OK;
  }
  if (nq == 1 && h->single_radius != 0) {
    // planner fast path (node sets of the brute-force regime): see bf_single_radius_kernel
    POMP_TRY(nn_flush(h, st));
    const int64_t n = h->n_dev;
    bool use_grid = false;
    POMP_TRY(want_grid(h, nq, &use_grid, st));
    if (!use_grid) {
      if (n == 0) {
        offsets_h[0] = offsets_h[1] = 0;
        if (needed_h) *needed_h = 0;
        return POMP_OK;
      }
      QueryArg qa;
      for (int j = 0; j < POMP_MAX_DIM; j++) qa.v[j] = j < h->D ? q_h[j] : 0.0;
      int blocks = (int)std::min<int64_t>((n + 1023) / 1024, (int64_t)sm_count(h->device) * 2);
      blocks = std::max(blocks, 1);
      const int threads = n <= 256 ? 64 : 256;
      LAUNCH(bf_single_radius_kernel, blocks, threads, 0, st, h->d_pts, (long long)n, h->metric, h->skip_ptr(), qa,
             radius, inclusive, h->rad_counters, (RadBox*)h->radbox_d);
      CHECK_LAUNCH();
      CUDA_TRY(cudaStreamSynchronize(st));
      RadBox* rb = (RadBox*)h->radbox_h;
      const int64_t cnt = rb->count;
      if (cnt <= RADBOX_CAP) {
        if (needed_h) *needed_h = cnt;
        offsets_h[0] = 0;
        offsets_h[1] = cnt;
        if (cnt > idx_capacity) {
          set_error("radius query produced %lld hits, capacity %lld", (long long)cnt, (long long)idx_capacity);
          return POMP_ERR_CAPACITY;
        }
        std::sort(rb->hits, rb->hits + cnt);   // neighbors() reports hits in insertion order
        for (int64_t j = 0; j < cnt; j++) idx_h[j] = rb->hits[j];
        return POMP_OK;
      }
      // more hits than the box holds: the general path answers
    }
  }
  POMP_TRY(h->q_dev.reserve((size_t)nq * h->D));
  POMP_TRY(h->out_idx.reserve((size_t)std::max<int64_t>(idx_capacity, 1)));
  DevBuf<int64_t>& off = h->rad_off;   // kept in the handle: no allocation per call
  POMP_TRY(off.reserve((size_t)nq + 1));
  CUDA_TRY(cudaMemcpyAsync(h->q_dev.p, q_h, sizeof(double) * (size_t)nq * h->D, cudaMemcpyHostToDevice, st));
  int64_t needed = 0;
  int rc = radius_device(h, h->q_dev.p, nq, radius, inclusive, off.p, h->out_idx.p, idx_capacity, &needed, st);
  if (needed_h) *needed_h = needed;
  if (rc != POMP_OK && rc != POMP_ERR_CAPACITY) return rc;
  CUDA_TRY(cudaMemcpyAsync(offsets_h, off.p, sizeof(int64_t) * (size_t)(nq + 1), cudaMemcpyDeviceToHost, st));
  if (rc == POMP_OK && needed > 0)
    CUDA_TRY(cudaMemcpyAsync(idx_h, h->out_idx.p, sizeof(int64_t) * (size_t)needed, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return rc;
}

POMP_API int pomp_nn_merge_candidates(const int64_t* idx_in, const double* dist_in, int32_t n_shards, int64_t nq,
                                      int32_t k, int64_t* idx_out, double* dist_out, int32_t* count_out,
                                      void* stream) {
  POMP_REQUIRE(idx_in && dist_in && idx_out && dist_out && n_shards >= 1 && nq >= 0, "bad arguments to merge");
  POMP_REQUIRE(k >= 1 && k <= 128, "k=%d outside 1..128", k);
  if (nq == 0) return POMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((nq * 32 + 255) / 256);
  // layout [shard][query][k]
#define MK(KL)                                                                                              \
  LAUNCH((merge_lists_kernel<KL, int64_t>), blocks, 256, 0, st, dist_in, idx_in, n_shards, (long long)nq, k, \
         (long long)k, (long long)nq * k, idx_out, dist_out, count_out)
  if (k <= 32) MK(1);
  else if (k <= 64) MK(2);
  else MK(4);
#undef MK
  CHECK_LAUNCH();
  return POMP_OK;
}
