// grid_fast.cuh -- register-resident grid search for the plain Euclidean metric with every
// coordinate gridded (the 2-D..6-D point-cloud configurations of BASELINE.json).
//
// Same algorithm and the same exactness argument as grid.cuh, specialised so that nothing is
// indexed dynamically (axes are unrolled by template recursion, the k-list lives in registers):
//   * candidates carry their SORTED POSITION; the insertion index is fetched from sidx[] only when
//     two distances are exactly equal (index tie-break) and once per result at the end;
//   * pruning is done on squared quantities against tsq = fl(T*T)*(1+2^-50): s > tsq implies
//     fl(sqrt(s)) > T, so no sqrt is spent on rows, cells or far candidates; everything that
//     survives is compared on the rooted value exactly as the reference does (vectorops.py:101).
#pragma once
#include "grid.cuh"
#include <type_traits>

#include "topk.cuh"

namespace pomp {

struct IdxResolver {
  const int* __restrict__ sidx;
  int n_indexed;
  // sorted slot -> insertion index; tail points (never sorted) use their insertion index as slot
  __device__ __forceinline__ int operator()(int pos) const { return pos < n_indexed ? __ldg(sidx + pos) : pos; }
};

__device__ __forceinline__ double tsq_of(double T) { return (T * T) * (1.0 + 0x1p-50); }

// Sorted list of the K best (distance, position); ties resolved lazily through the resolver.
// The list always keeps K entries (a runtime k <= K is served by its first k entries: the k best
// are a prefix of the K best).  No runtime slot arithmetic: with it, LLVM folds the unrolled
// selects back into a dynamically indexed local array and the list leaves the register file.
template <int K, class Res = IdxResolver>
struct PosTopK {
  static_assert(K <= 32, "register-resident list: larger k goes through the generic kernel");
  double d[K];
  int p[K];
  double T, tsq;  // current K-th distance and its squared prune bound
  Res res;

  __device__ __forceinline__ void init(Res r) {
    res = r;
#pragma unroll
    for (int j = 0; j < K; j++) {
      d[j] = POMP_INF;
      p[j] = -1;
    }
    T = POMP_INF;
    tsq = POMP_INF;
  }
  __device__ __forceinline__ double threshold() const { return T; }
  __device__ __forceinline__ double tsq_up() const { return tsq; }
  __device__ __forceinline__ bool full() const { return T < POMP_INF; }
  // (cd,cp) < slot j ?
  __device__ __forceinline__ bool less_slot(double cd, int cp, int j) const {
    if (cd < d[j]) return true;
    if (cd == d[j] && p[j] >= 0) return res(cp) < res(p[j]);
    return false;
  }
  // s = squared distance
  __device__ __forceinline__ void offer_sq(double s, int pos) {
    if (!(s <= tsq)) return;  // also rejects NaN
    double cd = sqrt(s);
    if (!(cd < POMP_INF)) return;
    if (!less_slot(cd, pos, K - 1)) return;  // must beat the current worst kept entry
#pragma unroll
    for (int j = K - 1; j >= 1; j--) {
      bool lt_prev = less_slot(cd, pos, j - 1);
      if (lt_prev) {
        d[j] = d[j - 1];
        p[j] = p[j - 1];
      } else if (less_slot(cd, pos, j)) {
        d[j] = cd;
        p[j] = pos;
      }
    }
    if (less_slot(cd, pos, 0)) {
      d[0] = cd;
      p[0] = pos;
    }
    T = d[K - 1];
    tsq = tsq_of(T);
  }
  __device__ __forceinline__ void finish(int (&idx)[K]) {
#pragma unroll
    for (int j = 0; j < K; j++) idx[j] = p[j] >= 0 ? res(p[j]) : IDX_NONE;
  }
};

// Larger k: keep the K best UNSORTED and track the worst kept entry (what KNearestResult does,
// knn.py:14-23): an insertion is "overwrite the worst slot, rescan for the new worst" (~3 ops per
// slot) instead of a shifting sorted insert (~12 per slot); one sort at the very end.
template <int K, class Res = IdxResolver>
struct PosMaxK {
  static_assert(K <= 32, "register-resident list");
  double d[K];
  int p[K];
  double T, tsq;  // worst kept distance (+inf while slots are free) and its squared prune bound
  int wp;         // position of the worst kept entry (-1: a free slot)
  int wslot;
  Res res;

  __device__ __forceinline__ void init(Res r) {
    res = r;
#pragma unroll
    for (int j = 0; j < K; j++) {
      d[j] = POMP_INF;
      p[j] = -1;
    }
    T = POMP_INF;
    tsq = POMP_INF;
    wp = -1;
    wslot = 0;
  }
  __device__ __forceinline__ double threshold() const { return T; }
  __device__ __forceinline__ double tsq_up() const { return tsq; }
  __device__ __forceinline__ bool full() const { return T < POMP_INF; }
  __device__ __forceinline__ void offer_sq(double s, int pos) {
    if (!(s <= tsq)) return;
    double cd = sqrt(s);
    if (!(cd < POMP_INF)) return;
    // must beat the worst kept entry on (distance, insertion index)
    if (!(cd < T)) {
      if (!(cd == T && wp >= 0 && res(pos) < res(wp))) return;
    }
#pragma unroll
    for (int j = 0; j < K; j++)
      if (j == wslot) {
        d[j] = cd;
        p[j] = pos;
      }
    // new worst
    double wd = d[0];
    int wq = p[0], ws = 0;
#pragma unroll
    for (int j = 1; j < K; j++) {
      bool worse = d[j] > wd;
      if (d[j] == wd && d[j] < POMP_INF) worse = res(p[j]) > res(wq);
      if (worse) {
        wd = d[j];
        wq = p[j];
        ws = j;
      }
    }
    T = wd;
    wp = wq;
    wslot = ws;
    tsq = tsq_of(T);
  }
  // sort by (distance, insertion index); idx[] receives the resolved indices
  __device__ __forceinline__ void finish(int (&idx)[K]) {
#pragma unroll
    for (int j = 0; j < K; j++) idx[j] = p[j] >= 0 ? res(p[j]) : IDX_NONE;
#pragma unroll
    for (int a = 0; a < K; a++) {
#pragma unroll
      for (int b = (a & 1); b + 1 < K; b += 2) {  // odd-even transposition sort: K passes, branch-free
        bool sw = d[b] > d[b + 1] || (d[b] == d[b + 1] && idx[b] > idx[b + 1]);
        double td = sw ? d[b + 1] : d[b], ud = sw ? d[b] : d[b + 1];
        int ti = sw ? idx[b + 1] : idx[b], ui = sw ? idx[b] : idx[b + 1];
        d[b] = td;
        d[b + 1] = ud;
        idx[b] = ti;
        idx[b + 1] = ui;
      }
    }
  }
};

// k >= 8: the list lives in SHARED memory (slot-major, one column per thread: conflict-free for
// equal slots), kept sorted by plain insertion -- dynamic indexing is free there, an insertion is a
// short shift loop instead of 2K unrolled selects, and the kernel needs ~40 registers instead of
// ~140 (3x the occupancy).
template <int K, class Res = IdxResolver>
struct SmemTopK {
  double* sd;  // sd[j * stride]
  int* sp;
  int stride, filled;
  double T, tsq;
  Res res;
  __device__ __forceinline__ void init(Res r, unsigned char* smem, int nthreads, int tid) {
    res = r;
    sd = reinterpret_cast<double*>(smem) + tid;
    sp = reinterpret_cast<int*>(smem + (size_t)K * nthreads * sizeof(double)) + tid;
    stride = nthreads;
    filled = 0;
    T = POMP_INF;
    tsq = POMP_INF;
  }
  __device__ __forceinline__ double threshold() const { return T; }
  __device__ __forceinline__ double tsq_up() const { return tsq; }
  __device__ __forceinline__ bool full() const { return filled == K; }
  __device__ __forceinline__ bool less_slot(double cd, int cp, int j) const {
    double dj = sd[j * stride];
    if (cd < dj) return true;
    if (cd == dj) return res(cp) < res(sp[j * stride]);
    return false;
  }
  __device__ __forceinline__ void offer_sq(double s, int pos) {
    if (!(s <= tsq)) return;
    double cd = sqrt(s);
    if (!(cd < POMP_INF)) return;
    int j;
    if (filled < K) {
      j = filled++;
    } else {
      if (!less_slot(cd, pos, K - 1)) return;
      j = K - 1;
    }
    while (j > 0 && less_slot(cd, pos, j - 1)) {
      sd[j * stride] = sd[(j - 1) * stride];
      sp[j * stride] = sp[(j - 1) * stride];
      j--;
    }
    sd[j * stride] = cd;
    sp[j * stride] = pos;
    if (filled == K) {
      T = sd[(K - 1) * stride];
      tsq = tsq_of(T);
    }
  }
  __device__ __forceinline__ double dist_at(int j) const { return j < filled ? sd[j * stride] : POMP_INF; }
  __device__ __forceinline__ int index_at(int j) const { return res(sp[j * stride]); }
};

template <int K, class Res = IdxResolver>
struct BestList {  // sorted insertion for tiny k, worst-tracking for the rest
  using type = typename std::conditional<(K <= 4), PosTopK<K, Res>, PosMaxK<K, Res>>::type;
};

// fixed-radius collector (count or fill)
template <bool FILL>
struct RadiusFast {
  double radius, tsq;
  int inclusive;
  long long cnt;
  IdxResolver res;
  int64_t* out;
  __device__ __forceinline__ double threshold() const { return radius; }
  __device__ __forceinline__ double tsq_up() const { return tsq; }
  __device__ __forceinline__ void offer_sq(double s, int pos) {
    if (!(s <= tsq)) return;
    double d = sqrt(s);
    bool hit = inclusive ? (d <= radius) : (d < radius);
    if (hit) {
      if (FILL) out[cnt] = res(pos);
      cnt++;
    }
  }
};

template <int D>
__device__ __forceinline__ double sqdist(const double (&p)[D], const double (&q)[D]) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    double t = p[i] - q[i];
    s = s + t * t;
  }
  return s;
}

template <class C>
struct small_list : std::false_type {};
template <int K, class Res>
struct small_list<PosTopK<K, Res>> : std::integral_constant<bool, (K < 8)> {};

template <int D, class C>
__device__ __forceinline__ void scan_run(const GridDev& g, const uint8_t* __restrict__ skip, const double (&q)[D],
                                         C& c, int s, int t) {
  if (s >= t) return;
  if constexpr (small_list<C>::value && D >= 6) {
    // tiny list, wide points: the extra D registers of the look-ahead cost more than the latency they hide
    // (6-D k = 1: 9.2 ms without, 10.1 ms with)
    for (int pos = s; pos < t; pos++) {
      if (skip && skip[__ldg(g.sidx + pos)]) continue;
      double p[D];
      load_point<D>(g.spts, pos, p);
      c.offer_sq(sqdist<D>(p, q), pos);
    }
    return;
  }
  // the point of position pos+1 is fetched while position pos is offered: with a large list update in the body
  // the compiler does not pipeline the loop, and every candidate would expose its load latency
  double pn[D];
  load_point<D>(g.spts, s, pn);
  for (int pos = s; pos < t; pos++) {
    double p[D];
#pragma unroll
    for (int j = 0; j < D; j++) p[j] = pn[j];
    if (pos + 1 < t) load_point<D>(g.spts, pos + 1, pn);
    if (skip && skip[__ldg(g.sidx + pos)]) continue;
    c.offer_sq(sqdist<D>(p, q), pos);
  }
}

// squared lower bound with gaps on axes >= A (compile-time), summed in coordinate order
template <int D, int A>
__device__ __forceinline__ double gap_sq(const double (&gap)[D]) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    double gi = i >= A ? gap[i] : 0.0;
    s = s + gi * gi;
  }
  return s;
}

// CO (centre-out: 0, -1, +1, -2, ...): near cells shrink the threshold before the far ones are looked
// at, and a direction closes at the first cell whose gap bound exceeds it (gaps grow with the
// offset, the threshold only shrinks).  Pays when the first bound is loose -- k >= 8 in 4-D / 6-D:
// 73 -> 49 ms for k = 16 in 6-D; for k = 1 the plain ascending sweep over the range is cheaper.
template <int D, int A, bool CO = false>
struct FastLevel {
  template <class C>
  static __device__ __forceinline__ void run(const GridDev& g, const uint8_t* __restrict__ skip,
                                             const double (&q)[D], double (&gap)[D], long long base, C& c, int w0,
                                             int w1) {
    if constexpr (CO) {
      const int home = grid_cell(g, A, q[A]);
      const int nc = g.ncell[A];
      bool dlo = false, dhi = false;
      for (int t = 0; !(dlo && dhi); t++) {
        const int off = t == 0 ? 0 : ((t & 1) ? -((t + 1) >> 1) : (t >> 1));
        if (off < 0 ? dlo : dhi) continue;
        const int cc = home + off;
        if (cc < 0) {
          dlo = true;
          continue;
        }
        if (cc >= nc) {
          dhi = true;
          continue;
        }
        gap[A] = grid_gap(g, A, cc, q[A]);
        if (gap_sq<D, A>(gap) > c.tsq_up()) {
          if (off < 0) dlo = true;
          else dhi = true;
          continue;
        }
        FastLevel<D, A - 1, CO>::run(g, skip, q, gap, base + (long long)cc * g.stride[A], c, w0, w1);
      }
    } else {
      int lo, hi;
      grid_range(g, A, q[A], c.threshold(), &lo, &hi);
      for (int cc = lo; cc <= hi; cc++) {
        gap[A] = grid_gap(g, A, cc, q[A]);
        if (gap_sq<D, A>(gap) > c.tsq_up()) continue;
        FastLevel<D, A - 1, CO>::run(g, skip, q, gap, base + (long long)cc * g.stride[A], c, w0, w1);
      }
    }
  }
};

template <int D, bool CO>
struct FastLevel<D, 0, CO> {
  template <class C>
  static __device__ __forceinline__ void run(const GridDev& g, const uint8_t* __restrict__ skip,
                                             const double (&q)[D], double (&gap)[D], long long base, C& c, int w0,
                                             int w1) {
    int lo, hi;
    grid_range(g, 0, q[0], c.threshold(), &lo, &hi);
    while (lo < hi) {
      gap[0] = grid_gap(g, 0, lo, q[0]);
      if (gap_sq<D, 0>(gap) > c.tsq_up()) lo++;
      else break;
    }
    while (hi > lo) {
      gap[0] = grid_gap(g, 0, hi, q[0]);
      if (gap_sq<D, 0>(gap) > c.tsq_up()) hi--;
      else break;
    }
    int s = __ldg(g.cell_start + base + lo);
    int t = __ldg(g.cell_start + base + hi + 1);
    // positions [w0,w1) were offered in phase A
    if (s < w1 && t > w0) {
      scan_run<D>(g, skip, q, c, s, min(t, w0));
      scan_run<D>(g, skip, q, c, max(s, w1), t);
    } else {
      scan_run<D>(g, skip, q, c, s, t);
    }
  }
};

// points appended after the last index build: their "position" is their insertion index
template <int D, class C>
__device__ __forceinline__ void scan_tail(const GridDev& g, const double* __restrict__ pts, long long n,
                                          const uint8_t* __restrict__ skip, const double (&q)[D], C& c) {
  for (long long i = g.n_indexed; i < n; i++) {
    if (skip && skip[i]) continue;
    double p[D];
    load_point<D>(pts, i, p);
    c.offer_sq(sqdist<D>(p, q), (int)i);
  }
}

// one query through the pruned traversal (phase A + FastLevel recursion + tail)
template <int D, int K>
__device__ __forceinline__ void knn_pruned_one(const GridDev& g, const double* __restrict__ pts, long long n,
                                               const uint8_t* __restrict__ skip, const double* __restrict__ q,
                                               long long qi, int k, int64_t* __restrict__ out_idx,
                                               double* __restrict__ out_dist, int32_t* __restrict__ out_cnt) {
  double qq[D];
  load_point<D>(q, qi, qq);
  PosTopK<K> top;
  top.init(IdxResolver{g.sidx, g.n_indexed});
  // phase A: the query's home cell (its points are close in every coordinate), widened to at
  // least K sorted-order neighbours; grown further only while masked points leave the list short
  long long home = 0;
#pragma unroll
  for (int a = 0; a < D; a++) home += (long long)grid_cell(g, a, qq[a]) * g.stride[a];
  int hs = __ldg(g.cell_start + home), he = __ldg(g.cell_start + home + 1);
  int center = (hs + he) >> 1;
  int w0 = min(hs, max(0, center - K)), w1 = max(he, min(g.n_indexed, center + K));
  scan_run<D>(g, skip, qq, top, w0, w1);
  while (!top.full() && (w0 > 0 || w1 < g.n_indexed)) {
    int len = max(w1 - w0, 1);
    if (w0 > 0) {
      int nw0 = max(0, w0 - len);
      scan_run<D>(g, skip, qq, top, nw0, w0);
      w0 = nw0;
    } else {
      int nw1 = min(g.n_indexed, w1 + len);
      scan_run<D>(g, skip, qq, top, w1, nw1);
      w1 = nw1;
    }
  }
  double gap[D];
#pragma unroll
  for (int i = 0; i < D; i++) gap[i] = 0.0;
  FastLevel<D, D - 1, (K >= 8)>::run(g, skip, qq, gap, 0LL, top, w0, w1);
  scan_tail<D>(g, pts, n, skip, qq, top);
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < K; j++)
    if (j < k) {
      bool fin = top.d[j] < POMP_INF;
      out_idx[qi * k + j] = fin ? (int64_t)top.res(top.p[j]) : -1;
      out_dist[qi * k + j] = top.d[j];
      cnt += fin;
    }
  if (out_cnt) out_cnt[qi] = cnt;
}

// qlist == nullptr: thread t answers query qperm[t] (or t).  qlist != nullptr: grid-stride over the
// *qcount queries listed there (the block kernel's overflow list).
template <int D, int K>
__global__ void __launch_bounds__(128)
grid_knn_fast_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                     const double* __restrict__ q, long long nq, const int* __restrict__ qperm, int k,
                     int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                     const int* __restrict__ qlist, const int* __restrict__ qcount) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qlist) {
    long long cnt = *qcount;
    for (; t < cnt; t += (long long)gridDim.x * blockDim.x)
      knn_pruned_one<D, K>(g, pts, n, skip, q, (long long)qlist[t], k, out_idx, out_dist, out_cnt);
    return;
  }
  if (t >= nq) return;
  long long qi = qperm ? (long long)__ldg(qperm + t) : t;
  knn_pruned_one<D, K>(g, pts, n, skip, q, qi, k, out_idx, out_dist, out_cnt);
}

// Fixed-block search for 2-D / 3-D: scan the (2R+1)^D block of cells around the home cell (each
// row of the block is one contiguous run of points), then CERTIFY: if the K-th distance is smaller
// than the distance from the query to the outside of the block, nothing outside can matter and the
// answer is exact; otherwise the query goes to the overflow list and is redone by the pruned
// traversal.  No per-row pruning logic => short, almost divergence-free code for the common case.
template <int D, int K>
__global__ void __launch_bounds__(128)
grid_knn_block_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                      const double* __restrict__ q, long long nq, const int* __restrict__ qperm, int k, int R,
                      int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                      int* __restrict__ ovf_list, int* __restrict__ ovf_count) {
  static_assert(D == 2 || D == 3, "block kernel is for 2-D and 3-D");
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  long long qi = qperm ? (long long)__ldg(qperm + t) : t;
  double qq[D];
  load_point<D>(q, qi, qq);
  int c[D];
#pragma unroll
  for (int a = 0; a < D; a++) c[a] = grid_cell(g, a, qq[a]);
  PosTopK<K> top;
  top.init(IdxResolver{g.sidx, g.n_indexed});
  const int w = 2 * R + 1;
  const int nrow = (D == 2) ? w : w * w;
  const int x0 = max(c[0] - R, 0), x1 = min(c[0] + R, g.ncell[0] - 1);
  for (int r = 0; r < nrow; r++) {
    int cy = c[1] + (D == 2 ? r : r % w) - R;
    if (cy < 0 || cy >= g.ncell[1]) continue;
    long long base = (long long)cy * g.stride[1];
    if (D == 3) {
      int cz = c[2] + r / w - R;
      if (cz < 0 || cz >= g.ncell[2]) continue;
      base += (long long)cz * g.stride[2];
    }
    int s = __ldg(g.cell_start + base + x0), e = __ldg(g.cell_start + base + x1 + 1);
    scan_run<D>(g, skip, qq, top, s, e);
  }
  scan_tail<D>(g, pts, n, skip, qq, top);
  // certificate: distance from the query to anything bucketed outside the block
  double margin = POMP_INF;
#pragma unroll
  for (int a = 0; a < D; a++) {
    int lo_c = c[a] - R, hi_c = c[a] + R;
    if (lo_c > 0) margin = fmin(margin, qq[a] - ((g.lo[a] + (double)lo_c * g.h[a]) + g.slack[a]));
    if (hi_c < g.ncell[a] - 1) margin = fmin(margin, ((g.lo[a] + (double)(hi_c + 1) * g.h[a]) - g.slack[a]) - qq[a]);
  }
  bool exact = !(margin < POMP_INF) || (top.T * (1.0 + 1e-9) < margin);
  if (!exact) {
    int o = atomicAdd(ovf_count, 1);
    ovf_list[o] = (int)qi;
    return;
  }
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < K; j++)
    if (j < k) {
      bool fin = top.d[j] < POMP_INF;
      out_idx[qi * k + j] = fin ? (int64_t)top.res(top.p[j]) : -1;
      out_dist[qi * k + j] = top.d[j];
      cnt += fin;
    }
  if (out_cnt) out_cnt[qi] = cnt;
}

// 2-D fixed-block search with the block's rows FLATTENED into one candidate stream: all
// 2(2R+1) cell_start loads are issued together, then candidates are fetched U at a time
// (U independent 16-byte loads in flight) -- fewer dependent memory round trips per thread and a
// single loop whose trip count varies little across the warp.  Returns whether the result is
// certified exact (same certificate as grid_knn_block_kernel).
template <int K, int R, class L, int U = 4>
__device__ __forceinline__ bool flat2d_search(const GridDev& g, const double* __restrict__ pts, long long n,
                                              const uint8_t* __restrict__ skip, const double (&qq)[2], int cx,
                                              int cy, L& top) {
  constexpr int NR = 2 * R + 1;
  const int W = g.ncell[0], H = g.ncell[1];
  const int x0 = max(cx - R, 0), x1 = min(cx + R, W - 1);
  int s[NR], pre[NR + 1];
  pre[0] = 0;
#pragma unroll
  for (int r = 0; r < NR; r++) {
    // centre row first, then alternately below/above: near candidates tighten the bound early
    int y = cy + ((r & 1) ? -((r + 1) >> 1) : (r >> 1));
    bool valid = y >= 0 && y < H;
    long long base = (long long)(valid ? y : 0) * W;
    int ss = __ldg(g.cell_start + base + x0), ee = __ldg(g.cell_start + base + x1 + 1);
    s[r] = ss;
    pre[r + 1] = pre[r] + (valid ? ee - ss : 0);
  }
  const int total = pre[NR];
  const double2* __restrict__ P = reinterpret_cast<const double2*>(g.spts);
  if (K >= 8) {
    // The k-list update is ~1.5 K instructions when unrolled for k >= 8; keep exactly ONE copy of it
    // in the kernel (one candidate per iteration, the unindexed tail appended to the same stream):
    // with several inlined copies the kernel thrashes the instruction cache.
    const double2* __restrict__ PT = reinterpret_cast<const double2*>(pts);
    const int ntail = (int)(n - g.n_indexed);
    // candidate i+1 is fetched while candidate i goes through the list update (~100 cycles): with the load
    // issued and consumed in the same iteration every candidate exposed a full load latency (28 % of the stall
    // samples of the k = 16 kernel, profiles/r1_grid_knn_flat2d_D2_k16_details.csv)
    auto fetch = [&](int i, int& pos, double2& pv) {
      if (i < total) {
        pos = s[0] + i;
#pragma unroll
        for (int r = 1; r < NR; r++)
          if (i >= pre[r]) pos = s[r] + (i - pre[r]);
        pv = __ldg(P + pos);
      } else {
        pos = g.n_indexed + (i - total);
        pv = __ldg(PT + pos);
      }
    };
    const int nall = total + ntail;
    int pos_n = 0;
    double2 pv_n = make_double2(0.0, 0.0);
    if (nall > 0) fetch(0, pos_n, pv_n);
#pragma unroll 1
    for (int i = 0; i < nall; i++) {
      const int pos = pos_n;
      const double2 pv = pv_n;
      if (i + 1 < nall) fetch(i + 1, pos_n, pv_n);
      if (skip && skip[i < total ? __ldg(g.sidx + pos) : pos]) continue;
      double dx = pv.x - qq[0], dy = pv.y - qq[1];
      double sq = 0.0;
      sq = sq + dx * dx;
      sq = sq + dy * dy;
      top.offer_sq(sq, pos);
    }
  } else {
    for (int i = 0; i < total; i += U) {
      double2 pv[U];
      int pp[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        int ii = min(i + u, total - 1);
        int pos = s[0] + ii;
#pragma unroll
        for (int r = 1; r < NR; r++)
          if (ii >= pre[r]) pos = s[r] + (ii - pre[r]);
        pp[u] = pos;
        pv[u] = __ldg(P + pos);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (i + u < total) {
          if (skip && skip[__ldg(g.sidx + pp[u])]) continue;
          double dx = pv[u].x - qq[0], dy = pv[u].y - qq[1];
          double sq = 0.0;
          sq = sq + dx * dx;
          sq = sq + dy * dy;
          top.offer_sq(sq, pp[u]);
        }
      }
    }
    scan_tail<2>(g, pts, n, skip, qq, top);
  }
  double margin = POMP_INF;
  {
    int lo_c = cx - R, hi_c = cx + R;
    if (lo_c > 0) margin = fmin(margin, qq[0] - ((g.lo[0] + (double)lo_c * g.h[0]) + g.slack[0]));
    if (hi_c < W - 1) margin = fmin(margin, ((g.lo[0] + (double)(hi_c + 1) * g.h[0]) - g.slack[0]) - qq[0]);
    lo_c = cy - R;
    hi_c = cy + R;
    if (lo_c > 0) margin = fmin(margin, qq[1] - ((g.lo[1] + (double)lo_c * g.h[1]) + g.slack[1]));
    if (hi_c < H - 1) margin = fmin(margin, ((g.lo[1] + (double)(hi_c + 1) * g.h[1]) - g.slack[1]) - qq[1]);
  }
  return !(margin < POMP_INF) || (top.T * (1.0 + 1e-9) < margin);
}

// the rare retry: plain nested loops over a (2R+1)^2 block, R at run time -- no per-row arrays, so
// it adds code but no registers to the common path
template <int K, class L>
__device__ __forceinline__ bool block2d_search_nested(const GridDev& g, const double* __restrict__ pts, long long n,
                                                      const uint8_t* __restrict__ skip, const double (&qq)[2],
                                                      int cx, int cy, int R, L& top) {
  const int W = g.ncell[0], H = g.ncell[1];
  const int x0 = max(cx - R, 0), x1 = min(cx + R, W - 1);
  top.init(IdxResolver{g.sidx, g.n_indexed});
  for (int y = max(cy - R, 0); y <= min(cy + R, H - 1); y++) {
    long long base = (long long)y * W;
    int s = __ldg(g.cell_start + base + x0), e = __ldg(g.cell_start + base + x1 + 1);
    scan_run<2>(g, skip, qq, top, s, e);
  }
  scan_tail<2>(g, pts, n, skip, qq, top);
  double margin = POMP_INF;
  int lo_c = cx - R, hi_c = cx + R;
  if (lo_c > 0) margin = fmin(margin, qq[0] - ((g.lo[0] + (double)lo_c * g.h[0]) + g.slack[0]));
  if (hi_c < W - 1) margin = fmin(margin, ((g.lo[0] + (double)(hi_c + 1) * g.h[0]) - g.slack[0]) - qq[0]);
  lo_c = cy - R;
  hi_c = cy + R;
  if (lo_c > 0) margin = fmin(margin, qq[1] - ((g.lo[1] + (double)lo_c * g.h[1]) + g.slack[1]));
  if (hi_c < H - 1) margin = fmin(margin, ((g.lo[1] + (double)(hi_c + 1) * g.h[1]) - g.slack[1]) - qq[1]);
  return !(margin < POMP_INF) || (top.T * (1.0 + 1e-9) < margin);
}

// qs != nullptr: queries already gathered in bucket order (coalesced).  An uncertified query is
// retried once in place with the next larger block (rare: ~exp(-pi*ppc*R^2) of the queries for
// k=1); only if that fails too does it go to the overflow list for the pruned traversal.
template <int K, int R, bool RETRY = true>
__device__ __forceinline__ void flat2d_one(const GridDev& g, const double* __restrict__ pts, long long n,
                                           const uint8_t* __restrict__ skip, const double (&qq)[2], long long qi,
                                           int k, int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                           int32_t* __restrict__ out_cnt, int* __restrict__ ovf_list,
                                           int* __restrict__ ovf_count) {
  const int cx = grid_cell(g, 0, qq[0]), cy = grid_cell(g, 1, qq[1]);
  if constexpr (K >= 8) {
    extern __shared__ __align__(16) unsigned char flat_smem[];
    SmemTopK<K> top;
    top.init(IdxResolver{g.sidx, g.n_indexed}, flat_smem, blockDim.x, threadIdx.x);
    bool exact = flat2d_search<K, R>(g, pts, n, skip, qq, cx, cy, top);
    if (!exact) {
      int o = atomicAdd(ovf_count, 1);
      ovf_list[o] = (int)qi;
      return;
    }
    int cnt = 0;
    for (int j = 0; j < k; j++) {
      double dj = top.dist_at(j);
      bool fin = dj < POMP_INF;
      out_idx[qi * k + j] = fin ? (int64_t)top.index_at(j) : -1;
      out_dist[qi * k + j] = dj;
      cnt += fin;
    }
    if (out_cnt) out_cnt[qi] = cnt;
  } else {
    typename BestList<K>::type top;
    top.init(IdxResolver{g.sidx, g.n_indexed});
    bool exact = flat2d_search<K, R>(g, pts, n, skip, qq, cx, cy, top);
    if constexpr (RETRY) {
      if (!exact) exact = block2d_search_nested<K>(g, pts, n, skip, qq, cx, cy, R + 1, top);
    }
    if (!exact) {
      int o = atomicAdd(ovf_count, 1);
      ovf_list[o] = (int)qi;
      return;
    }
    int idx[K];
    top.finish(idx);
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < K; j++)
      if (j < k) {
        bool fin = top.d[j] < POMP_INF;
        out_idx[qi * k + j] = fin ? (int64_t)idx[j] : -1;
        out_dist[qi * k + j] = top.d[j];
        cnt += fin;
      }
    if (out_cnt) out_cnt[qi] = cnt;
  }
}

// Warp-cooperative retry for k = 1: the few queries whose first block gave no certificate
// (~exp(-pi*ppc*R^2): one warp in ~15 has one) are re-searched over the (2*R2+1)^2 block by the
// WHOLE warp, 32 candidates per pass, instead of by their own lane alone while 31 lanes idle.
// Same arithmetic per candidate and the same (distance, insertion index) order as the lists, so
// the answer is the one the single-thread search would give.  Every lane of the warp must call.
__device__ __forceinline__ void warp_retry_nn2d(const GridDev& g, const double* __restrict__ pts, long long n,
                                                const uint8_t* __restrict__ skip, bool need, double qx, double qy,
                                                int cx, int cy, long long qi, int R2,
                                                int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                                int32_t* __restrict__ out_cnt, int* __restrict__ ovf_list,
                                                int* __restrict__ ovf_count) {
  const unsigned FULL = 0xffffffffu;
  unsigned m = __ballot_sync(FULL, need);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  const int W = g.ncell[0], H = g.ncell[1];
  const double2* __restrict__ P = reinterpret_cast<const double2*>(g.spts);
  const double2* __restrict__ PT = reinterpret_cast<const double2*>(pts);
  const int nrow = 2 * R2 + 1;
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const double sx = __shfl_sync(FULL, qx, src), sy = __shfl_sync(FULL, qy, src);
    const int scx = __shfl_sync(FULL, cx, src), scy = __shfl_sync(FULL, cy, src);
    const int x0 = max(scx - R2, 0), x1 = min(scx + R2, W - 1);
    const int y = scy - R2 + lane;
    const bool rv = lane < nrow && y >= 0 && y < H;
    int rs = 0, rc = 0;
    if (rv) {
      rs = __ldg(g.cell_start + (long long)y * W + x0);
      rc = __ldg(g.cell_start + (long long)y * W + x1 + 1) - rs;
    }
    int incl = rc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int b = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += b;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    const int ntail = (int)(n - g.n_indexed);
    double bd = POMP_INF;
    int bi = 0x7fffffff;
    for (int base = 0; base < total + ntail; base += 32) {
      const int i = base + lane;
      int pos = -1;
      for (int r = 0; r < nrow; r++) {
        const int e = __shfl_sync(FULL, incl, r), c = __shfl_sync(FULL, rc, r), s0 = __shfl_sync(FULL, rs, r);
        if (i < e && i >= e - c) pos = s0 + (i - (e - c));
      }
      int oi = -1;
      double2 pv = make_double2(0.0, 0.0);
      if (pos >= 0) {
        oi = __ldg(g.sidx + pos);
        pv = __ldg(P + pos);
      } else if (i >= total && i < total + ntail) {
        oi = g.n_indexed + (i - total);
        pv = __ldg(PT + oi);
      }
      if (oi >= 0 && !(skip && skip[oi])) {
        double dx = pv.x - sx, dy = pv.y - sy;
        double sq = 0.0;
        sq = sq + dx * dx;
        sq = sq + dy * dy;
        const double d = sqrt(sq);
        if (d < bd || (d == bd && oi < bi)) {
          bd = d;
          bi = oi;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(FULL, bd, o);
      const int ob = __shfl_xor_sync(FULL, bi, o);
      if (od < bd || (od == bd && ob < bi)) {
        bd = od;
        bi = ob;
      }
    }
    if (lane == src) {
      double margin = POMP_INF;
      int lo_c = cx - R2, hi_c = cx + R2;
      if (lo_c > 0) margin = fmin(margin, qx - ((g.lo[0] + (double)lo_c * g.h[0]) + g.slack[0]));
      if (hi_c < W - 1) margin = fmin(margin, ((g.lo[0] + (double)(hi_c + 1) * g.h[0]) - g.slack[0]) - qx);
      lo_c = cy - R2;
      hi_c = cy + R2;
      if (lo_c > 0) margin = fmin(margin, qy - ((g.lo[1] + (double)lo_c * g.h[1]) + g.slack[1]));
      if (hi_c < H - 1) margin = fmin(margin, ((g.lo[1] + (double)(hi_c + 1) * g.h[1]) - g.slack[1]) - qy);
      const bool exact = !(margin < POMP_INF) || (bd * (1.0 + 1e-9) < margin);
      if (exact) {
        const bool fin = bd < POMP_INF;
        out_idx[qi] = fin ? (int64_t)bi : -1;
        out_dist[qi] = bd;
        if (out_cnt) out_cnt[qi] = fin ? 1 : 0;
      } else {
        int o = atomicAdd(ovf_count, 1);
        ovf_list[o] = (int)qi;
      }
    }
  }
}

template <int K, int R, bool RETRY = true, int U = 4>
__global__ void __launch_bounds__(128, (K <= 4 ? (U <= 2 ? 10 : (U <= 4 ? 8 : (U <= 6 ? 6 : 5))) : 4))  // 64 regs at U = 4; 48 (10 blocks/SM) spills and measured slower
grid_knn_flat2d_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                       const double* __restrict__ q, const double* __restrict__ qs, long long nq,
                       const int* __restrict__ qperm, int k, int64_t* __restrict__ out_idx,
                       double* __restrict__ out_dist, int32_t* __restrict__ out_cnt, int* __restrict__ ovf_list,
                       int* __restrict__ ovf_count, const int* __restrict__ qlist, const int* __restrict__ qcount) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qlist) {
    // second-chance pass: grid-stride over the overflow list of the smaller block
    const long long cnt = *qcount;
    for (; t < cnt; t += (long long)gridDim.x * blockDim.x) {
      const long long qi = qlist[t];
      double2 v = __ldg(reinterpret_cast<const double2*>(q) + qi);
      double qq[2] = {v.x, v.y};
      flat2d_one<K, R, RETRY>(g, pts, n, skip, qq, qi, k, out_idx, out_dist, out_cnt, ovf_list, ovf_count);
    }
    return;
  }
  if constexpr (K == 1 && RETRY) {
    // k = 1: no in-thread retry; uncertified queries are retried by the whole warp afterwards
    const bool active = t < nq;
    long long qi = 0;
    double qq[2] = {0.0, 0.0};
    int cx = 0, cy = 0;
    bool need = false;
    if (active) {
      qi = qperm ? (long long)__ldg(qperm + t) : t;
      double2 v =
          qs ? __ldg(reinterpret_cast<const double2*>(qs) + t) : __ldg(reinterpret_cast<const double2*>(q) + qi);
      qq[0] = v.x;
      qq[1] = v.y;
      cx = grid_cell(g, 0, qq[0]);
      cy = grid_cell(g, 1, qq[1]);
      typename BestList<1>::type top;
      top.init(IdxResolver{g.sidx, g.n_indexed});
      if (flat2d_search<1, R, typename BestList<1>::type, U>(g, pts, n, skip, qq, cx, cy, top)) {
        int idx[1];
        top.finish(idx);
        const bool fin = top.d[0] < POMP_INF;
        out_idx[qi] = fin ? (int64_t)idx[0] : -1;
        out_dist[qi] = top.d[0];
        if (out_cnt) out_cnt[qi] = fin ? 1 : 0;
      } else {
        need = true;
      }
    }
    __syncwarp();
    warp_retry_nn2d(g, pts, n, skip, need, qq[0], qq[1], cx, cy, qi, R + 1, out_idx, out_dist, out_cnt, ovf_list,
                    ovf_count);
    return;
  }
  if (t >= nq) return;
  const long long qi = qperm ? (long long)__ldg(qperm + t) : t;
  double2 v = qs ? __ldg(reinterpret_cast<const double2*>(qs) + t) : __ldg(reinterpret_cast<const double2*>(q) + qi);
  double qq[2] = {v.x, v.y};
  flat2d_one<K, R, RETRY>(g, pts, n, skip, qq, qi, k, out_idx, out_dist, out_cnt, ovf_list, ovf_count);
}

// Radius queries in ONE traversal: the count pass also parks the (unresolved) positions of up to
// RADIUS_STAGE_CAP hits per query in a staging buffer; after the scan of the counts a warp per query
// resolves them to insertion indices, sorts them (shuffle bitonic network) and writes its CSR segment
// with coalesced stores.  Only queries with more hits than that are traversed a second time.
constexpr int RADIUS_STAGE_CAP = 32;

struct RadiusStage {
  double radius, tsq;
  int inclusive;
  int cnt;
  int* stage;  // this query's RADIUS_STAGE_CAP slots
  IdxResolver res;
  __device__ __forceinline__ double threshold() const { return radius; }
  __device__ __forceinline__ double tsq_up() const { return tsq; }
  __device__ __forceinline__ void offer_sq(double s, int pos) {
    if (!(s <= tsq)) return;
    double d = sqrt(s);
    bool hit = inclusive ? (d <= radius) : (d < radius);
    if (hit) {
      // parked as the INSERTION index: its gather overlaps the traversal's other loads here, in the finalize
      // kernel it was a dependent round trip per query
      if (cnt < RADIUS_STAGE_CAP) stage[cnt] = res(pos);
      cnt++;
    }
  }
};

// thread t answers query qperm[t] (bucket-sorted batches: neighbouring threads walk neighbouring cells) or t
template <int D>
__global__ void __launch_bounds__(128)
grid_radius_stage_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                         const double* __restrict__ q, const double* __restrict__ qsorted,
                         const int* __restrict__ qperm, long long nq, double radius, int inclusive,
                         int* __restrict__ counts, int* __restrict__ stage, int* __restrict__ n_long) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  const long long qi = qperm ? (long long)__ldg(qperm + t) : t;
  double qq[D];
  if (qsorted) load_point<D>(qsorted, t, qq);
  else load_point<D>(q, qi, qq);
  RadiusStage c{radius, tsq_of(radius), inclusive, 0, stage + qi * RADIUS_STAGE_CAP, IdxResolver{g.sidx, g.n_indexed}};
  double gap[D];
#pragma unroll
  for (int i = 0; i < D; i++) gap[i] = 0.0;
  FastLevel<D, D - 1>::run(g, skip, qq, gap, 0LL, c, 0, 0);
  scan_tail<D>(g, pts, n, skip, qq, c);
  counts[qi] = c.cnt;
  if (c.cnt > RADIUS_STAGE_CAP) atomicAdd(n_long, 1);
}

// one warp per query: the parked insertion indices (32-bit) through a shuffle bitonic network, coalesced store
static __global__ void __launch_bounds__(256)
radius_finalize_kernel(const int* __restrict__ counts, const int64_t* __restrict__ offsets,
                       const int* __restrict__ stage, long long nq, int64_t* __restrict__ out) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nq) return;
  const int len = counts[w];
  if (len == 0 || len > RADIUS_STAGE_CAP) return;
  int v = lane < len ? stage[w * RADIUS_STAGE_CAP + lane] : 0x7fffffff;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = (lane & k) == 0, lower = (lane & j) == 0;
      v = (lower == up) ? min(v, o) : max(v, o);
    }
  }
  if (lane < len) out[offsets[w] + lane] = (int64_t)v;
}

template <int D, bool FILL>
__global__ void __launch_bounds__(128)
grid_radius_fast_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                        const double* __restrict__ q, long long nq, double radius, int inclusive,
                        int* __restrict__ counts, const int64_t* __restrict__ offsets, int64_t* __restrict__ out) {
  long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  if (FILL && counts && counts[qi] <= RADIUS_STAGE_CAP) return;  // finished from the staging buffer already
  double qq[D];
  load_point<D>(q, qi, qq);
  RadiusFast<FILL> c{radius, tsq_of(radius), inclusive, 0, IdxResolver{g.sidx, g.n_indexed},
                     FILL ? out + offsets[qi] : nullptr};
  double gap[D];
#pragma unroll
  for (int i = 0; i < D; i++) gap[i] = 0.0;
  FastLevel<D, D - 1>::run(g, skip, qq, gap, 0LL, c, 0, 0);
  scan_tail<D>(g, pts, n, skip, qq, c);
  if (FILL) {
    if (c.cnt > 32) heapsort_i64(out + offsets[qi], c.cnt);  // short segments: segment_sort32_kernel, warp-wide
  } else {
    counts[qi] = (int)c.cnt;
  }
}

// Hits of a radius query are reported in insertion-index order.  One warp per query sorts a segment of
// up to 32 hits with a shuffle bitonic network (coalesced load, 15 compare-exchange stages, coalesced
// store); longer segments were heap-sorted by the thread that produced them.  (A per-thread heap sort in
// global memory for every query made the fill pass 4x as long as the count pass.)
static __global__ void __launch_bounds__(256)
segment_sort32_kernel(const int64_t* __restrict__ offsets, int64_t* __restrict__ vals, long long nq) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nq) return;
  const int64_t b = offsets[w];
  const long long len = offsets[w + 1] - b;
  if (len < 2 || len > 32) return;
  long long v = lane < len ? vals[b + lane] : 0x7fffffffffffffffLL;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const long long o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = (lane & k) == 0, lower = (lane & j) == 0;
      v = (lower == up) ? (v < o ? v : o) : (v > o ? v : o);
    }
  }
  if (lane < len) vals[b + lane] = v;
}

// EST Gaussian density for the plain Euclidean metric: the radius traversal with the hits folded into
// sum exp(-(d/sigma)^2) (kinodynamicplanner.py:688-693)
struct DensityFast {
  double radius, tsq, inv_sigma;
  int inclusive;
  double sum;
  int cnt;
  __device__ __forceinline__ double threshold() const { return radius; }
  __device__ __forceinline__ double tsq_up() const { return tsq; }
  __device__ __forceinline__ void offer_sq(double s, int) {
    if (!(s <= tsq)) return;
    double d = sqrt(s);
    bool hit = inclusive ? (d <= radius) : (d < radius);
    if (hit) {
      double t = d * inv_sigma;
      sum += exp(-(t * t));
      cnt++;
    }
  }
};

template <int D>
__global__ void __launch_bounds__(128)
grid_density_fast_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                         const double* __restrict__ q, long long nq, double radius, double sigma, int inclusive,
                         double* __restrict__ dens, int32_t* __restrict__ cnt) {
  long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  double qq[D];
  load_point<D>(q, qi, qq);
  DensityFast c{radius, tsq_of(radius), 1.0 / sigma, inclusive, 0.0, 0};
  double gap[D];
#pragma unroll
  for (int i = 0; i < D; i++) gap[i] = 0.0;
  FastLevel<D, D - 1>::run(g, skip, qq, gap, 0LL, c, 0, 0);
  scan_tail<D>(g, pts, n, skip, qq, c);
  dens[qi] = c.sum;
  if (cnt) cnt[qi] = c.cnt;
}

}  // namespace pomp
