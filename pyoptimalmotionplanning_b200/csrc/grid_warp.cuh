// grid_warp.cuh -- k-nearest neighbours with a GROUP OF LANES PER QUERY (W = 8, 16 or 32 lanes; 32 / W queries per
// warp) over the uniform-grid index (plain Euclidean metric, every coordinate a grid axis: the 2-D..6-D point clouds
// of BASELINE.json, k <= 32).
//
// Why: the thread-per-query kernels (grid_fast.cuh) are latency-bound once a query has more than a handful of
// candidates -- every lane walks its own pruned cell sequence (divergent), reads its own 8*D-byte points
// (uncoalesced) and pushes every hit through a private sorted list (6-D k = 16: 29 ms per 2^20 queries, 0.006 of the
// HBM roofline).  Here the lanes of a group share one query:
//   * SHELLS.  Cells are visited in shells of increasing lower bound.  Shell i holds the cells whose squared gap
//     bound s_b (the metric's own sum of squared per-axis gaps, coordinate order, see grid.cuh) satisfies
//     tsq(R_{i-1}) < s_b <= tsq(R_i).  After a shell, if the k-th distance T <= R_i the answer is exact: every
//     unvisited cell has fl(sqrt(s_b)) > R_i >= T, and no point of a cell is closer than its bound.  Otherwise
//     R_{i+1} = T (one more shell finishes it) or 2 R_i while fewer than k points were seen.  R_1 is the radius that
//     holds `need` points at the mean density; where the first rows of that shell hold far more points than that
//     (a cluster), the k-th distance among the query's neighbours in sorted order replaces it.
//   * ROWS, LANE-PARALLEL.  Per shell the squared gaps of every (axis, cell) of the bounding box are tabulated once
//     in shared memory (<= WQ_EXT cells per axis, computed directly beyond that); then each lane takes one row of
//     cells along axis 0 (centre-out digit order, so near rows come first), sums its bound, trims the row to the
//     cells inside the shell and reads the run bounds from cell_start.
//   * ONE CANDIDATE QUEUE.  The runs of W rows are concatenated by a prefix sum over the group; the lanes then take
//     W consecutive candidates at a time (binary search of the prefix through shuffles), two such slabs in flight,
//     each candidate's coordinates and insertion index fetched together: all lanes evaluate distances whatever the
//     run lengths, and a run of cells is read by neighbouring lanes.
//   * SELECTION.  k = 1: every lane keeps its own best (distance, insertion index); a float upper bound of the
//     group minimum (one redux.sync) prunes rows; one shuffle arg-min per shell.  k > 1: a sorted list of 32
//     (distance, insertion index) keys spread over the group (position p in lane p % W, slot p / W).  Candidates
//     are pre-filtered on the squared distance against tsq(T) and inserted by ballot / shuffle; with W = 32 a slab
//     with many survivors is merged in bulk (bitonic sort of the slab, min against the reversed list, bitonic
//     merge: ~250 instructions whatever the count).
// Narrow groups put 32 / W queries into every instruction of the bookkeeping (shell set-up, row decoding, prefix
// sums, list insertions -- 2/3 of the instructions of the W = 32 version, profiles/r2_warp_D3_k16) and 32 / W times
// as many queries in flight per SM; the groups of a warp loop independently (loops are group-uniform, every
// collective names only the group's lanes).
// Distances are rooted once per surviving candidate and compared as (distance, insertion index) exactly like
// KNearestResult (knn.py:7-30) with its ties broken by insertion order: bit-identical to the other kernels.
#pragma once
#include "grid_fast.cuh"

namespace pomp {

constexpr int WQ_EXT = 8;        // tabulated cells per axis (a shell's box is 3..5 cells wide)
constexpr int WQ_WARPS = 8;      // warps per block
#ifndef WQ_BULK
#define WQ_BULK 7
#endif
constexpr int WQ_BULK_MIN = WQ_BULK;   // W = 32: survivors in a slab from which the bulk merge beats single insertions
constexpr int WQ_DENSE = 24;     // candidates PER LANE in the first rows of the first shell beyond which the mean
                                 // density is taken to be useless (a cluster): the radius then comes from the
                                 // neighbours in sorted order instead
#ifndef WQ_OCC
#define WQ_OCC 4
#endif
#ifndef WQ_OCC6
#define WQ_OCC6 3
#endif

__device__ __forceinline__ bool wkey_less(double d1, int i1, double d2, int i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

// centre-out enumeration of the cells lo..hi of an axis around `home` (lo <= home <= hi): t = 0, 1, 2, ... gives
// home, home-1, home+1, home-2, ... until one side is exhausted, then the rest of the other side
__device__ __forceinline__ int centre_out(int t, int home, int lo, int hi) {
  const int L = home - lo, U = hi - home;
  const int m = min(L, U);
  if (t <= 2 * m) return home + ((t & 1) ? -((t + 1) >> 1) : (t >> 1));
  return L > U ? home - (t - U) : home + (t - L);
}

// ---- the lanes of one query -------------------------------------------------------------------------------------------
template <int W>
struct Group {
  unsigned gm;  // lanes of the group
  int gsh;      // first lane of the group
  int gl;       // lane within the group
  __device__ __forceinline__ void init() {
    const int lane = threadIdx.x & 31;
    gl = lane & (W - 1);
    gsh = lane & ~(W - 1);
    gm = W == 32 ? 0xffffffffu : (((1u << (W & 31)) - 1u) << gsh);
  }
  __device__ __forceinline__ unsigned ballot(bool p) const {
    const unsigned b = __ballot_sync(gm, p);
    return W == 32 ? b : ((b >> gsh) & ((1u << (W & 31)) - 1u));
  }
  template <typename T>
  __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(gm, v, src, W); }
  template <typename T>
  __device__ __forceinline__ T shfl_up(T v, int o) const { return __shfl_up_sync(gm, v, o, W); }
  template <typename T>
  __device__ __forceinline__ T shfl_xor(T v, int o) const { return __shfl_xor_sync(gm, v, o, W); }
  __device__ __forceinline__ void sync() const { __syncwarp(gm); }
};

// ---- sorted list of W * KL keys spread over the group: position p in lane p % W, slot p / W ------------------------
// KL = 1 (one key per lane: k <= W) has the bulk merge; KL > 1 only the single insertions
template <int W, int KL>
struct GroupList {
  double d[KL];
  int i[KL];
  double kd;  // k-th key, group-uniform
  int ki;
  double ktsq;
  int k;
  __device__ __forceinline__ void init(int k_) {
    k = k_;
#pragma unroll
    for (int s = 0; s < KL; s++) {
      d[s] = POMP_INF;
      i[s] = IDX_NONE;
    }
    kd = POMP_INF;
    ki = IDX_NONE;
    ktsq = POMP_INF;
  }
  __device__ __forceinline__ void refresh(const Group<W>& G) {
    const int p = k - 1;
    double vd = d[0];
    int vi = i[0];
#pragma unroll
    for (int s = 1; s < KL; s++)
      if (p / W == s) {
        vd = d[s];
        vi = i[s];
      }
    kd = G.shfl(vd, p % W);
    ki = G.shfl(vi, p % W);
    ktsq = tsq_of(kd);  // tsq_of(inf) = inf
  }
  __device__ __forceinline__ bool empty(const Group<W>& G) const { return !(G.shfl(d[0], 0) < POMP_INF); }
  // one key, the same in all lanes of the group
  __device__ __forceinline__ void insert_uniform(const Group<W>& G, double cd, int ci) {
    int P = 0;
#pragma unroll
    for (int s = 0; s < KL; s++) P += __popc(G.ballot(wkey_less(d[s], i[s], cd, ci)));
#pragma unroll
    for (int s = KL - 1; s >= 0; s--) {
      double ud = G.shfl_up(d[s], 1);
      int ui = G.shfl_up(i[s], 1);
      if (s > 0) {
        const double pd = G.shfl(d[s - 1], W - 1);
        const int pi = G.shfl(i[s - 1], W - 1);
        if (G.gl == 0) {
          ud = pd;
          ui = pi;
        }
      }
      const int p = s * W + G.gl;
      if (p > P) {
        d[s] = ud;
        i[s] = ui;
      } else if (p == P) {
        d[s] = cd;
        i[s] = ci;
      }
    }
    refresh(G);
  }
  // one candidate per lane (cd = +inf: none)
  __device__ __forceinline__ void offer_each(const Group<W>& G, double cd, int ci) {
    bool pass = (cd < POMP_INF) && wkey_less(cd, ci, kd, ki);
    unsigned mask = G.ballot(pass);
    while (mask) {
      const int src = __ffs(mask) - 1;
      const double bd = G.shfl(cd, src);
      const int bi = G.shfl(ci, src);
      insert_uniform(G, bd, bi);
      if (G.gl == src) pass = false;
      pass = pass && wkey_less(cd, ci, kd, ki);
      mask = G.ballot(pass);
    }
  }
  static __device__ __forceinline__ void cmpx(const Group<W>& G, double& kd_, int& ki_, int stride, bool keep_min) {
    const double od = G.shfl_xor(kd_, stride);
    const int oi = G.shfl_xor(ki_, stride);
    const bool oless = wkey_less(od, oi, kd_, ki_);
    if (keep_min ? oless : !oless) {
      kd_ = od;
      ki_ = oi;
    }
  }
  // KL = 1 only.  One candidate per lane, many expected to pass: sort the slab (bitonic network over the W lanes),
  // keep the W smallest of slab + list
  __device__ __forceinline__ void offer_bulk(const Group<W>& G, double cd, int ci, bool list_empty) {
    static_assert(KL == 1, "bulk merge needs one key per lane");
    const int lane = G.gl;
#pragma unroll
    for (int size = 2; size <= W; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const bool up = (lane & size) == 0 || size == W;
        const bool lower = (lane & stride) == 0;
        cmpx(G, cd, ci, stride, lower == up);
      }
    }
    if (!list_empty) {
      // min(list[l], slab[W - 1 - l]) is a bitonic sequence holding the W smallest keys of the union
      const double rd = G.shfl(cd, W - 1 - lane);
      const int ri = G.shfl(ci, W - 1 - lane);
      if (wkey_less(d[0], i[0], rd, ri)) {
        cd = d[0];
        ci = i[0];
      } else {
        cd = rd;
        ci = ri;
      }
#pragma unroll
      for (int stride = W / 2; stride > 0; stride >>= 1) cmpx(G, cd, ci, stride, (lane & stride) == 0);
    }
    d[0] = cd;
    i[0] = ci;
    refresh(G);
  }
};

// ---- per-query shell geometry ------------------------------------------------------------------------------------------
template <int D, int W>
struct WarpShell {
  int lo[D], hi[D];
  __device__ __forceinline__ void set(const GridDev& g, const Group<W>& G, const double (&q)[D], double R, double* tab) {
#pragma unroll
    for (int a = 0; a < D; a++) grid_range(g, a, q[a], R, &lo[a], &hi[a]);
    G.sync();
    for (int e = G.gl; e < D * WQ_EXT; e += W) {
      const int a = e / WQ_EXT, o = e - a * WQ_EXT;
      int la = 0, ha = 0;
      double qa = 0.0;
#pragma unroll
      for (int b = 0; b < D; b++)
        if (b == a) {
          la = lo[b];
          ha = hi[b];
          qa = q[b];
        }
      if (la + o <= ha) {
        const double gp = grid_gap(g, a, la + o, qa);
        tab[e] = gp * gp;
      }
    }
    G.sync();
  }
  template <int A>
  __device__ __forceinline__ double gsq(const GridDev& g, const double (&q)[D], const double* tab, int c) const {
    const int o = c - lo[A];
    if (o < WQ_EXT) return tab[A * WQ_EXT + o];
    const double gp = grid_gap(g, A, c, q[A]);
    return gp * gp;
  }
};

template <int D, int W, int A>
struct RowDecode {
  // digits of the row number, axis 1 fastest; accumulates the cell id and the squared gaps in coordinate order
  static __device__ __forceinline__ void run(const GridDev& g, const WarpShell<D, W>& sh, const double (&q)[D],
                                             const int (&home)[D], const double* tab, unsigned j, long long& base,
                                             double (&gs)[D]) {
    if constexpr (A < D) {
      const int ext = sh.hi[A] - sh.lo[A] + 1;
      unsigned jq;
      int t;
      if (j < (1u << 22)) {
        // float reciprocal instead of an integer division (j and ext are exact in float; the quotient is off by
        // at most one)
        jq = (unsigned)(__uint2float_rz(j) * __frcp_rn((float)ext));
        t = (int)(j - jq * (unsigned)ext);
        if (t < 0) {
          t += ext;
          jq--;
        } else if (t >= ext) {
          t -= ext;
          jq++;
        }
      } else {
        jq = j / (unsigned)ext;
        t = (int)(j - jq * (unsigned)ext);
      }
      const int c = centre_out(t, home[A], sh.lo[A], sh.hi[A]);
      base += (long long)c * g.stride[A];
      gs[A] = sh.template gsq<A>(g, q, tab, c);
      RowDecode<D, W, A + 1>::run(g, sh, q, home, tab, jq, base, gs);
    }
  }
};

template <int D>
__device__ __forceinline__ double fold_bound(double g0sq, const double (&gs)[D]) {
  double s = 0.0;
  s = s + g0sq;
#pragma unroll
  for (int a = 1; a < D; a++) s = s + gs[a];
  return s;
}

// one candidate of a slab: coordinates and insertion index are fetched together (the index speculatively: waiting
// for the distance first would put a second DRAM round trip behind every slab)
template <int D>
struct WCand {
  double p[D];
  int id;
  bool cv;
  __device__ __forceinline__ void fetch(bool valid, const double* __restrict__ base, const int* __restrict__ sidx,
                                        long long pos) {
    cv = valid;
    id = IDX_NONE;
    if (valid) {
      load_point<D>(base, pos, p);
      id = sidx ? __ldg(sidx + pos) : (int)pos;
    }
  }
};

// K1 = true: nearest neighbour (lane-local bests); false: sorted list of W * KL keys, k <= W * KL.  W lanes per query.
template <int D, bool K1, int W, int KL>
__global__ void __launch_bounds__(WQ_WARPS * 32, (D >= 6 ? WQ_OCC6 : WQ_OCC))
grid_knn_warp_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                     const double* __restrict__ q, const double* __restrict__ qsorted, long long nq,
                     const int* __restrict__ qperm, int k, double r0, double rmax, int64_t* __restrict__ out_idx,
                     double* __restrict__ out_dist, int32_t* __restrict__ out_cnt) {
  constexpr int GPB = WQ_WARPS * 32 / W;  // queries per block
  __shared__ double s_tab[GPB][D * WQ_EXT];
  Group<W> G;
  G.init();
  const int gib = threadIdx.x / W;
  const long long w = (long long)blockIdx.x * GPB + gib;
  if (w >= nq) return;
  double* tab = s_tab[gib];
  double qq[D];
  long long qi = w;
  if (qsorted) {
    load_point<D>(qsorted, w, qq);
    qi = (long long)__ldg(qperm + w);  // only needed for the result stores
  } else {
    if (qperm) qi = (long long)__ldg(qperm + w);
    load_point<D>(q, qi, qq);
  }
  bool finite = true;
#pragma unroll
  for (int a = 0; a < D; a++) finite = finite && (fabs(qq[a]) < POMP_INF);  // false for NaN
  if (!finite) {
    // every distance is inf or NaN: `d < dbest` never holds (nearestneighbors.py:88-96)
    for (int j = G.gl; j < k; j += W) {
      out_idx[qi * k + j] = -1;
      out_dist[qi * k + j] = POMP_INF;
    }
    if (out_cnt && G.gl == 0) out_cnt[qi] = 0;
    return;
  }
  int home[D];
#pragma unroll
  for (int a = 0; a < D; a++) home[a] = grid_cell(g, a, qq[a]);

  GroupList<W, KL> top;
  top.init(k);
  // K1 state: lane-local best and the group-uniform (conservative) threshold
  double bd = POMP_INF, btsq = POMP_INF, utsq = POMP_INF;
  int bi = IDX_NONE;
  double kd = POMP_INF;  // k-th distance so far, group-uniform (K1: after the reductions)
  int ki = IDX_NONE;

  // offers the fetched candidate of this lane; all lanes of the group call
  auto offer = [&](const WCand<D>& c) {
    double s = POMP_INF;
    if (c.cv) s = sqdist<D>(c.p, qq);
    if constexpr (K1) {
      if (c.cv && s <= btsq) {
        const double d = sqrt(s);
        if (!(skip && skip[c.id]) && wkey_less(d, c.id, bd, bi)) {
          bd = d;
          bi = c.id;
          btsq = tsq_of(d);
        }
      }
    } else {
      bool pass = c.cv && s <= top.ktsq;  // NaN fails
      double d = POMP_INF;
      if (pass) {
        d = sqrt(s);
        if (skip && skip[c.id]) d = POMP_INF;
        pass = d < POMP_INF && wkey_less(d, c.id, top.kd, top.ki);
        if (!pass) d = POMP_INF;
      }
      const unsigned m = G.ballot(pass);
      if (m) {
        if constexpr (KL == 1) {
          if (__popc(m) >= (W >= 32 ? WQ_BULK_MIN : (W >= 16 ? 5 : 4))) top.offer_bulk(G, d, c.id, top.empty(G));
          else top.offer_each(G, d, c.id);
        } else {
          top.offer_each(G, d, c.id);
        }
      }
    }
  };
  // K1: float upper bound of the group's best distance -> squared prune bound for rows
  auto k1_uniform = [&]() {
    if constexpr (K1) {
      const float f = __double2float_ru(bd);  // +inf stays +inf; >= bd
      const unsigned u = __reduce_min_sync(G.gm, __float_as_uint(f));
      utsq = tsq_of((double)__uint_as_float(u));
    }
  };
  auto prune_tsq = [&]() -> double {
    if constexpr (K1) return utsq;
    else return top.ktsq;
  };
  auto reduce_k = [&]() {
    if constexpr (K1) {
      double rd = bd;
      int ri = bi;
#pragma unroll
      for (int o = W / 2; o > 0; o >>= 1) {
        const double od = G.shfl_xor(rd, o);
        const int oi = G.shfl_xor(ri, o);
        if (wkey_less(od, oi, rd, ri)) {
          rd = od;
          ri = oi;
        }
      }
      kd = rd;
      ki = ri;
    } else {
      kd = top.kd;
      ki = top.ki;
    }
  };

  // ---- points appended after the last index build (position = insertion index)
  for (long long b = g.n_indexed; b < n; b += W) {
    WCand<D> c;
    c.fetch(b + G.gl < n, pts, nullptr, b + G.gl);
    offer(c);
  }
  if (n > g.n_indexed) {
    k1_uniform();
    reduce_k();
  }

  // ---- shells
  WarpShell<D, W> sh;
  int w0 = 0, w1 = 0;    // sorted positions offered by the dense-mode window (never again)
  bool windowed = false;
  double R = fmin(r0, kd);  // kd = inf while fewer than k points were seen
  double tq_prev = -1.0;
  while (true) {
    const double tq = tsq_of(R);
    sh.set(g, G, qq, R, tab);
    unsigned nrows = 1;  // < 2^31: the launcher requires n_cells < 2^31 - 64
#pragma unroll
    for (int a = 1; a < D; a++) nrows *= (unsigned)(sh.hi[a] - sh.lo[a] + 1);
    bool restart = false;
    for (unsigned rb = 0; rb < nrows; rb += W) {
      const unsigned j = rb + G.gl;
      int s1 = 0, n1 = 0, s2 = 0, n2 = 0;
      if (j < nrows) {
        long long base = 0;
        double gs[D];
        gs[0] = 0.0;
        RowDecode<D, W, 1>::run(g, sh, qq, home, tab, j, base, gs);
        const double tcur = fmin(tq, prune_tsq());
        const double srow = fold_bound<D>(0.0, gs);
        if (!(srow > tcur)) {
          // cells of the row inside the shell (bound <= tcur): the home column always is
          int a0 = sh.lo[0], b0 = sh.hi[0];
          while (a0 < home[0] && fold_bound<D>(sh.template gsq<0>(g, qq, tab, a0), gs) > tcur) a0++;
          while (b0 > home[0] && fold_bound<D>(sh.template gsq<0>(g, qq, tab, b0), gs) > tcur) b0--;
          if (tq_prev >= 0.0 && !(srow > tq_prev)) {
            // cells with bound <= tq_prev were visited (or pruned for good) by the previous shell
            int a1 = a0, b1 = b0;
            while (a1 < home[0] && fold_bound<D>(sh.template gsq<0>(g, qq, tab, a1), gs) > tq_prev) a1++;
            while (b1 > home[0] && fold_bound<D>(sh.template gsq<0>(g, qq, tab, b1), gs) > tq_prev) b1--;
            if (a1 > a0) {
              s1 = __ldg(g.cell_start + base + a0);
              n1 = __ldg(g.cell_start + base + a1) - s1;
            }
            if (b1 < b0) {
              s2 = __ldg(g.cell_start + base + b1 + 1);
              n2 = __ldg(g.cell_start + base + b0 + 1) - s2;
            }
          } else {
            s1 = __ldg(g.cell_start + base + a0);
            n1 = __ldg(g.cell_start + base + b0 + 1) - s1;
          }
        }
      }
      // concatenate the runs of the W rows
      const int nl = n1 + n2;
      int incl = nl;
#pragma unroll
      for (int o = 1; o < W; o <<= 1) {
        const int t = G.shfl_up(incl, o);
        if (G.gl >= o) incl += t;
      }
      const int total = G.shfl(incl, W - 1);
      const int excl = incl - nl;
      if (rb == 0 && tq_prev < 0.0 && !windowed && total > WQ_DENSE * W) {
        // a cluster: the k-th distance among 32 neighbours in sorted order bounds the radius much better
        long long hcell = 0;
#pragma unroll
        for (int a = 0; a < D; a++) hcell += (long long)home[a] * g.stride[a];
        const int hs = __ldg(g.cell_start + hcell), he = __ldg(g.cell_start + hcell + 1);
        const int center = (hs + he) >> 1;
        w0 = max(0, min(center - 16, g.n_indexed - 32));
        w1 = min(g.n_indexed, w0 + 32);
        for (int o = G.gl; o < 32; o += W) {  // group-uniform trip count
          WCand<D> c;
          c.fetch(w0 + o < w1, g.spts, g.sidx, w0 + o);
          offer(c);
        }
        k1_uniform();
        reduce_k();
        windowed = true;
        if (kd < R) {
          R = kd;
          restart = true;
          break;
        }
      }
      auto locate = [&](int item, bool& cv) -> int {
        int own = 0;
#pragma unroll
        for (int step = W / 2; step > 0; step >>= 1) {
          const int v = G.shfl(incl, own + step - 1);
          if (v <= item) own += step;
        }
        own = min(own, W - 1);
        const int os1 = G.shfl(s1, own), on1 = G.shfl(n1, own);
        const int os2 = G.shfl(s2, own), oex = G.shfl(excl, own);
        const int local = item - oex;
        const int pos = local < on1 ? os1 + local : os2 + (local - on1);
        cv = item < total && !(pos >= w0 && pos < w1);
        return pos;
      };
      // two slabs of W candidates in flight
      for (int e = 0; e < total; e += 2 * W) {
        WCand<D> c0, c1;
        bool v0, v1;
        const int p0 = locate(e + G.gl, v0);
        c0.fetch(v0, g.spts, g.sidx, p0);
        const bool second = e + W < total;
        if (second) {
          const int p1 = locate(e + W + G.gl, v1);
          c1.fetch(v1, g.spts, g.sidx, p1);
        }
        offer(c0);
        if (second) offer(c1);
      }
      k1_uniform();
    }
    if (restart) continue;
    reduce_k();
    if (kd <= R || !(R < POMP_INF)) break;  // exact, or everything was visited
    tq_prev = tq;
    if (kd < POMP_INF) R = kd;
    else {
      R = R * 2.0;
      if (!(R < rmax) || !(R > 0.0)) R = POMP_INF;
    }
  }

  // ---- results
  if constexpr (K1) {
    if (G.gl == 0) {
      out_idx[qi * k] = kd < POMP_INF ? (int64_t)ki : -1;
      out_dist[qi * k] = kd;
      if (out_cnt) out_cnt[qi] = kd < POMP_INF ? 1 : 0;
    }
    for (int j = 1 + G.gl; j < k; j += W) {  // k > 1 never comes here (host dispatch); defensive
      out_idx[qi * k + j] = -1;
      out_dist[qi * k + j] = POMP_INF;
    }
  } else {
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < KL; s++) {
      const int p = s * W + G.gl;
      const bool fin = top.d[s] < POMP_INF && p < k;
      if (p < k) {
        out_idx[qi * k + p] = fin ? (int64_t)top.i[s] : -1;
        out_dist[qi * k + p] = top.d[s];
      }
      cnt += __popc(G.ballot(fin));
    }
    if (out_cnt && G.gl == 0) out_cnt[qi] = cnt;
  }
}

}  // namespace pomp
