// grid_warp.cuh -- k-nearest neighbours with ONE WARP PER QUERY over the uniform-grid index (plain Euclidean metric,
// every coordinate a grid axis: the 2-D..6-D point clouds of BASELINE.json, k <= 32).
//
// Why: the thread-per-query kernels (grid_fast.cuh) are latency-bound once a query has more than a handful of
// candidates -- every lane walks its own pruned cell sequence (divergent), reads its own 8*D-byte points
// (uncoalesced) and pushes every hit through a private sorted list (6-D k = 16: 29 ms per 2^20 queries, 0.006 of the
// HBM roofline).  Here a warp shares one query:
//   * SHELLS.  Cells are visited in shells of increasing lower bound.  Shell i holds the cells whose squared gap
//     bound s_b (the metric's own sum of squared per-axis gaps, coordinate order, see grid.cuh) satisfies
//     tsq(R_{i-1}) < s_b <= tsq(R_i).  After a shell, if the k-th distance T <= R_i the answer is exact: every
//     unvisited cell has fl(sqrt(s_b)) > R_i >= T, and no point of a cell is closer than its bound.  Otherwise
//     R_{i+1} = T (one more shell finishes it) or 2 R_i while fewer than k points were seen.  R_1 is the radius that
//     holds `need` points at the mean density, capped by the k-th distance among the query's neighbours in sorted
//     order (a valid upper bound, and much smaller than R_1 inside clusters).
//   * ROWS, LANE-PARALLEL.  Per shell the squared gaps of every (axis, cell) of the bounding box are tabulated once
//     in shared memory (<= WQ_EXT cells per axis, computed directly beyond that); then each lane takes one row of
//     cells along axis 0 (centre-out digit order, so near rows come first), sums its bound, trims the row to the
//     cells inside the shell and reads the run bounds from cell_start.
//   * ONE CANDIDATE QUEUE.  The runs of 32 rows are concatenated by a warp prefix sum; the lanes then take 32
//     consecutive candidates at a time (binary search of the prefix through shuffles), so all lanes evaluate
//     distances whatever the run lengths, and a run of cells is read by neighbouring lanes.
//   * SELECTION.  k = 1: every lane keeps its own best (distance, insertion index); a float upper bound of the
//     warp minimum (one redux.sync) prunes rows; one shuffle arg-min per shell.  k > 1: a sorted list of <= 32
//     (distance, insertion index) keys, one per lane.  Candidates are pre-filtered on the squared distance against
//     tsq(T); a slab with many survivors is merged in bulk (bitonic sort of the slab, min against the reversed list,
//     bitonic merge: ~250 instructions whatever the count), a slab with few by ballot/shuffle insertion.
// Distances are rooted once per surviving candidate and compared as (distance, insertion index) exactly like
// KNearestResult (knn.py:7-30) with its ties broken by insertion order: bit-identical to the other kernels.
#pragma once
#include "grid_fast.cuh"

namespace pomp {

constexpr int WQ_EXT = 8;        // tabulated cells per axis (a shell's box is 3..5 cells wide)
constexpr int WQ_WARPS = 8;      // warps (= queries in flight) per block
constexpr int WQ_BULK_MIN = 7;   // survivors in a slab from which the bulk merge is cheaper than single insertions

struct WKey {
  double d;
  int i;
};
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

// ---- sorted warp list (one key per lane) with bulk merge -----------------------------------------------------------
struct WarpList {
  double d;   // lane l holds position l
  int i;
  double kd;  // k-th key, warp-uniform
  int ki;
  double ktsq;
  int k;
  __device__ __forceinline__ void init(int k_) {
    k = k_;
    d = POMP_INF;
    i = IDX_NONE;
    kd = POMP_INF;
    ki = IDX_NONE;
    ktsq = POMP_INF;
  }
  __device__ __forceinline__ void refresh() {
    kd = __shfl_sync(FULL_MASK, d, k - 1);
    ki = __shfl_sync(FULL_MASK, i, k - 1);
    ktsq = tsq_of(kd);  // tsq_of(inf) = inf
  }
  // one key, the same in all lanes
  __device__ __forceinline__ void insert_uniform(double cd, int ci, int lane) {
    const int P = __popc(__ballot_sync(FULL_MASK, wkey_less(d, i, cd, ci)));
    const double ud = __shfl_up_sync(FULL_MASK, d, 1);
    const int ui = __shfl_up_sync(FULL_MASK, i, 1);
    if (lane > P) {
      d = ud;
      i = ui;
    } else if (lane == P) {
      d = cd;
      i = ci;
    }
    refresh();
  }
  // one candidate per lane (cd = +inf: none), few of them expected to pass
  __device__ __forceinline__ void offer_each(double cd, int ci, int lane) {
    bool pass = (cd < POMP_INF) && wkey_less(cd, ci, kd, ki);
    unsigned mask = __ballot_sync(FULL_MASK, pass);
    while (mask) {
      const int src = __ffs(mask) - 1;
      const double bd = __shfl_sync(FULL_MASK, cd, src);
      const int bi = __shfl_sync(FULL_MASK, ci, src);
      insert_uniform(bd, bi, lane);
      if (lane == src) pass = false;
      pass = pass && wkey_less(cd, ci, kd, ki);
      mask = __ballot_sync(FULL_MASK, pass);
    }
  }
  static __device__ __forceinline__ void cmpx(double& kd_, int& ki_, int stride, bool keep_min) {
    const double od = __shfl_xor_sync(FULL_MASK, kd_, stride);
    const int oi = __shfl_xor_sync(FULL_MASK, ki_, stride);
    const bool oless = wkey_less(od, oi, kd_, ki_);
    if (keep_min ? oless : !oless) {
      kd_ = od;
      ki_ = oi;
    }
  }
  // one candidate per lane, many expected to pass: sort the slab, keep the 32 smallest of slab + list
  __device__ __forceinline__ void offer_bulk(double cd, int ci, int lane, bool list_empty) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const bool up = (lane & size) == 0 || size == 32;
        const bool lower = (lane & stride) == 0;
        cmpx(cd, ci, stride, lower == up);
      }
    }
    if (!list_empty) {
      // min(list[l], slab[31 - l]) is a bitonic sequence holding the 32 smallest keys of the union
      const double rd = __shfl_sync(FULL_MASK, cd, 31 - lane);
      const int ri = __shfl_sync(FULL_MASK, ci, 31 - lane);
      if (wkey_less(d, i, rd, ri)) {
        cd = d;
        ci = i;
      } else {
        cd = rd;
        ci = ri;
      }
#pragma unroll
      for (int stride = 16; stride > 0; stride >>= 1) cmpx(cd, ci, stride, (lane & stride) == 0);
    }
    d = cd;
    i = ci;
    refresh();
  }
};

// ---- per-warp shell geometry ---------------------------------------------------------------------------------------
template <int D>
struct WarpShell {
  int lo[D], hi[D];
  __device__ __forceinline__ void set(const GridDev& g, const double (&q)[D], double R, double* tab, int lane) {
#pragma unroll
    for (int a = 0; a < D; a++) grid_range(g, a, q[a], R, &lo[a], &hi[a]);
    __syncwarp();
    for (int e = lane; e < D * WQ_EXT; e += 32) {
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
    __syncwarp();
  }
  template <int A>
  __device__ __forceinline__ double gsq(const GridDev& g, const double (&q)[D], const double* tab, int c) const {
    const int o = c - lo[A];
    if (o < WQ_EXT) return tab[A * WQ_EXT + o];
    const double gp = grid_gap(g, A, c, q[A]);
    return gp * gp;
  }
};

template <int D, int A>
struct RowDecode {
  // digits of the row number, axis 1 fastest; accumulates the cell id and the squared gaps in coordinate order
  static __device__ __forceinline__ void run(const GridDev& g, const WarpShell<D>& sh, const double (&q)[D],
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
      RowDecode<D, A + 1>::run(g, sh, q, home, tab, jq, base, gs);
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

#ifndef WQ_OCC
#define WQ_OCC 4
#endif
#ifndef WQ_OCC6
#define WQ_OCC6 3
#endif
constexpr int WQ_DENSE = 768;  // candidates in the first 32 rows of the first shell beyond which the mean density
                               // is taken to be useless (a cluster): the radius then comes from the neighbours in
                               // sorted order instead

// K1 = true: nearest neighbour (lane-local bests); false: sorted warp list, k <= 32
template <int D, bool K1>
__global__ void __launch_bounds__(WQ_WARPS * 32, (D >= 6 ? WQ_OCC6 : WQ_OCC))
grid_knn_warp_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                     const double* __restrict__ q, const double* __restrict__ qsorted, long long nq,
                     const int* __restrict__ qperm, int k, double r0, double rmax, int64_t* __restrict__ out_idx,
                     double* __restrict__ out_dist, int32_t* __restrict__ out_cnt) {
  __shared__ double s_tab[WQ_WARPS][D * WQ_EXT];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long w = (long long)blockIdx.x * WQ_WARPS + wib;
  if (w >= nq) return;
  double* tab = s_tab[wib];
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
    for (int j = lane; j < k; j += 32) {
      out_idx[qi * k + j] = -1;
      out_dist[qi * k + j] = POMP_INF;
    }
    if (out_cnt && lane == 0) out_cnt[qi] = 0;
    return;
  }
  int home[D];
#pragma unroll
  for (int a = 0; a < D; a++) home[a] = grid_cell(g, a, qq[a]);

  WarpList top;
  top.init(k);
  // K1 state: lane-local best and the warp-uniform (conservative) threshold
  double bd = POMP_INF, btsq = POMP_INF, utsq = POMP_INF;
  int bi = IDX_NONE;
  double kd = POMP_INF;  // k-th distance so far, warp-uniform (K1: after the reductions)
  int ki = IDX_NONE;

  // offers the fetched candidate of this lane; all lanes call
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
      const unsigned m = __ballot_sync(FULL_MASK, pass);
      if (m) {
        if (__popc(m) >= WQ_BULK_MIN)
          top.offer_bulk(d, c.id, lane, !(__shfl_sync(FULL_MASK, top.d, 0) < POMP_INF));  // sorted: empty iff slot 0 is
        else
          top.offer_each(d, c.id, lane);
      }
    }
  };
  // K1: float upper bound of the warp's best distance -> squared prune bound for rows
  auto k1_uniform = [&]() {
    if constexpr (K1) {
      const float f = __double2float_ru(bd);  // +inf stays +inf; >= bd
      const unsigned u = __reduce_min_sync(FULL_MASK, __float_as_uint(f));
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
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(FULL_MASK, rd, o);
        const int oi = __shfl_xor_sync(FULL_MASK, ri, o);
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
  for (long long b = g.n_indexed; b < n; b += 32) {
    WCand<D> c;
    c.fetch(b + lane < n, pts, nullptr, b + lane);
    offer(c);
  }
  if (n > g.n_indexed) {
    k1_uniform();
    reduce_k();
  }

  // ---- shells
  WarpShell<D> sh;
  int w0 = 0, w1 = 0;    // sorted positions offered by the dense-mode window (never again)
  bool windowed = false;
  double R = fmin(r0, kd);  // kd = inf while fewer than k points were seen
  double tq_prev = -1.0;
  while (true) {
    const double tq = tsq_of(R);
    sh.set(g, qq, R, tab, lane);
    unsigned nrows = 1;  // < 2^31: the launcher requires n_cells < 2^31 - 64
#pragma unroll
    for (int a = 1; a < D; a++) nrows *= (unsigned)(sh.hi[a] - sh.lo[a] + 1);
    bool restart = false;
    for (unsigned rb = 0; rb < nrows; rb += 32) {
      const unsigned j = rb + lane;
      int s1 = 0, n1 = 0, s2 = 0, n2 = 0;
      if (j < nrows) {
        long long base = 0;
        double gs[D];
        gs[0] = 0.0;
        RowDecode<D, 1>::run(g, sh, qq, home, tab, j, base, gs);
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
      // concatenate the runs of the 32 rows
      const int nl = n1 + n2;
      int incl = nl;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
      }
      const int total = __shfl_sync(FULL_MASK, incl, 31);
      const int excl = incl - nl;
      if (rb == 0 && tq_prev < 0.0 && !windowed && total > WQ_DENSE) {
        // a cluster: the k-th distance among 32 neighbours in sorted order bounds the radius much better
        long long hcell = 0;
#pragma unroll
        for (int a = 0; a < D; a++) hcell += (long long)home[a] * g.stride[a];
        const int hs = __ldg(g.cell_start + hcell), he = __ldg(g.cell_start + hcell + 1);
        const int center = (hs + he) >> 1;
        w0 = max(0, min(center - 16, g.n_indexed - 32));
        w1 = min(g.n_indexed, w0 + 32);
        WCand<D> c;
        c.fetch(w0 + lane < w1, g.spts, g.sidx, w0 + lane);
        offer(c);
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
        for (int step = 16; step > 0; step >>= 1) {
          const int v = __shfl_sync(FULL_MASK, incl, own + step - 1);
          if (v <= item) own += step;
        }
        own = min(own, 31);
        const int os1 = __shfl_sync(FULL_MASK, s1, own), on1 = __shfl_sync(FULL_MASK, n1, own);
        const int os2 = __shfl_sync(FULL_MASK, s2, own), oex = __shfl_sync(FULL_MASK, excl, own);
        const int local = item - oex;
        const int pos = local < on1 ? os1 + local : os2 + (local - on1);
        cv = item < total && !(pos >= w0 && pos < w1);
        return pos;
      };
      // two slabs of 32 candidates in flight
      for (int e = 0; e < total; e += 64) {
        WCand<D> c0, c1;
        bool v0, v1;
        const int p0 = locate(e + lane, v0);
        c0.fetch(v0, g.spts, g.sidx, p0);
        const bool second = e + 32 < total;
        if (second) {
          const int p1 = locate(e + 32 + lane, v1);
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
    if (lane == 0) {
      out_idx[qi * k] = kd < POMP_INF ? (int64_t)ki : -1;
      out_dist[qi * k] = kd;
      if (out_cnt) out_cnt[qi] = kd < POMP_INF ? 1 : 0;
    }
    for (int j = 1 + lane; j < k; j += 32) {  // k > 1 never comes here (host dispatch); defensive
      out_idx[qi * k + j] = -1;
      out_dist[qi * k + j] = POMP_INF;
    }
  } else {
    const bool fin = top.d < POMP_INF && lane < k;
    if (lane < k) {
      out_idx[qi * k + lane] = fin ? (int64_t)top.i : -1;
      out_dist[qi * k + lane] = top.d;
    }
    const unsigned fm = __ballot_sync(FULL_MASK, fin);
    if (out_cnt && lane == 0) out_cnt[qi] = __popc(fm);
  }
}

}  // namespace pomp
