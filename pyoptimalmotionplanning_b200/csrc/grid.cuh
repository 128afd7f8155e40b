// grid.cuh -- uniform-grid bucket index and its exact search.
//
// Replaces the reference's kd-tree (pomp/structures/kdtree.py) for large node sets.  Points are
// bucketed into axis-aligned cells over the coordinates the metric is monotone in ("grid axes");
// within the index they are stored sorted by cell id (row-major, axis 0 fastest), so every row of
// cells along axis 0 is ONE contiguous run of points.  A query prunes rows/cells with a lower
// bound that evaluates the metric's own combine() on per-axis gaps -- rigorous in floating point
// because every operation involved is monotone -- so results equal the brute-force scan
// (nearestneighbors.py:88-141) bit for bit, ties included.
#pragma once
#include "common.cuh"

namespace pomp {

struct GridDev {
  int G;                        // number of grid axes (>= 1)
  int gdim[POMP_MAX_DIM];       // coordinate index of each axis, increasing
  int ncell[POMP_MAX_DIM];
  long long stride[POMP_MAX_DIM];  // linear cell id = sum c_a * stride_a, stride_0 = 1
  double lo[POMP_MAX_DIM];        // bounding box of the indexed points on each axis
  double hi[POMP_MAX_DIM];
  double inv_h[POMP_MAX_DIM];
  double h[POMP_MAX_DIM];
  double slack[POMP_MAX_DIM];   // absolute widening of cell bounds (covers cell-assignment rounding)
  double rscale[POMP_MAX_DIM];  // |diff_axis| <= T * rscale for every candidate with distance <= T
  const int* cell_start;        // n_cells + 1
  const double* spts;           // n_indexed x D, sorted by cell
  const int* sidx;              // insertion index of each sorted slot
  int n_indexed;
  long long n_cells;
};

// monotone in x by construction (fl subtraction, fl multiply, truncation, clamping)
__device__ __forceinline__ int grid_cell(const GridDev& g, int a, double x) {
  double t = (x - g.lo[a]) * g.inv_h[a];
  int n = g.ncell[a];
  if (t >= (double)n) return n - 1;
  return t > 0.0 ? (int)t : 0;  // NaN -> 0
}

// conservative bounds of the coordinates bucketed into cell c of axis a (all indexed points lie
// inside [lo, hi], so the outermost cells end at the bounding box)
__device__ __forceinline__ double grid_gap(const GridDev& g, int a, int c, double x) {
  double clo = (c == 0) ? g.lo[a] : (g.lo[a] + (double)c * g.h[a]);
  double chi = (c == g.ncell[a] - 1) ? g.hi[a] : (g.lo[a] + (double)(c + 1) * g.h[a]);
  clo -= g.slack[a];
  chi += g.slack[a];
  double gap = 0.0;
  if (x < clo) gap = clo - x;
  else if (x > chi) gap = x - chi;
  return gap;
}

__device__ __forceinline__ void grid_range(const GridDev& g, int a, double x, double T, int* clo, int* chi) {
  if (!(T < POMP_INF)) {
    *clo = 0;
    *chi = g.ncell[a] - 1;
    return;
  }
  double R = T * g.rscale[a] + g.slack[a];
  *clo = grid_cell(g, a, x - R);
  *chi = grid_cell(g, a, x + R);
}

template <int D>
__device__ __forceinline__ void load_point(const double* __restrict__ base, long long pos, double* p) {
  const double* src = base + pos * D;
  if (D % 2 == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int j = 0; j < D / 2; j++) {
      double2 v = __ldg(s2 + j);
      p[2 * j] = v.x;
      p[2 * j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; j++) p[j] = __ldg(src + j);
  }
}

// distance of stored point p to query q
template <int D, bool EUCLID>
__device__ __forceinline__ double point_dist(const pomp_metric_desc& m, const double* p, const double* q) {
  if (EUCLID) return euclid_fixed<D>(p, q);
  return metric_eval(m, p, q);
}

// lower bound of the distance from q to anything whose per-axis |difference| is >= e
template <int D, bool EUCLID>
__device__ __forceinline__ double gap_bound(const pomp_metric_desc& m, const double* e) {
  if (EUCLID) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) s = s + e[i] * e[i];
    return sqrt(s);
  }
  return metric_combine(m, e, m.has_cost ? m.dim - 1 : m.dim);
}

// Visit every indexed point that can have distance <= coll.threshold() (threshold may shrink while
// visiting).  Sorted positions in [w0,w1) are skipped (already offered by the caller).
// Coll: double threshold(); void offer(double d, long long pos)  [pos = sorted slot]
template <int D, bool EUCLID, class Coll>
__device__ __forceinline__ void grid_visit(const GridDev& g, const pomp_metric_desc& m,
                                           const uint8_t* __restrict__ skip, const double* q, Coll& coll,
                                           int w0, int w1) {
  const int G = g.G;
  constexpr int DA = D ? D : POMP_MAX_DIM;  // D = 0: runtime dimension (generic metrics)
  const int dim = D ? D : m.dim;
  (void)dim;
  int clo[POMP_MAX_DIM], chi[POMP_MAX_DIM], cur[POMP_MAX_DIM];
  double e[POMP_MAX_DIM];
#pragma unroll
  for (int i = 0; i < POMP_MAX_DIM; i++) e[i] = 0.0;
  double T = coll.threshold();
  for (int a = 1; a < G; a++) {
    grid_range(g, a, q[g.gdim[a]], T, &clo[a], &chi[a]);
    cur[a] = clo[a];
  }
  const int d0 = g.gdim[0];
  while (true) {
    long long base = 0;
    for (int a = 1; a < G; a++) {
      e[g.gdim[a]] = grid_gap(g, a, cur[a], q[g.gdim[a]]);
      base += (long long)cur[a] * g.stride[a];
    }
    e[d0] = 0.0;
    T = coll.threshold();
    double lb = gap_bound<DA, EUCLID>(m, e);
    if (!(lb > T)) {
      int c0lo, c0hi;
      grid_range(g, 0, q[d0], T, &c0lo, &c0hi);
      // tighten both ends with the exact per-cell bound
      while (c0lo < c0hi) {
        e[d0] = grid_gap(g, 0, c0lo, q[d0]);
        if (gap_bound<DA, EUCLID>(m, e) > T) c0lo++;
        else break;
      }
      while (c0hi > c0lo) {
        e[d0] = grid_gap(g, 0, c0hi, q[d0]);
        if (gap_bound<DA, EUCLID>(m, e) > T) c0hi--;
        else break;
      }
      int s = __ldg(g.cell_start + base + c0lo);
      int t = __ldg(g.cell_start + base + c0hi + 1);
      for (int pos = s; pos < t; pos++) {
        if (pos >= w0 && pos < w1) {
          pos = w1 - 1;
          continue;
        }
        if (skip && skip[__ldg(g.sidx + pos)]) continue;
        double p[DA];
        if (D) load_point<DA>(g.spts, pos, p);
        else
          for (int j = 0; j < dim; j++) p[j] = __ldg(g.spts + (long long)pos * dim + j);
        coll.offer(point_dist<DA, EUCLID>(m, p, q), pos);
      }
    }
    int a = 1;
    for (; a < G; a++) {
      if (++cur[a] <= chi[a]) break;
      cur[a] = clo[a];
    }
    if (a >= G) break;
  }
}

}  // namespace pomp
