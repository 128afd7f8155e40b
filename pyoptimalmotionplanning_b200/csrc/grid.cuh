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
  int periodic[POMP_MAX_DIM];   // axis is an SO(2) angle: bucketed by so2_normalize(x) over [0, 2 pi), wraps around
};

// blocked ("strip") copy of a 2-D grid index for the streaming nearest-neighbour kernel, see strip2d.cuh
struct StripDev {
  int W, H;            // cells of the underlying grid
  int nstrips, nty;    // strips along x, tiles per strip
  int ntiles;
  const double2* spts2;      // replicated, strip-major points
  const int* sidx2;          // insertion index per slot
  const int2* tile_info;     // (first slot, slots) of tile + halo
  const uint16_t* tile_cs;   // ST_CS_STRIDE offsets per tile
  long long nslots;
};

// monotone in x by construction (fl subtraction, fl multiply, truncation, clamping)
__device__ __forceinline__ int grid_cell(const GridDev& g, int a, double x) {
  double t = (x - g.lo[a]) * g.inv_h[a];
  int n = g.ncell[a];
  if (t >= (double)n) return n - 1;
  return t > 0.0 ? (int)t : 0;  // NaN -> 0
}

// periodic-aware variant for the generic paths (the Euclidean kernels never see a periodic axis)
__device__ __forceinline__ int grid_cell_p(const GridDev& g, int a, double x) {
  return grid_cell(g, a, g.periodic[a] ? so2_normalize(x) : x);
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

// periodic axis: lower bound of |so2.diff| between x and any angle bucketed into cell c
__device__ __forceinline__ double grid_gap_p(const GridDev& g, int a, int c, double x) {
  if (!g.periodic[a]) return grid_gap(g, a, c, x);
  const double phi = so2_normalize(x);
  const double clo = (double)c * g.h[a], chi = (c == g.ncell[a] - 1) ? POMP_TWOPI : (double)(c + 1) * g.h[a];
  if (phi >= clo && phi <= chi) return 0.0;
  double d1 = fabs(phi - clo), d2 = fabs(phi - chi);
  d1 = fmin(d1, POMP_TWOPI - d1);
  d2 = fmin(d2, POMP_TWOPI - d2);
  const double gap = fmin(d1, d2) - g.slack[a];
  return gap > 0.0 ? gap : 0.0;
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
  // Axes G-1 .. 1 are walked CENTRE-OUT from the query's own cell (offsets 0, -1, +1, -2, ...):
  // near cells come first, so the threshold has shrunk to the true neighbour distance long before
  // the far cells are looked at, and a direction is closed for good at the first cell whose gap
  // bound exceeds the threshold (gaps grow with the offset, the threshold only shrinks).  Metrics
  // with a non-gridded coordinate (SO(2) components) start from a poor bound: walking the rows of
  // the initial range end to end cost 26 s for 2^20 queries against 2^24 SE(2) states; this order
  // needs about the true ball.  Axis 0 is a contiguous run of the sorted points, as before.
  int home[POMP_MAX_DIM], cur[POMP_MAX_DIM], step[POMP_MAX_DIM];
  bool dlo[POMP_MAX_DIM], dhi[POMP_MAX_DIM];
  double e[POMP_MAX_DIM];
#pragma unroll
  for (int i = 0; i < POMP_MAX_DIM; i++) e[i] = 0.0;
  for (int a = 1; a < G; a++) home[a] = grid_cell_p(g, a, q[g.gdim[a]]);
  const int d0 = g.gdim[0];
  int lvl = G - 1;
  if (lvl >= 1) {
    step[lvl] = 0;
    dlo[lvl] = dhi[lvl] = false;
  }
  while (true) {
    if (G > 1) {
      // next open cell on level lvl (levels below it restart at their centre)
      bool found = false;
      while (!(dlo[lvl] && dhi[lvl])) {
        const int tt = step[lvl]++;
        const int off = tt == 0 ? 0 : ((tt & 1) ? -((tt + 1) >> 1) : (tt >> 1));
        if (off < 0 ? dlo[lvl] : dhi[lvl]) continue;
        int c = home[lvl] + off;
        if (g.periodic[lvl]) {
          // the first ncell offsets of the sequence are distinct modulo ncell: every cell once
          if (tt >= g.ncell[lvl]) {
            dlo[lvl] = dhi[lvl] = true;
            continue;
          }
          if (c < 0) c += g.ncell[lvl];
          else if (c >= g.ncell[lvl]) c -= g.ncell[lvl];
        } else {
          if (c < 0) {
            dlo[lvl] = true;
            continue;
          }
          if (c >= g.ncell[lvl]) {
            dhi[lvl] = true;
            continue;
          }
        }
        for (int b2 = 0; b2 < lvl; b2++) e[g.gdim[b2]] = 0.0;
        e[g.gdim[lvl]] = grid_gap_p(g, lvl, c, q[g.gdim[lvl]]);
        if (gap_bound<DA, EUCLID>(m, e) > coll.threshold()) {
          if (off < 0) dlo[lvl] = true;
          else dhi[lvl] = true;
          continue;
        }
        cur[lvl] = c;
        found = true;
        break;
      }
      if (!found) {
        if (++lvl >= G) break;
        continue;
      }
      if (lvl > 1) {
        lvl--;
        step[lvl] = 0;
        dlo[lvl] = dhi[lvl] = false;
        continue;
      }
    }
    // one row along axis 0
    long long base = 0;
    for (int a = 1; a < G; a++) base += (long long)cur[a] * g.stride[a];
    e[d0] = 0.0;
    const double T = coll.threshold();
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
    e[d0] = 0.0;
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
    if (G == 1) break;
  }
}

}  // namespace pomp
