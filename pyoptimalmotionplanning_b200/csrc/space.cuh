// space.cuh -- device-side feasibility test shared by the feasibility, edge-check, propagation and
// fused extend kernels.
//
// Semantics (reference): BoxSet.contains (pomp/spaces/sets.py:177-182) on every coordinate,
// then "outside every obstacle" (example_problems/geometric.py:56-60) with closed boxes and
// closed discs (geometric.py:15-16).  NaN coordinates follow the Python comparisons: they pass a
// box test and fail a disc test.
//
// Obstacles are bucketed into a uniform 2-D grid built on the host with the SAME monotone cell
// function the device uses, so the grid only prunes candidates and never changes a boolean.
#pragma once
#include "common.cuh"

namespace pomp {

struct SpaceDev {
  int dim, od0, od1, n_boxes, n_circles;
  int use_grid, gnx, gny;
  double lo[POMP_MAX_DIM], hi[POMP_MAX_DIM];
  double glo0, glo1, ghi0, ghi1, ginv0, ginv1;  // obstacle grid domain = bounding box of all obstacles
  const double* boxes;    // n_boxes x 4
  const double* circles;  // n_circles x 3
  const int* cell_start;  // gnx*gny + 1
  const int* cell_items;  // obstacle ids: boxes first, then n_boxes + circle id
  const double* cell_geom;  // the same items with their geometry inline, 4 doubles each: box (x0,y0,x1,y1) or
                            // circle (cx,cy,r,NaN) -- one dependent load less per candidate than id -> table
  // fine occupancy map over the same domain, 2 bits per cell (16 cells per word): OBS_EMPTY -- no obstacle touches
  // the cell; OBS_FULL -- the cell lies inside one box (every point of it is infeasible); OBS_PART -- look at the
  // bucket.  Most point tests end here with ONE load from a table of ~1 MB (the bucket path is two dependent loads).
  const unsigned* fine;     // NULL: no map
  int fnx, fny;
  double finv0, finv1;
  int* overflow;            // mapped pinned flag: an edge asked for more than POMP_MAX_EDGE_SAMPLES samples
};

constexpr int OBS_EMPTY = 0, OBS_FULL = 1, OBS_PART = 2;

__host__ __device__ __forceinline__ int obs_cell(double x, double lo, double inv, int n) {
  double t = (x - lo) * inv;
  if (t >= (double)n) return n - 1;
  return t > 0.0 ? (int)t : 0;
}

__device__ __forceinline__ bool in_box(const double* __restrict__ b, double px, double py) {
  // BoxSet.contains: outside iff (xi < a or xi > b) on some coordinate
  if (px < b[0] || px > b[2]) return false;
  if (py < b[1] || py > b[3]) return false;
  return true;
}

__device__ __forceinline__ bool in_circle(const double* __restrict__ c, double px, double py) {
  double dx = px - c[0], dy = py - c[1];
  double s = 0.0;
  s = s + dx * dx;
  s = s + dy * dy;
  return sqrt(s) <= c[2];
}

// bucket of the coarse obstacle grid that holds (px, py): its item range, then the scan (two steps, so that a caller
// can put the range loads of several points in flight together)
__device__ __forceinline__ void bucket_range(const SpaceDev& s, double px, double py, int& b, int& e) {
  int cx = obs_cell(px, s.glo0, s.ginv0, s.gnx), cy = obs_cell(py, s.glo1, s.ginv1, s.gny);
  int c = cy * s.gnx + cx;
  b = __ldg(s.cell_start + c);
  e = __ldg(s.cell_start + c + 1);
}
__device__ __forceinline__ bool bucket_scan(const SpaceDev& s, int b, int e, double px, double py) {
  const double2* __restrict__ G = reinterpret_cast<const double2*>(s.cell_geom);
  for (int it = b; it < e; it++) {
    const double2 g0 = __ldg(G + 2 * it), g1 = __ldg(G + 2 * it + 1);
    if (isnan(g1.y)) {  // circle (centre, radius): geometry.py:15-16
      const double c[3] = {g0.x, g0.y, g1.x};
      if (in_circle(c, px, py)) return true;
    } else {
      const double bx[4] = {g0.x, g0.y, g1.x, g1.y};
      if (in_box(bx, px, py)) return true;
    }
  }
  return false;
}
__device__ __forceinline__ bool hits_bucket(const SpaceDev& s, double px, double py) {
  int b, e;
  bucket_range(s, px, py, b, e);
  return bucket_scan(s, b, e, px, py);
}

// classification of a point by the fine map: OBS_EMPTY / OBS_FULL decide the test, OBS_PART needs hits_bucket().
// Only for spaces with use_grid and non-NaN coordinates (callers check); without a map everything is OBS_PART.
__device__ __forceinline__ int obs_class(const SpaceDev& s, double px, double py) {
  if (px < s.glo0 || px > s.ghi0 || py < s.glo1 || py > s.ghi1) return OBS_EMPTY;  // outside every obstacle's box
  if (!s.fine) return OBS_PART;
  const int cx = obs_cell(px, s.glo0, s.finv0, s.fnx), cy = obs_cell(py, s.glo1, s.finv1, s.fny);
  const long long c = (long long)cy * s.fnx + cx;
  return (int)((__ldg(s.fine + (c >> 4)) >> ((int)(c & 15) * 2)) & 3u);
}

__device__ __forceinline__ bool hits_obstacle(const SpaceDev& s, double px, double py) {
  if (s.n_boxes + s.n_circles == 0) return false;
  bool use_grid = s.use_grid && !(isnan(px) || isnan(py));
  if (!use_grid) {
    for (int o = 0; o < s.n_boxes; o++)
      if (in_box(s.boxes + 4 * o, px, py)) return true;
    for (int o = 0; o < s.n_circles; o++)
      if (in_circle(s.circles + 3 * o, px, py)) return true;
    return false;
  }
  const int cls = obs_class(s, px, py);
  if (cls == OBS_EMPTY) return false;
  if (cls == OBS_FULL) return true;
  return hits_bucket(s, px, py);
}

__device__ __forceinline__ bool in_bounds(const SpaceDev& s, const double* x) {
  for (int i = 0; i < s.dim; i++)
    if (x[i] < s.lo[i] || x[i] > s.hi[i]) return false;
  return true;
}

__device__ __forceinline__ bool feasible_pt(const SpaceDev& s, const double* x) {
  for (int i = 0; i < s.dim; i++)
    if (x[i] < s.lo[i] || x[i] > s.hi[i]) return false;
  return !hits_obstacle(s, x[s.od0], x[s.od1]);
}

// EpsilonEdgeChecker sample count (edgechecker.py:21-22): k = int(ceil(l / resolution)).
// More than POMP_MAX_EDGE_SAMPLES samples (a tiny resolution, or an infinite / huge edge in a space
// with unbounded bounds) would keep one thread busy until the watchdog fires; the reference raises
// OverflowError or never returns there.  Such an edge is reported infeasible, returns -1 here and
// raises the space's overflow flag (mapped pinned memory: the *_h entry points turn it into
// POMP_ERR_INVALID, device-pointer callers read it with pomp_space_get_stat("edge_overflow")).
constexpr double POMP_MAX_EDGE_SAMPLES = 1e7;
__device__ __forceinline__ long long edge_samples(const SpaceDev& s, double l, double res) {
  double c = ceil(l / res);
  if (!(c > 0.0)) return 0;  // also NaN
  if (c > POMP_MAX_EDGE_SAMPLES) {
    if (s.overflow) *s.overflow = 1;
    return -1;
  }
  return (long long)c;
}

}  // namespace pomp

constexpr size_t POMP_MAP_BYTES = 16384;

struct pomp_space {
  int device = 0;
  pomp::SpaceDev dev;
  double* d_boxes = nullptr;
  double* d_circles = nullptr;
  int* d_cell_start = nullptr;
  int* d_cell_items = nullptr;
  double* d_cell_geom = nullptr;
  unsigned* d_fine = nullptr;
  cudaStream_t stream = nullptr;
  pomp::DevBuf<double> in_a, in_b;
  pomp::DevBuf<int32_t> in_i;
  pomp::DevBuf<uint8_t> out_ok;
  // scratch of pomp_roadmap_build (roadmap.cu)
  pomp::DevBuf<int64_t> rm_kidx;
  pomp::DevBuf<double> rm_kdist;
  pomp::DevBuf<int32_t> rm_ij, rm_out;
  pomp::DevBuf<uint8_t> rm_keep, rm_ok, rm_node_ok;
  pomp::DevBuf<int> rm_flag, rm_pos;
  pomp::DevBuf<long long> rm_scan;
  int* overflow_h = nullptr;  // host alias of dev.overflow
  unsigned char* map_h = nullptr;  // mapped pinned scratch (POMP_MAP_BYTES) of the single-edge host calls, host / device alias
  unsigned char* map_d = nullptr;
};

// after a stream synchronisation: turns a raised overflow flag into POMP_ERR_INVALID (and clears it)
inline int space_overflow_check(pomp_space* sp) {
  if (sp && sp->overflow_h && *sp->overflow_h) {
    *sp->overflow_h = 0;
    pomp::set_error("an edge asks for more than %.0f samples (length / resolution)", pomp::POMP_MAX_EDGE_SAMPLES);
    return POMP_ERR_INVALID;
  }
  return POMP_OK;
}
