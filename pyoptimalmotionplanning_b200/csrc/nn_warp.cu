// nn_warp.cu -- launcher of the warp-per-query k-NN kernel (grid_warp.cuh).
#include "common.cuh"
#include "grid.cuh"
#include "grid_warp.cuh"
#include "nn_handle.cuh"

using namespace pomp;

// volume of the unit ball in D dimensions
static double unit_ball(int D) {
  switch (D) {
    case 1: return 2.0;
    case 2: return 3.141592653589793;
    case 3: return 4.1887902047863905;
    case 4: return 4.934802200544679;
    case 5: return 5.263789013914325;
    default: return 5.167712780049970;
  }
}

bool pomp::warp_knn_eligible(const pomp_nn* h, int k) {
  if (h->warp_search == 0 || k > 32) return false;
  if (!(h->D == 2 || h->D == 3 || h->D == 4 || h->D == 6)) return false;
  if (h->grid.n_cells >= 0x7fffffffLL - 64) return false;
  if (h->warp_search >= 2) return true;
  // measured on B200 (2^20 queries vs 2^24 points, thread-per-query -> this kernel; profiles/knn_sweep.py):
  //   3-D  k = 4: 1.00 -> 1.41 ms   k = 8: 1.44 -> 1.80 ms   k = 16: 4.61 -> 2.27 ms (16 lanes per query)
  //   4-D  k = 4: 2.01 -> 2.60 ms   k = 8: 2.97 -> 3.05 ms   k = 16: 7.68 -> 4.04 ms
  //   6-D  k = 1: 9.2 -> 7.2 ms   k = 4: 23.3 -> 11.8 ms   k = 8: 15.0 -> 15.7 ms   k = 16: 29.2 -> 21.8 ms (4 points
  //        per cell; 6.9 / 10.8 / 14.2 / 19.4 ms at the 8 the index now uses in 6-D)
  //   2-D  k = 4 / 8 / 16: 0.32 / 0.54 / 1.20 -> 0.68 / 0.79 / 1.18 ms: the thread kernels keep 2-D
  if (h->D == 2) return false;
  if (h->D == 6) return true;  // with the 8 points per cell the index is built with there (k = 8: 14.2 vs 20.3 ms)
  return k > 8;
}

int pomp::launch_grid_knn_warp(pomp_nn* h, const double* q, const double* qsorted, int64_t nq, const int* qperm, int k, int64_t* idx,
                               double* dist, int32_t* cnt, cudaStream_t st) {
  const GridDev& g = h->grid;
  const int D = h->D;
  // first shell: the radius expected to hold `need` points at the mean density of the bounding box
  double vol = 1.0, diag = 0.0, hmin = 1e300;
  for (int a = 0; a < D; a++) {
    const double ext = g.hi[a] - g.lo[a];
    vol *= ext > 0.0 ? ext : 1.0;
    diag += ext > 0.0 ? ext : 0.0;
    if (g.h[a] > 0.0) hmin = std::min(hmin, g.h[a]);
  }
  if (!(hmin < 1e300)) hmin = 1.0;
  const double need = h->warp_need > 0 ? h->warp_need : (k == 1 ? 3.0 : (double)k + 1.5 * sqrt((double)k) + 1.0);
  double r0 = pow(need * vol / (std::max(1, g.n_indexed) * unit_ball(D)), 1.0 / D);
  if (!(r0 > 0.0) || !std::isfinite(r0)) r0 = hmin;
  r0 = std::max(r0, 1e-3 * hmin);
  const double rmax = 4.0 * diag + 4.0 * r0;
  const uint8_t* skip = h->skip_ptr();
  // lanes per query and keys per lane.  A group of W lanes holding k <= W keys (one per lane) merges a slab in bulk
  // (bitonic network over the group), and 32 / W queries share every instruction of the bookkeeping -- measured
  // (narrowest fitting group vs 32 lanes): 2-D k = 16: 1.18 vs 1.53 ms, 3-D k = 16: 2.27 vs 2.43 ms, but 4-D k = 16:
  // 4.82 vs 4.04 ms, 4-D k = 8: 4.39 vs 3.05 ms, 6-D k = 4 / 8 / 16: 16.9 / 22.6 / 26.1 vs 11.8 / 15.7 / 21.8 ms (many
  // rows and candidates per query: the groups of a warp diverge), so 16 lanes only for k in 9..16 up to 3-D.
  // "warp_lanes" = 8 / 16 / 32 forces a width (keys per lane = 32 / W when k does not fit: single insertions only,
  // measured slow).  k = 1, W = 8 / 16 / 32: 3-D 0.85 / 1.07 / 1.47 ms, 4-D 1.57 / 1.69 / 1.99 ms, 6-D 9.7 / 8.5 / 7.2 ms.
  int W = (int)h->warp_lanes;
  if (!(W == 8 || W == 16 || W == 32)) W = (D <= 3 && k > 8 && k <= 16) ? 16 : 32;
#define WK(DD, K1, WW, KLL)                                                                                           \
  LAUNCH((grid_knn_warp_kernel<DD, K1, WW, KLL>), (int)((nq + (WQ_WARPS * 32 / WW) - 1) / (WQ_WARPS * 32 / WW)),      \
         WQ_WARPS * 32, 0, st, g, h->d_pts, (long long)h->n_dev, skip, q, qsorted, (long long)nq, qperm, k, r0,     \
         rmax, idx, dist, cnt)
#define WKW(DD)                               \
  do {                                        \
    if (k == 1) {                             \
      if (W == 8) WK(DD, true, 8, 1);         \
      else if (W == 16) WK(DD, true, 16, 1);  \
      else WK(DD, true, 32, 1);               \
    } else if (W == 8) {                      \
      if (k <= 8) WK(DD, false, 8, 1);        \
      else WK(DD, false, 8, 4);               \
    } else if (W == 16) {                     \
      if (k <= 16) WK(DD, false, 16, 1);      \
      else WK(DD, false, 16, 2);              \
    } else WK(DD, false, 32, 1);              \
  } while (0)
  prof_begin(h, st);
  if (D == 2) WKW(2);
  else if (D == 3) WKW(3);
  else if (D == 4) WKW(4);
  else WKW(6);
#undef WKW
#undef WK
  prof_end(h, st);
  CHECK_LAUNCH();
  return POMP_OK;
}
