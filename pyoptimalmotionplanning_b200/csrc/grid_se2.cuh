// grid_se2.cuh -- k-NN for the SE(2) product metric (planar position x heading), compile-time layout.
//
// MultiGeodesicSpace.distance over (R^2, SO(2)) with weights (w0, w1) -- the metric of the Dubins-car
// problems (pomp/spaces/geodesicspace.py:45-52, so2space.py:9, example_problems/dubins.py):
//     d = sqrt( (sqrt(ex^2 + ey^2))^2 * w0 + (|so2.diff|)^2 * w1 )
// The generic kernel (grid_knn_kernel<0,K,false>) evaluates it through the run-time metric descriptor and
// walks the grid with dynamically indexed per-level state: 7 of 32 lanes active, 30 % instruction-cache
// stalls, 13 % of the samples on local memory (profiles/r1_grid_knn_generic_SE2_k1_details.csv).  Here
// the layout is fixed, so the same arithmetic (same operations, same order: results are bit-identical
// to metric_eval) is straight-line code, the three grid levels are nested loops with their state in
// registers, and the list is the register-resident PosTopK of the Euclidean kernels.
//
// Grid: axes 0, 1 = x, y; axis 2 = heading, periodic (cells wrap, gap = angular distance to the cell).
// Pruning uses the SAME combine on per-axis gaps, on the squared value against fl(T^2)(1+2^-50), as
// the Euclidean kernels do (every operation is monotone in each |difference|).
#pragma once
#include "grid_fast.cuh"

namespace pomp {

// squared SE(2) distance of stored point p to query (qx, qy, qn = so2.normalize(q_theta)), operation for
// operation what metric_diffs + metric_combine compute before the final sqrt
__device__ __forceinline__ double se2_sq(double px, double py, double pth, double qx, double qy, double qn, double w0,
                                         double w1) {
  double ex = px - qx, ey = py - qy;
  double s = 0.0;
  s = s + ex * ex;
  s = s + ey * ey;
  double dc = sqrt(s);
  double res = 0.0;
  res += (dc * dc) * w0;
  double d = so2_normalize(pth) - qn;  // so2.diff (klampt/so2.py:30-37)
  if (d < -POMP_PI) d = d + POMP_TWOPI;
  else if (d > POMP_PI) d = d - POMP_TWOPI;
  double da = fabs(d);
  res += (da * da) * w1;
  return res;
}

// the same combine on per-axis gaps: a lower bound of se2_sq over everything bucketed behind those gaps
__device__ __forceinline__ double se2_bound_sq(double gx, double gy, double gth, double w0, double w1) {
  double s = 0.0;
  s = s + gx * gx;
  s = s + gy * gy;
  double dc = sqrt(s);
  double res = 0.0;
  res += (dc * dc) * w0;
  res += (gth * gth) * w1;
  return res;
}

template <class L>
__device__ __forceinline__ void se2_scan(const GridDev& g, const uint8_t* __restrict__ skip, double qx, double qy,
                                         double qn, double w0, double w1, L& top, int s, int t) {
  if (s >= t) return;
  const double* __restrict__ P = g.spts;
  double nx = __ldg(P + 3LL * s), ny = __ldg(P + 3LL * s + 1), nt = __ldg(P + 3LL * s + 2);
  for (int pos = s; pos < t; pos++) {
    const double px = nx, py = ny, pt = nt;
    if (pos + 1 < t) {  // look-ahead: the next point is in flight while this one is evaluated
      nx = __ldg(P + 3LL * (pos + 1));
      ny = __ldg(P + 3LL * (pos + 1) + 1);
      nt = __ldg(P + 3LL * (pos + 1) + 2);
    }
    if (skip && skip[__ldg(g.sidx + pos)]) continue;
    top.offer_sq(se2_sq(px, py, pt, qx, qy, qn, w0, w1), pos);
  }
}

template <int K>
__global__ void __launch_bounds__(128)
grid_knn_se2_kernel(GridDev g, double w0, double w1, const double* __restrict__ pts, long long n,
                    const uint8_t* __restrict__ skip, const double* __restrict__ q, long long nq,
                    const int* __restrict__ qperm, int k, int64_t* __restrict__ out_idx,
                    double* __restrict__ out_dist, int32_t* __restrict__ out_cnt) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  const long long qi = qperm ? (long long)__ldg(qperm + t) : t;
  const double qx = q[qi * 3], qy = q[qi * 3 + 1], qth = q[qi * 3 + 2];
  const double qn = so2_normalize(qth);
  PosTopK<K> top;
  top.init(IdxResolver{g.sidx, g.n_indexed});
  const int hx = grid_cell(g, 0, qx), hy = grid_cell(g, 1, qy), hz = grid_cell(g, 2, qn);
  // phase A: the home cell, widened to at least K sorted-order neighbours, bounds the k-th distance
  const long long home = (long long)hx + (long long)hy * g.stride[1] + (long long)hz * g.stride[2];
  const int hs = __ldg(g.cell_start + home), he = __ldg(g.cell_start + home + 1);
  const int center = (hs + he) >> 1;
  int w0p = min(hs, max(0, center - K)), w1p = max(he, min(g.n_indexed, center + K));
  se2_scan(g, skip, qx, qy, qn, w0, w1, top, w0p, w1p);
  while (!top.full() && (w0p > 0 || w1p < g.n_indexed)) {
    int len = max(w1p - w0p, 1);
    if (w0p > 0) {
      int nw = max(0, w0p - len);
      se2_scan(g, skip, qx, qy, qn, w0, w1, top, nw, w0p);
      w0p = nw;
    } else {
      int nw = min(g.n_indexed, w1p + len);
      se2_scan(g, skip, qx, qy, qn, w0, w1, top, w1p, nw);
      w1p = nw;
    }
  }
  // phase B: heading cells centre-out with wrap-around, rows centre-out, x as one contiguous run
  const int nz = g.ncell[2], ny = g.ncell[1];
  bool zlo = false, zhi = false;
  for (int tz = 0; !(zlo && zhi); tz++) {
    if (tz >= nz) break;  // the first nz offsets of the sequence are distinct modulo nz: every cell once
    const int oz = tz == 0 ? 0 : ((tz & 1) ? -((tz + 1) >> 1) : (tz >> 1));
    if (oz < 0 ? zlo : zhi) continue;
    int cz = hz + oz;
    if (cz < 0) cz += nz;
    else if (cz >= nz) cz -= nz;
    const double gz = grid_gap_p(g, 2, cz, qth);
    if (se2_bound_sq(0.0, 0.0, gz, w0, w1) > top.tsq_up()) {
      if (oz < 0) zlo = true;
      else zhi = true;
      continue;
    }
    bool ylo = false, yhi = false;
    for (int ty = 0; !(ylo && yhi); ty++) {
      const int oy = ty == 0 ? 0 : ((ty & 1) ? -((ty + 1) >> 1) : (ty >> 1));
      if (oy < 0 ? ylo : yhi) continue;
      const int cy = hy + oy;
      if (cy < 0) {
        ylo = true;
        continue;
      }
      if (cy >= ny) {
        yhi = true;
        continue;
      }
      const double gy = grid_gap(g, 1, cy, qy);
      if (se2_bound_sq(0.0, gy, gz, w0, w1) > top.tsq_up()) {
        if (oy < 0) ylo = true;
        else yhi = true;
        continue;
      }
      int lo, hi;
      grid_range(g, 0, qx, top.threshold(), &lo, &hi);
      while (lo < hi && se2_bound_sq(grid_gap(g, 0, lo, qx), gy, gz, w0, w1) > top.tsq_up()) lo++;
      while (hi > lo && se2_bound_sq(grid_gap(g, 0, hi, qx), gy, gz, w0, w1) > top.tsq_up()) hi--;
      const long long base = (long long)cy * g.stride[1] + (long long)cz * g.stride[2];
      const int s = __ldg(g.cell_start + base + lo), e = __ldg(g.cell_start + base + hi + 1);
      if (s < w1p && e > w0p) {  // positions [w0p, w1p) were offered in phase A
        se2_scan(g, skip, qx, qy, qn, w0, w1, top, s, min(e, w0p));
        se2_scan(g, skip, qx, qy, qn, w0, w1, top, max(s, w1p), e);
      } else {
        se2_scan(g, skip, qx, qy, qn, w0, w1, top, s, e);
      }
    }
  }
  // points appended after the last index build
  for (long long i = g.n_indexed; i < n; i++) {
    if (skip && skip[i]) continue;
    top.offer_sq(se2_sq(__ldg(pts + i * 3), __ldg(pts + i * 3 + 1), __ldg(pts + i * 3 + 2), qx, qy, qn, w0, w1), (int)i);
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

}  // namespace pomp
