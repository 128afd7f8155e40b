// grid_pipe.cuh -- 2-D Euclidean nearest neighbour (k = 1) as a per-thread SOFTWARE PIPELINE.
//
// The flattened block kernel (grid_fast.cuh) is bound by exposed memory latency: each query is a
// chain of dependent round trips (query -> cell_start bounds -> points -> sidx), and the warp sits
// through every one of them.  Here every thread walks a sequence of bucket-sorted queries and
// keeps FOUR of them in flight at different stages, so each load is issued a whole iteration before
// its value is needed:
//
//   iteration i:  A  load the coordinates of query i+3                          (registers)
//                 B  query i+2: home cell -> issue its 6 cell_start loads        (registers)
//                 C  query i+1: cell_start arrived -> runs -> cp.async its candidate points into
//                    the thread's private shared-memory slot (i+1)&1             (LDGSTS, no registers)
//                 D  query i: wait for its cp.async group, scan its candidates from shared memory,
//                    certify, issue the sidx load of the winner
//                 E  query i-1: sidx arrived -> write the result
//
// Exactness: same candidate set ((2R+1)^2 block, R = 1) and the same certificate as
// grid_knn_flat2d_kernel; queries whose block holds more than CAP points or whose certificate fails
// go to the overflow list (second chance in the flat kernel, then the pruned traversal).
#pragma once
#include "grid_fast.cuh"

namespace pomp {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct PipeRuns {   // the three row runs of one query's 3x3 block, centre row first
  int s0, s1, s2;   // first sorted position of each run
  int p1, p2, p3;   // prefix counts: run0 = [0,p1), run1 = [p1,p2), run2 = [p2,p3)
  __device__ __forceinline__ int pos_of(int c) const {
    int pos = s0 + c;
    if (c >= p1) pos = s1 + (c - p1);
    if (c >= p2) pos = s2 + (c - p2);
    return pos;
  }
};

template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
grid_nn_pipe2d_kernel(GridDev g, const double* __restrict__ qs, const int* __restrict__ qperm, long long nq,
                      int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                      int* __restrict__ ovf_list, int* __restrict__ ovf_count) {
  extern __shared__ __align__(16) double2 pipe_slots[];  // [2][CAP][THREADS]
  const int tid = threadIdx.x;
  const long long stride = (long long)gridDim.x * THREADS;
  const long long j0 = (long long)blockIdx.x * THREADS + tid;
  const int W = g.ncell[0], H = g.ncell[1];
  const double2* __restrict__ Q = reinterpret_cast<const double2*>(qs);
  const double2* __restrict__ P = reinterpret_cast<const double2*>(g.spts);
  auto slot = [&](int sl, int c) -> double2* { return pipe_slots + ((size_t)(sl * CAP + c) * THREADS + tid); };

  // ---- stage helpers ---------------------------------------------------------------------------
  // B: home cell + the six cell_start loads of the 3x3 block (rows: centre, below, above)
  auto issue_cs = [&](double2 q, bool valid, int& cx, int& cy, int (&cs)[6]) {
    cx = grid_cell(g, 0, q.x);
    cy = grid_cell(g, 1, q.y);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, W - 1);
#pragma unroll
    for (int r = 0; r < 3; r++) {
      int y = cy + (r == 0 ? 0 : (r == 1 ? -1 : 1));
      bool ok = valid && y >= 0 && y < H;
      long long base = (long long)(ok ? y : 0) * W;
      int a = __ldg(g.cell_start + base + x0), b = __ldg(g.cell_start + base + x1 + 1);
      cs[2 * r] = a;
      cs[2 * r + 1] = ok ? b : a;  // invalid row: empty run
    }
  };
  // C: runs from the arrived cell_start values; async copy of the candidates into the slot
  auto issue_points = [&](const int (&cs)[6], bool valid, int sl, PipeRuns& ru) {
    ru.s0 = cs[0];
    ru.s1 = cs[2];
    ru.s2 = cs[4];
    ru.p1 = valid ? cs[1] - cs[0] : 0;
    ru.p2 = ru.p1 + (valid ? cs[3] - cs[2] : 0);
    ru.p3 = ru.p2 + (valid ? cs[5] - cs[4] : 0);
    const int total = min(ru.p3, CAP);
    for (int c = 0; c < total; c++) cp_async16(slot(sl, c), P + ru.pos_of(c));
    cp_async_commit();
  };

  // ---- pipeline state, indexed by (query number mod 4 / mod 2) with COMPILE-TIME indices only -----
  // (a register that is the destination of a pending load must never be copied: the copy would
  // wait for the load and serialise the pipeline; the loop is unrolled by 4 so that the roles
  // rotate by name instead)
  double2 q[4];
  int qi[4], cx[4], cy[4];
  bool v[4];
  int cs[2][6];
  PipeRuns ru[2];
  int pidx[2], pqi[2];  // winner of query n: its sidx load stays in flight through iteration n+1
  double pd[2];
  bool pv[2];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    q[a] = make_double2(0, 0);
    qi[a] = cx[a] = cy[a] = 0;
    v[a] = false;
  }
#pragma unroll
  for (int a = 0; a < 2; a++) {
    pidx[a] = -1;
    pqi[a] = 0;
    pd[a] = 0.0;
    pv[a] = false;
    ru[a] = PipeRuns{0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 6; r++) cs[a][r] = 0;
  }
  auto load_q = [&](long long j, double2& qq, int& qq_i, bool& vv) {
    vv = j < nq;
    long long jj = vv ? j : 0;
    qq = __ldg(Q + jj);
    qq_i = __ldg(qperm + jj);
  };

  // prologue: queries 0, 1, 2 reach the state iteration 0 expects
  load_q(j0, q[0], qi[0], v[0]);
  load_q(j0 + stride, q[1], qi[1], v[1]);
  load_q(j0 + 2 * stride, q[2], qi[2], v[2]);
  issue_cs(q[0], v[0], cx[0], cy[0], cs[0]);
  issue_points(cs[0], v[0], 0, ru[0]);
  issue_cs(q[1], v[1], cx[1], cy[1], cs[1]);

  // one iteration for query number n with n % 4 == PH; returns false once nothing is left in flight
  auto step = [&](auto ph, long long j) -> bool {
    constexpr int PH = decltype(ph)::value;
    constexpr int N0 = PH & 3, N1 = (PH + 1) & 3, N2 = (PH + 2) & 3, N3 = (PH + 3) & 3;
    constexpr int S0 = PH & 1, S1 = (PH + 1) & 1;
    // A: coordinates of query n+3
    load_q(j + 3 * stride, q[N3], qi[N3], v[N3]);
    // B: query n+2 -> its cell_start loads into buffer (n+2)&1 = S0 (query n's copy was consumed last iteration)
    issue_cs(q[N2], v[N2], cx[N2], cy[N2], cs[S0]);
    // C: query n+1 -> cp.async into slot S1 (query n-1 was scanned in the previous iteration)
    issue_points(cs[S1], v[N1], S1, ru[S1]);
    // D: query n
    cp_async_wait<1>();  // everything but the newest group (query n+1) has landed
    bool w_valid = false;
    int w_pos = -1;
    double w_d = POMP_INF;
    if (v[N0]) {
      const double2 qq = q[N0];
      const int total = ru[S0].p3;
      bool w_ovf = false;
      if (total > CAP) {
        w_ovf = true;
      } else {
        double best_s = POMP_INF;  // squared, for the cheap first cut
        for (int c = 0; c < total; c++) {
          const double2 pvv = *slot(S0, c);
          double dx = pvv.x - qq.x, dy = pvv.y - qq.y;
          double sq = 0.0;
          sq = sq + dx * dx;
          sq = sq + dy * dy;
          if (sq <= best_s * (1.0 + 0x1p-50)) {  // possible winner or tie after the square root
            double d = sqrt(sq);
            int pos = ru[S0].pos_of(c);
            bool better = d < w_d;
            if (!better && d == w_d && w_pos >= 0) better = __ldg(g.sidx + pos) < __ldg(g.sidx + w_pos);
            if (better) {
              w_d = d;
              w_pos = pos;
              best_s = sq;
            }
          }
        }
        // certificate (same as grid_knn_flat2d_kernel, R = 1)
        double margin = POMP_INF;
        int lo_c = cx[N0] - 1, hi_c = cx[N0] + 1;
        if (lo_c > 0) margin = fmin(margin, qq.x - ((g.lo[0] + (double)lo_c * g.h[0]) + g.slack[0]));
        if (hi_c < W - 1) margin = fmin(margin, ((g.lo[0] + (double)(hi_c + 1) * g.h[0]) - g.slack[0]) - qq.x);
        lo_c = cy[N0] - 1;
        hi_c = cy[N0] + 1;
        if (lo_c > 0) margin = fmin(margin, qq.y - ((g.lo[1] + (double)lo_c * g.h[1]) + g.slack[1]));
        if (hi_c < H - 1) margin = fmin(margin, ((g.lo[1] + (double)(hi_c + 1) * g.h[1]) - g.slack[1]) - qq.y);
        bool exact = !(margin < POMP_INF) || (w_d * (1.0 + 1e-9) < margin);
        if (exact) w_valid = true;
        else w_ovf = true;
      }
      if (w_ovf) {
        int o = atomicAdd(ovf_count, 1);
        ovf_list[o] = qi[N0];
      }
    }
    pidx[S0] = (w_valid && w_pos >= 0) ? __ldg(g.sidx + w_pos) : -1;  // consumed one iteration later
    pqi[S0] = qi[N0];
    pd[S0] = w_d;
    pv[S0] = w_valid;
    // E: result of query n-1 (its sidx load was issued one iteration ago)
    if (pv[S1]) {
      bool fin = pidx[S1] >= 0;
      out_idx[pqi[S1]] = fin ? (int64_t)pidx[S1] : -1;
      out_dist[pqi[S1]] = fin ? pd[S1] : POMP_INF;
      if (out_cnt) out_cnt[pqi[S1]] = fin ? 1 : 0;
      pv[S1] = false;
    }
    // queries are valid for a prefix of the sequence only: stop when the next one is invalid and
    // this iteration left no result pending
    return v[N1] || pv[S0];
  };

  for (long long j = j0;; j += 4 * stride) {
    if (!step(std::integral_constant<int, 0>{}, j)) break;
    if (!step(std::integral_constant<int, 1>{}, j + stride)) break;
    if (!step(std::integral_constant<int, 2>{}, j + 2 * stride)) break;
    if (!step(std::integral_constant<int, 3>{}, j + 3 * stride)) break;
  }
  cp_async_wait<0>();
}

}  // namespace pomp
