// grid_tile.cuh -- 2-D Euclidean k-NN as a STREAMING kernel: the bucket-sorted point array is
// pulled through shared memory in large contiguous pieces with the bulk async copy engine (TMA,
// cp.async.bulk + mbarrier), so HBM sees long sequential reads instead of per-thread 32 B sectors.
//
// Work decomposition: one CTA per tile of TW x TH grid cells.  Queries are bucket-sorted by home
// cell beforehand (the same counting sort the index uses), so the queries of a tile are TH
// contiguous ranges of the sorted query array.  The CTA
//   1. reads the cell_start entries bounding its halo rows (tile +- R cells): every halo row is ONE
//      contiguous run of the sorted point array;
//   2. one elected thread issues one cp.async.bulk per row into shared memory, all completing on
//      one mbarrier (expect_tx = total bytes);
//   3. each thread answers one query against the (2R+1)^2 block of cells around its home cell,
//      entirely from shared memory, keeps the k best in registers, and certifies exactness with the
//      distance to the outside of the block (same argument as grid_knn_block_kernel); uncertified
//      queries and tiles that do not fit the buffer go to the overflow list (pruned traversal).
// Several CTAs per SM overlap one tile's copy with another tile's search.
#pragma once
#include "grid_fast.cuh"

namespace pomp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

constexpr int TILE_RMAX = 3;

template <int TW, int TH>
struct TileShared {
  static constexpr int ROWS = TH + 2 * TILE_RMAX;
  static constexpr int COLS = TW + 2 * TILE_RMAX + 1;
  uint64_t bar;
  int row_gs[ROWS];   // first sorted position of each halo row
  int row_off[ROWS];  // its offset in the shared point buffer
  int qbeg[TH + 1];   // sorted-query ranges of the tile rows (prefix form in qpre)
  int qlen[TH + 1];
  int qpre[TH + 1];
  int total_pts, fits;
  int cs[ROWS * COLS];  // cell_start of every halo cell (+1 sentinel column)
};

// sorted queries: qs[t] (coordinates) and qperm[t] (original index), t in bucket order;
// qstart[y * nbx + tile column] = first sorted query of that 64-cell bin (the coarse query sort)
template <int K, int TW, int TH>
__global__ void __launch_bounds__(128)
grid_knn_tile2d_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                       const double* __restrict__ qs, const int* __restrict__ qperm, const int* __restrict__ qstart,
                       int nbx, int k, int R, int cap_pts, int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                       int32_t* __restrict__ out_cnt, int* __restrict__ ovf_list, int* __restrict__ ovf_count) {
  using SH = TileShared<TW, TH>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* spts = reinterpret_cast<double2*>(smem_raw);                       // cap_pts points
  SH& sh = *reinterpret_cast<SH*>(smem_raw + (size_t)cap_pts * sizeof(double2));

  const int W = g.ncell[0], H = g.ncell[1];
  const int ntx = (W + TW - 1) / TW;
  const int tx0 = (blockIdx.x % ntx) * TW, ty0 = (blockIdx.x / ntx) * TH;
  const int tx1 = min(tx0 + TW, W) - 1, ty1 = min(ty0 + TH, H) - 1;  // inclusive
  const int tid = threadIdx.x;

  // ---- queries of this tile: bin (y, tile column) of the coarse query sort, one range per tile row
  const int tcol = blockIdx.x % ntx;
  if (tid < TH) {
    int y = ty0 + tid;
    int b = 0, e = 0;
    if (y <= ty1) {
      b = __ldg(qstart + (long long)y * nbx + tcol);
      e = __ldg(qstart + (long long)y * nbx + tcol + 1);
    }
    sh.qbeg[tid] = b;
    sh.qlen[tid] = e - b;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r < TH; r++) {
      sh.qpre[r] = acc;
      acc += sh.qlen[r];
    }
    sh.qpre[TH] = acc;
  }
  __syncthreads();
  const int nqt = sh.qpre[TH];
  if (nqt == 0) return;  // nothing to answer here: no bytes moved

  // ---- halo rows: [ylo, yhi] x [xlo, xhi]
  const int xlo = max(tx0 - R, 0), xhi = min(tx1 + R, W - 1);
  const int ylo = max(ty0 - R, 0), yhi = min(ty1 + R, H - 1);
  const int nrows = yhi - ylo + 1, ncols = xhi - xlo + 1;
  for (int i = tid; i < nrows * (ncols + 1); i += blockDim.x) {
    int r = i / (ncols + 1), cidx = i % (ncols + 1);
    sh.cs[r * SH::COLS + cidx] = __ldg(g.cell_start + (long long)(ylo + r) * W + xlo + cidx);
  }
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r < nrows; r++) {
      int s = sh.cs[r * SH::COLS], e = sh.cs[r * SH::COLS + ncols];
      sh.row_gs[r] = s;
      sh.row_off[r] = acc;
      acc += e - s;
    }
    sh.total_pts = acc;
    sh.fits = acc <= cap_pts;
    if (sh.fits && acc > 0) {
      mbar_expect_tx(&sh.bar, (uint32_t)acc * 16u);
      for (int r = 0; r < nrows; r++) {
        int s = sh.row_gs[r], cnt = sh.cs[r * SH::COLS + ncols] - s;
        if (cnt > 0) bulk_g2s(spts + sh.row_off[r], g.spts + (long long)s * 2, (uint32_t)cnt * 16u, &sh.bar);
      }
    }
  }
  __syncthreads();
  const bool fits = sh.fits != 0;
  if (fits && sh.total_pts > 0) mbar_wait(&sh.bar, 0);

  // ---- one query per thread
  for (int t = tid; t < nqt; t += blockDim.x) {
    int r = 0;
#pragma unroll
    for (int j = 1; j < TH; j++)
      if (t >= sh.qpre[j]) r = j;
    const int spos = sh.qbeg[r] + (t - sh.qpre[r]);  // position in the sorted query order
    const long long qi = __ldg(qperm + spos);
    if (!fits) {
      int o = atomicAdd(ovf_count, 1);
      ovf_list[o] = (int)qi;
      continue;
    }
    double2 qv = __ldg(reinterpret_cast<const double2*>(qs) + spos);
    double qq[2] = {qv.x, qv.y};
    const int cx = grid_cell(g, 0, qq[0]), cy = grid_cell(g, 1, qq[1]);
    PosTopK<K> top;
    top.init(IdxResolver{g.sidx, g.n_indexed});
    const int bx0 = max(cx - R, 0), bx1 = min(cx + R, W - 1);
    for (int y = max(cy - R, 0); y <= min(cy + R, H - 1); y++) {
      const int rr = y - ylo;
      const int s = sh.cs[rr * SH::COLS + (bx0 - xlo)], e = sh.cs[rr * SH::COLS + (bx1 + 1 - xlo)];
      const int base = sh.row_off[rr] - sh.row_gs[rr];
      for (int pos = s; pos < e; pos++) {
        if (skip && skip[__ldg(g.sidx + pos)]) continue;
        double2 pv = spts[base + pos];
        double dx = pv.x - qq[0], dy = pv.y - qq[1];
        double sq = 0.0;
        sq = sq + dx * dx;
        sq = sq + dy * dy;
        top.offer_sq(sq, pos);
      }
    }
    scan_tail<2>(g, pts, n, skip, qq, top);
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
    bool exact = !(margin < POMP_INF) || (top.T * (1.0 + 1e-9) < margin);
    if (!exact) {
      int o = atomicAdd(ovf_count, 1);
      ovf_list[o] = (int)qi;
      continue;
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
}

}  // namespace pomp
