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

// =====================================================================================================
// Persistent, warp-specialised streaming kernel (the headline path).
//
// One CTA per SM loops over tiles of TW x TH cells.  A PRODUCER warp runs ahead: for each tile it
// reads the few cell_start / qstart entries that bound the tile's halo rows and query ranges, then
// issues cp.async.bulk (TMA) copies -- the halo rows of the bucket-sorted point array, the matching
// slices of cell_start and the tile's bucket-sorted queries -- into one stage of a shared-memory
// ring; all copies of a stage complete on its `full` mbarrier.  Two CONSUMER groups (4 warps each)
// alternate tiles: one query per thread, searched entirely in shared memory (flattened (2R+1)^2
// block + certificate, like grid_knn_flat2d_kernel), then `empty` is signalled and the producer
// refills the stage.  HBM therefore sees a continuous stream of multi-KB sequential reads kept in
// flight by the ring, independent of the consumers' latency.
// =====================================================================================================
// candidates of the streaming kernel are identified by their slot in the stage's sidx slice; points
// appended after the index build (never staged) carry their insertion index with a tag bit
struct StageResolver {
  const int* ssidx;
  __device__ __forceinline__ int operator()(int p) const { return (p & 0x40000000) ? (p & 0x3fffffff) : ssidx[p]; }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int ST_TW = 64;             // tile width in cells = query-sort bin width (shift 6)
constexpr int ST_TH = 8;              // tile height in cell rows
constexpr int ST_ROWS = ST_TH + 2 * TILE_RMAX;
constexpr int ST_CSW = ST_TW + 16;    // cell_start ints copied per halo row (16-byte aligned slices)
constexpr int ST_STAGES = 5;
constexpr int ST_PROD_WARPS = 4;      // producers take tiles round-robin: 4 metadata round trips in flight
constexpr int ST_CONS_WARPS = 12;     // consumer warp pool
constexpr int ST_SLOTS = 5;           // 32-query chunk slots per tile (cap_q <= 160)
constexpr int ST_THREADS = 32 * (ST_PROD_WARPS + ST_CONS_WARPS);

struct StreamHeader {
  int nq, fits, nrows, ylo, xal, tx0, ty0, tx1, ty1, pad;
  int row_gs[ST_ROWS];   // first sorted position of each halo row
  int row_off[ST_ROWS];  // offset of the row in the stage's point buffer
  int qbeg[ST_TH];       // first sorted query of each tile row
  int qpre[ST_TH + 1];   // prefix of per-row query counts (= offset in the stage's query buffer)
  int row_ioff[ST_ROWS]; // sidx slice of halo row r starts at ssidx[row_ioff[r]] and holds sidx[row_gs[r] & ~3 ...]
  int qp_off[ST_TH];     // qperm slice of tile row r starts at sqperm[qp_off[r]] and holds qperm[qbeg[r] & ~3 ...]
};

struct StreamLayout {  // byte offsets inside one stage
  int cap_pts, cap_q;
  int off_pts, off_cs, off_q, off_sidx, off_qperm, off_hdr, stage_bytes;
};

__host__ __device__ inline StreamLayout stream_layout(int cap_pts, int cap_q) {
  StreamLayout L;
  L.cap_pts = cap_pts;
  L.cap_q = cap_q;
  L.off_pts = 0;
  L.off_cs = L.off_pts + cap_pts * 16;
  L.off_q = L.off_cs + ST_ROWS * ST_CSW * 4;
  L.off_sidx = L.off_q + cap_q * 16;                       // per halo row: aligned slice of sidx (+3 slack each)
  L.off_qperm = L.off_sidx + (cap_pts + 8 * ST_ROWS) * 4;  // per tile row: aligned slice of qperm
  L.off_hdr = L.off_qperm + (cap_q + 8 * ST_TH) * 4;
  L.stage_bytes = (L.off_hdr + (int)sizeof(StreamHeader) + 127) / 128 * 128;
  return L;
}

template <int K>
__global__ void __launch_bounds__(ST_THREADS, 1)
grid_knn_stream2d_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                         const double* __restrict__ qs, const int* __restrict__ qperm,
                         const int* __restrict__ qstart, int nbx, int k, int R, int cap_pts, int cap_q,
                         int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                         int* __restrict__ ovf_list, int* __restrict__ ovf_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const StreamLayout L = stream_layout(cap_pts, cap_q);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [ST_STAGES]
  uint64_t* empty = full + ST_STAGES;                      // [ST_STAGES]
  unsigned char* stages = smem_raw + 128;

  const int W = g.ncell[0], H = g.ncell[1];
  const int ntx = (W + ST_TW - 1) / ST_TW, nty = (H + ST_TH - 1) / ST_TH;
  const int ntiles = ntx * nty;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ST_CONS_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp < ST_PROD_WARPS) {
    // ------------------------------------------------------------------ producers
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      if (it % ST_PROD_WARPS != warp) continue;
      const int s = it % ST_STAGES;
      if (it >= ST_STAGES) mbar_wait(&empty[s], ((it / ST_STAGES) - 1) & 1);
      unsigned char* st = stages + (size_t)s * L.stage_bytes;
      StreamHeader* hd = reinterpret_cast<StreamHeader*>(st + L.off_hdr);
      const int tcol = tile % ntx;
      const int tx0 = tcol * ST_TW, ty0 = (tile / ntx) * ST_TH;
      const int tx1 = min(tx0 + ST_TW, W) - 1, ty1 = min(ty0 + ST_TH, H) - 1;
      const int xlo = max(tx0 - R, 0), xhi = min(tx1 + R, W - 1);
      const int ylo = max(ty0 - R, 0), yhi = min(ty1 + R, H - 1);
      const int nrows = yhi - ylo + 1;
      const int xal = max(tx0 - 4, 0);
      // query ranges of the tile rows (lanes 0..TH-1) and point ranges of the halo rows (lanes 0..nrows-1)
      int qb = 0, qn = 0;
      if (lane < ST_TH && ty0 + lane <= ty1) {
        qb = __ldg(qstart + (long long)(ty0 + lane) * nbx + tcol);
        qn = __ldg(qstart + (long long)(ty0 + lane) * nbx + tcol + 1) - qb;
      }
      int gs = 0, gn = 0;
      if (lane < nrows) {
        gs = __ldg(g.cell_start + (long long)(ylo + lane) * W + xlo);
        gn = __ldg(g.cell_start + (long long)(ylo + lane) * W + xhi + 1) - gs;
      }
      int qincl = qn, gincl = gn;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, qincl, o), b = __shfl_up_sync(0xffffffffu, gincl, o);
        if (lane >= o) {
          qincl += a;
          gincl += b;
        }
      }
      const int nq = __shfl_sync(0xffffffffu, qincl, 31), npts = __shfl_sync(0xffffffffu, gincl, 31);
      // 16-byte aligned slices of the 4-byte arrays (sidx per halo row, qperm per tile row)
      const int gal = gs & ~3, galn = gn > 0 ? (((gs + gn + 3) & ~3) - gal) : 0;
      const int qal = qb & ~3, qaln = qn > 0 ? (((qb + qn + 3) & ~3) - qal) : 0;
      int iincl = galn, pincl = qaln;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int a = __shfl_up_sync(0xffffffffu, iincl, o), b = __shfl_up_sync(0xffffffffu, pincl, o);
        if (lane >= o) {
          iincl += a;
          pincl += b;
        }
      }
      const int nidx = __shfl_sync(0xffffffffu, iincl, 31), nqp = __shfl_sync(0xffffffffu, pincl, 31);
      const bool fits = npts <= cap_pts && nq <= cap_q && nidx <= cap_pts + 8 * ST_ROWS && nqp <= cap_q + 8 * ST_TH;
      if (lane < ST_TH) {
        hd->qbeg[lane] = qb;
        hd->qpre[lane] = qincl - qn;
        hd->qp_off[lane] = pincl - qaln;
      }
      if (lane == ST_TH) hd->qpre[ST_TH] = nq;
      if (lane < nrows) {
        hd->row_gs[lane] = gs;
        hd->row_off[lane] = gincl - gn;
        hd->row_ioff[lane] = iincl - galn;
      }
      if (lane == 0) {
        hd->nq = nq;
        hd->fits = fits ? 1 : 0;
        hd->nrows = nrows;
        hd->ylo = ylo;
        hd->xal = xal;
        hd->tx0 = tx0;
        hd->ty0 = ty0;
        hd->tx1 = tx1;
        hd->ty1 = ty1;
      }
      __syncwarp();
      if (nq == 0 || !fits) {
        if (lane == 0) mbar_arrive(&full[s]);  // header only
        continue;
      }
      if (lane == 0) {
        uint32_t bytes = (uint32_t)npts * 16u + (uint32_t)nq * 16u + (uint32_t)nrows * (ST_CSW * 4u) +
                         (uint32_t)nidx * 4u + (uint32_t)nqp * 4u;
        mbar_expect_tx(&full[s], bytes);
      }
      __syncwarp();
      double2* spts = reinterpret_cast<double2*>(st + L.off_pts);
      int* scs = reinterpret_cast<int*>(st + L.off_cs);
      double2* sq = reinterpret_cast<double2*>(st + L.off_q);
      if (lane < nrows) {
        if (gn > 0) bulk_g2s(spts + (gincl - gn), g.spts + (long long)gs * 2, (uint32_t)gn * 16u, &full[s]);
        bulk_g2s(scs + lane * ST_CSW, g.cell_start + (long long)(ylo + lane) * W + xal, ST_CSW * 4u, &full[s]);
      }
      if (lane < ST_TH && qn > 0) bulk_g2s(sq + (qincl - qn), qs + (long long)qb * 2, (uint32_t)qn * 16u, &full[s]);
      int* ssidx = reinterpret_cast<int*>(st + L.off_sidx);
      int* sqperm = reinterpret_cast<int*>(st + L.off_qperm);
      if (lane < nrows && galn > 0) bulk_g2s(ssidx + (iincl - galn), g.sidx + gal, (uint32_t)galn * 4u, &full[s]);
      if (lane < ST_TH && qaln > 0) bulk_g2s(sqperm + (pincl - qaln), qperm + qal, (uint32_t)qaln * 4u, &full[s]);
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  // A pool of warps flows through the stages: the 32-query chunks of tile `it` are dealt to warps
  // (rot + chunk) % NC with a rotating offset, so consecutive tiles occupy different warps and
  // several stages are searched concurrently while others are still loading.
  const int cw = warp - ST_PROD_WARPS;
  constexpr int NC = ST_CONS_WARPS;
  // A pool of warps flows through the stages IN LOCKSTEP ORDER (every warp waits on every tile's
  // `full` and arrives on its `empty`, so no warp can run a whole ring ahead of the barrier phases);
  // the 32-query chunks of a tile are dealt to warps (rot + chunk) % NC with a rotating offset, so
  // consecutive tiles keep different warps busy.  With points, cell_start, queries, sidx and qperm all
  // staged, a chunk is ~1 us of shared-memory work, short against the ring's refill latency.
  int it = 0, rot = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const int s = it % ST_STAGES;
    mbar_wait(&full[s], (it / ST_STAGES) & 1);
    unsigned char* st = stages + (size_t)s * L.stage_bytes;
    const StreamHeader* hd = reinterpret_cast<const StreamHeader*>(st + L.off_hdr);
    const int nq = hd->nq;
    const int nchunk = (nq + 31) >> 5;
    const int first = ((cw - rot) % NC + NC) % NC;  // my first chunk of this tile
    rot = (rot + nchunk) % NC;
    if (first < nchunk)
    {
      const bool fits = hd->fits != 0;
      const double2* spts = reinterpret_cast<const double2*>(st + L.off_pts);
      const int* scs = reinterpret_cast<const int*>(st + L.off_cs);
      const double2* sq = reinterpret_cast<const double2*>(st + L.off_q);
      const int* ssidx = reinterpret_cast<const int*>(st + L.off_sidx);
      const int* sqperm = reinterpret_cast<const int*>(st + L.off_qperm);
      const int ylo = hd->ylo, xal = hd->xal;
      for (int t = first * 32 + lane; t < nq; t += NC * 32) {
        int r = 0;
#pragma unroll
        for (int j = 1; j < ST_TH; j++)
          if (t >= hd->qpre[j]) r = j;
        const int spos = hd->qbeg[r] + (t - hd->qpre[r]);
        const long long qi = fits ? (long long)sqperm[hd->qp_off[r] + (spos - (hd->qbeg[r] & ~3))]
                                  : (long long)__ldg(qperm + spos);
        if (!fits) {
          int o = atomicAdd(ovf_count, 1);
          ovf_list[o] = (int)qi;
          continue;
        }
        const double2 qv = sq[t];
        double qq[2] = {qv.x, qv.y};
        const int cx = grid_cell(g, 0, qq[0]), cy = grid_cell(g, 1, qq[1]);
        typename BestList<K, StageResolver>::type top;
        top.init(StageResolver{ssidx});
        const int bx0 = max(cx - R, 0), bx1 = min(cx + R, W - 1);
        for (int rr = 0; rr <= 2 * R; rr++) {
          int y = cy + ((rr & 1) ? -((rr + 1) >> 1) : (rr >> 1));  // centre row first
          if (y < 0 || y >= H) continue;
          const int hr = y - ylo;
          const int s0 = scs[hr * ST_CSW + (bx0 - xal)], e0 = scs[hr * ST_CSW + (bx1 + 1 - xal)];
          const int gs0 = hd->row_gs[hr];
          const int base = hd->row_off[hr] - gs0;
          const int ibase = hd->row_ioff[hr] - (gs0 & ~3);
          for (int pos = s0; pos < e0; pos++) {
            if (skip && skip[ssidx[ibase + pos]]) continue;
            const double2 pv = spts[base + pos];
            double dx = pv.x - qq[0], dy = pv.y - qq[1];
            double sqd = 0.0;
            sqd = sqd + dx * dx;
            sqd = sqd + dy * dy;
            top.offer_sq(sqd, ibase + pos);
          }
        }
        for (long long i = g.n_indexed; i < n; i++) {  // unindexed tail, straight from global memory
          if (skip && skip[i]) continue;
          double p2[2];
          load_point<2>(pts, i, p2);
          top.offer_sq(sqdist<2>(p2, qq), (int)i | 0x40000000);
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
        const bool exact = !(margin < POMP_INF) || (top.T * (1.0 + 1e-9) < margin);
        if (!exact) {
          int o = atomicAdd(ovf_count, 1);
          ovf_list[o] = (int)qi;
          continue;
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
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

}  // namespace pomp
