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

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
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
// Persistent, warp-specialised streaming kernel, v2.
//
// One CTA per SM loops over tiles of ST_TW x ST_TH cells.  PRODUCER warps run ahead: for each tile
// they read the cell_start entries that bound its halo rows and the tile's range in the TILE-MAJOR
// bucket-sorted query array, then issue cp.async.bulk (TMA) copies -- one per halo row of the sorted
// point array (~2 KB each), one for the tile's queries, one for their original indices -- into one
// stage of a shared-memory ring; all copies of a stage complete on its `full` mbarrier.  A pool of
// CONSUMER warps searches the staged tiles (one query per thread, flattened 3x3 block + certificate
// as in grid_knn_flat2d_kernel, candidates read from shared memory) and signals `empty`.
// profiles/microbench: the bulk-copy engine sustains the full HBM rate as long as a stage is <= 16
// copies of >= 2 KB; v1 staged cell_start, sidx and per-row query slices as well (~46 small copies
// per tile) and stalled at ~3 TB/s.
// =====================================================================================================
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int ST_TW = 64;             // tile width in cells = query-sort bin width (shift 6)
constexpr int ST_TH = 8;              // tile height in cell rows = query-sort bin height
constexpr int ST_ROWS = ST_TH + 2 * TILE_RMAX;
constexpr int ST_STAGES = 6;
constexpr int ST_PROD_WARPS = 4;      // producers take tiles round-robin: 4 metadata round trips in flight
constexpr int ST_CONS_WARPS = 12;     // consumer warp pool
constexpr int ST_SLOTS = 5;           // 32-query chunk slots per tile (cap_q <= 160)
constexpr int ST_CSTRIDE = 2;         // slot c of tile `it` belongs to consumer warp (it*ST_CSTRIDE + c) % ST_CONS_WARPS
constexpr int ST_THREADS = 32 * (ST_PROD_WARPS + ST_CONS_WARPS);
static_assert(ST_PROD_WARPS <= ST_STAGES, "producer run-ahead must stay within one ring revolution");

struct StreamHeader {
  int nq, fits, nrows, ylo, qbeg, qp_shift, pad0, pad1;
  int row_gs[ST_ROWS];   // first sorted position of each halo row
  int row_off[ST_ROWS];  // offset of the row in the stage's point buffer
};

struct StreamLayout {  // byte offsets inside one stage
  int cap_pts, cap_q;
  int off_pts, off_q, off_qperm, off_hdr, stage_bytes;
};

__host__ __device__ inline StreamLayout stream_layout(int cap_pts, int cap_q) {
  StreamLayout L;
  L.cap_pts = cap_pts;
  L.cap_q = cap_q;
  L.off_pts = 0;
  L.off_q = L.off_pts + cap_pts * 16;
  L.off_qperm = L.off_q + cap_q * 16;          // aligned slice of qperm (+ up to 3 leading, 3 trailing ints)
  L.off_hdr = L.off_qperm + (cap_q + 8) * 4;
  L.stage_bytes = (L.off_hdr + (int)sizeof(StreamHeader) + 127) / 128 * 128;
  return L;
}

// qs / qperm: queries bucket-sorted by TILE (bin = tile_row * ntx + tile_col); qstart[bin] = first query
template <int K>
__global__ void __launch_bounds__(ST_THREADS, 1)
grid_knn_stream2d_kernel(GridDev g, const double* __restrict__ pts, long long n, const uint8_t* __restrict__ skip,
                         const double* __restrict__ qs, const int* __restrict__ qperm,
                         const int* __restrict__ qstart, int k, int R, int cap_pts, int cap_q,
                         int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                         int* __restrict__ ovf_list, int* __restrict__ ovf_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const StreamLayout L = stream_layout(cap_pts, cap_q);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [ST_STAGES]
  uint64_t* empty = full + ST_STAGES;                      // [ST_STAGES]
  volatile int* seq = reinterpret_cast<volatile int*>(empty + ST_STAGES);  // tile iteration staged in each slot
  int* done = const_cast<int*>(seq) + ST_STAGES;            // chunks of the staged tile already searched
  unsigned char* stages = smem_raw + 256;

  const int W = g.ncell[0], H = g.ncell[1];
  const int ntx = (W + ST_TW - 1) / ST_TW, nty = (H + ST_TH - 1) / ST_TH;
  const int ntiles = ntx * nty;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      seq[s] = -1;
      done[s] = 0;
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
      if (it >= ST_STAGES) {
        // slots are released out of order, so the parity of `empty` alone is ambiguous (fill j-2
        // still in use looks like fill j-1 released): first wait until the previous fill of this
        // slot has at least been STARTED (its producer published seq), then for its release
        while (seq[s] != it - ST_STAGES) __nanosleep(32);
        mbar_wait(&empty[s], ((it / ST_STAGES) - 1) & 1);
      }
      unsigned char* st = stages + (size_t)s * L.stage_bytes;
      StreamHeader* hd = reinterpret_cast<StreamHeader*>(st + L.off_hdr);
      const int tcol = tile % ntx;
      const int tx0 = tcol * ST_TW, ty0 = (tile / ntx) * ST_TH;
      const int tx1 = min(tx0 + ST_TW, W) - 1, ty1 = min(ty0 + ST_TH, H) - 1;
      const int xlo = max(tx0 - R, 0), xhi = min(tx1 + R, W - 1);
      const int ylo = max(ty0 - R, 0), yhi = min(ty1 + R, H - 1);
      const int nrows = yhi - ylo + 1;
      int qb = 0, qn = 0;
      if (lane == 31) {
        qb = __ldg(qstart + tile);
        qn = __ldg(qstart + tile + 1) - qb;
      }
      int gs = 0, gn = 0;
      if (lane < nrows) {
        gs = __ldg(g.cell_start + (long long)(ylo + lane) * W + xlo);
        gn = __ldg(g.cell_start + (long long)(ylo + lane) * W + xhi + 1) - gs;
      }
      qb = __shfl_sync(0xffffffffu, qb, 31);
      qn = __shfl_sync(0xffffffffu, qn, 31);
      int gincl = gn;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int b = __shfl_up_sync(0xffffffffu, gincl, o);
        if (lane >= o) gincl += b;
      }
      const int npts = __shfl_sync(0xffffffffu, gincl, 31);
      const int qal = qb & ~3, qaln = qn > 0 ? (((qb + qn + 3) & ~3) - qal) : 0;
      const bool fits = npts <= cap_pts && qn <= cap_q;
      // seqlock: publish "slot s now belongs to tile iteration `it`" BEFORE touching the header
      if (lane == 0) {
        done[s] = 0;
        seq[s] = it;
      }
      __threadfence_block();
      __syncwarp();
      if (lane < nrows) {
        hd->row_gs[lane] = gs;
        hd->row_off[lane] = gincl - gn;
      }
      if (lane == 0) {
        hd->nq = qn;
        hd->fits = fits ? 1 : 0;
        hd->nrows = nrows;
        hd->ylo = ylo;
        hd->qbeg = qb;
        hd->qp_shift = qb - qal;
      }
      __syncwarp();
      if (qn == 0 || !fits) {
        if (lane == 0) {
          mbar_arrive(&full[s]);               // header only (unfit tiles: consumers read qperm from global)
          if (qn == 0) mbar_arrive(&empty[s]);  // nobody will search an empty tile: release the slot ourselves
        }
        continue;
      }
      if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)npts * 16u + (uint32_t)qn * 16u + (uint32_t)qaln * 4u);
      __syncwarp();
      double2* spts = reinterpret_cast<double2*>(st + L.off_pts);
      if (lane < nrows && gn > 0)
        bulk_g2s(spts + (gincl - gn), g.spts + (long long)gs * 2, (uint32_t)gn * 16u, &full[s]);
      if (lane == 30) bulk_g2s(st + L.off_q, qs + (long long)qb * 2, (uint32_t)qn * 16u, &full[s]);
      if (lane == 31) bulk_g2s(st + L.off_qperm, qperm + qal, (uint32_t)qaln * 4u, &full[s]);
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  // Chunk slot c (32 queries) of tile iteration `it` belongs to warp (it*ST_CSTRIDE + c) % NC -- a
  // static deal, so a warp visits only the tiles where it may have work and a busy warp never
  // holds back the release of other tiles.  Because warps skip tiles, the mbarrier parity alone
  // cannot tell "fill j done" from "fill j-1 not started": seq[s] (written by the producer before
  // it arms the barrier) pins the ring revolution; a slot is released by whichever owner finishes
  // the tile's last real chunk (done[s] counter), BEFORE the sidx gather and the result stores.
  const int cw = warp - ST_PROD_WARPS;
  constexpr int NC = ST_CONS_WARPS;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const int c = (((cw - it * ST_CSTRIDE) % NC) + NC) % NC;
    if (c >= ST_SLOTS) continue;
    const int s = it % ST_STAGES;
    int sv;
    while ((sv = seq[s]) < it) __nanosleep(32);
    if (sv > it) continue;  // released and refilled without us: we had no real chunk there
    {
      // a warp without a real chunk here may lag by a whole ring revolution, where the parity
      // repeats: give up as soon as the slot moves on
      bool moved = false;
      while (!mbar_try_wait(&full[s], (it / ST_STAGES) & 1))
        if (seq[s] != it) {
          moved = true;
          break;
        }
      if (moved) continue;
    }
    unsigned char* st = stages + (size_t)s * L.stage_bytes;
    const StreamHeader* hd = reinterpret_cast<const StreamHeader*>(st + L.off_hdr);
    const int nq = hd->nq;
    const int fits_i = hd->fits, ylo = hd->ylo, qbeg = hd->qbeg, qp_shift = hd->qp_shift;
    __threadfence_block();
    if (seq[s] != it) continue;  // header already belongs to a later tile (again: no real chunk of ours)
    const int nchunk = min((nq + 31) >> 5, ST_SLOTS);  // real owners of this tile
    if (c >= nchunk) continue;
    if (!fits_i) {
      // tile too dense for the stage (nothing was copied): hand all its queries to the overflow pass
      for (int tt = c * 32 + lane; tt < nq; tt += ST_SLOTS * 32) {
        int o = atomicAdd(ovf_count, 1);
        ovf_list[o] = __ldg(qperm + qbeg + tt);
      }
      __syncwarp();
      if (lane == 0) {
        int dprev = atomicAdd(&done[s], 1);
        if (dprev == nchunk - 1) mbar_arrive(&empty[s]);
      }
      continue;
    }
    const double2* spts = reinterpret_cast<const double2*>(st + L.off_pts);
    const double2* sq = reinterpret_cast<const double2*>(st + L.off_q);
    const int* sqperm = reinterpret_cast<const int*>(st + L.off_qperm) + qp_shift;
    const int t = c * 32 + lane;
    const bool active = t < nq;
    long long qi = 0;
    bool ovf = false, have = false;
    typename BestList<K>::type top;
    top.init(IdxResolver{g.sidx, g.n_indexed});
    double qq[2] = {0.0, 0.0};
    if (active) {
      qi = sqperm[t];
      const double2 qv = sq[t];
      qq[0] = qv.x;
      qq[1] = qv.y;
      const int cx = grid_cell(g, 0, qq[0]), cy = grid_cell(g, 1, qq[1]);
      const int x0 = max(cx - R, 0), x1 = min(cx + R, W - 1);
      // all cell_start bounds of the block first (one round trip), then the staged candidates
      constexpr int NRMAX = 2 * TILE_RMAX + 1;
      int rs[NRMAX], re[NRMAX];
#pragma unroll
      for (int rr = 0; rr < NRMAX; rr++) {
        int y = cy + ((rr & 1) ? -((rr + 1) >> 1) : (rr >> 1));  // centre row first
        bool ok = rr <= 2 * R && y >= 0 && y < H;
        long long base = (long long)(ok ? y : 0) * W;
        int a = __ldg(g.cell_start + base + x0), b = __ldg(g.cell_start + base + x1 + 1);
        rs[rr] = a;
        re[rr] = ok ? b : a;
      }
#pragma unroll
      for (int rr = 0; rr < NRMAX; rr++) {
        if (rr > 2 * R) break;
        int y = cy + ((rr & 1) ? -((rr + 1) >> 1) : (rr >> 1));
        const int hr = min(max(y - ylo, 0), ST_ROWS - 1);
        const int sbase = hd->row_off[hr] - hd->row_gs[hr];
        for (int pos = rs[rr]; pos < re[rr]; pos++) {
          if (skip && skip[__ldg(g.sidx + pos)]) continue;
#ifdef POMP_STREAM_DEBUG
          if (sbase + pos < 0 || sbase + pos >= cap_pts || y - ylo < 0 || y - ylo >= hd->nrows) {
            printf("stream OOB: tile %d it %d s %d c %d t %d nq %d q(%g,%g) cx %d cy %d y %d ylo %d nrows %d hr %d "
                   "row_off %d row_gs %d pos %d [%d,%d) seq %d\n",
                   tile, it, s, c, t, nq, qq[0], qq[1], cx, cy, y, ylo, hd->nrows, hr, hd->row_off[hr],
                   hd->row_gs[hr], pos, rs[rr], re[rr], seq[s]);
            break;
          }
#endif
          const double2 pv = spts[sbase + pos];
          double dx = pv.x - qq[0], dy = pv.y - qq[1];
          double sqd = 0.0;
          sqd = sqd + dx * dx;
          sqd = sqd + dy * dy;
          top.offer_sq(sqd, pos);
        }
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
      // the unindexed tail (global memory) can only tighten T, so the certificate may use it
      scan_tail<2>(g, pts, n, skip, qq, top);
      const bool exact = !(margin < POMP_INF) || (top.T * (1.0 + 1e-9) < margin);
      have = exact;
      ovf = !exact;
    }
    // the staged data is no longer needed: release the slot before the global gathers / stores
    __syncwarp();
    if (lane == 0) {
      int dprev = atomicAdd(&done[s], 1);
      if (dprev == nchunk - 1) mbar_arrive(&empty[s]);
    }
    if (ovf) {
      int o = atomicAdd(ovf_count, 1);
      ovf_list[o] = (int)qi;
    }
    if (have) {
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
}

}  // namespace pomp
