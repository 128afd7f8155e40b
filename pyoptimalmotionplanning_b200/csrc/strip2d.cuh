// strip2d.cuh -- the 2-D Euclidean nearest-neighbour search as a STREAMING kernel over a blocked index.
//
// Why: the per-query gather kernels (grid_fast.cuh) touch the sorted point array through scattered 16-byte
// loads -- 20 of 32 lanes busy, 60 % of the stalls on memory latency, 0.46 of the HBM roofline.  A large batch
// (2^20 queries against 2^24 points) reads almost every point anyway, so the point array is better pulled
// through shared memory ONCE, in large contiguous pieces, by the bulk copy engine, with the queries answered
// out of shared memory.
//
// Layout ("strip index", derived from the row-major grid index by one segmented copy):
//   * the W x H cell grid is cut into vertical STRIPS of ST_TW cells; a strip stores, row by row, its own
//     cells plus ONE replicated halo cell column on either side ((ST_TW+2)/ST_TW = 1.03x the points);
//   * inside a strip the rows follow each other, so a TILE (ST_TW x ST_TH cells) together with its full
//     one-cell halo -- rows y0-1 .. y0+ST_TH of the strip -- is ONE contiguous run of points: one bulk copy;
//   * per tile, the offsets of its (ST_TH+2) x (ST_TW+2) local cells inside that run, as uint16
//     (tile-local: one more bulk copy, 4.5 KB);
//   * sidx2[slot] = insertion index of every (replicated) slot.
// Queries are bucketed by tile in ONE pass (fixed-capacity buckets, one padded counter per tile: no scan, no
// second pass); tile ids are strip-major, so consecutive tiles share halo rows in L2.
//
// Kernel: persistent, one CTA per SM, 1 producer warp + up to ST_MAX_STAGES consumer GROUPS of ST_WG warps, one
// stage buffer (header, points, cell offsets) per group, guarded by header/full/empty mbarriers.  The producer
// walks the CTA's tiles (metadata of 32 tiles per fetch), skips those without queries and hands each staged tile
// to the first free group; the header lands first, so the group fetches its queries while the bulk copies are
// in flight, and a warp only ever visits tiles it has work on.  (First version: one ring shared by all warps,
// every warp walking every stage -- 34 M of 63 M executed instructions were that walk, 0.100 ms.)  A consumer
// lane owns one query: the 3 row segments of the 3x3 block around its home cell are read from shared memory,
// the smallest and second-smallest SQUARED distances are tracked branch-free (no sqrt per candidate), the
// winner is rooted once.  Exactness: (1) the same certificate as grid_knn_flat2d_kernel (distance to the
// outside of the block); (2) if the runner-up could round to the same rooted distance (sec <= tsq_of(d), which
// includes exact ties) the index tie-break of the reference matters and the query is deferred.  Deferred
// queries -- uncertified, ambiguous, tiles too full for a stage, buckets that overflowed -- go to the overflow
// list and are answered by the warp-cooperative / gather kernels with the sequential rule.  Same arithmetic
// per candidate (one rounding per operation) => bit-identical answers.
#pragma once
#include "grid_fast.cuh"
#include "tma.cuh"

namespace pomp {

// Tile geometry, measured on B200 (2^20 queries vs 2^24 points, ms per step; tile height / consumer warps per group /
// stage buffers): 16/4/6 0.1126 (the first tuned version), 16/3/6 0.1149, 16/5/6 0.1104, 8/2/12 0.1258, 24/6/4
// 0.1104, 24/7/4 0.1094, 32/8/3 0.1096, 48/12/2 0.1128, 32/10/3 0.1079: taller tiles replicate fewer halo rows
// (34/32 instead of 18/16) and 31 consumer warps (64 registers each) hide more of the per-query latency.
#ifndef POMP_ST_TH
#define POMP_ST_TH 32
#endif
#ifndef POMP_ST_WG
#define POMP_ST_WG 10
#endif
#ifndef POMP_ST_STAGES
#define POMP_ST_STAGES 3
#endif
constexpr int ST_TW = 64;   // tile width in cells (= strip width)
constexpr int ST_TH = POMP_ST_TH;   // tile height in cells
constexpr int ST_LW = ST_TW + 2;  // local cell row incl. halo
constexpr int ST_LH = ST_TH + 2;
constexpr int ST_NCELL = ST_LW * ST_LH;            // 2244 local cells (66 x 34)
constexpr int ST_CS_STRIDE = (ST_NCELL + 1 + 7) / 8 * 8;  // uint16 entries per tile, 16-byte multiple (2248 for 64 x 32 cells)
constexpr int ST_CS_BYTES = ST_CS_STRIDE * 2;
constexpr int ST_OLIST = 1024;   // CTA-local deferred list (expected ~20 entries per CTA at the bench density)
constexpr int ST_WG = POMP_ST_WG; // consumer warps per group (one stage buffer per group): 320 queries per sweep
constexpr int ST_MAX_STAGES = POMP_ST_STAGES; // groups = stage buffers; how many fit is decided at launch from the cell occupancy
constexpr int ST_THREADS = (ST_MAX_STAGES * ST_WG + 1) * 32;
constexpr int ST_HDR_BYTES = 96;   // stage header: tile, first slot, next tile, sub-bucket prefix sums of both
#ifndef POMP_ST_SUB
#define POMP_ST_SUB 8
#endif
constexpr int ST_SUB = POMP_ST_SUB;          // sub-buckets per tile: same-address atomics serialise at ~0.25 us each in L2
                                   // (95 queries on one counter: 23 us for the bucketing pass; 12 per counter: hidden)
constexpr int ST_CNT_STRIDE = 8;   // ints between two counters: one counter per 32-byte sector


// ---------------------------------------------------------------------------- index build
// length of strip row (s, y): cells max(s*TW-1,0) .. min(s*TW+TW, W-1) of grid row y
static __global__ void strip_row_len_kernel(const int* __restrict__ cell_start, int W, int H, int nstrips,
                                            int* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nstrips * H) return;
  const int s = i / H, y = i - s * H;
  const int xa = max(s * ST_TW - 1, 0), xb = min(s * ST_TW + ST_TW, W - 1);
  const long long base = (long long)y * W;
  len[i] = __ldg(cell_start + base + xb + 1) - __ldg(cell_start + base + xa);
}

// one warp copies one strip row (points + insertion indices), coalesced on both sides
static __global__ void __launch_bounds__(256)
strip_copy_kernel(const int* __restrict__ cell_start, const double2* __restrict__ spts, const int* __restrict__ sidx,
                  int W, int H, int nstrips, const int* __restrict__ srow_start, double2* __restrict__ spts2,
                  int* __restrict__ sidx2) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= nstrips * H) return;
  const int s = i / H, y = i - s * H;
  const int xa = max(s * ST_TW - 1, 0);
  const int src = __ldg(cell_start + (long long)y * W + xa);
  const int dst = __ldg(srow_start + i), n = __ldg(srow_start + i + 1) - dst;
  for (int j = lane; j < n; j += 32) {
    spts2[dst + j] = __ldg(spts + src + j);
    sidx2[dst + j] = __ldg(sidx + src + j);
  }
}

// one block per tile: (first slot, slot count) and the uint16 offsets of its local cells
static __global__ void __launch_bounds__(128)
strip_tile_kernel(const int* __restrict__ cell_start, int W, int H, int nstrips, int nty,
                  const int* __restrict__ srow_start, int2* __restrict__ tile_info, uint16_t* __restrict__ tile_cs) {
  const int tile = blockIdx.x;
  const int s = tile / nty, ty = tile - s * nty;
  const int y0 = ty * ST_TH - 1;  // grid row of local row 0
  const int ya = max(y0, 0), yb = min(y0 + ST_LH - 1, H - 1);
  const int t0 = __ldg(srow_start + s * H + ya), t1 = __ldg(srow_start + s * H + yb + 1);
  const int len = t1 - t0;
  if (threadIdx.x == 0) tile_info[tile] = make_int2(t0, len);
  const int xa = max(s * ST_TW - 1, 0), xb = min(s * ST_TW + ST_TW, W - 1);
  uint16_t* cs = tile_cs + (long long)tile * ST_CS_STRIDE;
  for (int c = threadIdx.x; c < ST_CS_STRIDE; c += blockDim.x) {
    int off;
    if (c >= ST_NCELL) {
      off = len;
    } else {
      const int ly = c / ST_LW, lx = c - ly * ST_LW;
      const int y = y0 + ly, x = s * ST_TW - 1 + lx;
      if (y < 0) off = 0;
      else if (y >= H) off = len;
      else {
        const int rb = __ldg(srow_start + s * H + y) - t0;
        const long long base = (long long)y * W;
        if (x < xa) off = rb;
        else if (x > xb) off = rb + (__ldg(cell_start + base + xb + 1) - __ldg(cell_start + base + xa));
        else off = rb + (__ldg(cell_start + base + x) - __ldg(cell_start + base + xa));
      }
    }
    cs[c] = (uint16_t)min(off, 65535);  // tiles longer than ST_CAP are never staged
  }
}

// ---------------------------------------------------------------------------- query bucketing (one pass)
__device__ __forceinline__ int strip_tile_of(const GridDev& g, int nty, double x, double y) {
  const int cx = grid_cell(g, 0, x), cy = grid_cell(g, 1, y);
  return (cx / ST_TW) * nty + cy / ST_TH;
}

// ctrl[0], ctrl[1]: overflow counters; ctrl[ST_CNT_STRIDE * (1 + tile * ST_SUB + sub)]: queries in sub-bucket `sub`.
// Sub-bucket (t, sub): slots [(t*ST_SUB + sub)*Cs, +Cs) of qsorted / qperm; queries beyond Cs go to the overflow list.
// One 32-byte record per bucketed query: ONE full-sector store.  Separate 16-byte (x, y) and 4-byte (index) stores
// are partial-sector writes into a sparsely filled array -- 44 us for 2^20 queries against 25 us for the record
// (profiles/microbench/qbucket_variants.cu); the consumer also gets its query in one stream.
struct __align__(32) StripQRec {
  double x, y;
  long long qi;
  long long pad;
};
__device__ __forceinline__ void strip_store_rec(StripQRec* dst, double x, double y, long long qi) {
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(x), "d"(y), "l"(qi), "l"(0LL) : "memory");
}

#ifndef POMP_ST_QPT
#define POMP_ST_QPT 4
#endif
constexpr int ST_QPT = POMP_ST_QPT;  // queries per thread: independent load -> atomic -> store chains in flight
static __global__ void __launch_bounds__(256)
strip_qbucket_kernel(GridDev g, int nty, const double2* __restrict__ q, long long nq, int Cs, int* __restrict__ ctrl,
                     StripQRec* __restrict__ qrec, int* __restrict__ ovf_list) {
  const long long i0 = (long long)blockIdx.x * (blockDim.x * ST_QPT) + threadIdx.x;
  double2 v[ST_QPT];
  long long b[ST_QPT];
  int slot[ST_QPT];
#pragma unroll
  for (int u = 0; u < ST_QPT; u++) {
    const long long i = i0 + (long long)u * blockDim.x;
    v[u] = i < nq ? __ldg(q + i) : make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int u = 0; u < ST_QPT; u++) {
    const long long i = i0 + (long long)u * blockDim.x;
    const int t = strip_tile_of(g, nty, v[u].x, v[u].y);
    b[u] = (long long)t * ST_SUB + ((threadIdx.x + u) & (ST_SUB - 1));
    slot[u] = i < nq ? atomicAdd(ctrl + ST_CNT_STRIDE * (1 + b[u]), 1) : 0;
  }
#pragma unroll
  for (int u = 0; u < ST_QPT; u++) {
    const long long i = i0 + (long long)u * blockDim.x;
    if (i >= nq) continue;
    if (slot[u] < Cs) strip_store_rec(qrec + b[u] * Cs + slot[u], v[u].x, v[u].y, i);
    else ovf_list[atomicAdd(ctrl + 1, 1)] = (int)i;  // bucket full: straight to the pruned traversal
  }
}

// ---------------------------------------------------------------------------- the streaming search
struct StripSmem {
  uint64_t full[ST_MAX_STAGES];   // header + points + cell offsets of the stage have landed
  uint64_t empty[ST_MAX_STAGES];  // every consumer warp is done with the stage
  uint64_t hdr[ST_MAX_STAGES];    // the stage header alone is written (lets consumers fetch queries ahead)
  int uses[ST_MAX_STAGES];        // producer only: hand-overs to each group so far
  int ocount;                     // deferred queries of this CTA (uncertified / ambiguous), answered at its tail
  int olist[ST_OLIST];
};
constexpr int ST_SMEM_HEAD = 4608;
__host__ __device__ inline size_t strip_stage_bytes(int cap) { return ST_HDR_BYTES + (size_t)cap * 16 + ST_CS_BYTES; }

// slot of the t-th query of a tile whose sub-bucket prefix sums are pre[] (shared memory header)
__device__ __forceinline__ long long strip_query_slot(const int* pre, int tile, int Cs, int t) {
  int seg = 0;
#pragma unroll
  for (int j = 1; j < ST_SUB; j++) seg += (t >= pre[j]) ? 1 : 0;
  return ((long long)tile * ST_SUB + seg) * Cs + (t - pre[seg]);
}

__global__ void __launch_bounds__(ST_THREADS, 1)
strip_nn_kernel(GridDev g, StripDev sd, const double* __restrict__ pts, const double2* __restrict__ q,
                const StripQRec* __restrict__ qrec, int* __restrict__ ctrl, int Cs, int cap, int ngroups,
                int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                int* __restrict__ ovf_list, int spin_ns, int probe) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StripSmem& sh = *reinterpret_cast<StripSmem*>(smem_raw);
  unsigned char* stage0 = smem_raw + ST_SMEM_HEAD;
  const size_t stage_bytes = strip_stage_bytes(cap);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ngroups; s++) {
      mbar_init(&sh.full[s], 1);
      mbar_init(&sh.empty[s], ST_WG);
      mbar_init(&sh.hdr[s], 1);
      sh.uses[s] = 0;
    }
    sh.ocount = 0;
    fence_mbar_init();
  }
  __syncthreads();
  grid_dependency_wait();      // launched programmatically behind the query bucketing: its counters and records
  grid_launch_dependents();    // the overflow pass may be scheduled; it waits for this grid to complete

  if (warp == ST_MAX_STAGES * ST_WG) {
    // ---------------- producer warp.  Metadata of 32 tiles is fetched at once (one lane each: the latency of the
    // counter / tile_info loads is paid once per batch, not once per tile); then the lanes whose tile gets a stage
    // take turns, in tile order: each hands its tile to the first consumer GROUP whose (single) stage buffer is
    // free -- header first (the group starts fetching its queries), then the two bulk copies.
    int gnext = 0;
    for (int base = 0;; base += 32) {
      const long long tl = (long long)blockIdx.x + (long long)(base + lane) * gridDim.x;
      const bool valid = tl < sd.ntiles;
      if (__ballot_sync(0xffffffffu, valid) == 0u) break;
      const int tile = valid ? (int)tl : 0;
      int pre[ST_SUB + 1];
      pre[0] = 0;
      int2 info = make_int2(0, 0);
      if (valid) {
        int c[ST_SUB];
#pragma unroll
        for (int j = 0; j < ST_SUB; j++) {
          // read and clear: the counters go back to the host side zeroed (no memset launch per step); plain
          // loads -- the read-only path must not see a location this kernel writes
          int* cp = ctrl + ST_CNT_STRIDE * (1 + (long long)tile * ST_SUB + j);
          c[j] = __ldcg(cp);
          if (c[j]) *cp = 0;
        }
        info = __ldg(sd.tile_info + tile);
#pragma unroll
        for (int j = 0; j < ST_SUB; j++) pre[j + 1] = pre[j] + min(c[j], Cs);
      } else {
#pragma unroll
        for (int j = 0; j < ST_SUB; j++) pre[j + 1] = 0;
      }
      const bool hasq = valid && pre[ST_SUB] > 0;
      const bool staged = hasq && info.y > 0 && info.y <= cap;
      // tiles with queries that cannot be staged (empty tile + halo, or more points than a stage holds)
      unsigned um = __ballot_sync(0xffffffffu, hasq && !staged);
      while (um) {
        const int src = __ffs(um) - 1;
        um &= um - 1;
        const int utile = __shfl_sync(0xffffffffu, tile, src);
#pragma unroll
        for (int j = 0; j < ST_SUB; j++) {
          const int cnt = __shfl_sync(0xffffffffu, pre[j + 1] - pre[j], src);
          for (int t = lane; t < cnt; t += 32)
            ovf_list[atomicAdd(ctrl + 1, 1)] = (int)__ldg(&qrec[((long long)utile * ST_SUB + j) * Cs + t].qi);
        }
      }
      const unsigned sm = __ballot_sync(0xffffffffu, staged);
      for (unsigned m = sm; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        if (lane == src) {
          // first free group, round robin from the one after the last hand-over
          int gsel = gnext;
          while (true) {
            const int use = sh.uses[gsel];
            if (use == 0 || (probe ? mbar_test_wait(&sh.empty[gsel], (use - 1) & 1)
                                   : mbar_try_wait(&sh.empty[gsel], (use - 1) & 1)))
              break;
            gsel = gsel + 1 == ngroups ? 0 : gsel + 1;
          }
          sh.uses[gsel] += 1;
          unsigned char* buf = stage0 + (size_t)gsel * stage_bytes;
          int* hdr = reinterpret_cast<int*>(buf);
          const int s = tile / sd.nty;
          hdr[0] = tile;
          hdr[1] = info.x;
          hdr[2] = s * ST_TW - 1;                   // grid cell of local cell (0, 0)
          hdr[3] = (tile - s * sd.nty) * ST_TH - 1;
          hdr[4 + ST_SUB + 1] = 0;  // an opaque zero, see the consumer
#pragma unroll
          for (int j = 0; j <= ST_SUB; j++) hdr[4 + j] = pre[j];
          mbar_arrive(&sh.hdr[gsel]);
          const uint32_t pbytes = (uint32_t)info.y * 16u;
          mbar_expect_tx(&sh.full[gsel], pbytes + ST_CS_BYTES);
          bulk_g2s(buf + ST_HDR_BYTES, sd.spts2 + info.x, pbytes, &sh.full[gsel]);
          bulk_g2s(buf + ST_HDR_BYTES + (size_t)cap * 16, sd.tile_cs + (long long)tile * ST_CS_STRIDE, ST_CS_BYTES,
                   &sh.full[gsel]);
          gnext = gsel + 1 == ngroups ? 0 : gsel + 1;
        }
        __syncwarp();  // sh.uses is passed from lane to lane
        gnext = __shfl_sync(0xffffffffu, gnext, src);
      }
    }
    if (lane == 0) {  // end marker for every group: a header with tile = -1
      for (int gsel = 0; gsel < ngroups; gsel++) {
        const int use = sh.uses[gsel];
        if (use > 0) mbar_wait(&sh.empty[gsel], (use - 1) & 1);
        int* hdr = reinterpret_cast<int*>(stage0 + (size_t)gsel * stage_bytes);
        hdr[0] = -1;
        mbar_arrive(&sh.hdr[gsel]);
      }
    }
    return;
  }

  // ---------------- consumers: group = warp / ST_WG owns stage buffer `group`; its ST_WG warps split the passes
  const int group = warp / ST_WG, wg = warp - group * ST_WG;
  if (group >= ngroups) return;
  const int W = g.ncell[0], H = g.ncell[1];
  const unsigned char* buf = stage0 + (size_t)group * stage_bytes;
  const int* hdr = reinterpret_cast<const int*>(buf);
  const double2* __restrict__ P = reinterpret_cast<const double2*>(buf + ST_HDR_BYTES);
  const uint16_t* __restrict__ CS = reinterpret_cast<const uint16_t*>(buf + ST_HDR_BYTES + (size_t)cap * 16);
  // the result of a pass is stored one pass later: its insertion-index gather is in flight meanwhile
  bool pend = false;
  int pend_qi = 0, pend_idx = 0;
  double pend_d = 0.0;
  auto flush = [&]() {
    if (pend) {
      out_idx[pend_qi] = (int64_t)pend_idx;
      out_dist[pend_qi] = pend_d;
      if (out_cnt) out_cnt[pend_qi] = 1;
    }
    pend = false;
  };
  // Deferred queries (no certificate from the 3x3 block, or a runner-up that could tie after rounding) are parked
  // in a CTA-wide list and answered at the tail of the kernel.  (Answering them whenever a warp waits for its
  // group's next tile was measured slower, 0.135 vs 0.122 ms per step: a 3 us warp-wide search delays the group.)
  for (unsigned use = 0;; use++) {
    while (!mbar_try_wait(&sh.hdr[group], use & 1))
      if (spin_ns) __nanosleep(spin_ns);  // waiting warps share their scheduler with the producer warp
    const int tile = hdr[0];
    if (tile < 0) break;
    const int start = hdr[1], x0 = hdr[2], y0 = hdr[3];
    const int* pre = hdr + 4;
    const int qn = pre[ST_SUB];
    // queries of this warp's first pass: on their way while the bulk copies land
    int t = (wg << 5) + lane;
    double2 v = make_double2(0.0, 0.0);
    int qi = 0;
    if (t < qn) {
      const double2* rec = reinterpret_cast<const double2*>(qrec + strip_query_slot(pre, tile, Cs, t));
      v = __ldg(rec);
      qi = (int)__double_as_longlong(__ldg(rec + 1).x);
    }
    while (!mbar_try_wait(&sh.full[group], use & 1))
      if (spin_ns) __nanosleep(spin_ns);
    for (int tb = wg << 5; tb < qn; tb += ST_WG * 32) {
      t = tb + lane;
      const bool active = t < qn;
      if (tb != (wg << 5) && active) {
        const double2* rec = reinterpret_cast<const double2*>(qrec + strip_query_slot(pre, tile, Cs, t));
        v = __ldg(rec);
        qi = (int)__double_as_longlong(__ldg(rec + 1).x);
      }
      double bsq = POMP_INF, sec = POMP_INF;
      // bpos starts as (qi & 0) with a zero the compiler cannot see through: qi's load is thereby CONSUMED here,
      // before the candidate loop.  Otherwise its first use is next to the insertion-index gather of the epilogue,
      // ptxas gives both loads the same scoreboard, and the move of qi waits for the gather it was meant to
      // overlap (17 % of all stall samples, profiles/r2_strip_v3).
      int bpos = qi & hdr[4 + ST_SUB + 1], cx = 0, cy = 0;
      if (active) {
        cx = grid_cell(g, 0, v.x);
        cy = grid_cell(g, 1, v.y);
        const int lx = cx - x0, ly = cy - y0;  // in [1, TW] x [1, TH]
        const int c0 = (ly - 1) * ST_LW + lx - 1;
        const int a0 = CS[c0], b0 = CS[c0 + 3];
        const int a1 = CS[c0 + ST_LW], b1 = CS[c0 + ST_LW + 3];
        const int a2 = CS[c0 + 2 * ST_LW], b2 = CS[c0 + 2 * ST_LW + 3];
        const int n0 = b0 - a0, n01 = n0 + (b1 - a1), total = n01 + (b2 - a2);
        const int off0 = a0, off1 = a1 - n0, off2 = a2 - n01;
        for (int i = 0; i < total; i += 4) {
          double2 pv[4];
          int pp[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int ii = min(i + u, total - 1);
            pp[u] = ii + (ii < n0 ? off0 : (ii < n01 ? off1 : off2));
            pv[u] = P[pp[u]];
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const double dx = pv[u].x - v.x, dy = pv[u].y - v.y;
            double sq = 0.0;
            sq = sq + dx * dx;
            sq = sq + dy * dy;
            if (i + u >= total) sq = POMP_INF;  // padding of the last group: never a winner, never a runner-up
            // smallest and second smallest by compare + select (sm_100 has no FP64 min/max instruction: fmin / fmax
            // expand to DSETP + selects + NaN fix-ups).  NaN never wins and never becomes the runner-up, like
            // `d < dbest` in the reference
            const bool lt = sq < bsq;
            const double hi = lt ? bsq : sq;
            sec = hi < sec ? hi : sec;
            bsq = lt ? sq : bsq;
            bpos = lt ? pp[u] : bpos;
          }
        }
      }
      if (tb + ST_WG * 32 >= qn) {  // this warp's last pass over the stage: hand the buffer back before the epilogue
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty[group]);
      }
      flush();
      bool defer = false;
      if (active) {
        const double d = sqrt(bsq);
        double margin = POMP_INF;
        int lo_c = cx - 1, hi_c = cx + 1;
        if (lo_c > 0) margin = fmin(margin, v.x - ((g.lo[0] + (double)lo_c * g.h[0]) + g.slack[0]));
        if (hi_c < W - 1) margin = fmin(margin, ((g.lo[0] + (double)(hi_c + 1) * g.h[0]) - g.slack[0]) - v.x);
        lo_c = cy - 1;
        hi_c = cy + 1;
        if (lo_c > 0) margin = fmin(margin, v.y - ((g.lo[1] + (double)lo_c * g.h[1]) + g.slack[1]));
        if (hi_c < H - 1) margin = fmin(margin, ((g.lo[1] + (double)(hi_c + 1) * g.h[1]) - g.slack[1]) - v.y);
        // certified: nothing outside the block can tie or beat d; unique: no other candidate can round to d
        const bool ok = (bsq < POMP_INF) && (d * (1.0 + 1e-9) < margin) && (sec > tsq_of(d));
        if (ok) {
          pend = true;
          pend_qi = qi;
          pend_d = d;
          pend_idx = __ldg(sd.sidx2 + start + bpos);
        } else {
          defer = true;
        }
      }
      if (defer) {
        const int o = atomicAdd(&sh.ocount, 1);
        if (o < ST_OLIST) sh.olist[o] = qi;
        else ovf_list[atomicAdd(ctrl + 1, 1)] = qi;
      }
    }
    if ((wg << 5) >= qn) {  // no pass for this warp on this tile (fewer than wg*32 queries)
      if (lane == 0) mbar_arrive(&sh.empty[group]);
    }
  }
  flush();
  // ---------------- tail: the CTA's deferred queries, dealt round robin to its consumer warps; ONE WARP per query
  // searches the 5x5 block of the row-major index with the sequential (distance, insertion index) rule; what that
  // cannot certify goes to the global list (pruned traversal).  As separate launches the two overflow passes cost
  // 11 us of latency for ~0.3 % of the queries.
  const int ncw = ngroups * ST_WG;
  asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
  const int no = min(sh.ocount, ST_OLIST);
  for (int e = warp; e < no; e += ncw) {
    const int oq = sh.olist[e];
    const double2 ov = __ldg(q + oq);
    const int ocx = grid_cell(g, 0, ov.x), ocy = grid_cell(g, 1, ov.y);
    warp_retry_nn2d(g, pts, (long long)g.n_indexed, nullptr, lane == 0, ov.x, ov.y, ocx, ocy, (long long)oq, 2,
                    out_idx, out_dist, out_cnt, ovf_list, ctrl + 1);
  }
}

}  // namespace pomp
