// shard.cu -- device helpers of the SPATIALLY sharded nearest-neighbour path (sharded.py):
//   * route: counting sort of a query batch by owning rank (slab of one coordinate), so that ONE all-to-all
//     moves every query to the rank that holds its neighbourhood;
//   * finalize: local k-NN answers -> global insertion indices, packed 16-byte (distance, index) records for the
//     return all-to-all, plus the exactness certificate of the slab (k-th distance < distance to the outside of the
//     region the rank covers, halo included); halo copies can be blanked for the all-ranks fallback;
//   * unpack: records back into the caller's (idx, dist, count) arrays at the original query positions.
// The collectives themselves are issued by the host binding through torch.distributed (NCCL over NVLink).
#include <algorithm>

#include "common.cuh"

using namespace pomp;

namespace {

constexpr int SHARD_MAX_RANKS = 64;
constexpr int ROUTE_QPT = 8;  // queries per thread of the peer-memory routing kernel

struct Bounds {
  double b[SHARD_MAX_RANKS];  // b[r] = lower bound of rank r's slab (b[0] = -inf is implied), n_ranks - 1 used
};

__device__ __forceinline__ int owner_of(const Bounds& B, int n_ranks, double y) {
  int r = 0;  // slabs are [b[r-1], b[r]): first r with y < b[r]; NaN goes to the last rank
  while (r < n_ranks - 1 && !(y < B.b[r])) r++;
  return r;
}

__global__ void __launch_bounds__(256)
route_count_kernel(const double* __restrict__ q, long long nq, int D, int axis, Bounds B, int n_ranks,
                   int* __restrict__ dest, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int sc[SHARD_MAX_RANKS];
  if (threadIdx.x < SHARD_MAX_RANKS) sc[threadIdx.x] = 0;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) {
    const int r = owner_of(B, n_ranks, __ldg(q + i * D + axis));
    dest[i] = r;
    atomicAdd(&sc[r], 1u);
  }
  __syncthreads();
  if (threadIdx.x < n_ranks && sc[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
}

// cursor[r] starts at the first slot of rank r; blocks reserve contiguous pieces
__global__ void __launch_bounds__(256)
route_scatter_kernel(const double* __restrict__ q, long long nq, int D, const int* __restrict__ dest, int n_ranks,
                     unsigned long long* __restrict__ cursor, double* __restrict__ q_out, int64_t* __restrict__ perm) {
  __shared__ unsigned int sc[SHARD_MAX_RANKS];
  __shared__ unsigned long long base[SHARD_MAX_RANKS];
  if (threadIdx.x < SHARD_MAX_RANKS) sc[threadIdx.x] = 0;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int r = 0;
  unsigned int slot = 0;
  if (i < nq) {
    r = __ldg(dest + i);
    slot = atomicAdd(&sc[r], 1u);
  }
  __syncthreads();
  if (threadIdx.x < n_ranks && sc[threadIdx.x]) base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
  __syncthreads();
  if (i < nq) {
    const long long t = (long long)(base[r] + slot);
    for (int j = 0; j < D; j++) q_out[t * D + j] = __ldg(q + i * D + j);
    perm[t] = i;
  }
}

// rec[(i*k + j)*2 + 0] = distance bits, [+1] = global index (-1: none).  flags[i] = 1 when query i has no
// certificate (its k-th distance reaches outside [lo_cov, hi_cov] on `axis`, or fewer than k hits).
__global__ void __launch_bounds__(256)
finalize_kernel(const double* __restrict__ q, long long nq, int D, int axis, double lo_cov, double hi_cov, int k,
                const int64_t* __restrict__ idx_local, const double* __restrict__ dist,
                const int64_t* __restrict__ l2g, const uint8_t* __restrict__ halo, int drop_halo,
                int64_t* __restrict__ rec, uint8_t* __restrict__ flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  double kth = POMP_INF;
  bool full = true;
  for (int j = 0; j < k; j++) {
    const int64_t li = idx_local[i * k + j];
    double d = dist[i * k + j];
    int64_t gi = -1;
    if (li >= 0 && !(drop_halo && halo && halo[li])) gi = l2g ? l2g[li] : li;
    if (li < 0) full = false;
    if (gi < 0) d = POMP_INF;
    if (li >= 0) kth = dist[i * k + j];  // lists are sorted: the last finite entry is the k-th (or the last one)
    rec[(i * k + j) * 2] = __double_as_longlong(d);
    rec[(i * k + j) * 2 + 1] = gi;
  }
  if (flags) {
    const double y = q[i * D + axis];
    // every point within kth of the query lies inside the covered region  <=>  exact (ties included: a point
    // outside is strictly farther than the margin, hence than kth)
    const double margin = fmin(y - lo_cov, hi_cov - y);
    flags[i] = (full && kth < margin) ? 0 : 1;
  }
}

__global__ void __launch_bounds__(256)
unpack_kernel(const int64_t* __restrict__ rec, long long nq, int k, const int64_t* __restrict__ perm,
              int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const long long o = perm ? perm[i] : i;
  int c = 0;
  for (int j = 0; j < k; j++) {
    const int64_t gi = rec[(i * k + j) * 2 + 1];
    out_idx[o * k + j] = gi;
    out_dist[o * k + j] = __longlong_as_double(rec[(i * k + j) * 2]);
    c += gi >= 0;
  }
  if (out_cnt) out_cnt[o] = c;
}

// ---- sync-free variant: fixed capacity per destination rank (equal-split all-to-all, no counts on the host).
// q_out [n_ranks][cap][D] and perm_out [n_ranks][cap] arrive pre-filled with NaN / -1; queries that do not fit their
// block (or are NaN on the axis) go to the deferred list.
__global__ void __launch_bounds__(256)
route_fixed_kernel(const double* __restrict__ q, long long nq, int D, int axis, Bounds B, int n_ranks, long long cap,
                   unsigned long long* __restrict__ cursor, double* __restrict__ q_out, int64_t* __restrict__ perm,
                   int64_t* __restrict__ defer_list, unsigned long long* __restrict__ defer_count, long long defer_cap) {
  __shared__ unsigned int sc[SHARD_MAX_RANKS];
  __shared__ unsigned long long base[SHARD_MAX_RANKS];
  if (threadIdx.x < SHARD_MAX_RANKS) sc[threadIdx.x] = 0;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int r = 0;
  unsigned int slot = 0;
  if (i < nq) {
    r = owner_of(B, n_ranks, __ldg(q + i * D + axis));
    slot = atomicAdd(&sc[r], 1u);
  }
  __syncthreads();
  if (threadIdx.x < n_ranks && sc[threadIdx.x]) base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
  __syncthreads();
  if (i < nq) {
    const unsigned long long t = base[r] + slot;
    if (t < (unsigned long long)cap) {
      const long long o = (long long)r * cap + (long long)t;
      for (int j = 0; j < D; j++) q_out[o * D + j] = __ldg(q + i * D + j);
      perm[o] = i;
    } else {
      const unsigned long long d = atomicAdd(defer_count, 1ull);
      if ((long long)d < defer_cap) defer_list[d] = i;
    }
  }
}

// records with a row stride (row = k records of 2 words + 1 flag word) for one packed return exchange
__global__ void __launch_bounds__(256)
finalize_rows_kernel(const double* __restrict__ q, long long nq, int D, int axis, double lo_cov, double hi_cov, int k,
                     const int64_t* __restrict__ idx_local, const double* __restrict__ dist,
                     const int64_t* __restrict__ l2g, const uint8_t* __restrict__ halo, int drop_halo,
                     int64_t* __restrict__ rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int64_t* row = rows + i * (2 * k + 1);
  double kth = POMP_INF;
  bool full = true;
  for (int j = 0; j < k; j++) {
    const int64_t li = idx_local[i * k + j];
    double d = dist[i * k + j];
    int64_t gi = -1;
    if (li >= 0 && !(drop_halo && halo[li])) gi = l2g[li];
    if (li < 0) full = false;
    if (gi < 0) d = POMP_INF;
    if (li >= 0) kth = dist[i * k + j];
    row[2 * j] = __double_as_longlong(d);
    row[2 * j + 1] = gi;
  }
  const double y = q[i * D + axis];
  const double margin = fmin(y - lo_cov, hi_cov - y);
  row[2 * k] = (full && kth < margin) ? 0 : 1;  // padding queries (NaN) are flagged too; their perm is -1
}

// rows back at the source: write certified answers at perm[i]; collect the rest (flag set) in the deferred list
__global__ void __launch_bounds__(256)
unpack_rows_kernel(const int64_t* __restrict__ rows, long long n, int k, const int64_t* __restrict__ perm,
                   int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int64_t* __restrict__ defer_list,
                   unsigned long long* __restrict__ defer_count, long long defer_cap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long o = perm[i];
  if (o < 0) return;  // padding slot
  const int64_t* row = rows + i * (2 * k + 1);
  if (row[2 * k]) {
    const unsigned long long d = atomicAdd(defer_count, 1ull);
    if ((long long)d < defer_cap) defer_list[d] = o;
    return;
  }
  for (int j = 0; j < k; j++) {
    out_idx[o * k + j] = row[2 * j + 1];
    out_dist[o * k + j] = __longlong_as_double(row[2 * j]);
  }
}

// deferred queries -> a fixed-size, NaN-padded block for the all-gather
__global__ void __launch_bounds__(256)
gather_deferred_kernel(const double* __restrict__ q, int D, const int64_t* __restrict__ defer_list,
                       const unsigned long long* __restrict__ defer_count, long long F, const double* __restrict__ pad,
                       double* __restrict__ qf) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= F) return;
  const long long cnt = (long long)min((unsigned long long)F, *defer_count);
  // unused slots carry an ordinary point (a NaN query would make every rank scan its whole shard for nothing)
  for (int c = 0; c < D; c++) qf[j * D + c] = j < cnt ? q[defer_list[j] * D + c] : pad[c];
}

// merged fallback answers -> the caller's arrays at the deferred positions
__global__ void __launch_bounds__(256)
scatter_deferred_kernel(const int64_t* __restrict__ midx, const double* __restrict__ mdist, int k,
                        const int64_t* __restrict__ defer_list, const unsigned long long* __restrict__ defer_count,
                        long long F, int64_t* __restrict__ out_idx, double* __restrict__ out_dist) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cnt = (long long)min((unsigned long long)F, *defer_count);
  if (j >= cnt) return;
  const long long o = defer_list[j];
  for (int c = 0; c < k; c++) {
    out_idx[o * k + c] = midx[j * k + c];
    out_dist[o * k + c] = mdist[j * k + c];
  }
}

// ---- peer-memory variant (NVLink loads/stores instead of NCCL): every rank maps the receive buffers of its peers
// (torch symmetric memory).  The route kernel stores each query STRAIGHT into the owner's receive block
// rq[dest][src_rank][slot]; after the owner's search, the finalize kernel stores the answer row straight into the
// source's result block rb[src][my_rank][slot].  Two device barriers per step, no collective on the data path.
struct PeerPtrs {
  void* p[SHARD_MAX_RANKS];
};

__global__ void __launch_bounds__(256)
route_p2p_kernel(const double* __restrict__ q, long long nq, int D, int axis, Bounds B, int n_ranks, int my_rank,
                 long long cap, unsigned long long* __restrict__ cursor, PeerPtrs rq, int64_t* __restrict__ perm,
                 int64_t* __restrict__ defer_list, unsigned long long* __restrict__ defer_count, long long defer_cap,
                 int qpt) {
  // A block takes 256 * qpt queries, sorts them by destination rank in SHARED memory and then writes one contiguous
  // run per destination: the NVLink stores are full 128-byte lines (one query per thread wrote 16-byte pieces in
  // smem-atomic order -- 0.11 ms for 2^20 queries on 8 GPUs), and a block needs ONE global atomic per destination.
  extern __shared__ __align__(16) unsigned char route_smem[];
  __shared__ unsigned int sc[SHARD_MAX_RANKS], pre[SHARD_MAX_RANKS + 1];
  __shared__ unsigned long long base[SHARD_MAX_RANKS];
  const int nQ = blockDim.x * qpt;
  double* sq = reinterpret_cast<double*>(route_smem);                       // [nQ][D]
  long long* si = reinterpret_cast<long long*>(sq + (size_t)nQ * D);         // [nQ] original index
  unsigned char* sr = reinterpret_cast<unsigned char*>(si + nQ);             // [nQ] destination rank
  if (threadIdx.x < SHARD_MAX_RANKS) sc[threadIdx.x] = 0;
  __syncthreads();
  const long long i0 = (long long)blockIdx.x * nQ + threadIdx.x;
  int r[ROUTE_QPT];
  unsigned int slot[ROUTE_QPT];
#pragma unroll
  for (int u = 0; u < ROUTE_QPT; u++) {
    const long long i = i0 + (long long)u * blockDim.x;
    r[u] = -1;
    slot[u] = 0;
    if (u < qpt && i < nq) r[u] = owner_of(B, n_ranks, __ldg(q + i * D + axis));
    // warp-aggregated: one shared-memory atomic per destination and warp instead of one per query (8 hot counters)
    const unsigned peers = __match_any_sync(0xffffffffu, r[u]);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    unsigned int first = 0;
    if (lane == leader && r[u] >= 0) first = atomicAdd(&sc[r[u]], (unsigned)__popc(peers));
    first = __shfl_sync(0xffffffffu, first, leader);
    slot[u] = first + (unsigned)__popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int acc = 0;
    for (int j = 0; j < n_ranks; j++) {
      pre[j] = acc;
      acc += sc[j];
    }
    pre[n_ranks] = acc;
  }
  if (threadIdx.x < n_ranks && sc[threadIdx.x]) base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
  __syncthreads();
#pragma unroll
  for (int u = 0; u < ROUTE_QPT; u++) {
    if (r[u] < 0) continue;
    const long long i = i0 + (long long)u * blockDim.x;
    const unsigned int e = pre[r[u]] + slot[u];
    for (int j = 0; j < D; j++) sq[(size_t)e * D + j] = __ldg(q + i * D + j);
    si[e] = i;
    sr[e] = (unsigned char)r[u];
  }
  __syncthreads();
  const unsigned int total = pre[n_ranks];
  // coordinates: consecutive threads write consecutive doubles of one destination's run
  for (unsigned int w = threadIdx.x; w < total * (unsigned)D; w += blockDim.x) {
    const unsigned int e = w / (unsigned)D, j = w - e * (unsigned)D;
    const int rr = sr[e];
    const unsigned long long t = base[rr] + (e - pre[rr]);
    if (t < (unsigned long long)cap)
      reinterpret_cast<double*>(rq.p[rr])[((long long)my_rank * cap + (long long)t) * D + j] = sq[w];  // peer store
  }
  for (unsigned int e = threadIdx.x; e < total; e += blockDim.x) {
    const int rr = sr[e];
    const unsigned long long t = base[rr] + (e - pre[rr]);
    if (t < (unsigned long long)cap) {
      perm[(long long)rr * cap + (long long)t] = si[e];
    } else {
      const unsigned long long d = atomicAdd(defer_count, 1ull);
      if ((long long)d < defer_cap) defer_list[d] = si[e];
    }
  }
}

// local answers for slot (src, t) -> row in the SOURCE rank's result block rb[src] at (my_rank, t)
__global__ void __launch_bounds__(256)
finalize_p2p_kernel(const double* __restrict__ q, long long n_slots, long long cap, int my_rank, int D, int axis,
                    double lo_cov, double hi_cov, int k, const int64_t* __restrict__ idx_local,
                    const double* __restrict__ dist, const int64_t* __restrict__ l2g, PeerPtrs rb,
                    unsigned long long* __restrict__ fails) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const long long src = i / cap, t = i - src * cap;
  // word-major block [2k][n_ranks*cap]: consecutive lanes store consecutive 8-byte words (coalesced NVLink stores)
  int64_t* row = reinterpret_cast<int64_t*>(rb.p[src]) + ((long long)my_rank * cap + t);
  const long long ws = n_slots;
  double kth = POMP_INF;
  bool full = true;
  double d0 = POMP_INF;
  int64_t g0 = -1;
  for (int j = 0; j < k; j++) {
    const int64_t li = idx_local[i * k + j];
    double d = dist[i * k + j];
    const int64_t gi = li >= 0 ? l2g[li] : -1;
    if (li < 0) {
      full = false;
      d = POMP_INF;
    } else {
      kth = d;
    }
    if (j > 0) {  // entry 0 is stored last: its index word carries the certificate
      row[(2 * j) * ws] = __double_as_longlong(d);
      row[(2 * j + 1) * ws] = gi;
    } else {
      d0 = d;
      g0 = gi;
    }
  }
  const double y = q[i * D + axis];
  const double margin = fmin(y - lo_cov, hi_cov - y);
  // no certificate (the k-th distance reaches beyond the region this rank covers, or fewer than k points): the index
  // word of entry 0 says so (-2; a missing neighbour is -1) -- two words per neighbour over NVLink instead of 2k + 1
  const bool certified = full && kth < margin;
  row[0] = __double_as_longlong(d0);
  row[ws] = certified ? g0 : -2;
  if (!certified && fails) atomicAdd(fails + src, 1ull);  // per asking rank (rare)
}

// counts -> the same place in EVERY peer's flag block (thread p stores to peer p)
__global__ void share_counts_kernel(const unsigned long long* __restrict__ vals, int ncols, PeerPtrs fl, long long dst_off,
                                    int n_ranks) {
  const int p = threadIdx.x;
  if (p >= n_ranks) return;
  int64_t* dst = reinterpret_cast<int64_t*>(fl.p[p]) + dst_off;
  for (int a = 0; a < ncols; a++) dst[a] = (int64_t)vals[a];
}

// flags[rows][cols]: column sums (per asking rank), their maximum -> *worst
__global__ void worst_count_kernel(const int64_t* __restrict__ flags, int rows, int cols, int64_t* __restrict__ worst) {
  long long best = 0;
  for (int a = threadIdx.x; a < cols; a += 32) {
    long long sum = 0;
    for (int r = 0; r < rows; r++) sum += flags[(long long)r * cols + a];
    best = sum > best ? sum : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
    best = ob > best ? ob : best;
  }
  if (threadIdx.x == 0) *worst = best;
}

// result block [2k][n_ranks][cap] (word-major; allocated with 2k+1 rows) -> the caller's arrays; slot (r, t) is valid for t < min(cursor[r], cap)
__global__ void __launch_bounds__(256)
unpack_p2p_kernel(const int64_t* __restrict__ rows, int n_ranks, long long cap, int k, const int64_t* __restrict__ perm,
                  const unsigned long long* __restrict__ cursor, int64_t* __restrict__ out_idx,
                  double* __restrict__ out_dist, int64_t* __restrict__ defer_list,
                  unsigned long long* __restrict__ defer_count, long long defer_cap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_ranks * cap) return;
  const long long r = i / cap, t = i - r * cap;
  if ((unsigned long long)t >= cursor[r]) return;
  const long long o = perm[i];
  const long long ws = (long long)n_ranks * cap;
  const int64_t* row = rows + i;
  if (row[ws] == -2) {  // the owner could not certify this query (finalize_p2p_kernel)
    const unsigned long long d = atomicAdd(defer_count, 1ull);
    if ((long long)d < defer_cap) defer_list[d] = o;
    return;
  }
  for (int j = 0; j < k; j++) {
    out_idx[o * k + j] = row[(2 * j + 1) * ws];
    out_dist[o * k + j] = __longlong_as_double(row[(2 * j) * ws]);
  }
}

}  // namespace

POMP_API int pomp_shard_route_p2p(const double* q, int64_t nq, int32_t dim, int32_t axis, const double* bounds_h,
                                  int32_t n_ranks, int32_t my_rank, int64_t cap, uint64_t* counters,
                                  const uint64_t* peer_query_blocks_h, int64_t* perm_out, int64_t* defer_list,
                                  int64_t defer_cap, void* stream) {
  POMP_REQUIRE(n_ranks >= 1 && n_ranks <= SHARD_MAX_RANKS && dim >= 1 && dim <= POMP_MAX_DIM && axis >= 0 && axis < dim &&
                   cap >= 0 && my_rank >= 0 && my_rank < n_ranks, "bad arguments to pomp_shard_route_p2p");
  if (nq == 0) return POMP_OK;
  POMP_REQUIRE(q && counters && peer_query_blocks_h && perm_out && defer_list, "NULL argument to pomp_shard_route_p2p");
  Bounds B;
  PeerPtrs P;
  for (int r = 0; r < SHARD_MAX_RANKS; r++) {
    B.b[r] = (r < n_ranks - 1 && bounds_h) ? bounds_h[r] : INFINITY;
    P.p[r] = r < n_ranks ? (void*)(uintptr_t)peer_query_blocks_h[r] : nullptr;
  }
  {
    // queries per thread: as many as 48 KB of staging (D doubles + index + rank per query) allow, at most ROUTE_QPT
    const size_t per_q = (size_t)dim * 8 + 8 + 1;
    int qpt = (int)std::min<size_t>(ROUTE_QPT, std::max<size_t>(1, (48 * 1024 - 64) / (256 * per_q)));
    const size_t smem = (size_t)256 * qpt * per_q + 16;
    const long long per_block = 256LL * qpt;
    LAUNCH(route_p2p_kernel, (int)((nq + per_block - 1) / per_block), 256, smem, (cudaStream_t)stream, q, (long long)nq, dim,
           axis, B, n_ranks, my_rank, (long long)cap, (unsigned long long*)counters, P, perm_out, defer_list,
           (unsigned long long*)(counters + SHARD_MAX_RANKS), (long long)defer_cap, qpt);
  }
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_finalize_p2p(const double* q, int64_t cap, int32_t n_ranks, int32_t my_rank, int32_t dim,
                                     int32_t axis, double lo_cov, double hi_cov, int32_t k, const int64_t* idx_local,
                                     const double* dist, const int64_t* local_to_global,
                                     const uint64_t* peer_result_blocks_h, uint64_t* fails_per_asker, void* stream) {
  POMP_REQUIRE(cap >= 0 && k >= 1 && dim >= 1 && axis >= 0 && axis < dim && n_ranks >= 1 && n_ranks <= SHARD_MAX_RANKS,
               "bad arguments to pomp_shard_finalize_p2p");
  if (cap == 0) return POMP_OK;
  POMP_REQUIRE(q && idx_local && dist && local_to_global && peer_result_blocks_h, "NULL argument to pomp_shard_finalize_p2p");
  PeerPtrs P;
  for (int r = 0; r < SHARD_MAX_RANKS; r++) P.p[r] = r < n_ranks ? (void*)(uintptr_t)peer_result_blocks_h[r] : nullptr;
  const long long n = (long long)n_ranks * cap;
  LAUNCH(finalize_p2p_kernel, (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream, q, n, (long long)cap, my_rank, dim,
         axis, lo_cov, hi_cov, k, idx_local, dist, local_to_global, P, (unsigned long long*)fails_per_asker);
  CHECK_LAUNCH();
  return POMP_OK;
}

// vals[ncols] (device) -> words [dst_offset, dst_offset + ncols) of every peer's flag block
POMP_API int pomp_shard_share_counts_p2p(const uint64_t* vals, int32_t ncols, const uint64_t* peer_flag_blocks_h,
                                         int64_t dst_offset, int32_t n_ranks, void* stream) {
  POMP_REQUIRE(vals && peer_flag_blocks_h && ncols >= 1 && dst_offset >= 0 && n_ranks >= 1 && n_ranks <= SHARD_MAX_RANKS,
               "bad arguments to pomp_shard_share_counts_p2p");
  PeerPtrs P;
  for (int r = 0; r < SHARD_MAX_RANKS; r++) P.p[r] = r < n_ranks ? (void*)(uintptr_t)peer_flag_blocks_h[r] : nullptr;
  LAUNCH(share_counts_kernel, 1, SHARD_MAX_RANKS, 0, (cudaStream_t)stream, (const unsigned long long*)vals, ncols, P,
         (long long)dst_offset, n_ranks);
  CHECK_LAUNCH();
  return POMP_OK;
}

// flags[rows][cols] (device): *worst = max over columns of the column sums
POMP_API int pomp_shard_worst_count(const int64_t* flags, int32_t rows, int32_t cols, int64_t* worst, void* stream) {
  POMP_REQUIRE(flags && worst && rows >= 1 && cols >= 1, "bad arguments to pomp_shard_worst_count");
  LAUNCH(worst_count_kernel, 1, 32, 0, (cudaStream_t)stream, flags, rows, cols, worst);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_unpack_p2p(const int64_t* rows, int32_t n_ranks, int64_t cap, int32_t k, const int64_t* perm,
                                   const uint64_t* counters, int64_t* idx, double* dist, int64_t* defer_list,
                                   int64_t defer_cap, void* stream) {
  POMP_REQUIRE(n_ranks >= 1 && n_ranks <= SHARD_MAX_RANKS && cap >= 0 && k >= 1, "bad arguments to pomp_shard_unpack_p2p");
  if (cap == 0) return POMP_OK;
  POMP_REQUIRE(rows && perm && counters && idx && dist && defer_list, "NULL argument to pomp_shard_unpack_p2p");
  const long long n = (long long)n_ranks * cap;
  LAUNCH(unpack_p2p_kernel, (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream, rows, n_ranks, (long long)cap, k, perm,
         (const unsigned long long*)counters, idx, dist, defer_list,
         (unsigned long long*)(const_cast<uint64_t*>(counters) + SHARD_MAX_RANKS), (long long)defer_cap);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_route_fixed(const double* q, int64_t nq, int32_t dim, int32_t axis, const double* bounds_h,
                                    int32_t n_ranks, int64_t cap, uint64_t* counters, double* q_out, int64_t* perm_out,
                                    int64_t* defer_list, int64_t defer_cap, void* stream) {
  POMP_REQUIRE(n_ranks >= 1 && n_ranks <= SHARD_MAX_RANKS && dim >= 1 && dim <= POMP_MAX_DIM && axis >= 0 && axis < dim &&
                   cap >= 0, "bad arguments to pomp_shard_route_fixed");
  if (nq == 0) return POMP_OK;
  POMP_REQUIRE(q && counters && q_out && perm_out && defer_list, "NULL argument to pomp_shard_route_fixed");
  Bounds B;
  for (int r = 0; r < SHARD_MAX_RANKS; r++) B.b[r] = (r < n_ranks - 1 && bounds_h) ? bounds_h[r] : INFINITY;
  // counters[0..63]: per-rank cursors, counters[64]: deferred count -- zeroed by the caller together with the fills
  LAUNCH(route_fixed_kernel, (int)((nq + 255) / 256), 256, 0, (cudaStream_t)stream, q, (long long)nq, dim, axis, B,
         n_ranks, (long long)cap, (unsigned long long*)counters, q_out, perm_out, defer_list,
         (unsigned long long*)(counters + SHARD_MAX_RANKS), (long long)defer_cap);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_finalize_rows(const double* q, int64_t nq, int32_t dim, int32_t axis, double lo_cov,
                                      double hi_cov, int32_t k, const int64_t* idx_local, const double* dist,
                                      const int64_t* local_to_global, const uint8_t* halo_flag, int32_t drop_halo,
                                      int64_t* rows, void* stream) {
  POMP_REQUIRE(nq >= 0 && k >= 1 && dim >= 1 && axis >= 0 && axis < dim, "bad arguments to pomp_shard_finalize_rows");
  if (nq == 0) return POMP_OK;
  POMP_REQUIRE(q && idx_local && dist && local_to_global && halo_flag && rows, "NULL argument to pomp_shard_finalize_rows");
  LAUNCH(finalize_rows_kernel, (int)((nq + 255) / 256), 256, 0, (cudaStream_t)stream, q, (long long)nq, dim, axis, lo_cov,
         hi_cov, k, idx_local, dist, local_to_global, halo_flag, drop_halo, rows);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_unpack_rows(const int64_t* rows, int64_t n, int32_t k, const int64_t* perm, int64_t* idx,
                                    double* dist, int64_t* defer_list, uint64_t* defer_count, int64_t defer_cap,
                                    void* stream) {
  POMP_REQUIRE(n >= 0 && k >= 1, "bad arguments to pomp_shard_unpack_rows");
  if (n == 0) return POMP_OK;
  POMP_REQUIRE(rows && perm && idx && dist && defer_list && defer_count, "NULL argument to pomp_shard_unpack_rows");
  LAUNCH(unpack_rows_kernel, (int)((n + 255) / 256), 256, 0, (cudaStream_t)stream, rows, (long long)n, k, perm, idx, dist,
         defer_list, (unsigned long long*)defer_count, (long long)defer_cap);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_gather_deferred(const double* q, int32_t dim, const int64_t* defer_list,
                                        const uint64_t* defer_count, int64_t F, const double* pad_point, double* qf,
                                        void* stream) {
  POMP_REQUIRE(q && defer_list && defer_count && qf && pad_point && F > 0 && dim >= 1,
               "bad arguments to pomp_shard_gather_deferred");
  LAUNCH(gather_deferred_kernel, (int)((F + 255) / 256), 256, 0, (cudaStream_t)stream, q, dim, defer_list,
         (const unsigned long long*)defer_count, (long long)F, pad_point, qf);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_scatter_deferred(const int64_t* merged_idx, const double* merged_dist, int32_t k,
                                         const int64_t* defer_list, const uint64_t* defer_count, int64_t F,
                                         int64_t* idx, double* dist, void* stream) {
  POMP_REQUIRE(merged_idx && merged_dist && defer_list && defer_count && idx && dist && F > 0 && k >= 1,
               "bad arguments to pomp_shard_scatter_deferred");
  LAUNCH(scatter_deferred_kernel, (int)((F + 255) / 256), 256, 0, (cudaStream_t)stream, merged_idx, merged_dist, k,
         defer_list, (const unsigned long long*)defer_count, (long long)F, idx, dist);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_route(const double* q, int64_t nq, int32_t dim, int32_t axis, const double* bounds_h,
                              int32_t n_ranks, int32_t* dest_scratch, uint64_t* counts_dev, int64_t* counts_h,
                              double* q_out, int64_t* perm_out, void* stream) {
  POMP_REQUIRE(n_ranks >= 1 && n_ranks <= SHARD_MAX_RANKS && dim >= 1 && dim <= POMP_MAX_DIM && axis >= 0 && axis < dim,
               "bad arguments to pomp_shard_route");
  POMP_REQUIRE(nq == 0 || (q && dest_scratch && counts_dev && counts_h && q_out && perm_out), "NULL argument to pomp_shard_route");
  cudaStream_t st = (cudaStream_t)stream;
  Bounds B;
  for (int r = 0; r < SHARD_MAX_RANKS; r++) B.b[r] = (r < n_ranks - 1 && bounds_h) ? bounds_h[r] : INFINITY;
  CUDA_TRY(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * 2 * SHARD_MAX_RANKS, st));
  const int blocks = (int)((nq + 255) / 256);
  if (nq > 0) LAUNCH(route_count_kernel, blocks, 256, 0, st, q, (long long)nq, dim, axis, B, n_ranks, dest_scratch, (unsigned long long*)counts_dev);
  uint64_t c[SHARD_MAX_RANKS];
  CUDA_TRY(cudaMemcpyAsync(c, counts_dev, sizeof(uint64_t) * n_ranks, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // the split sizes of the all-to-all are needed on the host anyway
  uint64_t cur[SHARD_MAX_RANKS], acc = 0;
  for (int r = 0; r < n_ranks; r++) {
    counts_h[r] = (int64_t)c[r];
    cur[r] = acc;
    acc += c[r];
  }
  CUDA_TRY(cudaMemcpyAsync(counts_dev + SHARD_MAX_RANKS, cur, sizeof(uint64_t) * n_ranks, cudaMemcpyHostToDevice, st));
  if (nq > 0) LAUNCH(route_scatter_kernel, blocks, 256, 0, st, q, (long long)nq, dim, dest_scratch, n_ranks,
                     (unsigned long long*)(counts_dev + SHARD_MAX_RANKS), q_out, perm_out);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_finalize(const double* q, int64_t nq, int32_t dim, int32_t axis, double lo_cov, double hi_cov,
                                 int32_t k, const int64_t* idx_local, const double* dist, const int64_t* local_to_global,
                                 const uint8_t* halo_flag, int32_t drop_halo, int64_t* records, uint8_t* uncertified,
                                 void* stream) {
  POMP_REQUIRE(nq >= 0 && k >= 1 && dim >= 1 && axis >= 0 && axis < dim, "bad arguments to pomp_shard_finalize");
  if (nq == 0) return POMP_OK;
  POMP_REQUIRE(q && idx_local && dist && records, "NULL argument to pomp_shard_finalize");
  LAUNCH(finalize_kernel, (int)((nq + 255) / 256), 256, 0, (cudaStream_t)stream, q, (long long)nq, dim, axis, lo_cov,
         hi_cov, k, idx_local, dist, local_to_global, halo_flag, drop_halo, records, uncertified);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_shard_unpack(const int64_t* records, int64_t nq, int32_t k, const int64_t* perm, int64_t* idx,
                               double* dist, int32_t* count, void* stream) {
  POMP_REQUIRE(nq >= 0 && k >= 1, "bad arguments to pomp_shard_unpack");
  if (nq == 0) return POMP_OK;
  POMP_REQUIRE(records && idx && dist, "NULL argument to pomp_shard_unpack");
  LAUNCH(unpack_kernel, (int)((nq + 255) / 256), 256, 0, (cudaStream_t)stream, records, (long long)nq, k, perm, idx,
         dist, count);
  CHECK_LAUNCH();
  return POMP_OK;
}
