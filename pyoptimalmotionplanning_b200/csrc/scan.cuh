// scan.cuh -- exclusive prefix sum over int32 counts (bucket offsets of the grid build, CSR offsets
// of radius queries).  Three short kernels: tile sums, scan of the tile sums, apply.
#pragma once
#include "common.cuh"

namespace pomp {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long warp_incl_scan(long long v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns the prefix, *total = block sum
__device__ __forceinline__ long long block_excl_scan(long long v, long long* total) {
  __shared__ long long wsum[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  long long inc = warp_incl_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    long long x = lane < nw ? wsum[lane] : 0;
    long long xi = warp_incl_scan(x);
    wsum[lane] = xi - x;  // exclusive prefix of warp sums
    if (lane == 31) wsum[31] = xi - x;
  }
  __syncthreads();
  long long res = wsum[w] + inc - v;
  __shared__ long long tot;
  if (threadIdx.x == blockDim.x - 1) tot = res + v;
  __syncthreads();
  *total = tot;
  __syncthreads();
  return res;
}

static __global__ void scan_tile_sums(const int* __restrict__ in, long long n, long long* __restrict__ tile_sums) {
  long long base = (long long)blockIdx.x * SCAN_TILE;
  long long s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    long long i = base + (long long)j * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  long long tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of the tile sums
static __global__ void scan_sums_inplace(long long* sums, int nb) {
  long long carry = 0;
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    long long v = i < nb ? sums[i] : 0;
    long long tot;
    long long ex = block_excl_scan(v, &tot);
    if (i < nb) sums[i] = carry + ex;
    carry += tot;
  }
}

template <typename OutT>
__global__ void scan_apply(const int* __restrict__ in, long long n, const long long* __restrict__ tile_sums,
                           OutT* __restrict__ out) {
  long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    long long i = base + j;
    v[j] = i < n ? in[i] : 0;
    s += v[j];
  }
  long long tot;
  long long ex = block_excl_scan(s, &tot) + tile_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    long long i = base + j;
    if (i < n) out[i] = (OutT)ex;
    ex += v[j];
    if (i == n - 1) out[n] = (OutT)ex;  // grand total in the extra slot
  }
}

// out has n+1 entries; tile_sums scratch needs ceil(n/SCAN_TILE) long longs
template <typename OutT>
inline int exclusive_scan(const int* in, long long n, OutT* out, long long* tile_sums, cudaStream_t st) {
  if (n <= 0) {
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(OutT), st));
    return POMP_OK;
  }
  int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  LAUNCH(scan_tile_sums, nb, SCAN_THREADS, 0, st, in, n, tile_sums);
  LAUNCH(scan_sums_inplace, 1, 1024, 0, st, tile_sums, nb);
  LAUNCH((scan_apply<OutT>), nb, SCAN_THREADS, 0, st, in, n, tile_sums, out);
  CHECK_LAUNCH();
  return POMP_OK;
}

// single-CTA exclusive scan for small arrays (query bins), coalesced: 4 consecutive items per
// thread per pass, running total carried across passes.  out[n] = grand total
template <typename OutT>
__global__ void scan_small_kernel(const int* __restrict__ in, int n, OutT* __restrict__ out) {
  long long carry = 0;
  for (int base = 0; base < n; base += blockDim.x * 4) {
    int i0 = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = (i0 + j < n) ? in[i0 + j] : 0;
    long long s = (long long)v[0] + v[1] + v[2] + v[3];
    long long tot;
    long long ex = carry + block_excl_scan(s, &tot);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (i0 + j < n) out[i0 + j] = (OutT)ex;
      ex += v[j];
    }
    carry += tot;
  }
  if (threadIdx.x == 0) out[n] = (OutT)carry;
}
inline size_t scan_scratch_elems(long long n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1; }

}  // namespace pomp
