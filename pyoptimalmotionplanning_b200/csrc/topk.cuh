// topk.cuh -- exact top-k selection with (distance, index) ordering.
//
// KNearestResult (pomp/structures/knn.py:7-30) keeps the k smallest distances, replacing the
// current maximum only on strict '<', and returns them sorted by (distance, item).  Scanning
// points in insertion order that is exactly "the k smallest keys (distance, insertion index)".
// Candidates at +inf (CostMetric above costMax) or NaN are never admitted (d < inf fails).
#pragma once
#include "common.cuh"

namespace pomp {

#define FULL_MASK 0xffffffffu
#define IDX_NONE 0x7fffffff
template <typename T>
__host__ __device__ constexpr T idx_none() {
  return sizeof(T) == 8 ? (T)0x7fffffffffffffffLL : (T)0x7fffffff;
}

// ---- warp-distributed sorted list: position p lives in lane p%32, slot p/32 -----------------
template <int KL, typename IdxT>
struct WarpTopK {
  double d[KL];
  IdxT i[KL];
  double kth_d;  // key at position k-1 (broadcast)
  IdxT kth_i;
  int k;

  __device__ __forceinline__ void init(int k_) {
    k = k_;
#pragma unroll
    for (int s = 0; s < KL; s++) {
      d[s] = POMP_INF;
      i[s] = idx_none<IdxT>();
    }
    kth_d = POMP_INF;
    kth_i = idx_none<IdxT>();
  }

  static __device__ __forceinline__ bool less(double d1, IdxT i1, double d2, IdxT i2) {
    return d1 < d2 || (d1 == d2 && i1 < i2);
  }

  __device__ __forceinline__ void refresh_kth() {
    int p = k - 1;
    double vd = POMP_INF;
    IdxT vi = idx_none<IdxT>();
#pragma unroll
    for (int s = 0; s < KL; s++)
      if (p / 32 == s) {
        vd = d[s];
        vi = i[s];
      }
    kth_d = __shfl_sync(FULL_MASK, vd, p % 32);
    kth_i = __shfl_sync(FULL_MASK, vi, p % 32);
  }

  // all lanes call with the same (cd, ci)
  __device__ __forceinline__ void insert_uniform(double cd, IdxT ci) {
    int lane = threadIdx.x & 31;
    int P = 0;
#pragma unroll
    for (int s = 0; s < KL; s++) P += __popc(__ballot_sync(FULL_MASK, less(d[s], i[s], cd, ci)));
#pragma unroll
    for (int s = KL - 1; s >= 0; s--) {
      double up_d = __shfl_up_sync(FULL_MASK, d[s], 1);
      IdxT up_i = __shfl_up_sync(FULL_MASK, i[s], 1);
      if (s > 0) {
        double pd = __shfl_sync(FULL_MASK, d[s - 1], 31);
        IdxT pi = __shfl_sync(FULL_MASK, i[s - 1], 31);
        if (lane == 0) {
          up_d = pd;
          up_i = pi;
        }
      }
      int p = s * 32 + lane;
      if (p > P) {
        d[s] = up_d;
        i[s] = up_i;
      } else if (p == P) {
        d[s] = cd;
        i[s] = ci;
      }
    }
    refresh_kth();
  }

  // each lane offers one candidate (cd = +inf for "none")
  __device__ __forceinline__ void offer(double cd, IdxT ci) {
    int lane = threadIdx.x & 31;
    bool pass = (cd < POMP_INF) && less(cd, ci, kth_d, kth_i);
    unsigned mask = __ballot_sync(FULL_MASK, pass);
    while (mask) {
      int src = __ffs(mask) - 1;
      double bd = __shfl_sync(FULL_MASK, cd, src);
      IdxT bi = __shfl_sync(FULL_MASK, ci, src);
      insert_uniform(bd, bi);
      if (lane == src) pass = false;
      pass = pass && less(cd, ci, kth_d, kth_i);
      mask = __ballot_sync(FULL_MASK, pass);
    }
  }

  // write positions [0,k) ; returns through lane-uniform count of finite entries
  template <typename OutI>
  __device__ __forceinline__ int store(double* od, OutI* oi) const {
    int lane = threadIdx.x & 31;
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < KL; s++) {
      int p = s * 32 + lane;
      bool fin = d[s] < POMP_INF;
      if (p < k) {
        od[p] = d[s];
        oi[p] = fin ? (OutI)i[s] : (OutI)-1;
      }
      cnt += __popc(__ballot_sync(FULL_MASK, fin && p < k));
    }
    return cnt;
  }
};

// ---- per-thread sorted list (generic grid search: one query per thread) ----------------------
// d[0..K) ascending by (distance, index).  Always keeps K entries; a runtime k <= K is served by
// the first k (the k best are a prefix of the K best).  K <= 32: unrolled, register resident;
// larger K: rolled loops over a local-memory list.
template <int K>
struct ThreadTopK {
  double d[K];
  int i[K];

  __device__ __forceinline__ void init(int) {
    if (K <= 32) {
#pragma unroll
      for (int j = 0; j < K; j++) {
        d[j] = POMP_INF;
        i[j] = IDX_NONE;
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < K; j++) {
        d[j] = POMP_INF;
        i[j] = IDX_NONE;
      }
    }
  }
  __device__ __forceinline__ double threshold() const { return d[K - 1]; }
  __device__ __forceinline__ bool full() const { return d[K - 1] < POMP_INF; }
  // cheap pre-test before the index is loaded
  __device__ __forceinline__ bool maybe(double cd) const { return cd < POMP_INF && cd <= d[K - 1]; }
  __device__ __forceinline__ void insert(double cd, int ci) {
    if (!(cd < POMP_INF) || !key_less(cd, ci, d[K - 1], i[K - 1])) return;
    if (K <= 32) {
#pragma unroll
      for (int j = K - 1; j >= 1; j--) {
        bool lt_prev = key_less(cd, ci, d[j - 1], i[j - 1]);
        bool lt_cur = key_less(cd, ci, d[j], i[j]);
        if (lt_prev) {
          d[j] = d[j - 1];
          i[j] = i[j - 1];
        } else if (lt_cur) {
          d[j] = cd;
          i[j] = ci;
        }
      }
      if (key_less(cd, ci, d[0], i[0])) {
        d[0] = cd;
        i[0] = ci;
      }
    } else {
      int j = K - 1;
#pragma unroll 1
      while (j > 0 && key_less(cd, ci, d[j - 1], i[j - 1])) {
        d[j] = d[j - 1];
        i[j] = i[j - 1];
        j--;
      }
      d[j] = cd;
      i[j] = ci;
    }
  }
};

// in-place ascending sort of one query's hit list (radius queries return hits by insertion index)
__device__ __forceinline__ void heapsort_i64(int64_t* a, long long n) {
  if (n < 2) return;
  for (long long start = n / 2 - 1; start >= 0; start--) {
    long long root = start;
    while (true) {
      long long child = 2 * root + 1;
      if (child >= n) break;
      if (child + 1 < n && a[child] < a[child + 1]) child++;
      if (a[root] >= a[child]) break;
      int64_t t = a[root];
      a[root] = a[child];
      a[child] = t;
      root = child;
    }
  }
  for (long long end = n - 1; end > 0; end--) {
    int64_t t = a[0];
    a[0] = a[end];
    a[end] = t;
    long long root = 0;
    while (true) {
      long long child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && a[child] < a[child + 1]) child++;
      if (a[root] >= a[child]) break;
      int64_t t2 = a[root];
      a[root] = a[child];
      a[child] = t2;
      root = child;
    }
  }
}


}  // namespace pomp
