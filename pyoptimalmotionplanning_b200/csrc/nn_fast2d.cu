// nn_fast2d.cu -- launchers of the 2-D Euclidean gather kernels: the flattened fixed-block kernel
// (grid_fast.cuh) followed by the overflow passes (next larger block, then the pruned traversal).
#include "common.cuh"
#include "grid.cuh"
#include "grid_fast.cuh"
#include "nn_handle.cuh"
#include "topk.cuh"

using namespace pomp;

// 2-D: flattened fixed-block kernel (grid_knn_flat2d_kernel) + overflow pass
int pomp::launch_grid_knn_flat2d(pomp_nn* h, const double* q, const double* qsorted, int64_t nq, const int* qperm,
                                 int k, int R, int64_t* idx, double* dist, int32_t* cnt, cudaStream_t st) {
  const int KK = k == 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32)));
  SortScratch& S = h->ss[h->cur];
  // two overflow lists: [0] count1, [1..nq] list1, [nq+1] count2, [nq+2..] list2
  POMP_TRY(S.ovf.reserve(2 * (size_t)nq + 2));
  int* c1 = S.ovf.p;
  int* l1 = S.ovf.p + 1;
  int* c2 = S.ovf.p + nq + 1;
  int* l2 = S.ovf.p + nq + 2;
  if (!S.ovf_zeroed) CUDA_TRY(cudaMemsetAsync(c1, 0, sizeof(int), st));
  h->ovf_count_dev = c1;
  // k >= 8: small block first, R+1 (list mode) for the few that fail (k < 8 retries inside the thread / warp)
  const bool two_stage = KK >= 8 && R < 3;
  if (two_stage) CUDA_TRY(cudaMemsetAsync(c2, 0, sizeof(int), st));
  int blocks = (int)((nq + 127) / 128);
  int oblocks = std::min(blocks, sm_count(h->device) * 4);
  const uint8_t* skip = h->skip_ptr();
  const int* np = nullptr;
  const double* npd = nullptr;
#define FLAT(KC, RC, NB, QS, QP, OL, OC, QL, QC)                                                                   \
  do {                                                                                                             \
    size_t fsm = KC >= 8 ? (size_t)KC * 128 * 12 : 0; /* SmemTopK: K x 128 threads x (8 + 4) bytes */             \
    if (fsm > 48 * 1024)                                                                                           \
      CUDA_TRY(cudaFuncSetAttribute(grid_knn_flat2d_kernel<KC, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                    (int)fsm));                                                                    \
    LAUNCH((grid_knn_flat2d_kernel<KC, RC>), NB, 128, fsm, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q, QS, \
           (long long)nq, QP, k, idx, dist, cnt, OL, OC, QL, QC);                                                  \
  } while (0)
#define PRUNED(KC, OL, OC)                                                                                         \
  LAUNCH((grid_knn_fast_kernel<2, KC>), oblocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,     \
         (long long)nq, np, k, idx, dist, cnt, (const int*)OL, (const int*)OC)
#define GKR(KC, RC)                                                                                                \
  do {                                                                                                             \
    FLAT(KC, RC, blocks, qsorted, qperm, l1, c1, np, np);                                                          \
    prof_end(h, st);                                                                                               \
    if (two_stage) {                                                                                               \
      /* list length is unknown on the host: launch enough blocks for 1/8 of the batch, rest -> pruned */        \
      FLAT(KC, (RC < 3 ? RC + 1 : 3), (blocks + 7) / 8, npd, np, l2, c2, (const int*)l1, (const int*)c1);          \
      PRUNED(KC, l2, c2);                                                                                          \
    } else {                                                                                                       \
      PRUNED(KC, l1, c1);                                                                                          \
    }                                                                                                              \
  } while (0)
#define GK(KC)                   \
  do {                           \
    if (R == 1) GKR(KC, 1);      \
    else if (R == 2) GKR(KC, 2); \
    else GKR(KC, 3);             \
  } while (0)
  prof_begin(h, st);
  if (KK == 1) GK(1);
  else if (KK == 4) GK(4);
  else if (KK == 8) GK(8);
  else if (KK == 16) GK(16);
  else GK(32);
#undef GK
#undef GKR
#undef PRUNED
#undef FLAT
  CHECK_LAUNCH();
  return POMP_OK;
}

// Overflow lists of a first-stage search (l1/c1: uncertified queries): flat kernel with the 5x5 block in list mode,
// whatever that cannot certify either (l2/c2) through the pruned traversal.  c2 must be zero on entry.
int pomp::launch_overflow_2d(pomp_nn* h, const double* q, int64_t nq, int k, int* l1, int* c1, int* l2, int* c2,
                             int64_t* idx, double* dist, int32_t* cnt, cudaStream_t st) {
  const int KK = k == 1 ? 1 : 4;
  POMP_REQUIRE(k <= 4, "launch_overflow_2d serves k <= 4");
  const int blocks = (int)((nq + 127) / 128);
  const int oblocks = std::min(blocks, sm_count(h->device) * 4);
  const uint8_t* skip = h->skip_ptr();
  const int* np = nullptr;
  const double* npd = nullptr;
#define OV(KC)                                                                                                       \
  do {                                                                                                               \
    LAUNCH((grid_knn_flat2d_kernel<KC, 2>), std::max(1, (blocks + 7) / 8), 128, 0, st, h->grid, h->d_pts,            \
           (long long)h->n_dev, skip, q, npd, (long long)nq, np, k, idx, dist, cnt, l2, c2, (const int*)l1,          \
           (const int*)c1);                                                                                          \
    LAUNCH((grid_knn_fast_kernel<2, KC>), oblocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,     \
           (long long)nq, np, k, idx, dist, cnt, (const int*)l2, (const int*)c2);                                    \
  } while (0)
  if (KK == 1) OV(1);
  else OV(4);
#undef OV
  CHECK_LAUNCH();
  return POMP_OK;
}
