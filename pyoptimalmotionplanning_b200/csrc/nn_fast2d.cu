// nn_fast2d.cu -- launchers of the 2-D Euclidean large-batch k-NN kernels (the headline path):
// the flattened fixed-block kernel (grid_fast.cuh) and the streaming TMA tile kernel
// (grid_tile.cuh), each followed by the overflow pass through the pruned traversal.
#include "common.cuh"
#include "grid.cuh"
#include "grid_fast.cuh"
#include "grid_pipe.cuh"
#include "grid_tile.cuh"
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
  // small block first, R+1 (list mode) for the few that fail.  k >= 8 always; k < 8 when the
  // two_stage knob is set (otherwise those retry inside the thread)
  const bool small_two = KK < 8 && h->two_stage != 0 && R < 3;
  const bool two_stage = (KK >= 8 && R < 3) || small_two;
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
#define FLATNR(KC, RC, NB, QS, QP, OL, OC, QL, QC)                                                                 \
  LAUNCH((grid_knn_flat2d_kernel<(KC < 8 ? KC : 1), RC, false>), NB, 128, 0, st, h->grid, h->d_pts,                \
         (long long)h->n_dev, skip, q, QS, (long long)nq, QP, k, idx, dist, cnt, OL, OC, QL, QC)
#define GKR(KC, RC)                                                                                                \
  do {                                                                                                             \
    if (small_two) {                                                                                               \
      FLATNR(KC, RC, blocks, qsorted, qperm, l1, c1, np, np);                                                      \
      prof_end(h, st);                                                                                             \
      FLATNR(KC, (RC < 3 ? RC + 1 : 3), (blocks + 63) / 64, npd, np, l2, c2, (const int*)l1, (const int*)c1);      \
      PRUNED(KC, l2, c2);                                                                                          \
      break;                                                                                                       \
    }                                                                                                              \
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
  if (KK == 1 && k == 1 && R == 1 && h->pipe_search != 0 && qsorted && qperm && !skip &&
      h->n_dev == (int64_t)h->grid.n_indexed) {
    // software-pipelined nearest-neighbour kernel; its overflow (block fuller than CAP, or no
    // certificate) gets a second chance in the flat kernel (list mode), then the pruned traversal
    constexpr int CAP = 24, TH = 128;
    size_t psm = (size_t)2 * CAP * TH * sizeof(double2);
    auto kern = grid_nn_pipe2d_kernel<CAP, TH>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
    CUDA_TRY(cudaMemsetAsync(c2, 0, sizeof(int), st));
    int pblocks = std::min(blocks, sm_count(h->device) * 2);
    LAUNCH(kern, pblocks, TH, psm, st, h->grid, qsorted, qperm, (long long)nq, idx, dist, cnt, l1, c1);
    prof_end(h, st);
    FLAT(1, 1, (blocks + 7) / 8, npd, np, l2, c2, (const int*)l1, (const int*)c1);
    PRUNED(1, l2, c2);
  } else if (KK == 1 && R == 1 && !small_two && h->flat_u != 4) {
    // experiment knob: candidate loads in flight per thread (registers vs round trips)
    if (h->flat_u == 2)
      LAUNCH((grid_knn_flat2d_kernel<1, 1, true, 2>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,
             qsorted, (long long)nq, qperm, k, idx, dist, cnt, l1, c1, np, np);
    else if (h->flat_u == 6)
      LAUNCH((grid_knn_flat2d_kernel<1, 1, true, 6>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,
             qsorted, (long long)nq, qperm, k, idx, dist, cnt, l1, c1, np, np);
    else
      LAUNCH((grid_knn_flat2d_kernel<1, 1, true, 8>), blocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,
             qsorted, (long long)nq, qperm, k, idx, dist, cnt, l1, c1, np, np);
    prof_end(h, st);
    PRUNED(1, l1, c1);
  } else if (KK == 1) GK(1);
  else if (KK == 4) GK(4);
  else if (KK == 8) GK(8);
  else if (KK == 16) GK(16);
  else GK(32);
#undef GK
#undef GKR
#undef FLATNR
#undef PRUNED
#undef FLAT
  CHECK_LAUNCH();
  return POMP_OK;
}

// 2-D Euclidean, bucket-sorted queries: streaming tile kernel (grid_tile.cuh) + overflow pass
template <int TW>
static int launch_grid_knn_tile2d(pomp_nn* h, const double* q, int64_t nq, const int* qperm, const double* qsorted,
                                  const int* qstart, int nbx, int k, int R, int64_t* idx, double* dist, int32_t* cnt,
                                  cudaStream_t st) {
  constexpr int TH = 8;
  const int KK = k == 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32)));
  double ppc = (double)h->grid.n_indexed / (double)h->grid.n_cells;
  int cap = (int)(1.5 * (TH + 2 * R) * (TW + 2 * R) * ppc) + 128;
  size_t smem = (size_t)cap * 16 + sizeof(TileShared<TW, TH>);
  if (smem > 200 * 1024) return 1;  // caller falls back
  POMP_TRY(h->ss[h->cur].ovf.reserve((size_t)nq + 1));
  int* ovf_count = h->ss[h->cur].ovf.p;
  int* ovf_list = h->ss[h->cur].ovf.p + 1;
  if (!h->ss[h->cur].ovf_zeroed) CUDA_TRY(cudaMemsetAsync(ovf_count, 0, sizeof(int), st));
  const int W = h->grid.ncell[0], H = h->grid.ncell[1];
  int tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH);
  int oblocks = std::min((int)((nq + 127) / 128), sm_count(h->device) * 4);
  const uint8_t* skip = h->skip_ptr();
#define GK(KC)                                                                                                       \
  do {                                                                                                               \
    auto kern = grid_knn_tile2d_kernel<KC, TW, TH>;                                                                  \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    LAUNCH(kern, tiles, 128, smem, st, h->grid, h->d_pts, (long long)h->n_dev, skip, qsorted, qperm, qstart, nbx, k, R, \
           cap, idx, dist, cnt, ovf_list, ovf_count);                                                                \
    prof_end(h, st);                                                                                                 \
    LAUNCH((grid_knn_fast_kernel<2, KC>), oblocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,     \
           (long long)nq, (const int*)nullptr, k, idx, dist, cnt, (const int*)ovf_list, (const int*)ovf_count);      \
  } while (0)
  prof_begin(h, st);
  if (KK == 1) GK(1);
  else if (KK == 4) GK(4);
  else if (KK == 8) GK(8);
  else if (KK == 16) GK(16);
  else GK(32);
#undef GK
  CHECK_LAUNCH();
  return POMP_OK;
}


int pomp::launch_grid_knn_tile2d_64(pomp_nn* h, const double* q, int64_t nq, const int* qperm, const double* qsorted,
                                    const int* qstart, int nbx, int k, int R, int64_t* idx, double* dist,
                                    int32_t* cnt, cudaStream_t st) {
  return launch_grid_knn_tile2d<64>(h, q, nq, qperm, qsorted, qstart, nbx, k, R, idx, dist, cnt, st);
}

// persistent TMA-pipelined kernel (grid_knn_stream2d_kernel v2): needs the TILE-major query sort
// (sort_shift 6, row_div ST_TH: nn.cu sets both when stream_search is on)
int pomp::launch_grid_knn_stream2d(pomp_nn* h, const double* q, int64_t nq, const int* qperm, const double* qsorted,
                                   const int* qstart, int nbx, int k, int R, int64_t* idx, double* dist,
                                   int32_t* cnt, cudaStream_t st) {
  (void)nbx;
  const int KK = k == 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : (k <= 16 ? 16 : 32)));
  if (KK > 4) return 1;  // register lists only; larger k stays on the flat kernel
  double ppc = (double)h->grid.n_indexed / (double)h->grid.n_cells;
  double qdens = (double)nq / (double)h->grid.n_cells;
  int cap_pts = ((int)(1.2 * (ST_TH + 2 * R) * (ST_TW + 2 * R) * ppc) + 96 + 3) / 4 * 4;
  int cap_q = ((int)(1.5 * ST_TH * ST_TW * qdens) + 64 + 3) / 4 * 4;
  StreamLayout L = stream_layout(cap_pts, cap_q);
  size_t smem = 256 + (size_t)ST_STAGES * L.stage_bytes;
  if (smem > 225 * 1024 || h->n_dev >= (1LL << 31) || R > TILE_RMAX || cap_q > 32 * ST_SLOTS)
    return 1;  // caller falls back
  SortScratch& S = h->ss[h->cur];
  POMP_TRY(S.ovf.reserve(2 * (size_t)nq + 2));
  int* c1 = S.ovf.p;
  int* l1 = S.ovf.p + 1;
  int* c2 = S.ovf.p + nq + 1;
  int* l2 = S.ovf.p + nq + 2;
  if (!S.ovf_zeroed) CUDA_TRY(cudaMemsetAsync(c1, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(c2, 0, sizeof(int), st));
  int ctas = sm_count(h->device);
  int blocks = (int)((nq + 127) / 128);
  int oblocks = std::min(blocks, sm_count(h->device) * 4);
  const uint8_t* skip = h->skip_ptr();
  const int* np = nullptr;
  const double* npd = nullptr;
#define GK(KC)                                                                                                       \
  do {                                                                                                               \
    auto kern = grid_knn_stream2d_kernel<KC>;                                                                        \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    LAUNCH(kern, ctas, ST_THREADS, smem, st, h->grid, h->d_pts, (long long)h->n_dev, skip, qsorted, qperm, qstart,   \
           k, R, cap_pts, cap_q, idx, dist, cnt, l1, c1);                                                            \
    prof_end(h, st);                                                                                                 \
    /* second chance: flat kernel with the next larger block over the overflow list, then pruned */                \
    if (R == 1)                                                                                                      \
      LAUNCH((grid_knn_flat2d_kernel<KC, 2>), (blocks + 7) / 8, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev,  \
             skip, q, npd, (long long)nq, np, k, idx, dist, cnt, l2, c2, (const int*)l1, (const int*)c1);            \
    else                                                                                                             \
      LAUNCH((grid_knn_flat2d_kernel<KC, 3>), (blocks + 7) / 8, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev,  \
             skip, q, npd, (long long)nq, np, k, idx, dist, cnt, l2, c2, (const int*)l1, (const int*)c1);            \
    LAUNCH((grid_knn_fast_kernel<2, KC>), oblocks, 128, 0, st, h->grid, h->d_pts, (long long)h->n_dev, skip, q,     \
           (long long)nq, np, k, idx, dist, cnt, (const int*)l2, (const int*)c2);                                    \
  } while (0)
  prof_begin(h, st);
  if (KK == 1) GK(1);
  else GK(4);
#undef GK
  CHECK_LAUNCH();
  return POMP_OK;
}
