// nn_handle.cuh -- the NearestNeighbors handle (shared by nn.cu and the fused extend kernel).
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "grid.cuh"

using namespace pomp;

// scratch of one bucket-sorted query batch (two sets: sort of chunk c+1 overlaps search of chunk c)
struct SortScratch {
  pomp::DevBuf<int> counts, qcell, qperm, ovf;  // ovf[0] = overflow count, ovf[1..] = overflow query ids
  pomp::DevBuf<double> qsorted;
  pomp::DevBuf<long long> scan_tmp;
  bool ovf_zeroed = false;
  // strip path (strip2d.cu): bucket counters that the kernels hand back ZEROED (the producer warp clears a tile's
  // counters after reading them, the overflow pass clears the deferred count), so a step needs no memset launch
  pomp::DevBuf<int> strip_ctrl;
  bool strip_ctrl_clean = false;
  size_t strip_ctrl_n = 0;
};

// ============================================================================ handle
struct pomp_nn {
  int device = 0;
  pomp_metric_desc metric;
  int D = 0;
  // canonical store, insertion order
  double* d_pts = nullptr;
  uint8_t* d_skip = nullptr;  // dead | rejected, zero-filled up to cap
  int64_t cap = 0;
  int64_t n_dev = 0;  // points resident on the device
  std::vector<double> pending;  // host-side adds not yet uploaded
  std::vector<uint8_t> h_dead;
  std::vector<uint8_t> h_reject;
  bool any_dead = false, has_reject = false;
  // grid index
  bool grid_valid = false;
  bool grid_unusable = false;  // metric gives no grid axis
  GridDev grid;
  DevBuf<int> cell_start;
  DevBuf<double> spts;
  DevBuf<int> sidx;
  DevBuf<int> tmp_i32;  // slot per point / counts
  DevBuf<long long> scan_tmp;
  // strip index (strip2d.cuh): blocked copy of the 2-D grid index, built on demand for large nearest() batches
  uint64_t strip_build = 0;  // grid_builds value it was derived from (0 = none)
  uint64_t strip_launches = 0;
  const int* ovf_count_dev = nullptr;  // device counter of the queries the last batch deferred to the overflow passes
  StripDev strip;
  DevBuf<double> spts2;
  DevBuf<int> sidx2, tile_info, srow;
  DevBuf<uint16_t> tile_cs;
  double strip_search = 1;         // 2-D Euclidean k = 1 batches: streaming kernel over the strip index
  double strip_spin_ns = 0, strip_probe = 0;  // tuning: back-off of waiting consumer warps, non-blocking producer probe
  double strip_min_per_tile = 24;  // ... when the batch brings at least this many queries per tile on average
  // knobs
  double grid_min_points = 16384;
  double points_per_cell = 0;  // 0 = auto by number of grid axes
  double sort_queries = 1;
  double sort_min_queries = 32768;
  double sort_shift = 6;
  double grid_min_queries = 64;
  double max_tail = 2048;
  double periodic_axes = 1;  // SO(2) components of a product metric are grid axes (wrap-around cells)
  double bf_tiled = 1;      // brute-force regime, Euclidean k = 1 batches: shared-memory tiled kernel
  double se2_fast = 1;      // SE(2) product metric: compile-time kernel (grid_se2.cuh) instead of the generic one
  double block_need = 0;    // k = 1: expected points in the certified ball that selects the block radius (0 = default)
  double warp_search = 1;   // Euclidean k-NN: warp-per-query shell kernel (grid_warp.cuh); 0 off, 1 where it wins, 2 always
  double warp_lanes = 0;    // lanes per query of that kernel (0 = by k: 8 / 16 / 32 lanes hold k <= 8 / 16 / 32 keys)
  double warp_need = 0;     // expected points inside its first shell (0 = k + 1.5 sqrt(k) + 1; 3 for k = 1)
  double block_search = 1;  // 2-D Euclidean: fixed block of cells + certificate (0: pruned traversal; 2: also in 3-D)
  // scratch for queries
  DevBuf<double> q_dev;
  DevBuf<double> part_d;
  DevBuf<int> part_i;
  DevBuf<int> counts;
  DevBuf<int> qperm;
  DevBuf<int> qcell;
  DevBuf<double> qsorted;
  SortScratch ss[2];
  int cur = 0;  // scratch set of the chunk being searched
  cudaStream_t aux_stream = nullptr;  // runs the query sort of the next chunk
  std::vector<cudaEvent_t> ovl_ev;
  double overlap_chunks = 1;          // chunks per large device batch; >1 overlaps sort and search on two
                                      // streams -- measured SLOWER on B200 (host launch rate bound), kept as a knob
  double overlap_min_queries = 524288;
  DevBuf<int64_t> out_idx;
  DevBuf<double> out_dist;
  DevBuf<int32_t> out_cnt;
  DevBuf<int64_t> out_off;
  DevBuf<int64_t> rad_off;  // CSR offsets of pomp_nn_neighbors_h
  DevBuf<int> rstage;       // radius queries: parked hit positions, RADIUS_STAGE_CAP per query
  DevBuf<double> rw_ab;     // pomp_rewire_candidates_h: the 2k edges (a then b)
  DevBuf<uint8_t> rw_ok;    // packed result record of pomp_rewire_candidates_h
  void* rw_pin = nullptr;   // its pinned host landing buffer
  size_t rw_pin_cap = 0;
  // pinned mailbox for the single-query fast path
  void* mailbox_h = nullptr;
  void* mailbox_d = nullptr;
  unsigned int* ticket = nullptr;
  // single radius query fast path: hit list in mapped pinned memory, device counters {hits, ticket}
  void* radbox_h = nullptr;
  void* radbox_d = nullptr;
  unsigned int* rad_counters = nullptr;
  double single_radius = 1;   // 0: every neighbors() call takes the general (count / scan / fill) path
  cudaStream_t stream = nullptr;  // used by the *_h variants
  cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H streams of the pipelined *_h path
  std::vector<cudaEvent_t> pipe_ev;
  int pipe_chunks_active = 0;   // > 1 while a chunk of the pipelined *_h path is being searched
  double pipeline_chunks = 6;   // measured on B200 (2^20 2-D queries): 4: 0.575, 6: 0.552, 8: 0.567 ms end to end
  double pipeline_shape = 1;    // chunk sizes: 0 = shrinking (C+1, C, ..., 2), 1 = triangle
  // CUDA graphs of the per-chunk searches of the pipelined *_h path: with both copy engines busy every
  // separate kernel launch costs ~8 us extra (profiles/chunk_probe.py), and a chunk is 8 small launches
  struct ChunkGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t sig = 0;        // launch signature the graph was captured under
    uint64_t seen_sig = 0;   // signature of the previous direct run with this shape
    const double* q = nullptr;
    int64_t m = 0;
    int k = 0;
    void *idx = nullptr, *dist = nullptr, *cnt = nullptr;
    int64_t launches = 0;
  };
  std::vector<ChunkGraph> chunk_graphs;
  double use_graphs = 1;
  uint64_t grid_builds = 0, mut_count = 0;
  // optional timing of the dominant search kernel with CUDA events on the launching stream
  // (bench.py's roofline numerator); enabled with set_param("profile", 1)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs
  size_t prof_used = 0;

  int64_t size() const { return n_dev + (int64_t)(pending.size() / (size_t)std::max(D, 1)); }
  const uint8_t* skip_ptr() const { return (any_dead || has_reject) ? d_skip : nullptr; }
};


namespace pomp {
// ---- optional event timing of the dominant kernel ---------------------------------------------
inline void prof_begin(pomp_nn* h, cudaStream_t st) {
  if (!h->profile) return;
  if (h->prof_used + 2 > h->prof_ev.size()) {
    for (int i = 0; i < 2; i++) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      h->prof_ev.push_back(e);
    }
  }
  cudaEventRecord(h->prof_ev[h->prof_used], st);
}
inline void prof_end(pomp_nn* h, cudaStream_t st) {
  if (!h->profile || h->prof_used + 2 > h->prof_ev.size()) return;
  cudaEventRecord(h->prof_ev[h->prof_used + 1], st);
  h->prof_used += 2;
}


int nn_grow(pomp_nn* h, int64_t need, cudaStream_t st);
int nn_flush(pomp_nn* h, cudaStream_t st);

struct QueryArg {
  double v[POMP_MAX_DIM];
};
struct Mailbox {
  int64_t idx;
  double dist;
  int64_t aux[6];
  double x[POMP_MAX_DIM];
  double len;
};
}  // namespace pomp

namespace pomp {
// nn_fast2d.cu: 2-D Euclidean gather kernels (flattened block kernel + overflow passes)
int launch_grid_knn_flat2d(pomp_nn* h, const double* q, const double* qsorted, int64_t nq, const int* qperm, int k,
                           int R, int64_t* idx, double* dist, int32_t* cnt, cudaStream_t st);
// overflow lists of a first-stage 2-D search: flat kernel with the next larger block, then the pruned traversal
int launch_overflow_2d(pomp_nn* h, const double* q, int64_t nq, int k, int* l1, int* c1, int* l2, int* c2, int64_t* idx,
                       double* dist, int32_t* cnt, cudaStream_t st);
// nn_warp.cu: warp-per-query shell kernel (grid_warp.cuh)
bool warp_knn_eligible(const pomp_nn* h, int k);
int launch_grid_knn_warp(pomp_nn* h, const double* q, const double* qsorted, int64_t nq, const int* qperm, int k, int64_t* idx, double* dist,
                         int32_t* cnt, cudaStream_t st);
// strip2d.cu: streaming nearest-neighbour search over the strip index
bool strip_eligible(const pomp_nn* h, int64_t nq, int k);
int strip_nearest(pomp_nn* h, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* cnt,
                  cudaStream_t st);
}  // namespace pomp
