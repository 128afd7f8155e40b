// roadmap.cu -- device-resident roadmap construction over a NearestNeighbors node set (BASELINE config 3: "Kink 2D
// roadmap: 4M samples, batched edge visibility checks vs 10K box obstacles").
//
// The reference has no roadmap planner of its own for this -- its planners grow trees and export them with
// TreePlanner.getRoadmap (kinodynamicplanner.py:211-238) as (V, E) -- but every piece is a call on the path:
// ConfigurationSpace.feasible per sample (geometric.py:56-60), NearestNeighbors.knearest per sample
// (nearestneighbors.py:98-116), EpsilonEdgeChecker.feasible per candidate edge (edgechecker.py:20-30).  Here the
// whole chain runs on the device without a host round trip: V flags, the k-NN self-join, the edge list (optionally
// one direction of every mutual pair only), the sampled edge checks, and a STABLE compaction of the surviving
// edges (edge order = sample order, then neighbour rank -- what a host loop over the samples would append).
#include <algorithm>

#include "common.cuh"
#include "nn_handle.cuh"
#include "scan.cuh"
#include "space.cuh"

using namespace pomp;

namespace {

// sample i -> its k neighbours out of the k+1 nearest (the sample itself is dropped: the entry equal to i, or the
// last one when duplicates pushed i out of its own list).  undirected: (i, j) with j < i is dropped when i is in
// j's list, i.e. when (j, i) is emitted by sample j.
__global__ void __launch_bounds__(256)
roadmap_edges_kernel(const int64_t* __restrict__ kidx, long long n, int k1, int undirected, int* __restrict__ ij,
                     uint8_t* __restrict__ keep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t* row = kidx + i * k1;
  int self = k1 - 1;
  for (int t = 0; t < k1; t++)
    if (row[t] == i) {
      self = t;
      break;
    }
  int o = 0;
  for (int t = 0; t < k1; t++) {
    if (t == self) continue;
    const int64_t j = row[t];
    bool use = j >= 0;
    if (use && undirected && j < i) {
      const int64_t* rj = kidx + j * k1;
      for (int u = 0; u < k1; u++)
        if (rj[u] == i) use = false;
    }
    const long long e = i * (k1 - 1) + o;
    ij[2 * e] = use ? (int)i : 0;
    ij[2 * e + 1] = use ? (int)j : 0;
    keep[e] = use ? 1 : 0;
    o++;
  }
}

__global__ void __launch_bounds__(256)
roadmap_flags_kernel(const uint8_t* __restrict__ keep, const uint8_t* __restrict__ ok, long long ne,
                     int* __restrict__ flag) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < ne) flag[e] = (keep[e] && ok[e]) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
roadmap_scatter_kernel(const int* __restrict__ ij, const int* __restrict__ flag, const int* __restrict__ pos,
                       long long ne, long long cap, int* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || !flag[e]) return;
  const long long p = pos[e];
  if (p < cap) {
    out[2 * p] = ij[2 * e];
    out[2 * p + 1] = ij[2 * e + 1];
  }
}

}  // namespace

// Builds the k-nearest-neighbour roadmap of the node set of `nn` in the space `s`.
//   node_ok[n]          (device, may be NULL)  ConfigurationSpace.feasible of every node
//   edges_out[cap][2]   (device)               the feasible edges, stable order
//   *n_edges_h          (host)                 their number (may exceed cap: the caller re-runs with more room)
//   *n_candidates_h     (host, may be NULL)    candidate edges that were checked
POMP_API int pomp_roadmap_build(pomp_nn* nn, pomp_space* s, int32_t k, double resolution, int32_t undirected,
                                uint8_t* node_ok, int32_t* edges_out, int64_t cap, int64_t* n_edges_h,
                                int64_t* n_candidates_h, void* stream) {
  POMP_REQUIRE(nn && s && n_edges_h && (cap == 0 || edges_out), "NULL argument to pomp_roadmap_build");
  POMP_REQUIRE(k >= 1 && k < 128 && resolution > 0.0 && cap >= 0, "pomp_roadmap_build: k in 1..127, resolution > 0");
  POMP_REQUIRE(nn->D == s->dev.dim && nn->device == s->device, "pomp_roadmap_build: node set and space differ in dimension or device");
  DeviceGuard g(nn->device);
  if (!g.ok) return POMP_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  POMP_TRY(nn_flush(nn, st));
  const int64_t n = nn->n_dev;
  *n_edges_h = 0;
  if (n_candidates_h) *n_candidates_h = 0;
  if (n == 0) return POMP_OK;
  const int k1 = k + 1;
  const int64_t ne = n * k;
  POMP_REQUIRE(ne < 0x7fffffffLL && n < 0x7fffffffLL, "pomp_roadmap_build: more than 2^31 candidate edges");
  POMP_TRY(s->rm_kidx.reserve((size_t)n * k1));
  POMP_TRY(s->rm_kdist.reserve((size_t)n * k1));
  POMP_TRY(s->rm_ij.reserve((size_t)ne * 2));
  POMP_TRY(s->rm_keep.reserve((size_t)ne));
  POMP_TRY(s->rm_ok.reserve((size_t)ne));
  POMP_TRY(s->rm_flag.reserve((size_t)ne + 1));
  POMP_TRY(s->rm_pos.reserve((size_t)ne + 1));
  POMP_TRY(s->rm_scan.reserve(scan_scratch_elems(ne)));
  POMP_TRY(s->rm_node_ok.reserve((size_t)n));
  uint8_t* nok = node_ok ? node_ok : s->rm_node_ok.p;
  POMP_TRY(pomp_nn_knearest(nn, nn->d_pts, n, k1, s->rm_kidx.p, s->rm_kdist.p, nullptr, stream));
  LAUNCH(roadmap_edges_kernel, (int)((n + 255) / 256), 256, 0, st, s->rm_kidx.p, (long long)n, k1, undirected,
         s->rm_ij.p, s->rm_keep.p);
  // node feasibility once per node; the edges gather the flags and test their interior samples only
  POMP_TRY(pomp_edge_check_indexed_nodes(s, nn->d_pts, n, s->rm_ij.p, ne, resolution, nok, s->rm_ok.p, stream));
  const int eb = (int)((ne + 255) / 256);
  LAUNCH(roadmap_flags_kernel, eb, 256, 0, st, s->rm_keep.p, s->rm_ok.p, (long long)ne, s->rm_flag.p);
  POMP_TRY(exclusive_scan<int>(s->rm_flag.p, ne, s->rm_pos.p, s->rm_scan.p, st));
  LAUNCH(roadmap_scatter_kernel, eb, 256, 0, st, s->rm_ij.p, s->rm_flag.p, s->rm_pos.p, (long long)ne, (long long)cap,
         edges_out);
  CHECK_LAUNCH();
  int total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, s->rm_pos.p + ne, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  POMP_TRY(space_overflow_check(s));
  *n_edges_h = total;
  if (n_candidates_h) *n_candidates_h = ne;
  return POMP_OK;
}

// same with host outputs (node_ok_h[n] may be NULL; edges_h[cap][2])
POMP_API int pomp_roadmap_build_h(pomp_nn* nn, pomp_space* s, int32_t k, double resolution, int32_t undirected,
                                  uint8_t* node_ok_h, int32_t* edges_h, int64_t cap, int64_t* n_edges_h,
                                  int64_t* n_candidates_h) {
  POMP_REQUIRE(nn && s && n_edges_h && (cap == 0 || edges_h) && cap >= 0, "bad arguments to pomp_roadmap_build_h");
  DeviceGuard g(nn->device);
  if (!g.ok) return POMP_ERR_CUDA;
  POMP_TRY(nn_flush(nn, s->stream));
  const int64_t n = nn->n_dev;
  POMP_TRY(s->rm_out.reserve((size_t)std::max<int64_t>(cap, 1) * 2));
  if (node_ok_h) POMP_TRY(s->rm_node_ok.reserve((size_t)std::max<int64_t>(n, 1)));
  POMP_TRY(pomp_roadmap_build(nn, s, k, resolution, undirected, node_ok_h ? s->rm_node_ok.p : nullptr, s->rm_out.p,
                              cap, n_edges_h, n_candidates_h, s->stream));
  const int64_t m = std::min<int64_t>(*n_edges_h, cap);
  if (m > 0) CUDA_TRY(cudaMemcpyAsync(edges_h, s->rm_out.p, sizeof(int32_t) * 2 * m, cudaMemcpyDeviceToHost, s->stream));
  if (node_ok_h && n > 0) CUDA_TRY(cudaMemcpyAsync(node_ok_h, s->rm_node_ok.p, (size_t)n, cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return POMP_OK;
}
