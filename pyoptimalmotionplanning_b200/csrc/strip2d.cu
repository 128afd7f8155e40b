// strip2d.cu -- host side of the streaming 2-D nearest-neighbour path: builds the strip index from the row-major
// grid index, counting-sorts a query batch by tile, launches the persistent kernel (strip2d.cuh) and the overflow
// passes.
#include <algorithm>

#include "common.cuh"
#include "grid.cuh"
#include "nn_handle.cuh"
#include "scan.cuh"
#include "strip2d.cuh"

using namespace pomp;

static bool strip_geometry_ok(const pomp_nn* h) {
  const GridDev& g = h->grid;
  return h->grid_valid && h->D == 2 && g.G == 2 && g.gdim[0] == 0 && g.gdim[1] == 1 && !g.periodic[0] &&
         !g.periodic[1] && g.ncell[0] >= ST_TW && g.ncell[1] >= ST_TH;
}

bool pomp::strip_eligible(const pomp_nn* h, int64_t nq, int k) {
  if (h->strip_search == 0 || k != 1 || !strip_geometry_ok(h)) return false;
  if (h->metric.kind != POMP_METRIC_EUCLIDEAN || h->metric.has_cost) return false;
  if (h->skip_ptr() != nullptr || h->n_dev != (int64_t)h->grid.n_indexed) return false;  // masks / unindexed tail
  const GridDev& g = h->grid;
  const long long ntiles = (long long)((g.ncell[0] + ST_TW - 1) / ST_TW) * ((g.ncell[1] + ST_TH - 1) / ST_TH);
  // the kernel streams every tile that holds a query: it only pays when the batch is dense in the index.  The chunks of
  // the pipelined host path each touch nearly every tile again: below 64 queries per tile the gather kernels serve them
  // better (end to end, 2^20 queries in 6 chunks, on a box with slow PCIe: 0.66 -> 0.58 ms; no difference on a fast one)
  const double per_tile = h->pipe_chunks_active > 1 ? std::max(h->strip_min_per_tile, 64.0) : h->strip_min_per_tile;
  return nq < 0x7fffffffLL && (double)nq >= per_tile * (double)ntiles;
}

static int build_strip_index(pomp_nn* h, cudaStream_t st) {
  const GridDev& g = h->grid;
  StripDev sd;
  memset(&sd, 0, sizeof(sd));
  sd.W = g.ncell[0];
  sd.H = g.ncell[1];
  sd.nstrips = (sd.W + ST_TW - 1) / ST_TW;
  sd.nty = (sd.H + ST_TH - 1) / ST_TH;
  sd.ntiles = sd.nstrips * sd.nty;
  const long long nrows = (long long)sd.nstrips * sd.H;
  POMP_REQUIRE(nrows < 0x7fffffffLL, "strip index: too many strip rows");
  POMP_TRY(h->srow.reserve((size_t)2 * nrows + 2));
  POMP_TRY(h->scan_tmp.reserve(scan_scratch_elems(nrows)));
  int* len = h->srow.p + nrows + 1;
  int* start = h->srow.p;
  LAUNCH(strip_row_len_kernel, (int)((nrows + 255) / 256), 256, 0, st, g.cell_start, sd.W, sd.H, sd.nstrips, len);
  POMP_TRY(exclusive_scan<int>(len, nrows, start, h->scan_tmp.p, st));
  int total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, start + nrows, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  POMP_REQUIRE(total >= 0, "strip index: slot count overflow");
  sd.nslots = total;
  POMP_TRY(h->spts2.reserve((size_t)2 * total + 2));
  POMP_TRY(h->sidx2.reserve((size_t)total + 1));
  POMP_TRY(h->tile_info.reserve((size_t)2 * sd.ntiles));
  POMP_TRY(h->tile_cs.reserve((size_t)sd.ntiles * ST_CS_STRIDE));
  LAUNCH(strip_copy_kernel, (int)((nrows * 32 + 255) / 256), 256, 0, st, g.cell_start,
         reinterpret_cast<const double2*>(g.spts), g.sidx, sd.W, sd.H, sd.nstrips, start,
         reinterpret_cast<double2*>(h->spts2.p), h->sidx2.p);
  LAUNCH(strip_tile_kernel, sd.ntiles, 128, 0, st, g.cell_start, sd.W, sd.H, sd.nstrips, sd.nty, start,
         reinterpret_cast<int2*>(h->tile_info.p), h->tile_cs.p);
  CHECK_LAUNCH();
  sd.spts2 = reinterpret_cast<const double2*>(h->spts2.p);
  sd.sidx2 = h->sidx2.p;
  sd.tile_info = reinterpret_cast<const int2*>(h->tile_info.p);
  sd.tile_cs = h->tile_cs.p;
  h->strip = sd;
  h->strip_build = h->grid_builds;
  return POMP_OK;
}

// overflow pass of the strip path: grid-stride over the deferred list with the pruned traversal (the qlist branch of
// grid_knn_fast_kernel), then the last block to finish moves the count to ctrl[2] and clears ctrl[1]
static __global__ void __launch_bounds__(128)
strip_overflow_kernel(GridDev g, const double* __restrict__ pts, long long n, const double* __restrict__ q,
                      int64_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ out_cnt,
                      const int* __restrict__ qlist, int* __restrict__ ctrl) {
  grid_dependency_wait();
  const long long cnt = *(volatile int*)(ctrl + 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += (long long)gridDim.x * blockDim.x)
    knn_pruned_one<2, 1>(g, pts, n, nullptr, q, (long long)qlist[t], 1, out_idx, out_dist, out_cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned tk = atomicInc(reinterpret_cast<unsigned*>(ctrl + 3), gridDim.x - 1);  // wraps back to 0
    if (tk == gridDim.x - 1) {
      ctrl[2] = (int)cnt;
      ctrl[1] = 0;
    }
  }
}

int pomp::strip_nearest(pomp_nn* h, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* cnt,
                        cudaStream_t st) {
  (void)k;  // k == 1 (strip_eligible)
  if (h->strip_build != h->grid_builds) POMP_TRY(build_strip_index(h, st));
  const StripDev& sd = h->strip;
  const GridDev& g = h->grid;
  h->cur = 0;
  SortScratch& S = h->ss[0];
  const int nt = sd.ntiles;
  // sub-bucket capacity: 2x the mean population + slack (Poisson tail ~1e-6); the rest goes to the overflow list
  const double avg = (double)nq / ((double)nt * ST_SUB);
  const int Cs = (int)std::min<double>(1 << 17, ((long long)(2.0 * avg + 8.0) + 3) / 4 * 4);
  const size_t slots = (size_t)nt * ST_SUB * Cs;
  const size_t nctrl = (size_t)ST_CNT_STRIDE * (1 + (size_t)nt * ST_SUB);
  {
    const int* before = S.strip_ctrl.p;
    POMP_TRY(S.strip_ctrl.reserve(nctrl));
    if (S.strip_ctrl.p != before || S.strip_ctrl_n != nctrl) S.strip_ctrl_clean = false;
  }
  POMP_TRY(S.qsorted.reserve(slots * 4));  // 32-byte records
  POMP_TRY(S.ovf.reserve((size_t)nq + 2));
  // [1] deferred-to-pruned count, [2] its value at the end of the last step (statistics), [3] ticket of the overflow
  // pass, [ST_CNT_STRIDE (1 + t ST_SUB + sub)] bucket counters
  int* ctrl = S.strip_ctrl.p;
  int* l2 = S.ovf.p;
  StripQRec* qrec = reinterpret_cast<StripQRec*>(S.qsorted.p);
  // the kernels leave the counters zeroed; only a fresh / resized buffer or a step that failed half-way is cleared here
  if (!S.strip_ctrl_clean) CUDA_TRY(cudaMemsetAsync(ctrl, 0, sizeof(int) * nctrl, st));
  S.strip_ctrl_clean = false;
  S.strip_ctrl_n = nctrl;
  const double2* q2 = reinterpret_cast<const double2*>(q);
  LAUNCH(strip_qbucket_kernel, (int)((nq + 256 * ST_QPT - 1) / (256 * ST_QPT)), 256, 0, st, g, sd.nty, q2, (long long)nq,
         Cs, ctrl, qrec, l2);
  // ---- stage geometry from the cell occupancy: expected tile (+halo) population + 6 sigma (Poisson) + slack
  const double ppc = (double)g.n_indexed / (double)g.n_cells;
  const double mean = ST_NCELL * ppc;
  int cap = ((int)(mean + 6.0 * sqrt(mean) + 32.0) + 15) / 16 * 16;
  const size_t budget = 224 * 1024;
  while (ST_SMEM_HEAD + 2 * strip_stage_bytes(cap) > budget) cap -= 16;  // always at least two groups
  if (cap > 65520) cap = 65520;                                        // uint16 cell offsets
  const int nstages = (int)std::min<size_t>(ST_MAX_STAGES, (budget - ST_SMEM_HEAD) / strip_stage_bytes(cap));
  const size_t smem = ST_SMEM_HEAD + (size_t)nstages * strip_stage_bytes(cap);
  CUDA_TRY(cudaFuncSetAttribute(strip_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(h, st);
  LAUNCH_PDL(strip_nn_kernel, sm_count(h->device), ST_THREADS, smem, st, g, sd, (const double*)h->d_pts, q2,
             (const StripQRec*)qrec, ctrl, Cs, cap, nstages, idx, dist, cnt, l2, (int)h->strip_spin_ns,
             (int)h->strip_probe);
  prof_end(h, st);
  h->strip_launches++;
  h->ovf_count_dev = ctrl + 2;
  // ---- whatever neither the tiles nor the 5x5 retries could certify: pruned traversal; its last block hands the
  // deferred counter back zeroed (copy in ctrl[2] for the statistics)
  const int oblocks = sm_count(h->device) * 4;
  LAUNCH_PDL(strip_overflow_kernel, oblocks, 128, 0, st, g, (const double*)h->d_pts, (long long)h->n_dev, q, idx, dist,
             cnt, (const int*)l2, ctrl);
  CHECK_LAUNCH();
  S.ovf_zeroed = false;
  S.strip_ctrl_clean = true;
  return POMP_OK;
}
