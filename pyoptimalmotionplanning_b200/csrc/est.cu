// est.cu -- the projection-hash density of ESTWithProjections (pomp/planners/kinodynamicplanner.py:746-868).
//
// Reference: every node is hashed, per projection basis b (a set of <= 4 coordinates k with scales v_bk), to the
// integer cell key_b(x) = (int(x[k]*v_bk))_k (onAddNode :817-832; int() truncates towards zero), kept in one Python
// dictionary per basis; density(x) (:834-856) = sum_b count_b(key_b(x)) * pow(level_b, 1) / n_bases, the sums taken
// in basis order.  Here the node set lives on the device as coordinates; a warp per query counts, per basis, the
// nodes whose key equals the query's (keys are recomputed on the fly: <= 4 multiplies and truncations per node and
// basis), lanes striding the nodes, and lane 0 folds the counts in the reference's order.  Counts are integers and
// the fold uses the same float operations, so the result is bit-identical.
#include "common.cuh"

using namespace pomp;

namespace {

constexpr int EST_MAX_BASES = 64;
constexpr int EST_BASIS_DIM = 4;

struct EstBases {
  int n_bases;
  int dim[EST_MAX_BASES];                      // coordinates per basis (<= 4)
  int idx[EST_MAX_BASES][EST_BASIS_DIM];
  double scale[EST_MAX_BASES][EST_BASIS_DIM];
  double weight[EST_MAX_BASES];                // pow(level, len(basisvector)) = level
};

__device__ __forceinline__ long long trunc_ll(double v) {
  // Python int(float): towards zero.  Values beyond the int64 range saturate (the reference would make a big int)
  if (!(v == v)) return 0x7ffffffffffffffeLL;
  if (v >= 9.2e18) return 0x7fffffffffffffffLL;
  if (v <= -9.2e18) return -0x7fffffffffffffffLL;
  return (long long)v;
}

__global__ void __launch_bounds__(256)
est_density_kernel(const double* __restrict__ pts, long long n, int D, const EstBases* __restrict__ Bp,
                   const double* __restrict__ q, long long nq, double* __restrict__ out) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nq) return;
  const EstBases& B = *Bp;
  double c = 0.0;
  for (int b = 0; b < B.n_bases; b++) {
    long long key[EST_BASIS_DIM];
    const int bd = B.dim[b];
    for (int i = 0; i < bd; i++) key[i] = trunc_ll(0.0 + q[w * D + B.idx[b][i]] * B.scale[b][i]);
    int cnt = 0;
    for (long long j = lane; j < n; j += 32) {
      bool same = true;
      for (int i = 0; i < bd; i++) same = same && (trunc_ll(0.0 + __ldg(pts + j * D + B.idx[b][i]) * B.scale[b][i]) == key[i]);
      cnt += same ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    c += (double)cnt * B.weight[b];
  }
  if (lane == 0) out[w] = c / (double)B.n_bases;
}

}  // namespace

struct pomp_est {
  int device = 0;
  int D = 0;
  EstBases bases;
  EstBases* d_bases = nullptr;
  DevBuf<double> pts, q, out;
  int64_t n = 0, cap = 0;
  cudaStream_t stream = nullptr;
};

POMP_API int pomp_est_create(int32_t dim, int32_t n_bases, const int32_t* basis_dim_h, const int32_t* basis_idx_h,
                             const double* basis_scale_h, const double* basis_weight_h, int device, pomp_est** out) {
  POMP_REQUIRE(out && basis_dim_h && basis_idx_h && basis_scale_h && basis_weight_h, "NULL argument to pomp_est_create");
  POMP_REQUIRE(dim >= 1 && dim <= POMP_MAX_DIM && n_bases >= 1 && n_bases <= EST_MAX_BASES, "bad sizes for pomp_est_create");
  DeviceGuard g(device);
  if (!g.ok) return POMP_ERR_CUDA;
  pomp_est* e = new pomp_est();
  e->device = device;
  e->D = dim;
  memset(&e->bases, 0, sizeof(e->bases));
  e->bases.n_bases = n_bases;
  for (int b = 0; b < n_bases; b++) {
    const int bd = basis_dim_h[b];
    if (bd < 1 || bd > EST_BASIS_DIM) {
      delete e;
      set_error("basis %d has %d coordinates (1..4 supported)", b, bd);
      return POMP_ERR_INVALID;
    }
    e->bases.dim[b] = bd;
    e->bases.weight[b] = basis_weight_h[b];
    for (int i = 0; i < bd; i++) {
      const int k = basis_idx_h[b * EST_BASIS_DIM + i];
      if (k < 0 || k >= dim) {
        delete e;
        set_error("basis %d refers to coordinate %d", b, k);
        return POMP_ERR_INVALID;
      }
      e->bases.idx[b][i] = k;
      e->bases.scale[b][i] = basis_scale_h[b * EST_BASIS_DIM + i];
    }
  }
  if (cudaMalloc((void**)&e->d_bases, sizeof(EstBases)) != cudaSuccess ||
      cudaMemcpy(e->d_bases, &e->bases, sizeof(EstBases), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("pomp_est_create: %s", cudaGetErrorString(cudaGetLastError()));
    if (e->d_bases) cudaFree(e->d_bases);
    delete e;
    return POMP_ERR_CUDA;
  }
  *out = e;
  return POMP_OK;
}

POMP_API int pomp_est_destroy(pomp_est* e) {
  if (!e) return POMP_OK;
  DeviceGuard g(e->device);
  if (e->d_bases) cudaFree(e->d_bases);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return POMP_OK;
}

POMP_API int pomp_est_reset(pomp_est* e) {
  POMP_REQUIRE(e, "NULL handle");
  e->n = 0;
  return POMP_OK;
}

POMP_API int64_t pomp_est_size(const pomp_est* e) { return e ? e->n : 0; }

// onAddNode: append nodes
POMP_API int pomp_est_add_h(pomp_est* e, const double* pts_h, int64_t n) {
  POMP_REQUIRE(e && (n == 0 || pts_h) && n >= 0, "bad arguments to pomp_est_add_h");
  if (n == 0) return POMP_OK;
  DeviceGuard g(e->device);
  if (!g.ok) return POMP_ERR_CUDA;
  if (e->n + n > e->cap) {
    const int64_t ncap = std::max<int64_t>(2 * e->cap, e->n + n + 1024);
    DevBuf<double> nb;
    POMP_TRY(nb.reserve((size_t)ncap * e->D));
    if (e->n) CUDA_TRY(cudaMemcpy(nb.p, e->pts.p, sizeof(double) * e->n * e->D, cudaMemcpyDeviceToDevice));
    std::swap(nb.p, e->pts.p);
    std::swap(nb.cap, e->pts.cap);
    e->cap = ncap;
  }
  CUDA_TRY(cudaMemcpyAsync(e->pts.p + e->n * e->D, pts_h, sizeof(double) * n * e->D, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->n += n;
  return POMP_OK;
}

// density(x) for a batch of states (device pointers; nq warps)
POMP_API int pomp_est_density(pomp_est* e, const double* q, int64_t nq, double* out, void* stream) {
  POMP_REQUIRE(e && (nq == 0 || (q && out)) && nq >= 0, "bad arguments to pomp_est_density");
  if (nq == 0) return POMP_OK;
  DeviceGuard g(e->device);
  if (!g.ok) return POMP_ERR_CUDA;
  LAUNCH(est_density_kernel, (int)((nq * 32 + 255) / 256), 256, 0, (cudaStream_t)stream, e->pts.p, (long long)e->n,
         e->D, e->d_bases, q, (long long)nq, out);
  CHECK_LAUNCH();
  return POMP_OK;
}

POMP_API int pomp_est_density_h(pomp_est* e, const double* q_h, int64_t nq, double* out_h) {
  POMP_REQUIRE(e && (nq == 0 || (q_h && out_h)) && nq >= 0, "bad arguments to pomp_est_density_h");
  if (nq == 0) return POMP_OK;
  DeviceGuard g(e->device);
  if (!g.ok) return POMP_ERR_CUDA;
  POMP_TRY(e->q.reserve((size_t)nq * e->D));
  POMP_TRY(e->out.reserve((size_t)nq));
  CUDA_TRY(cudaMemcpyAsync(e->q.p, q_h, sizeof(double) * nq * e->D, cudaMemcpyHostToDevice, e->stream));
  POMP_TRY(pomp_est_density(e, e->q.p, nq, e->out.p, e->stream));
  CUDA_TRY(cudaMemcpyAsync(out_h, e->out.p, sizeof(double) * nq, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return POMP_OK;
}
