// Microbenchmark behind the query bucketing pass of strip2d.cuh: which part of "load query -> atomic slot ->
// scattered 16 B + 4 B stores" bounds the pass?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o qb qbucket_variants.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <random>

template <int MODE, int QPT>
__global__ void __launch_bounds__(256) k(const double2* __restrict__ q, long long nq, int W, int nty, int Cs, int sub,
                                         int stride, int* __restrict__ ctrl, double2* __restrict__ qs,
                                         int* __restrict__ qp, int* sink) {
  const long long i0 = (long long)blockIdx.x * (blockDim.x * QPT) + threadIdx.x;
  double2 v[QPT];
  long long b[QPT];
  int slot[QPT];
#pragma unroll
  for (int u = 0; u < QPT; u++) {
    long long i = i0 + (long long)u * blockDim.x;
    v[u] = i < nq ? __ldg(q + i) : make_double2(0, 0);
  }
  int acc = 0;
#pragma unroll
  for (int u = 0; u < QPT; u++) {
    long long i = i0 + (long long)u * blockDim.x;
    int cx = min(W - 1, (int)(v[u].x * W)), cy = min(W - 1, (int)(v[u].y * W));
    int t = (cx / 64) * nty + cy / 16;
    b[u] = (long long)t * sub + ((threadIdx.x + u) & (sub - 1));
    if (MODE == 1) slot[u] = (int)((i * 2654435761u) >> 8) & (Cs - 1) & 15;  // no atomic
    else slot[u] = i < nq ? atomicAdd(ctrl + (long long)stride * (1 + b[u]), 1) : 0;
  }
#pragma unroll
  for (int u = 0; u < QPT; u++) {
    long long i = i0 + (long long)u * blockDim.x;
    if (i >= nq) continue;
    if (MODE == 2) { acc += slot[u]; continue; }           // atomics only
    if (MODE == 4 || MODE == 5) {                          // one 32-byte record per query: full-sector writes
      if (slot[u] < Cs) {
        long long d = b[u] * Cs + slot[u];
        double4* rec = reinterpret_cast<double4*>(qs) + d;
        double w = __longlong_as_double((long long)i);
        if (MODE == 4) {
          asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec), "d"(v[u].x), "d"(v[u].y), "d"(w), "d"(0.0) : "memory");
        } else {
          reinterpret_cast<double2*>(rec)[0] = v[u];
          reinterpret_cast<double2*>(rec)[1] = make_double2(w, 0.0);
        }
      }
      continue;
    }
    if (slot[u] < Cs) {
      long long d = b[u] * Cs + slot[u];
      if (MODE != 3) qs[d] = v[u];
      qp[d] = (int)i;
    }
  }
  if (MODE == 2 && acc == -12345) *sink = acc;
}

int main() {
  const long long nq = 1 << 20;
  const int W = 3344, nty = 209, nt = 53 * 209;
  std::vector<double2> h(nq);
  std::mt19937_64 rng(1);
  std::uniform_real_distribution<double> U(0, 1);
  for (auto& p : h) p = make_double2(U(rng), U(rng));
  double2* q; cudaMalloc(&q, nq * 16); cudaMemcpy(q, h.data(), nq * 16, cudaMemcpyHostToDevice);
  int* ctrl; size_t nctrl = 8ull * (1 + (size_t)nt * 8); cudaMalloc(&ctrl, nctrl * 4);
  double2* qs; int* qp; size_t slots = (size_t)nt * 8 * 64; cudaMalloc(&qs, slots * 32); cudaMalloc(&qp, slots * 4);
  int* sink; cudaMalloc(&sink, 4);
  char* flush; cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, int qpt, int Cs, int sub, int stride) {
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
      cudaMemset(ctrl, 0, nctrl * 4);
      cudaMemset(flush, r, 256 << 20);  // flush L2
      cudaEventRecord(e0);
      kern<<<(unsigned)((nq + 256 * qpt - 1) / (256 * qpt)), 256>>>(q, nq, W, nty, Cs, sub, stride, ctrl, qs, qp, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-52s %.1f us  (%s)\n", name, best * 1e3, cudaGetErrorString(cudaGetLastError()));
  };
  run("full: sub=8 Cs=32 stride=8 qpt=4", k<0, 4>, 4, 32, 8, 8);
  run("full: sub=8 Cs=32 stride=8 qpt=1", k<0, 1>, 1, 32, 8, 8);
  run("full: sub=8 Cs=32 stride=1 qpt=4", k<0, 4>, 4, 32, 8, 1);
  run("full: sub=1 Cs=256 stride=8 qpt=4", k<0, 4>, 4, 256, 1, 8);
  run("full: sub=8 Cs=16 stride=8 qpt=4 (dense)", k<0, 4>, 4, 16, 8, 8);
  run("no atomic, stores only: sub=8 Cs=32", k<1, 4>, 4, 32, 8, 8);
  run("atomics only: sub=8 stride=8", k<2, 4>, 4, 32, 8, 8);
  run("atomics only: sub=8 stride=1", k<2, 4>, 4, 32, 8, 1);
  run("atomics only: sub=1 stride=8", k<2, 4>, 4, 256, 1, 8);
  run("atomics only: sub=16 stride=1", k<2, 4>, 4, 16, 16, 1);
  run("atomics + qperm store only: sub=8", k<3, 4>, 4, 32, 8, 8);
  run("32 B record, one v4.b64 store: sub=8 Cs=32", k<4, 4>, 4, 32, 8, 8);
  run("32 B record, two 16 B stores: sub=8 Cs=32", k<5, 4>, 4, 32, 8, 8);
  run("32 B record, one v4.b64 store: sub=8 Cs=32 qpt=1", k<4, 1>, 1, 32, 8, 8);
  run("32 B record, v4 store: sub=4 Cs=64", k<4, 4>, 4, 64, 4, 8);
  return 0;
}
