// microbench_quad.cu -- column-step time of the warp-autonomous TSQR leaf with 1, 4 and 8 active warps on ONE SM
// (1 warp = the pure dependency chain, 4 = one warp per scheduler, 8 = the production occupancy), to separate the
// critical path from FP64-pipe contention.  Includes the library's tsqr.cu.  Development aid.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I gauss_newton_via_generalized_krylov_subspaces_b200/csrc \
//        -o /tmp/mb_quad tools/microbench_quad.cu && /tmp/mb_quad
#include "../gauss_newton_via_generalized_krylov_subspaces_b200/csrc/tsqr.cu"
#include <vector>
#include <cstdlib>
void gnk_set_error(const std::string&) {}
int gnk_fail(const char* w, cudaError_t e, const char*, int) { printf("fail %s %s\n", w, cudaGetErrorString(e)); return -1; }
int gnk_comm_allgather_doubles(gnk_ctx*, const double*, double*, int64_t, void*) { return 0; }
template <int CPL, int RPL, int MINB>
void run(int k, const double* dA, int64_t lda, const double* dy, double* dR, double ghz) {
  using P_t = QuadPanel<CPL, RPL, false>;
  const size_t smem = sizeof(double) * (size_t)NWARP * (P_t::CP * P_t::CP + 2 * P_t::RT + P_t::STAGE);
  auto kern = tsqr_quad_kernel<CPL, RPL, false, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int T = 100;
  for (int W : {1, 4, 8}) {
    const int64_t n_tiles = (int64_t)T * W, n_rows = n_tiles * P_t::RT;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      kern<<<1, TPB, smem>>>(dA, lda, dy, -1.0, k, n_rows, n_tiles, T, dR);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double clk_per_step = best * 1e-3 * ghz * 1e9 / ((double)T * k);
    printf("k=%2d <%d,%d>: %d active warp(s) on one SM: %8.1f us per kernel, %7.0f clk per column step and warp, %6.1f clk per step and row-of-64 tile throughput/SM: %6.0f clk\n",
           k, CPL, RPL, W, best * 1e3, clk_per_step, clk_per_step / 1.0, clk_per_step / W);
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  const int64_t lda = 64 * 100 * 8;
  std::vector<double> h((size_t)lda * 32);
  for (auto& x : h) x = rand() / (double)RAND_MAX - 0.5;
  double *dA, *dy, *dR;
  cudaMalloc(&dA, h.size() * 8); cudaMalloc(&dy, lda * 8); cudaMalloc(&dR, 32 * 32 * 8 * 16);
  cudaMemcpy(dA, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, h.data(), lda * 8, cudaMemcpyHostToDevice);
  printf("%s, %.3f GHz\n", p.name, ghz);
  run<4, 16, 1>(30, dA, lda, dy, dR, ghz);
  run<4, 16, 1>(24, dA, lda, dy, dR, ghz);
  run<3, 16, 1>(20, dA, lda, dy, dR, ghz);
  run<2, 16, 2>(15, dA, lda, dy, dR, ghz);
  run<2, 16, 2>(8, dA, lda, dy, dR, ghz);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
