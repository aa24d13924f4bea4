// microbench_tree.cu -- where do the clocks of one tree-level column step go?  Includes the library's tsqr.cu with
// GNK_TREE_CLOCK defined (per-warp phase timers in tsqr_tree_kernel) and runs one level on a stack of 8 triangles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I gauss_newton_via_generalized_krylov_subspaces_b200/csrc \
//        -o /tmp/mb_tree tools/microbench_tree.cu && /tmp/mb_tree
// Development aid.
#define GNK_TREE_CLOCK 1
#include "../gauss_newton_via_generalized_krylov_subspaces_b200/csrc/tsqr.cu"
#include <vector>
#include <cstdlib>
void gnk_set_error(const std::string&) {}
int gnk_fail(const char* w, cudaError_t e, const char*, int) { printf("fail %s %s\n", w, cudaGetErrorString(e)); return -1; }
int gnk_comm_allgather_doubles(gnk_ctx*, const double*, double*, int64_t, void*) { return 0; }
int main() {
  for (int variant = 0; variant < 2; ++variant)
  for (int k : {8, 15, 30}) {
    const int c = k + 1, fan = 256 / c, cnt = fan;
    std::vector<double> h((size_t)cnt * c * c, 0.0);
    for (int t = 0; t < cnt; ++t)
      for (int r = 0; r < c; ++r)
        for (int cc = r; cc < c; ++cc) h[((size_t)t * c + r) * c + cc] = (rand() / (double)RAND_MAX) - 0.5 + (r == cc ? 3.0 : 0.0);
    double *din, *dout, *dres;
    cudaMalloc(&din, h.size() * 8); cudaMalloc(&dout, c * c * 8 * 4); cudaMalloc(&dres, 4096);
    cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(double) * ((size_t)c * c + 2 * 256 + 2 + 2 * c);
    long long zero[64] = {0};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemcpyToSymbol(g_tree_clk, zero, sizeof(zero));
      cudaEventRecord(e0);
      if (variant == 0) {
        if (c <= 16) tsqr_tree_kernel<2, false><<<1, TPB, smem>>>(din, cnt, fan, k, dout, 0, dres);
        else tsqr_tree_kernel<4, false><<<1, TPB, smem>>>(din, cnt, fan, k, dout, 0, dres);
      } else {
        if (c <= 16) tsqr_tree_kernel<2, true><<<1, TPB, smem>>>(din, cnt, fan, k, dout, 0, dres);
        else tsqr_tree_kernel<4, true><<<1, TPB, smem>>>(din, cnt, fan, k, dout, 0, dres);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk[64]; cudaMemcpyFromSymbol(clk, g_tree_clk, sizeof(clk));
    printf("%s k=%d: kernel %.1f us (%s), %d steps; per step clocks [lds, scalars+dots, finish, update+publish, barrier]:\n", variant ? "SHFL" : "DMMA", k, ms * 1e3,
           cudaGetErrorString(cudaGetLastError()), c - 1);
    for (int w = 0; w < 8; ++w) {
      printf("  warp %d:", w);
      long long tot = 0;
      for (int s = 0; s < 5; ++s) { printf(" %7.0f", (double)clk[w * 8 + s] / (c - 1)); tot += clk[w * 8 + s]; }
      printf("   total %7.0f\n", (double)tot / (c - 1));
    }
  }
  return 0;
}
