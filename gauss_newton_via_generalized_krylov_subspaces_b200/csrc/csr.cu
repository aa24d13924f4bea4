// csr.cu -- generic sparse Jacobians (rosenbrock_problem.py:14-19 and Jacobians returned by foreign
// Python callables: CSR / COO / dense are all uploaded as CSR of J and CSR of J^T).
//   gnk_spmm_csr       out[:, j] = sign * A * in[:, j]   (J @ V_k, gauss_newton_krylow.py:86; with the CSR
//                                                        of A^T: -J^T r, krylow.py:62)
//   gnk_csr_row_sumsq  per-row sum of squares (on the CSR of A^T: diag(A^T A), gauss_newton.py:50-52)
// These problems are tiny (p = 1000, 2, 1): the kernels are latency-bound by construction; one thread
// walks one row in index order, which is also scipy's csr_matvec summation order.
#include "common.cuh"

namespace {
constexpr int TPB = 128;

__global__ void __launch_bounds__(TPB) spmm_csr_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ col,
                                                        const double* __restrict__ val, const double* __restrict__ in,
                                                        int64_t in_ld, int64_t in_off, double sign,
                                                        double* __restrict__ out, int64_t out_ld, int64_t out_off) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const double* x = in + (int64_t)blockIdx.y * in_ld + in_off;
  double acc = 0.0;
  const int e1 = rowptr[row + 1];
  for (int e = rowptr[row]; e < e1; ++e) acc = fma(val[e], x[col[e]], acc);
  out[(int64_t)blockIdx.y * out_ld + out_off + row] = sign * acc;
}

__global__ void __launch_bounds__(TPB) row_sumsq_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr,
                                                         const double* __restrict__ val, double* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  double acc = 0.0;
  const int e1 = rowptr[row + 1];
  for (int e = rowptr[row]; e < e1; ++e) acc = fma(val[e], val[e], acc);
  out[row] = acc;
}
}  // namespace

extern "C" {

int gnk_spmm_csr(gnk_ctx* ctx, int64_t n_rows, const int32_t* d_rowptr, const int32_t* d_col, const double* d_val,
                 const double* d_in, int64_t in_ld, int64_t in_off, int k, double sign, double* d_out, int64_t out_ld,
                 int64_t out_off, void* stream) {
  GNK_REQUIRE(ctx && d_rowptr && d_in && d_out, "gnk_spmm_csr: null argument");
  GNK_REQUIRE(k >= 1 && k <= 65535 && n_rows >= 0, "gnk_spmm_csr: bad sizes");
  if (n_rows == 0) return 0;
  dim3 grid((unsigned)ceil_div(n_rows, TPB), k);
  spmm_csr_kernel<<<grid, TPB, 0, (cudaStream_t)stream>>>(n_rows, d_rowptr, d_col, d_val, d_in, in_ld, in_off, sign,
                                                          d_out, out_ld, out_off);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

int gnk_csr_row_sumsq(gnk_ctx* ctx, int64_t n_rows, const int32_t* d_rowptr, const double* d_val, double* d_out,
                      void* stream) {
  GNK_REQUIRE(ctx && d_rowptr && d_out && n_rows >= 0, "gnk_csr_row_sumsq: bad argument");
  if (n_rows == 0) return 0;
  row_sumsq_kernel<<<(unsigned)ceil_div(n_rows, TPB), TPB, 0, (cudaStream_t)stream>>>(n_rows, d_rowptr, d_val, d_out);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
