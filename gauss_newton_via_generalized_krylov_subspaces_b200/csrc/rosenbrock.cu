// rosenbrock.cu -- device-native callbacks of the chained Rosenbrock problem (rosenbrock_problem.py:8-19), so that
// the rosenbrock_test.py scenarios run without a host round trip per residual / Jacobian evaluation.
//   res(x)  = sqrt2 * [10 (x[1:] - x[:-1]^2) ; 1 - x[:-1]]                       (2p-2 residuals)
//   jac(x)  = COO block array, 3(p-1) non-zeros; here the VALUES of its CSR form and of the CSR form of its transpose
//             (fixed sparsity structure, uploaded once by the host: rosenbrock_problem.py in this package).
// Every product is rounded separately (no FMA contraction), in numpy's order, so F and the Jacobian values are
// bit-identical to the reference's.  p = 1000: launch-latency regime, one thread per row.
#include "common.cuh"

namespace {
constexpr int TPB = 256;

__global__ void __launch_bounds__(TPB) rosenbrock_residual_kernel(int64_t p, double sqrt2, const double* __restrict__ x,
                                                                   double* __restrict__ F) {
  const int64_t q = p - 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q) return;
  const double xi = x[i];
  F[i] = __dmul_rn(sqrt2, __dmul_rn(10.0, __dsub_rn(x[i + 1], __dmul_rn(xi, xi))));
  F[q + i] = __dmul_rn(sqrt2, __dsub_rn(1.0, xi));
}

// CSR of J  (2q x p): row i < q: (i, -20 sqrt2 x_i), (i+1, 10 sqrt2);  row q+i: (i, -sqrt2)
// CSR of J^T (p x 2q): row 0: (0, v_0), (q, -sqrt2);  row j = 1..q-1: (j-1, 10 sqrt2), (j, v_j), (q+j, -sqrt2);
//                      row q: (q-1, 10 sqrt2);   row pointer of J^T: 0, 2, 5, ..., 3j-1, ..., 3q
__global__ void __launch_bounds__(TPB) rosenbrock_jacobian_kernel(int64_t p, double sqrt2, const double* __restrict__ x,
                                                                   double* __restrict__ val,
                                                                   double* __restrict__ val_t) {
  const int64_t q = p - 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q) return;
  const double v = __dmul_rn(sqrt2, __dmul_rn(-20.0, x[i]));
  const double up = __dmul_rn(sqrt2, 10.0), lo = __dmul_rn(sqrt2, -1.0);
  val[2 * i] = v;
  val[2 * i + 1] = up;
  val[2 * q + i] = lo;
  const int64_t b = (i == 0) ? 0 : 3 * i - 1;   // first entry of row i of J^T
  if (i > 0) val_t[b] = up;                     // from row i-1 of J
  val_t[b + (i > 0)] = v;
  val_t[b + (i > 0) + 1] = lo;
  if (i == q - 1) val_t[3 * q - 1] = up;        // row q of J^T
}
}  // namespace

extern "C" {

int gnk_rosenbrock_residual(gnk_ctx* ctx, int64_t p, double sqrt2, const double* d_x, double* d_F, void* stream) {
  GNK_REQUIRE(ctx && d_x && d_F && p >= 2, "gnk_rosenbrock_residual: bad argument");
  rosenbrock_residual_kernel<<<(unsigned)ceil_div(p - 1, TPB), TPB, 0, (cudaStream_t)stream>>>(p, sqrt2, d_x, d_F);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

int gnk_rosenbrock_jacobian(gnk_ctx* ctx, int64_t p, double sqrt2, const double* d_x, double* d_val, double* d_val_t,
                            void* stream) {
  GNK_REQUIRE(ctx && d_x && d_val && d_val_t && p >= 2, "gnk_rosenbrock_jacobian: bad argument");
  rosenbrock_jacobian_kernel<<<(unsigned)ceil_div(p - 1, TPB), TPB, 0, (cudaStream_t)stream>>>(p, sqrt2, d_x, d_val,
                                                                                             d_val_t);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
