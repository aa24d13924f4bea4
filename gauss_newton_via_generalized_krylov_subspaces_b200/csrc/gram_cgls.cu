// gram_cgls.cu -- the projected least squares of GNK solved by CGLS (BASELINE config 5: "Krylov dim 50 with CGLS inner
// solve"): Jacobi-preconditioned conjugate gradients on the normal equations  A^T A d = A^T y  of the dense n x k panel
// A = sign * J V_k, i.e. the reference's cg_least_squares(A, y, cg_rtol, preconditioner=True) (gauss_newton.py:11-60 on
// scipy's cg: x0 = 0, stop when |r|_2 < rtol |A^T y|_2 tested at the top of every iteration, at most 10 k iterations,
// M = 1 / diag(A^T A)).
//
// scipy's cg touches A only through  p -> A^T (A p): 16 n k bytes per CG iteration, up to k iterations per outer step --
// at k = 50 that is ~50 sweeps over a 27 GB panel.  Here the k x k operator G = A^T A (and A^T y, y^T y) is formed ONCE
// per outer iteration -- one sweep over the panel with FP64 tensor-core MMAs, 5..7 column blocks wide, the same fragment
// scheme as cholqr_gram_kernel -- and the whole CG recurrence then runs inside ONE single-CTA kernel on G (k^2 flops per
// iteration, no host round trips).  In exact arithmetic the iterates are those of scipy's cg on the LinearOperator; in
// floating point both are CG on the normal equations (conditioning cond(A)^2 either way), and the stopping rule is
// the loose cg_rtol of the inexact solve.  ||A d||^2 = d^T G d for the Armijo rule comes out of G exactly.
// Multi-GPU: every rank forms the Gram matrix of its slab; the single CTA sums them over the ranks in rank order through
// the peer mailboxes (bit-identical on all ranks) before it iterates.
#include <stdint.h>

#include "common.cuh"

int gnk_comm_allgather_doubles(gnk_ctx* ctx, const double* d_send, double* d_recv, int64_t count, void* stream);

namespace {

constexpr int WG = 8;            // warps per CTA (one CTA per SM, up to 255 registers per thread)
constexpr int WT = 32 * WG;
constexpr int WMAXB = 7;         // widest panel: 56 columns (k <= 55)
constexpr int WMAXC = 8 * WMAXB;

__host__ __device__ constexpr int nblocks(int NB) { return NB * (NB + 1) / 2; }
__host__ __device__ constexpr int blk_index(int NB, int I, int J) { return I * NB - I * (I - 1) / 2 + (J - I); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

struct Panel {
  const double* A;
  int64_t lda;
  const double* y;
  int k;
  int64_t n_rows;
};

// G = P^T P for the panel P = [A | y] with NB column blocks of 8; lane (g, t) holds rows 2t, 2t+1 of column 8 I + g of
// an 8-row group: both the A and the B fragment of the DMMAs (see cholqr.cu).  A warp owns 8 rows per step.
template <int NB>
__global__ void __launch_bounds__(WT, 1)
    gram_wide_kernel(Panel src, int64_t rows_per_cta, double* __restrict__ partials, unsigned int* ticket,
                     double* __restrict__ Gout) {
  constexpr int NBLK = nblocks(NB);
  constexpr int NE = NBLK * 64;
  __shared__ __align__(16) double red[NE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const double* cp[NB];
#pragma unroll
  for (int I = 0; I < NB; ++I) {
    const int col = 8 * I + g;
    cp[I] = col < src.k ? src.A + (int64_t)col * src.lda : (col == src.k ? src.y : nullptr);
  }
  double acc[NBLK][2];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) acc[b][0] = acc[b][1] = 0.0;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t limit = row0 + rows_per_cta;
  if (limit > src.n_rows) limit = src.n_rows;
  // 8 rows per warp and step, the next step's rows already in flight (two register tiles): with a single tile the
  // loads and the DMMAs of a warp alternated and the kernel ran at 0.33 of the HBM roofline (config 5, k <= 50)
  constexpr int64_t STEP = 8 * WG;
  auto load8 = [&](double2 (&v)[NB], int64_t r) {
#pragma unroll
    for (int I = 0; I < NB; ++I) {
      const int64_t rr = r + 2 * t;
      v[I] = (cp[I] != nullptr && rr < limit) ? __ldcs(reinterpret_cast<const double2*>(cp[I] + rr))
                                              : make_double2(0.0, 0.0);
    }
  };
  auto mma8 = [&](const double2 (&v)[NB]) {
#pragma unroll
    for (int I = 0; I < NB; ++I)
#pragma unroll
      for (int J = I; J < NB; ++J) {
        const int b = blk_index(NB, I, J);
        dmma(acc[b][0], acc[b][1], v[I].x, v[J].x);
        dmma(acc[b][0], acc[b][1], v[I].y, v[J].y);
      }
  };
  {
    double2 va[NB], vb[NB];
    int64_t r = row0 + 8 * warp;
    load8(va, r);
    while (r < limit) {
      load8(vb, r + STEP);
      mma8(va);
      r += STEP;
      if (r >= limit) break;
      load8(va, r + STEP);
      mma8(vb);
      r += STEP;
    }
  }
  // CTA sum in warp order, CTA partial, last CTA adds the partials in CTA order (deterministic)
  for (int w = 0; w < WG; ++w) {
    if (warp == w) {
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        double2* p = reinterpret_cast<double2*>(red + b * 64 + lane * 2);
        double2 x = make_double2(acc[b][0], acc[b][1]);
        if (w > 0) {
          const double2 o = *p;
          x.x += o.x;
          x.y += o.y;
        }
        *p = x;
      }
    }
    __syncthreads();
  }
  double* mine = partials + (int64_t)blockIdx.x * NE;
  for (int e = threadIdx.x; e < NE; e += WT) mine[e] = red[e];
  __threadfence();
  if (!grid_arrive_last(ticket)) return;
  const int nb = gridDim.x;
  for (int e = threadIdx.x; e < NE; e += WT) {
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += __ldcg(partials + (int64_t)b * NE + e);
    Gout[e] = s;
  }
}

// ---- the CG recurrence on the k x k system, one CTA of 1024 threads ---------------------------------------------------
constexpr int CT = 1024;
constexpr int GLD = WMAXC + 1;

// sum over the ranks (rank order) of `count` doubles in vals, count may exceed the block size
__device__ void p2p_allreduce_long(const gnk_p2p_dev& pd, double* vals, int count) {
  const int parity = (int)(pd.seq & 1ull);
  for (int r = 0; r < pd.nranks; ++r) {
    double* dst = reinterpret_cast<double*>(static_cast<char*>(pd.peers[r]) + p2p_gather_off(pd.nranks, parity, pd.rank));
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  char* mine = static_cast<char*>(pd.peers[pd.rank]);
  if ((int)threadIdx.x < pd.nranks) {
    st_release_sys(reinterpret_cast<unsigned long long*>(pd.peers[threadIdx.x]) + pd.rank, pd.seq);
    wait_flag(reinterpret_cast<const unsigned long long*>(mine) + threadIdx.x, pd.seq);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double a = ld_volatile(reinterpret_cast<const double*>(mine + p2p_gather_off(pd.nranks, parity, 0)) + i);
    for (int r = 1; r < pd.nranks; ++r)
      a += ld_volatile(reinterpret_cast<const double*>(mine + p2p_gather_off(pd.nranks, parity, r)) + i);
    vals[i] = a;
  }
  __syncthreads();
}

__device__ __forceinline__ double cta_sum(double v, double* sh) {  // result in every thread
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < CT / 32; ++w) r += sh[w];  // fixed order
  return r;
}

__global__ void __launch_bounds__(CT) gram_pcg_kernel(double* __restrict__ parts, int nparts, int NB, int k, double sign,
                                                       double rtol, double* __restrict__ out, gnk_p2p_dev pd) {
  __shared__ double Gs[WMAXC * GLD];   // dense G of [sign A | y]
  __shared__ double xs[WMAXC], rs[WMAXC], ps[WMAXC], qs[WMAXC], zs[WMAXC], minv[WMAXC], bs[WMAXC];
  __shared__ double sh[CT / 32];
  const int c = k + 1;
  const int NE = nblocks(NB) * 64;
  if (pd.peers) {
    p2p_allreduce_long(pd, parts, NE);  // in place in this rank's scratch: the sum over all ranks, identical everywhere
    nparts = 1;
  }
  // fragment order -> dense symmetric matrix (ranks, if gathered by NCCL, added in rank order)
  for (int e = threadIdx.x; e < c * c; e += CT) {
    const int i = e / c, l = e - i * c;
    const int a = i < l ? i : l, b = i < l ? l : i;
    const int I = a >> 3, J = b >> 3;
    const int idx = blk_index(NB, I, J) * 64 + ((a & 7) * 4 + ((b & 7) >> 1)) * 2 + (b & 1);
    double s = parts[idx];
    for (int r = 1; r < nparts; ++r) s += parts[(int64_t)r * NE + idx];
    if ((i < k) != (l < k)) s *= sign;  // (sign A)^T y
    Gs[i * GLD + l] = s;
  }
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < k) {
    bs[tid] = Gs[tid * GLD + k];
    minv[tid] = 1.0 / Gs[tid * GLD + tid];
    xs[tid] = 0.0;
    rs[tid] = Gs[tid * GLD + k];
    ps[tid] = 0.0;
  }
  __syncthreads();
  const double bb = cta_sum(tid < k ? bs[tid] * bs[tid] : 0.0, sh);
  const double bn = sqrt(bb);
  int its = 0;
  if (bn != 0.0) {
    const double atol = rtol * bn;
    double rho_prev = 0.0;
    const int warp = tid >> 5, lane = tid & 31;
    for (int it = 0; it < 10 * k; ++it) {
      const double rr = cta_sum(tid < k ? rs[tid] * rs[tid] : 0.0, sh);
      if (sqrt(rr) < atol) break;
      if (tid < k) zs[tid] = minv[tid] * rs[tid];
      __syncthreads();
      const double rho = cta_sum(tid < k ? rs[tid] * zs[tid] : 0.0, sh);
      if (tid < k) ps[tid] = (it == 0) ? zs[tid] : zs[tid] + (rho / rho_prev) * ps[tid];
      __syncthreads();
      // q = G p: warp w owns rows w, w + 32; lanes run over the columns
      for (int i = warp; i < k; i += CT / 32) {
        double a = 0.0;
        for (int l = lane; l < k; l += 32) a = fma(Gs[i * GLD + l], ps[l], a);
        a = warp_sum(a);
        if (lane == 0) qs[i] = a;
      }
      __syncthreads();
      const double pq = cta_sum(tid < k ? ps[tid] * qs[tid] : 0.0, sh);
      const double alpha = rho / pq;
      if (tid < k) {
        xs[tid] = fma(alpha, ps[tid], xs[tid]);
        rs[tid] = fma(-alpha, qs[tid], rs[tid]);
      }
      __syncthreads();
      rho_prev = rho;
      ++its;
    }
  }
  // result block of gnk_tsqr_ls: d, ||A d||^2 = d^T G d, ||y - A d||^2, 0, ||d||^2, sqrt(diag G) in place of diag R,
  // and the CG iteration count behind it
  for (int i = (tid >> 5); i < k; i += CT / 32) {
    double a = 0.0;
    for (int l = (tid & 31); l < k; l += 32) a = fma(Gs[i * GLD + l], xs[l], a);
    a = warp_sum(a);
    if ((tid & 31) == 0) qs[i] = a;  // (G d)_i
  }
  __syncthreads();
  const double dGd = cta_sum(tid < k ? xs[tid] * qs[tid] : 0.0, sh);
  const double db = cta_sum(tid < k ? xs[tid] * bs[tid] : 0.0, sh);
  const double d2 = cta_sum(tid < k ? xs[tid] * xs[tid] : 0.0, sh);
  if (tid < k) {
    out[tid] = xs[tid];
    out[k + 4 + tid] = sqrt(Gs[tid * GLD + tid]);
  }
  if (tid == 0) {
    out[k] = dGd;
    out[k + 1] = Gs[k * GLD + k] - 2.0 * db + dGd;
    out[k + 2] = 0.0;
    out[k + 3] = d2;
    out[2 * k + 4] = (double)its;
  }
}

template <int NB>
int launch_gram_wide(gnk_ctx* ctx, const Panel& src, double* scratch, cudaStream_t st) {
  constexpr int64_t GRAN = 8 * WG;
  int64_t ctas = ctx->sm_count;
  if (ctas * GRAN > src.n_rows) ctas = ceil_div(src.n_rows, GRAN);
  if (ctas < 1) ctas = 1;
  const int64_t rows_per_cta = ceil_div(ceil_div(src.n_rows, ctas), GRAN) * GRAN;
  ctas = ceil_div(src.n_rows, rows_per_cta);
  if (ctas < 1) ctas = 1;
  constexpr int NE = nblocks(NB) * 64;
  // scratch: [this rank's Gram matrix NE | all ranks' (NCCL path) P2P_MAXR * NE | per-CTA partials]
  gram_wide_kernel<NB><<<(unsigned)ctas, WT, 0, st>>>(src, rows_per_cta, scratch + (int64_t)(1 + P2P_MAXR) * NE,
                                                    ctx->d_tickets + TK_GRAMW, scratch);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}


// ======================================================================================================================
// Wide panels on the QR path: gnk_tsqr_ls for 33..56 panel columns (krylow_restart up to 55 with
// projected_least_squares = "qr").  Same scheme as cholqr.cu's refinement form, with the pieces that do not fit its
// 32 x 32 thread layout re-done for a dense 56 x 56 system:
//   gram_wide_kernel<5..7>      G = [A | y]^T [A | y] in one sweep over the panel (above)
//   wide_factor1_kernel         (cross-rank sum,) R1 = chol(G) in shared memory, normal-equation solution d0, status
//   wide_refine_kernel<KC>      rho = y - A d0 and g = A^T rho in one sweep (the panel's second and last read)
//   wide_factor2_kernel         (cross-rank sum,) d = d0 + (R1^T R1)^{-1} g, the scalar block of gnk_tsqr_ls
// Refusals (pivot floor, pivot ratios below the refinement form's floor, correction not small) are reported with the same
// sentinel as cholqr.cu; the caller re-issues the solve with the Householder TSQR.
// ======================================================================================================================
constexpr int QT = 1024;                 // threads of the single-CTA factor kernels
constexpr int QPAD = 64;                 // [g (k values) | sum rho^2], padded
constexpr double Q_PIVOT_FLOOR = 1e-12;  // as cholqr.cu: reduced pivot / diagonal entry ~ 1 / cond^2 of the leading columns
constexpr double Q_REFINE_FLOOR = 1e-10;
constexpr double Q_REFINE_ACCEPT = 1e-5;
// scratch behind the Gram scratch in gnk_ctx::d_gramw (doubles)
constexpr int64_t WQ_R = 0;                                  // R1, WMAXC x GLD
constexpr int64_t WQ_D0 = WQ_R + WMAXC * GLD;                // normal-equation solution d0
constexpr int64_t WQ_B = WQ_D0 + WMAXC;                      // b = (sign A)^T y
constexpr int64_t WQ_GLOC = WQ_B + WMAXC;                    // this rank's [g | sum rho^2]
constexpr int64_t WQ_GALL = WQ_GLOC + QPAD;                  // all ranks' (NCCL path)
constexpr int64_t WQ_STATUS = WQ_GALL + P2P_MAXR * QPAD;     // int status word: 0 refine, 1 refused, 2 too ill-conditioned
constexpr int64_t WQ_TOTAL = WQ_STATUS + 8;
constexpr int64_t GRAMW_NE_MAX = nblocks(WMAXB) * 64;
constexpr int64_t GRAMW_DOUBLES = (1 + P2P_MAXR + 512) * GRAMW_NE_MAX;   // the Gram scratch proper (gnk_gram_cgls)

__device__ void wide_write_refusal(int k, double* out) {
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    out[i] = 0.0;
    out[k + 4 + i] = 1.0;
  }
  if (threadIdx.x == 0) {
    out[k] = 0.0;
    out[k + 1] = 0.0;
    out[k + 2] = -1.0;
    out[k + 3] = 0.0;
  }
}

// Back substitution R x = z with the upper triangular k x k block of Rs (row * GLD + column), executed by ONE warp:
// lane l owns rows l and l + 32.  Once x_j is known every row above subtracts R_ij x_j (two shuffles and an FMA per
// step; the reciprocals of the diagonal are formed up front).
__device__ __forceinline__ void warp_solve_upper(const double* Rs, int k, double (&z)[2], double (&x)[2]) {
  const int l = threadIdx.x & 31;
  double rinv[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int row = l + 32 * s;
    rinv[s] = row < k ? 1.0 / Rs[row * GLD + row] : 1.0;
    x[s] = 0.0;
  }
  for (int j = k - 1; j >= 0; --j) {
    const int src = j & 31;
    double xj;
    if (j >= 32) xj = __shfl_sync(0xffffffffu, z[1], src) * __shfl_sync(0xffffffffu, rinv[1], src);
    else xj = __shfl_sync(0xffffffffu, z[0], src) * __shfl_sync(0xffffffffu, rinv[0], src);
    if (l == src) {
      if (j >= 32) x[1] = xj;
      else x[0] = xj;
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = l + 32 * s;
      if (row < j) z[s] = fma(-Rs[row * GLD + j], xj, z[s]);
    }
  }
}
// Forward substitution R^T u = z (R upper triangular): once u_j is known every row below subtracts R_ji u_j.
__device__ __forceinline__ void warp_solve_upper_transposed(const double* Rs, int k, double (&z)[2], double (&u)[2]) {
  const int l = threadIdx.x & 31;
  double rinv[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int row = l + 32 * s;
    rinv[s] = row < k ? 1.0 / Rs[row * GLD + row] : 1.0;
    u[s] = 0.0;
  }
  for (int j = 0; j < k; ++j) {
    const int src = j & 31;
    double uj;
    if (j >= 32) uj = __shfl_sync(0xffffffffu, z[1], src) * __shfl_sync(0xffffffffu, rinv[1], src);
    else uj = __shfl_sync(0xffffffffu, z[0], src) * __shfl_sync(0xffffffffu, rinv[0], src);
    if (l == src) {
      if (j >= 32) u[1] = uj;
      else u[0] = uj;
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = l + 32 * s;
      if (row > j && row < k) z[s] = fma(-Rs[j * GLD + row], uj, z[s]);
    }
  }
}

__global__ void __launch_bounds__(QT) wide_factor1_kernel(double* __restrict__ parts, int nparts, int NB, int k, double sign,
                                                           double* __restrict__ Rg, double* __restrict__ d0g,
                                                           double* __restrict__ bg, int* __restrict__ status,
                                                           gnk_p2p_dev pd) {
  __shared__ double Gs[WMAXC * GLD];
  __shared__ double diag[WMAXC];
  const int c = k + 1, tid = threadIdx.x;
  const int NE = nblocks(NB) * 64;
  if (pd.peers) {
    // compute step + collective in one kernel, as in gram_pcg_kernel: the rank's Gram matrix is summed over the ranks in
    // rank order through the peer mailboxes (in place; bit-identical on every rank, so are all decisions below)
    p2p_allreduce_long(pd, parts, NE);
    nparts = 1;
  }
  // fragment order -> dense upper triangle (ranks, if gathered by NCCL, added in rank order); zero below the diagonal
  for (int e = tid; e < c * c; e += QT) {
    const int i = e / c, l = e - i * c;
    double s = 0.0;
    if (l >= i) {
      const int idx = blk_index(NB, i >> 3, l >> 3) * 64 + ((i & 7) * 4 + ((l & 7) >> 1)) * 2 + (l & 1);
      s = parts[idx];
      for (int r = 1; r < nparts; ++r) s += parts[(int64_t)r * NE + idx];
      if (l == k && i < k) s *= sign;  // (sign A)^T y
    }
    Gs[i * GLD + l] = s;
    if (i == l) diag[i] = s;
  }
  __syncthreads();
  if (tid < k) bg[tid] = Gs[tid * GLD + k];  // ||A d||^2 = b^T d at the solution (wide_factor2_kernel)
  __syncthreads();  // column k is updated from step 0 on
  // right-looking Cholesky G = R^T R on the upper triangle: step j subtracts a_ji a_jl / a_jj from the trailing block
  // (row j is final by then); the square roots and the divisions by them are taken after the loop
  bool ok = true;
  double min_ratio = 1.0;  // min over the A columns of reduced pivot / diagonal entry ~ 1 / cond^2 of [A] alone
  for (int j = 0; j < c; ++j) {
    const double ajj = Gs[j * GLD + j], dj = diag[j];
    if (!(ajj > Q_PIVOT_FLOOR * dj) || !(dj > 0.0)) {  // every thread reads the same two numbers
      ok = false;
      break;
    }
    if (j < c - 1) min_ratio = fmin(min_ratio, ajj / dj);
    const double inv = __drcp_rn(ajj);
    const int rem = c - 1 - j;
    for (int e = tid; e < rem * rem; e += QT) {
      const int ii = e / rem, ll = e - ii * rem;
      if (ll >= ii) {
        const int i = j + 1 + ii, l = j + 1 + ll;
        Gs[i * GLD + l] = fma(-Gs[j * GLD + i], Gs[j * GLD + l] * inv, Gs[i * GLD + l]);
      }
    }
    __syncthreads();
  }
  if (tid == 0) *status = ok ? (min_ratio >= Q_REFINE_FLOOR ? 0 : 2) : 1;
  if (!ok) return;  // uniform
  {
    double v[4];
    static_assert(4 * QT >= WMAXC * WMAXC, "four matrix entries per thread");
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * QT;
      v[q] = 0.0;
      if (e < c * c) {
        const int i = e / c, l = e - i * c;
        if (l >= i) {
          const double piv = sqrt(Gs[i * GLD + i]);
          v[q] = (l == i) ? piv : Gs[i * GLD + l] / piv;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * QT;
      if (e < c * c) {
        const int i = e / c, l = e - i * c;
        Gs[i * GLD + l] = v[q];
        Rg[i * GLD + l] = v[q];
      }
    }
    __syncthreads();
  }
  if (tid < 32) {
    // normal-equation solution R1[:k,:k] d0 = R1[:k,k]
    double z[2], d[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = tid + 32 * s;
      z[s] = row < k ? Gs[row * GLD + k] : 0.0;
    }
    warp_solve_upper(Gs, k, z, d);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = tid + 32 * s;
      if (row < k) d0g[row] = d[s];
    }
  }
}

// rho = y - (sign A) d0,  g = (sign A)^T rho,  sum rho^2: the panel's second read.  A thread owns a row pair and keeps
// the KC partial sums of g in registers; the columns pass through in chunks of CHK loads in flight, twice per row pair --
// first for rho (from HBM), then for g (the lines just loaded: L1 / L2 hits).
constexpr int WRT = 256;
template <int KC, int CHK>
__global__ void __launch_bounds__(WRT, 1)
    wide_refine_kernel(Panel src, double sign, const double* __restrict__ d0g, const int* __restrict__ status,
                       int64_t rows_per_cta, double* __restrict__ partials, unsigned int* ticket,
                       double* __restrict__ gout) {
  static_assert(KC % CHK == 0, "whole chunks");
  __shared__ double ds[KC];
  __shared__ double wsum[WRT / 32][KC + 1];
  if (*status != 0) return;  // refused or too ill-conditioned for the refinement form (uniform across the grid)
  const int k = src.k;
  for (int j = threadIdx.x; j < KC; j += WRT) ds[j] = (j < k) ? sign * d0g[j] : 0.0;
  __syncthreads();
  double g[KC];
#pragma unroll
  for (int j = 0; j < KC; ++j) g[j] = 0.0;
  double rr = 0.0;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t limit = row0 + rows_per_cta;
  if (limit > src.n_rows) limit = src.n_rows;
  for (int64_t r = row0 + 2 * threadIdx.x; r < limit; r += 2 * WRT) {
    double2 rho = __ldcs(reinterpret_cast<const double2*>(src.y + r));
#pragma unroll
    for (int j0 = 0; j0 < KC; j0 += CHK) {
      if (j0 < k) {  // uniform
        double2 a[CHK];
#pragma unroll
        for (int jj = 0; jj < CHK; ++jj)
          a[jj] = (j0 + jj < k) ? __ldg(reinterpret_cast<const double2*>(src.A + (int64_t)(j0 + jj) * src.lda + r))
                                : make_double2(0.0, 0.0);
#pragma unroll
        for (int jj = 0; jj < CHK; ++jj) {
          rho.x = fma(-ds[j0 + jj], a[jj].x, rho.x);
          rho.y = fma(-ds[j0 + jj], a[jj].y, rho.y);
        }
      }
    }
#pragma unroll
    for (int j0 = 0; j0 < KC; j0 += CHK) {
      if (j0 < k) {
        double2 a[CHK];
#pragma unroll
        for (int jj = 0; jj < CHK; ++jj)
          a[jj] = (j0 + jj < k) ? __ldcs(reinterpret_cast<const double2*>(src.A + (int64_t)(j0 + jj) * src.lda + r))
                                : make_double2(0.0, 0.0);
#pragma unroll
        for (int jj = 0; jj < CHK; ++jj) g[j0 + jj] = fma(a[jj].x, rho.x, fma(a[jj].y, rho.y, g[j0 + jj]));
      }
    }
    rr = fma(rho.x, rho.x, fma(rho.y, rho.y, rr));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < KC; ++j) {
    const double v = warp_sum(g[j]);
    if (lane == 0) wsum[warp][j] = v;
  }
  rr = warp_sum(rr);
  if (lane == 0) wsum[warp][KC] = rr;
  __syncthreads();
  double* mine = partials + (int64_t)blockIdx.x * (KC + 1);
  if (threadIdx.x <= KC) {
    double v = wsum[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < WRT / 32; ++w) v += wsum[w][threadIdx.x];
    mine[threadIdx.x] = v;
    __threadfence();
  }
  if (!grid_arrive_last(ticket)) return;
  if (threadIdx.x <= KC) {
    double v = 0.0;
    const int nb = gridDim.x;
    for (int b = 0; b < nb; ++b) v += __ldcg(partials + (int64_t)b * (KC + 1) + threadIdx.x);
    if (threadIdx.x < k) gout[threadIdx.x] = sign * v;   // (sign A)^T rho
    if (threadIdx.x == KC) gout[k] = v;                  // sum rho^2
  }
}

__global__ void __launch_bounds__(QT) wide_factor2_kernel(const double* __restrict__ parts, int nparts, int k,
                                                           const double* __restrict__ Rg,
                                                           const double* __restrict__ d0g,
                                                           const double* __restrict__ bg,
                                                           const int* __restrict__ status, double* __restrict__ out,
                                                           gnk_p2p_dev pd) {
  __shared__ double Rs[WMAXC * GLD];
  __shared__ double gs[QPAD];
  __shared__ double dsum[2];
  const int c = k + 1, tid = threadIdx.x;
  // the cross-rank sum of [g | sum rho^2] comes first and unconditionally: every collective of the channel must be
  // executed by every rank, whatever the status word says (a refused solve sums stale numbers nobody reads)
  if (tid < QPAD) {
    double g = 0.0;
    if (tid < c) {
      g = parts[tid];
      for (int r = 1; r < nparts; ++r) g += parts[(int64_t)r * QPAD + tid];
    }
    gs[tid] = g;
  }
  __syncthreads();
  if (pd.peers) {
    p2p_tail_allreduce(pd, gs, c, 0);
    __syncthreads();
  }
  if (*status != 0) {
    wide_write_refusal(k, out);
    return;
  }
  for (int e = tid; e < c * GLD; e += QT) Rs[e] = Rg[e];
  __syncthreads();
  if (tid < 32) {
    // delta = (R1^T R1)^{-1} g by two triangular solves, d = d0 + delta.  (sign^2 = 1: g = (sign A)^T rho already
    // carries the sign, the Gram block of the A columns does not.)
    double z[2], u[2], dl[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = tid + 32 * s;
      z[s] = row < k ? gs[row] : 0.0;
    }
    warp_solve_upper_transposed(Rs, k, z, u);
    warp_solve_upper(Rs, k, u, dl);
    double d[2], rii[2];
    double d2 = 0.0, dl2 = 0.0, ndef = 0.0, bd = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int row = tid + 32 * s;
      const bool in = row < k;
      d[s] = in ? d0g[row] + dl[s] : 0.0;
      rii[s] = in ? Rs[row * GLD + row] : 1.0;
      d2 = fma(d[s], d[s], d2);
      dl2 = in ? fma(dl[s], dl[s], dl2) : dl2;
      ndef += (in && fabs(rii[s]) <= 1e-8) ? 1.0 : 0.0;
      bd = in ? fma(bg[row], d[s], bd) : bd;
    }
    d2 = warp_sum(d2);
    dl2 = warp_sum(dl2);
    ndef = warp_sum(ndef);
    bd = warp_sum(bd);
    if (tid == 0) {
      dsum[0] = d2;
      dsum[1] = dl2;
    }
    if (dl2 <= Q_REFINE_ACCEPT * Q_REFINE_ACCEPT * d2) {  // uniform
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int row = tid + 32 * s;
        if (row < k) {
          out[row] = d[s];
          out[k + 4 + row] = rii[s];
        }
      }
      if (tid == 0) {
        out[k] = bd;          // ||A d||^2 = b^T d at the least-squares solution (see cholqr_factor2_kernel)
        out[k + 1] = gs[k];   // sum rho^2
        out[k + 2] = ndef;
        out[k + 3] = d2;
      }
    }
  }
  __syncthreads();
  if (!(dsum[1] <= Q_REFINE_ACCEPT * Q_REFINE_ACCEPT * dsum[0])) wide_write_refusal(k, out);
}

template <int KC, int CHK>
int launch_wide_refine(gnk_ctx* ctx, const Panel& src, double sign, double* partials, double* qr, cudaStream_t st) {
  constexpr int64_t GRAN = 2 * WRT;
  int64_t ctas = ctx->sm_count < 512 ? ctx->sm_count : 512;
  if (ctas * GRAN > src.n_rows) ctas = ceil_div(src.n_rows, GRAN);
  if (ctas < 1) ctas = 1;
  const int64_t rows_per_cta = ceil_div(ceil_div(src.n_rows, ctas), GRAN) * GRAN;
  ctas = ceil_div(src.n_rows, rows_per_cta);
  if (ctas < 1) ctas = 1;
  wide_refine_kernel<KC, CHK><<<(unsigned)ctas, WRT, 0, st>>>(src, sign, qr + WQ_D0, reinterpret_cast<int*>(qr + WQ_STATUS),
                                                           rows_per_cta, partials, ctx->d_tickets + TK_GRAMW,
                                                           qr + WQ_GLOC);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

int ensure_gramw(gnk_ctx* ctx, cudaStream_t st) {
  if (!ctx->d_gramw) {
    const size_t doubles = (size_t)(GRAMW_DOUBLES + WQ_TOTAL);
    GNK_CUDA(cudaMalloc(&ctx->d_gramw, sizeof(double) * doubles));
    GNK_CUDA(cudaMemsetAsync(ctx->d_gramw, 0, sizeof(double) * doubles, st));
  }
  return 0;
}

template <int NB>
int run_wide_qr(gnk_ctx* ctx, const Panel& src, double sign, double* d_out, cudaStream_t st) {
  if (int rc = ensure_gramw(ctx, st)) return rc;
  double* scratch = ctx->d_gramw;
  double* qr = scratch + GRAMW_DOUBLES;
  int* status = reinterpret_cast<int*>(qr + WQ_STATUS);
  constexpr int NE = nblocks(NB) * 64;
  double* partials = scratch + (int64_t)(1 + P2P_MAXR) * NE;
  if (int rc = launch_gram_wide<NB>(ctx, src, scratch, st)) return rc;
  const bool multi = ctx->nranks > 1;
  const gnk_p2p_dev none{nullptr, 0, 1, 0ull};
  const gnk_p2p_dev pd1 = multi ? p2p_next(ctx) : none;
  const bool gather1 = multi && !pd1.peers;
  if (gather1)
    if (int rc = gnk_comm_allgather_doubles(ctx, scratch, scratch + NE, NE, st)) return rc;
  wide_factor1_kernel<<<1, QT, 0, st>>>(gather1 ? scratch + NE : scratch, gather1 ? ctx->nranks : 1, NB, src.k, sign,
                                        qr + WQ_R, qr + WQ_D0, qr + WQ_B, status, pd1);
  GNK_LAUNCH_CHECK(ctx);
  int rc;
  if (src.k <= 42) rc = launch_wide_refine<42, 14>(ctx, src, sign, partials, qr, st);
  else rc = launch_wide_refine<56, 14>(ctx, src, sign, partials, qr, st);
  if (rc) return rc;
  const gnk_p2p_dev pd2 = multi ? p2p_next(ctx) : none;
  const bool gather2 = multi && !pd2.peers;
  if (gather2)
    if (int rc2 = gnk_comm_allgather_doubles(ctx, qr + WQ_GLOC, qr + WQ_GALL, QPAD, st)) return rc2;
  wide_factor2_kernel<<<1, QT, 0, st>>>(gather2 ? qr + WQ_GALL : qr + WQ_GLOC, gather2 ? ctx->nranks : 1, src.k, qr + WQ_R,
                                        qr + WQ_D0, qr + WQ_B, status, d_out, pd2);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace

// Called by gnk_tsqr_ls (tsqr.cu) for the panels it found eligible (33 <= k+1 <= 56 columns, >= 16384 rows, even row
// count and leading dimension, 16-byte aligned, sign = +-1); returns 1 for a panel it does not take.
int gnk_cholqr_wide_try(gnk_ctx* ctx, const double* d_A, int64_t lda, int64_t n_rows, int k, const double* d_y,
                        double sign_a, double* d_out, void* stream) {
  const int c = k + 1;
  if (c <= 32 || c > WMAXC || ctx->nranks > P2P_MAXR) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  Panel src{d_A, lda, d_y, k, n_rows};
  if (c <= 40) return run_wide_qr<5>(ctx, src, sign_a, d_out, st);
  if (c <= 48) return run_wide_qr<6>(ctx, src, sign_a, d_out, st);
  return run_wide_qr<7>(ctx, src, sign_a, d_out, st);
}

extern "C" int gnk_gram_cgls(gnk_ctx* ctx, const double* d_A, int64_t lda, int64_t n_rows, int k, const double* d_y,
                             double sign_a, double rtol, double* d_out, void* stream) {
  GNK_REQUIRE(ctx && d_A && d_y && d_out, "gnk_gram_cgls: null argument");
  GNK_REQUIRE(k >= 1 && k + 1 <= WMAXC, "gnk_gram_cgls: at most 55 columns");
  GNK_REQUIRE(n_rows >= 0 && lda >= n_rows && lda % 2 == 0 && n_rows % 2 == 0, "gnk_gram_cgls: even row count / lda");
  GNK_REQUIRE((uintptr_t)d_A % 16 == 0 && (uintptr_t)d_y % 16 == 0, "gnk_gram_cgls: 16-byte aligned panel");
  GNK_REQUIRE(sign_a == 1.0 || sign_a == -1.0, "gnk_gram_cgls: sign must be +-1");
  GNK_REQUIRE(ctx->nranks <= P2P_MAXR, "gnk_gram_cgls: too many ranks");
  cudaStream_t st = (cudaStream_t)stream;
  const int NB = (k + 1 + 7) / 8;
  if (int rc0 = ensure_gramw(ctx, st)) return rc0;
  double* scratch = ctx->d_gramw;
  Panel src{d_A, lda, d_y, k, n_rows};
  int rc = 0;
  switch (NB) {
    case 1: rc = launch_gram_wide<1>(ctx, src, scratch, st); break;
    case 2: rc = launch_gram_wide<2>(ctx, src, scratch, st); break;
    case 3: rc = launch_gram_wide<3>(ctx, src, scratch, st); break;
    case 4: rc = launch_gram_wide<4>(ctx, src, scratch, st); break;
    case 5: rc = launch_gram_wide<5>(ctx, src, scratch, st); break;
    case 6: rc = launch_gram_wide<6>(ctx, src, scratch, st); break;
    default: rc = launch_gram_wide<7>(ctx, src, scratch, st); break;
  }
  if (rc) return rc;
  const int NE = nblocks(NB) * 64;
  const bool multi = ctx->nranks > 1;
  gnk_p2p_dev pd{nullptr, 0, 1, 0ull};
  if (multi) pd = p2p_next(ctx);
  const bool gather = multi && !pd.peers;
  if (gather)
    if (int rc2 = gnk_comm_allgather_doubles(ctx, scratch, scratch + NE, NE, stream)) return rc2;
  gram_pcg_kernel<<<1, CT, 0, st>>>(gather ? scratch + NE : scratch, gather ? ctx->nranks : 1, NB, k, sign_a, rtol, d_out,
                                    pd);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}
