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

}  // namespace

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
  const int64_t ne_max = nblocks(WMAXB) * 64;
  if (!ctx->d_gramw) {
    const size_t doubles = (size_t)(1 + P2P_MAXR + 512) * ne_max;
    GNK_CUDA(cudaMalloc(&ctx->d_gramw, sizeof(double) * doubles));
    GNK_CUDA(cudaMemsetAsync(ctx->d_gramw, 0, sizeof(double) * doubles, st));
  }
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
