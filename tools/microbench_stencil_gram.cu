// microbench_stencil_gram.cu -- DRAFT for the next step of DESIGN.md section 7 (1): one kernel that applies the Bratu
// stencil to all k basis columns of a tile (J V_k, gauss_newton_krylow.py:86), stores J V_k, AND accumulates the Gram
// matrix of [J V_k | y] with DMMAs in the fragment layout of csrc/cholqr.cu pass 1 -- so that the least-squares panel is
// read once instead of twice (-8 n k bytes per outer iteration, ~10 ms of the 88 ms step at 4096^2).
//
// STATUS (end of round 1, profiles/r01s3_microbench_stencil_gram.txt): CORRECT -- J V is bitwise identical to the
// naive stencil and G agrees to 3e-16 at 256^2 / k = 15 and 512^2 / k = 31 -- but SLOW as laid out here: 4.72 ms at
// 4096^2, k = 30 (1.76 TB/s over the 8.3 GB it moves) against 1.32 + 0.85 ms for apply_kernel + cholqr_gram_kernel.
// Not profiled yet (the GPU budget ended with this run), but the sector arithmetic explains the factor: per warp and
// step the tile loads ask for 4 x 16 = 64 sectors and the stores for 64, while the sixteen 8-byte left/right
// neighbour loads touch ~2.5 sectors per column each = 160 sectors -- if they miss L1 (the tile rows are loaded with
// ld.global.cs, evict first, so the lines are probably gone when the neighbours are asked for) the kernel moves
// (64 + 160 + 64) / 128 = 2.3 x the L2 sectors it should, at a 32 KB stride (64 bytes per column and step; apply_kernel
// reads 4 KB per column and row).  Next: neighbours by quad shuffle (only lanes t = 0 / 3 load, 64 sectors instead of
// 160), wider segments, or -- the robust fix -- rows x 32-column tiles with a one-cell halo staged in shared memory as
// tsqr_stencil_kernel does (5 shared loads per output: 0.6 ms of shared-memory time per solve step at k = 30, under
// the 1.3 ms of HBM time).
// It is a stand-alone program (not part of libgnk_b200.so, not built by __graft_entry__.build()):
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tools/microbench_stencil_gram \
//        tools/microbench_stencil_gram.cu
//   tools/microbench_stencil_gram 256 15        # validation against naive kernels (J V bitwise, G to 1e-12)
//   tools/microbench_stencil_gram 4096 30       # timing: ms and GB/s over the 16 n k + 16 n bytes it moves
//   tools/microbench_stencil_gram 4096 30 2     # version 2 (neighbours by shuffle; compiled, not run yet)
//
// Design (numbers from profiles/r01s3_cholqr_ncu.txt and DESIGN.md section 3):
//  * a warp owns an 8-wide j-segment and marches down a strip of TR grid rows; lane (g, t) holds, for every column block
//    I, the two grid points j = 8 seg + 2t, 2t+1 of basis column 8I+g, i.e. exactly the pass-1 fragment (the DMMA k
//    index runs over the 4 lanes of a quad; .x and .y feed two DMMAs).  The (up, mid, down, down+1) rows of the tile
//    live in registers (4 x NB double2), so every basis element is loaded from HBM once; the left/right neighbours are
//    8-byte loads that hit L1 (the neighbouring lanes / warps loaded those lines);
//  * arithmetic of J V is apply_refbits of csrc/common.cuh (scipy's order, no FMA contraction): bit-identical to
//    apply_kernel, which the parity tests require;
//  * FP64 pipe per 8 grid points and 32 columns: 20 DMMAs + ~80 DFMA-rate instructions = 1.5 x pass 1 today, ~0.9 ms at
//    k = 30 -- under the 1.3 ms the 16 n k bytes take at the measured HBM peak; two rows are in flight per warp
//    (12 warps x 2 x 2 KB = 48 KB per SM; one row gives 3.5 TB/s);
//  * tasks (strip, segment) are numbered with the segment fastest and dealt to the warps round-robin, so that at any
//    time the GPU sweeps a few consecutive grid rows and the CTAs' 768-byte pieces are adjacent in every column.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);      \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int GW = 12, GT = 32 * GW;
constexpr int TR = 64;  // grid rows per strip

struct Problem {
  int m, rows, k;          // m x rows owned grid points (single rank: rows = m), k basis columns
  int64_t ldv, off, ldjv;  // stored column: [2 halo rows | rows | 2 halo rows] x m, off = 2 m
  double c_lap, c_adv, lam, sign;
};

__host__ __device__ constexpr int nblocks(int NB) { return NB * (NB + 1) / 2; }
__host__ __device__ constexpr int blk_index(int NB, int I, int J) { return I * NB - I * (I - 1) / 2 + (J - I); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// same as csrc/common.cuh
__device__ __forceinline__ double apply_refbits(double cu, double cl, double dg, double cd, double up, double lf,
                                                double mid, double rt, double dn) {
  double s = __dadd_rn(__dmul_rn(cu, up), __dmul_rn(cl, lf));
  s = __dadd_rn(s, __dmul_rn(dg, mid));
  s = __dadd_rn(s, __dmul_rn(cl, rt));
  s = __dadd_rn(s, __dmul_rn(cd, dn));
  return s;
}

// ---- naive reference kernels (validation only) ---------------------------------------------------------------------
__global__ void ref_apply(Problem p, const double* V, const double* expu, double* JV) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)p.rows * p.m;
  if (idx >= n * p.k) return;
  const int col = (int)(idx / n);
  const int64_t r = idx - (int64_t)col * n;
  const int i = (int)(r / p.m), j = (int)(r - (int64_t)i * p.m);
  const double* v = V + (int64_t)col * p.ldv + p.off;
  const double d0 = __dadd_rn(4.0 * p.c_lap, -p.c_adv), cd = __dadd_rn(-p.c_lap, p.c_adv), cu = -p.c_lap, cl = -p.c_lap;
  const double dg = __dadd_rn(d0, __dmul_rn(p.lam, expu[p.off + r]));
  const double lf = j > 0 ? v[r - 1] : 0.0, rt = j + 1 < p.m ? v[r + 1] : 0.0;
  JV[(int64_t)col * p.ldjv + r] = p.sign * apply_refbits(cu, cl, dg, cd, v[r - p.m], lf, v[r], rt, v[r + p.m]);
}
__global__ void ref_gram(Problem p, const double* JV, const double* y, double* G, int c) {
  // one CTA per (a, b) entry
  const int a = blockIdx.x / c, b = blockIdx.x % c;
  const int64_t n = (int64_t)p.rows * p.m;
  const double* pa = a < p.k ? JV + (int64_t)a * p.ldjv : y + p.off;
  const double* pb = b < p.k ? JV + (int64_t)b * p.ldjv : y + p.off;
  double s = 0.0;
  for (int64_t r = threadIdx.x; r < n; r += blockDim.x) s = fma(pa[r], pb[r], s);
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) G[a * c + b] = sh[0];
}

// ---- the fused kernel ------------------------------------------------------------------------------------------------
template <int NB>
struct Row {
  double2 v[NB];
};

// SHFL = false: version 1 (the measured one): evict-first tile loads, all neighbours by 8-byte global loads.
// SHFL = true : version 2 (compiles, NOT run yet): default-cached tile loads, left/right neighbours from the adjacent
//               lane of the quad by shuffle; only lanes t = 0 / t = 3 load the element of the neighbouring segment.
template <int NB, bool SHFL>
__device__ __forceinline__ void load_row(Row<NB>& R, const double* const (&vb)[NB], int64_t ro) {
#pragma unroll
  for (int I = 0; I < NB; ++I) {
    if (!vb[I])
      R.v[I] = make_double2(0.0, 0.0);
    else if (SHFL)
      R.v[I] = __ldg(reinterpret_cast<const double2*>(vb[I] + ro));
    else
      R.v[I] = __ldcs(reinterpret_cast<const double2*>(vb[I] + ro));
  }
}

template <int NB, bool SHFL>
__global__ void __launch_bounds__(GT, 1)
    stencil_gram_kernel(Problem p, const double* __restrict__ V, const double* __restrict__ expu,
                        const double* __restrict__ y, double* __restrict__ JV, double* __restrict__ partials) {
  constexpr int NBLK = nblocks(NB);
  __shared__ __align__(16) double red[NBLK * 64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m = p.m;
  const int nseg = m / 8;  // requires m % 8 == 0 (the library would fall back to the two kernels otherwise)
  const int nstrip = (p.rows + TR - 1) / TR;
  const int64_t ntask = (int64_t)nseg * nstrip;
  const int64_t gw = (int64_t)blockIdx.x * GW + warp, nw = (int64_t)gridDim.x * GW;
  const double d0 = __dadd_rn(4.0 * p.c_lap, -p.c_adv), cd = __dadd_rn(-p.c_lap, p.c_adv), cu = -p.c_lap, cl = -p.c_lap;

  double acc[NBLK][2];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) acc[b][0] = acc[b][1] = 0.0;

  for (int64_t task = gw; task < ntask; task += nw) {
    const int strip = (int)(task / nseg), seg = (int)(task - (int64_t)strip * nseg);
    const int j = 8 * seg + 2 * t;
    const int i0 = strip * TR, i1 = min(i0 + TR, p.rows);
    const bool has_l = j > 0, has_r = j + 2 < m;
    // per column block: stencil columns read V, column k reads y (no stencil), columns beyond k are zero
    const double* vb[NB];
    double* ob[NB];
    bool sten[NB];
#pragma unroll
    for (int I = 0; I < NB; ++I) {
      const int col = 8 * I + g;
      sten[I] = col < p.k;
      vb[I] = col < p.k ? V + (int64_t)col * p.ldv + p.off + j : (col == p.k ? y + p.off + j : nullptr);
      ob[I] = col < p.k ? JV + (int64_t)col * p.ldjv + j : nullptr;
    }
    const double* eb = expu + p.off + j;
    Row<NB> up, mid, dn, dn2;
    load_row<NB, SHFL>(up, vb, (int64_t)(i0 - 1) * m);  // halo rows exist (zero at the domain boundary)
    load_row<NB, SHFL>(mid, vb, (int64_t)i0 * m);
    load_row<NB, SHFL>(dn, vb, (int64_t)(i0 + 1) * m);
    for (int i = i0; i < i1; ++i) {
      const int64_t ro = (int64_t)i * m;
      load_row<NB, SHFL>(dn2, vb, ro + 2 * m);  // two rows ahead; the halo has two rows, so i + 2 <= rows + 1 is always stored
      const double2 e = __ldg(reinterpret_cast<const double2*>(eb + ro));
      const double dga = __dadd_rn(d0, __dmul_rn(p.lam, e.x)), dgb = __dadd_rn(d0, __dmul_rn(p.lam, e.y));
      double2 tile[NB];
#pragma unroll
      for (int I = 0; I < NB; ++I) {
        double lf_q = 0.0, rt_q = 0.0;
        if (SHFL) {  // all lanes take part; lane - 1 / lane + 1 is the same column, j - 2 / j + 2, for t > 0 / t < 3
          lf_q = __shfl_up_sync(0xffffffffu, mid.v[I].y, 1);
          rt_q = __shfl_down_sync(0xffffffffu, mid.v[I].x, 1);
        }
        if (sten[I]) {
          double lf, rt;
          if (SHFL) {
            lf = (t > 0) ? lf_q : (has_l ? vb[I][ro - 1] : 0.0);
            rt = (t < 3) ? rt_q : (has_r ? vb[I][ro + 2] : 0.0);
          } else {
            lf = has_l ? vb[I][ro - 1] : 0.0;
            rt = has_r ? vb[I][ro + 2] : 0.0;
          }
          const double oa = apply_refbits(cu, cl, dga, cd, up.v[I].x, lf, mid.v[I].x, mid.v[I].y, dn.v[I].x);
          const double obv = apply_refbits(cu, cl, dgb, cd, up.v[I].y, mid.v[I].x, mid.v[I].y, rt, dn.v[I].y);
          tile[I] = make_double2(p.sign * oa, p.sign * obv);
          __stcs(reinterpret_cast<double2*>(ob[I] + ro), tile[I]);
        } else {
          tile[I] = mid.v[I];  // the y column, or zeros
        }
      }
#pragma unroll
      for (int I = 0; I < NB; ++I)
#pragma unroll
        for (int J = I; J < NB; ++J) {
          const int b = blk_index(NB, I, J);
          dmma(acc[b][0], acc[b][1], tile[I].x, tile[J].x);
          dmma(acc[b][0], acc[b][1], tile[I].y, tile[J].y);
        }
      up = mid;
      mid = dn;
      dn = dn2;
    }
  }
  // CTA sum in warp order, then one partial per CTA (fragment order, as cholqr.cu: block * 64 + lane * 2 + {0, 1})
  for (int w = 0; w < GW; ++w) {
    if (warp == w) {
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        double2* q = reinterpret_cast<double2*>(red + b * 64 + lane * 2);
        double2 v = make_double2(acc[b][0], acc[b][1]);
        if (w > 0) {
          v.x += q->x;
          v.y += q->y;
        }
        *q = v;
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NBLK * 64; e += GT) partials[(int64_t)blockIdx.x * NBLK * 64 + e] = red[e];
}


// ---- version 3: TMA-staged row slots (tensor maps), warp-specialised ----------------------------------------------------
// One producer warp (lane 0) streams grid rows of a (TI x 64) task through a ring of shared-memory slots with
// cp.async.bulk.tensor: one 3-D box {72 j, 1 row, 8 NB columns} of V, one 2-D box each of y and e^u, all completing on
// the slot's `full` mbarrier.  Out-of-range j (j0 - 4 < 0, j0 + 68 > m) is zero-filled by the TMA unit = the Dirichlet
// condition.  Eight consumer warps each own an 8-point j-segment, keep the (up, mid, dn) window of their fragment in
// registers, take left/right neighbours and e^u from the mid row's slot, apply the stencil (apply_refbits, bit-identical
// to apply_kernel), store J V and feed the Gram DMMAs.  A slot is released (`empty`, one arrival per consumer warp)
// when its row has been the mid row.  Plane stride 576 B = 64 mod 128: the quarter-warp's two 64-byte pieces of an
// LDS.128 fall into disjoint banks.
namespace v3 {
// CW consumer warps -> TJ = 8 CW grid points per row slot; plane = TJ + 8 doubles (j0 - 4 .. j0 + TJ + 3); for CW = 8,
// 12, 16 the plane stride is 576 / 832 / 1088 bytes = 64 mod 128
template <int CW> __host__ __device__ constexpr int plane_bytes() { return (8 * CW + 8) * 8; }
template <int CW> __host__ __device__ constexpr int ye_bytes() { return (plane_bytes<CW>() + 127) / 128 * 128; }
template <int NB, int CW> __host__ __device__ constexpr int slot_bytes() { return 8 * NB * plane_bytes<CW>() + 2 * ye_bytes<CW>(); }
template <int NB, int CW> __host__ __device__ constexpr int nslot() {
  return (200 * 1024 / slot_bytes<NB, CW>()) < 16 ? (200 * 1024 / slot_bytes<NB, CW>()) : 16;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ double2 lds2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds1(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

template <int NB, bool STORE, int CW>
__global__ void __launch_bounds__(32 * (CW + 1), 1)
    stencil_gram_tma(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmE, Problem p, int TI, double* __restrict__ JV,
                     double* __restrict__ partials) {
  constexpr int NBLK = nblocks(NB);
  constexpr int SB = slot_bytes<NB, CW>();
  constexpr int NS = nslot<NB, CW>();
  constexpr int TJ = 8 * CW, PB = plane_bytes<CW>(), YE = ye_bytes<CW>();
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(16) double red[NBLK * 64];
  __shared__ __align__(8) unsigned long long bars[2 * NS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m = p.m;
  const int ntj = (m + TJ - 1) / TJ;
  const int nstrip = (p.rows + TI - 1) / TI;
  const int64_t ntask = (int64_t)ntj * nstrip;
  const uint32_t ring0 = smem_u32(ring), full0 = smem_u32(bars), empty0 = full0 + 8 * NS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc[NBLK][2];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) acc[b][0] = acc[b][1] = 0.0;

  if (warp == CW) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t task = blockIdx.x; task < ntask; task += gridDim.x) {
        const int strip = (int)(task / ntj), tj = (int)(task - (int64_t)strip * ntj);
        const int i0 = strip * TI, i1 = min(i0 + TI, p.rows), j0 = tj * TJ;
        for (int r = i0 - 1; r <= i1; ++r, ++it) {
          const uint32_t s = it % NS, ph = (it / NS) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const uint32_t dst = ring0 + s * SB, fb = full0 + 8 * s;
          mbar_expect_tx(fb, (8 * NB + 2) * PB);
          tma_load_3d(dst, &tmV, j0 - 4, r + 2, 0, fb);
          tma_load_2d(dst + 8 * NB * PB, &tmY, j0 - 4, r + 2, fb);
          tma_load_2d(dst + 8 * NB * PB + YE, &tmE, j0 - 4, r + 2, fb);
        }
      }
    }
  } else {
    const double d0 = __dadd_rn(4.0 * p.c_lap, -p.c_adv), cd = __dadd_rn(-p.c_lap, p.c_adv), cu = -p.c_lap, cl = -p.c_lap;
    // per column block: byte offset of this lane's pair inside a slot, and what the column is
    uint32_t poff[NB];
    int kind[NB];  // 0: stencil column, 1: the y column (copied), 2: beyond the panel (zeros)
#pragma unroll
    for (int I = 0; I < NB; ++I) {
      const int col = 8 * I + g;
      kind[I] = col < p.k ? 0 : (col == p.k ? 1 : 2);
      poff[I] = (uint32_t)((col < p.k ? col * PB : (col == p.k ? 8 * NB * PB : 0)) + 8 * (4 + 8 * warp + 2 * t));
    }
    const uint32_t eoff = (uint32_t)(8 * NB * PB + YE + 8 * (4 + 8 * warp + 2 * t));
    uint32_t it = 0;
    for (int64_t task = blockIdx.x; task < ntask; task += gridDim.x) {
      const int strip = (int)(task / ntj), tj = (int)(task - (int64_t)strip * ntj);
      const int i0 = strip * TI, i1 = min(i0 + TI, p.rows);
      const int j = tj * TJ + 8 * warp + 2 * t;
      const bool active = tj * TJ + 8 * warp < m;  // the last tile of a row may be narrower than TJ (whole segments)
      double2 up[NB], mid[NB], dn[NB];
      {  // row i0 - 1
        const uint32_t s = it % NS, ph = (it / NS) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        const uint32_t base = ring0 + s * SB;
#pragma unroll
        for (int I = 0; I < NB; ++I) up[I] = lds2(base + poff[I]);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        ++it;
      }
      uint32_t mid_base, mid_slot;
      {  // row i0
        const uint32_t s = it % NS, ph = (it / NS) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        mid_base = ring0 + s * SB;
        mid_slot = s;
#pragma unroll
        for (int I = 0; I < NB; ++I) mid[I] = lds2(mid_base + poff[I]);
        ++it;
      }
      for (int i = i0; i < i1; ++i) {
        const uint32_t s = it % NS, ph = (it / NS) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        const uint32_t base = ring0 + s * SB;
#pragma unroll
        for (int I = 0; I < NB; ++I) dn[I] = lds2(base + poff[I]);
        ++it;
        // neighbours and e^u of the mid row, then give its slot back
        double lf[NB], rt[NB];
#pragma unroll
        for (int I = 0; I < NB; ++I) {
          lf[I] = lds1(mid_base + poff[I] - 8);
          rt[I] = lds1(mid_base + poff[I] + 16);
        }
        const double2 e = lds2(mid_base + eoff);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * mid_slot);
        const double dga = __dadd_rn(d0, __dmul_rn(p.lam, e.x)), dgb = __dadd_rn(d0, __dmul_rn(p.lam, e.y));
        const int64_t ro = (int64_t)i * m + j;
        double2 tile[NB];
#pragma unroll
        for (int I = 0; I < NB; ++I) {
          if (kind[I] == 0 && active) {
            const double oa = apply_refbits(cu, cl, dga, cd, up[I].x, lf[I], mid[I].x, mid[I].y, dn[I].x);
            const double obv = apply_refbits(cu, cl, dgb, cd, up[I].y, mid[I].x, mid[I].y, rt[I], dn[I].y);
            tile[I] = make_double2(p.sign * oa, p.sign * obv);
            if (STORE) __stcs(reinterpret_cast<double2*>(JV + (int64_t)(8 * I + g) * p.ldjv + ro), tile[I]);
          } else if (kind[I] == 1 && active) {
            tile[I] = mid[I];
          } else {
            tile[I] = make_double2(0.0, 0.0);
          }
        }
#pragma unroll
        for (int I = 0; I < NB; ++I)
#pragma unroll
          for (int J = I; J < NB; ++J) {
            const int b = blk_index(NB, I, J);
            dmma(acc[b][0], acc[b][1], tile[I].x, tile[J].x);
            dmma(acc[b][0], acc[b][1], tile[I].y, tile[J].y);
          }
#pragma unroll
        for (int I = 0; I < NB; ++I) {
          up[I] = mid[I];
          mid[I] = dn[I];
        }
        mid_base = base;
        mid_slot = s;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * mid_slot);  // row i1 was only ever a dn row
    }
  }
  // CTA sum in warp order (consumer warps), one partial per CTA in fragment order
  for (int w = 0; w < CW; ++w) {
    if (warp == w) {
#pragma unroll
      for (int b = 0; b < NBLK; ++b) {
        double2* q = reinterpret_cast<double2*>(red + b * 64 + lane * 2);
        double2 v = make_double2(acc[b][0], acc[b][1]);
        if (w > 0) {
          v.x += q->x;
          v.y += q->y;
        }
        *q = v;
      }
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NBLK * 64; e += blockDim.x) partials[(int64_t)blockIdx.x * NBLK * 64 + e] = red[e];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    if (!f || q != cudaDriverEntryPointSuccess) {
      printf("cuTensorMapEncodeTiled not available\n");
      exit(1);
    }
    fn = (EncodeFn)f;
  }
  return fn;
}
// stored columns: [rows + 4][m] doubles, column stride ld; box {PW, 1, ncol_box}
static CUtensorMap make_map(const double* base, int m, int rows, int64_t ld, int ncols, int ncol_box, int PW) {
  CUtensorMap tm;
  if (ncols > 1 || ncol_box > 1) {
    cuuint64_t dims[3] = {(cuuint64_t)m, (cuuint64_t)(rows + 4), (cuuint64_t)ncols};
    cuuint64_t strides[2] = {(cuuint64_t)m * 8, (cuuint64_t)ld * 8};
    cuuint32_t box[3] = {(cuuint32_t)PW, 1, (cuuint32_t)ncol_box};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(3d) failed: %d\n", (int)r); exit(1); }
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)m, (cuuint64_t)(rows + 4)};
    cuuint64_t strides[1] = {(cuuint64_t)m * 8};
    cuuint32_t box[2] = {(cuuint32_t)PW, 1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled(2d) failed: %d\n", (int)r); exit(1); }
  }
  return tm;
}
}  // namespace v3

// fragment order -> dense c x c (upper part), CTA partials added in CTA order
__global__ void gather_kernel(const double* partials, int nctas, int NB, int c, double* G) {
  const int i = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (!(i <= l && l < c)) return;
  const int I = i >> 3, J = l >> 3;
  const int b = I * NB - I * (I - 1) / 2 + (J - I);
  const int e = b * 64 + ((i & 7) * 4 + ((l & 7) >> 1)) * 2 + (l & 1);
  const int NE = nblocks(NB) * 64;
  double s = 0.0;
  for (int q = 0; q < nctas; ++q) s += partials[(int64_t)q * NE + e];
  G[i * c + l] = s;
  G[l * c + i] = s;
}

template <int NB, bool SHFL>
float run_fused(const Problem& p, const double* V, const double* expu, const double* y, double* JV, double* partials,
                double* G, int ctas, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaEventRecord(e0));
    stencil_gram_kernel<NB, SHFL><<<ctas, GT>>>(p, V, expu, y, JV, partials);
    gather_kernel<<<1, 1024>>>(partials, ctas, NB, p.k + 1, G);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2 && ms < best) best = ms;
  }
  return best;
}

template <int NB, bool STORE, int CW>
float run_tma(const Problem& p, const double* V, const double* expu, const double* y, double* JV, double* partials,
              double* G, int ctas, int reps, int TI) {
  using namespace v3;
  constexpr int PW = 8 * CW + 8;
  const CUtensorMap tmV = make_map(V, p.m, p.rows, p.ldv, p.k, 8 * NB, PW);
  const CUtensorMap tmY = make_map(y, p.m, p.rows, p.ldv, 1, 1, PW);
  const CUtensorMap tmE = make_map(expu, p.m, p.rows, p.ldv, 1, 1, PW);
  auto kern = stencil_gram_tma<NB, STORE, CW>;
  const int dyn = slot_bytes<NB, CW>() * nslot<NB, CW>();
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 2; ++r) {
    CK(cudaEventRecord(e0));
    kern<<<ctas, 32 * (CW + 1), dyn>>>(tmV, tmY, tmE, p, TI, JV, partials);
    gather_kernel<<<1, 1024>>>(partials, ctas, NB, p.k + 1, G);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2 && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 256;
  const int k = argc > 2 ? atoi(argv[2]) : 15;
  if (m % 8 || k < 1 || k > 31) {
    printf("usage: %s m(multiple of 8) k(1..31)\n", argv[0]);
    return 1;
  }
  Problem p;
  p.m = m;
  p.rows = m;
  p.k = k;
  p.off = 2 * (int64_t)m;
  p.ldv = ((int64_t)(m + 4) * m + 15) / 16 * 16;
  p.ldjv = ((int64_t)m * m + 15) / 16 * 16;
  const double h = 6.0 / (m + 1);
  p.c_lap = 1.0 / (h * h);
  p.c_adv = 5.0 / h;
  p.lam = 10.0;
  p.sign = -1.0;
  const int c = k + 1, NB = (c + 7) / 8;
  const int64_t n = (int64_t)m * m;

  std::vector<double> hV((size_t)k * p.ldv, 0.0), hE(p.ldv, 0.0), hY(p.ldv, 0.0);
  srand(1);
  auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * 2e-3; };
  for (int col = 0; col < k; ++col)
    for (int64_t r = 0; r < n; ++r) hV[(size_t)col * p.ldv + p.off + r] = rnd();  // halo rows stay zero (Dirichlet)
  for (int64_t r = 0; r < n; ++r) {
    hE[p.off + r] = exp(rnd());
    hY[p.off + r] = rnd() * 1e3;
  }
  double *V, *E, *Y, *JV, *JVr, *part, *G, *Gr;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int ctas = prop.multiProcessorCount;
  CK(cudaMalloc(&V, sizeof(double) * hV.size()));
  CK(cudaMalloc(&E, sizeof(double) * hE.size()));
  CK(cudaMalloc(&Y, sizeof(double) * hY.size()));
  CK(cudaMalloc(&JV, sizeof(double) * (size_t)k * p.ldjv));
  CK(cudaMalloc(&part, sizeof(double) * (size_t)ctas * 640));
  CK(cudaMalloc(&G, sizeof(double) * 1024));
  CK(cudaMemcpy(V, hV.data(), sizeof(double) * hV.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(E, hE.data(), sizeof(double) * hE.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(Y, hY.data(), sizeof(double) * hY.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(JV, 0, sizeof(double) * (size_t)k * p.ldjv));
  CK(cudaMemset(G, 0, sizeof(double) * 1024));

  const int variant = argc > 3 ? atoi(argv[3]) : 1;  // 1 = version 1 (measured), 2 = version 2 (shuffle neighbours)
  float ms = 0;
  const int TI = argc > 4 ? atoi(argv[4]) : 64;
  if (variant == 3 || variant == 4) {  // 3: TMA-staged, stores J V; 4: the same without the store (Gram only)
    if (m % 8) {
      printf("version 3 needs m %% 8 == 0\n");
      return 1;
    }
    const int cw = argc > 5 ? atoi(argv[5]) : 8;
    printf("consumer warps %d: ", cw);
#define RUN_TMA(NBV, ST)                                                                  \
  if (NB == NBV) {                                                                        \
    if (cw == 12) ms = run_tma<NBV, ST, 12>(p, V, E, Y, JV, part, G, ctas, 5, TI);        \
    else if (cw == 10) ms = run_tma<NBV, ST, 10>(p, V, E, Y, JV, part, G, ctas, 5, TI);   \
    else if (cw == 16) ms = run_tma<NBV, ST, 16>(p, V, E, Y, JV, part, G, ctas, 5, TI);   \
    else ms = run_tma<NBV, ST, 8>(p, V, E, Y, JV, part, G, ctas, 5, TI);                  \
  }
    if (variant == 3) {
      RUN_TMA(1, true) RUN_TMA(2, true) RUN_TMA(3, true) RUN_TMA(4, true)
    } else {
      RUN_TMA(1, false) RUN_TMA(2, false) RUN_TMA(3, false) RUN_TMA(4, false)
    }
  } else if (variant == 2) {
    if (NB == 1) ms = run_fused<1, true>(p, V, E, Y, JV, part, G, ctas, 5);
    if (NB == 2) ms = run_fused<2, true>(p, V, E, Y, JV, part, G, ctas, 5);
    if (NB == 3) ms = run_fused<3, true>(p, V, E, Y, JV, part, G, ctas, 5);
    if (NB == 4) ms = run_fused<4, true>(p, V, E, Y, JV, part, G, ctas, 5);
  } else {
    if (NB == 1) ms = run_fused<1, false>(p, V, E, Y, JV, part, G, ctas, 5);
    if (NB == 2) ms = run_fused<2, false>(p, V, E, Y, JV, part, G, ctas, 5);
    if (NB == 3) ms = run_fused<3, false>(p, V, E, Y, JV, part, G, ctas, 5);
    if (NB == 4) ms = run_fused<4, false>(p, V, E, Y, JV, part, G, ctas, 5);
  }
  const double bytes = 16.0 * n * k + 16.0 * n;  // read V, write J V, read e^u and y
  printf("fused stencil + Gram v%d  m=%d k=%d TI=%d  %.4f ms  %.1f GB/s over %.3f GB (16nk+16n)\n", variant, m, k, TI, ms, bytes / ms / 1e6, bytes / 1e9);

  if (m <= 1024) {  // validation against the naive kernels
    CK(cudaMalloc(&JVr, sizeof(double) * (size_t)k * p.ldjv));
    CK(cudaMalloc(&Gr, sizeof(double) * 1024));
    CK(cudaMemset(JVr, 0, sizeof(double) * (size_t)k * p.ldjv));
    ref_apply<<<(unsigned)((n * k + 255) / 256), 256>>>(p, V, E, JVr);
    ref_gram<<<c * c, 256>>>(p, JVr, Y, Gr, c);
    CK(cudaDeviceSynchronize());
    std::vector<double> a((size_t)k * p.ldjv), b((size_t)k * p.ldjv), ga(1024), gb(1024);
    CK(cudaMemcpy(a.data(), JV, sizeof(double) * a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), JVr, sizeof(double) * b.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ga.data(), G, sizeof(double) * 1024, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gb.data(), Gr, sizeof(double) * 1024, cudaMemcpyDeviceToHost));
    const bool same = memcmp(a.data(), b.data(), sizeof(double) * a.size()) == 0;
    double worst = 0.0;
    for (int i = 0; i < c; ++i)
      for (int l = 0; l < c; ++l) {
        const double den = sqrt(fabs(gb[i * c + i] * gb[l * c + l])) + 1e-300;
        worst = fmax(worst, fabs(ga[i * c + l] - gb[i * c + l]) / den);
      }
    printf("J V bitwise identical to the naive stencil: %s;  max |G - G_ref| / sqrt(G_ii G_ll) = %.3e (expect < 1e-12)\n",
           same ? "yes" : "NO", worst);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
