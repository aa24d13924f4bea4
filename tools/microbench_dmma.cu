// microbench_dmma.cu -- latency/throughput of mma.sync.m8n8k4.f64 on this device, and a check that two chained
// DMMAs with an all-ones A operand reproduce a 32-lane sum.  Development aid.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ double warp_sum_dmma(double v) {
  double d0, d1, e0, e1;
  dmma(d0, d1, 1.0, v, 0.0, 0.0);        // d = sums over groups of 4 lanes; lane l holds groups 2(l%4), 2(l%4)+1
  dmma(e0, e1, 1.0, d0 + d1, 0.0, 0.0);  // sum of the 8 group sums, in every lane
  return e0;
}
__global__ void lat(double* out, long long* clk) {
  double x = out[threadIdx.x], d0 = 0, d1 = 0;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { dmma(d0, d1, 1.0, x, d0, d1); x = d0; }
  long long t1 = clock64();
  out[threadIdx.x] = x + d1;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int ILP>
__global__ void thr(double* out, long long* clk) {
  double d0[ILP], d1[ILP];
  double x = out[threadIdx.x];
  for (int q = 0; q < ILP; ++q) { d0[q] = q; d1[q] = -q; }
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) dmma(d0[q], d1[q], 1.0, x, d0[q], d1[q]);
  }
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < ILP; ++q) s += d0[q] + d1[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_sum(double* out, long long* clk) {
  double x = out[threadIdx.x] + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = warp_sum_dmma(x) * 1e-3 + threadIdx.x;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void check(double* out) {
  double v = 1.0 + threadIdx.x * 0.37 + (threadIdx.x % 3) * 1e-9;
  double s = warp_sum_dmma(v);
  double r = v;
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  out[threadIdx.x] = s;
  out[32 + threadIdx.x] = r;
}
// quad all-reduce: data in the A operand (lane l <-> A[l/4][l%4]), B all ones -> D[m][n] = sum of quad m, in every lane of the quad
__global__ void check_quad(double* out) {
  double v = 1.0 + threadIdx.x * 0.37;
  double d0, d1;
  dmma(d0, d1, v, 1.0, 0.0, 0.0);
  double r = v;
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  out[threadIdx.x] = d0;
  out[32 + threadIdx.x] = d1;
  out[64 + threadIdx.x] = r;
}
// DFMA and DMMA interleaved: do they share one pipe?
template <int NF, int NM>
__global__ void thr_mix(double* out, long long* clk) {
  double d0[4], d1[4], f[8];
  double x = out[threadIdx.x];
  for (int q = 0; q < 4; ++q) { d0[q] = q; d1[q] = -q; }
  for (int q = 0; q < 8; ++q) f[q] = q + x;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int q = 0; q < NM; ++q) dmma(d0[q & 3], d1[q & 3], 1.0, x, d0[q & 3], d1[q & 3]);
#pragma unroll
    for (int q = 0; q < NF; ++q) f[q & 7] = fma(f[q & 7], 1.0000001, 1e-9);
  }
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 4; ++q) s += d0[q] + d1[q];
  for (int q = 0; q < 8; ++q) s += f[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24); cudaMalloc(&clk, 8 * 4096);
  long long h[4096]; double hv[64];
  lat<<<1, 32>>>(out, clk); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("DMMA m8n8k4 dependent latency        %8.2f clk\n", (double)h[0] / N);
  lat_sum<<<1, 32>>>(out, clk); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("warp_sum via 2 DMMA (+mul,add) chain %8.2f clk\n", (double)h[0] / N);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int warps = 4; warps <= 32; warps *= 2) {
    thr<4><<<p.multiProcessorCount, warps * 32>>>(out, clk);
    cudaMemcpy(h, clk, 8 * p.multiProcessorCount, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < p.multiProcessorCount; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("DMMA throughput %2d warps/SM ILP4: %7.3f DMMA/clk/SM = %6.1f FMA/clk/SM\n", warps, (double)N * 4 * warps / mx,
           256.0 * N * 4 * warps / mx);
  }
  check<<<1, 32>>>(out); cudaMemcpy(hv, out, 64 * 8, cudaMemcpyDeviceToHost);
  double mxd = 0; for (int i = 0; i < 32; ++i) { double d = hv[i] - hv[32 + i]; if (d < 0) d = -d; if (d > mxd) mxd = d; }
  printf("sum via DMMA %.17g  via shuffles %.17g  max |diff| %.3g  (all lanes equal: %d)\n", hv[0], hv[32], mxd,
         hv[0] == hv[31]);
  check_quad<<<1, 32>>>(out); { double q[96]; cudaMemcpy(q, out, 96 * 8, cudaMemcpyDeviceToHost);
    int ok = 1; for (int i = 0; i < 32; ++i) ok &= (q[i] == q[32 + i]) && (q[i] == q[4 * (i / 4)]);
    double md = 0; for (int i = 0; i < 32; ++i) { double d = q[i] - q[64 + i]; if (d < 0) d = -d; if (d > md) md = d; }
    printf("quad-sum via DMMA(A=data,B=ones): d0==d1 and uniform in quad: %d, max|diff| vs shuffles %.3g (q0 %.6f q7 %.6f)\n", ok, md, q[0], q[28]); }
  {
    auto runmix = [&](auto kern, const char* nm, int nf, int nm_) {
      for (int warps = 8; warps <= 16; warps *= 2) {
        kern<<<p.multiProcessorCount, warps * 32>>>(out, clk);
        cudaMemcpy(h, clk, 8 * p.multiProcessorCount, cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < p.multiProcessorCount; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%s %2d warps/SM: %8.1f clk per iteration per SM-wide warp set -> %.2f clk/SMSP per (DFMA x%d + DMMA x%d)\n", nm, warps,
               mx / N, mx / N / (warps / 4.0), nf, nm_);
      }
    };
    runmix(thr_mix<16, 0>, "mix 16 DFMA + 0 DMMA", 16, 0);
    runmix(thr_mix<0, 2>, "mix  0 DFMA + 2 DMMA", 0, 2);
    runmix(thr_mix<16, 2>, "mix 16 DFMA + 2 DMMA", 16, 2);
    runmix(thr_mix<16, 1>, "mix 16 DFMA + 1 DMMA", 16, 1);
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
