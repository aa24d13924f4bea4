// microbench_fp64.cu -- dependent-chain latency and throughput of the fp64 building blocks used by the TSQR leaf
// (DFMA, SHFL.64, rsqrt, rcp, LDS, __syncthreads) on the device at hand.  Development aid.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void lat_dfma(double* out, long long* clk, double a, double b) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void thr_dfma(double* out, long long* clk, double a, double b) {
  double x[ILP];
  for (int q = 0; q < ILP; ++q) x[q] = out[threadIdx.x] + q;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) x[q] = fma(x[q], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < ILP; ++q) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_shfl(double* out, long long* clk) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_rsqrt(double* out, long long* clk) {
  double x = out[threadIdx.x] + 2.0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_rcp(double* out, long long* clk) {
  double x = out[threadIdx.x] + 2.0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = __drcp_rn(x) + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_div(double* out, long long* clk) {
  double x = out[threadIdx.x] + 2.0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 3.0 / x + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_sqrt(double* out, long long* clk) {
  double x = out[threadIdx.x] + 2.0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_lds(double* out, long long* clk) {
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  double x = out[threadIdx.x];
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { double v = sh[idx]; idx = (int)v; x += v; }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void lat_bar(double* out, long long* clk) {
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 1 << 24); cudaMemset(out, 0, 1 << 24); cudaMalloc(&clk, 8 * 4096);
  long long h[4096];
  auto rep = [&](const char* name, double per) { printf("%-34s %8.2f clk\n", name, per); };
#define RUN1(k, nm, ...) k<<<1, 32>>>(out, clk, ##__VA_ARGS__); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost); rep(nm, (double)h[0] / N);
  RUN1(lat_dfma, "DFMA dependent latency", 1.0000001, 1e-9)
  RUN1(lat_shfl, "SHFL.64+DADD dependent latency")
  RUN1(lat_rsqrt, "rsqrt()+DADD dependent latency")
  RUN1(lat_rcp, "__drcp_rn()+DADD dependent latency")
  RUN1(lat_div, "3.0/x+DADD dependent latency")
  RUN1(lat_sqrt, "sqrt()+DADD dependent latency")
  RUN1(lat_lds, "LDS.64->cvt->LDS pointer chase")
  lat_bar<<<1, 256>>>(out, clk); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost); rep("__syncthreads (256 thr, 1 CTA)", (double)h[0] / N);
  // throughput: 148*k CTAs x 256 threads, ILP 8: lane-FMAs per clk per SM
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int warps = 4; warps <= 32; warps *= 2) {
    thr_dfma<8><<<p.multiProcessorCount, warps * 32>>>(out, clk, 1.0000001, 1e-9);
    cudaMemcpy(h, clk, 8 * p.multiProcessorCount, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < p.multiProcessorCount; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("DFMA throughput, %2d warps/SM ILP8: %7.2f lane-FMA/clk/SM  (%.1f TFLOP/s at %d MHz x %d SMs)\n", warps,
           (double)N * 8 * warps * 32 / mx, 2.0 * N * 8 * warps * 32 / mx * p.clockRate * 1e3 * p.multiProcessorCount / 1e12,
           p.clockRate / 1000, p.multiProcessorCount);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
