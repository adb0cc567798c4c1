// DMMA (mma.sync m8n8k4 f64) issue/latency micro-benchmark: cycles per DMMA per warp for C independent accumulator
// chains and W warps per CTA (one CTA).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_lat tools/dmma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, long long* cyc, int iters) {
  double c0[C], c1[C];
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < C; ++i) c0[i] = c1[i] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < C; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < C; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int C>
void run(int warps) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<C><<<1, 32 * warps>>>(out, cyc, iters);
  k<C><<<1, 32 * warps>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("chains %d warps %2d (per SMSP %d): %.1f cycles per DMMA per warp, %.1f cycles per DMMA per SMSP\n", C, warps, (warps + 3) / 4,
         (double)h / (iters * C), (double)h / (iters * C) / ((warps + 3) / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
