// Microbenchmark: fp64 tensor-core MMA (mma.sync m8n8k4 f64) rate on this GPU, against the DFMA pipe
// (tools/fp64_peak.cu).  One DMMA = 8x8x4 = 256 FMA per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
__global__ void dmma_chain(double* out, long long* cyc, int iters) {
  double a = 1.0 + threadIdx.x * 1e-3, b = 1e-3;
  double c[2] = {0, 0};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = c[0] + c[1];
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4); cudaMallocManaged(&cyc, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {128, 256, 512, 1024}) {
    const int iters = 20000;
    dmma_kernel<<<148, threads>>>(out, 100);
    cudaEventRecord(e0); dmma_kernel<<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * (threads / 32) * iters * 4.0 * 256.0;
    printf("dmma m8n8k4 threads/SM=%4d : %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz) %s\n", threads, 2 * fma / ms * 1e-9,
           fma / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  }
  dmma_chain<<<1, 32>>>(out, cyc, 2000); cudaDeviceSynchronize();
  printf("dependent dmma latency: %.1f cycles\n", cyc[0] / 2000.0);
  return 0;
}
