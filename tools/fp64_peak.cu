// Microbenchmark: achievable fp64 FMA rate and shared-memory fp64 update rate per SM on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void smem_update_kernel(double* out, int iters) {
  __shared__ double sm[32 * 129];
  for (int t = threadIdx.x; t < 32 * 129; t += blockDim.x) sm[t] = t;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double l = 1e-3 * lane;
  for (int i = 0; i < iters; ++i) {
    double* col = sm + (w % 32) * 129;
    const double u = col[128];
    col[lane] = fma(-l, u, col[lane]);
    col[lane + 32] = fma(-l, u, col[lane + 32]);
    col[lane + 64] = fma(-l, u, col[lane + 64]);
    col[lane + 96] = fma(-l, u, col[lane + 96]);
  }
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}
int main() {
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) {
    const int iters = 20000;
    dfma_kernel<<<148, threads>>>(out, 100);
    cudaEventRecord(e0); dfma_kernel<<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * threads * iters * 8;
    printf("dfma  threads/SM=%4d : %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", threads, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
    smem_update_kernel<<<148, threads>>>(out, 100);
    cudaEventRecord(e0); smem_update_kernel<<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double upd = 148.0 * threads * iters * 4;
    printf("smem rank-1 update threads/SM=%4d : %.1f FMA/clk/SM\n", threads, upd / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
