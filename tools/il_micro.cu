// il_micro.cu -- microbenchmark for the batch-interleaved (lane = scenario) trailing update:
// C[i][j] -= sum_k L[i][k] * U[k][j] where every element is a 256-byte line of 32 scenarios.
// One CTA per group of 32 scenarios, warps own register tiles of RT x CT lines.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int RT, int CT, int NB, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
upd_kernel(double* __restrict__ Cg, const double* __restrict__ Lg, const double* __restrict__ Ug, int u, int reps,
           long long* cyc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  double* C = Cg + (size_t)blockIdx.x * u * u * 32 + lane;
  const double* L = Lg + (size_t)blockIdx.x * u * NB * 32 + lane;   // column-major u x NB
  const double* U = Ug + (size_t)blockIdx.x * u * NB * 32 + lane;   // row-major-in-k: U[k + NB*j]
  const int tr = u / RT, tc = u / CT;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int t = w; t < tr * tc; t += W) {
      const int ti = t % tr, tj = t / tr;
      double acc[RT][CT];
#pragma unroll
      for (int j = 0; j < CT; ++j)
#pragma unroll
        for (int i = 0; i < RT; ++i) acc[i][j] = C[((size_t)(tj * CT + j) * u + ti * RT + i) * 32];
#pragma unroll 2
      for (int k = 0; k < NB; ++k) {
        double a[RT], b[CT];
#pragma unroll
        for (int i = 0; i < RT; ++i) a[i] = L[((size_t)k * u + ti * RT + i) * 32];
#pragma unroll
        for (int j = 0; j < CT; ++j) b[j] = U[((size_t)(tj * CT + j) * NB + k) * 32];
#pragma unroll
        for (int j = 0; j < CT; ++j)
#pragma unroll
          for (int i = 0; i < RT; ++i) acc[i][j] = fma(-a[i], b[j], acc[i][j]);
      }
#pragma unroll
      for (int j = 0; j < CT; ++j)
#pragma unroll
        for (int i = 0; i < RT; ++i) C[((size_t)(tj * CT + j) * u + ti * RT + i) * 32] = acc[i][j];
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int RT, int CT, int NB, int MAXT>
void run(int u, int threads, int grid, double* C, double* L, double* U, long long* cyc) {
  const int reps = 8;
  upd_kernel<RT, CT, NB, MAXT><<<grid, threads>>>(C, L, U, u, 1, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  upd_kernel<RT, CT, NB, MAXT><<<grid, threads>>>(C, L, U, u, reps, cyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[1024]; cudaMemcpy(hc, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
  const double fma_per_cta = (double)(u / RT) * (u / CT) * RT * CT * NB * 32.0 * reps;
  printf("RT=%d CT=%d NB=%2d u=%3d thr=%4d grid=%3d : %8.3f ms  max cyc %9lld  FMA/clk/SM %6.2f  (%s)\n", RT, CT, NB, u,
         threads, grid, ms, mx, fma_per_cta / (double)mx, cudaGetErrorString(e));
}

int main() {
  const int grid = 148, umax = 128, nbmax = 32;
  double *C, *L, *U; long long* cyc;
  cudaMalloc(&C, sizeof(double) * grid * umax * umax * 32);
  cudaMalloc(&L, sizeof(double) * grid * umax * nbmax * 32);
  cudaMalloc(&U, sizeof(double) * grid * umax * nbmax * 32);
  cudaMalloc(&cyc, sizeof(long long) * 1024);
  cudaMemset(C, 0, sizeof(double) * grid * umax * umax * 32);
  cudaMemset(L, 0, sizeof(double) * grid * umax * nbmax * 32);
  cudaMemset(U, 0, sizeof(double) * grid * umax * nbmax * 32);
  for (int u : {48, 96, 128}) {
    run<4, 4, 8, 512>(u, 512, grid, C, L, U, cyc);
    run<8, 4, 8, 512>(u, 512, grid, C, L, U, cyc);
    run<8, 4, 16, 512>(u, 512, grid, C, L, U, cyc);
    run<8, 4, 32, 512>(u, 512, grid, C, L, U, cyc);
    run<4, 8, 8, 512>(u, 512, grid, C, L, U, cyc);
    run<8, 4, 8, 512>(u, 256, grid, C, L, U, cyc);
    run<8, 4, 8, 1024>(u, 1024, grid, C, L, U, cyc);
    run<4, 4, 8, 1024>(u, 1024, grid, C, L, U, cyc);
    run<8, 8, 8, 256>(u, 256, grid, C, L, U, cyc);
    run<8, 8, 16, 256>(u, 256, grid, C, L, U, cyc);
    run<8, 8, 32, 256>(u, 256, grid, C, L, U, cyc);
    run<8, 8, 8, 384>(u, 384, grid, C, L, U, cyc);
  }
  return 0;
}
