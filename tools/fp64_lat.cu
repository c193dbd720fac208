// Microbenchmark: dependent-chain latencies (cycles) of the primitives the LU kernels are made of.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int iters) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) sm[t] = 1.0 + 1e-3 * t;
  __syncthreads();
  double a = 1.0 + lane * 1e-3, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) a = fma(a, b, c);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  // division chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) a = 1.0 / a + 0.5;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  // rcp approx + 2 newton (fast reciprocal) chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double x = a; double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r);
    a = r + 0.5;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  // shuffle (double = 2 shfl) chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  // LDS chain (pointer chase through doubles)
  int idx = lane;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { double v = sm[idx]; idx = (int)(v) + ((idx + 1) & 1023) - 1; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  a += idx;
  // STS -> syncwarp -> LDS round trip
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] ; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  // __syncthreads
  t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  // STS -> syncthreads -> LDS
  t0 = clock64();
  for (int i = 0; i < iters; ++i) { sm[threadIdx.x] = a; __syncthreads(); a = sm[(threadIdx.x + 33) % blockDim.x]; __syncthreads(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
  // float rcp + dfma refine
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    double x = a; double r = (double)(__frcp_rn((float)x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r); 
    a = r + 0.5;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
  out[threadIdx.x] = a;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMallocManaged(&cyc, 128);
  const int iters = 2000;
  for (int threads : {32, 256, 1024}) {
    lat<<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("threads %4d: dfma %.1f  div(+add) %.1f  rcp.approx+2NR(+add) %.1f  shfl.f64 %.1f  lds-chase %.1f  sts/syncwarp/lds/syncwarp %.1f  syncthreads %.1f  sts/sync/lds/sync %.1f  frcp+2NR(+add) %.1f  (%s)\n", threads,
           cyc[0] / (double)iters, cyc[1] / (double)iters, cyc[2] / (double)iters, cyc[3] / (double)iters, cyc[4] / (double)iters,
           cyc[5] / (double)iters, cyc[6] / (double)iters, cyc[7] / (double)iters, cyc[8] / (double)iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
