// micro-timing of the register panel factorisation (development aid)
#include "../solverz_b200/csrc/szb_lu.cu"
#include <cstdio>
namespace szb { namespace {
__global__ void micro(double* out, long long* cyc, int f, int ld, int nb, int mode) {
  extern __shared__ double sm[];
  double* Fm = sm; double* lt = sm + ld * (f + 1);
  for (int t = threadIdx.x; t < ld * (f + 1); t += blockDim.x) Fm[t] = 1.0 / (1 + (t % 7)) + ((t % (ld + 1)) == 0 ? 10.0 : 0.0);
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) lt[t] = 1e-3;
  __syncthreads();
  int flag = 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long t0 = clock64();
  if (w == 0) {
    panel_next<4>(Fm, lt, lt + 512, 0, nb, nb, f, ld, lane, &flag);
  } else if (mode == 1) {
    trailing_worker<4>(Fm, lt, 0, nb, f, ld, 8, f + 1, w - 1, 31, lane);
  }
  long long t1 = clock64();
  if (lane == 0) cyc[w] = t1 - t0;
  __syncthreads();
  out[threadIdx.x] = Fm[threadIdx.x] + lt[threadIdx.x % 64];
}
}}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMallocManaged(&cyc, 512);
  cudaFuncSetAttribute(szb::micro, cudaFuncAttributeMaxDynamicSharedMemorySize, 160000);
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 2; ++rep) {
      szb::micro<<<1, 1024, 160000>>>(out, cyc, 124, 125, 4, mode);
      cudaDeviceSynchronize();
      printf("mode %d: panel_next (warp 0) %lld cycles; workers: w1 %lld w8 %lld w31 %lld (%s)\n", mode, cyc[0], cyc[1], cyc[8], cyc[31], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
