// Microbenchmark of the building blocks of the blocked (DMMA) front elimination of csrc/szb_lu.cu
// on one SM: cycles (clock64) of the diagonal-block LU + inverse, one phase-A tile, one phase-B strip,
// the dependent DMMA latency, and a whole front.  Build (on the GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include -I solverz_b200/csrc tools/lu_micro.cu -o tools/lu_micro
#include "../solverz_b200/csrc/szb_lu.cu"

using namespace szb;

__global__ void __launch_bounds__(512, 1) micro_kernel(long long* out, int f, int s, int reps) {
  extern __shared__ double sm[];
  const int ld = ld_blk(f), lts = lt_stride(f);
  const int fm_off = 0, lt_off = front_elems_blk(f);
  double* Fm = sm;
  double* lt = sm + lt_off;
  double* dinv = lt + kNB * lts;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __shared__ int flag;
  auto fill = [&]() {
    for (int t = threadIdx.x; t < ld * cols_blk(f); t += blockDim.x) {
      const int r = t % ld, c = t / ld;
      Fm[t] = (r < f && c <= f) ? (r == c ? 4.0 + f : 0.0) + 0.01 * ((r * 7 + c * 13) % 17) : 0.0;
    }
    __syncthreads();
  };
  fill();
  long long t0, t1;
  // (0) dependent DMMA latency
  if (w == 0) {
    double c0 = 0, c1 = 0, a = 1e-3 * lane, b = 1e-3;
    t0 = clock64();
    for (int i = 0; i < 256; ++i) dmma(c0, c1, a, b);
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / 256;
    if (c0 == 123.0) out[63] = 1;
  }
  __syncthreads();
  // (1) diag_block alone, one warp
  if (w == 0) {
    t0 = clock64();
    for (int i = 0; i < reps; ++i) { diag_block(Fm, dinv, kNB * (i % 4), kNB, ld, lane, &flag); __syncwarp(); }
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / reps;
  }
  __syncthreads();
  fill();
  // (2) phase A tile, one warp
  if (w == 0) {
    const int q = lane & 3, g8 = lane >> 2;
    const double bd = g8 < kNB ? dinv[q + kNB * g8] : 0.0;
    t0 = clock64();
    for (int i = 0; i < reps; ++i) lt_tile(Fm, lt, lts, nullptr, 0, kNB, f, ld, i % 8, bd, 0.0, lane);
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / reps;
  }
  __syncthreads();
  // (3) phase B: one warp alone, 4 column tiles of a 32-row strip
  if (w == 0) {
    t0 = clock64();
    for (int i = 0; i < reps; ++i) trailing_strips<4>(Fm, lt, lts, 0, kNB, ld, 0, 0, 1, 4, lane);
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / reps;
  }
  __syncthreads();
  // (4) phase B: all 16 warps, 4 column tiles each (the f = 128 panel-0 shape)
  t0 = clock64();
  for (int i = 0; i < reps; ++i) trailing_rows(Fm, lt, lts, 0, kNB, f - kNB, ld, (w & 3) << 5, w >> 2, 4, 16, lane);
  __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = (t1 - t0) / reps;
  // (5) whole front
  fill();
  PanelGrp g; g.id = 0; g.nw = blockDim.x >> 5; g.w0 = 0;
  t0 = clock64();
  eliminate_blocked<true>(g, Fm, fm_off, lt_off, lts, nullptr, s, f, f + 1, ld, &flag);
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = t1 - t0;
  // (6) shuffle latency chain, LDS latency chain
  if (w == 0) {
    double v = lane;
    t0 = clock64();
    for (int i = 0; i < 256; ++i) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0) / 256;
    if (v == 1234.5) out[62] = 1;
    int idx = lane;
    int* ism = reinterpret_cast<int*>(sm);
    for (int t = threadIdx.x; t < 1024; t += 32) ism[t] = (t * 33 + 7) & 1023;
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < 256; ++i) idx = ism[idx];
    t1 = clock64();
    if (lane == 0) out[7] = (t1 - t0) / 256;
    if (idx == -5) out[61] = 1;
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64 * sizeof(long long));
  for (int f : {128, 72}) {
    const int s = f == 128 ? 124 : 20;
    const size_t smem = sizeof(double) * group_elems(f);
    cudaFuncSetAttribute(micro_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int i = 0; i < 64; ++i) out[i] = 0;
    micro_kernel<<<1, f == 128 ? 512 : 128, smem>>>(out, f, s, 50);
    cudaError_t e = cudaDeviceSynchronize();
    printf("f=%d s=%d (%s): dmma dep latency %lld | diag_block %lld | lt_tile %lld | strip x4 (1 warp) %lld | strips (all warps) %lld | "
           "whole front %lld | shfl %lld | lds %lld cycles\n",
           f, s, cudaGetErrorString(e), out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
  }
  return 0;
}
