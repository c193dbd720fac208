// szb_lu.cu -- batched fixed-pattern multifrontal sparse LU for sm_100a, batch-interleaved.
//
// Replaces, for a whole batch of scenarios that share one sparsity pattern, the per-call
// SuperLU factorisation behind the reference's linear-solve seam
//   solve(A,b) -> scipy.sparse.linalg.spsolve      (Solverz/solvers/laesolver.py:11-15)
//   lu_decomposition(A) -> splu ; .solve(b)        (Solverz/solvers/laesolver.py:18-45)
//
// Design (DESIGN.md "LU"; the per-front arithmetic is csrc/szb_lu_front.cuh):
//   * one CTA per GROUP of 32 scenarios, lane = scenario; factors, contribution blocks and the
//     pivot-order vectors are stored interleaved (element e of a group at [e][lane]), so every warp
//     access is one coalesced 256-byte line and every index of the host schedule is decoded once per
//     32 scenarios; lanes never communicate (no shuffles, no atomics => bitwise reproducible and
//     independent of a scenario's position in the batch);
//   * the active scenarios of a call are compacted into dense groups first (lu_map_kernel), so a
//     ragged active set (Newton stragglers, Rodas scenarios that finished) costs what is active;
//   * the host analysis (lu/symbolic.py) fixes pivots, ordering, supernodes and the split of the
//     elimination tree into warp-level subtrees (each warp of the CTA walks its own list of small
//     fronts without any synchronisation) and the CTA-level top (all warps share a front, CTA
//     barriers between the phases);
//   * fronts are assembled and eliminated in place in the factor storage / contribution workspace;
//     the right-hand side can ride along as an extra front column, which fuses the forward
//     substitution L y = b into the factorisation (Newton);
//   * the triangular sweeps follow the same tree with the same lane = scenario layout: the backward
//     sweep streams U once, coalesced.
#include "szb_lu.cuh"

#include <cstdio>

#include "szb_lu_front.cuh"

namespace szb {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
// shared memory of the factor kernels in lines of 32 doubles: a level-0 front (front_lines(f) <= kWarpLines;
// lu/symbolic.py cuts those subtrees at warp_front_max = 10: 10 * 11 = 110 lines) is assembled and eliminated
// in its warp's buffer; the tiled levels use kTileScratch lines per warp (+ one shared diagonal tile)
constexpr int kWarpLines = 110;
constexpr int kSmemLines = kWarps * kWarpLines;   // 880 lines = 225 280 bytes >= kNB * kNB + kWarps * kTileScratch

__device__ __forceinline__ Tabs make_tabs(const LuDev& S) {
  Tabs T;
  T.fh = S.fh; T.ct = S.ct; T.a_src = S.a_src; T.a_dst = S.a_dst; T.rel = S.rel; T.idx = S.idx;
  T.rowp = S.rowp; T.colp = S.colp; T.atp = S.atp; T.tcp = S.tcp; T.tce = S.tce;
  return T;
}

// ---- compaction of the active scenarios into dense groups ---------------------------------------
__global__ void __launch_bounds__(1024) lu_map_kernel(const int* __restrict__ active, int B, int* __restrict__ map,
                                                      int* __restrict__ npos) {
  __shared__ int wsum[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int a = b < B && (!active || active[b]) ? 1 : 0;
    const unsigned m = __ballot_sync(0xffffffffu, a);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int off = base;
    for (int i = 0; i < w; ++i) off += wsum[i];
    if (a) map[off + __popc(m & ((1u << lane) - 1u))] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < 32; ++i) t += wsum[i];
      base += t;
    }
    __syncthreads();
  }
  const int n = base;
  const int end = (n + kGroup - 1) / kGroup * kGroup;
  for (int p = n + threadIdx.x; p < end; p += 1024) map[p] = -1;
  if (threadIdx.x == 0) *npos = n;
}

#ifdef SZB_LU_TIMING
__device__ unsigned long long g_lu_cycles[24];
#endif

// ---- factorisation: one kernel per level of the schedule --------------------------------------------
//   LEVEL 0: each warp walks its list of small subtrees, a front is assembled and eliminated in the warp's
//            shared-memory buffer and written out once;
//   LEVEL 1: each warp walks its list of medium subtrees by register tiles (operands from L2 through the
//            warp's asynchronous-copy ring);
//   LEVEL 2: the top of the tree, all warps of the CTA share a front (same tiles, dealt to the warps).
// Separate launches: each body gets the whole register file for its inner loop (merged into one kernel ptxas
// spilled 6 kB), and contribution blocks that cross levels live in the per-group workspace.
template <int LEVEL>
__global__ void __launch_bounds__(kThreads, 1)
lu_factor_kernel(const LuDev S, const double* __restrict__ JV, double* LU, double* CBW,
                 const double* __restrict__ mvals, const double* __restrict__ slot_mask,
                 const double* __restrict__ scale_b, double scale_const, const double* __restrict__ RHS,
                 double* YI, const int* __restrict__ map, const int* __restrict__ npos, double* MON,
                 int* __restrict__ flags, double growth_tol) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (*npos + kGroup - 1) / kGroup;
  const Tabs T = make_tabs(S);
  int wq0 = 0, wq1 = 0;
  if (LEVEL == 0 && w < S.n_warp_lists) { wq0 = S.wl_ptr[w]; wq1 = S.wl_ptr[w + 1]; }
  if (LEVEL == 1 && w < S.n_mid_lists) { wq0 = S.n_warp_sn + S.ml_ptr[w]; wq1 = S.n_warp_sn + S.ml_ptr[w + 1]; }
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int b = map[g * kGroup + lane];
    Lane L;
    L.valid = b >= 0;
    L.lu = LU + static_cast<size_t>(g) * S.lu_size * kGroup + lane;
    L.cb = CBW + static_cast<size_t>(g) * S.cb_size * kGroup + lane;
    L.cbsrc = L.cb;
    L.y = YI + static_cast<size_t>(g) * S.n * kGroup + lane;
    L.jv = L.valid ? JV + static_cast<size_t>(b) * S.nnz_a : nullptr;
    L.with_rhs = RHS != nullptr;
    L.rhs = (L.valid && RHS) ? RHS + static_cast<size_t>(b) * S.n : nullptr;
    L.mvals = mvals;
    L.smask = slot_mask;
    L.scale = (L.valid && scale_b) ? scale_b[b] : scale_const;
    L.plain = mvals == nullptr && slot_mask == nullptr && scale_b == nullptr && scale_const == 1.0;
    L.amax = 0.0; L.umax = 0.0; L.bad = 0;
#ifdef SZB_LU_TIMING
    long long tacc[8];
    for (int i = 0; i < 8; ++i) tacc[i] = 0;
    L.tacc = tacc; L.tbase = 0;
    const long long tl0 = clock64();
#endif
    if (LEVEL == 0) {
      const Coop C{0, 1, false};
      double* buf = sm + static_cast<size_t>(w) * kWarpLines * kGroup + lane;
      for (int q = wq0; q < wq1; ++q) factor_front_buffered(L, load_front(T.fh, q), T, C, buf);
    } else if (LEVEL == 1) {
      const Coop C{0, 1, false};
      double* tbuf = sm + static_cast<size_t>(w) * kTileScratch * kGroup + lane;
      for (int q = wq0; q < wq1; ++q) factor_front_tiled(L, load_front(T.fh, q), T, C, tbuf, nullptr);
    } else {
      const Coop C{w, kWarps, true};
      double* tbuf = sm + static_cast<size_t>(kNB * kNB + w * kTileScratch) * kGroup + lane;
      for (int q = S.n_warp_sn + S.n_mid_sn; q < S.n_sched; ++q)
        factor_front_tiled(L, load_front(T.fh, q), T, C, tbuf, sm + lane);
    }
    __syncthreads();
#ifdef SZB_LU_TIMING
    tacc[7] = clock64() - tl0;
    if (threadIdx.x == 0)
      for (int i = 0; i < 8; ++i) atomicAdd(&g_lu_cycles[8 * LEVEL + i], static_cast<unsigned long long>(tacc[i]));
#endif
    // pivot / growth monitor of the scenario: reduction over the warps, carried across the level kernels in MON
    sm[(3 * w + 0) * kGroup + lane] = L.amax;
    sm[(3 * w + 1) * kGroup + lane] = L.umax;
    sm[(3 * w + 2) * kGroup + lane] = L.bad ? 1.0 : 0.0;
    __syncthreads();
    if (w == 0) {
      double am = 0.0, um = 0.0, bad = 0.0;
      for (int i = 0; i < kWarps; ++i) {
        const double a_ = sm[(3 * i + 0) * kGroup + lane], u_ = sm[(3 * i + 1) * kGroup + lane];
        am = a_ > am ? a_ : am;
        um = u_ > um ? u_ : um;
        bad += sm[(3 * i + 2) * kGroup + lane];
      }
      double* mon = MON + (static_cast<size_t>(g) * kGroup + lane) * 3;
      if (LEVEL > 0) {
        am = mon[0] > am ? mon[0] : am;
        um = mon[1] > um ? mon[1] : um;
        bad += mon[2];
      }
      if (LEVEL < 2) { mon[0] = am; mon[1] = um; mon[2] = bad; }
      else if (L.valid && (bad > 0.0 || !(um <= growth_tol * am))) flags[b] = 2;
    }
    __syncthreads();
  }
}

// ---- forward sweep with stored factors --------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
lu_forward_kernel(const LuDev S, const double* LU, const double* __restrict__ RHS, double* YI, double* CVW,
                  const int* __restrict__ map, const int* __restrict__ npos, const int* __restrict__ active) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (*npos + kGroup - 1) / kGroup;
  const Tabs T = make_tabs(S);
  int wq0 = 0, wq1 = 0, mq0 = 0, mq1 = 0;
  if (w < S.n_warp_lists) { wq0 = S.wl_ptr[w]; wq1 = S.wl_ptr[w + 1]; }
  if (w < S.n_mid_lists) { mq0 = S.n_warp_sn + S.ml_ptr[w]; mq1 = S.n_warp_sn + S.ml_ptr[w + 1]; }
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int b = map[g * kGroup + lane];
    Lane L;
    L.valid = b >= 0 && (!active || active[b]);
    L.lu = const_cast<double*>(LU) + static_cast<size_t>(g) * S.lu_size * kGroup + lane;
    L.cb = nullptr;
    L.y = YI + static_cast<size_t>(g) * S.n * kGroup + lane;
    L.jv = nullptr;
    L.rhs = L.valid ? RHS + static_cast<size_t>(b) * S.n : nullptr;
    L.with_rhs = true;
    double* cv = CVW + static_cast<size_t>(blockIdx.x) * S.cv_size * kGroup + lane;
    {
      const Coop C{0, 1, false};
      for (int q = wq0; q < wq1; ++q) forward_front(L, cv, load_front(T.fh, q), T, C);
      __syncthreads();
      for (int q = mq0; q < mq1; ++q) forward_front(L, cv, load_front(T.fh, q), T, C);
    }
    __syncthreads();
    {
      const Coop C{w, kWarps, true};
      for (int q = S.n_warp_sn + S.n_mid_sn; q < S.n_sched; ++q) forward_front(L, cv, load_front(T.fh, q), T, C);
    }
    __syncthreads();
  }
}

// ---- backward sweep (+ scatter to the caller's scenario-major vector) ---------------------------------
__global__ void __launch_bounds__(kThreads, 1)
lu_backward_kernel(const LuDev S, const double* LU, double* YI, double* __restrict__ X, int mode,
                   const int* __restrict__ map, const int* __restrict__ npos, const int* __restrict__ active) {
  __shared__ double red[kWarps * kNB * kGroup];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (*npos + kGroup - 1) / kGroup;
  const Tabs T = make_tabs(S);
  int wq0 = 0, wq1 = 0, mq0 = 0, mq1 = 0;
  if (w < S.n_warp_lists) { wq0 = S.wl_ptr[w]; wq1 = S.wl_ptr[w + 1]; }
  if (w < S.n_mid_lists) { mq0 = S.n_warp_sn + S.ml_ptr[w]; mq1 = S.n_warp_sn + S.ml_ptr[w + 1]; }
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int b = map[g * kGroup + lane];
    Lane L;
    L.valid = b >= 0 && (!active || active[b]);
    L.lu = const_cast<double*>(LU) + static_cast<size_t>(g) * S.lu_size * kGroup + lane;
    L.cb = nullptr;
    L.y = YI + static_cast<size_t>(g) * S.n * kGroup + lane;
    {
      const Coop C{w, kWarps, true};
      for (int q = S.n_sched - 1; q >= S.n_warp_sn + S.n_mid_sn; --q)
        backward_front(L, load_front(T.fh, q), T, C, red + lane);
    }
    __syncthreads();
    {
      const Coop C{0, 1, false};
      for (int q = mq1 - 1; q >= mq0; --q) backward_front(L, load_front(T.fh, q), T, C, red + lane);
      __syncthreads();
      for (int q = wq1 - 1; q >= wq0; --q) backward_front(L, load_front(T.fh, q), T, C, red + lane);
    }
    __syncthreads();
    if (L.valid) {
      double* xb = X + static_cast<size_t>(b) * S.n;
      for (int k = w; k < S.n; k += kWarps) {
        const double v = SZB_EL(L.y, k);
        const int col = __ldg(S.colp + k);
        if (mode == 1) xb[col] -= v; else xb[col] = v;
      }
    }
    __syncthreads();
  }
}

}  // namespace

cudaError_t lu_plan(const LuDev& S, int B, LuLaunchCfg* cfg) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (S.n_warp_lists > kWarps || S.n_mid_lists > kWarps || S.warp_lines != kWarpLines ||
      S.f_warp_max * (S.f_warp_max + 1) > kWarpLines)
    return cudaErrorInvalidConfiguration;
  cfg->threads = kThreads;
  cfg->smem = static_cast<size_t>(kSmemLines) * kGroup * sizeof(double);
  const int sm_bytes = static_cast<int>(cfg->smem);
  if ((e = cudaFuncSetAttribute(lu_factor_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lu_factor_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lu_factor_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_bytes)) != cudaSuccess) return e;
  cfg->groups = (B + kGroup - 1) / kGroup;
  cfg->grid = sms;                       // one CTA per SM (register file and shared memory are both full)
  if (cfg->grid > cfg->groups) cfg->grid = cfg->groups > 0 ? cfg->groups : 1;
  return cudaSuccess;
}

#ifdef SZB_LU_TIMING
extern "C" __attribute__((visibility("default"))) int szb_debug_lu_cycles(unsigned long long* out24, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out24, g_lu_cycles, sizeof(unsigned long long) * 24);
  if (reset) {
    unsigned long long z[24] = {0};
    cudaMemcpyToSymbol(g_lu_cycles, z, sizeof z);
  }
  return 0;
}
#endif

cudaError_t lu_map_launch(const int* active, int B, int* map, int* npos, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  lu_map_kernel<<<1, 1024, 0, st>>>(active, B, map, npos);
  return cudaGetLastError();
}

cudaError_t lu_factor_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* JV, double* LU, double* CBW,
                             const double* mvals, const double* slot_mask, const double* scale_b,
                             double scale_const, const double* RHS, double* YI, const int* map, const int* npos,
                             double* MON, int* flags, double growth_tol, cudaStream_t st) {
  lu_factor_kernel<0><<<cfg.grid, cfg.threads, cfg.smem, st>>>(S, JV, LU, CBW, mvals, slot_mask, scale_b, scale_const,
                                                               RHS, YI, map, npos, MON, flags, growth_tol);
  lu_factor_kernel<1><<<cfg.grid, cfg.threads, cfg.smem, st>>>(S, JV, LU, CBW, mvals, slot_mask, scale_b, scale_const,
                                                               RHS, YI, map, npos, MON, flags, growth_tol);
  lu_factor_kernel<2><<<cfg.grid, cfg.threads, cfg.smem, st>>>(S, JV, LU, CBW, mvals, slot_mask, scale_b, scale_const,
                                                               RHS, YI, map, npos, MON, flags, growth_tol);
  return cudaGetLastError();
}

cudaError_t lu_solve_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* LU, const double* RHS, double* X,
                            double* CVW, double* YI, int do_fwd, int mode, const int* map, const int* npos,
                            const int* active, cudaStream_t st) {
  if (do_fwd) {
    lu_forward_kernel<<<cfg.grid, cfg.threads, 0, st>>>(S, LU, RHS, YI, CVW, map, npos, active);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  lu_backward_kernel<<<cfg.grid, cfg.threads, 0, st>>>(S, LU, YI, X, mode, map, npos, active);
  return cudaGetLastError();
}

}  // namespace szb
