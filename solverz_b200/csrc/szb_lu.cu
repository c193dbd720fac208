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

__device__ __forceinline__ Tabs make_tabs(const LuDev& S) {
  Tabs T;
  T.fh = S.fh; T.ct = S.ct; T.a_src = S.a_src; T.a_dst = S.a_dst; T.rel = S.rel; T.idx = S.idx;
  T.rowp = S.rowp; T.colp = S.colp;
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

// ---- factorisation ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
lu_factor_kernel(const LuDev S, const double* __restrict__ JV, double* LU, double* CBW,
                 const double* __restrict__ mvals, const double* __restrict__ slot_mask,
                 const double* __restrict__ scale_b, double scale_const, const double* __restrict__ RHS,
                 double* YI, const int* __restrict__ map, const int* __restrict__ npos, int* __restrict__ flags,
                 double growth_tol) {
  __shared__ double red_a[kWarps][32], red_u[kWarps][32];
  __shared__ int red_b[kWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (*npos + kGroup - 1) / kGroup;
  const Tabs T = make_tabs(S);
  int wq0 = 0, wq1 = 0;
  if (w < S.n_warp_lists) { wq0 = S.wl_ptr[w]; wq1 = S.wl_ptr[w + 1]; }
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int b = map[g * kGroup + lane];
    Lane L;
    L.valid = b >= 0;
    L.lu = LU + static_cast<size_t>(g) * S.lu_size * kGroup + lane;
    L.cb = CBW + static_cast<size_t>(blockIdx.x) * S.cb_size * kGroup + lane;
    L.y = YI + static_cast<size_t>(g) * S.n * kGroup + lane;
    L.jv = L.valid ? JV + static_cast<size_t>(b) * S.nnz_a : nullptr;
    L.with_rhs = RHS != nullptr;
    L.rhs = (L.valid && RHS) ? RHS + static_cast<size_t>(b) * S.n : nullptr;
    L.mvals = mvals;
    L.smask = slot_mask;
    L.scale = (L.valid && scale_b) ? scale_b[b] : scale_const;
    L.plain = mvals == nullptr && slot_mask == nullptr && scale_b == nullptr && scale_const == 1.0;
    L.amax = 0.0; L.umax = 0.0; L.bad = 0;
    // level 0: every warp walks its own subtrees
    {
      const Coop C{0, 1, false};
      for (int q = wq0; q < wq1; ++q) {
        const Front F = load_front(T.fh, q);
        factor_front(L, F, T, C);
      }
    }
    __syncthreads();
    // level 1: the top of the tree, all warps per front
    {
      const Coop C{w, kWarps, true};
      for (int q = S.n_warp_sn; q < S.n_sched; ++q) {
        const Front F = load_front(T.fh, q);
        factor_front(L, F, T, C);
      }
    }
    // pivot / growth monitor of the scenario = reduction over the warps of the lane's values
    red_a[w][lane] = L.amax; red_u[w][lane] = L.umax; red_b[w][lane] = L.bad;
    __syncthreads();
    if (w == 0 && L.valid) {
      double am = 0.0, um = 0.0;
      int bad = 0;
      for (int i = 0; i < kWarps; ++i) {
        am = red_a[i][lane] > am ? red_a[i][lane] : am;
        um = red_u[i][lane] > um ? red_u[i][lane] : um;
        bad |= red_b[i][lane];
      }
      if (bad || !(um <= growth_tol * am)) flags[b] = 2;
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
  int wq0 = 0, wq1 = 0;
  if (w < S.n_warp_lists) { wq0 = S.wl_ptr[w]; wq1 = S.wl_ptr[w + 1]; }
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
    }
    __syncthreads();
    {
      const Coop C{w, kWarps, true};
      for (int q = S.n_warp_sn; q < S.n_sched; ++q) forward_front(L, cv, load_front(T.fh, q), T, C);
    }
    __syncthreads();
  }
}

// ---- backward sweep (+ scatter to the caller's scenario-major vector) ---------------------------------
__global__ void __launch_bounds__(kThreads, 1)
lu_backward_kernel(const LuDev S, const double* LU, double* YI, double* __restrict__ X, int mode,
                   const int* __restrict__ map, const int* __restrict__ npos, const int* __restrict__ active) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (*npos + kGroup - 1) / kGroup;
  const Tabs T = make_tabs(S);
  int wq0 = 0, wq1 = 0;
  if (w < S.n_warp_lists) { wq0 = S.wl_ptr[w]; wq1 = S.wl_ptr[w + 1]; }
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int b = map[g * kGroup + lane];
    Lane L;
    L.valid = b >= 0 && (!active || active[b]);
    L.lu = const_cast<double*>(LU) + static_cast<size_t>(g) * S.lu_size * kGroup + lane;
    L.cb = nullptr;
    L.y = YI + static_cast<size_t>(g) * S.n * kGroup + lane;
    {
      const Coop C{w, kWarps, true};
      for (int q = S.n_sched - 1; q >= S.n_warp_sn; --q) backward_front(L, load_front(T.fh, q), T, C);
    }
    __syncthreads();
    {
      const Coop C{0, 1, false};
      for (int q = wq1 - 1; q >= wq0; --q) backward_front(L, load_front(T.fh, q), T, C);
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
  if (S.n_warp_lists > kWarps) return cudaErrorInvalidConfiguration;
  cfg->threads = kThreads;
  cfg->smem = 0;
  cfg->groups = (B + kGroup - 1) / kGroup;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lu_factor_kernel, kThreads, cfg->smem);
  if (occ < 1) occ = 1;
  cfg->grid = sms * occ;
  if (cfg->grid > cfg->groups) cfg->grid = cfg->groups > 0 ? cfg->groups : 1;
  return cudaSuccess;
}

cudaError_t lu_map_launch(const int* active, int B, int* map, int* npos, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  lu_map_kernel<<<1, 1024, 0, st>>>(active, B, map, npos);
  return cudaGetLastError();
}

cudaError_t lu_factor_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* JV, double* LU, double* CBW,
                             const double* mvals, const double* slot_mask, const double* scale_b,
                             double scale_const, const double* RHS, double* YI, const int* map, const int* npos,
                             int* flags, double growth_tol, cudaStream_t st) {
  lu_factor_kernel<<<cfg.grid, cfg.threads, cfg.smem, st>>>(S, JV, LU, CBW, mvals, slot_mask, scale_b, scale_const,
                                                            RHS, YI, map, npos, flags, growth_tol);
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
