// szb_lu.cu -- batched fixed-pattern multifrontal sparse LU for sm_100a.
//
// Replaces, for a whole batch of scenarios that share one sparsity pattern, the per-call
// SuperLU factorisation behind the reference's linear-solve seam
//   solve(A,b) -> scipy.sparse.linalg.spsolve      (Solverz/solvers/laesolver.py:11-15)
//   lu_decomposition(A) -> splu ; .solve(b)        (Solverz/solvers/laesolver.py:18-45)
//
// Design (see DESIGN.md "LU"):
//   * one CTA per scenario (persistent grid, grid-stride over scenarios); values are
//     scenario-major so every panel of the factors is one contiguous, coalesced stream;
//   * the host analysis (lu/symbolic.py) fixes pivots, ordering, supernodes and all index
//     maps once; the kernels below are pure value arithmetic driven by int32 tables that are
//     identical for all scenarios (L1/L2 resident);
//   * phase 1: every warp walks its own list of small subtrees (fronts <= f_warp_max) with
//     warp-synchronous dense partial LU in a private shared-memory front -- no CTA barrier;
//     phase 2: the CTA eliminates the top of the tree together, front in shared memory;
//   * contribution blocks travel through a small per-CTA stack in global memory (L2 resident,
//     reused across scenarios by the persistent CTA), factors are written exactly once;
//   * triangular solves follow the same two-phase tree walk; the forward sweep passes
//     contribution vectors up the tree (no atomics => bitwise reproducible), the backward
//     sweep only reads ancestors.  No tensor cores: nothing here is a dense contraction large
//     enough to amortise them, the arithmetic is fp64 FMA fed from shared memory.
#include "szb_lu.cuh"

#include <cstdio>

namespace szb {

namespace {

constexpr int kWarp = 32;

template <bool WARP>
struct Grp {
  __device__ static __forceinline__ int rank() { return WARP ? (threadIdx.x & 31) : threadIdx.x; }
  __device__ static __forceinline__ int size() { return WARP ? kWarp : blockDim.x; }
  __device__ static __forceinline__ void sync() {
    if (WARP) __syncwarp(); else __syncthreads();
  }
};

// t / m for 0 <= t < 2^22 without an integer division (exact)
__device__ __forceinline__ int fdiv(int t, float invm) {
  return __float2int_rz((static_cast<float>(t) + 0.5f) * invm);
}

// ---------------------------------------------------------------------------------------
// numeric factorisation of one supernode
// ---------------------------------------------------------------------------------------
template <bool WARP>
__device__ void factor_supernode(const LuDev& S, int k, double* __restrict__ Fm,
                                 double* __restrict__ dinv, const double* __restrict__ jv,
                                 const double* __restrict__ mvals, double scale,
                                 double* __restrict__ lu, double* __restrict__ cb, int* flag) {
  const int s = S.sn_s[k], f = S.sn_f[k], u = f - s;
  const int r = Grp<WARP>::rank(), G = Grp<WARP>::size();
  const int ff = f * f;
  for (int t = r; t < ff; t += G) Fm[t] = 0.0;
  Grp<WARP>::sync();
  // original entries
  {
    const int a1 = S.a_ptr[k + 1];
    if (mvals) {
      for (int e = S.a_ptr[k] + r; e < a1; e += G) {
        const int src = S.a_src[e];
        Fm[S.a_dst[e]] = fma(scale, jv[src], mvals[src]);
      }
    } else {
      for (int e = S.a_ptr[k] + r; e < a1; e += G) Fm[S.a_dst[e]] = scale * jv[S.a_src[e]];
    }
  }
  Grp<WARP>::sync();
  // extend-add of the children's contribution blocks
  for (int ci = S.child_ptr[k]; ci < S.child_ptr[k + 1]; ++ci) {
    const int c = S.child[ci];
    const int uc = S.sn_f[c] - S.sn_s[c];
    if (uc == 0) continue;
    const int* __restrict__ rel = S.rel + S.rel_ptr[c];
    const double* __restrict__ C = cb + S.cb_off[c];
    const float inv = 1.0f / static_cast<float>(uc);
    const int n2 = uc * uc;
    for (int t = r; t < n2; t += G) {
      const int j = fdiv(t, inv);
      const int i = t - j * uc;
      Fm[rel[i] + rel[j] * f] += C[t];
    }
    Grp<WARP>::sync();
  }
  // right-looking elimination of the s pivots; L is kept unscaled until write-out
  for (int p = 0; p < s; ++p) {
    const double piv = Fm[p + p * f];
    const double inv = 1.0 / piv;
    if (r == 0) {
      dinv[p] = inv;
      if (!(fabs(piv) > 0.0) || isinf(piv)) *flag = 2;
    }
    const int m = f - p - 1;
    double* __restrict__ colp = Fm + (p + 1) + p * f;        // L column (unscaled)
    double* __restrict__ A22 = Fm + (p + 1) + (p + 1) * f;   // trailing block
    const double* __restrict__ rowp = Fm + p + (p + 1) * f;  // U row, stride f
    if (WARP) {
      const float invm = 1.0f / static_cast<float>(m);
      const int n2 = m * m;
      for (int t = r; t < n2; t += kWarp) {
        const int j = fdiv(t, invm);
        const int i = t - j * m;
        A22[i + j * f] = fma(-(colp[i] * inv), rowp[j * f], A22[i + j * f]);
      }
    } else {
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
      for (int j = w; j < m; j += nw) {
        const double ukj = rowp[j * f];
        for (int i = lane; i < m; i += kWarp)
          A22[i + j * f] = fma(-(colp[i] * inv), ukj, A22[i + j * f]);
      }
    }
    Grp<WARP>::sync();
  }
  // write-out: [L11\U11; L21] panel (f x s, column-major), U12^T (u x s), contribution (u x u)
  {
    double* __restrict__ Lp = lu + S.lp_off[k];
    const int fs = f * s;
    const float invf = 1.0f / static_cast<float>(f);
    for (int t = r; t < fs; t += G) {
      const int p = fdiv(t, invf);
      const int i = t - p * f;
      const double v = Fm[t];
      Lp[t] = (i > p) ? v * dinv[p] : v;
    }
    if (u > 0) {
      double* __restrict__ Up = lu + S.up_off[k];
      const float invu = 1.0f / static_cast<float>(u);
      const int us = u * s;
      for (int t = r; t < us; t += G) {
        const int rr = fdiv(t, invu);
        const int c = t - rr * u;
        Up[t] = Fm[rr + (s + c) * f];
      }
      double* __restrict__ C = cb + S.cb_off[k];
      const int uu = u * u;
      for (int t = r; t < uu; t += G) {
        const int j = fdiv(t, invu);
        const int i = t - j * u;
        C[t] = Fm[(s + i) + (s + j) * f];
      }
    }
  }
  Grp<WARP>::sync();
}

__global__ void __launch_bounds__(1024, 1)
lu_factor_kernel(const LuDev S, const double* __restrict__ JV, double* __restrict__ LU,
                 double* __restrict__ CBW, const double* __restrict__ mvals,
                 const double* __restrict__ scale_b, double scale_const,
                 const int* __restrict__ active, int* __restrict__ flags, int B) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5;
  const int fw2 = S.f_warp_max * S.f_warp_max;
  const int ftop2 = S.f_top_max * S.f_top_max;
  const int front_doubles = max(S.n_warp_lists * fw2, ftop2);
  double* dinv_all = sm + front_doubles;  // n_warp_lists*s_warp_max  |  s_top_max
  double* cb = CBW + static_cast<size_t>(blockIdx.x) * S.cb_size;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (active && !active[b]) continue;
    const double* jv = JV + static_cast<size_t>(b) * S.nnz_a;
    double* lu = LU + static_cast<size_t>(b) * S.lu_size;
    const double scale = scale_b ? scale_b[b] : scale_const;
    int* flag = flags + b;
    if (warp < S.n_warp_lists) {
      double* Fm = sm + warp * fw2;
      double* dinv = dinv_all + warp * S.s_warp_max;
      for (int q = S.wl_ptr[warp]; q < S.wl_ptr[warp + 1]; ++q)
        factor_supernode<true>(S, S.wl[q], Fm, dinv, jv, mvals, scale, lu, cb, flag);
    }
    __syncthreads();
    for (int q = 0; q < S.n_top; ++q)
      factor_supernode<false>(S, S.top[q], sm, dinv_all, jv, mvals, scale, lu, cb, flag);
  }
}

// ---------------------------------------------------------------------------------------
// triangular solves
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// forward: r = [b_piv ; 0] + sum_children cv_c ;  L11 x = r_piv ; cv_k = r_upd - L21 x
template <bool WARP>
__device__ void forward_supernode(const LuDev& S, int k, const double* __restrict__ lu,
                                  double* __restrict__ xb, double* __restrict__ cv,
                                  double* __restrict__ rv, double* __restrict__ T) {
  const int s = S.sn_s[k], f = S.sn_f[k], u = f - s, p0 = S.sn_piv0[k];
  const int r = Grp<WARP>::rank(), G = Grp<WARP>::size();
  const double* __restrict__ Lp = lu + S.lp_off[k];
  for (int t = r; t < f; t += G) rv[t] = (t < s) ? xb[p0 + t] : 0.0;
  // stage the s x s diagonal block
  {
    const float invs = 1.0f / static_cast<float>(s);
    const int ss = s * s;
    for (int t = r; t < ss; t += G) {
      const int p = fdiv(t, invs);
      const int i = t - p * s;
      T[t] = Lp[i + p * f];
    }
  }
  Grp<WARP>::sync();
  for (int ci = S.child_ptr[k]; ci < S.child_ptr[k + 1]; ++ci) {
    const int c = S.child[ci];
    const int c0 = S.rel_ptr[c], uc = S.rel_ptr[c + 1] - c0;
    if (uc == 0) continue;
    for (int t = r; t < uc; t += G) rv[S.rel[c0 + t]] += cv[c0 + t];
    Grp<WARP>::sync();
  }
  for (int p = 0; p + 1 < s; ++p) {
    const double xp = rv[p];
    for (int i = p + 1 + r; i < s; i += G) rv[i] = fma(-T[i + p * s], xp, rv[i]);
    Grp<WARP>::sync();
  }
  for (int t = r; t < s; t += G) xb[p0 + t] = rv[t];
  if (u > 0) {
    const int c0 = S.rel_ptr[k];
    for (int i = r; i < u; i += G) {
      double acc = rv[s + i];
      const double* __restrict__ row = Lp + s + i;
      for (int p = 0; p < s; ++p) acc = fma(-row[p * f], rv[p], acc);
      cv[c0 + i] = acc;
    }
  }
  Grp<WARP>::sync();
}

// backward: x_piv = U11^{-1} ( y_piv - U12 x_upd )
template <bool WARP>
__device__ void backward_supernode(const LuDev& S, int k, const double* __restrict__ lu,
                                   double* __restrict__ xb, double* __restrict__ rv,
                                   double* __restrict__ gv, double* __restrict__ T) {
  const int s = S.sn_s[k], f = S.sn_f[k], u = f - s, p0 = S.sn_piv0[k];
  const int r = Grp<WARP>::rank(), G = Grp<WARP>::size();
  const double* __restrict__ Lp = lu + S.lp_off[k];
  const double* __restrict__ Ut = lu + S.up_off[k];
  const int* __restrict__ ids = S.idx + S.idx_ptr[k] + s;
  for (int t = r; t < u; t += G) gv[t] = xb[ids[t]];
  {
    const float invs = 1.0f / static_cast<float>(s);
    const int ss = s * s;
    for (int t = r; t < ss; t += G) {
      const int p = fdiv(t, invs);
      const int i = t - p * s;
      T[t] = Lp[i + p * f];
    }
  }
  Grp<WARP>::sync();
  {
    const int lane = threadIdx.x & 31;
    const int w = WARP ? 0 : (threadIdx.x >> 5);
    const int nw = WARP ? 1 : (blockDim.x >> 5);
    for (int rr = w; rr < s; rr += nw) {
      double acc = 0.0;
      const double* __restrict__ col = Ut + rr * u;
      for (int c = lane; c < u; c += kWarp) acc = fma(col[c], gv[c], acc);
      acc = warp_sum(acc);
      if (lane == 0) rv[rr] = xb[p0 + rr] - acc;
    }
  }
  Grp<WARP>::sync();
  for (int p = s - 1; p >= 0; --p) {
    const double xp = rv[p] / T[p + p * s];
    for (int i = r; i < p; i += G) rv[i] = fma(-T[i + p * s], xp, rv[i]);
    if (r == 0) xb[p0 + p] = xp;
    Grp<WARP>::sync();
  }
}

__global__ void __launch_bounds__(1024, 1)
lu_solve_kernel(const LuDev S, const double* __restrict__ LU, const double* __restrict__ RHS,
                double* __restrict__ X, double* __restrict__ CVW, double* __restrict__ XW,
                int x_in_smem, int mode, const int* __restrict__ active, int B) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5;
  const int n = S.n;
  double* xb = x_in_smem ? sm : XW + static_cast<size_t>(blockIdx.x) * n;
  double* scratch = x_in_smem ? sm + n : sm;
  const int per_warp = 2 * S.f_warp_max + S.s_warp_max * S.s_warp_max;
  double* cv = CVW + static_cast<size_t>(blockIdx.x) * S.cv_size;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (active && !active[b]) continue;
    const double* lu = LU + static_cast<size_t>(b) * S.lu_size;
    const double* rhs = RHS + static_cast<size_t>(b) * n;
    double* x = X + static_cast<size_t>(b) * n;
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) xb[t] = rhs[S.rowp[t]];
    __syncthreads();
    // forward
    if (warp < S.n_warp_lists) {
      double* rv = scratch + warp * per_warp;
      double* T = rv + 2 * S.f_warp_max;
      for (int q = S.wl_ptr[warp]; q < S.wl_ptr[warp + 1]; ++q)
        forward_supernode<true>(S, S.wl[q], lu, xb, cv, rv, T);
    }
    __syncthreads();
    {
      double* rv = scratch;
      double* T = rv + 2 * S.f_top_max;
      for (int q = 0; q < S.n_top; ++q) forward_supernode<false>(S, S.top[q], lu, xb, cv, rv, T);
      // backward
      double* gv = rv + S.f_top_max;
      for (int q = S.n_top - 1; q >= 0; --q)
        backward_supernode<false>(S, S.top[q], lu, xb, rv, gv, T);
    }
    __syncthreads();
    if (warp < S.n_warp_lists) {
      double* rv = scratch + warp * per_warp;
      double* gv = rv + S.f_warp_max;
      double* T = rv + 2 * S.f_warp_max;
      for (int q = S.wl_ptr[warp + 1] - 1; q >= S.wl_ptr[warp]; --q)
        backward_supernode<true>(S, S.wl[q], lu, xb, rv, gv, T);
    }
    __syncthreads();
    if (mode == 0) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) x[S.colp[t]] = xb[t];
    } else {
      for (int t = threadIdx.x; t < n; t += blockDim.x) x[S.colp[t]] -= xb[t];
    }
  }
}

}  // namespace

cudaError_t lu_plan(const LuDev& S, int B, LuLaunchCfg* cfg) {
  int dev = 0, sms = 0, smem_max = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int threads = 32 * (S.n_warp_lists > 0 ? S.n_warp_lists : 1);
  if (S.n_top > 0 && threads < 256) threads = 256;
  if (threads > 1024) threads = 1024;
  cfg->threads = threads;
  const size_t fw2 = static_cast<size_t>(S.f_warp_max) * S.f_warp_max;
  const size_t ftop2 = static_cast<size_t>(S.f_top_max) * S.f_top_max;
  size_t front = S.n_warp_lists * fw2;
  if (ftop2 > front) front = ftop2;
  size_t dinv = static_cast<size_t>(S.n_warp_lists) * S.s_warp_max;
  if (static_cast<size_t>(S.s_top_max) > dinv) dinv = S.s_top_max;
  cfg->smem_factor = (front + dinv) * sizeof(double);
  if (cfg->smem_factor > static_cast<size_t>(smem_max)) return cudaErrorInvalidConfiguration;
  size_t per_warp = 2 * static_cast<size_t>(S.f_warp_max) + static_cast<size_t>(S.s_warp_max) * S.s_warp_max;
  size_t top = 2 * static_cast<size_t>(S.f_top_max) + static_cast<size_t>(S.s_top_max) * S.s_top_max;
  size_t scratch = S.n_warp_lists * per_warp;
  if (top > scratch) scratch = top;
  size_t with_x = (scratch + S.n) * sizeof(double);
  if (with_x <= static_cast<size_t>(smem_max)) {
    cfg->x_in_smem = 1;
    cfg->smem_solve = with_x;
  } else {
    cfg->x_in_smem = 0;
    cfg->smem_solve = scratch * sizeof(double);
    if (cfg->smem_solve > static_cast<size_t>(smem_max)) return cudaErrorInvalidConfiguration;
  }
  e = cudaFuncSetAttribute(lu_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(cfg->smem_factor));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(lu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(cfg->smem_solve));
  if (e != cudaSuccess) return e;
  int occ_f = 1, occ_s = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, lu_factor_kernel, threads, cfg->smem_factor);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, lu_solve_kernel, threads, cfg->smem_solve);
  if (occ_f < 1) occ_f = 1;
  if (occ_s < 1) occ_s = 1;
  cfg->grid_factor = sms * occ_f;
  cfg->grid_solve = sms * occ_s;
  if (cfg->grid_factor > B) cfg->grid_factor = B > 0 ? B : 1;
  if (cfg->grid_solve > B) cfg->grid_solve = B > 0 ? B : 1;
  return cudaSuccess;
}

cudaError_t lu_factor_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* JV, double* LU,
                             double* CBW, const double* mvals, const double* scale_b,
                             double scale_const, const int* active, int* flags, int B,
                             cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  lu_factor_kernel<<<cfg.grid_factor, cfg.threads, cfg.smem_factor, st>>>(
      S, JV, LU, CBW, mvals, scale_b, scale_const, active, flags, B);
  return cudaGetLastError();
}

cudaError_t lu_solve_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* LU,
                            const double* RHS, double* X, double* CVW, double* XW, int mode,
                            const int* active, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  lu_solve_kernel<<<cfg.grid_solve, cfg.threads, cfg.smem_solve, st>>>(
      S, LU, RHS, X, CVW, XW, cfg.x_in_smem, mode, active, B);
  return cudaGetLastError();
}

}  // namespace szb
