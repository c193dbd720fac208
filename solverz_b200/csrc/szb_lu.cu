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
//   * three levels of the elimination tree run with three kinds of thread groups:
//       warp  : every warp walks its own list of small subtrees, front in a warp-private
//               shared-memory buffer, __syncwarp only;
//       group : a few warps (named barrier) share one medium front; groups run independent
//               subtrees concurrently;
//       CTA   : the few big fronts at the top use all threads.
//     Inside a front: lane = row (conflict-free column-major shared memory with an odd leading
//     dimension), warps stride over columns, right-looking rank-1 updates, one barrier/pivot;
//   * the right-hand side can ride along as an extra front column, which fuses the forward
//     substitution L y = b into the factorisation (Newton never needs L again);
//   * contribution blocks travel through a per-CTA stack in global memory (L2 resident, reused
//     across scenarios by the persistent CTA); factors are written exactly once;
//   * triangular solves follow the same tree; the forward sweep passes contribution vectors up
//     the tree (no atomics => bitwise reproducible), the backward sweep only reads ancestors.
//   No tensor cores: fp64 FMA fed from shared memory; nothing here is a large dense contraction.
#include "szb_lu.cuh"

#include <cstdio>

namespace szb {

namespace {

constexpr int kWarp = 32;

// t / m for 0 <= t < 2^22 without an integer division: (t + 0.5)/m is never within 0.5/m of an
// integer, far more than the 2^-23 relative rounding of the float product
__device__ __forceinline__ int fdiv(int t, float invm) {
  return __float2int_rz((static_cast<float>(t) + 0.5f) * invm);
}

// ---- thread groups ------------------------------------------------------------------------
struct WarpGrp {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int warp() const { return 0; }
  __device__ __forceinline__ int nwarps() const { return 1; }
  __device__ __forceinline__ int rank() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int size() const { return kWarp; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
struct NamedGrp {  // `nw` consecutive warps, hardware named barrier `id` (1..15)
  int id, nw, w0;
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int warp() const { return (threadIdx.x >> 5) - w0; }
  __device__ __forceinline__ int nwarps() const { return nw; }
  __device__ __forceinline__ int rank() const { return threadIdx.x - (w0 << 5); }
  __device__ __forceinline__ int size() const { return nw << 5; }
  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nw << 5) : "memory");
  }
};
struct BlockGrp {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
  __device__ __forceinline__ int nwarps() const { return blockDim.x >> 5; }
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

struct FactorArgs {
  const double* __restrict__ jv;     // matrix values of this scenario (CSC slots)
  const double* __restrict__ mvals;  // optional mass values (shared)
  const double* __restrict__ smask;  // optional 0/1 mask per slot applied to jv (shared)
  double scale;
  const double* __restrict__ rhs;    // optional right-hand side of this scenario (original order)
  double* __restrict__ yout;         // forward-substituted rhs in pivot order (when rhs != null)
  double* __restrict__ lu;           // factor storage of this scenario
  double* __restrict__ cb;           // contribution workspace of this CTA
  int store_l;                       // write the L21 rows of the panels
  int* flag;
};

__device__ __forceinline__ void check_pivot(const double piv, int* flag) {
  if (!(fabs(piv) > 0.0) || isinf(piv)) *flag = 2;
}

// ---- single warp, fronts with at most 32 rows: right-looking rank-1 steps, lane = row.  When the
// trailing block has m <= 16 (<= 8) rows, two (four) columns are updated per warp instruction so that
// all lanes stay busy on the tiny fronts that make up most of the tree.
__device__ __forceinline__ void eliminate_warp(double* __restrict__ Fm, const int s, const int f,
                                               const int ncol, const int ld, const int lane, int* flag) {
  for (int p = 0; p < s; ++p) {
    const int m = f - p - 1;
    const int nc = ncol - p - 1;
    double* __restrict__ colp = Fm + (p + 1) + p * ld;
    const double piv = colp[-1];
    const double inv = 1.0 / piv;
    if (lane == 0) check_pivot(piv, flag);
    const int sh = m <= 8 ? 3 : (m <= 16 ? 4 : 5);     // rows per column slot = 1 << sh
    const int i = lane & ((1 << sh) - 1);
    const int cj = lane >> sh;                          // column slot of this lane
    const int cpw = 32 >> sh;                           // columns per warp instruction
    const bool row_ok = i < m;
    double l = 0.0;
    if (row_ok) l = colp[i] * inv;
    __syncwarp();
    if (cj == 0 && row_ok) colp[i] = l;                 // L stored scaled
    double* __restrict__ col = Fm + (p + 1) + (p + 1 + cj) * ld;
    const int step = cpw * ld;
    for (int j = cj; j < nc; j += cpw, col += step) {
      const double u = col[-1];
      if (row_ok) col[i] = fma(-l, u, col[i]);
    }
    __syncwarp();
  }
}

// ---- several warps: blocked right-looking elimination with look-ahead, NB = 4 pivots per panel.
//
//   panel t = pivots [p0, p0+nb).  After its factorisation  A22 -= L21 U12  with U12 = L11^{-1} A12
//   is applied as  A22 -= Lt A12,  Lt = L21 L11^{-1}:  the update then only reads the *unsolved* top
//   rows A12, so no per-column triangular solve sits in the inner loop; the rows of U (and the rhs
//   entries of the pivot rows) are solved once, after the last panel (finalize_u).
//   Warp 0 is the panel warp: in iteration t it first updates the columns of panel t+1 with panel
//   t, factors them in registers (pivots exchanged with shuffles), and publishes L (in place) and
//   Lt (double-buffered side panel) while the other warps update the rest of the trailing block --
//   the panel factorisation is off the critical path and there is ONE barrier per 4 pivots.
//   Workers split the trailing block 2-D: one 32-row chunk per warp, columns strided.
#ifdef SZB_PHASE_TIMING
__device__ long long g_dbg[16];
#define SZB_T(i) if (nw > 8 && r == 0 && blockIdx.x == 0) { long long now_ = clock64(); g_dbg[i] += now_ - tprev_; tprev_ = now_; }
#else
#define SZB_T(i)
#endif

constexpr int kNB = 4;
constexpr int kLtStride = 128;   // rows per Lt column (fronts of multi-warp groups have <= 128 rows)

// Factor the nb columns held in P (rows r = p0 + lane + 32c), write L in place, Lt to `lt`.
template <int NCH>
__device__ __forceinline__ void panel_finish(double (&P)[NCH][kNB], double* __restrict__ Fm,
                                             double* __restrict__ lt, const int p0, const int nb,
                                             const int f, const int ld, const int lane, int* flag) {
#pragma unroll
  for (int k = 0; k < kNB; ++k) {
    if (k < nb) {
      const double piv = __shfl_sync(0xffffffffu, P[0][k], k);
      const double inv = 1.0 / piv;
      if (lane == 0) check_pivot(piv, flag);
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c > 0 || lane > k) P[c][k] *= inv;
#pragma unroll
      for (int k2 = k + 1; k2 < kNB; ++k2) {
        const double u = __shfl_sync(0xffffffffu, P[0][k2], k);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (c > 0 || lane > k) P[c][k2] = fma(-P[c][k], u, P[c][k2]);
      }
    }
  }
  // L (scaled) and the U11 block go back in place
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = p0 + lane + 32 * c;
#pragma unroll
    for (int k = 0; k < kNB; ++k)
      if (i < f && k < nb) Fm[i + (p0 + k) * ld] = P[c][k];
  }
  // Lt = L21 L11^{-1} for the rows below the panel (x L11 = l, L11 unit lower)
  const double L10 = __shfl_sync(0xffffffffu, P[0][0], 1);
  const double L20 = __shfl_sync(0xffffffffu, P[0][0], 2), L21 = __shfl_sync(0xffffffffu, P[0][1], 2);
  const double L30 = __shfl_sync(0xffffffffu, P[0][0], 3), L31 = __shfl_sync(0xffffffffu, P[0][1], 3),
               L32 = __shfl_sync(0xffffffffu, P[0][2], 3);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = lane + 32 * c - nb;            // row index relative to r0 = p0 + nb
    if (i >= 0 && p0 + nb + i < f) {
      const double x3 = nb > 3 ? P[c][3] : 0.0;
      const double x2 = nb > 2 ? fma(-x3, L32, P[c][2]) : 0.0;
      const double x1 = nb > 1 ? fma(-x3, L31, fma(-x2, L21, P[c][1])) : 0.0;
      const double x0 = fma(-x3, L30, fma(-x2, L20, fma(-x1, L10, P[c][0])));
      lt[i] = x0;
      lt[kLtStride + i] = x1;
      lt[2 * kLtStride + i] = x2;
      lt[3 * kLtStride + i] = x3;
    }
  }
}

// first panel of a supernode: load and factor
template <int NCH>
__device__ __forceinline__ void panel_first(double* __restrict__ Fm, double* __restrict__ lt, const int nb,
                                            const int f, const int ld, const int lane, int* flag) {
  double P[NCH][kNB];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = lane + 32 * c;
#pragma unroll
    for (int k = 0; k < kNB; ++k) P[c][k] = (i < f && k < nb) ? Fm[i + k * ld] : 0.0;
  }
  panel_finish<NCH>(P, Fm, lt, 0, nb, f, ld, lane, flag);
}

// look-ahead: update the nb1 columns of the next panel (rows >= r0) with panel [p0, p0+nb) using
// its Lt in `lt_cur`, then factor them and publish Lt of the next panel in `lt_next`.
template <int NCH>
__device__ __forceinline__ void panel_next(double* __restrict__ Fm, const double* __restrict__ lt_cur,
                                           double* __restrict__ lt_next, const int p0, const int nb,
                                           const int nb1, const int f, const int ld, const int lane,
                                           int* flag) {
  const int r0 = p0 + nb;
  double P[NCH][kNB];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = r0 + lane + 32 * c;
#pragma unroll
    for (int q = 0; q < kNB; ++q) P[c][q] = (i < f && q < nb1) ? Fm[i + (r0 + q) * ld] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < kNB; ++k) {
    if (k < nb) {
      double top[kNB];
#pragma unroll
      for (int q = 0; q < kNB; ++q) top[q] = q < nb1 ? Fm[(p0 + k) + (r0 + q) * ld] : 0.0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int i = lane + 32 * c;
        const double l = (r0 + i < f) ? lt_cur[k * kLtStride + i] : 0.0;
#pragma unroll
        for (int q = 0; q < kNB; ++q) P[c][q] = fma(-l, top[q], P[c][q]);
      }
    }
  }
  panel_finish<NCH>(P, Fm, lt_next, r0, nb1, f, ld, lane, flag);
}

// worker: all NCH row chunks of columns jbeg + cg, jbeg + cg + CW, ...  The four unsolved top values
// of a column are fetched once (broadcast) and reused for every chunk; multipliers live in registers.
template <int NCH>
__device__ __forceinline__ void trailing_worker(double* __restrict__ Fm, const double* __restrict__ lt,
                                                const int p0, const int nb, const int f, const int ld,
                                                const int jbeg, const int ncol, const int cg, const int CW,
                                                const int lane) {
  const int r0 = p0 + nb;
  const int m = f - r0;
  double l[NCH][kNB];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = lane + 32 * c;
#pragma unroll
    for (int k = 0; k < kNB; ++k) l[c][k] = (i < m && k < nb) ? lt[k * kLtStride + i] : 0.0;
  }
  const int k1 = nb > 1 ? 1 : 0, k2 = nb > 2 ? 2 : 0, k3 = nb > 3 ? 3 : 0;   // stay inside the panel rows
  double* __restrict__ col = Fm + (jbeg + cg) * ld;
  const int step = CW * ld;
  for (int j = jbeg + cg; j < ncol; j += CW, col += step) {
    const double a0 = col[p0], a1 = col[p0 + k1], a2 = col[p0 + k2], a3 = col[p0 + k3];
    double a[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int i = lane + 32 * c;
      a[c] = (c + 1 < NCH || i < m) ? col[r0 + i] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      a[c] = fma(-l[c][0], a0, a[c]);
      a[c] = fma(-l[c][1], a1, a[c]);
      a[c] = fma(-l[c][2], a2, a[c]);
      a[c] = fma(-l[c][3], a3, a[c]);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int i = lane + 32 * c;
      if (c + 1 < NCH || i < m) col[r0 + i] = a[c];
    }
  }
}

// deferred triangular solves of the U rows: for every panel and every column right of it
//   u = L11^{-1} a   (4 values per column, lanes over columns, warps over panels)
__device__ __forceinline__ void finalize_u(double* __restrict__ Fm, const int s, const int ncol, const int ld,
                                           const int lane, const int w, const int nw) {
  const int npan = (s + kNB - 1) / kNB;
  for (int t = w; t < npan; t += nw) {
    const int p0 = t * kNB;
    const int nb = min(kNB, s - p0);
    if (nb < 2) continue;
    const double* __restrict__ blk = Fm + p0 + p0 * ld;
    const double L10 = blk[1];
    const double L20 = nb > 2 ? blk[2] : 0.0, L21 = nb > 2 ? blk[2 + ld] : 0.0;
    const double L30 = nb > 3 ? blk[3] : 0.0, L31 = nb > 3 ? blk[3 + ld] : 0.0, L32 = nb > 3 ? blk[3 + 2 * ld] : 0.0;
    for (int j = p0 + nb + lane; j < ncol; j += kWarp) {
      double* __restrict__ col = Fm + p0 + j * ld;
      const double u0 = col[0];
      const double u1 = fma(-L10, u0, col[1]);
      col[1] = u1;
      if (nb > 2) {
        const double u2 = fma(-L21, u1, fma(-L20, u0, col[2]));
        col[2] = u2;
        if (nb > 3) col[3] = fma(-L32, u2, fma(-L31, u1, fma(-L30, u0, col[3])));
      }
    }
  }
}

template <class G>
__device__ __forceinline__ void eliminate_blocked(const G g, double* __restrict__ Fm, double* __restrict__ ltbuf,
                                                  const int s, const int f, const int ncol, const int ld,
                                                  int* flag) {
  const int lane = g.lane(), w = g.warp(), nw = g.nwarps();
  const int npan = (s + kNB - 1) / kNB;
#ifdef SZB_PHASE_TIMING
  const int r = g.rank();
  long long tprev_ = clock64();
#endif
  if (w == 0) {
    const int nb = min(kNB, s);
    switch ((f + 31) >> 5) {
      case 1: panel_first<1>(Fm, ltbuf, nb, f, ld, lane, flag); break;
      case 2: panel_first<2>(Fm, ltbuf, nb, f, ld, lane, flag); break;
      case 3: panel_first<3>(Fm, ltbuf, nb, f, ld, lane, flag); break;
      default: panel_first<4>(Fm, ltbuf, nb, f, ld, lane, flag); break;
    }
  }
  g.sync();
  SZB_T(4)
  for (int t = 0; t < npan; ++t) {
    const int p0 = t * kNB;
    const int nb = min(kNB, s - p0);
    const int r0 = p0 + nb;
    const bool has_next = t + 1 < npan;
    const int nb1 = has_next ? min(kNB, s - r0) : 0;
    const double* lt_cur = ltbuf + (t & 1) * (kNB * kLtStride);
    double* lt_next = ltbuf + ((t + 1) & 1) * (kNB * kLtStride);
    if (has_next && w == 0) {
      switch ((f - r0 + 31) >> 5) {
        case 1: panel_next<1>(Fm, lt_cur, lt_next, p0, nb, nb1, f, ld, lane, flag); break;
        case 2: panel_next<2>(Fm, lt_cur, lt_next, p0, nb, nb1, f, ld, lane, flag); break;
        case 3: panel_next<3>(Fm, lt_cur, lt_next, p0, nb, nb1, f, ld, lane, flag); break;
        default: panel_next<4>(Fm, lt_cur, lt_next, p0, nb, nb1, f, ld, lane, flag); break;
      }
    } else {
      const int W = has_next ? nw - 1 : nw;          // worker warps
      const int ww = has_next ? w - 1 : w;
      const int m = f - r0;
      const int jbeg = r0 + nb1;
      if (jbeg < ncol) {
        switch ((m + 31) >> 5) {                      // row chunks (0: only the rhs column is left)
          case 0: break;
          case 1: trailing_worker<1>(Fm, lt_cur, p0, nb, f, ld, jbeg, ncol, ww, W, lane); break;
          case 2: trailing_worker<2>(Fm, lt_cur, p0, nb, f, ld, jbeg, ncol, ww, W, lane); break;
          case 3: trailing_worker<3>(Fm, lt_cur, p0, nb, f, ld, jbeg, ncol, ww, W, lane); break;
          default: trailing_worker<4>(Fm, lt_cur, p0, nb, f, ld, jbeg, ncol, ww, W, lane); break;
        }
      }
    }
    SZB_T(5)
    g.sync();
    SZB_T(6)
  }
  finalize_u(Fm, s, ncol, ld, lane, w, nw);
  g.sync();
  SZB_T(7)
}

// ---------------------------------------------------------------------------------------
// numeric factorisation of one supernode; the front is column-major with leading dimension
// ld = f | 1 and has f (+1 when a rhs rides along) columns.  Fronts of multi-warp groups are
// limited to 128 rows (4 row chunks per lane); lu_plan() rejects larger ones.
// ---------------------------------------------------------------------------------------
template <class G>
__device__ void factor_supernode(const G g, const LuDev& S, const int k, double* __restrict__ Fm,
                                 double* __restrict__ ltbuf, const FactorArgs& A) {
  const int s = S.sn_s[k], f = S.sn_f[k], u = f - s, ld = f | 1;
#ifdef SZB_PHASE_TIMING
  long long tprev_ = clock64();
#endif
  const int lane = g.lane(), w = g.warp(), nw = g.nwarps(), r = g.rank(), G_ = g.size();
  const int with_rhs = A.rhs != nullptr;
  const int ncol = f + with_rhs;
  // 1. clear
  for (int t = r; t < ld * ncol; t += G_) Fm[t] = 0.0;
  g.sync();
  // 2. original entries (+ right-hand side of the pivot rows).  Global loads are issued in batches
  // of four before the dependent shared-memory stores so that their latencies overlap.
  {
    const int a0 = S.a_ptr[k], a1 = S.a_ptr[k + 1];
    for (int e = a0 + r; e < a1; e += 4 * G_) {
      int src[4], dst[4];
      double v[4], mv[4], mk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ee = e + q * G_;
        const bool ok = ee < a1;
        src[q] = ok ? S.a_src[ee] : -1;
        dst[q] = ok ? S.a_dst[ee] : 0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = src[q] >= 0 ? A.jv[src[q]] : 0.0;
        mv[q] = (A.mvals && src[q] >= 0) ? A.mvals[src[q]] : 0.0;
        mk[q] = (A.smask && src[q] >= 0) ? A.smask[src[q]] : 1.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (src[q] >= 0) Fm[dst[q]] = fma(A.scale * mk[q], v[q], mv[q]);
    }
    if (with_rhs) {
      const int p0 = S.sn_piv0[k];
      for (int t = r; t < s; t += G_) Fm[t + f * ld] = A.rhs[S.rowp[p0 + t]];
    }
  }
  g.sync();
  SZB_T(0)
  // 3. extend-add of the children's contribution blocks (u_c x (u_c + 1), column-major); the
  // destination of every element was computed on the host (ea_dst).  Loads batched as above.
  for (int ci = S.child_ptr[k]; ci < S.child_ptr[k + 1]; ++ci) {
    const int c = S.child[ci];
    const int uc = S.sn_f[c] - S.sn_s[c];
    if (uc == 0) continue;
    const int* __restrict__ dstp = S.ea_dst + S.ea_ptr[c];
    const double* __restrict__ C = A.cb + S.cb_off[c];
    const int cnt = uc * (uc + with_rhs);
    for (int t = r; t < cnt; t += 4 * G_) {
      int d[4];
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int tt = t + q * G_;
        d[q] = tt < cnt ? dstp[tt] : -1;
        v[q] = tt < cnt ? C[tt] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (d[q] >= 0) Fm[d[q]] += v[q];
    }
    g.sync();
  }
  SZB_T(1)
  // 4. elimination of the s pivots
  if (nw == 1) eliminate_warp(Fm, s, f, ncol, ld, lane, A.flag);
  else eliminate_blocked(g, Fm, ltbuf, s, f, ncol, ld, A.flag);
  SZB_T(2)
  // 5. write-out
  {
    double* __restrict__ Lp = A.lu + S.lp_off[k];           // f x s, column-major, ld = f
    const int rows = A.store_l ? f : s;
    if (nw == 1) {
      const float invr = 1.0f / static_cast<float>(rows);
      for (int t = r; t < rows * s; t += kWarp) {
        const int p = fdiv(t, invr);
        const int i = t - p * rows;
        Lp[i + p * f] = Fm[i + p * ld];
      }
    } else {
      for (int p = w; p < s; p += nw)
        for (int i = lane; i < rows; i += kWarp) Lp[i + p * f] = Fm[i + p * ld];
    }
    if (u > 0) {
      double* __restrict__ Up = A.lu + S.up_off[k];         // U12 rows: Up[c + r*u] = U12[r][c]
      double* __restrict__ C = A.cb + S.cb_off[k];          // u x (u+1), column-major
      const int nc = u + with_rhs;
      if (nw == 1) {
        const float invu = 1.0f / static_cast<float>(u);
        for (int t = r; t < u * s; t += kWarp) {
          const int rr = fdiv(t, invu);
          const int c = t - rr * u;
          Up[t] = Fm[rr + (s + c) * ld];
        }
        for (int t = r; t < u * nc; t += kWarp) {
          const int j = fdiv(t, invu);
          const int i = t - j * u;
          C[t] = Fm[(s + i) + (s + j) * ld];
        }
      } else {
        for (int rr = w; rr < s; rr += nw)
          for (int c = lane; c < u; c += kWarp) Up[c + rr * u] = Fm[rr + (s + c) * ld];
        for (int j = w; j < nc; j += nw)
          for (int i = lane; i < u; i += kWarp) C[i + j * u] = Fm[(s + i) + (s + j) * ld];
      }
    }
    if (with_rhs) {
      const int p0 = S.sn_piv0[k];
      for (int t = r; t < s; t += G_) A.yout[p0 + t] = Fm[t + f * ld];
    }
  }
  g.sync();
  SZB_T(3)
}

__global__ void __launch_bounds__(1024, 1)
lu_factor_kernel(const LuDev S, const double* __restrict__ JV, double* __restrict__ LU,
                 double* __restrict__ CBW, const double* __restrict__ mvals,
                 const double* __restrict__ slot_mask,
                 const double* __restrict__ scale_b, double scale_const,
                 const double* __restrict__ RHS, double* __restrict__ YOUT, int store_l,
                 const int* __restrict__ active, int* __restrict__ flags, int B) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5;
  const int wstride = (S.f_warp_max | 1) * (S.f_warp_max + 1);
  const int gfront = (S.f_group_max | 1) * (S.f_group_max + 1);
  const int gstride = gfront + 2 * kNB * kLtStride;
  const int tfront = (S.f_top_max | 1) * (S.f_top_max + 1);
  const int wpg = (blockDim.x >> 5) / max(S.n_group_lists, 1);   // warps per group
#ifdef SZB_PHASE_TIMING
  long long tph[4] = {0, 0, 0, 0};
#endif
  FactorArgs A;
  A.mvals = mvals;
  A.smask = slot_mask;
  A.store_l = store_l;
  A.cb = CBW + static_cast<size_t>(blockIdx.x) * S.cb_size;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (active && !active[b]) continue;
    A.jv = JV + static_cast<size_t>(b) * S.nnz_a;
    A.lu = LU + static_cast<size_t>(b) * S.lu_size;
    A.scale = scale_b ? scale_b[b] : scale_const;
    A.rhs = RHS ? RHS + static_cast<size_t>(b) * S.n : nullptr;
    A.yout = YOUT ? YOUT + static_cast<size_t>(b) * S.n : nullptr;
    A.flag = flags + b;
#ifdef SZB_PHASE_TIMING
    long long t0 = clock64();
#endif
    // level 0: warps
    if (warp < S.n_warp_lists) {
      double* Fm = sm + warp * wstride;
      const WarpGrp g;
      for (int q = S.wl_ptr[warp]; q < S.wl_ptr[warp + 1]; ++q) factor_supernode(g, S, S.wl[q], Fm, nullptr, A);
    }
    __syncthreads();
#ifdef SZB_PHASE_TIMING
    long long t1 = clock64();
#endif
    // level 1: groups of warps
    if (S.n_group_lists > 0) {
      const int gi = warp / wpg;
      if (gi < S.n_group_lists) {
        double* Fm = sm + gi * gstride;
        NamedGrp g;
        g.id = 1 + gi;
        g.nw = wpg;
        g.w0 = gi * wpg;
        for (int q = S.gl_ptr[gi]; q < S.gl_ptr[gi + 1]; ++q)
          factor_supernode(g, S, S.gl[q], Fm, Fm + gfront, A);
      }
      __syncthreads();
    }
#ifdef SZB_PHASE_TIMING
    long long t2 = clock64();
#endif
    // level 2: the whole CTA
    {
      const BlockGrp g;
      for (int q = 0; q < S.n_top; ++q) factor_supernode(g, S, S.top[q], sm, sm + tfront, A);
    }
#ifdef SZB_PHASE_TIMING
    long long t3 = clock64();
    tph[0] += t1 - t0; tph[1] += t2 - t1; tph[2] += t3 - t2; tph[3] += 1;
#endif
  }
#ifdef SZB_PHASE_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("[lu_factor phases] warp %lld group %lld top %lld cycles/scenario (%lld scenarios) | top: asm %lld ea %lld elim %lld write %lld\n",
           tph[0] / max(tph[3], 1LL), tph[1] / max(tph[3], 1LL), tph[2] / max(tph[3], 1LL), tph[3],
           g_dbg[0] / max(tph[3], 1LL), g_dbg[1] / max(tph[3], 1LL), g_dbg[2] / max(tph[3], 1LL), g_dbg[3] / max(tph[3], 1LL));
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("   elim split (thread 0): first-panel %lld  panel_next/own work %lld  barrier wait %lld  finalize %lld\n", g_dbg[4] / max(tph[3], 1LL),
           g_dbg[5] / max(tph[3], 1LL), g_dbg[6] / max(tph[3], 1LL), g_dbg[7] / max(tph[3], 1LL));
  if (threadIdx.x == 0 && blockIdx.x == 0) for (int i = 0; i < 8; ++i) g_dbg[i] = 0;
#endif
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
// rv: f doubles of scratch, T: s*s doubles of scratch (diagonal block staged once)
template <class G>
__device__ void forward_supernode(const G g, const LuDev& S, const int k,
                                  const double* __restrict__ lu, double* __restrict__ xb,
                                  double* __restrict__ cv, double* __restrict__ rv,
                                  double* __restrict__ T) {
  const int s = S.sn_s[k], f = S.sn_f[k], u = f - s, p0 = S.sn_piv0[k];
  const int lane = g.lane(), w = g.warp(), nw = g.nwarps(), r = g.rank(), G_ = g.size();
  const double* __restrict__ Lp = lu + S.lp_off[k];
  for (int t = r; t < f; t += G_) rv[t] = (t < s) ? xb[p0 + t] : 0.0;
  for (int p = w; p < s; p += nw)
    for (int i = lane; i < s; i += kWarp) T[i + p * s] = Lp[i + p * f];
  g.sync();
  for (int ci = S.child_ptr[k]; ci < S.child_ptr[k + 1]; ++ci) {
    const int c = S.child[ci];
    const int c0 = S.rel_ptr[c], uc = S.rel_ptr[c + 1] - c0;
    if (uc == 0) continue;
    for (int t = r; t < uc; t += G_) rv[S.rel[c0 + t]] += cv[c0 + t];
    g.sync();
  }
  for (int p = 0; p + 1 < s; ++p) {
    const double xp = rv[p];
    for (int i = p + 1 + r; i < s; i += G_) rv[i] = fma(-T[i + p * s], xp, rv[i]);
    g.sync();
  }
  for (int t = r; t < s; t += G_) xb[p0 + t] = rv[t];
  if (u > 0) {
    const int c0 = S.rel_ptr[k];
    for (int i = r; i < u; i += G_) {
      double acc = rv[s + i];
      const double* __restrict__ row = Lp + s + i;
      for (int p = 0; p < s; ++p) acc = fma(-row[p * f], rv[p], acc);
      cv[c0 + i] = acc;
    }
  }
  g.sync();
}

// backward: x_piv = U11^{-1} ( y_piv - U12 x_upd )
// rv: s doubles, gv: u doubles, T: s*s doubles
template <class G>
__device__ void backward_supernode(const G g, const LuDev& S, const int k,
                                   const double* __restrict__ lu, double* __restrict__ xb,
                                   double* __restrict__ rv, double* __restrict__ gv,
                                   double* __restrict__ T) {
  const int s = S.sn_s[k], f = S.sn_f[k], u = f - s, p0 = S.sn_piv0[k];
  const int lane = g.lane(), w = g.warp(), nw = g.nwarps(), r = g.rank(), G_ = g.size();
  const double* __restrict__ Lp = lu + S.lp_off[k];
  const double* __restrict__ Ut = lu + S.up_off[k];
  const int* __restrict__ ids = S.idx + S.idx_ptr[k] + s;
  for (int t = r; t < u; t += G_) gv[t] = xb[ids[t]];
  // stage the upper triangle (incl. diagonal) of the s x s block: column p, rows 0..p
  for (int p = w; p < s; p += nw)
    for (int i = lane; i <= p; i += kWarp) T[i + p * s] = Lp[i + p * f];
  g.sync();
  for (int rr = w; rr < s; rr += nw) {
    double acc = 0.0;
    const double* __restrict__ row = Ut + rr * u;
    for (int c = lane; c < u; c += kWarp) acc = fma(row[c], gv[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) rv[rr] = xb[p0 + rr] - acc;
  }
  g.sync();
  if (s <= kWarp) {
    // one warp finishes the small triangle with shuffles, no further barriers
    if (w == 0) {
      double mine = (lane < s) ? rv[lane] : 0.0;
      for (int p = s - 1; p >= 0; --p) {
        const double xp = __shfl_sync(0xffffffffu, mine, p) / T[p + p * s];
        if (lane == p) mine = xp;
        else if (lane < p) mine = fma(-T[lane + p * s], xp, mine);
      }
      if (lane < s) xb[p0 + lane] = mine;
    }
  } else {
    for (int p = s - 1; p >= 0; --p) {
      const double xp = rv[p] / T[p + p * s];
      for (int i = r; i < p; i += G_) rv[i] = fma(-T[i + p * s], xp, rv[i]);
      if (r == 0) xb[p0 + p] = xp;
      g.sync();
    }
  }
  g.sync();
}

struct SolveSmem {
  int wstride, gstride;
  int w_rv, w_T, g_rv, g_T, t_rv, t_T;
};

__device__ __forceinline__ SolveSmem solve_layout(const LuDev& S) {
  SolveSmem L;
  L.w_rv = 2 * S.f_warp_max;
  L.w_T = S.s_warp_max * S.s_warp_max;
  L.wstride = L.w_rv + L.w_T;
  L.g_rv = 2 * S.f_group_max;
  L.g_T = S.s_group_max * S.s_group_max;
  L.gstride = L.g_rv + L.g_T;
  L.t_rv = 2 * S.f_top_max;
  L.t_T = S.s_top_max * S.s_top_max;
  return L;
}

// do_fwd: run the forward sweep (RHS in original order); otherwise RHS already holds y in pivot
// order (produced by the fused factorisation).
__global__ void __launch_bounds__(1024, 1)
lu_solve_kernel(const LuDev S, const double* __restrict__ LU, const double* __restrict__ RHS,
                double* __restrict__ X, double* __restrict__ CVW, double* __restrict__ XW,
                int x_in_smem, int do_fwd, int mode, const int* __restrict__ active, int B) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5;
  const int n = S.n;
  double* xb = x_in_smem ? sm : XW + static_cast<size_t>(blockIdx.x) * n;
  double* scratch = x_in_smem ? sm + n : sm;
  const SolveSmem L = solve_layout(S);
  const int wpg = (blockDim.x >> 5) / max(S.n_group_lists, 1);
  double* cv = CVW + static_cast<size_t>(blockIdx.x) * S.cv_size;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (active && !active[b]) continue;
    const double* lu = LU + static_cast<size_t>(b) * S.lu_size;
    const double* rhs = RHS + static_cast<size_t>(b) * n;
    double* x = X + static_cast<size_t>(b) * n;
    __syncthreads();
    if (do_fwd) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) xb[t] = rhs[S.rowp[t]];
    } else {
      for (int t = threadIdx.x; t < n; t += blockDim.x) xb[t] = rhs[t];
    }
    __syncthreads();
    if (do_fwd) {
      if (warp < S.n_warp_lists) {
        double* rv = scratch + warp * L.wstride;
        const WarpGrp g;
        for (int q = S.wl_ptr[warp]; q < S.wl_ptr[warp + 1]; ++q)
          forward_supernode(g, S, S.wl[q], lu, xb, cv, rv, rv + L.w_rv);
      }
      __syncthreads();
      if (S.n_group_lists > 0) {
        const int gi = warp / wpg;
        if (gi < S.n_group_lists) {
          double* rv = scratch + gi * L.gstride;
          NamedGrp g;
          g.id = 1 + gi; g.nw = wpg; g.w0 = gi * wpg;
          for (int q = S.gl_ptr[gi]; q < S.gl_ptr[gi + 1]; ++q)
            forward_supernode(g, S, S.gl[q], lu, xb, cv, rv, rv + L.g_rv);
        }
        __syncthreads();
      }
      {
        const BlockGrp g;
        for (int q = 0; q < S.n_top; ++q)
          forward_supernode(g, S, S.top[q], lu, xb, cv, scratch, scratch + L.t_rv);
      }
    }
    // backward: top, groups, warps
    {
      const BlockGrp g;
      for (int q = S.n_top - 1; q >= 0; --q)
        backward_supernode(g, S, S.top[q], lu, xb, scratch, scratch + S.f_top_max, scratch + L.t_rv);
    }
    __syncthreads();
    if (S.n_group_lists > 0) {
      const int gi = warp / wpg;
      if (gi < S.n_group_lists) {
        double* rv = scratch + gi * L.gstride;
        NamedGrp g;
        g.id = 1 + gi; g.nw = wpg; g.w0 = gi * wpg;
        for (int q = S.gl_ptr[gi + 1] - 1; q >= S.gl_ptr[gi]; --q)
          backward_supernode(g, S, S.gl[q], lu, xb, rv, rv + S.f_group_max, rv + L.g_rv);
      }
      __syncthreads();
    }
    if (warp < S.n_warp_lists) {
      double* rv = scratch + warp * L.wstride;
      const WarpGrp g;
      for (int q = S.wl_ptr[warp + 1] - 1; q >= S.wl_ptr[warp]; --q)
        backward_supernode(g, S, S.wl[q], lu, xb, rv, rv + S.f_warp_max, rv + L.w_rv);
    }
    __syncthreads();
    if (mode == 0) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) x[S.colp[t]] = xb[t];
    } else {
      for (int t = threadIdx.x; t < n; t += blockDim.x) x[S.colp[t]] -= xb[t];
    }
  }
}

size_t smax(size_t a, size_t b) { return a > b ? a : b; }

}  // namespace

cudaError_t lu_plan(const LuDev& S, int B, LuLaunchCfg* cfg) {
  int dev = 0, sms = 0, smem_max = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  // threads: one warp per warp list, and a whole number of warps per group
  int warps = S.n_warp_lists > 0 ? S.n_warp_lists : 1;
  if (S.n_group_lists > 0) {
    const int per = (warps + S.n_group_lists - 1) / S.n_group_lists;
    const int wpg = per < 2 ? 2 : per;
    warps = wpg * S.n_group_lists;
  }
  if (S.n_top > 0 && warps < 8) warps = S.n_group_lists > 0 ? ((8 + S.n_group_lists - 1) / S.n_group_lists) * S.n_group_lists : 8;
  if (warps > 32) return cudaErrorInvalidConfiguration;
  cfg->threads = 32 * warps;
  if (S.f_group_max > 128 || S.f_top_max > 128) return cudaErrorInvalidConfiguration;  // 4 row chunks per lane
  if (S.f_warp_max > 32) return cudaErrorInvalidConfiguration;
  auto front = [](int f) { return static_cast<size_t>(f | 1) * (f + 1); };
  const size_t ltb = 2 * 4 * 128;   // two Lt side panels (kNB x kLtStride doubles)
  size_t fsm = S.n_warp_lists * front(S.f_warp_max);
  fsm = smax(fsm, S.n_group_lists * (front(S.f_group_max) + ltb));
  fsm = smax(fsm, S.n_top > 0 ? front(S.f_top_max) + ltb : 0);
  cfg->smem_factor = fsm * sizeof(double);
  if (cfg->smem_factor > static_cast<size_t>(smem_max)) return cudaErrorInvalidConfiguration;
  size_t ssm = S.n_warp_lists * (2 * static_cast<size_t>(S.f_warp_max) + static_cast<size_t>(S.s_warp_max) * S.s_warp_max);
  ssm = smax(ssm, S.n_group_lists * (2 * static_cast<size_t>(S.f_group_max) + static_cast<size_t>(S.s_group_max) * S.s_group_max));
  ssm = smax(ssm, 2 * static_cast<size_t>(S.f_top_max) + static_cast<size_t>(S.s_top_max) * S.s_top_max);
  const size_t with_x = (ssm + S.n) * sizeof(double);
  if (with_x <= static_cast<size_t>(smem_max)) {
    cfg->x_in_smem = 1;
    cfg->smem_solve = with_x;
  } else {
    cfg->x_in_smem = 0;
    cfg->smem_solve = ssm * sizeof(double);
    if (cfg->smem_solve > static_cast<size_t>(smem_max)) return cudaErrorInvalidConfiguration;
  }
  e = cudaFuncSetAttribute(lu_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(cfg->smem_factor));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(lu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(cfg->smem_solve));
  if (e != cudaSuccess) return e;
  int occ_f = 1, occ_s = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, lu_factor_kernel, cfg->threads, cfg->smem_factor);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, lu_solve_kernel, cfg->threads, cfg->smem_solve);
  if (occ_f < 1) occ_f = 1;
  if (occ_s < 1) occ_s = 1;
  cfg->grid_factor = sms * occ_f;
  cfg->grid_solve = sms * occ_s;
  if (cfg->grid_factor > B) cfg->grid_factor = B > 0 ? B : 1;
  if (cfg->grid_solve > B) cfg->grid_solve = B > 0 ? B : 1;
  return cudaSuccess;
}

cudaError_t lu_factor_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* JV, double* LU,
                             double* CBW, const double* mvals, const double* slot_mask,
                             const double* scale_b, double scale_const, const double* RHS, double* YOUT,
                             int store_l, const int* active, int* flags, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  lu_factor_kernel<<<cfg.grid_factor, cfg.threads, cfg.smem_factor, st>>>(
      S, JV, LU, CBW, mvals, slot_mask, scale_b, scale_const, RHS, YOUT, store_l, active, flags, B);
  return cudaGetLastError();
}

cudaError_t lu_solve_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* LU,
                            const double* RHS, double* X, double* CVW, double* XW, int do_fwd,
                            int mode, const int* active, int B, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  lu_solve_kernel<<<cfg.grid_solve, cfg.threads, cfg.smem_solve, st>>>(
      S, LU, RHS, X, CVW, XW, cfg.x_in_smem, do_fwd, mode, active, B);
  return cudaGetLastError();
}

}  // namespace szb
