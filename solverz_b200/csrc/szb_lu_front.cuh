// szb_lu_front.cuh -- per-front arithmetic of the batch-interleaved multifrontal LU (lane = scenario).
//
// One CTA works on a GROUP of 32 scenarios that share the sparsity pattern: lane l of every warp owns
// scenario l of the group, and every value array the factorisation touches is stored interleaved,
// element e of the group at  base[e * 32 + lane]  (a "line" of 256 bytes).  Consequences:
//   * every access of a warp is one fully coalesced line, whatever the index pattern of the front;
//   * all index work (descriptors, relative indices, gather offsets) is warp-uniform and is paid once per
//     32 scenarios;
//   * lanes never exchange data: no shuffles, no atomics; warps only meet at CTA barriers between the
//     phases of a CTA-level front, and a warp that walks its own subtree needs no synchronisation at all.
// A front is never copied: its pivot panel [L11\U11; L21] and its U12 rows are assembled and eliminated IN
// PLACE in the factor storage, its Schur block (plus one right-hand-side column) in the contribution
// workspace of the CTA.  The logical f x (f+1) front of a supernode with s pivots, u = f - s update rows:
//
//        columns 0..s-1            columns s..f-1            column f (right-hand side)
//   rows 0..s-1   P[r + c*f]       U[r + (c-s)*s]            y[piv0 + r]
//   rows s..f-1   P[r + c*f]       CB[(r-s) + (c-s)*u]       CB[(r-s) + u*u]
//
// This file is plain C++ over those pointers, so that the same source builds (i) the device functions of
// csrc/szb_lu.cu and (ii) with -DSZB_HOST_EMU a one-lane host harness under tests/ that walks exactly this
// code on a machine without a GPU (test infrastructure only: the product has no CPU path).
//
// Reference seam replaced: solve(A, b) -> spsolve and lu_decomposition(A) -> splu / .solve(b)
// (Solverz/solvers/laesolver.py:11-45).
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef SZB_HOST_EMU
#include <cmath>
#define SZB_D static inline
#define SZB_UNROLL
namespace szb {
constexpr int kLanes = 1;
struct i4 { int x, y, z, w; };
template <typename T> static inline T ldg(const T* p) { return *p; }
static inline double dabs(double x) { return std::fabs(x); }
static inline double dfma(double a, double b, double c) { return std::fma(a, b, c); }
static inline bool dinf(double x) { return std::isinf(x); }
}  // namespace szb
#else
#define SZB_D __device__ __forceinline__
#define SZB_UNROLL _Pragma("unroll")
namespace szb {
constexpr int kLanes = 32;
using i4 = int4;
template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }
__device__ __forceinline__ double dabs(double x) { return fabs(x); }
__device__ __forceinline__ double dfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ bool dinf(double x) { return isinf(x); }
}  // namespace szb
#endif

namespace szb {

#define SZB_EL(p, e) (p)[static_cast<ptrdiff_t>(e) * kLanes]

constexpr int kNB = 8;   // pivots per sub-panel = rows / columns of a register tile

// one front of the schedule (lu/schedule.py: 16 int32 per front)
struct Front {
  int s, f, piv0, lp, up, cbo;   // pivots, order, first pivot, offsets of P / U12 / contribution block
  int a0, a1;                    // matrix entries of the front: a_src / a_dst [a0, a1)
  int c0, c1;                    // children: rows [c0, c1) of the child table
  int idx0;                      // front index list (pivot-order indices; the first s are the pivots)
  int cvo;                       // offset of the contribution VECTOR (forward sweep with stored factors)
};

SZB_D Front load_front(const i4* __restrict__ fh, const int q) {
  const i4 a = ldg(fh + 4 * q), b = ldg(fh + 4 * q + 1), c = ldg(fh + 4 * q + 2);
  Front F;
  F.s = a.x; F.f = a.y; F.piv0 = a.z; F.lp = a.w;
  F.up = b.x; F.cbo = b.y; F.a0 = b.z; F.a1 = b.w;
  F.c0 = c.x; F.c1 = c.y; F.idx0 = c.z; F.cvo = c.w;
  return F;
}

struct Tabs {
  const i4* fh;      // front descriptors, schedule order
  const i4* ct;      // child table: (cb_off, u, rel0, cv_off)
  const int* a_src;  // Jacobian slot of a matrix entry
  const int* a_dst;  // its offset in the factor storage
  const int* rel;    // positions of a child's update rows in the parent front
  const int* idx;    // front index lists
  const int* rowp;   // pivot k <- original row
  const int* colp;   // pivot k <- original column
};

// what one lane (= one scenario of the group) works on
struct Lane {
  double* lu;            // factor storage of the group, already offset by the lane
  double* cb;            // contribution workspace of the CTA, offset by the lane
  double* y;             // vector in pivot order (interleaved), offset by the lane
  const double* jv;      // scenario-major matrix values of this lane's scenario
  const double* rhs;     // scenario-major right-hand side of this lane's scenario, or null
  const double* mvals;   // optional mass values per slot (shared)
  const double* smask;   // optional 0/1 mask per slot (shared)
  double scale;          // A = scale * mask * J + mvals   (scale == 1 and no mvals / mask: A = J)
  bool plain;
  bool valid;            // the lane carries a scenario
  bool with_rhs;         // a right-hand side rides along as column f
  double amax, umax;     // max |a_ij| assembled, max |u_ij| produced (element-growth monitor)
  int bad;               // zero / non-finite pivot met
};

// how the warps of the CTA split a front: warp-level fronts are walked by ONE warp (W = 1, no barrier),
// CTA-level fronts by all of them with a CTA barrier between phases
struct Coop {
  int w, W;
  bool cta;
};

SZB_D void coop_sync(const Coop& c) {
#ifndef SZB_HOST_EMU
  if (c.cta) __syncthreads();
#else
  (void)c;
#endif
}

// lane pointer to column j of the TOP part of the front (rows r < s at top[r])
SZB_D double* col_top(const Lane& L, const Front& F, const int j) {
  if (j < F.s) return L.lu + static_cast<ptrdiff_t>(F.lp + j * F.f) * kLanes;
  if (j < F.f) return L.lu + static_cast<ptrdiff_t>(F.up + (j - F.s) * F.s) * kLanes;
  return L.y + static_cast<ptrdiff_t>(F.piv0) * kLanes;
}
// lane pointer to column j of the BOTTOM part, pre-offset so that row r >= s is at bot[r]
SZB_D double* col_bot(const Lane& L, const Front& F, const int j) {
  if (j < F.s) return L.lu + static_cast<ptrdiff_t>(F.lp + j * F.f) * kLanes;
  return L.cb + (static_cast<ptrdiff_t>(F.cbo) + static_cast<ptrdiff_t>(j - F.s) * (F.f - F.s) - F.s) * kLanes;
}

SZB_D double entry_value(const Lane& L, const int slot) {
  if (!L.valid) return 0.0;
  const double v = L.jv[slot];
  if (L.plain) return v;
  const double mv = L.mvals ? ldg(L.mvals + slot) : 0.0;
  const double mk = L.smask ? ldg(L.smask + slot) : 1.0;
  return dfma(L.scale * mk, v, mv);
}

// ---------------------------------------------------------------------------------------------------
// assembly: zero the front, scatter the matrix entries and the right-hand side, extend-add the children
// ---------------------------------------------------------------------------------------------------
SZB_D void front_assemble(Lane& L, const Front& F, const Tabs& T, const Coop& C) {
  const int s = F.s, f = F.f, u = f - s;
  const int n_pu = f * s + s * u;                      // P and U12 are adjacent in the factor storage
  const int n_cb = u * (u + (L.with_rhs ? 1 : 0));
  double* pu = L.lu + static_cast<ptrdiff_t>(F.lp) * kLanes;
  double* cb = L.cb + static_cast<ptrdiff_t>(F.cbo) * kLanes;
  for (int e = C.w; e < n_pu; e += C.W) SZB_EL(pu, e) = 0.0;
  for (int e = C.w; e < n_cb; e += C.W) SZB_EL(cb, e) = 0.0;
  if (L.with_rhs) {
    for (int r = C.w; r < s; r += C.W) {
      const int row = ldg(T.rowp + F.piv0 + r);
      SZB_EL(L.y, F.piv0 + r) = L.valid ? L.rhs[row] : 0.0;
    }
  }
  coop_sync(C);
  // matrix entries: unique destinations inside P / U12
  for (int e = F.a0 + C.w; e < F.a1; e += C.W) {
    const int src = ldg(T.a_src + e), dst = ldg(T.a_dst + e);
    const double v = entry_value(L, src);
    const double a = dabs(v);
    L.amax = a > L.amax ? a : L.amax;
    SZB_EL(L.lu, dst) = v;
  }
  coop_sync(C);
  // children, one after the other: a child column goes to ONE parent column, so the columns of a child
  // can be dealt to the warps without conflicts
  const int ncc = L.with_rhs ? 1 : 0;
  for (int c = F.c0; c < F.c1; ++c) {
    const i4 ch = ldg(T.ct + c);
    const int uc = ch.y;
    const int* __restrict__ rel = T.rel + ch.z;
    const double* src = L.cb + static_cast<ptrdiff_t>(ch.x) * kLanes;
    for (int jc = C.w; jc < uc + ncc; jc += C.W) {
      const int pc = jc < uc ? ldg(rel + jc) : f;
      double* top = col_top(L, F, pc);
      double* bot = col_bot(L, F, pc);
      const double* sc = src + static_cast<ptrdiff_t>(jc) * uc * kLanes;
      int ic = 0;
      for (; ic + 4 <= uc; ic += 4) {
        int pr[4];
        double v[4], d[4];
        double* dp[4];
        SZB_UNROLL
        for (int q = 0; q < 4; ++q) pr[q] = ldg(rel + ic + q);
        SZB_UNROLL
        for (int q = 0; q < 4; ++q) {
          dp[q] = (pr[q] < s ? top : bot) + static_cast<ptrdiff_t>(pr[q]) * kLanes;
          v[q] = SZB_EL(sc, ic + q);
          d[q] = *dp[q];
        }
        SZB_UNROLL
        for (int q = 0; q < 4; ++q) *dp[q] = d[q] + v[q];
      }
      for (; ic < uc; ++ic) {
        const int pr = ldg(rel + ic);
        double* dp = (pr < s ? top : bot) + static_cast<ptrdiff_t>(pr) * kLanes;
        *dp += SZB_EL(sc, ic);
      }
    }
    coop_sync(C);
  }
}

// ---------------------------------------------------------------------------------------------------
// elimination of the s pivots, right-looking in sub-panels of kNB pivots
// ---------------------------------------------------------------------------------------------------
SZB_D void pivot_check(Lane& L, const double piv) {
  if (!(dabs(piv) > 0.0) || dinf(piv)) L.bad = 1;
}

// LU of the nb x nb diagonal block at P[k0 + k0*f] (no pivoting); on return L below the diagonal (unit),
// U above it and the RECIPROCAL of the pivots on it
SZB_D void diag_block_factor(Lane& L, const Front& F, const int k0, const int nb) {
  double* d = L.lu + static_cast<ptrdiff_t>(F.lp + k0 + k0 * F.f) * kLanes;
  double D[kNB][kNB];
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c)
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r) D[r][c] = (r < nb && c < nb) ? SZB_EL(d, r + c * F.f) : 0.0;
  SZB_UNROLL
  for (int k = 0; k < kNB; ++k) {
    if (k < nb) {
      const double piv = D[k][k];
      pivot_check(L, piv);
      const double a = dabs(piv);
      L.umax = a > L.umax ? a : L.umax;
      const double rp = 1.0 / piv;
      D[k][k] = rp;
      SZB_UNROLL
      for (int r = k + 1; r < kNB; ++r) D[r][k] *= rp;
      SZB_UNROLL
      for (int c = k + 1; c < kNB; ++c) {
        SZB_UNROLL
        for (int r = k + 1; r < kNB; ++r) D[r][c] = dfma(-D[r][k], D[k][c], D[r][c]);
        const double a2 = dabs(D[k][c]);
        L.umax = a2 > L.umax ? a2 : L.umax;
      }
    }
  }
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c)
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r)
      if (r < nb && c < nb) SZB_EL(d, r + c * F.f) = D[r][c];
}

// rows below the block: L[i, k0..k0+nb) = A[i, k0..) U11^{-1};  columns right of it (U12 rows, right-hand
// side): U[k0..k0+nb, j) = L11^{-1} A[k0.., j).  Items (rows, then columns) are dealt to the warps.
SZB_D void panel_solve(Lane& L, const Front& F, const int k0, const int nb, const int ncols, const Coop& C) {
  const int f = F.f, k1 = k0 + nb;
  const double* d = L.lu + static_cast<ptrdiff_t>(F.lp + k0 + k0 * f) * kLanes;
  double D[kNB][kNB];
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c)
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r) D[r][c] = (r < nb && c < nb) ? SZB_EL(d, r + c * f) : 0.0;
  const int n_rows = f - k1, n_cols = ncols - k1;
  for (int it = C.w; it < n_rows + n_cols; it += C.W) {
    double a[kNB];
    if (it < n_rows) {
      double* p = L.lu + static_cast<ptrdiff_t>(F.lp + k0 * f + k1 + it) * kLanes;
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c) a[c] = c < nb ? SZB_EL(p, c * f) : 0.0;
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c) {
        SZB_UNROLL
        for (int m = 0; m < c; ++m) a[c] = dfma(-a[m], D[m][c], a[c]);
        a[c] *= D[c][c];
      }
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c)
        if (c < nb) SZB_EL(p, c * f) = a[c];
    } else {
      double* p = col_top(L, F, k1 + (it - n_rows)) + static_cast<ptrdiff_t>(k0) * kLanes;
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) a[r] = r < nb ? SZB_EL(p, r) : 0.0;
      SZB_UNROLL
      for (int r = 1; r < kNB; ++r)
        SZB_UNROLL
        for (int m = 0; m < r; ++m) a[r] = dfma(-D[r][m], a[m], a[r]);
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r)
        if (r < nb) {
          SZB_EL(p, r) = a[r];
          const double v = dabs(a[r]);
          L.umax = v > L.umax ? v : L.umax;
        }
    }
  }
}

// trailing update  A[i, j] -= sum_k L[i, k0+k] U[k0+k, j]  for i >= k1, j >= k1 in register tiles of
// kNB x kNB lines; the rows are cut at s so that a tile lies in one storage region per column
template <bool FULL>
SZB_D void update_tile(const Lane& L, const Front& F, const int k0, const int nb, const int i0, const int nr,
                       const bool top_rows, const int j0, const int nc) {
  double* cp[kNB];
  const double* up[kNB];
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c) {
    const int j = j0 + (FULL || c < nc ? c : 0);
    const double* t = col_top(L, F, j);
    up[c] = t + static_cast<ptrdiff_t>(k0) * kLanes;
    cp[c] = (top_rows ? const_cast<double*>(t) : col_bot(L, F, j)) + static_cast<ptrdiff_t>(i0) * kLanes;
  }
  const double* lp = L.lu + static_cast<ptrdiff_t>(F.lp + k0 * F.f + i0) * kLanes;
  double acc[kNB][kNB];
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c)
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r) acc[r][c] = (FULL || (r < nr && c < nc)) ? SZB_EL(cp[c], r) : 0.0;
  SZB_UNROLL
  for (int k = 0; k < kNB; ++k) {
    if (FULL || k < nb) {
      double a[kNB], b[kNB];
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) a[r] = (FULL || r < nr) ? SZB_EL(lp, k * F.f + r) : 0.0;
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c) b[c] = (FULL || c < nc) ? SZB_EL(up[c], k) : 0.0;
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c)
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r) acc[r][c] = dfma(-a[r], b[c], acc[r][c]);
    }
  }
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c)
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r)
      if (FULL || (r < nr && c < nc)) SZB_EL(cp[c], r) = acc[r][c];
}

SZB_D void trailing_update(const Lane& L, const Front& F, const int k0, const int nb, const int ncols,
                           const Coop& C) {
  const int s = F.s, f = F.f, k1 = k0 + nb;
  const int rt_top = (s - k1 + kNB - 1) / kNB, rt_bot = (f - s + kNB - 1) / kNB;
  const int rt = rt_top + rt_bot, ctn = (ncols - k1 + kNB - 1) / kNB;
  for (int t = C.w; t < rt * ctn; t += C.W) {
    const int ti = t % rt, tj = t / rt;
    const bool top = ti < rt_top;
    const int i0 = top ? k1 + ti * kNB : s + (ti - rt_top) * kNB;
    const int iend = top ? s : f;
    const int nr = iend - i0 < kNB ? iend - i0 : kNB;
    const int j0 = k1 + tj * kNB;
    const int nc = ncols - j0 < kNB ? ncols - j0 : kNB;
    if (nb == kNB && nr == kNB && nc == kNB) update_tile<true>(L, F, k0, nb, i0, nr, top, j0, nc);
    else update_tile<false>(L, F, k0, nb, i0, nr, top, j0, nc);
  }
}

SZB_D void front_eliminate(Lane& L, const Front& F, const Coop& C) {
  const int ncols = F.f + (L.with_rhs ? 1 : 0);
  for (int k0 = 0; k0 < F.s; k0 += kNB) {
    const int nb = F.s - k0 < kNB ? F.s - k0 : kNB;
    if (C.w == 0) diag_block_factor(L, F, k0, nb);
    coop_sync(C);
    panel_solve(L, F, k0, nb, ncols, C);
    coop_sync(C);
    if (k0 + nb < ncols) trailing_update(L, F, k0, nb, ncols, C);
    coop_sync(C);
  }
}

SZB_D void factor_front(Lane& L, const Front& F, const Tabs& T, const Coop& C) {
  front_assemble(L, F, T, C);
  front_eliminate(L, F, C);
}

// ---------------------------------------------------------------------------------------------------
// forward sweep with stored factors:  y = L^{-1} P b, contribution VECTORS passed up the tree through
// the workspace `cv` (lane pointer; a front's vector of u lines at F.cvo)
// ---------------------------------------------------------------------------------------------------
SZB_D void forward_front(const Lane& L, double* cv, const Front& F, const Tabs& T, const Coop& C) {
  const int s = F.s, f = F.f, u = f - s;
  double* mine = cv + static_cast<ptrdiff_t>(F.cvo) * kLanes;
  for (int r = C.w; r < s; r += C.W) {
    const int row = ldg(T.rowp + F.piv0 + r);
    SZB_EL(L.y, F.piv0 + r) = L.valid ? L.rhs[row] : 0.0;
  }
  for (int r = C.w; r < u; r += C.W) SZB_EL(mine, r) = 0.0;
  coop_sync(C);
  for (int c = F.c0; c < F.c1; ++c) {
    const i4 ch = ldg(T.ct + c);
    const int uc = ch.y;
    const int* __restrict__ rel = T.rel + ch.z;
    const double* src = cv + static_cast<ptrdiff_t>(ch.w) * kLanes;
    for (int ic = C.w; ic < uc; ic += C.W) {
      const int pr = ldg(rel + ic);
      double* dp = pr < s ? L.y + static_cast<ptrdiff_t>(F.piv0 + pr) * kLanes
                          : mine + static_cast<ptrdiff_t>(pr - s) * kLanes;
      *dp += SZB_EL(src, ic);
    }
    coop_sync(C);
  }
  double* ytop = L.y + static_cast<ptrdiff_t>(F.piv0) * kLanes;
  for (int k0 = 0; k0 < s; k0 += kNB) {
    const int nb = s - k0 < kNB ? s - k0 : kNB, k1 = k0 + nb;
    const double* d = L.lu + static_cast<ptrdiff_t>(F.lp + k0 + k0 * f) * kLanes;
    if (C.w == 0) {   // unit lower triangle of the diagonal block
      double a[kNB];
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) a[r] = r < nb ? SZB_EL(ytop, k0 + r) : 0.0;
      SZB_UNROLL
      for (int r = 1; r < kNB; ++r)
        SZB_UNROLL
        for (int m = 0; m < r; ++m)
          if (r < nb) a[r] = dfma(-SZB_EL(d, r + m * f), a[m], a[r]);
      SZB_UNROLL
      for (int r = 1; r < kNB; ++r)
        if (r < nb) SZB_EL(ytop, k0 + r) = a[r];
    }
    coop_sync(C);
    double yk[kNB];
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c) yk[c] = c < nb ? SZB_EL(ytop, k0 + c) : 0.0;
    const double* lcol = L.lu + static_cast<ptrdiff_t>(F.lp + k0 * f) * kLanes;
    for (int i = k1 + C.w; i < f; i += C.W) {
      double* dp = i < s ? ytop + static_cast<ptrdiff_t>(i) * kLanes : mine + static_cast<ptrdiff_t>(i - s) * kLanes;
      double acc = *dp;
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c)
        if (c < nb) acc = dfma(-SZB_EL(lcol, c * f + i), yk[c], acc);
      *dp = acc;
    }
    coop_sync(C);
  }
}

// ---------------------------------------------------------------------------------------------------
// backward sweep:  x_piv = U11^{-1} (y_piv - U12 x_upd), ancestors first; x overwrites y
// ---------------------------------------------------------------------------------------------------
SZB_D void backward_front(const Lane& L, const Front& F, const Tabs& T, const Coop& C) {
  const int s = F.s, f = F.f, u = f - s;
  double* ytop = L.y + static_cast<ptrdiff_t>(F.piv0) * kLanes;
  const int* __restrict__ ids = T.idx + F.idx0 + s;
  // y_piv -= U12 x_upd, pivot rows in chunks of kNB dealt to the warps
  if (u > 0) {
    const int chunks = (s + kNB - 1) / kNB;
    for (int ch = C.w; ch < chunks; ch += C.W) {
      const int r0 = ch * kNB;
      const int nr = s - r0 < kNB ? s - r0 : kNB;
      const double* up = L.lu + static_cast<ptrdiff_t>(F.up + r0) * kLanes;
      double acc[kNB];
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) acc[r] = r < nr ? SZB_EL(ytop, r0 + r) : 0.0;
      for (int j = 0; j < u; ++j) {
        const double xj = SZB_EL(L.y, ldg(ids + j));
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r)
          if (r < nr) acc[r] = dfma(-SZB_EL(up, j * s + r), xj, acc[r]);
      }
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r)
        if (r < nr) SZB_EL(ytop, r0 + r) = acc[r];
    }
    coop_sync(C);
  }
  // U11 x = y: diagonal blocks from the last to the first
  const int last = ((s - 1) / kNB) * kNB;
  for (int k0 = last; k0 >= 0; k0 -= kNB) {
    const int nb = s - k0 < kNB ? s - k0 : kNB;
    const double* d = L.lu + static_cast<ptrdiff_t>(F.lp + k0 + k0 * f) * kLanes;
    if (C.w == 0) {
      double a[kNB];
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) a[r] = r < nb ? SZB_EL(ytop, k0 + r) : 0.0;
      SZB_UNROLL
      for (int r = kNB - 1; r >= 0; --r) {
        if (r < nb) {
          SZB_UNROLL
          for (int m = r + 1; m < kNB; ++m)
            if (m < nb) a[r] = dfma(-SZB_EL(d, r + m * f), a[m], a[r]);
          a[r] *= SZB_EL(d, r + r * f);   // reciprocal of the pivot
        }
      }
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r)
        if (r < nb) SZB_EL(ytop, k0 + r) = a[r];
    }
    coop_sync(C);
    if (k0 > 0) {
      double xk[kNB];
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c) xk[c] = c < nb ? SZB_EL(ytop, k0 + c) : 0.0;
      const double* ucol = L.lu + static_cast<ptrdiff_t>(F.lp + k0 * f) * kLanes;
      for (int r = C.w; r < k0; r += C.W) {
        double acc = SZB_EL(ytop, r);
        SZB_UNROLL
        for (int c = 0; c < kNB; ++c)
          if (c < nb) acc = dfma(-SZB_EL(ucol, c * f + r), xk[c], acc);
        SZB_EL(ytop, r) = acc;
      }
      coop_sync(C);
    }
  }
}

}  // namespace szb
