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
static inline void cp_async8(double* dst, const double* src) { *dst = *src; }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
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
// 8-byte asynchronous copy global -> shared of THIS lane's value (LDGSTS): no register is tied up while the
// line is in flight, and because a lane only ever reads back what it copied itself, cp.async.wait_group is
// all the synchronisation it needs
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace szb
#endif

namespace szb {

#define SZB_EL(p, e) (p)[static_cast<ptrdiff_t>(e) * kLanes]

// optional phase timing of the factor kernel (tools/lu_timing.sh builds a -DSZB_LU_TIMING library)
#if defined(SZB_LU_TIMING) && !defined(SZB_HOST_EMU)
#define SZB_TIC() long long t__ = clock64()
#define SZB_TOC(L, i) do { const long long n__ = clock64(); (L).tacc[(i) + (L).tbase] += n__ - t__; t__ = n__; } while (0)
#else
#define SZB_TIC() do {} while (0)
#define SZB_TOC(L, i) do {} while (0)
#endif

constexpr int kNB = 8;   // pivots per sub-panel = rows / columns of a register tile

// one front of the schedule (lu/schedule.py: 16 int32 per front)
struct Front {
  int s, f, piv0, lp, up, cbo;   // pivots, order, first pivot, offsets of P / U12 / contribution block
  int a0, a1;                    // matrix entries of the front: a_src / a_dst [a0, a1)
  int c0, c1;                    // children: rows [c0, c1) of the child table
  int idx0;                      // front index list (pivot-order indices; the first s are the pivots)
  int cvo;                       // offset of the contribution VECTOR (forward sweep with stored factors)
  int atp0, tcp0;                // tiled fronts: start of the per-tile pointers in Tabs::atp / Tabs::tcp, else -1
  int yo;                        // offset of the pivot rows of the right-hand-side column in Lane::y (= piv0 in place)
  int abias;                     // subtracted from a_dst (0 in place; lp_off when the front lives in shared memory)
};

SZB_D Front load_front(const i4* __restrict__ fh, const int q) {
  const i4 a = ldg(fh + 4 * q), b = ldg(fh + 4 * q + 1), c = ldg(fh + 4 * q + 2), d = ldg(fh + 4 * q + 3);
  Front F;
  F.s = a.x; F.f = a.y; F.piv0 = a.z; F.lp = a.w;
  F.up = b.x; F.cbo = b.y; F.a0 = b.z; F.a1 = b.w;
  F.c0 = c.x; F.c1 = c.y; F.idx0 = c.z; F.cvo = c.w;
  F.atp0 = d.x; F.tcp0 = d.y;
  F.yo = F.piv0; F.abias = 0;
  return F;
}

struct Tabs {
  const i4* fh;      // front descriptors, schedule order
  const i4* ct;      // child table, 2 x int4 per edge: (cb_off, u, rel0, cv_off), (0, 0, 0, 0)
  const int* atp;    // tiled fronts: per tile (tr + tc * nt) the run of its matrix entries in a_src / a_dst
  const int* tcp;    // tiled fronts: per tile (rhs column: tc = nt) the run of its contributors in tce
  const i4* tce;     // contributors: (workspace offset of the piece, u of the child, row nibbles, column nibbles)
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
  double* cb;            // contribution workspace of the CTA, offset by the lane (where this front's block goes)
  const double* cbsrc;   // contribution workspace the children's blocks are read from (global)
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
#if defined(SZB_LU_TIMING) && !defined(SZB_HOST_EMU)
  long long* tacc;       // cycle counters per phase
  int tbase;             // 0: buffered fronts, 8: in-place fronts
#endif
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
  return L.y + static_cast<ptrdiff_t>(F.yo) * kLanes;
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
  SZB_TIC();
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
      SZB_EL(L.y, F.yo + r) = L.valid ? L.rhs[row] : 0.0;
    }
  }
  coop_sync(C);
  SZB_TOC(L, 0);
  // matrix entries: unique destinations inside P / U12; kAsm independent loads in flight per lane (the values
  // sit in scenario-major HBM: one sector per lane)
  {
    constexpr int kAsm = 8;
    for (int e0 = F.a0 + C.w * kAsm; e0 < F.a1; e0 += C.W * kAsm) {
      int src[kAsm], dst[kAsm];
      double v[kAsm];
      SZB_UNROLL
      for (int q = 0; q < kAsm; ++q) {
        const bool in = e0 + q < F.a1;
        src[q] = in ? ldg(T.a_src + e0 + q) : -1;
        dst[q] = in ? ldg(T.a_dst + e0 + q) : 0;
      }
      SZB_UNROLL
      for (int q = 0; q < kAsm; ++q) v[q] = src[q] >= 0 ? entry_value(L, src[q]) : 0.0;
      SZB_UNROLL
      for (int q = 0; q < kAsm; ++q) {
        if (src[q] >= 0) {
          const double a = dabs(v[q]);
          L.amax = a > L.amax ? a : L.amax;
          SZB_EL(L.lu, dst[q] - F.abias) = v[q];
        }
      }
    }
  }
  coop_sync(C);
  SZB_TOC(L, 1);
  // children, one after the other: a child column goes to ONE parent column, so the columns of a child
  // can be dealt to the warps without conflicts
  const int ncc = L.with_rhs ? 1 : 0;
  for (int c = F.c0; c < F.c1; ++c) {
    const i4 ch = ldg(T.ct + 2 * c);
    const int uc = ch.y;
    const int* __restrict__ rel = T.rel + ch.z;
    const double* src = L.cbsrc + static_cast<ptrdiff_t>(ch.x) * kLanes;
    for (int jc = C.w; jc < uc + ncc; jc += C.W) {
      const int pc = jc < uc ? ldg(rel + jc) : f;
      double* top = col_top(L, F, pc);
      double* bot = col_bot(L, F, pc);
      const double* sc = src + static_cast<ptrdiff_t>(jc) * uc * kLanes;
      int ic = 0;
      for (; ic + 8 <= uc; ic += 8) {
        int pr[8];
        double v[8], d[8];
        double* dp[8];
        SZB_UNROLL
        for (int q = 0; q < 8; ++q) pr[q] = ldg(rel + ic + q);
        SZB_UNROLL
        for (int q = 0; q < 8; ++q) {
          dp[q] = (pr[q] < s ? top : bot) + static_cast<ptrdiff_t>(pr[q]) * kLanes;
          v[q] = SZB_EL(sc, ic + q);
          d[q] = *dp[q];
        }
        SZB_UNROLL
        for (int q = 0; q < 8; ++q) *dp[q] = d[q] + v[q];
      }
      for (; ic < uc; ++ic) {
        const int pr = ldg(rel + ic);
        double* dp = (pr < s ? top : bot) + static_cast<ptrdiff_t>(pr) * kLanes;
        *dp += SZB_EL(sc, ic);
      }
    }
    coop_sync(C);
  }
  SZB_TOC(L, 2);
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
    SZB_TIC();
    if (C.w == 0) diag_block_factor(L, F, k0, nb);
    coop_sync(C);
    SZB_TOC(L, 3);
    panel_solve(L, F, k0, nb, ncols, C);
    coop_sync(C);
    SZB_TOC(L, 4);
    if (k0 + nb < ncols) trailing_update(L, F, k0, nb, ncols, C);
    coop_sync(C);
    SZB_TOC(L, 5);
  }
}

SZB_D void factor_front(Lane& L, const Front& F, const Tabs& T, const Coop& C) {
  front_assemble(L, F, T, C);
  front_eliminate(L, F, C);
}

// lines of a buffer that holds a whole front: P, U12, the contribution block with its right-hand-side
// column, and the pivot rows of the right-hand side
SZB_D int front_lines(const int f) { return f * (f + 1); }

// The same front, but assembled and eliminated in a (shared-memory) buffer `buf` (lane pointer) laid out
// [P | U12 | CB | y_top], then written ONCE to the factor storage / contribution workspace: the children's
// blocks are read once, every factor and contribution line is written once, nothing else leaves the SM.
SZB_D void factor_front_buffered(Lane& L, const Front& F, const Tabs& T, const Coop& C, double* buf) {
  const int s = F.s, f = F.f, u = f - s;
  Front Fb = F;
  Fb.lp = 0; Fb.up = f * s; Fb.cbo = f * s + s * u; Fb.yo = f * s + s * u + u * (u + 1);
  Fb.abias = F.lp;
  Lane Lb = L;
  Lb.lu = buf; Lb.cb = buf; Lb.y = buf;
  front_assemble(Lb, Fb, T, C);
  front_eliminate(Lb, Fb, C);
  L.amax = Lb.amax; L.umax = Lb.umax; L.bad = Lb.bad;
  SZB_TIC();
  const int n_pu = f * s + s * u, n_cb = u * (u + (L.with_rhs ? 1 : 0));
  double* pu = L.lu + static_cast<ptrdiff_t>(F.lp) * kLanes;
  for (int e = C.w; e < n_pu; e += C.W) SZB_EL(pu, e) = SZB_EL(buf, e);
  double* cb = L.cb + static_cast<ptrdiff_t>(F.cbo) * kLanes;
  const double* cs = buf + static_cast<ptrdiff_t>(Fb.cbo) * kLanes;
  for (int e = C.w; e < n_cb; e += C.W) SZB_EL(cb, e) = SZB_EL(cs, e);
  if (L.with_rhs)
    for (int r = C.w; r < s; r += C.W) SZB_EL(L.y, F.piv0 + r) = SZB_EL(buf, Fb.yo + r);
  coop_sync(C);
  SZB_TOC(L, 6);
}

// ---------------------------------------------------------------------------------------------------
// Large fronts (they do not fit the shared-memory buffer): Crout by register tiles of kNB x kNB lines.
// The front is cut into tiles on the pivot grid: tile rows / columns 0..Tt-1 cover the s pivots (the last
// one may be short), Tt..nt-1 the u update rows from s on, column nt is the right-hand side.  A tile is
//   1. ASSEMBLED on chip: its matrix entries and the matching pieces of the children's contribution blocks
//      are gathered into a shared-memory scratch tile (host tables atp / rsp say which), so nothing is
//      zero-filled or read-modified-written in global memory and every child line is read once;
//   2. updated in registers with ALL the pivots it depends on, k < min(row origin, column origin, s):
//      acc[r][c] -= L[i0 + r, k] * U[k, j0 + c], one operand private to the warp and streamed, the other
//      shared by the warps that work on the same strip (L1);
//   3. finished in registers (diagonal tile: LU; below it: * U_KK^{-1}; right of it: L_KK^{-1} *) with the
//      factored diagonal tile read from shared memory, and STORED once to its final place.
// Step K = diagonal tile (K, K), then the column strip below and the row strip right of it, dealt to the
// warps; after the last pivot tile the Schur block.  Every front element crosses the memory system twice
// (children -> tile, tile -> factor / contribution storage) plus 1/8 line of operand per multiply-add.
// ---------------------------------------------------------------------------------------------------
SZB_D int tile_start(const Front& F, const int Tt, const int t) { return t < Tt ? t * kNB : F.s + (t - Tt) * kNB; }
SZB_D int tile_size(const Front& F, const int Tt, const int t) {
  const int n = (t < Tt ? F.s : F.f) - tile_start(F, Tt, t);
  return n < kNB ? n : kNB;
}

// The k-loop of a tile.  Per pivot k a warp needs kNB lines of L (its rows, private) and kNB lines of U (its
// columns).  With 64 accumulators per lane the register file cannot hold more than two pivots of operands in
// flight, far too few for the ~1.5 k cycles of an L2 / HBM round trip, so the operands are prefetched kRing
// pivots ahead into a per-warp shared-memory ring with asynchronous copies and read back from there.
constexpr int kRing = 6;                       // pivots in flight per warp
constexpr int kRingLines = kRing * 2 * kNB;    // 96 lines of the warp's scratch

template <bool FULL>
SZB_D void tile_kloop(double (&acc)[kNB][kNB], const double* ap, const int f, const double* const (&bp)[kNB],
                      const int kend, const int nr, const int nc, double* ring) {
  auto issue = [&](const int k) {
    double* slot = ring + static_cast<ptrdiff_t>((k % kRing) * 2 * kNB) * kLanes;
    const double* a = ap + static_cast<ptrdiff_t>(k) * f * kLanes;
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r)
      if (FULL || r < nr) cp_async8(slot + static_cast<ptrdiff_t>(r) * kLanes, a + static_cast<ptrdiff_t>(r) * kLanes);
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c)
      if (FULL || c < nc) cp_async8(slot + static_cast<ptrdiff_t>(kNB + c) * kLanes, bp[c] + static_cast<ptrdiff_t>(k) * kLanes);
    cp_async_commit();
  };
  for (int k = 0; k < kRing; ++k) {
    if (k < kend) issue(k);
    else cp_async_commit();                    // empty groups keep the group count of the wait below fixed
  }
  for (int k = 0; k < kend; ++k) {
    cp_async_wait<kRing - 1>();                // the group of pivot k has landed
    const double* slot = ring + static_cast<ptrdiff_t>((k % kRing) * 2 * kNB) * kLanes;
    double a[kNB], b[kNB];
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r) a[r] = (FULL || r < nr) ? SZB_EL(slot, r) : 0.0;
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c) b[c] = (FULL || c < nc) ? SZB_EL(slot, kNB + c) : 0.0;
    if (k + kRing < kend) issue(k + kRing);
    else cp_async_commit();
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c)
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) acc[r][c] = dfma(-a[r], b[c], acc[r][c]);
  }
  cp_async_wait<0>();
}

enum TileMode { kTileSchur = 0, kTileDiag = 1, kTileBelow = 2, kTileRight = 3 };

// tbuf: this warp's scratch (kRingLines >= kNB * kNB lines, lane pointer): the tile is gathered there, then the
// same lines are the operand ring of the k-loop
SZB_D void tile_accumulate(Lane& L, const Front& F, const Tabs& T, const int Tt, const int nt, const int tr,
                           const int tc, double* tbuf, double (&acc)[kNB][kNB]) {
  const int s = F.s, f = F.f;
  const bool rhs_col = tc == nt;
  const int i0 = tile_start(F, Tt, tr), nr = tile_size(F, Tt, tr);
  const int j0 = rhs_col ? f : tile_start(F, Tt, tc), nc = rhs_col ? 1 : tile_size(F, Tt, tc);
  // ---- 1. assemble the tile in the scratch
  SZB_TIC();
  SZB_UNROLL
  for (int e = 0; e < kNB * kNB; ++e) SZB_EL(tbuf, e) = 0.0;
  if (!rhs_col) {
    const int e0 = ldg(T.atp + F.atp0 + tr + tc * nt), e1 = ldg(T.atp + F.atp0 + tr + tc * nt + 1);
    for (int eb = e0; eb < e1; eb += kNB) {
      int src[kNB], dst[kNB];
      double v[kNB];
      SZB_UNROLL
      for (int q = 0; q < kNB; ++q) {
        const bool in = eb + q < e1;
        src[q] = in ? ldg(T.a_src + eb + q) : -1;
        dst[q] = in ? ldg(T.a_dst + eb + q) : 0;
      }
      SZB_UNROLL
      for (int q = 0; q < kNB; ++q) v[q] = src[q] >= 0 ? entry_value(L, src[q]) : 0.0;
      SZB_UNROLL
      for (int q = 0; q < kNB; ++q) {
        if (src[q] >= 0) {
          const double a = dabs(v[q]);
          L.amax = a > L.amax ? a : L.amax;
          SZB_EL(tbuf, dst[q]) = v[q];
        }
      }
    }
  } else if (tr < Tt) {
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r)
      if (r < nr) SZB_EL(tbuf, r) = L.valid ? L.rhs[ldg(T.rowp + F.piv0 + i0 + r)] : 0.0;
  }
  SZB_TOC(L, 0);
  // the pieces of the children's contribution blocks that land in this tile (host table: one int4 each)
  {
    const int e0 = ldg(T.tcp + F.tcp0 + tr + tc * nt), e1 = ldg(T.tcp + F.tcp0 + tr + tc * nt + 1);
    for (int e = e0; e < e1; ++e) {
      const i4 ent = ldg(T.tce + e);
      const double* src = L.cbsrc + static_cast<ptrdiff_t>(ent.x) * kLanes;
      const int uc = ent.y;
      const unsigned rp = static_cast<unsigned>(ent.z), cpk = static_cast<unsigned>(ent.w);
      double v[kNB][kNB];
      SZB_UNROLL
      for (int p = 0; p < kNB; ++p)
        SZB_UNROLL
        for (int q = 0; q < kNB; ++q)
          v[p][q] = (((cpk >> (4 * p)) & 0xFu) != 0xFu && ((rp >> (4 * q)) & 0xFu) != 0xFu)
                        ? SZB_EL(src, q + static_cast<ptrdiff_t>(p) * uc) : 0.0;
      SZB_UNROLL
      for (int p = 0; p < kNB; ++p)
        SZB_UNROLL
        for (int q = 0; q < kNB; ++q) {
          const unsigned cl = (cpk >> (4 * p)) & 0xFu, rl = (rp >> (4 * q)) & 0xFu;
          if (cl != 0xFu && rl != 0xFu) SZB_EL(tbuf, rl + kNB * cl) += v[p][q];
        }
    }
  }
  SZB_TOC(L, 1);
  // ---- 2. into registers, 3. all the pivots the tile depends on
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c)
    SZB_UNROLL
    for (int r = 0; r < kNB; ++r) acc[r][c] = SZB_EL(tbuf, r + kNB * c);
  const int ki = tr < Tt ? i0 : s, kj = (rhs_col || tc >= Tt) ? s : j0;
  const int kend = ki < kj ? ki : kj;
  if (kend > 0) {
    const double* bp[kNB];
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c) bp[c] = col_top(L, F, j0 + (c < nc ? c : 0));
    const double* ap = L.lu + static_cast<ptrdiff_t>(F.lp + i0) * kLanes;
    if (nr == kNB && nc == kNB) tile_kloop<true>(acc, ap, f, bp, kend, nr, nc, tbuf);
    else tile_kloop<false>(acc, ap, f, bp, kend, nr, nc, tbuf);
  }
  SZB_TOC(L, 2);
}

// dbuf / ldd: the factored diagonal tile of the step, element (r, c) at dbuf[r + ldd * c]: a shared-memory copy
// (ldd = kNB; written here when the tile is the diagonal one) or its final place in P (ldd = f)
SZB_D void tile_finish(Lane& L, const Front& F, const int Tt, const int nt, const int tr, const int tc,
                       double* dbuf, const int ldd, const bool write_d, const int mode, double (&acc)[kNB][kNB]) {
  const bool rhs_col = tc == nt;
  const int i0 = tile_start(F, Tt, tr), nr = tile_size(F, Tt, tr);
  const int j0 = rhs_col ? F.f : tile_start(F, Tt, tc), nc = rhs_col ? 1 : tile_size(F, Tt, tc);
  SZB_TIC();
  // ---- 4. finish
  if (mode == kTileDiag) {
    SZB_UNROLL
    for (int k = 0; k < kNB; ++k) {
      if (k < nr) {
        const double piv = acc[k][k];
        pivot_check(L, piv);
        const double a = dabs(piv);
        L.umax = a > L.umax ? a : L.umax;
        const double rp = 1.0 / piv;
        acc[k][k] = rp;
        SZB_UNROLL
        for (int r = k + 1; r < kNB; ++r) acc[r][k] *= rp;
        SZB_UNROLL
        for (int c = k + 1; c < kNB; ++c) {
          SZB_UNROLL
          for (int r = k + 1; r < kNB; ++r) acc[r][c] = dfma(-acc[r][k], acc[k][c], acc[r][c]);
          const double a2 = dabs(acc[k][c]);
          L.umax = a2 > L.umax ? a2 : L.umax;
        }
      }
    }
    if (write_d) {
      SZB_UNROLL
      for (int c = 0; c < kNB; ++c)
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r) SZB_EL(dbuf, r + ldd * c) = acc[r][c];
    }
  } else if (mode == kTileBelow) {   // rows of L: x U_KK = a
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c) {
      if (c < nc) {
        SZB_UNROLL
        for (int m = 0; m < c; ++m) {
          const double d = SZB_EL(dbuf, m + ldd * c);
          SZB_UNROLL
          for (int r = 0; r < kNB; ++r) acc[r][c] = dfma(-acc[r][m], d, acc[r][c]);
        }
        const double rp = SZB_EL(dbuf, c + ldd * c);
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r) acc[r][c] *= rp;
      }
    }
  } else if (mode == kTileRight) {   // columns of U: L_KK y = a
    SZB_UNROLL
    for (int r = 1; r < kNB; ++r) {
      if (r < nr) {
        SZB_UNROLL
        for (int m = 0; m < r; ++m) {
          const double d = SZB_EL(dbuf, r + ldd * m);
          SZB_UNROLL
          for (int c = 0; c < kNB; ++c) acc[r][c] = dfma(-d, acc[m][c], acc[r][c]);
        }
      }
    }
    SZB_UNROLL
    for (int c = 0; c < kNB; ++c)
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r)
        if (r < nr && c < nc) {
          const double v = dabs(acc[r][c]);
          L.umax = v > L.umax ? v : L.umax;
        }
  }
  // ---- 5. store once
  SZB_UNROLL
  for (int c = 0; c < kNB; ++c) {
    if (c < nc) {
      double* cp = (tr < Tt ? col_top(L, F, j0 + c) : col_bot(L, F, j0 + c)) + static_cast<ptrdiff_t>(i0) * kLanes;
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r)
        if (r < nr) SZB_EL(cp, r) = acc[r][c];
    }
  }
  SZB_TOC(L, 6);
}

// tbuf: this warp's scratch of kTileScratch lines; dsm: kNB * kNB lines of shared memory for the diagonal tile
// of the step when the warps of a CTA share the front, null when one warp walks it alone (it then reads the
// diagonal tile back from its final place in P)
constexpr int kTileScratch = 100;   // >= kRingLines and >= kNB * kNB
SZB_D void factor_front_tiled(Lane& L, const Front& F, const Tabs& T, const Coop& C, double* tbuf, double* dsm) {
  const int s = F.s, f = F.f, u = f - s;
  const int Tt = (s + kNB - 1) / kNB, Tb = (u + kNB - 1) / kNB, nt = Tt + Tb;
  const int ntc = nt + (L.with_rhs ? 1 : 0);
  double acc[kNB][kNB];
  for (int K = 0; K < Tt; ++K) {
    double* dbuf = dsm ? dsm : L.lu + static_cast<ptrdiff_t>(F.lp + K * kNB + K * kNB * f) * kLanes;
    const int ldd = dsm ? kNB : f;
    // items of the step: the diagonal tile, the column strip below it, the row strip right of it.  Every warp
    // accumulates its first item while warp 0 factors the diagonal tile; the finishes wait for that.
    const int n_col = nt - K - 1, n_items = 1 + n_col + (ntc - K - 1);
    auto where = [&](const int it, int& tr, int& tc, int& mode) {
      if (it == 0) { tr = K; tc = K; mode = kTileDiag; }
      else if (it <= n_col) { tr = K + it; tc = K; mode = kTileBelow; }
      else { tr = K; tc = K + (it - n_col); mode = kTileRight; }
    };
    int tr = 0, tc = 0, mode = 0;
    const bool have = C.w < n_items;
    if (have) {
      where(C.w, tr, tc, mode);
      tile_accumulate(L, F, T, Tt, nt, tr, tc, tbuf, acc);
      if (mode == kTileDiag) tile_finish(L, F, Tt, nt, tr, tc, dbuf, ldd, dsm != nullptr, mode, acc);
    }
    coop_sync(C);
    if (have && mode != kTileDiag) tile_finish(L, F, Tt, nt, tr, tc, dbuf, ldd, false, mode, acc);
    for (int it = C.w + C.W; it < n_items; it += C.W) {
      where(it, tr, tc, mode);
      tile_accumulate(L, F, T, Tt, nt, tr, tc, tbuf, acc);
      tile_finish(L, F, Tt, nt, tr, tc, dbuf, ldd, false, mode, acc);
    }
    coop_sync(C);
  }
  const int ncb = ntc - Tt;
  for (int t = C.w; t < Tb * ncb; t += C.W) {
    const int tr = Tt + t % Tb, tc = Tt + t / Tb;
    tile_accumulate(L, F, T, Tt, nt, tr, tc, tbuf, acc);
    tile_finish(L, F, Tt, nt, tr, tc, nullptr, kNB, false, kTileSchur, acc);
  }
  coop_sync(C);
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
    const i4 ch = ldg(T.ct + 2 * c);
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
// `red`: lane pointer to W * kNB lines of scratch (shared memory) for the partial sums of split columns
SZB_D void backward_front(const Lane& L, const Front& F, const Tabs& T, const Coop& C, double* red) {
  const int s = F.s, f = F.f, u = f - s;
  double* ytop = L.y + static_cast<ptrdiff_t>(F.piv0) * kLanes;
  const int* __restrict__ ids = T.idx + F.idx0 + s;
  // y_piv -= U12 x_upd.  Work items = (chunk of kNB pivot rows) x (segment of the u columns): when the front
  // has fewer row chunks than warps the columns are split and the partial sums meet in `red`, summed in
  // segment order (deterministic).
  if (u > 0) {
    const int chunks = (s + kNB - 1) / kNB;
    int nseg = C.W / chunks;
    nseg = nseg < 1 ? 1 : nseg;
    const int seg = (u + nseg - 1) / nseg;
    for (int it = C.w; it < chunks * nseg; it += C.W) {
      const int ch = it / nseg, sg = it - ch * nseg;
      const int r0 = ch * kNB;
      const int nr = s - r0 < kNB ? s - r0 : kNB;
      const int j0 = sg * seg, j1 = j0 + seg < u ? j0 + seg : u;
      const double* up = L.lu + static_cast<ptrdiff_t>(F.up + r0) * kLanes;
      double acc[kNB];
      SZB_UNROLL
      for (int r = 0; r < kNB; ++r) acc[r] = (sg == 0 && r < nr) ? SZB_EL(ytop, r0 + r) : 0.0;
      int j = j0;
      for (; j + 4 <= j1; j += 4) {
        double xj[4];
        SZB_UNROLL
        for (int q = 0; q < 4; ++q) xj[q] = SZB_EL(L.y, ldg(ids + j + q));
        SZB_UNROLL
        for (int q = 0; q < 4; ++q)
          SZB_UNROLL
          for (int r = 0; r < kNB; ++r)
            if (r < nr) acc[r] = dfma(-SZB_EL(up, (j + q) * s + r), xj[q], acc[r]);
      }
      for (; j < j1; ++j) {
        const double xj = SZB_EL(L.y, ldg(ids + j));
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r)
          if (r < nr) acc[r] = dfma(-SZB_EL(up, j * s + r), xj, acc[r]);
      }
      if (nseg == 1) {
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r)
          if (r < nr) SZB_EL(ytop, r0 + r) = acc[r];
      } else {
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r) SZB_EL(red, it * kNB + r) = acc[r];
      }
    }
    coop_sync(C);
    if (nseg > 1) {
      for (int ch = C.w; ch < chunks; ch += C.W) {
        const int r0 = ch * kNB;
        const int nr = s - r0 < kNB ? s - r0 : kNB;
        SZB_UNROLL
        for (int r = 0; r < kNB; ++r) {
          if (r < nr) {
            double a = SZB_EL(red, (ch * nseg) * kNB + r);
            for (int sg = 1; sg < nseg; ++sg) a += SZB_EL(red, (ch * nseg + sg) * kNB + r);
            SZB_EL(ytop, r0 + r) = a;
          }
        }
      }
      coop_sync(C);
    }
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
    }
    coop_sync(C);
  }
}

}  // namespace szb
