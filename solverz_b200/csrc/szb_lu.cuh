// szb_lu.cuh -- device-side view of the multifrontal schedule and launch prototypes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace szb {

// All arrays are device pointers (int32).  Built once per handle from szb_lu_symbolic.
struct LuDev {
  int n, nnz_a, ns;
  int lu_size, cb_size, cv_size;
  int n_warp_lists, n_group_lists, n_top;
  int n_warp_sn, n_group_sn;     // fronts in the warp lists / group lists (schedule order: warp, group, top)
  int threads, wpg;              // threads per CTA, warps per group (fixed by the host schedule)
  int f_warp_max, f_group_max, f_top_max, s_warp_max, s_group_max, s_top_max;
  int f_max;                     // largest front order
  const int *rowp, *colp;
  const int *sn_piv0, *sn_s, *sn_f;
  const int *idx_ptr, *idx, *rel_ptr, *rel, *child_ptr, *child;
  const int *a_ptr, *a_src, *a_dst;
  const int *lp_off, *up_off, *cb_off;
  const int *ea_ptr, *ea_dst;   // per child: destination of every contribution-block element in the parent front
  const int *wl_ptr, *wl, *gl_ptr, *gl, *top;
  const int4* fh;                // front headers, 4 x int4 each, schedule order (lu/schedule.py)
  const int2* ga;                // gather table: (tagged source, front position)
};

struct LuLaunchCfg {
  int threads;        // threads per block (multiple of 32, >= 32*n_warp_lists)
  int grid_factor;    // resident blocks (persistent grid size)
  size_t smem_factor; // dynamic shared memory of the factor kernel
  int grid_solve;
  size_t smem_solve;
  int top_in_global;  // CTA-level fronts do not fit shared memory: they live in a per-CTA global workspace
  size_t big_elems;   // doubles of that workspace per resident block
  int outer_nb;       // pivots per outer block of the global-front elimination (side panel columns in shared memory)
  int threads_fwd;    // forward sweep: threads per block
  int x_in_smem;      // forward kernel keeps the permuted vector in shared memory
  int threads_bwd;    // backward sweep: threads per block, resident blocks, dynamic shared memory
  int grid_bwd;
  size_t smem_bwd;
  int bwd_x_in_smem;
};

// Chooses block/grid/smem for this schedule on the current device; returns cudaError_t.
cudaError_t lu_plan(const LuDev& S, int B, LuLaunchCfg* cfg);

// A(b) = scale(b) * JV[b] (+ mvals)  ->  LU[b];  scale_b may be null (then scale_const).
// RHS / YOUT non-null: the right-hand side rides along (forward substitution fused), YOUT gets
// y = L^{-1} P b in pivot order.  store_l == 0 skips writing the L21 rows (Newton never reuses L).
cudaError_t lu_factor_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* JV,
                             double* LU, double* CBW, double* FBW, const double* mvals, const double* slot_mask,
                             const double* scale_b, double scale_const, const double* RHS, double* YOUT, int store_l,
                             const int* active, int* flags, int B, cudaStream_t st);

// mode 0:  X[b] <- A^{-1} RHS[b]         (X may alias RHS)
// mode 1:  X[b] <- X[b] - A^{-1} RHS[b]  (Newton update y -= dx; X is Y)
// do_fwd 1: RHS is b in original order (forward + backward sweep)
// do_fwd 0: RHS is y in pivot order from a fused factorisation (backward sweep only)
// YW: [B][n] scratch for y in pivot order (used when do_fwd == 1)
cudaError_t lu_solve_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* LU,
                            const double* RHS, double* X, double* CVW, double* XW, double* YW, int do_fwd,
                            int mode, const int* active, int B, cudaStream_t st);

}  // namespace szb
