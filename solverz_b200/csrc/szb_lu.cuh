// szb_lu.cuh -- device-side view of the multifrontal schedule and launch prototypes of the
// batch-interleaved LU (csrc/szb_lu.cu, csrc/szb_lu_front.cuh).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace szb {

constexpr int kGroup = 32;   // scenarios per group = lanes of a warp

// All arrays are device pointers (int32).  Built once per handle from szb_lu_symbolic.
struct LuDev {
  int n, nnz_a, ns;
  int lu_size, cb_size, cv_size;   // lines (of kGroup doubles) per group: factors, contribution ws, fwd ws
  int n_warp_lists, n_warp_sn, n_mid_lists, n_mid_sn, n_sched;
  int f_max, f_warp_max, warp_lines;
  const int *rowp, *colp;
  const int *idx, *rel;
  const int *a_src, *a_dst;
  const int *wl_ptr, *ml_ptr;
  const int *atp, *tcp;   // tile tables of the tiled fronts (levels 1 and 2)
  const int4* tce;
  const int4* fh;   // front descriptors, 4 x int4 each, schedule order (lu/schedule.py)
  const int4* ct;   // child table
};

struct LuLaunchCfg {
  int threads;      // threads per CTA
  int grid;         // resident CTAs (persistent grid over the groups)
  int groups;       // ceil(B / kGroup): groups the value arrays are sized for
  size_t smem;
};

// Chooses block/grid for this schedule on the current device; returns cudaError_t.
cudaError_t lu_plan(const LuDev& S, int B, LuLaunchCfg* cfg);

// map[p] = scenario at position p of the interleaved storage (active scenarios compacted in ascending
// order, -1 behind them up to the end of the last group), *npos = number of active scenarios.
// active == null: all B scenarios.
cudaError_t lu_map_launch(const int* active, int B, int* map, int* npos, cudaStream_t st);

// A(b) = scale(b) * slot_mask * JV[b] + mvals  ->  LU (interleaved by position, see lu_map_launch);
// scale_b may be null (then scale_const).  RHS non-null: the right-hand side rides along (forward
// substitution fused) and YI gets y = L^{-1} P b in pivot order (interleaved).  flags[b] = 2 for a zero /
// non-finite pivot or an element growth max|u| / max|a| above growth_tol (MON: 3 doubles per position of
// scratch for the monitors, carried across the three level kernels).
cudaError_t lu_factor_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* JV, double* LU, double* CBW,
                             const double* mvals, const double* slot_mask, const double* scale_b,
                             double scale_const, const double* RHS, double* YI, const int* map, const int* npos,
                             double* MON, int* flags, double growth_tol, cudaStream_t st);

// mode 0:  X[b] <- A^{-1} RHS[b]         (X may alias RHS)
// mode 1:  X[b] <- X[b] - A^{-1} RHS[b]  (Newton update y -= dx; X is Y)
// do_fwd 1: RHS is b in original order, scenario-major (forward + backward sweep)
// do_fwd 0: YI holds y in pivot order from a fused factorisation (backward sweep only)
// map / npos: as left by the factorisation; active (may be null) masks scenarios on top of it
cudaError_t lu_solve_launch(const LuDev& S, const LuLaunchCfg& cfg, const double* LU, const double* RHS, double* X,
                            double* CVW, double* YI, int do_fwd, int mode, const int* map, const int* npos,
                            const int* active, cudaStream_t st);

}  // namespace szb
