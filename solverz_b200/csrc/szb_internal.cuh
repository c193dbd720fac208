// szb_internal.cuh -- the handle behind include/szb.h (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "szb.h"
#include "szb_lu.cuh"
#include "szb_model.h"

struct szb_handle {
  using launch_fn = int (*)(const szbm_args*, void*);

  int device = 0;
  int n = 0, nnz_j = 0, B = 0;
  std::string err;
  // LU schedule (device copies owned by `owned`)
  szb::LuDev S{};
  szb::LuLaunchCfg cfg{};
  const double* mass_vals = nullptr;
  // model library
  void* model_dl = nullptr;
  launch_fn launch_F = nullptr, launch_J = nullptr;
  std::vector<szbm_param_info> params;
  std::vector<void*> param_ptr;   // what the kernels see (bound or own)
  std::vector<void*> param_own;
  std::vector<char> param_bound;
  // per-batch buffers
  double *own_Y = nullptr, *own_F = nullptr, *own_JV = nullptr, *own_RHS = nullptr;
  double *bound_Y = nullptr, *bound_F = nullptr, *bound_JV = nullptr, *bound_RHS = nullptr;
  double *Y = nullptr, *F = nullptr, *JV = nullptr, *RHS = nullptr;
  int64_t rhs_len = 0;
  double *LU = nullptr, *CBW = nullptr, *FBW = nullptr, *CVW = nullptr, *XW = nullptr, *ybuf = nullptr;
  double *fnorm = nullptr, *scale_b = nullptr;
  int *active = nullptr, *nstep = nullptr, *flags = nullptr, *counter = nullptr;
  int* pinned = nullptr;  // 16 ints of pinned host memory
  double time = 0.0;   // t used by szb_newton for F(t,y,p) / J(t,y,p) (FDAE steps)
  bool factored = false;
  bool has_lu = false;
  // bookkeeping
  std::vector<void*> owned;        // freed in destroy
  std::vector<void*> batch_owned;  // freed in set_batch / destroy
  int64_t ws_bytes = 0;
  int64_t launches[5] = {0, 0, 0, 0, 0};  // F, J, LU, solve, bookkeeping
  bool profile = false;
  double prof_ms[5] = {0, 0, 0, 0, 0};
  int64_t prof_units[5] = {0, 0, 0, 0, 0};  // scenarios processed per kernel class
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int last_sweeps = 0;
  // Rodas workspaces (allocated on first use)
  std::vector<void*> rodas_owned;
  int rodas_B = 0, rodas_stages = 0;
  double *rd_K = nullptr, *rd_y1 = nullptr, *rd_sum2 = nullptr, *rd_dfdt = nullptr, *rd_F1 = nullptr;
  double *rd_dbl = nullptr, *rd_tout = nullptr, *rd_ev = nullptr;
  double *rd_y0s = nullptr, *rd_F0s = nullptr;   // Hermite dense output (rodas3): y and F at the start of the step
  int *rd_int = nullptr, *rd_mcol = nullptr;
  bool rodas_started = false;   // a szb_rodas call has initialised the controller state of this batch
  int rd_tout_cap = 0;
  bool rd_mcol_ready = false;
  std::vector<int> mass_row_col;   // row -> column of the unit mass entry, -1 for algebraic rows
  // implicit one-step mode of szb_newton (szb_set_implicit)
  int imp_mode = 0;
  double imp_c = 0.0;
  const double *imp_y0 = nullptr, *imp_f0 = nullptr;
  double* imp_G = nullptr;         // [B][n] residual of the step equation (batch-owned)
  int* imp_mcol = nullptr;         // device copy of mass_row_col (owned)
  int imp_G_B = 0;
};

void szb_set_error(szb_handle* h, const char* fmt, ...);

#define SZB_CUDA(h, call)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      szb_set_error((h), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return SZB_ECUDA;                                                                \
    }                                                                                  \
  } while (0)

// shared by the drivers
int szb_eval_F_masked(szb_handle* h, double t, const int* active, cudaStream_t st);
int szb_eval_J_masked(szb_handle* h, double t, const int* active, cudaStream_t st);
int szb_lu_refactor_masked(szb_handle* h, double h_gamma, const double* scale_b, const int* active,
                           const double* rhs, double* yout, int store_l, cudaStream_t st,
                           const double* slot_mask = nullptr);
int szb_launch_F(szb_handle* h, double t, const double* t_b, const double* Y, double* F, double* fnorm,
                 const int* active, cudaStream_t st);
int szb_launch_J(szb_handle* h, double t, const double* t_b, const double* Y, double* JV, const int* active,
                 cudaStream_t st);
