/* szb_model.h -- interface between libszb.so and a generated model library.
 *
 * `module_printer(..., lang='cuda').render()` (solverz_b200/code_printer/cuda) prints one .cu file
 * per model from the sympy trees of F(t,y,p) and of the sparse J(t,y,p) value array, compiles it
 * with nvcc for sm_100a and libszb.so dlopen()s the result.  This mirrors what the reference's
 * python renderer emits as num_func.py (code_printer/python/module/module_generator.py:58-221):
 * F_ / inner_F<N> become szbm_launch_F, J_ / inner_J<N> + the COO->CSC conversion become
 * szbm_launch_J writing straight into CSC slot order.
 */
#ifndef SZB_MODEL_H
#define SZB_MODEL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SZBM_API __attribute__((visibility("default")))
#else
#define SZBM_API
#endif

#define SZBM_ABI_VERSION 1

typedef struct szbm_param_info {
  const char* name;
  int32_t dtype;        /* SZB_DTYPE_F64 / SZB_DTYPE_I32 */
  int32_t per_scenario; /* 1: [B][len], 0: [len] shared by all scenarios */
  int64_t len;
} szbm_param_info;

typedef struct szbm_args {
  int32_t B;              /* scenarios */
  double t;               /* time (DAE/FDAE), ignored by AE models */
  const double* t_b;      /* [B] per-scenario time, or NULL to use t */
  const double* Y;        /* [B][n] */
  double* F;              /* [B][n]       (launch_F) */
  double* JV;             /* [B][nnz_j]   (launch_J) */
  double* fnorm;          /* [B] max|F| per scenario, may be NULL (launch_F) */
  const void* const* params; /* device pointers, one per entry of the parameter table */
  const int32_t* active;  /* [B] or NULL: scenarios with active[b]==0 are skipped */
} szbm_args;

SZBM_API int szbm_abi_version(void);
SZBM_API int szbm_n(void);
SZBM_API int szbm_nnz_j(void);
SZBM_API int szbm_n_params(void);
SZBM_API const szbm_param_info* szbm_params(void);
/* both return a cudaError_t as int; they only enqueue work on `stream` */
SZBM_API int szbm_launch_F(const szbm_args* a, void* stream);
SZBM_API int szbm_launch_J(const szbm_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif
