/* szb.h -- C-ABI of the B200-native batched Newton / Rosenbrock backend for Solverz.
 *
 * One shared library, libszb.so, exports exactly the entry points below.  They are what a
 * binding inside the reference would call in place of the three native pieces on its hot
 * path (all paths relative to the Solverz 0.8.5 tree):
 *
 *   reference piece                                              replaced by
 *   ------------------------------------------------------------ ---------------------------
 *   generated F_/inner_F<N> (@njit)                              szb_eval_F
 *     code_printer/python/module/module_printer.py:683-870
 *   generated J_/inner_J<N> + coo_array(...).tocsc()             szb_eval_J
 *     module_printer.py:140-458, sym_algebra/symbols.py:297-298
 *   solve(A,b) -> scipy spsolve (SuperLU gssv)                   szb_lu_refactor + szb_lu_solve
 *     solvers/laesolver.py:11-15
 *   lu_decomposition(A) -> splu; .solve(b)                       szb_lu_refactor ; szb_lu_solve
 *     solvers/laesolver.py:18-45
 *   nr_method main loop                                          szb_newton / szb_newton_host
 *     solvers/nlaesolver/nr.py:45-64
 *   Rodas step loop                                              szb_rodas
 *     solvers/daesolver/rodas/rodas.py:123-327
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary.
 *   - every function returns 0 on success or a negative SZB_E* code and never throws; the
 *     text of the last error of a handle is returned by szb_last_error (thread-unsafe per
 *     handle, like the reference's module-global scratch buffers).
 *   - run-time numerical failure is soft, as in the reference (nr.py:50-52, rodas.py:127-135):
 *     it is reported per scenario in flags[] and never as an error code.
 *   - layouts: scenario-major, row-major.  Y[B][n], F[B][n], JV[B][nnz_j] (CSC slot order of
 *     the frozen pattern, jac.py:106-126 after duplicate summation), per-scenario parameter
 *     p[B][len], shared parameter p[len].
 *   - ownership: the caller owns every buffer passed to szb_bind and must keep it alive until
 *     it is re-bound or the handle is destroyed.  The library owns only the workspaces it
 *     allocates in szb_create / szb_set_batch and frees them in szb_destroy.
 *   - all kernels are stream-ordered on the stream given; the only host synchronisations are
 *     inside szb_newton / szb_rodas (reading the active-scenario count) and the *_host calls.
 */
#ifndef SZB_H
#define SZB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SZB_API __attribute__((visibility("default")))
#else
#define SZB_API
#endif

#define SZB_ABI_VERSION 4

/* error codes */
#define SZB_OK 0
#define SZB_EINVAL (-1)   /* bad argument / unbound slot                        */
#define SZB_ECUDA (-2)    /* CUDA runtime error, see szb_last_error             */
#define SZB_ENOMEM (-3)   /* host or device allocation failed                   */
#define SZB_EMODEL (-4)   /* generated model library missing / ABI mismatch     */
#define SZB_ESTATE (-5)   /* call sequence error (e.g. solve before refactor)   */

/* per-scenario status written to flags[] by szb_newton / szb_rodas */
#define SZB_FLAG_CONVERGED 0
#define SZB_FLAG_MAXIT 1      /* nr.py:50-52                                   */
#define SZB_FLAG_SINGULAR 2   /* zero / non-finite static pivot                */
#define SZB_FLAG_DTMIN 3      /* rodas.py:127-130                              */
#define SZB_FLAG_REJECTS 4    /* rodas.py:131-135                              */
#define SZB_FLAG_NAN 5        /* NaN/Inf residual                              */
#define SZB_FLAG_RUNNING 6    /* not finished (internal)                       */
#define SZB_FLAG_NOTCONV 7    /* loop left with max|F| == tol (not < tol)      */
#define SZB_FLAG_EVENT 8      /* Rodas stopped by a terminal event, rodas.py:267-271 (a success) */
#define SZB_MAX_EVENTS 8

/* szb_bind slots */
#define SZB_SLOT_Y 0        /* double[B][n]     unknowns, updated in place               */
#define SZB_SLOT_F 1        /* double[B][n]     residual                                 */
#define SZB_SLOT_JV 2       /* double[B][nnz_j] Jacobian values                          */
#define SZB_SLOT_RHS 3      /* double[B][n]     right-hand side / solution of lu_solve   */
#define SZB_SLOT_PARAM0 16  /* + index in the model's parameter table                    */

#define SZB_DTYPE_F64 0
#define SZB_DTYPE_I32 1

typedef struct szb_handle szb_handle;
typedef void* szb_stream; /* cudaStream_t */

/* Result of the host-side symbolic analysis (solverz_b200/lu/symbolic.py).  All arrays are
 * host pointers, copied to the device by szb_create.  Index arrays are int32. */
typedef struct szb_lu_symbolic {
  int32_t n, nnz_a, ns;
  int32_t lu_size, cb_size, cv_size; /* doubles per scenario: factors, contribution ws, fwd ws */
  int32_t n_warp_lists, n_group_lists, n_top;
  int32_t f_warp_max, f_group_max, f_top_max, s_warp_max, s_group_max, s_top_max;
  int32_t threads, wpg;     /* threads per CTA and warps per group the tables below were laid out for */
  int32_t n_sched, n_ga;    /* fronts in schedule order; (src, dst) pairs in the gather table */
  const int32_t* rowp;      /* [n]   pivot k <- original row     */
  const int32_t* colp;      /* [n]   pivot k <- original column  */
  const int32_t* sn_piv0;   /* [ns]  */
  const int32_t* sn_s;      /* [ns]  */
  const int32_t* sn_f;      /* [ns]  */
  const int32_t* idx_ptr;   /* [ns+1] */
  const int32_t* idx;       /* front index lists */
  const int32_t* rel_ptr;   /* [ns+1] */
  const int32_t* rel;
  const int32_t* child_ptr; /* [ns+1] */
  const int32_t* child;
  const int32_t* a_ptr;     /* [ns+1] */
  const int32_t* a_src;     /* [nnz_a] */
  const int32_t* a_dst;     /* [nnz_a] */
  const int32_t* lp_off;    /* [ns] */
  const int32_t* up_off;    /* [ns] */
  const int32_t* cb_off;    /* [ns] */
  const int32_t* ea_ptr;    /* [ns+1] extend-add tables: ea_dst[ea_ptr[c] + t] = position in the parent */
  const int32_t* ea_dst;    /*        front of element t of the u_c x (u_c+1) contribution block of c   */
  const int32_t* wl_ptr;    /* [n_warp_lists+1] */
  const int32_t* wl;
  const int32_t* gl_ptr;    /* [n_group_lists+1] */
  const int32_t* gl;
  const int32_t* top;       /* [n_top] */
  const int32_t* fh;        /* [n_sched][16] front descriptors in schedule order (warp lists, group lists,
                               top): s, f, piv0, lp_off, up_off, cb_off, gaJ, itJ, gaC, itC, gaR, itR, k, n_src,
                               idx_ptr, 0 */
  const int32_t* ga;        /* [n_ga][2] gather table: per front three [iterations][group threads] blocks
                               (J slots, contribution offsets, right-hand side) of (source, front position); see solverz_b200/lu/schedule.py */
} szb_lu_symbolic;

typedef struct szb_desc {
  int32_t abi_version;     /* SZB_ABI_VERSION */
  int32_t device;          /* CUDA device ordinal */
  int32_t n;               /* unknowns per scenario */
  int32_t nnz_j;           /* Jacobian value slots per scenario */
  const char* model_lib;   /* path of the generated model library (szbm_* symbols), or NULL for
                              a handle that only factors / solves */
  szb_lu_symbolic lu;
  const double* mass_vals; /* [nnz_j] mass matrix M scattered on the J pattern (DAE), or NULL */
} szb_desc;

/* options of szb_rodas -- field meanings follow solvers/option.py:4-51; the tableau arrays are the
 * derived coefficients of Rodas_param (solvers/daesolver/rodas/param.py:4-202) in its final layout:
 * alpha[l*s + j] and gammatilde[l*s + j] multiply stage l in stage j (param.py stores the transposes) */
typedef struct szb_rodas_opt {
  double t0, tend;
  double rtol, atol, f_savety, fac1, fac2, facmax;
  double hinit;  /* <= 0: 1e-6*(tend-t0), rodas.py:110-117 */
  double hmax;   /* <= 0: |tend - t0|, rodas.py:76-77 */
  int32_t fix_h;
  int32_t dfdt_mode;  /* 0: dF/dt == 0, what the reference's rendered-module path computes because its
                         two F calls alias one buffer; 1: forward difference
                         (rodas.py:347-352), what the inline made_numerical path computes */
  int32_t max_steps;  /* safety bound on accepted+rejected steps per scenario (0: 1000000) */
  int32_t n_out;      /* dense-output times (rodas.py:273-303); 0: none */
  const double* t_out; /* host [n_out], increasing, t_out[0] == t0 */
  int32_t record_cap; /* > 0: record every accepted step (what the reference returns for a
                         two-point tspan), at most record_cap rows per scenario */
  /* tableau */
  int32_t s, pord;
  double gamma;
  const double *alpha, *gammatilde; /* host [s*s] */
  const double *a, *g, *b, *bd, *c, *d, *e; /* host [s] */
  /* Device-side events: the reference's events(t, y) callback (rodas.py:99-107,
   * 200-271) restricted to component thresholds, value_i = y[ev_idx[i]] - ev_thr[i], so that detection,
   * the secant start + bisection on the dense output ntrp1 and the roll-back to the event time run per
   * scenario inside the step kernel.  n_events == 0: none.  Needs a scheme with ntrp1 (c, d, e given). */
  int32_t n_events;          /* <= SZB_MAX_EVENTS */
  const int32_t* ev_idx;     /* host [n_events] */
  const double* ev_thr;      /* host [n_events] */
  const int32_t* ev_dir;     /* host [n_events]: -1 falling, +1 rising, 0 both */
  const int32_t* ev_term;    /* host [n_events]: 1 = the scenario stops at this event */
  double event_duration;     /* events closer than this to the step start are dropped (rodas.py:257-259) */
  int32_t ev_cap;            /* events recorded per scenario (rows of ev_t / ev_i / ev_y) */
  /* Step-wise driving, for event functions that are Python callbacks (rodas.py:99-107: value, isterminal,
   * direction = events(t, y)): the host has to see every accepted step.  resume = 1 keeps the controller state
   * of the previous call on this handle (t, dt, facmax, reject counter, dense-output cursor, statistics) instead
   * of initialising from t0 / hinit; stop_after_accepted > 0 pauses a scenario (it stays SZB_FLAG_RUNNING) after
   * that many accepted steps of this call.  Between calls the host may move a scenario back to an event time
   * with szb_rodas_state_set and by writing its state into SLOT_Y. */
  int32_t resume;
  int32_t stop_after_accepted;
} szb_rodas_opt;

typedef struct szb_rodas_out {
  double* y_out;     /* device [n_out][B][n] or NULL */
  double* t_end;     /* device [B]: time reached */
  int32_t* stats;    /* device [B][8]: nstep, nfeval, nJeval, ndecomp, nreject, nsolve, flag, n_recorded */
  double* y_steps;   /* device [record_cap][B][n] or NULL */
  double* t_steps;   /* device [record_cap][B] or NULL */
  int32_t* ev_count; /* device [B]: events located (may exceed ev_cap; only the first ev_cap are stored) */
  double* ev_t;      /* device [B][ev_cap] */
  int32_t* ev_i;     /* device [B][ev_cap] */
  double* ev_y;      /* device [B][ev_cap][n] or NULL */
  int32_t* n_rows;   /* device [B] or NULL: rows of y_out filled (< n_out after a terminal event) */
} szb_rodas_out;

SZB_API int szb_abi_version(void);
SZB_API const char* szb_last_error(const szb_handle* h); /* h == NULL: error of the last failed create */

SZB_API int szb_create(const szb_desc* desc, szb_handle** out);
SZB_API int szb_destroy(szb_handle* h);
SZB_API int szb_set_batch(szb_handle* h, int32_t B);
SZB_API int szb_bind(szb_handle* h, int32_t slot, void* dev_ptr, int64_t len, int32_t dtype);

/* residual F(t, y, p) of every scenario into SLOT_F; max|F| per scenario is kept internally */
SZB_API int szb_eval_F(szb_handle* h, double t, szb_stream stream);
/* Jacobian values into SLOT_JV */
SZB_API int szb_eval_J(szb_handle* h, double t, szb_stream stream);
/* numeric refactorisation of  A = J            (h_gamma == 0)
 *                             A = M - h_gamma*J (h_gamma != 0; M from desc.mass_vals) */
SZB_API int szb_lu_refactor(szb_handle* h, double h_gamma, szb_stream stream);
/* per-scenario status of the last szb_lu_refactor / szb_lu_refactor_masked_rows into flags_dev (device
 * int32[B]): 0 ok, SZB_FLAG_SINGULAR = a zero / non-finite pivot or pivot growth beyond the monitor's bound
 * under the static pivot order -- what the reference reports as SuperLU's RuntimeError("Factor is exactly
 * singular") / MatrixRankWarning (laesolver.py:11-45 -> scipy linsolve.py) */
SZB_API int szb_lu_status(szb_handle* h, int32_t* flags_dev, szb_stream stream);
/* factor storage of scenario b of the last szb_lu_refactor into out_dev (device double[SZB_Q_LU_SIZE]): per
 * supernode the f x s panel [L11\U11; L21] (column-major, 1/u_pp on the diagonal) followed by the s x u panel
 * U12 -- what the host needs to present the reference object's .L / .U (laesolver.py:37-42) */
SZB_API int szb_lu_factors(szb_handle* h, int32_t b, double* out_dev, szb_stream stream);
/* a-posteriori check of a solve under the static pivot order:  relres_dev[b] = max|rhs_b - A_b x_b| /
 * max|rhs_b|  with A = J (h_gamma == 0) or M - h_gamma*J, assembled slot by slot from SLOT_JV
 * (slot_row_dev / slot_col_dev: device int32[nnz_j], row and column of every CSC slot).  rhs_dev / x_dev:
 * device double[B][n] in the original row / column order.  SuperLU's partial pivoting, which the reference
 * gets per call (laesolver.py:11-15), bounds this residual; a fixed pivot sequence does not -- callers
 * re-analyse the scenarios whose residual is large with their own values (solvers/laesolver.py, nr.py). */
SZB_API int szb_lu_residual(szb_handle* h, double h_gamma, const int32_t* slot_row_dev, const int32_t* slot_col_dev,
                            const double* rhs_dev, const double* x_dev, double* relres_dev, szb_stream stream);
/* SLOT_RHS <- A^{-1} SLOT_RHS for nrhs right-hand sides stored as [nrhs][B][n] */
SZB_API int szb_lu_solve(szb_handle* h, int32_t nrhs, szb_stream stream);
/* whole Newton solve, nr.py:45-64, on device buffers; iters/flags are device int32[B] */
SZB_API int szb_newton(szb_handle* h, double tol, int32_t max_it, int32_t* iters_dev, int32_t* flags_dev,
               szb_stream stream);
/* time t at which szb_newton evaluates F(t,y,p) and J(t,y,p) (0 by default); fdae_solver sets t0 + dt
 * before each step (fdesolver.py:104-108) */
SZB_API int szb_set_time(szb_handle* h, double t);
/* Implicit one-step mode of szb_newton: the residual and matrix of the algebraic system that
 * backward_euler (beuler.py:66-72) and implicit_trapezoid (trapezoidal.py:65-73) hand to nr_method,
 *     G(y) = M (y - y_prev) - c (F(t, y, p) + kappa F_prev),      A = M - c J(t, y, p),
 * mode 1: backward Euler (c = dt, kappa = 0, f_prev_dev may be NULL), mode 2: trapezoid (c = dt/2, kappa = 1;
 * the reference's residual is -G and its matrix -A: same iterates), mode 0: back to F(y) = 0.
 * y_prev_dev / f_prev_dev: caller-owned device double[B][n], read during szb_newton; t comes from
 * szb_set_time; needs szb_set_mass_rows. */
SZB_API int szb_set_implicit(szb_handle* h, int32_t mode, double c, const double* y_prev_dev,
                     const double* f_prev_dev);
/* same with HOST buffers: copies y0 and the per-scenario parameters in, solves, copies y,
 * iters and flags out.  params_host[i] may be NULL to keep what is already on the device. */
SZB_API int szb_newton_host(szb_handle* h, double tol, int32_t max_it, double* y_host,
                    const void* const* params_host, int32_t n_params, int32_t* iters_host,
                    int32_t* flags_host, szb_stream stream);
/* batched Rodas (rodas.py:7-344): integrates every scenario of SLOT_Y from t0 to tend with its own
 * adaptive step; SLOT_Y holds the final states on return. */
SZB_API int szb_rodas(szb_handle* h, const szb_rodas_opt* opt, const szb_rodas_out* out, szb_stream stream);
/* controller state after a szb_rodas call, host arrays of B entries each (NULL: skipped): time reached, start
 * and size of the last accepted step, dense-output cursor (index of the next requested time), the [B][8]
 * statistics block of szb_rodas_out.stats */
SZB_API int szb_rodas_state_get(szb_handle* h, double* t_host, double* told_host, double* dt_host,
                                int32_t* inext_host, int32_t* stats_host);
/* scenario b continues from time t (its state: SLOT_Y); inext >= 0 rewinds the dense-output cursor; flag >= 0
 * sets its status (SZB_FLAG_EVENT stops it, SZB_FLAG_RUNNING lets a resume call continue it) */
SZB_API int szb_rodas_state_set(szb_handle* h, int32_t b, double t, int32_t inext, int32_t flag);
/* stage vectors K [s][B][n] of the last attempted step into K_out_dev (device double[s*B*n]; s = stages of the
 * scheme of the last call) -- the data of the dense output ntrp1 (rodas.py:383-386) */
SZB_API int szb_rodas_stages(szb_handle* h, double* K_out_dev, szb_stream stream);
/* row_col[r] = column of the unit entry of M in row r, or -1 for an algebraic row (host int32[n]);
 * needed once before szb_rodas */
SZB_API int szb_set_mass_rows(szb_handle* h, const int32_t* row_col, int32_t n);
/* A = M + diag(row_mask) J  (row_mask_dev: device double[nnz_j] with 0/1 per Jacobian slot, or NULL
 * for all ones) -- the matrix of the consistent-initialisation Newton of daeic.py:24-68 */
SZB_API int szb_lu_refactor_masked_rows(szb_handle* h, const double* row_mask_dev, szb_stream stream);

/* introspection */
SZB_API int szb_query(const szb_handle* h, int32_t what, int64_t* out);
#define SZB_Q_N 0
#define SZB_Q_NNZ_J 1
#define SZB_Q_LU_SIZE 2
#define SZB_Q_N_PARAMS 3
#define SZB_Q_BATCH 4
#define SZB_Q_LAUNCHES 5       /* kernels launched by this handle since creation */
#define SZB_Q_WORKSPACE_BYTES 6
/* per-kernel-class device time of the last szb_newton call, milliseconds, from CUDA event pairs
 * recorded on the launching stream around every launch (no extra synchronisation):
 * out5[0..4] = F, J, LU refactor, solve+update, bookkeeping; launches5 = launches per class since
 * the handle was created; units5 = scenarios processed per class in the last call.  Only filled
 * when profiling was switched on with szb_set_profile(h, 1). */
SZB_API int szb_set_profile(szb_handle* h, int32_t on);
SZB_API int szb_get_profile(const szb_handle* h, double* out5, int64_t* launches5, int64_t* units5);

#ifdef __cplusplus
}
#endif
#endif /* SZB_H */
