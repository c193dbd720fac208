// szb_rodas.cu -- batched stiffly-accurate Rosenbrock (Rodas) driver.
//
// Restates Solverz/solvers/daesolver/rodas/rodas.py:123-327 per scenario: every scenario carries its
// own t, dt, reject counter and facmax; all active scenarios advance in lock-step over the STAGES of
// one step attempt (one batched F launch, one batched triangular solve per stage; one batched J
// launch and one batched refactorisation of M - dt*gamma*J per attempt), accepted / rejected
// individually.  Line references below are to rodas.py unless stated otherwise.
#include <cfloat>
#include <cmath>
#include <vector>

#include "szb.h"
#include "szb_internal.cuh"

namespace {

constexpr int kMaxStages = 8;

struct Tableau {
  int s, pord;
  double gamma;
  double alpha[kMaxStages * kMaxStages];   // alpha[l*s + j]: weight of K_l in stage j
  double gt[kMaxStages * kMaxStages];      // gammatilde, same layout
  double a[kMaxStages], g[kMaxStages], b[kMaxStages], bd[kMaxStages], c[kMaxStages], d[kMaxStages],
      e[kMaxStages];
};

struct Ctl {  // scalars of the step controller
  double tend, rtol, atol, f_savety, fac1, fac2, hmax, hmin, uround, hinit;
  int fix_h, n_out, record_cap, max_steps;
  int hermite;      // dense output by the cubic Hermite form ntrp2 (rodas3: no interpolation coefficients)
  int stop_after;   // > 0: a scenario pauses after this many accepted steps of the current call (szb_rodas_opt)
  // component-threshold events (szb.h): value_i = y[ev_idx[i]] - ev_thr[i]
  int n_events, ev_cap;
  double event_duration;
  int ev_idx[SZB_MAX_EVENTS], ev_dir[SZB_MAX_EVENTS], ev_term[SZB_MAX_EVENTS];
  double ev_thr[SZB_MAX_EVENTS];
};

// per-scenario state, struct of arrays
struct State {
  double *t, *dt, *told, *facmax, *scale, *tstage, *ddt, *dtacc;   // dtacc: step size of the last accepted step
  int *reject, *active, *jmask, *inext, *nrec, *stats;  // stats: [B][8]
  int* acc_left;   // accepted steps this scenario may still take in the current call (stop_after_accepted)
  int* acc_now;    // 1: this attempt was accepted (ntrp2 dense output runs after the F evaluation at the new point)
  double* ev_val;   // [B][SZB_MAX_EVENTS] event values at the end of the last accepted step (rodas.py:203)
};

// A scenario whose M - dt*gamma*J met a zero / non-finite static pivot stops with SZB_FLAG_SINGULAR -- the
// per-scenario form of the RuntimeError that lu_decomposition raises inside the reference's step (rodas.py:159-162)
__global__ void rodas_lu_check_kernel(State S, const int* __restrict__ lu_flags, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && S.active[b] && lu_flags[b] == SZB_FLAG_SINGULAR) {
    S.active[b] = 0;
    S.stats[8 * b + 6] = SZB_FLAG_SINGULAR;
  }
}

// rodas.py:127-146 -- failure exits, end stretch, fixed step; prepares LU scale and the J mask
__global__ void step_begin_kernel(State S, Ctl C, double gamma, int* __restrict__ n_active, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = 0;
  if (b < B && !S.active[b]) S.jmask[b] = 0;
  if (b < B && S.active[b]) {
    double dt = S.dt[b];
    const double t = S.t[b];
    int* st = S.stats + 8 * b;
    if (fabs(dt) < C.uround) {                       // :127-130
      S.active[b] = 0; st[6] = SZB_FLAG_DTMIN;
    } else if (S.reject[b] > 100) {                  // :132-135
      S.active[b] = 0; st[6] = SZB_FLAG_REJECTS;
    } else if (st[0] + st[4] >= C.max_steps) {
      S.active[b] = 0; st[6] = SZB_FLAG_MAXIT;
    } else {
      if (t + dt >= C.tend) dt = C.tend - t;         // :138-141
      else dt = fmin(dt, 0.5 * (C.tend - t));
      if (C.fix_h) dt = C.hinit;                     // :143-144
      S.dt[b] = dt;
      S.scale[b] = -dt * gamma;                      // A = M - dt*gamma*J, :159
      S.jmask[b] = (S.reject[b] == 0) ? 1 : 0;       // :150-152
      if (S.jmask[b]) st[2] += 1;
      st[3] += 1;                                    // ndecomp, :162
      const double tscale = fmax(0.1 * fabs(t), 1e-8);   // dfdt, :348-349
      S.ddt[b] = t + sqrt(DBL_EPSILON) * tscale - t;
      S.tstage[b] = t + S.ddt[b];
      mine = 1;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_active, __popc(m));
}

// dfdt0 = dt * (F(t+ddt) - F(t)) / ddt          (:154, :347-352)
__global__ void dfdt_kernel(const double* __restrict__ F0, const double* __restrict__ F1,
                            double* __restrict__ dfdt0, State S, int n, int B) {
  const int b = blockIdx.y;
  if (!S.active[b]) return;
  const double f = S.dt[b] / S.ddt[b];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = static_cast<size_t>(b) * n + i;
    dfdt0[q] = f * (F1[q] - F0[q]);
  }
}

// rhs = F + g_j * dfdt0 (+ M @ sum_2 for j > 0)    (:155, :171)
__global__ void stage_rhs_kernel(double* __restrict__ rhs, const double* __restrict__ F,
                                 const double* __restrict__ dfdt0, const double* __restrict__ sum2,
                                 const int* __restrict__ mcol, double gj, State S, int n, int B) {
  const int b = blockIdx.y;
  if (!S.active[b]) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = static_cast<size_t>(b) * n + i;
    double v = F[q];
    if (dfdt0) v += gj * dfdt0[q];
    if (sum2) {
      const int c = mcol[i];
      if (c >= 0) v += sum2[static_cast<size_t>(b) * n + c];
    }
    rhs[q] = v;
  }
}

// sum_1 = K @ alpha[:,j], sum_2 = K @ gammatilde[:,j], y1 = y0 + dt*sum_1, tstage = t + dt*a_j (:166-171)
__global__ void stage_prep_kernel(const double* __restrict__ K, const double* __restrict__ y0,
                                  double* __restrict__ y1, double* __restrict__ sum2, Tableau T, int j,
                                  State S, int n, int B) {
  const int b = blockIdx.y;
  if (!S.active[b]) return;
  const double dt = S.dt[b];
  if (blockIdx.x == 0 && threadIdx.x == 0) S.tstage[b] = S.t[b] + dt * T.a[j];
  const size_t stride = static_cast<size_t>(B) * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = static_cast<size_t>(b) * n + i;
    double s1 = 0.0, s2 = 0.0;
    for (int l = 0; l < j; ++l) {
      const double k = K[l * stride + q];
      s1 = fma(k, T.alpha[l * T.s + j], s1);
      s2 = fma(k, T.gt[l * T.s + j], s2);
    }
    y1[q] = y0[q] + dt * s1;
    sum2[q] = s2;
  }
}

// K_j = sol - sum_2   (:173); sum2 == null for stage 0
__global__ void stage_post_kernel(double* __restrict__ Kj, const double* __restrict__ sol,
                                  const double* __restrict__ sum2, State S, int n, int B) {
  const int b = blockIdx.y;
  if (!S.active[b]) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = static_cast<size_t>(b) * n + i;
    Kj[q] = sum2 ? sol[q] - sum2[q] : sol[q];
  }
}

__device__ __forceinline__ double nanmax(double a, double b) { return (a > b || a != a) ? a : b; }

// np.spacing(x): distance to the next double away from zero
__device__ __forceinline__ double spacing(const double x) {
  const double a = fabs(x);
  return __longlong_as_double(__double_as_longlong(a) + 1) - a;
}

// error estimate, step-size controller, accept / reject, dense output (:176-327); one CTA per scenario
__global__ void __launch_bounds__(256)
step_finish_kernel(const double* __restrict__ K, double* __restrict__ Y, Tableau T, Ctl C, State S,
                   szb_rodas_out O, const double* __restrict__ t_out, int n, int B) {
  __shared__ double red[8];
  __shared__ int s_bad;
  __shared__ int s_accept;
  __shared__ double s_told, s_tnew, s_dt;
  __shared__ int s_nev, s_ev_i[SZB_MAX_EVENTS];
  __shared__ double s_tau, s_ev_tau[SZB_MAX_EVENTS], s_ev_t[SZB_MAX_EVENTS];
  const int b = blockIdx.x;
  if (!S.active[b]) return;
  const size_t stride = static_cast<size_t>(B) * n;
  double* y = Y + static_cast<size_t>(b) * n;
  const double dt = S.dt[b];
  int* st = S.stats + 8 * b;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  double emax = 0.0;
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t q = static_cast<size_t>(b) * n + i;
    double s1 = 0.0, s2 = 0.0;
    for (int l = 0; l < T.s; ++l) {
      const double k = K[l * stride + q];
      s1 = fma(k, dt * T.b[l], s1);                 // :175
      s2 = fma(k, dt * T.bd[l], s2);                // :178
    }
    const double ynew = y[i] + s1;                  // :176
    if (isinf(ynew) || isnan(ynew)) bad = 1;        // :182
    const double sk = C.atol + C.rtol * fabs(ynew); // :179
    emax = nanmax(emax, fabs((s1 - s2) / sk));      // :180 (np.max propagates NaN)
  }
  for (int o = 16; o > 0; o >>= 1) emax = nanmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if (bad) s_bad = 1;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = emax;
  __syncthreads();
  if (threadIdx.x == 0) {
    double err = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) err = nanmax(err, red[w]);
    double dtnew;
    if (!C.fix_h) {
      if (s_bad) err = 1.0e6;                       // :182-184
      if (err != err) err = 1.0e6;                  // NaN estimate: treated like the NaN/Inf case
      err = fmax(err, 1.0e-6);                      // :185
      double fac = C.f_savety / pow(err, 1.0 / T.pord);       // :186
      fac = fmin(S.facmax[b], fmax(C.fac1, fac));             // :187
      dtnew = dt * fac;                                       // :188
    } else {
      err = 1.0;
      dtnew = dt;
    }
    st[1] += T.s;   // nfeval: one per stage (:156, :172)
    st[5] += T.s;   // nsolve (:174)
    const int accept = err <= 1.0;
    s_accept = accept;
    s_told = S.t[b];
    s_dt = dt;
    if (accept) {                                   // :195-199
      S.reject[b] = 0;
      S.told[b] = S.t[b];
      S.t[b] = S.t[b] + dt;
      S.dtacc[b] = dt;
      st[0] += 1;
      S.facmax[b] = C.fac2;                         // :321
      if (fabs(C.tend - S.t[b]) < C.uround) {       // :318
        S.active[b] = 0;
        st[6] = SZB_FLAG_CONVERGED;
      } else if (C.stop_after > 0 && --S.acc_left[b] <= 0) {
        S.active[b] = 0;                            // paused: still SZB_FLAG_RUNNING, a resume call continues it
      }
    } else {                                        // :323-326
      S.reject[b] += 1;
      st[4] += 1;
      S.facmax[b] = 1.0;
    }
    s_tnew = S.t[b];
    S.dt[b] = fmin(C.hmax, fmax(C.hmin, dtnew));    // :327
    s_nev = 0;
    s_tau = 1.0;
    // ---- events (:200-271), component thresholds; the dense output of ONE component is a scalar formula
    // in tau, so the whole secant + bisection search runs on this thread
    if (accept && C.n_events > 0) {
      const double told = s_told;
      double tcur = s_tnew;                           // `t` of the reference: rolls back to the event time
      double* vold = S.ev_val + static_cast<size_t>(b) * SZB_MAX_EVENTS;
      double vsave[SZB_MAX_EVENTS];
      for (int i = 0; i < C.n_events; ++i) {          // value = events(t, ynew), ynew = y0 + K (dt b)  (:175-176, :203)
        const int c = C.ev_idx[i];
        double s1 = 0.0;
        for (int l = 0; l < T.s; ++l) s1 = fma(K[l * stride + static_cast<size_t>(b) * n + c], dt * T.b[l], s1);
        vsave[i] = (y[c] + s1) - C.ev_thr[i];
      }
      int stop = 0;
      for (int i = 0; i < C.n_events && !stop; ++i) {
        double v0 = vold[i], v1 = vsave[i];
        if (!(v1 * v0 < 0.0)) continue;               // ff = where(value * valueold < 0)
        if (C.ev_dir[i] < 0 && v0 <= v1) continue;
        if (C.ev_dir[i] > 0 && v0 >= v1) continue;
        const int c = C.ev_idx[i];
        const size_t qc = static_cast<size_t>(b) * n + c;
        auto comp = [&](const double tau) {           // ntrp1 of component c (:383-386)
          double acc = 0.0;
          for (int l = 0; l < T.s; ++l)
            acc += K[l * stride + qc] * (T.b[l] + (tau - 1.0) * (T.c[l] + tau * (T.d[l] + tau * T.e[l])));
          return y[c] + tau * dt * acc;
        };
        int iterate = 1;
        double tL = told, tR = tcur, tevent;
        if (fabs(v1 - v0) > C.uround) tevent = told - v0 * dt / (v1 - v0);     // :222-223
        else { iterate = 0; tevent = tcur; }
        double tol = 128.0 * fmax(spacing(told), spacing(tcur));                 // np.spacing, :228
        tol = fmin(tol, fabs(tcur - told));
        while (iterate > 0) {                          // :231-254
          ++iterate;
          const double vi = comp((tevent - told) / dt) - C.ev_thr[i];
          if (v1 * vi < 0.0) { tL = tevent; tevent = 0.5 * (tevent + tR); v0 = vi; }
          else if (v0 * vi < 0.0) { tR = tevent; tevent = 0.5 * (tL + tevent); v1 = vi; }
          else iterate = 0;
          if ((tR - tL) < tol) iterate = 0;
          if (iterate > 100) break;                    // "Lost Event"
        }
        if (fabs(tevent - told) < C.event_duration) break;                      // :257-259
        tcur = tevent;
        s_ev_tau[s_nev] = (tevent - told) / dt;
        s_ev_t[s_nev] = tevent;
        s_ev_i[s_nev] = i;
        ++s_nev;
        if (C.ev_term[i]) stop = 1;                    // :267-271
      }
      for (int i = 0; i < C.n_events; ++i) vold[i] = vsave[i];                  // value = value_save (:266)
      if (s_nev > 0) {
        s_tau = s_ev_tau[s_nev - 1];
        s_tnew = tcur;
        S.t[b] = tcur;
        if (stop) { S.active[b] = 0; st[6] = SZB_FLAG_EVENT; }
        else if (st[6] == SZB_FLAG_CONVERGED && fabs(C.tend - tcur) >= C.uround) { S.active[b] = 1; st[6] = SZB_FLAG_RUNNING; }
      }
    }
  }
  __syncthreads();
  if (!s_accept) return;
  // states at the located events (ye) and their bookkeeping (:262-265)
  if (s_nev > 0) {
    const int base = O.ev_count[b];
    for (int k = 0; k < s_nev; ++k) {
      const int slot = base + k;
      if (slot >= C.ev_cap) break;
      const double tau = s_ev_tau[k];
      double w[kMaxStages];
      for (int l = 0; l < T.s; ++l)
        w[l] = tau * s_dt * (T.b[l] + (tau - 1.0) * (T.c[l] + tau * (T.d[l] + tau * T.e[l])));
      if (O.ev_y) {
        double* dst = O.ev_y + (static_cast<size_t>(b) * C.ev_cap + slot) * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const size_t q = static_cast<size_t>(b) * n + i;
          double acc = 0.0;
          for (int l = 0; l < T.s; ++l) acc += K[l * stride + q] * w[l];
          dst[i] = y[i] + acc;
        }
      }
      if (threadIdx.x == 0) {
        O.ev_t[static_cast<size_t>(b) * C.ev_cap + slot] = s_ev_t[k];
        O.ev_i[static_cast<size_t>(b) * C.ev_cap + slot] = s_ev_i[k];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) O.ev_count[b] = base + s_nev;
  }
  if (C.hermite && threadIdx.x == 0) S.acc_now[b] = 1;   // ntrp2 needs F at the new point: hermite_output_kernel
  // dense output at the requested times inside (told, t]   (:273-303, ntrp1 :383-386)
  if (C.n_out > 0 && !C.hermite) {
    int inext = S.inext[b];
    while (inext < C.n_out) {
      const double tn = t_out[inext];
      if (!(s_tnew >= tn && tn > s_told)) break;
      const double tau = (tn - s_told) / s_dt;
      double w[kMaxStages];
      for (int l = 0; l < T.s; ++l)
        w[l] = tau * s_dt * (T.b[l] + (tau - 1.0) * (T.c[l] + tau * (T.d[l] + tau * T.e[l])));
      double* dst = O.y_out + (static_cast<size_t>(inext) * B + b) * n;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t q = static_cast<size_t>(b) * n + i;
        double acc = y[i];
        for (int l = 0; l < T.s; ++l) acc = fma(K[l * stride + q], w[l], acc);
        dst[i] = acc;
      }
      ++inext;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      S.inext[b] = inext;
      if (O.n_rows) O.n_rows[b] = inext;
    }
  }
  // y0 = ynew (:320); after an event ynew = ynext = ntrp1 at the event time (:261)
  if (s_nev > 0) {
    const double tau = s_tau;
    double w[kMaxStages];
    for (int l = 0; l < T.s; ++l)
      w[l] = tau * s_dt * (T.b[l] + (tau - 1.0) * (T.c[l] + tau * (T.d[l] + tau * T.e[l])));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const size_t q = static_cast<size_t>(b) * n + i;
      double acc = 0.0;
      for (int l = 0; l < T.s; ++l) acc += K[l * stride + q] * w[l];
      y[i] += acc;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const size_t q = static_cast<size_t>(b) * n + i;
      double s1 = 0.0;
      for (int l = 0; l < T.s; ++l) s1 = fma(K[l * stride + q], s_dt * T.b[l], s1);
      y[i] += s1;
    }
  }
  // record the accepted step (two-point tspan behaviour of the reference, :305-309)
  if (C.record_cap > 0) {
    __syncthreads();
    const int r = S.nrec[b];
    if (r < C.record_cap) {
      double* dst = O.y_steps + (static_cast<size_t>(r) * B + b) * n;
      for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = y[i];
      if (threadIdx.x == 0) {
        O.t_steps[static_cast<size_t>(r) * B + b] = s_tnew;
        S.nrec[b] = r + 1;
        st[7] = r + 1;
      }
    }
  }
}

__global__ void rodas_init_kernel(State S, szb_rodas_out O, const double* __restrict__ Y, double t0,
                                  double dt0, double facmax, int n_out, int record_cap, Ctl C, int n, int B) {
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    if (C.n_events > 0) {                              // value = events(t0, y0), rodas.py:102-103
      for (int i = 0; i < C.n_events; ++i)
        S.ev_val[static_cast<size_t>(b) * SZB_MAX_EVENTS + i] = Y[static_cast<size_t>(b) * n + C.ev_idx[i]] - C.ev_thr[i];
      O.ev_count[b] = 0;
    }
    if (O.n_rows) O.n_rows[b] = n_out > 0 ? 1 : 0;
    S.t[b] = t0; S.told[b] = t0; S.dt[b] = dt0; S.facmax[b] = facmax;
    S.reject[b] = 0; S.active[b] = 1; S.jmask[b] = 1; S.inext[b] = n_out > 0 ? 1 : 0;
    S.nrec[b] = record_cap > 0 ? 1 : 0;
    int* st = S.stats + 8 * b;
    for (int i = 0; i < 8; ++i) st[i] = 0;
    st[6] = SZB_FLAG_RUNNING;
    st[7] = record_cap > 0 ? 1 : 0;
    if (record_cap > 0) O.t_steps[b] = t0;
  }
  const double* y = Y + static_cast<size_t>(b) * n;
  if (n_out > 0) {
    double* dst = O.y_out + static_cast<size_t>(b) * n;   // row 0 = y(t0)
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = y[i];
  }
  if (record_cap > 0) {
    double* dst = O.y_steps + static_cast<size_t>(b) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = y[i];
  }
}

// ---- rodas3 dense output (ntrp2, rodas.py:389-393): cubic Hermite over the accepted step from y0, y1 and
// F0 = F(told, y0), F1 = F(told + dt, y1).  The step start (y0, F0) is saved before the stages overwrite F; after an
// accepted step F is evaluated once more at the new point (only in this mode), then the requested times inside
// (told, t] are written.  alias_F = 1 reproduces the reference's rendered-module path, where both F calls of ntrp
// return the module's one shared buffer, so F0 IS F1 there (the same aliasing as in dfdt).
__global__ void hermite_save_kernel(const double* __restrict__ Y, const double* __restrict__ F, double* __restrict__ Y0s,
                                    double* __restrict__ F0s, State S, int n, int B) {
  const int b = blockIdx.y;
  if (!S.active[b]) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) S.acc_now[b] = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t q = static_cast<size_t>(b) * n + i;
    Y0s[q] = Y[q];
    F0s[q] = F[q];
  }
}

__global__ void __launch_bounds__(256)
hermite_output_kernel(const double* __restrict__ Y0s, const double* __restrict__ F0s, const double* __restrict__ Y1,
                      const double* __restrict__ F1, State S, Ctl C, szb_rodas_out O, const double* __restrict__ t_out,
                      int alias_F, int n, int B) {
  const int b = blockIdx.x;
  if (!S.acc_now[b]) return;
  const double told = S.told[b], dt = S.dtacc[b], tnew = S.t[b];
  const size_t o = static_cast<size_t>(b) * n;
  int inext = S.inext[b];
  while (inext < C.n_out) {
    const double tn = t_out[inext];
    if (!(tnew >= tn && tn > told)) break;                       // while t >= tnext > told (:274)
    const double tau = (tn - told) / dt;
    double* dst = O.y_out + (static_cast<size_t>(inext) * B + b) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double y0 = Y0s[o + i], y1 = Y1[o + i], f1 = F1[o + i];
      const double f0 = alias_F ? f1 : F0s[o + i];
      dst[i] = (1.0 - tau) * y0 + tau * y1 +
               tau * (tau - 1.0) * ((1.0 - 2.0 * tau) * (y1 - y0) + (tau - 1.0) * dt * f0 + tau * dt * f1);
    }
    ++inext;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    S.inext[b] = inext;
    if (O.n_rows) O.n_rows[b] = inext;
    S.acc_now[b] = 0;
  }
}

// resume = 1: the controller state of the previous call stays; scenarios that were only paused run again
__global__ void rodas_resume_kernel(State S, int stop_after, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    S.active[b] = S.stats[8 * b + 6] == SZB_FLAG_RUNNING ? 1 : 0;
    S.acc_left[b] = stop_after;
  }
}

__global__ void rodas_finalize_kernel(State S, double* t_end, int* stats_out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    if (t_end) t_end[b] = S.t[b];
    if (stats_out)
      for (int i = 0; i < 8; ++i) stats_out[8 * b + i] = S.stats[8 * b + i];
  }
}

}  // namespace

extern "C" int szb_rodas(szb_handle* h, const szb_rodas_opt* o, const szb_rodas_out* out, szb_stream stream) {
  if (!h || !o || !out) return SZB_EINVAL;
  if (!h->launch_F || h->B <= 0 || !h->has_lu) { szb_set_error(h, "szb_rodas: handle needs a model, an LU schedule and a batch"); return SZB_ESTATE; }
  if (!h->mass_vals) { szb_set_error(h, "szb_rodas: the model has no mass matrix (not a DAE)"); return SZB_EINVAL; }
  if (o->s < 1 || o->s > kMaxStages) { szb_set_error(h, "szb_rodas: unsupported stage count %d", o->s); return SZB_EINVAL; }
  if (o->t0 > o->tend) { szb_set_error(h, "t0: %g > tend: %g", o->t0, o->tend); return SZB_EINVAL; }
  if (o->n_out > 0 && (!o->t_out || !out->y_out)) { szb_set_error(h, "szb_rodas: n_out without t_out / y_out"); return SZB_EINVAL; }
  if (o->record_cap > 0 && (!out->y_steps || !out->t_steps)) { szb_set_error(h, "szb_rodas: record_cap without buffers"); return SZB_EINVAL; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = h->B, n = h->n, s = o->s;
  const size_t bn = static_cast<size_t>(B) * n;

  // ---- workspaces (kept on the handle between calls)
  if (h->rodas_B != B || h->rodas_stages < s) {
    for (void* p : h->rodas_owned) cudaFree(p);
    h->rodas_owned.clear();
    auto alloc = [&](size_t bytes) -> void* {
      void* p = nullptr;
      if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
      h->rodas_owned.push_back(p);
      h->ws_bytes += static_cast<int64_t>(bytes);
      return p;
    };
    h->rd_K = static_cast<double*>(alloc(sizeof(double) * bn * s));
    h->rd_y1 = static_cast<double*>(alloc(sizeof(double) * bn));
    h->rd_sum2 = static_cast<double*>(alloc(sizeof(double) * bn));
    h->rd_dfdt = static_cast<double*>(alloc(sizeof(double) * bn));
    h->rd_F1 = static_cast<double*>(alloc(sizeof(double) * bn));
    h->rd_y0s = nullptr;
    h->rd_F0s = nullptr;
    h->rd_dbl = static_cast<double*>(alloc(sizeof(double) * 8 * B));
    h->rd_int = static_cast<int*>(alloc(sizeof(int) * (7 + 8) * B));
    h->rd_mcol = static_cast<int*>(alloc(sizeof(int) * n));
    h->rd_ev = static_cast<double*>(alloc(sizeof(double) * SZB_MAX_EVENTS * B));
    h->rd_tout = nullptr;
    if (!h->rd_K || !h->rd_y1 || !h->rd_sum2 || !h->rd_dfdt || !h->rd_F1 || !h->rd_dbl || !h->rd_int || !h->rd_mcol ||
        !h->rd_ev) {
      szb_set_error(h, "szb_rodas: out of device memory for %d stages", s);
      return SZB_ENOMEM;
    }
    h->rodas_B = B;
    h->rodas_stages = s;
    h->rd_mcol_ready = false;
    h->rodas_started = false;
  }
  if (!h->rd_mcol_ready) {
    if (h->mass_row_col.size() != static_cast<size_t>(n)) { szb_set_error(h, "szb_rodas: mass row map missing; call szb_set_mass_rows"); return SZB_ESTATE; }
    SZB_CUDA(h, cudaMemcpyAsync(h->rd_mcol, h->mass_row_col.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    h->rd_mcol_ready = true;
  }
  if (o->n_out > 0 && !(o->c && o->d && o->e) && !h->rd_y0s) {   // Hermite dense output: start state of the step
    void *p0 = nullptr, *p1 = nullptr;
    if (cudaMalloc(&p0, sizeof(double) * bn) != cudaSuccess || cudaMalloc(&p1, sizeof(double) * bn) != cudaSuccess) {
      szb_set_error(h, "szb_rodas: out of device memory for the Hermite dense output");
      return SZB_ENOMEM;
    }
    h->rodas_owned.push_back(p0);
    h->rodas_owned.push_back(p1);
    h->rd_y0s = static_cast<double*>(p0);
    h->rd_F0s = static_cast<double*>(p1);
  }
  double* tout_dev = nullptr;
  if (o->n_out > 0) {
    if (h->rd_tout_cap < o->n_out) {
      void* p = nullptr;
      SZB_CUDA(h, cudaMalloc(&p, sizeof(double) * o->n_out));
      h->rodas_owned.push_back(p);
      h->rd_tout = static_cast<double*>(p);
      h->rd_tout_cap = o->n_out;
    }
    tout_dev = h->rd_tout;
    SZB_CUDA(h, cudaMemcpyAsync(tout_dev, o->t_out, sizeof(double) * o->n_out, cudaMemcpyHostToDevice, st));
  }

  Tableau T{};
  T.s = s; T.pord = o->pord; T.gamma = o->gamma;
  for (int i = 0; i < s * s; ++i) { T.alpha[i] = o->alpha[i]; T.gt[i] = o->gammatilde[i]; }
  for (int i = 0; i < s; ++i) {
    T.a[i] = o->a[i]; T.g[i] = o->g[i]; T.b[i] = o->b[i]; T.bd[i] = o->bd[i];
    T.c[i] = o->c ? o->c[i] : 0.0; T.d[i] = o->d ? o->d[i] : 0.0; T.e[i] = o->e ? o->e[i] : 0.0;
  }
  Ctl C{};
  C.tend = o->tend; C.rtol = o->rtol; C.atol = o->atol; C.f_savety = o->f_savety;
  C.fac1 = o->fac1; C.fac2 = o->fac2;
  C.hmax = o->hmax > 0 ? o->hmax : fabs(o->tend - o->t0);                // :76-77
  C.hmin = 16.0 * (nextafter(fabs(o->t0), INFINITY) - fabs(o->t0));      // 16*spacing(t0), :80
  C.uround = DBL_EPSILON;                                                // spacing(1.0), :81
  C.fix_h = o->fix_h; C.hinit = o->hinit; C.n_out = o->n_out; C.record_cap = o->record_cap;
  C.max_steps = o->max_steps > 0 ? o->max_steps : 1000000;
  C.stop_after = o->stop_after_accepted > 0 ? o->stop_after_accepted : 0;
  C.hermite = (o->n_out > 0 && !(o->c && o->d && o->e)) ? 1 : 0;
  C.n_events = o->n_events;
  C.ev_cap = o->ev_cap;
  C.event_duration = o->event_duration;
  if (o->n_events > 0) {
    if (o->n_events > SZB_MAX_EVENTS || !o->ev_idx || !o->ev_thr || !o->ev_dir || !o->ev_term || o->ev_cap < 1 ||
        !out->ev_count || !out->ev_t || !out->ev_i) {
      szb_set_error(h, "szb_rodas: %d events need ev_idx/ev_thr/ev_dir/ev_term, ev_cap >= 1 and ev_count/ev_t/ev_i (max %d)",
                    o->n_events, SZB_MAX_EVENTS);
      return SZB_EINVAL;
    }
    if (!o->c || !o->d || !o->e) { szb_set_error(h, "szb_rodas: events need a scheme with dense output (ntrp1)"); return SZB_EINVAL; }
    for (int i = 0; i < o->n_events; ++i) {
      if (o->ev_idx[i] < 0 || o->ev_idx[i] >= n) { szb_set_error(h, "szb_rodas: event %d watches component %d of %d", i, o->ev_idx[i], n); return SZB_EINVAL; }
      C.ev_idx[i] = o->ev_idx[i]; C.ev_thr[i] = o->ev_thr[i]; C.ev_dir[i] = o->ev_dir[i]; C.ev_term[i] = o->ev_term[i];
    }
  }
  double dt0 = o->hinit > 0 ? o->hinit : 1e-6 * (o->tend - o->t0);       // :110-114
  dt0 = fmax(dt0, C.hmin);                                               // :116
  dt0 = fmin(dt0, C.hmax);                                               // :117

  State S{};
  S.t = h->rd_dbl; S.dt = S.t + B; S.told = S.dt + B; S.facmax = S.told + B; S.scale = S.facmax + B;
  S.tstage = S.scale + B; S.ddt = S.tstage + B; S.dtacc = S.ddt + B;
  S.reject = h->rd_int; S.active = S.reject + B; S.jmask = S.active + B; S.inext = S.jmask + B;
  S.nrec = S.inext + B; S.acc_left = S.nrec + B; S.stats = S.acc_left + B; S.acc_now = S.stats + 8 * B;
  S.ev_val = h->rd_ev;

  if (B > 65535) {   // the stage kernels index the scenario through gridDim.y
    szb_set_error(h, "szb_rodas: at most 65535 scenarios per call (got %d); split the batch", B);
    return SZB_EINVAL;
  }
  const dim3 eg((n + 255) / 256 > 64 ? 64 : (n + 255) / 256, B);
  SZB_CUDA(h, cudaMemsetAsync(h->flags, 0, sizeof(int) * B, st));
  if (o->resume) {
    if (!h->rodas_started) { szb_set_error(h, "szb_rodas: resume without a previous call on this handle"); return SZB_ESTATE; }
    rodas_resume_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, C.stop_after, B);
  } else {
    rodas_init_kernel<<<B, 128, 0, st>>>(S, *out, h->Y, o->t0, dt0, o->facmax, o->n_out, o->record_cap, C, n, B);
    rodas_resume_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, C.stop_after, B);
    h->rodas_started = true;
  }
  h->launches[4]++;
  int rc;
  const bool use_dfdt = o->dfdt_mode != 0;
  for (long long attempt = 0;; ++attempt) {
    SZB_CUDA(h, cudaMemsetAsync(h->counter, 0, sizeof(int), st));
    step_begin_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, C, T.gamma, h->counter, B);
    h->launches[4]++;
    SZB_CUDA(h, cudaMemcpyAsync(h->pinned, h->counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    SZB_CUDA(h, cudaStreamSynchronize(st));
    if (h->pinned[0] <= 0) break;
    // J(t, y0) where the previous attempt was not rejected (:150-152)
    if ((rc = szb_launch_J(h, 0.0, S.t, h->Y, h->JV, S.jmask, st)) != SZB_OK) return rc;
    // F(t, y0) and the time derivative term (:154-156)
    if ((rc = szb_launch_F(h, 0.0, S.t, h->Y, h->F, nullptr, S.active, st)) != SZB_OK) return rc;
    if (C.hermite) {                                 // ntrp2: the step start (y0, F0) before the stages overwrite F
      hermite_save_kernel<<<eg, 256, 0, st>>>(h->Y, h->F, h->rd_y0s, h->rd_F0s, S, n, B);
      h->launches[4]++;
    }
    if (use_dfdt) {
      if ((rc = szb_launch_F(h, 0.0, S.tstage, h->Y, h->rd_F1, nullptr, S.active, st)) != SZB_OK) return rc;
      dfdt_kernel<<<eg, 256, 0, st>>>(h->F, h->rd_F1, h->rd_dfdt, S, n, B);
      h->launches[4]++;
    }
    stage_rhs_kernel<<<eg, 256, 0, st>>>(h->RHS, h->F, use_dfdt ? h->rd_dfdt : nullptr, nullptr, h->rd_mcol,
                                         T.g[0], S, n, B);
    h->launches[4]++;
    // lu = splu(M - dt*gamma*J) (:159), K[:,0] = lu.solve(rhs) (:163)
    if ((rc = szb_lu_refactor_masked(h, 0.0, S.scale, S.active, nullptr, nullptr, 1, st)) != SZB_OK) return rc;
    rodas_lu_check_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, h->flags, B);
    h->launches[4]++;
    {
      cudaError_t e = szb::lu_solve_launch(h->S, h->cfg, h->LU, h->RHS, h->rd_K, h->CVW, h->XW, h->ybuf, 1, 0, S.active, B, st);
      if (e != cudaSuccess) { szb_set_error(h, "LU solve launch: %s", cudaGetErrorString(e)); return SZB_ECUDA; }
      h->launches[3]++;
    }
    for (int j = 1; j < s; ++j) {                    // :165-173
      stage_prep_kernel<<<eg, 256, 0, st>>>(h->rd_K, h->Y, h->rd_y1, h->rd_sum2, T, j, S, n, B);
      h->launches[4]++;
      if ((rc = szb_launch_F(h, 0.0, S.tstage, h->rd_y1, h->F, nullptr, S.active, st)) != SZB_OK) return rc;
      stage_rhs_kernel<<<eg, 256, 0, st>>>(h->RHS, h->F, use_dfdt ? h->rd_dfdt : nullptr, h->rd_sum2,
                                           h->rd_mcol, T.g[j], S, n, B);
      h->launches[4]++;
      cudaError_t e = szb::lu_solve_launch(h->S, h->cfg, h->LU, h->RHS, h->RHS, h->CVW, h->XW, h->ybuf, 1, 0, S.active, B, st);
      if (e != cudaSuccess) { szb_set_error(h, "LU solve launch: %s", cudaGetErrorString(e)); return SZB_ECUDA; }
      h->launches[3]++;
      stage_post_kernel<<<eg, 256, 0, st>>>(h->rd_K + j * bn, h->RHS, h->rd_sum2, S, n, B);
      h->launches[4]++;
    }
    step_finish_kernel<<<B, 256, 0, st>>>(h->rd_K, h->Y, T, C, S, *out, tout_dev, n, B);
    h->launches[4]++;
    if (C.hermite) {                                 // F at the new point of the accepted scenarios, then ntrp2
      if ((rc = szb_launch_F(h, 0.0, S.t, h->Y, h->rd_F1, nullptr, S.acc_now, st)) != SZB_OK) return rc;
      hermite_output_kernel<<<B, 256, 0, st>>>(h->rd_y0s, h->rd_F0s, h->Y, h->rd_F1, S, C, *out, tout_dev,
                                               o->dfdt_mode == 0 ? 1 : 0, n, B);
      h->launches[4]++;
    }
    SZB_CUDA(h, cudaGetLastError());
  }
  rodas_finalize_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, out->t_end, out->stats, B);
  h->launches[4]++;
  SZB_CUDA(h, cudaStreamSynchronize(st));
  return SZB_OK;
}

// ---- step-wise driving (callback events, solvers/rodas_event.py): controller state of the last call
extern "C" int szb_rodas_state_get(szb_handle* h, double* t_host, double* told_host, double* dt_host,
                                   int32_t* inext_host, int32_t* stats_host) {
  if (!h) return SZB_EINVAL;
  if (!h->rodas_started || h->rodas_B != h->B) { szb_set_error(h, "szb_rodas_state_get before szb_rodas"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  const int B = h->B;
  const double* d = h->rd_dbl;      // t, dt, told, facmax, scale, tstage, ddt, dtacc
  const int* i = h->rd_int;         // reject, active, jmask, inext, nrec, acc_left, stats[8]
  if (t_host) SZB_CUDA(h, cudaMemcpy(t_host, d, sizeof(double) * B, cudaMemcpyDeviceToHost));
  if (told_host) SZB_CUDA(h, cudaMemcpy(told_host, d + 2 * B, sizeof(double) * B, cudaMemcpyDeviceToHost));
  if (dt_host) SZB_CUDA(h, cudaMemcpy(dt_host, d + 7 * B, sizeof(double) * B, cudaMemcpyDeviceToHost));
  if (inext_host) SZB_CUDA(h, cudaMemcpy(inext_host, i + 3 * B, sizeof(int) * B, cudaMemcpyDeviceToHost));
  if (stats_host) SZB_CUDA(h, cudaMemcpy(stats_host, i + 6 * B, sizeof(int) * 8 * B, cudaMemcpyDeviceToHost));
  return SZB_OK;
}

extern "C" int szb_rodas_state_set(szb_handle* h, int32_t b, double t, int32_t inext, int32_t flag) {
  if (!h) return SZB_EINVAL;
  if (!h->rodas_started || h->rodas_B != h->B) { szb_set_error(h, "szb_rodas_state_set before szb_rodas"); return SZB_ESTATE; }
  if (b < 0 || b >= h->B) { szb_set_error(h, "szb_rodas_state_set: scenario %d outside the batch", b); return SZB_EINVAL; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  const int B = h->B;
  SZB_CUDA(h, cudaMemcpy(h->rd_dbl + b, &t, sizeof(double), cudaMemcpyHostToDevice));
  if (inext >= 0) SZB_CUDA(h, cudaMemcpy(h->rd_int + 3 * B + b, &inext, sizeof(int), cudaMemcpyHostToDevice));
  if (flag >= 0) SZB_CUDA(h, cudaMemcpy(h->rd_int + 6 * B + 8 * b + 6, &flag, sizeof(int), cudaMemcpyHostToDevice));
  return SZB_OK;
}

extern "C" int szb_rodas_stages(szb_handle* h, double* K_out_dev, szb_stream stream) {
  if (!h || !K_out_dev) return SZB_EINVAL;
  if (!h->rodas_started) { szb_set_error(h, "szb_rodas_stages before szb_rodas"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  SZB_CUDA(h, cudaMemcpyAsync(K_out_dev, h->rd_K, sizeof(double) * static_cast<size_t>(h->rodas_stages) * h->B * h->n,
                              cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return SZB_OK;
}

extern "C" int szb_set_mass_rows(szb_handle* h, const int32_t* row_col, int32_t n) {
  if (!h || !row_col || n != h->n) { szb_set_error(h, "szb_set_mass_rows: bad arguments"); return SZB_EINVAL; }
  h->mass_row_col.assign(row_col, row_col + n);
  h->rd_mcol_ready = false;
  return SZB_OK;
}
