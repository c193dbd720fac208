// szb_core.cu -- handle management, buffer binding and the batched Newton driver behind the
// C-ABI of include/szb.h.  The Newton loop restates Solverz/solvers/nlaesolver/nr.py:45-64 per
// scenario with an active mask; see szb_newton below for the exact correspondence.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <mutex>
#include <vector>

#include "szb.h"
#include "szb_internal.cuh"
#include "szb_lu.cuh"
#include "szb_model.h"

namespace {
thread_local std::string g_create_error;
}

void szb_set_error(szb_handle* h, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
}

namespace {

template <typename T>
int upload(szb_handle* h, const T* host, size_t count, T** dev) {
  *dev = nullptr;
  const size_t alloc = count == 0 ? 1 : count;  // keep pointers non-null; never read past an empty host table
  void* p = nullptr;
  SZB_CUDA(h, cudaMalloc(&p, alloc * sizeof(T)));
  h->owned.push_back(p);
  if (host && count > 0) SZB_CUDA(h, cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
  else SZB_CUDA(h, cudaMemset(p, 0, alloc * sizeof(T)));
  *dev = static_cast<T*>(p);
  return SZB_OK;
}

int dev_alloc(szb_handle* h, size_t bytes, void** out, std::vector<void*>& bag) {
  void* p = nullptr;
  if (bytes == 0) bytes = 8;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    szb_set_error(h, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return SZB_ENOMEM;
  }
  bag.push_back(p);
  h->ws_bytes += static_cast<int64_t>(bytes);
  *out = p;
  return SZB_OK;
}

void free_bag(std::vector<void*>& bag) {
  for (void* p : bag) cudaFree(p);
  bag.clear();
}

// ---- small bookkeeping kernels ---------------------------------------------------------
__global__ void newton_init_kernel(int* nstep, int* active, int* flags, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    nstep[b] = 0;
    active[b] = 1;
    flags[b] = SZB_FLAG_RUNNING;
  }
}

// nr.py:48-56 and 63-64, per scenario.  `while max|df| > tol: if nstep > max_it: break; nstep += 1`
__global__ void newton_check_kernel(const double* __restrict__ fnorm, int* __restrict__ nstep,
                                    int* __restrict__ active, int* __restrict__ flags,
                                    int* __restrict__ n_active, double tol, int max_it, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int mine = 0;
  if (b < B && active[b]) {
    const double fn = fnorm[b];
    if (!(fn > tol)) {  // also taken for NaN, exactly like `nan > tol` in Python
      active[b] = 0;
      if (fn < tol) flags[b] = SZB_FLAG_CONVERGED;
      else if (flags[b] == SZB_FLAG_SINGULAR) {}
      else flags[b] = (fn != fn) ? SZB_FLAG_NAN : SZB_FLAG_NOTCONV;
    } else if (nstep[b] > max_it) {
      active[b] = 0;
      if (flags[b] != SZB_FLAG_SINGULAR) flags[b] = isinf(fn) ? SZB_FLAG_NAN : SZB_FLAG_MAXIT;
    } else {
      nstep[b] += 1;
      mine = 1;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_active, __popc(m));
}

}  // namespace

static int need_model(szb_handle* h) {
  if (!h) return SZB_EINVAL;
  if (!h->launch_F) { szb_set_error(h, "no model library attached to this handle"); return SZB_EMODEL; }
  if (h->B <= 0) { szb_set_error(h, "call szb_set_batch first"); return SZB_ESTATE; }
  return SZB_OK;
}

int szb_launch_F(szb_handle* h, double t, const double* t_b, const double* Y, double* F, double* fnorm,
                 const int* active, cudaStream_t st) {
  szbm_args a{};
  a.B = h->B; a.t = t; a.t_b = t_b; a.Y = Y; a.F = F; a.JV = nullptr; a.fnorm = fnorm;
  a.params = h->param_ptr.data(); a.active = active;
  const int e = h->launch_F(&a, st);
  if (e != 0) { szb_set_error(h, "F kernel launch: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return SZB_ECUDA; }
  h->launches[0]++;
  return SZB_OK;
}

int szb_launch_J(szb_handle* h, double t, const double* t_b, const double* Y, double* JV, const int* active,
                 cudaStream_t st) {
  szbm_args a{};
  a.B = h->B; a.t = t; a.t_b = t_b; a.Y = Y; a.F = nullptr; a.JV = JV; a.fnorm = nullptr;
  a.params = h->param_ptr.data(); a.active = active;
  const int e = h->launch_J(&a, st);
  if (e != 0) { szb_set_error(h, "J kernel launch: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return SZB_ECUDA; }
  h->launches[1]++;
  return SZB_OK;
}

int szb_eval_F_masked(szb_handle* h, double t, const int* active, cudaStream_t st) {
  szbm_args a{};
  a.B = h->B; a.t = t; a.Y = h->Y; a.F = h->F; a.JV = nullptr; a.fnorm = h->fnorm;
  a.params = h->param_ptr.data(); a.active = active;
  const int e = h->launch_F(&a, st);
  if (e != 0) { szb_set_error(h, "F kernel launch: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return SZB_ECUDA; }
  h->launches[0]++;
  return SZB_OK;
}

int szb_eval_J_masked(szb_handle* h, double t, const int* active, cudaStream_t st) {
  szbm_args a{};
  a.B = h->B; a.t = t; a.Y = h->Y; a.F = nullptr; a.JV = h->JV; a.fnorm = nullptr;
  a.params = h->param_ptr.data(); a.active = active;
  const int e = h->launch_J(&a, st);
  if (e != 0) { szb_set_error(h, "J kernel launch: %s", cudaGetErrorString(static_cast<cudaError_t>(e))); return SZB_ECUDA; }
  h->launches[1]++;
  return SZB_OK;
}

int szb_lu_refactor_masked(szb_handle* h, double h_gamma, const double* scale_b, const int* active,
                           const double* rhs, double* yout, int store_l, cudaStream_t st,
                           const double* slot_mask) {
  if (!h->has_lu) { szb_set_error(h, "this handle was created without an LU schedule"); return SZB_ESTATE; }
  const double* mv = nullptr;
  double scale = 1.0;
  if (h_gamma != 0.0 || scale_b || slot_mask) {
    if (!h->mass_vals) { szb_set_error(h, "szb_lu_refactor: h_gamma != 0 needs desc.mass_vals"); return SZB_EINVAL; }
    mv = h->mass_vals;
    scale = -h_gamma;
  }
  cudaError_t e = szb::lu_factor_launch(h->S, h->cfg, h->JV, h->LU, h->CBW, h->FBW, mv, slot_mask, scale_b, scale, rhs,
                                        yout, store_l, active, h->flags, h->B, st);
  if (e != cudaSuccess) { szb_set_error(h, "LU factor launch: %s", cudaGetErrorString(e)); return SZB_ECUDA; }
  h->launches[2]++;
  h->factored = store_l != 0;
  return SZB_OK;
}

// relres[b] = max_i |rhs_i - (A x)_i| / max(max_i |rhs_i|, tiny) with A = scale * JV[b] (+ mass) given slot by slot
// (row / column index per CSC slot).  One CTA per scenario, the residual vector in shared memory (fp64 atomicAdd:
// the order of the additions is irrelevant for a check).  This is the a-posteriori stability test of the static
// pivot order: SuperLU's partial pivoting (the reference, laesolver.py:11-45) guarantees a small residual, a fixed
// pivot sequence does not, so the solve is CHECKED instead.
__global__ void __launch_bounds__(256)
lin_residual_kernel(const double* __restrict__ JV, const double* __restrict__ mvals, double scale,
                    const int* __restrict__ srow, const int* __restrict__ scol, const double* __restrict__ RHS,
                    const double* __restrict__ X, double* __restrict__ relres, double* rs_global, int n, int nnz,
                    int B) {
  extern __shared__ double rs_shared[];
  // the residual vector of the scenario in flight: shared memory when n doubles fit, else this CTA's slice of a
  // global scratch (L2; the same atomics, slower)
  double* rs = rs_global ? rs_global + static_cast<size_t>(blockIdx.x) * n : rs_shared;
  __shared__ double red[2][8];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const double* __restrict__ jv = JV + static_cast<size_t>(b) * nnz;
    const double* __restrict__ rhs = RHS + static_cast<size_t>(b) * n;
    const double* __restrict__ x = X + static_cast<size_t>(b) * n;
    __syncthreads();
    double bm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double v = rhs[i];
      rs[i] = v;
      const double a = fabs(v);
      bm = (a > bm || a != a) ? a : bm;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < nnz; p += blockDim.x) {
      double a = scale * jv[p];
      if (mvals) a += mvals[p];
      atomicAdd(&rs[srow[p]], -a * x[scol[p]]);
    }
    __syncthreads();
    double rm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double a = fabs(rs[i]);
      rm = (a > rm || a != a) ? a : rm;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double t1 = __shfl_xor_sync(0xffffffffu, rm, o), t2 = __shfl_xor_sync(0xffffffffu, bm, o);
      rm = (t1 > rm || t1 != t1) ? t1 : rm;
      bm = (t2 > bm || t2 != t2) ? t2 : bm;
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = rm; red[1][threadIdx.x >> 5] = bm; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) {
        rm = (red[0][w] > rm || red[0][w] != red[0][w]) ? red[0][w] : rm;
        bm = (red[1][w] > bm || red[1][w] != red[1][w]) ? red[1][w] : bm;
      }
      relres[b] = rm / fmax(bm, 1e-300);
    }
  }
}

// G = M (y - y_prev) - c (F + kappa F_prev) and max|G| per scenario (NaN-propagating); one CTA per scenario
__global__ void __launch_bounds__(256)
implicit_residual_kernel(const double* __restrict__ F, const double* __restrict__ Y, const double* __restrict__ Y0,
                         const double* __restrict__ F0, const int* __restrict__ mcol, double c, double kappa,
                         double* __restrict__ G, double* __restrict__ fnorm, const int* __restrict__ active, int n) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  if (active && !active[b]) return;
  const size_t o = static_cast<size_t>(b) * n;
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int col = mcol[i];
    double f = F[o + i];
    if (F0) f += kappa * F0[o + i];
    const double g = (col >= 0 ? Y[o + col] - Y0[o + col] : 0.0) - c * f;
    G[o + i] = g;
    const double a = fabs(g);
    m = (a > m || a != a) ? a : m;
  }
  for (int s = 16; s > 0; s >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, s);
    m = (t > m || t != t) ? t : m;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = (red[w] > m || red[w] != red[w]) ? red[w] : m;
    fnorm[b] = m;
  }
}

// =========================================================================================
extern "C" {

int szb_abi_version(void) { return SZB_ABI_VERSION; }

const char* szb_last_error(const szb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int szb_destroy(szb_handle* h) {
  if (!h) return SZB_OK;
  cudaSetDevice(h->device);
  free_bag(h->owned);
  free_bag(h->batch_owned);
  free_bag(h->rodas_owned);
  if (h->pinned) cudaFreeHost(h->pinned);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_pool) cudaEventDestroy(e);
  if (h->model_dl) dlclose(h->model_dl);
  delete h;
  return SZB_OK;
}

int szb_create(const szb_desc* d, szb_handle** out) {
  if (!d || !out) { szb_set_error(nullptr, "szb_create: null argument"); return SZB_EINVAL; }
  *out = nullptr;
  if (d->abi_version != SZB_ABI_VERSION) {
    szb_set_error(nullptr, "szb_create: ABI version %d, library is %d", d->abi_version, SZB_ABI_VERSION);
    return SZB_EINVAL;
  }
  const szb_lu_symbolic& L = d->lu;
  if (d->n <= 0 || d->nnz_j < 0 || (L.ns != 0 && (L.n != d->n || L.nnz_a != d->nnz_j || L.ns < 0))) {
    szb_set_error(nullptr, "szb_create: inconsistent sizes (n=%d lu.n=%d nnz_j=%d lu.nnz_a=%d ns=%d)",
                  d->n, L.n, d->nnz_j, L.nnz_a, L.ns);
    return SZB_EINVAL;
  }
  szb_handle* h = new (std::nothrow) szb_handle();
  if (!h) { szb_set_error(nullptr, "out of host memory"); return SZB_ENOMEM; }
  h->device = d->device;
  h->n = d->n;
  h->nnz_j = d->nnz_j;
  int rc = SZB_OK;
  auto fail = [&](int code) { g_create_error = h->err; szb_destroy(h); return code; };
  {
    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) {
      szb_set_error(h, "cudaSetDevice(%d): %s", d->device, cudaGetErrorString(e));
      return fail(SZB_ECUDA);
    }
  }
  szb::LuDev& S = h->S;
  h->has_lu = L.ns > 0;
  if (h->has_lu) {
  S.n = L.n; S.nnz_a = L.nnz_a; S.ns = L.ns;
  S.lu_size = L.lu_size; S.cb_size = L.cb_size; S.cv_size = L.cv_size;
  S.n_warp_lists = L.n_warp_lists; S.n_group_lists = L.n_group_lists; S.n_top = L.n_top;
  S.f_warp_max = L.f_warp_max; S.f_group_max = L.f_group_max; S.f_top_max = L.f_top_max;
  S.s_warp_max = L.s_warp_max; S.s_group_max = L.s_group_max; S.s_top_max = L.s_top_max;
  S.threads = L.threads; S.wpg = L.wpg;
  S.f_max = L.f_warp_max > L.f_group_max ? L.f_warp_max : L.f_group_max;
  if (L.f_top_max > S.f_max) S.f_max = L.f_top_max;
  S.n_warp_sn = L.wl_ptr[L.n_warp_lists];
  S.n_group_sn = L.n_group_lists > 0 ? L.gl_ptr[L.n_group_lists] : 0;
  if (L.n_sched != S.n_warp_sn + S.n_group_sn + L.n_top || L.n_ga <= 0 || !L.fh || !L.ga) {
    szb_set_error(h, "szb_create: LU schedule tables missing or inconsistent (n_sched=%d)", L.n_sched);
    return fail(SZB_EINVAL);
  }
  const int ns = L.ns;
#define UP(field, cnt)                                                               \
  do {                                                                               \
    int* p_ = nullptr;                                                               \
    if ((rc = upload<int>(h, L.field, static_cast<size_t>(cnt), &p_)) != SZB_OK) return fail(rc); \
    S.field = p_;                                                                    \
  } while (0)
  UP(rowp, L.n); UP(colp, L.n); UP(sn_piv0, ns); UP(sn_s, ns); UP(sn_f, ns);
  UP(idx_ptr, ns + 1); UP(idx, L.idx_ptr[ns]); UP(rel_ptr, ns + 1); UP(rel, L.rel_ptr[ns]);
  UP(child_ptr, ns + 1); UP(child, L.child_ptr[ns]);
  UP(a_ptr, ns + 1); UP(a_src, L.nnz_a); UP(a_dst, L.nnz_a);
  UP(lp_off, ns); UP(up_off, ns); UP(cb_off, ns); UP(ea_ptr, ns + 1); UP(ea_dst, L.ea_ptr[ns]);
  UP(wl_ptr, L.n_warp_lists + 1); UP(wl, L.wl_ptr[L.n_warp_lists]);
  UP(gl_ptr, L.n_group_lists + 1); UP(gl, L.n_group_lists > 0 ? L.gl_ptr[L.n_group_lists] : 0); UP(top, L.n_top);
#undef UP
  {
    int* p_ = nullptr;
    if ((rc = upload<int>(h, L.fh, static_cast<size_t>(L.n_sched) * 16, &p_)) != SZB_OK) return fail(rc);
    S.fh = reinterpret_cast<const int4*>(p_);
    if ((rc = upload<int>(h, L.ga, static_cast<size_t>(L.n_ga) * 2, &p_)) != SZB_OK) return fail(rc);
    S.ga = reinterpret_cast<const int2*>(p_);
  }
  }
  if (d->mass_vals) {
    double* p = nullptr;
    if ((rc = upload<double>(h, d->mass_vals, d->nnz_j, &p)) != SZB_OK) return fail(rc);
    h->mass_vals = p;
  }
  if (d->model_lib) {
    h->model_dl = dlopen(d->model_lib, RTLD_NOW | RTLD_LOCAL);
    if (!h->model_dl) {
      szb_set_error(h, "dlopen(%s): %s", d->model_lib, dlerror());
      return fail(SZB_EMODEL);
    }
    auto sym = [&](const char* name) { return dlsym(h->model_dl, name); };
    auto abi = reinterpret_cast<int (*)()>(sym("szbm_abi_version"));
    auto fn = reinterpret_cast<int (*)()>(sym("szbm_n"));
    auto fnnz = reinterpret_cast<int (*)()>(sym("szbm_nnz_j"));
    auto fnp = reinterpret_cast<int (*)()>(sym("szbm_n_params"));
    auto fpi = reinterpret_cast<const szbm_param_info* (*)()>(sym("szbm_params"));
    h->launch_F = reinterpret_cast<szb_handle::launch_fn>(sym("szbm_launch_F"));
    h->launch_J = reinterpret_cast<szb_handle::launch_fn>(sym("szbm_launch_J"));
    if (!abi || !fn || !fnnz || !fnp || !fpi || !h->launch_F || !h->launch_J) {
      szb_set_error(h, "model library %s lacks szbm_* symbols", d->model_lib);
      return fail(SZB_EMODEL);
    }
    if (abi() != SZBM_ABI_VERSION || fn() != d->n || fnnz() != d->nnz_j) {
      szb_set_error(h, "model library mismatch: abi %d (want %d), n %d (want %d), nnz_j %d (want %d)",
                    abi(), SZBM_ABI_VERSION, fn(), d->n, fnnz(), d->nnz_j);
      return fail(SZB_EMODEL);
    }
    const int np = fnp();
    const szbm_param_info* pi = fpi();
    h->params.assign(pi, pi + np);
    h->param_ptr.assign(np, nullptr);
    h->param_bound.assign(np, 0);
  }
  {
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&h->pinned), 64);
    if (e != cudaSuccess) { szb_set_error(h, "cudaMallocHost: %s", cudaGetErrorString(e)); return fail(SZB_ECUDA); }
    for (auto& ev : h->ev) {
      e = cudaEventCreate(&ev);
      if (e != cudaSuccess) { szb_set_error(h, "cudaEventCreate: %s", cudaGetErrorString(e)); return fail(SZB_ECUDA); }
    }
  }
  *out = h;
  return SZB_OK;
}

int szb_set_batch(szb_handle* h, int32_t B) {
  if (!h || B <= 0) { szb_set_error(h, "szb_set_batch: bad batch %d", B); return SZB_EINVAL; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  free_bag(h->batch_owned);
  free_bag(h->rodas_owned);
  h->imp_G = nullptr; h->imp_G_B = 0;
  h->rodas_B = 0; h->rodas_stages = 0; h->rd_tout_cap = 0; h->rd_mcol_ready = false;
  h->ws_bytes = 0;
  h->B = 0;
  h->factored = false;
  if (h->has_lu) {
    cudaError_t e = szb::lu_plan(h->S, B, &h->cfg);
    if (e != cudaSuccess) {
      szb_set_error(h, "LU launch plan failed (front %d / %d too large for shared memory?): %s",
                    h->S.f_warp_max, h->S.f_top_max, cudaGetErrorString(e));
      return SZB_ECUDA;
    }
  }
  const size_t n = h->n, nz = h->nnz_j;
  int rc;
  void* p;
#define ALLOC(dst, type, count)                                                         \
  do {                                                                                  \
    if ((rc = dev_alloc(h, sizeof(type) * (count), &p, h->batch_owned)) != SZB_OK) return rc; \
    dst = static_cast<type*>(p);                                                        \
  } while (0)
  ALLOC(h->own_Y, double, B * n);
  ALLOC(h->own_F, double, B * n);
  ALLOC(h->own_JV, double, B * nz);
  ALLOC(h->own_RHS, double, B * n);
  if (h->has_lu) {
    ALLOC(h->LU, double, static_cast<size_t>(B) * h->S.lu_size);
    ALLOC(h->ybuf, double, B * n);
    int gmax = h->cfg.grid_factor > h->cfg.grid_solve ? h->cfg.grid_factor : h->cfg.grid_solve;
    if (h->cfg.grid_bwd > gmax) gmax = h->cfg.grid_bwd;
    ALLOC(h->CBW, double, static_cast<size_t>(gmax) * (h->S.cb_size > 0 ? h->S.cb_size : 1));
    ALLOC(h->FBW, double, h->cfg.top_in_global ? static_cast<size_t>(h->cfg.grid_factor) * h->cfg.big_elems : 1);
    ALLOC(h->CVW, double, static_cast<size_t>(gmax) * (h->S.cv_size > 0 ? h->S.cv_size : 1));
    ALLOC(h->XW, double, (h->cfg.x_in_smem && h->cfg.bwd_x_in_smem) ? 1 : static_cast<size_t>(gmax) * n);
  }
  ALLOC(h->fnorm, double, B);
  ALLOC(h->scale_b, double, B);
  ALLOC(h->active, int, B);
  ALLOC(h->nstep, int, B);
  ALLOC(h->flags, int, B);
  ALLOC(h->counter, int, 16);
  for (size_t i = 0; i < h->params.size(); ++i) {
    const auto& pi = h->params[i];
    const size_t esz = pi.dtype == SZB_DTYPE_F64 ? 8 : 4;
    const size_t cnt = static_cast<size_t>(pi.len) * (pi.per_scenario ? B : 1);
    if ((rc = dev_alloc(h, esz * cnt, &p, h->batch_owned)) != SZB_OK) return rc;
    SZB_CUDA(h, cudaMemset(p, 0, esz * (cnt ? cnt : 1)));
    h->param_own.resize(h->params.size());
    h->param_own[i] = p;
    if (!h->param_bound[i]) h->param_ptr[i] = p;
  }
#undef ALLOC
  SZB_CUDA(h, cudaMemset(h->flags, 0, sizeof(int) * B));
  SZB_CUDA(h, cudaMemset(h->own_JV, 0, sizeof(double) * B * nz));
  h->Y = h->bound_Y ? h->bound_Y : h->own_Y;
  h->F = h->bound_F ? h->bound_F : h->own_F;
  h->JV = h->bound_JV ? h->bound_JV : h->own_JV;
  h->RHS = h->bound_RHS ? h->bound_RHS : h->own_RHS;
  h->B = B;
  return SZB_OK;
}

int szb_bind(szb_handle* h, int32_t slot, void* dev_ptr, int64_t len, int32_t dtype) {
  if (!h) return SZB_EINVAL;
  if (h->B <= 0) { szb_set_error(h, "szb_bind: call szb_set_batch first"); return SZB_ESTATE; }
  const int64_t B = h->B, n = h->n;
  auto want = [&](int64_t need, int32_t dt) {
    if (dev_ptr && (len != need || dtype != dt)) {
      szb_set_error(h, "szb_bind(slot %d): got len %lld dtype %d, expected len %lld dtype %d", slot,
                    static_cast<long long>(len), dtype, static_cast<long long>(need), dt);
      return false;
    }
    return true;
  };
  switch (slot) {
    case SZB_SLOT_Y:
      if (!want(B * n, SZB_DTYPE_F64)) return SZB_EINVAL;
      h->bound_Y = static_cast<double*>(dev_ptr); h->Y = dev_ptr ? h->bound_Y : h->own_Y; return SZB_OK;
    case SZB_SLOT_F:
      if (!want(B * n, SZB_DTYPE_F64)) return SZB_EINVAL;
      h->bound_F = static_cast<double*>(dev_ptr); h->F = dev_ptr ? h->bound_F : h->own_F; return SZB_OK;
    case SZB_SLOT_JV:
      if (!want(B * h->nnz_j, SZB_DTYPE_F64)) return SZB_EINVAL;
      h->bound_JV = static_cast<double*>(dev_ptr); h->JV = dev_ptr ? h->bound_JV : h->own_JV; return SZB_OK;
    case SZB_SLOT_RHS:
      if (dev_ptr && (dtype != SZB_DTYPE_F64 || len < B * n || len % (B * n) != 0)) {
        szb_set_error(h, "szb_bind(RHS): len %lld is not a multiple of B*n", static_cast<long long>(len));
        return SZB_EINVAL;
      }
      h->bound_RHS = static_cast<double*>(dev_ptr); h->RHS = dev_ptr ? h->bound_RHS : h->own_RHS;
      h->rhs_len = dev_ptr ? len : B * n;
      return SZB_OK;
    default: break;
  }
  const int i = slot - SZB_SLOT_PARAM0;
  if (i < 0 || i >= static_cast<int>(h->params.size())) {
    szb_set_error(h, "szb_bind: unknown slot %d", slot);
    return SZB_EINVAL;
  }
  const auto& pi = h->params[i];
  if (!want(pi.len * (pi.per_scenario ? B : 1), pi.dtype)) return SZB_EINVAL;
  h->param_bound[i] = dev_ptr != nullptr;
  h->param_ptr[i] = dev_ptr ? dev_ptr : h->param_own[i];
  return SZB_OK;
}

int szb_eval_F(szb_handle* h, double t, szb_stream stream) {
  int rc = need_model(h);
  if (rc) return rc;
  SZB_CUDA(h, cudaSetDevice(h->device));
  return szb_eval_F_masked(h, t, nullptr, static_cast<cudaStream_t>(stream));
}

int szb_eval_J(szb_handle* h, double t, szb_stream stream) {
  int rc = need_model(h);
  if (rc) return rc;
  SZB_CUDA(h, cudaSetDevice(h->device));
  return szb_eval_J_masked(h, t, nullptr, static_cast<cudaStream_t>(stream));
}

int szb_lu_refactor(szb_handle* h, double h_gamma, szb_stream stream) {
  if (!h) return SZB_EINVAL;
  if (h->B <= 0) { szb_set_error(h, "call szb_set_batch first"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  SZB_CUDA(h, cudaMemsetAsync(h->flags, 0, sizeof(int) * h->B, static_cast<cudaStream_t>(stream)));
  return szb_lu_refactor_masked(h, h_gamma, nullptr, nullptr, nullptr, nullptr, 1,
                                static_cast<cudaStream_t>(stream));
}

int szb_lu_residual(szb_handle* h, double h_gamma, const int32_t* slot_row_dev, const int32_t* slot_col_dev,
                    const double* rhs_dev, const double* x_dev, double* relres_dev, szb_stream stream) {
  if (!h || !slot_row_dev || !slot_col_dev || !rhs_dev || !x_dev || !relres_dev) return SZB_EINVAL;
  if (h->B <= 0) { szb_set_error(h, "call szb_set_batch first"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  const double* mv = nullptr;
  double scale = 1.0;
  if (h_gamma != 0.0) {
    if (!h->mass_vals) { szb_set_error(h, "szb_lu_residual: h_gamma != 0 needs desc.mass_vals"); return SZB_EINVAL; }
    mv = h->mass_vals;
    scale = -h_gamma;
  }
  size_t smem = sizeof(double) * static_cast<size_t>(h->n);
  int smem_max = 0;
  cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  int grid = h->B < 4 * sms ? h->B : 4 * sms;
  double* scratch = nullptr;
  if (smem > static_cast<size_t>(smem_max)) {          // large n: one slice of global scratch per resident CTA
    grid = h->B < sms ? h->B : sms;
    SZB_CUDA(h, cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(double) * static_cast<size_t>(grid) * h->n,
                                static_cast<cudaStream_t>(stream)));
    smem = 0;
  }
  SZB_CUDA(h, cudaFuncSetAttribute(lin_residual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem > 0 ? smem : 8)));
  lin_residual_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      h->JV, mv, scale, slot_row_dev, slot_col_dev, rhs_dev, x_dev, relres_dev, scratch, h->n, h->nnz_j, h->B);
  if (scratch) SZB_CUDA(h, cudaFreeAsync(scratch, static_cast<cudaStream_t>(stream)));
  SZB_CUDA(h, cudaGetLastError());
  h->launches[4]++;
  return SZB_OK;
}

int szb_lu_factors(szb_handle* h, int32_t b, double* out_dev, szb_stream stream) {
  if (!h || !out_dev) return SZB_EINVAL;
  if (!h->factored) { szb_set_error(h, "szb_lu_factors before szb_lu_refactor"); return SZB_ESTATE; }
  if (b < 0 || b >= h->B) { szb_set_error(h, "szb_lu_factors: scenario %d outside the batch", b); return SZB_EINVAL; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  SZB_CUDA(h, cudaMemcpyAsync(out_dev, h->LU + static_cast<size_t>(b) * h->S.lu_size, sizeof(double) * h->S.lu_size,
                              cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return SZB_OK;
}

int szb_lu_status(szb_handle* h, int32_t* flags_dev, szb_stream stream) {
  if (!h || !flags_dev) return SZB_EINVAL;
  if (h->B <= 0) { szb_set_error(h, "call szb_set_batch first"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  SZB_CUDA(h, cudaMemcpyAsync(flags_dev, h->flags, sizeof(int) * h->B, cudaMemcpyDeviceToDevice,
                              static_cast<cudaStream_t>(stream)));
  return SZB_OK;
}

int szb_lu_refactor_masked_rows(szb_handle* h, const double* row_mask_dev, szb_stream stream) {
  if (!h) return SZB_EINVAL;
  if (h->B <= 0) { szb_set_error(h, "call szb_set_batch first"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  SZB_CUDA(h, cudaMemsetAsync(h->flags, 0, sizeof(int) * h->B, static_cast<cudaStream_t>(stream)));
  return szb_lu_refactor_masked(h, -1.0, nullptr, nullptr, nullptr, nullptr, 1,
                                static_cast<cudaStream_t>(stream), row_mask_dev);
}

int szb_lu_solve(szb_handle* h, int32_t nrhs, szb_stream stream) {
  if (!h) return SZB_EINVAL;
  if (!h->factored) { szb_set_error(h, "szb_lu_solve before szb_lu_refactor"); return SZB_ESTATE; }
  const int64_t per = static_cast<int64_t>(h->B) * h->n;
  const int64_t have = h->bound_RHS ? h->rhs_len : per;
  if (nrhs < 1 || nrhs * per > have) {
    szb_set_error(h, "szb_lu_solve: nrhs=%d does not fit the bound RHS buffer", nrhs);
    return SZB_EINVAL;
  }
  SZB_CUDA(h, cudaSetDevice(h->device));
  for (int r = 0; r < nrhs; ++r) {
    double* x = h->RHS + r * per;
    cudaError_t e = szb::lu_solve_launch(h->S, h->cfg, h->LU, x, x, h->CVW, h->XW, h->ybuf, 1, 0, nullptr, h->B,
                                         static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { szb_set_error(h, "LU solve launch: %s", cudaGetErrorString(e)); return SZB_ECUDA; }
    h->launches[3]++;
  }
  return SZB_OK;
}

// nr.py:45-64 for the whole batch.
//   df = F(y)                                   -> szbm_launch_F (+ max|F| per scenario)
//   while max|df| > tol:                        -> newton_check_kernel (active mask)
//       if nstep > max_it: break
//       nstep += 1
//       y = y - solve(J(y), df)                 -> szbm_launch_J, lu_factor, lu_solve(mode 1)
//       df = F(y)
//   succeed = max|df| < tol                     -> flags[b]
int szb_newton(szb_handle* h, double tol, int32_t max_it, int32_t* iters_dev, int32_t* flags_dev,
               szb_stream stream) {
  int rc = need_model(h);
  if (rc) return rc;
  if (!h->has_lu) { szb_set_error(h, "this handle was created without an LU schedule"); return SZB_ESTATE; }
  SZB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = h->B;
  const int tb = 256, gb = (B + tb - 1) / tb;
  const bool prof = h->profile;
  for (int i = 0; i < 5; ++i) { h->prof_ms[i] = 0.0; h->prof_units[i] = 0; }
  // sync-free timing: event pairs are recorded around every launch and resolved after the loop
  size_t ev_used = 0;
  std::vector<int> ev_cls;
  auto tic = [&]() {
    if (!prof) return;
    if (ev_used + 2 > h->ev_pool.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
    }
    cudaEventRecord(h->ev_pool[ev_used], st);
  };
  auto toc = [&](int cls) {
    if (!prof) return;
    cudaEventRecord(h->ev_pool[ev_used + 1], st);
    ev_used += 2;
    ev_cls.push_back(cls);
  };
  tic();
  newton_init_kernel<<<gb, tb, 0, st>>>(h->nstep, h->active, h->flags, B);
  h->launches[4]++;
  toc(4);
  const bool implicit_step = h->imp_mode != 0;
  if (implicit_step) {
    if (!h->mass_vals || h->mass_row_col.size() != static_cast<size_t>(h->n) || !h->imp_y0) {
      szb_set_error(h, "szb_newton (implicit step): needs a mass matrix, szb_set_mass_rows and y_prev");
      return SZB_ESTATE;
    }
    if (!h->imp_mcol) {
      void* p = nullptr;
      SZB_CUDA(h, cudaMalloc(&p, sizeof(int) * h->n));
      h->owned.push_back(p);
      h->imp_mcol = static_cast<int*>(p);
    }
    SZB_CUDA(h, cudaMemcpyAsync(h->imp_mcol, h->mass_row_col.data(), sizeof(int) * h->n, cudaMemcpyHostToDevice, st));
    if (!h->imp_G || h->imp_G_B != B) {
      void* p = nullptr;
      SZB_CUDA(h, cudaMalloc(&p, sizeof(double) * static_cast<size_t>(B) * h->n));
      h->batch_owned.push_back(p);
      h->imp_G = static_cast<double*>(p);
      h->imp_G_B = B;
    }
  }
  auto step_residual = [&]() {   // F -> G of the step equation (and its max-norm) when a one-step method drives Newton
    if (!implicit_step) return;
    implicit_residual_kernel<<<B, 256, 0, st>>>(h->F, h->Y, h->imp_y0, h->imp_mode == 2 ? h->imp_f0 : nullptr, h->imp_mcol,
                                                h->imp_c, 1.0, h->imp_G, h->fnorm, h->active, h->n);
    h->launches[4]++;
  };
  const double* resid = implicit_step ? h->imp_G : h->F;
  tic();
  if ((rc = szb_eval_F_masked(h, h->time, h->active, st)) != SZB_OK) return rc;
  step_residual();
  toc(0);
  h->prof_units[0] += B;
  int sweeps = 0;
  for (;;) {
    tic();
    SZB_CUDA(h, cudaMemsetAsync(h->counter, 0, sizeof(int), st));
    newton_check_kernel<<<gb, tb, 0, st>>>(h->fnorm, h->nstep, h->active, h->flags, h->counter, tol,
                                           max_it, B);
    h->launches[4]++;
    SZB_CUDA(h, cudaMemcpyAsync(h->pinned, h->counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    SZB_CUDA(h, cudaStreamSynchronize(st));
    toc(4);
    const int n_active = h->pinned[0];
    if (n_active <= 0) break;
    for (int i = 0; i < 4; ++i) h->prof_units[i] += n_active;
    tic();
    if ((rc = szb_eval_J_masked(h, h->time, h->active, st)) != SZB_OK) return rc;
    toc(1);
    tic();
    // factorisation with the residual riding along as an extra column: y = L^{-1} P F
    if ((rc = szb_lu_refactor_masked(h, implicit_step ? h->imp_c : 0.0, nullptr, h->active, resid, h->ybuf, 0, st)) != SZB_OK)
      return rc;
    toc(2);
    tic();
    {
      // backward substitution and Newton update  Y -= U^{-1} y
      cudaError_t e = szb::lu_solve_launch(h->S, h->cfg, h->LU, h->ybuf, h->Y, h->CVW, h->XW, nullptr, 0, 1,
                                           h->active, B, st);
      if (e != cudaSuccess) { szb_set_error(h, "LU solve launch: %s", cudaGetErrorString(e)); return SZB_ECUDA; }
      h->launches[3]++;
    }
    toc(3);
    tic();
    if ((rc = szb_eval_F_masked(h, h->time, h->active, st)) != SZB_OK) return rc;
    step_residual();
    toc(0);
    if (++sweeps > max_it + 2) {  // cannot happen: every scenario stops after max_it+1 sweeps
      szb_set_error(h, "szb_newton: internal sweep bound exceeded");
      return SZB_ESTATE;
    }
  }
  h->last_sweeps = sweeps;
  if (prof) {  // the loop ended right after a stream synchronisation: every event has completed
    for (size_t i = 0; i < ev_cls.size(); ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]) == cudaSuccess)
        h->prof_ms[ev_cls[i]] += ms;
    }
  }
  if (iters_dev) SZB_CUDA(h, cudaMemcpyAsync(iters_dev, h->nstep, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
  if (flags_dev) SZB_CUDA(h, cudaMemcpyAsync(flags_dev, h->flags, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
  SZB_CUDA(h, cudaGetLastError());
  return SZB_OK;
}

int szb_newton_host(szb_handle* h, double tol, int32_t max_it, double* y_host,
                    const void* const* params_host, int32_t n_params, int32_t* iters_host,
                    int32_t* flags_host, szb_stream stream) {
  int rc = need_model(h);
  if (rc) return rc;
  if (!y_host) { szb_set_error(h, "szb_newton_host: y_host is null"); return SZB_EINVAL; }
  if (n_params > static_cast<int>(h->params.size())) {
    szb_set_error(h, "szb_newton_host: %d parameters given, model has %zu", n_params, h->params.size());
    return SZB_EINVAL;
  }
  SZB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t B = h->B, n = h->n;
  {
    // Upload turnstile: handles driven concurrently from several host threads share one PCIe link, so
    // their uploads are taken one at a time; the handle that has its inputs starts iterating while the
    // next one uploads (copies of one handle overlap the kernels of the others).
    static std::mutex upload_turn;
    std::lock_guard<std::mutex> turn(upload_turn);
    SZB_CUDA(h, cudaMemcpyAsync(h->Y, y_host, B * n * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int i = 0; i < n_params; ++i) {
      if (!params_host || !params_host[i]) continue;
      const auto& pi = h->params[i];
      const size_t bytes = (pi.dtype == SZB_DTYPE_F64 ? 8 : 4) * static_cast<size_t>(pi.len) * (pi.per_scenario ? B : 1);
      SZB_CUDA(h, cudaMemcpyAsync(h->param_ptr[i], params_host[i], bytes, cudaMemcpyHostToDevice, st));
    }
    SZB_CUDA(h, cudaStreamSynchronize(st));
  }
  if ((rc = szb_newton(h, tol, max_it, nullptr, nullptr, stream)) != SZB_OK) return rc;
  SZB_CUDA(h, cudaMemcpyAsync(y_host, h->Y, B * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (iters_host) SZB_CUDA(h, cudaMemcpyAsync(iters_host, h->nstep, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (flags_host) SZB_CUDA(h, cudaMemcpyAsync(flags_host, h->flags, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  SZB_CUDA(h, cudaStreamSynchronize(st));
  return SZB_OK;
}

int szb_query(const szb_handle* h, int32_t what, int64_t* out) {
  if (!h || !out) return SZB_EINVAL;
  switch (what) {
    case SZB_Q_N: *out = h->n; return SZB_OK;
    case SZB_Q_NNZ_J: *out = h->nnz_j; return SZB_OK;
    case SZB_Q_LU_SIZE: *out = h->S.lu_size; return SZB_OK;
    case SZB_Q_N_PARAMS: *out = static_cast<int64_t>(h->params.size()); return SZB_OK;
    case SZB_Q_BATCH: *out = h->B; return SZB_OK;
    case SZB_Q_LAUNCHES: {
      int64_t s = 0;
      for (int i = 0; i < 5; ++i) s += h->launches[i];
      *out = s;
      return SZB_OK;
    }
    case SZB_Q_WORKSPACE_BYTES: *out = h->ws_bytes; return SZB_OK;
    default: return SZB_EINVAL;
  }
}

int szb_set_implicit(szb_handle* h, int32_t mode, double c, const double* y_prev_dev, const double* f_prev_dev) {
  if (!h) return SZB_EINVAL;
  if (mode < 0 || mode > 2) { szb_set_error(h, "szb_set_implicit: mode %d (0 off, 1 backward Euler, 2 trapezoid)", mode); return SZB_EINVAL; }
  if (mode != 0 && !y_prev_dev) { szb_set_error(h, "szb_set_implicit: y_prev_dev is null"); return SZB_EINVAL; }
  if (mode == 2 && !f_prev_dev) { szb_set_error(h, "szb_set_implicit: the trapezoid needs f_prev_dev"); return SZB_EINVAL; }
  h->imp_mode = mode;
  h->imp_c = c;
  h->imp_y0 = y_prev_dev;
  h->imp_f0 = f_prev_dev;
  return SZB_OK;
}

int szb_set_time(szb_handle* h, double t) {
  if (!h) return SZB_EINVAL;
  h->time = t;
  return SZB_OK;
}

int szb_set_profile(szb_handle* h, int32_t on) {
  if (!h) return SZB_EINVAL;
  h->profile = on != 0;
  return SZB_OK;
}

int szb_get_profile(const szb_handle* h, double* out5, int64_t* launches5, int64_t* units5) {
  if (!h) return SZB_EINVAL;
  for (int i = 0; i < 5; ++i) {
    if (out5) out5[i] = h->prof_ms[i];
    if (launches5) launches5[i] = h->launches[i];
    if (units5) units5[i] = h->prof_units[i];
  }
  return SZB_OK;
}

}  // extern "C"
