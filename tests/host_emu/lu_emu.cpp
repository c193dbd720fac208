// lu_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Builds the per-front arithmetic of the batch-interleaved LU
// (solverz_b200/csrc/szb_lu_front.cuh, the very source the CUDA kernels are made of) for the host with one
// lane per "group" and one "warp", so that the host-side schedule AND the index arithmetic of the device
// code can be walked on a machine without a GPU (tests/test_cpu_host.py).  Nothing in the product loads it.
#define SZB_HOST_EMU
#include "szb_lu_front.cuh"

#include <cstring>
#include <vector>

using namespace szb;

namespace {
Tabs tabs(const int* fh, const int* ct, const int* a_src, const int* a_dst, const int* rel, const int* idx,
          const int* rowp, const int* colp, const int* atp, const int* tcp, const int* tce) {
  Tabs T;
  T.fh = reinterpret_cast<const i4*>(fh); T.ct = reinterpret_cast<const i4*>(ct);
  T.a_src = a_src; T.a_dst = a_dst; T.rel = rel; T.idx = idx; T.rowp = rowp; T.colp = colp;
  T.atp = atp; T.tcp = tcp; T.tce = reinterpret_cast<const i4*>(tce);
  return T;
}
}  // namespace

extern "C" {

// factor one scenario; y (length n) gets L^{-1} P rhs when rhs != null; mon = {amax, umax, bad}
int emu_factor(int n_sched, const int* fh, const int* ct, const int* a_src, const int* a_dst, const int* rel,
               const int* idx, const int* rowp, const int* colp, const int* atp, const int* tcp, const int* tce, int cb_size,
               const double* jv,
               const double* mvals, const double* smask, double scale, const double* rhs, double* lu, double* y,
               double* mon, int buf_lines) {
  const Tabs T = tabs(fh, ct, a_src, a_dst, rel, idx, rowp, colp, atp, tcp, tce);
  std::vector<double> cb(static_cast<size_t>(cb_size) + 1, 0.0 / 0.0);   // NaN: reads of unwritten space show
  Lane L;
  L.lu = lu; L.cb = cb.data(); L.cbsrc = cb.data(); L.y = y; L.jv = jv; L.rhs = rhs; L.mvals = mvals; L.smask = smask;
  L.scale = scale; L.plain = !mvals && !smask && scale == 1.0; L.valid = true; L.with_rhs = rhs != nullptr;
  L.amax = 0; L.umax = 0; L.bad = 0;
  const Coop C{0, 1, false};
  // fronts that fit `buf_lines` go through the buffered (shared-memory) path, as in the kernel
  std::vector<double> buf(static_cast<size_t>(buf_lines) + 1, 0.0 / 0.0);
  for (int q = 0; q < n_sched; ++q) {
    const Front F = load_front(T.fh, q);
    double scratch[kNB * kNB + kTileScratch];
    // tiled fronts: alternately with the shared copy of the diagonal tile (CTA-level kernel) and without
    if (F.atp0 >= 0) factor_front_tiled(L, F, T, C, scratch + kNB * kNB, (q & 1) ? scratch : nullptr);
    else if (front_lines(F.f) <= buf_lines) factor_front_buffered(L, F, T, C, buf.data());
    else factor_front(L, F, T, C);
  }
  mon[0] = L.amax; mon[1] = L.umax; mon[2] = L.bad;
  return 0;
}

// x = A^{-1} rhs from stored factors (do_fwd = 1) or from y = L^{-1} P b in pivot order (do_fwd = 0, y in/out)
int emu_solve(int n, int n_sched, const int* fh, const int* ct, const int* rel, const int* idx, const int* rowp,
              const int* colp, int cv_size, const double* lu, const double* rhs, double* y, double* x, int do_fwd) {
  const Tabs T = tabs(fh, ct, nullptr, nullptr, rel, idx, rowp, colp, nullptr, nullptr, nullptr);
  std::vector<double> cv(static_cast<size_t>(cv_size) + 1, 0.0 / 0.0);
  Lane L;
  L.lu = const_cast<double*>(lu); L.cb = nullptr; L.y = y; L.jv = nullptr; L.rhs = rhs; L.valid = true;
  L.with_rhs = true;
  const Coop C{0, 1, false};
  if (do_fwd)
    for (int q = 0; q < n_sched; ++q) forward_front(L, cv.data(), load_front(T.fh, q), T, C);
  double red[kNB];
  for (int q = n_sched - 1; q >= 0; --q) backward_front(L, load_front(T.fh, q), T, C, red);
  for (int k = 0; k < n; ++k) x[colp[k]] = y[k];
  return 0;
}

}  // extern "C"
