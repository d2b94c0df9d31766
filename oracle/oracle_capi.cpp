// ORACLE — TEST INFRASTRUCTURE ONLY (see nbg_oracle.hpp header).
// extern "C" surface over the templated restatement, for ctypes use from tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include "nbg_oracle.hpp"
#include "nbg_oracle_ic.hpp"
#include <thread>
#include <atomic>

using namespace nbgo;

namespace {
template <class T> State<T> load_state(int n, const double* x, const double* v, const double* m, const uint8_t* pair, const double* xerr,
                                       const double* verr, const double* jac_step, const double* jac_err, const double* dqdt, const double* t) {
  State<T> s(n);
  const int M = 7 * n;
  for (int q = 0; q < 3 * n; ++q) { s.x[q] = T(x[q]); s.v[q] = T(v[q]); }
  for (int q = 0; q < n; ++q) s.m[q] = T(m[q]);
  if (pair) for (int q = 0; q < n * n; ++q) s.pair[q] = pair[q];
  if (xerr) for (int q = 0; q < 3 * n; ++q) s.xerror[q] = T(xerr[q]);
  if (verr) for (int q = 0; q < 3 * n; ++q) s.verror[q] = T(verr[q]);
  if (jac_step) for (size_t q = 0; q < (size_t)M * M; ++q) s.jac_step[q] = T(jac_step[q]);
  if (jac_err) for (size_t q = 0; q < (size_t)M * M; ++q) s.jac_error[q] = T(jac_err[q]);
  if (dqdt) for (int q = 0; q < M; ++q) s.dqdt[q] = T(dqdt[q]);
  if (t) s.t = T(*t);
  return s;
}
template <class T> void store_state(const State<T>& s, double* x, double* v, double* xerr, double* verr, double* jac_step, double* jac_err,
                                    double* dqdt, double* t) {
  const int n = s.n, M = s.M;
  for (int q = 0; q < 3 * n; ++q) { if (x) x[q] = (double)s.x[q]; if (v) v[q] = (double)s.v[q]; }
  if (xerr) for (int q = 0; q < 3 * n; ++q) xerr[q] = (double)s.xerror[q];
  if (verr) for (int q = 0; q < 3 * n; ++q) verr[q] = (double)s.verror[q];
  if (jac_step) for (size_t q = 0; q < (size_t)M * M; ++q) jac_step[q] = (double)s.jac_step[q];
  if (jac_err) for (size_t q = 0; q < (size_t)M * M; ++q) jac_err[q] = (double)s.jac_error[q];
  if (dqdt) for (int q = 0; q < M; ++q) dqdt[q] = (double)s.dqdt[q];
  if (t) *t = (double)s.t;
}

enum MapId { MAP_AHL21 = 0, MAP_KEPLER_DRIFTIJ = 1, MAP_PHISALPHA = 2, MAP_PHIC = 3, MAP_KICKFAST = 4 };
template <class T> void apply_map(int map, State<T>& s, T h, long nsteps, int i, int j, int drift_first) {
  switch (map) {
    case MAP_AHL21: for (long q = 0; q < nsteps; ++q) ahl21_nograd(s, h); break;
    case MAP_KEPLER_DRIFTIJ: kepler_driftij_gamma_nograd(s, i, j, h, drift_first != 0); break;
    case MAP_PHISALPHA: phisalpha_nograd(s, h, T(2)); break;
    case MAP_PHIC: phic_nograd(s, h); break;
    case MAP_KICKFAST: kickfast_nograd(s, h); break;
  }
}
}  // namespace

extern "C" {

int nbgo_version() { return 1; }
// test switch: 1 = a pair of two massless bodies applies the identity to jac_step (what libnbgrad_b200 does) instead of the reference's
// stale operator (quirk B-2, ahl21.jl:712-716); affects the calling thread only
int nbgo_set_b2_identity(int on) { b2_identity() = on != 0; return 0; }
double nbgo_gnewt() { return GNEWT; }
// CPU-baseline timing only: route the dense products (the reference's mul! calls) through a Fortran-interface dgemm; NULL = the built-in loops
int nbgo_set_dgemm(void* f) { blas_dgemm() = (dgemm_fn)f; return 0; }
int nbgo_has_dgemm() { return blas_dgemm() != nullptr; }

// ElementsIC(t0, H, elements) -> State(ic): x, v, jac_init.   elements is n x 7 column-major.
int nbgo_init_nbody(int n, const double* elements, double t0, const double* eps, double* x, double* v, double* jac_init) {
  ElementsIC<double> ic = make_elements_ic<double>(t0, n, elements, eps);
  std::vector<double> xx, vv, jj;
  init_nbody(ic, xx, vv, jj);
  std::memcpy(x, xx.data(), sizeof(double) * 3 * n);
  std::memcpy(v, vv.data(), sizeof(double) * 3 * n);
  if (jac_init) std::memcpy(jac_init, jj.data(), sizeof(double) * 49 * n * n);
  return 0;
}
// get_orbital_elements(s, ic) (src/outputs/elements.jl:108-137): x, v (3 x n column-major) of a State built from masses m and the
// hierarchy eps (NULL = fully nested) -> out[body][11] = (m, P, t0 = 0, ecosw, esinw, I, Omega, a, e, omega, tp).
int nbgo_orbital_elements(int n, const double* m, const double* eps, const double* x, const double* v, double* out) {
  std::vector<double> el((size_t)n * 7, 0.0);
  for (int i = 0; i < n; ++i) el[i] = m[i];
  ElementsIC<double> ic = make_elements_ic<double>(0.0, n, el.data(), eps);
  get_orbital_elements(ic, x, v, out);
  return 0;
}
int nbgo_ntt(double tmax, const double* periods, int n) { return ntt_from_periods(tmax, periods, n); }

// mode 0: (intr)(s,time) Integrator.jl:159-197 ; mode 1: (intr)(s,N) Integrator.jl:211-234
int nbgo_integrate(int n, double* x, double* v, const double* m, const uint8_t* pair, double* xerr, double* verr, double* jac_step, double* jac_err,
                   double* dqdt, double* t, double h, int mode, double time, long nsteps, int grad) {
  State<double> s = load_state<double>(n, x, v, m, pair, xerr, verr, jac_step, jac_err, dqdt, t);
  if (mode == 0) integrate_to(s, h, time, grad != 0); else integrate_nsteps(s, h, nsteps, grad != 0);
  store_state(s, x, v, xerr, verr, jac_step, jac_err, dqdt, t);
  return 0;
}

// (intr)(s,tt;grad) Transits.jl:140-180.  ntbv = 1 (TransitTiming) or 3 (TransitParameters).
// Output layouts are Julia's: tt[i + n*k] (ttbv[c + 3*(i + n*k)]), dtdq0[i + n*(k + ntt*(q + 7*p))] (with c fastest when ntbv=3).
// stats[0..3] (nullable) += findtransit Newton iterations, kepler-solver calls, gamma Newton iterations, transits that hit ITMAX.
int nbgo_transit_timing(int n, double* x, double* v, const double* m, const uint8_t* pair, const double* jac_init, double* t, double h, double tmax,
                        int ti, int ntt, int ntbv, int grad, double* tt, long* count, double* dtdq0, double* dtdelements, double* xerr,
                        double* verr, double* jac_step, double* jac_err, double* dqdt, long* stats_out) {
  State<double> s = load_state<double>(n, x, v, m, pair, xerr, verr, jac_step, jac_err, dqdt, t);
  if (jac_init) for (size_t q = 0; q < (size_t)49 * n * n; ++q) s.jac_init[q] = jac_init[q];
  TransitOut<double> to(n, ntt, ti, ntbv);
  long newton = 0;
  Stats before = stats();
  const long itmax_before = itmax_hits();
  integrate_transits(s, to, h, tmax, grad != 0, &newton);
  store_state(s, x, v, xerr, verr, jac_step, jac_err, dqdt, t);
  std::memcpy(tt, to.tt.data(), sizeof(double) * to.tt.size());
  for (int i = 0; i < n; ++i) count[i] = to.count[i];
  if (dtdq0) std::memcpy(dtdq0, to.dtdq0.data(), sizeof(double) * to.dtdq0.size());
  if (dtdelements) std::memcpy(dtdelements, to.dtdelements.data(), sizeof(double) * to.dtdelements.size());
  if (stats_out) {
    stats_out[0] += newton;
    stats_out[1] += stats().kepler_calls - before.kepler_calls;
    stats_out[2] += stats().newton_iters - before.newton_iters;
    stats_out[3] += itmax_hits() - itmax_before;   // transits whose findtransit! Newton loop ran into its 20-iteration cap
  }
  return 0;
}

// Batch form used for the CPU baseline: one system per thread, nthreads host threads.
// All arrays are the single-system layouts above with the system index slowest.  Final state is not returned
// except x, v (for checks).  Returns total findtransit Newton iterations in *newton_total.
int nbgo_batch_transit_timing(long nsys, int n, double* x, double* v, const double* m, const double* jac_init, double t0, double h, double tmax,
                              int ti, int ntt, int grad, double* tt, long* count, double* dtdq0, double* dtdelements, int nthreads,
                              long* newton_total, long* itmax_per_system /* nullable: transits of each system that hit ITMAX */) {
  std::atomic<long> next(0), newton(0);
  const int M = 7 * n;
  auto work = [&]() {
    for (;;) {
      long b = next.fetch_add(1);
      if (b >= nsys) break;
      double t = t0;
      long st[4] = {0, 0, 0, 0};
      nbgo_transit_timing(n, x + (size_t)b * 3 * n, v + (size_t)b * 3 * n, m + (size_t)b * n, nullptr, jac_init ? jac_init + (size_t)b * M * M : nullptr,
                          &t, h, tmax, ti, ntt, 1, grad, tt + (size_t)b * n * ntt, count + (size_t)b * n,
                          dtdq0 ? dtdq0 + (size_t)b * n * ntt * M : nullptr, dtdelements ? dtdelements + (size_t)b * n * ntt * M : nullptr, nullptr,
                          nullptr, nullptr, nullptr, nullptr, st);
      newton += st[0];
      if (itmax_per_system) itmax_per_system[b] = st[3];
    }
  };
  std::vector<std::thread> th;
  for (int q = 0; q < nthreads; ++q) th.emplace_back(work);
  for (auto& q : th) q.join();
  if (newton_total) *newton_total = newton.load();
  return 0;
}
// Batch plain integration (N steps), one system per thread; state arrays system-slowest. jac arrays nullable when grad=0.
int nbgo_batch_integrate(long nsys, int n, double* x, double* v, const double* m, double* xerr, double* verr, double* jac_step, double* jac_err,
                         double* dqdt, double h, long nsteps, int grad, int nthreads) {
  std::atomic<long> next(0);
  const int M = 7 * n;
  auto work = [&]() {
    for (;;) {
      long b = next.fetch_add(1);
      if (b >= nsys) break;
      double t = 0;
      nbgo_integrate(n, x + (size_t)b * 3 * n, v + (size_t)b * 3 * n, m + (size_t)b * n, nullptr, xerr ? xerr + (size_t)b * 3 * n : nullptr,
                     verr ? verr + (size_t)b * 3 * n : nullptr, jac_step ? jac_step + (size_t)b * M * M : nullptr,
                     jac_err ? jac_err + (size_t)b * M * M : nullptr, dqdt ? dqdt + (size_t)b * M : nullptr, &t, h, 1, 0.0, nsteps, grad);
    }
  };
  std::vector<std::thread> th;
  for (int q = 0; q < nthreads; ++q) th.emplace_back(work);
  for (auto& q : th) q.join();
  return 0;
}

// ---- unit pieces (reference test/test_kepler_driftij_gamma.jl, test_phisalpha.jl, test_phic.jl, test_kickfast.jl) ----
int nbgo_kepler_driftij(int n, double* x, double* v, const double* m, int i, int j, double h, int drift_first, int grad, double* jac_ij,
                        double* dqdt_ij) {
  State<double> s = load_state<double>(n, x, v, m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (grad) {
    Derivs<double> d(n);
    kepler_driftij_gamma(s, d, i, j, h, drift_first != 0);
    if (jac_ij) std::memcpy(jac_ij, d.jac_ij.data(), sizeof(double) * 196);
    if (dqdt_ij) std::memcpy(dqdt_ij, d.dqdt_ij.data(), sizeof(double) * 14);
  } else {
    kepler_driftij_gamma_nograd(s, i, j, h, drift_first != 0);
  }
  store_state(s, x, v, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  return 0;
}
// which: 2 phisalpha(alpha=2), 3 phic, 4 kickfast.  jac (M x M, identity removed) and dq (M) as left by the reference routine.
int nbgo_kick_piece(int which, int n, double* x, double* v, const double* m, const uint8_t* pair, double h, int grad, double* jac, double* dq) {
  State<double> s = load_state<double>(n, x, v, m, pair, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  const int M = 7 * n;
  if (grad) {
    Derivs<double> d(n);
    if (which == 2) { phisalpha(s, d, h, 2.0); std::memcpy(jac, d.jac_phi.data(), sizeof(double) * M * M); std::memcpy(dq, d.dqdt_phi.data(), sizeof(double) * M); }
    else if (which == 3) { phic(s, d, h); std::memcpy(jac, d.jac_phi.data(), sizeof(double) * M * M); std::memcpy(dq, d.dqdt_phi.data(), sizeof(double) * M); }
    else { kickfast(s, d, h); std::memcpy(jac, d.jac_kick.data(), sizeof(double) * M * M); std::memcpy(dq, d.dqdt_kick.data(), sizeof(double) * M); }
  } else {
    if (which == 2) phisalpha_nograd(s, h, 2.0); else if (which == 3) phic_nograd(s, h); else kickfast_nograd(s, h);
  }
  store_state(s, x, v, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  return 0;
}

// ---- __float128 finite differences (stand-in for the reference tests' BigFloat FD) ----
// Central differences of the no-grad map `map` w.r.t. every x, v (relative step dlnq, absolute if the entry is 0) and m,
// and w.r.t. h (dq = dlnq*h).  jac_num is M x M column-major with mass rows = identity; dqdt_num is M.
int nbgoq_fd_map(int map, int n, const double* x, const double* v, const double* m, const uint8_t* pair, double h, long nsteps, int i, int j,
                 int drift_first, double dlnq_d, double* jac_num, double* dqdt_num) {
  const int M = 7 * n;
  const quad dlnq = (quad)dlnq_d;
  skip_zero_gemm() = true;
  auto base = [&]() { return load_state<quad>(n, x, v, m, pair, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); };
  if (jac_num) {
    for (size_t q = 0; q < (size_t)M * M; ++q) jac_num[q] = 0.0;
    for (int b = 0; b < n; ++b) {
      for (int c = 0; c < 7; ++c) {
        State<quad> sm = base(), sp = base();
        quad* qm = c < 3 ? &sm.x[c + 3 * b] : (c < 6 ? &sm.v[c - 3 + 3 * b] : &sm.m[b]);
        quad* qp = c < 3 ? &sp.x[c + 3 * b] : (c < 6 ? &sp.v[c - 3 + 3 * b] : &sp.m[b]);
        quad dq = dlnq * (*qm);
        if (*qm != 0) { *qm -= dq; *qp += dq; } else { dq = dlnq; *qm = -dq; *qp = dq; }
        apply_map(map, sm, (quad)h, nsteps, i, j, drift_first);
        apply_map(map, sp, (quad)h, nsteps, i, j, drift_first);
        for (int bi = 0; bi < n; ++bi)
          for (int k = 0; k < 3; ++k) {
            jac_num[(7 * bi + k) + (size_t)M * (7 * b + c)] = (double)((quad)0.5 * (sp.x[k + 3 * bi] - sm.x[k + 3 * bi]) / dq);
            jac_num[(7 * bi + 3 + k) + (size_t)M * (7 * b + c)] = (double)((quad)0.5 * (sp.v[k + 3 * bi] - sm.v[k + 3 * bi]) / dq);
          }
      }
      jac_num[(7 * b + 6) + (size_t)M * (7 * b + 6)] = 1.0;
    }
  }
  if (dqdt_num) {
    for (int q = 0; q < M; ++q) dqdt_num[q] = 0.0;
    State<quad> sm = base(), sp = base();
    quad dq = (quad)h * dlnq;
    apply_map(map, sm, (quad)h - dq, nsteps, i, j, drift_first);
    apply_map(map, sp, (quad)h + dq, nsteps, i, j, drift_first);
    for (int bi = 0; bi < n; ++bi)
      for (int k = 0; k < 3; ++k) {
        dqdt_num[7 * bi + k] = (double)((quad)0.5 * (sp.x[k + 3 * bi] - sm.x[k + 3 * bi]) / dq);
        dqdt_num[7 * bi + 3 + k] = (double)((quad)0.5 * (sp.v[k + 3 * bi] - sm.v[k + 3 * bi]) / dq);
      }
  }
  skip_zero_gemm() = false;
  return 0;
}

// test/test_transit_timing.jl:26-62 (and test_transit_parameters.jl): d(tt)/d(elements) by central differences with
// absolute step dq0 on each element of each body, ICs and integration in __float128, grad=false.
// dtde_num layout = dtdelements: [c + ntbv*(i + n*(k + ntt*(iq + 7*jq)))], iq: 0..5 = P,t0,ecosw,esinw,I,Omega ; 6 = mass.
int nbgoq_fd_transit_elements(int n, const double* elements, double t0, double h, double tmax, int ti, int ntt, int ntbv, double dq0_d,
                              double* dtde_num, long* count_out) {
  const quad dq0 = (quad)dq0_d;
  skip_zero_gemm() = true;
  std::vector<quad> el((size_t)n * 7);
  for (size_t q = 0; q < el.size(); ++q) el[q] = (quad)elements[q];
  const size_t tot = (size_t)ntbv * n * ntt * 7 * n;
  for (size_t q = 0; q < tot; ++q) dtde_num[q] = 0.0;
  for (int jq = 0; jq < n; ++jq)
    for (int iq = 0; iq < 7; ++iq) {
      int ivary = (iq == 6) ? 0 : iq + 1;
      std::vector<quad> ep(el), em(el);
      ep[jq + (size_t)n * ivary] += dq0;
      em[jq + (size_t)n * ivary] -= dq0;
      ElementsIC<quad> icp = make_elements_ic<quad>((quad)t0, n, ep.data(), nullptr);
      ElementsIC<quad> icm = make_elements_ic<quad>((quad)t0, n, em.data(), nullptr);
      State<quad> sp = make_state(icp), sm = make_state(icm);
      TransitOut<quad> tp(n, ntt, ti, ntbv), tm(n, ntt, ti, ntbv);
      integrate_transits(sp, tp, (quad)h, (quad)tmax, false);
      integrate_transits(sm, tm, (quad)h, (quad)tmax, false);
      for (int i : tm.occs)
        for (long k = 0; k < tm.count[i] && k < ntt; ++k)
          for (int c = 0; c < ntbv; ++c) {
            size_t src = c + (size_t)ntbv * (i + (size_t)n * k);
            dtde_num[c + (size_t)ntbv * (i + (size_t)n * (k + (size_t)ntt * (iq + 7 * jq)))] = (double)((tp.tt[src] - tm.tt[src]) / (2 * dq0));
          }
      if (count_out) for (int i = 0; i < n; ++i) count_out[i] = tm.count[i];
    }
  skip_zero_gemm() = false;
  return 0;
}

// quad-precision single run of the transit driver from double elements (used to gauge double round-off in tt)
int nbgoq_transit_times(int n, const double* elements, double t0, double h, double tmax, int ti, int ntt, double* tt, long* count) {
  skip_zero_gemm() = true;
  std::vector<quad> el((size_t)n * 7);
  for (size_t q = 0; q < el.size(); ++q) el[q] = (quad)elements[q];
  ElementsIC<quad> ic = make_elements_ic<quad>((quad)t0, n, el.data(), nullptr);
  State<quad> s = make_state(ic);
  TransitOut<quad> to(n, ntt, ti, 1);
  integrate_transits(s, to, (quad)h, (quad)tmax, false);
  for (size_t q = 0; q < to.tt.size(); ++q) tt[q] = (double)to.tt[q];
  for (int i = 0; i < n; ++i) count[i] = to.count[i];
  skip_zero_gemm() = false;
  return 0;
}

// quad-precision run of the FULL gradient path (ahl21! with Derivatives, findtransit!, dtbvdq!, calc_dtdelements!) from the same
// double inputs the double oracle and the GPU get: "the reference algorithm evaluated without round-off".  Used to generate the
// golden of tests/golden/cfg2_quad_system0.npz (tools/gen_quad_golden.py) against which the full-length test checks that the GPU
// deviates from the exact map no more than the reference's own Float64 path does.  Outputs are rounded to double.
int nbgoq_transit_timing_grad(int n, const double* x, const double* v, const double* m, const double* jac_init, double t0, double h, double tmax,
                              int ti, int ntt, double* tt, long* count, double* dtdq0, double* dtdelements, double* x_out, double* v_out,
                              double* jac_step) {
  skip_zero_gemm() = true;
  State<quad> s = load_state<quad>(n, x, v, m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &t0);
  if (jac_init) for (size_t q = 0; q < (size_t)49 * n * n; ++q) s.jac_init[q] = (quad)jac_init[q];
  TransitOut<quad> to(n, ntt, ti, 1);
  integrate_transits(s, to, (quad)h, (quad)tmax, true);
  for (size_t q = 0; q < to.tt.size(); ++q) tt[q] = (double)to.tt[q];
  for (int i = 0; i < n; ++i) count[i] = to.count[i];
  if (dtdq0) for (size_t q = 0; q < to.dtdq0.size(); ++q) dtdq0[q] = (double)to.dtdq0[q];
  if (dtdelements) for (size_t q = 0; q < to.dtdelements.size(); ++q) dtdelements[q] = (double)to.dtdelements[q];
  store_state(s, x_out, v_out, nullptr, nullptr, jac_step, nullptr, nullptr, nullptr);
  skip_zero_gemm() = false;
  return 0;
}

}  // extern "C"
