// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement (C++17, templated on the scalar type) of the hot path of
// ericagol/NbodyGradient.jl v0.2.1: the AHL21 step with and without
// derivatives, the transit driver, findtransit! and dtbvdq!, plus the IC layer
// needed to generate inputs.  Every function cites the reference file:line it
// follows (paths relative to the reference tree).  Loop orders, Kahan
// sequences, repeat-terminated Newton loops and the documented quirks are kept
// as in the reference; evaluation order of the closed-form algebra follows
// Julia's left-to-right parse.
//
// Parity status: the reference (pure Julia) cannot be executed in this
// environment and ships no stored numeric vectors, so bit-level parity with it
// is UNPINNED.  The oracle is pinned instead by (i) the 15-digit end-to-end
// numbers printed in examples/ttv_example.ipynb, (ii) the known-answer test of
// test/test_findtransit.jl, (iii) a re-run of the reference's own derivative
// test programme with __float128 finite differences in place of BigFloat, and
// (iv) its two exact grad == no-grad equality tests.  See tests/test_oracle_*.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may use anything in this directory.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <type_traits>
#include <quadmath.h>

namespace nbgo {

typedef __float128 quad;

// ---- scalar math shims (Julia Base libm -> glibc / libquadmath) -------------
inline double m_sqrt(double x) { return std::sqrt(x); }
inline double m_cbrt(double x) { return std::cbrt(x); }
inline double m_sin(double x) { return std::sin(x); }
inline double m_cos(double x) { return std::cos(x); }
inline double m_sinh(double x) { return std::sinh(x); }
inline double m_cosh(double x) { return std::cosh(x); }
inline double m_exp(double x) { return std::exp(x); }
inline double m_abs(double x) { return std::fabs(x); }
inline double m_atan2(double y, double x) { return std::atan2(y, x); }
inline double m_acos(double x) { return std::acos(x); }
inline double m_fmod(double x, double y) { return std::fmod(x, y); }
inline double m_ceil(double x) { return std::ceil(x); }
inline quad m_sqrt(quad x) { return sqrtq(x); }
inline quad m_cbrt(quad x) { return cbrtq(x); }
inline quad m_sin(quad x) { return sinq(x); }
inline quad m_cos(quad x) { return cosq(x); }
inline quad m_sinh(quad x) { return sinhq(x); }
inline quad m_cosh(quad x) { return coshq(x); }
inline quad m_exp(quad x) { return expq(x); }
inline quad m_abs(quad x) { return fabsq(x); }
inline quad m_atan2(quad y, quad x) { return atan2q(y, x); }
inline quad m_acos(quad x) { return acosq(x); }
inline quad m_fmod(quad x, quad y) { return fmodq(x, y); }
inline quad m_ceil(quad x) { return ceilq(x); }

// Julia sign(): -1, 0, +1
template <class T> inline T jl_sign(T x) { return x > T(0) ? T(1) : (x < T(0) ? T(-1) : x); }

// src/NbodyGradient.jl:13-17
static const double YEAR = 365.242;
static const double GNEWT = 39.4845 / (YEAR * YEAR);
static const double THIRD = 1.0 / 3.0;
static const double PI = 3.141592653589793;

// ---- utils.jl:16-23  comp_sum ------------------------------------------------
template <class T> inline void comp_sum(T& val, T& err, T addend) {
  err += addend;
  T tmp = val + err;
  err = (val - tmp) + err;
  val = tmp;
}
// ---- utils.jl:36-46  comp_sum_matrix! ---------------------------------------
template <class T> inline void comp_sum_matrix(T* val, T* err, const T* addend, size_t len) {
  for (size_t i = 0; i < len; ++i) {
    err[i] += addend[i];
    T tmp = val[i] + err[i];
    err[i] += val[i] - tmp;
    val[i] = tmp;
  }
}
// ---- utils.jl:500-506 dot_fast ----------------------------------------------
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline T dot3(const T* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }

// ---- utils.jl:103-122 cubic1 -------------------------------------------------
template <class T> inline T cubic1(T a, T b, T c) {
  T a3 = a * T(THIRD);
  T Q = a3 * a3 - b * T(THIRD);
  T R = a3 * a3 * a3 + T(0.5) * (-a3 * b + c);
  T R2 = R * R, Q3 = Q * Q * Q;
  if (R2 < Q3) return -c / b;
  T A = -jl_sign(R) * m_cbrt(m_abs(R) + m_sqrt(R2 - Q3));
  T B = (A == T(0)) ? T(0) : Q / A;
  return A + B - a3;
}

// ---- utils.jl:124-396  G3, H1, H2, H3, H5, H6 (closed form gamma>=0.5, series below)
template <class T> inline T G3_series(T gamma, T beta, T sqb) {  // :137-165
  T x2 = -jl_sign(beta) * (gamma * gamma);
  T term = T(1), g3 = T(1), g31 = 2 * g3, g32 = 2 * g3;
  int n = 0, iter = 0;
  while (true) {
    g32 = g31; g31 = g3; n += 1;
    term *= x2 / T((2 * n + 3) * (2 * n + 2));
    g3 += term; iter += 1;
    if (iter >= 100 || g3 == g32 || g3 == g31) break;
  }
  g3 *= -x2 * gamma / (6 * beta * sqb);
  return g3;
}
template <class T> inline T G3(T gamma, T beta, T sqb) {  // :124-135
  if (gamma < T(0.5)) return G3_series(gamma, beta, sqb);
  if (beta >= T(0)) return (gamma - m_sin(gamma)) / (sqb * beta);
  return (gamma - m_sinh(gamma)) / (sqb * beta);
}
template <class T> inline T H1_series(T gamma, T beta) {  // :180-209
  T x2 = -jl_sign(beta) * (gamma * gamma);
  T term = T(1), h1 = T(1), h11 = 2 * h1, h12 = 2 * h1;
  int n = 0, iter = 0;
  while (true) {
    h12 = h11; h11 = h1; n += 1;
    term *= x2 * T(n + 1);
    term /= T((2 * n + 4) * (2 * n + 3) * n);
    h1 += term; iter += 1;
    if (iter >= 100 || h1 == h12 || h1 == h11) break;
  }
  h1 *= (x2 * x2) / (12 * (beta * beta));
  return h1;
}
template <class T> inline T H1(T gamma, T beta) {  // :167-178
  if (gamma < T(0.5)) return H1_series(gamma, beta);
  if (beta >= T(0)) { T s = m_sin(T(0.5) * gamma); return (4 * (s * s) - gamma * m_sin(gamma)) / (beta * beta); }
  T s = m_sinh(T(0.5) * gamma);
  return (-4 * (s * s) + gamma * m_sinh(gamma)) / (beta * beta);
}
template <class T> inline T H2_series(T gamma, T beta, T sqb) {  // :225-254
  T x2 = -jl_sign(beta) * (gamma * gamma);
  T term = T(1), h2 = T(1), h21 = 2 * h2, h22 = 2 * h2;
  int n = 0, iter = 0;
  while (true) {
    h22 = h21; h21 = h2; n += 1;
    term *= x2;
    term /= T((4 * n + 6) * n);
    h2 += term; iter += 1;
    if (iter >= 100 || h2 == h22 || h2 == h21) break;
  }
  h2 *= -x2 * gamma / (3 * beta * sqb);
  return h2;
}
template <class T> inline T H2(T gamma, T beta, T sqb) {  // :211-223
  if (gamma < T(0.5)) return H2_series(gamma, beta, sqb);
  if (beta >= T(0)) return (m_sin(gamma) - gamma * m_cos(gamma)) / (sqb * beta);
  return (m_sinh(gamma) - gamma * m_cosh(gamma)) / (sqb * beta);
}
template <class T> inline T H3_series(T gamma, T beta, T sqb) {  // :271-303
  T x2 = -jl_sign(beta) * (gamma * gamma);
  T term = T(1) / T(30), h3 = T(1) / T(10), h31 = 2 * h3, h32 = 2 * h3;
  int n = 0, iter = 0;
  T four2n = T(4);
  while (true) {
    h32 = h31; h31 = h3; n += 1;
    term *= x2;
    term /= T((2 * n + 4) * (2 * n + 5));
    four2n *= 4;
    h3 += term * (four2n - 1); iter += 1;
    if (iter >= 100 || h3 == h32 || h3 == h31) break;
  }
  h3 *= -(x2 * x2) * gamma / (beta * sqb);
  return h3;
}
template <class T> inline T H3(T gamma, T beta, T sqb) {  // :256-269
  if (gamma < T(0.5)) return H3_series(gamma, beta, sqb);
  if (beta >= T(0)) return (4 * m_sin(gamma) - m_sin(gamma) * m_cos(gamma) - 3 * gamma) / (beta * sqb);
  return (4 * m_sinh(gamma) - m_sinh(gamma) * m_cosh(gamma) - 3 * gamma) / (beta * sqb);
}
template <class T> inline T H5_series(T gamma, T beta, T sqb) {  // :320-349
  T x2 = -jl_sign(beta) * (gamma * gamma);
  T term = T(1) / T(60), h5 = term, h51 = 2 * h5, h52 = 2 * h5;
  int n = 0, iter = 0;
  while (true) {
    h52 = h51; h51 = h5; n += 1;
    term *= x2 * T(n + 1);
    term /= T((2 * n + 5) * (2 * n + 4) * n);
    h5 += term; iter += 1;
    if (iter >= 100 || h5 == h52 || h5 == h51) break;
  }
  h5 *= -(x2 * x2) * gamma / (beta * sqb);
  return h5;
}
template <class T> inline T H5(T gamma, T beta, T sqb) {  // :305-318
  if (gamma < T(0.5)) return H5_series(gamma, beta, sqb);
  if (beta >= T(0)) return (3 * m_sin(gamma) - 2 * gamma - gamma * m_cos(gamma)) / (beta * sqb);
  return (3 * m_sinh(gamma) - 2 * gamma - gamma * m_cosh(gamma)) / (beta * sqb);
}
template <class T> inline T H6_series(T gamma, T beta) {  // :364-396
  T x2 = -jl_sign(beta) * (gamma * gamma);
  T term = T(1) / T(360), h6 = T(1) / T(40), h61 = 2 * h6, h62 = 2 * h6;
  int n = 0, iter = 0;
  T four2n = T(16);
  while (true) {
    h62 = h61; h61 = h6; n += 1;
    term *= x2;
    term /= T((2 * n + 5) * (2 * n + 6));
    four2n *= 4;
    h6 += term * (four2n - T(3 * n) - 7); iter += 1;
    if (iter >= 100 || h6 == h62 || h6 == h61) break;
  }
  h6 *= -(x2 * x2 * x2) / (beta * beta);
  return h6;
}
template <class T> inline T H6(T gamma, T beta) {  // :351-362
  if (gamma < T(0.5)) return H6_series(gamma, beta);
  if (beta >= T(0)) return (9 - 8 * m_cos(gamma) - m_cos(2 * gamma) - 6 * gamma * m_sin(gamma)) / (2 * (beta * beta));
  return (9 - 8 * m_cosh(gamma) - m_cosh(2 * gamma) + 6 * gamma * m_sinh(gamma)) / (2 * (beta * beta));
}

// ---- State (Integrator.jl:49-103) and Derivatives (PreAllocArrays.jl:65-123) --
// Column-major, 0-based: x[k+3*i]; jac_step[r + M*c]; dadq[k + 3*(i + n*(p + 4*di))].
template <class T> struct State {
  int n = 0, M = 0;
  std::vector<T> x, v, m, jac_step, dqdt, jac_init, xerror, verror, dqdt_error, jac_error, a;
  std::vector<uint8_t> pair;  // pair[i + n*j]
  T t = T(0);
  T rij[3], aij[3], x0[3], v0[3], delxv[6], rtmp[3];
  State() {}
  explicit State(int n_) { init(n_); }
  void init(int n_) {
    n = n_; M = 7 * n;
    x.assign(3 * n, T(0)); v.assign(3 * n, T(0)); m.assign(n, T(0));
    xerror.assign(3 * n, T(0)); verror.assign(3 * n, T(0));
    jac_step.assign((size_t)M * M, T(0));
    for (int i = 0; i < M; ++i) jac_step[i + (size_t)M * i] = T(1);  // Integrator.jl:87
    jac_init = jac_step;
    jac_error.assign((size_t)M * M, T(0));
    dqdt.assign(M, T(0)); dqdt_error.assign(M, T(0));
    a.assign(3 * n, T(0));
    pair.assign((size_t)n * n, 0);
    for (int k = 0; k < 3; ++k) rij[k] = aij[k] = x0[k] = v0[k] = rtmp[k] = T(0);
    for (int k = 0; k < 6; ++k) delxv[k] = T(0);
  }
};

// Integrator.jl:123-134 set_state!
template <class T> inline void set_state(State<T>& dst, const State<T>& src) {
  dst.t = src.t; dst.x = src.x; dst.v = src.v; dst.jac_step = src.jac_step;
  dst.xerror = src.xerror; dst.verror = src.verror; dst.jac_error = src.jac_error;
  dst.dqdt = src.dqdt; dst.dqdt_error = src.dqdt_error;
}

template <class T> struct Derivs {
  int n = 0, M = 0;
  std::vector<T> jac_phi, jac_kick, jac_copy, jac_ij, jac_tmp1, jac_tmp2, jac_err1;
  std::vector<T> dqdt_phi, dqdt_kick, dqdt_ij, dqdt_tmp1, jac_kepler, jac_mass, dadq, dotdadq, tmp7n, tmp14;
  explicit Derivs(int n_) : n(n_), M(7 * n_) {
    jac_phi.resize((size_t)M * M); jac_kick.resize((size_t)M * M); jac_copy.resize((size_t)M * M);
    jac_ij.resize(14 * 14); jac_tmp1.resize(14 * M); jac_tmp2.resize(14 * M); jac_err1.resize(14 * M);
    dqdt_phi.resize(M); dqdt_kick.resize(M); dqdt_ij.resize(14); dqdt_tmp1.resize(14);
    jac_kepler.resize(6 * 8); jac_mass.resize(6); dadq.resize((size_t)3 * n * 4 * n); dotdadq.resize(4 * n);
    tmp7n.resize(M); tmp14.resize(14);
    zero_out();
  }
  static void z(std::vector<T>& a) { std::fill(a.begin(), a.end(), T(0)); }
  void zero_out() {  // PreAllocArrays.jl:116-123
    z(jac_phi); z(jac_kick); z(jac_copy); z(jac_ij); z(jac_tmp1); z(jac_tmp2); z(jac_err1);
    z(dqdt_phi); z(dqdt_kick); z(dqdt_ij); z(dqdt_tmp1); z(jac_kepler); z(jac_mass); z(dadq); z(dotdadq);
    z(tmp7n); z(tmp14);
  }
};

// dense column-major C = A(ma x ka) * B(ka x nb): stand-in for LinearAlgebra.mul! (OpenBLAS).
// Each C[r,c] accumulates its products in ascending k; the loops are ordered (c,k,r) so the
// inner loop is a contiguous axpy the compiler can vectorise (same per-element rounding as a
// k-ascending dot product when FP contraction is off).
// skip_zero_gemm(): set by the __float128 finite-difference drivers only -- a product with an
// all-zero left factor (jac_kick when pair is all false) is returned as zeros without the flops.
inline bool& skip_zero_gemm() { static thread_local bool f = false; return f; }
template <class T> inline bool all_zero(const T* A, size_t len) {
  for (size_t q = 0; q < len; ++q) if (A[q] != T(0)) return false;
  return true;
}
// CPU-BASELINE TIMING ONLY: the reference's mul! calls go to OpenBLAS (Julia's LinearAlgebra), so the timing build of this oracle can be
// handed a Fortran-interface dgemm (nbgo_set_dgemm; bench.py passes the OpenBLAS that ships with scipy, single-threaded: the batch is
// threaded over systems).  Never set in the reference-semantics build that the parity tests use (BLAS kernels round differently).
using dgemm_fn = void (*)(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*, const double*,
                          const int*, const double*, double*, const int*);
inline dgemm_fn& blas_dgemm() { static dgemm_fn f = nullptr; return f; }
template <class T> inline void gemm(T* C, const T* A, const T* B, int ma, int ka, int nb) {
  if constexpr (std::is_same<T, double>::value) {
    if (blas_dgemm() && !skip_zero_gemm()) {
      const double one = 1.0, zero = 0.0;
      blas_dgemm()("N", "N", &ma, &nb, &ka, &one, A, &ma, B, &ka, &zero, C, &ma);
      return;
    }
  }
  for (size_t q = 0; q < (size_t)ma * nb; ++q) C[q] = T(0);
  if (skip_zero_gemm() && all_zero(A, (size_t)ma * ka)) return;
  for (int c = 0; c < nb; ++c) {
    T* Cc = C + (size_t)ma * c;
    for (int k = 0; k < ka; ++k) {
      const T b = B[k + (size_t)ka * c];
      const T* Ak = A + (size_t)ma * k;
      for (int r = 0; r < ma; ++r) Cc[r] += Ak[r] * b;
    }
  }
}
template <class T> inline void gemv(T* y, const T* A, const T* x, int ma, int ka) {
  for (int r = 0; r < ma; ++r) y[r] = T(0);
  if (skip_zero_gemm() && all_zero(A, (size_t)ma * ka)) return;
  for (int k = 0; k < ka; ++k) {
    const T b = x[k];
    const T* Ak = A + (size_t)ma * k;
    for (int r = 0; r < ma; ++r) y[r] += Ak[r] * b;
  }
}

// counters for the instrumented-oracle op statistics (optional)
struct Stats { long kepler_calls = 0, newton_iters = 0, hyperbolic = 0; };
inline Stats& stats() { static thread_local Stats s; return s; }

// ---- ahl21.jl:766-890  jac_delxv_gamma! --------------------------------------
template <class T> struct KepParams {
  T gamma, g0, g1, g2, g3, h1, h2, dfdt, fm1, gmh, dgdtm1, r0, r, r0inv, rinv, k, h, beta, betainv, eta, sqb, zeta;
};
template <class T> inline KepParams<T> jac_delxv_gamma(State<T>& s, T k, T h, bool drift_first) {
  T r0;
  s.rtmp[0] = s.x0[0] - h * s.v0[0];
  s.rtmp[1] = s.x0[1] - h * s.v0[1];
  s.rtmp[2] = s.x0[2] - h * s.v0[2];
  r0 = drift_first ? m_sqrt(dot3(s.rtmp)) : m_sqrt(dot3(s.x0));
  T r0inv = T(1) / r0;
  T beta0 = 2 * k * r0inv - dot3(s.v0, s.v0);
  T beta0inv = T(1) / beta0;
  T signb = jl_sign(beta0);
  T sqb = m_sqrt(signb * beta0);
  T zeta = k - r0 * beta0;
  T gamma_guess = T(0);
  T eta = drift_first ? dot3(s.rtmp, s.v0) : dot3(s.x0, s.v0);
  if (zeta != T(0)) {
    T zinv = 6 / zeta;
    gamma_guess = cubic1(T(0.5) * eta * sqb * zinv, r0 * signb * beta0 * zinv, -h * signb * beta0 * sqb * zinv);
  } else {
    if (eta != T(0)) {
      T reta = r0 / eta;
      T disc = reta * reta + 2 * h / eta;
      gamma_guess = disc > T(0) ? sqb * (-reta + m_sqrt(disc)) : h * r0inv * sqb;
    } else {
      gamma_guess = h * r0inv * sqb;
    }
  }
  T gamma = gamma_guess;
  T gamma1 = 2 * gamma, gamma2 = 3 * gamma;
  int iter = 0;
  const int ITMAX = 20;
  T c2 = -2 * zeta, c3 = 2 * eta * signb * sqb, c4 = -sqb * h * beta0, c5 = 2 * eta * signb * sqb;
  T sx, cx, xx;
  stats().kepler_calls++;
  if (!(beta0 > T(0))) stats().hyperbolic++;
  while (iter < ITMAX) {
    gamma2 = gamma1;
    gamma1 = gamma;
    xx = T(0.5) * gamma;
    if (beta0 > T(0)) { sx = m_sin(xx); cx = m_cos(xx); }
    else { sx = m_sinh(xx); cx = m_exp(-xx) + sx; }
    gamma -= (k * gamma + c2 * sx * cx + c3 * (sx * sx) + c4) / (2 * signb * zeta * (sx * sx) + c5 * sx * cx + r0 * beta0);
    iter += 1;
    stats().newton_iters++;
    if (gamma == gamma2 || gamma == gamma1) break;
  }
  for (int j = 0; j < 6; ++j) s.delxv[j] = T(0);
  xx = T(0.5) * gamma;
  if (beta0 > T(0)) { sx = m_sin(xx); cx = m_cos(xx); }
  else { sx = m_sinh(xx); cx = m_exp(-xx) + sx; }
  T g1bs = 2 * sx * cx / sqb;
  T g2bs = 2 * signb * (sx * sx) * beta0inv;
  T g0bs = T(1) - beta0 * g2bs;
  T g3bs = G3(gamma, beta0, sqb);
  T h1 = T(0), h2 = T(0);
  T r = r0 * g0bs + eta * g1bs + k * g2bs;
  T rinv = T(1) / r;
  T dfdt = -k * g1bs * rinv * r0inv;
  T fm1, gmh, dgdtm1;
  if (drift_first) {
    fm1 = -k * r0inv * g2bs;
    gmh = k * r0inv * (h * g2bs - r0 * g3bs);
  } else {
    h1 = H1(gamma, beta0); h2 = H2(gamma, beta0, sqb);
    fm1 = k * rinv * (g2bs - k * r0inv * h1);
    gmh = k * rinv * (r0 * h2 + eta * h1);
  }
  if (drift_first) dgdtm1 = k * r0inv * rinv * (h * g1bs - r0 * g2bs);
  else dgdtm1 = -k * rinv * g2bs;
  for (int j = 0; j < 3; ++j) s.delxv[j] = fm1 * s.x0[j] + gmh * s.v0[j];
  for (int j = 0; j < 3; ++j) s.delxv[3 + j] = dfdt * s.x0[j] + dgdtm1 * s.v0[j];
  KepParams<T> p{gamma, g0bs, g1bs, g2bs, g3bs, h1, h2, dfdt, fm1, gmh, dgdtm1, r0, r, r0inv, rinv, k, h, beta0, beta0inv, eta, sqb, zeta};
  return p;
}

// ---- ahl21.jl:896-1139 compute_jacobian_gamma! (debug=false rows only) -------
// delxv_jac is 6x8 column-major (J(r,c) = dj[r+6*c]); jac_mass is 6.
template <class T>
inline void compute_jacobian_gamma(const KepParams<T>& P, const T* x0, const T* v0, T* dj, T* jac_mass, bool drift_first) {
  const T gamma = P.gamma, g0 = P.g0, g1 = P.g1, g2 = P.g2, g3 = P.g3, h1 = P.h1, dfdt = P.dfdt, fm1 = P.fm1, gmh = P.gmh,
          dgdtm1 = P.dgdtm1, r0 = P.r0, r = P.r, r0inv = P.r0inv, rinv = P.rinv, k = P.k, h = P.h, beta = P.beta,
          betainv = P.betainv, eta = P.eta, sqb = P.sqb, zeta = P.zeta;
  T h2 = P.h2;
  const T G = T(GNEWT);
#define DJ(r_, c_) dj[(r_) + 6 * (c_)]
  T r0inv2 = r0inv * r0inv;
  T r0inv3 = r0inv2 * r0inv;
  T rinv2 = rinv * rinv;
  T rinv3 = rinv2 * rinv;
  T hsq = h * h;
  T ksq = k * k;
  if (drift_first) {  // :906-997
    T d = (h + eta * g2 + 2 * k * g3) * betainv;
    T c1 = d - r0 * g3;
    T c2 = eta * g0 + g1 * zeta;
    T c3 = d * k + g1 * (r0 * r0);
    T c13 = g1 * h - g2 * r0;
    T c9 = 2 * g2 * h - 3 * g3 * r0;
    T c10 = k * (r0inv2 * r0inv2) * (-g2 * r0 * h + k * c9 * betainv - c3 * c13 * rinv);
    T c24 = r0inv3 * (r0 * (2 * k * r0inv - beta) * betainv - g1 * c3 * rinv / g2);
    T h6 = H6(gamma, beta);
    T dfm1dxx = fm1 * c24;
    T dfm1dxv = -fm1 * (g1 * rinv + h * c24);
    T dfm1dvx = dfm1dxv;
    T dfm1dvv = fm1 * rinv * (-r0 * g2 + k * h6 * betainv / g2 + h * (2 * g1 + h * r * c24));
    T dfm1dh = fm1 * (g1 * rinv * (1 / g2 + 2 * k * r0inv - beta) - eta * c24);
    T dfm1dk = fm1 * (1 / k + g1 * c1 * rinv * r0inv / g2 - 2 * betainv * r0inv);
    T h4 = -H1(gamma, beta) * beta;
    T h5 = H5(gamma, beta, sqb);
    T dfm1dk2 = (r0 * h4 + k * h6);
    T dgmhdxx = c10;
    T dgmhdxv = -g2 * k * c13 * rinv * r0inv - h * c10;
    T dgmhdvx = dgmhdxv;
    T h3 = H3(gamma, beta, sqb);
    T h8 = -2 * h3 + 3 * h5;
    T dgmhdvv = 2 * g2 * h * k * c13 * rinv * r0inv + hsq * c10 +
                k * betainv * rinv * r0inv * ((r0 * r0) * h8 - beta * h * r0 * (g2 * g2) + (h * k + eta * r0) * h6);
    T dgmhdh = g2 * k * r0inv + k * c13 * rinv * r0inv + g2 * k * (2 * k * r0inv - beta) * c13 * rinv * r0inv - eta * c10;
    T dgmhdk = r0inv * (k * c1 * c13 * rinv * r0inv + g2 * h - g3 * r0 - k * c9 * betainv * r0inv);
    T dgmhdk2 = (h6 * g3 * ksq + eta * r0 * (h6 + g2 * h4) + (r0 * r0) * g0 * h5 + k * eta * g2 * h6 + (g1 * h6 + g3 * h4) * k * r0);
    for (int j = 0; j < 3; ++j) {
      DJ(j, j) = fm1;
      DJ(j, 3 + j) = gmh;
      for (int i = 0; i < 3; ++i) {
        DJ(j, i) += (dfm1dxx * x0[i] + dfm1dxv * v0[i]) * x0[j] + (dgmhdxx * x0[i] + dgmhdxv * v0[i]) * v0[j];
        DJ(j, 3 + i) += (dfm1dvx * x0[i] + dfm1dvv * v0[i]) * x0[j] + (dgmhdvx * x0[i] + dgmhdvv * v0[i]) * v0[j];
      }
      DJ(j, 6) = dfm1dk * x0[j] + dgmhdk * v0[j];
      DJ(j, 7) = dfm1dh * x0[j] + dgmhdh * v0[j];
      jac_mass[j] = (G * r0inv) * (G * r0inv) * betainv * rinv * (dfm1dk2 * x0[j] - dgmhdk2 * v0[j]);
    }
    T c12 = g0 * h - g1 * r0;
    T c17 = r0 - r - g2 * k;
    T c21 = (g2 * k - r0) * (beta * c3 - k * g1 * r) * betainv * rinv2 * r0inv3 / g1 + eta * g1 * rinv * r0inv2 - 2 * r0inv2;
    T c22 = rinv * (-g1 - g0 * g2 / g1 + g2 * c2 * rinv);
    T c25 = k * rinv * r0inv2 *
            (-g2 + k * (c13 - g2 * r0) * betainv * r0inv2 - c13 * r0inv - c12 * c3 * rinv * r0inv2 +
             c13 * c2 * c3 * rinv2 * r0inv2 - c13 * (k * (g2 * k + r) - g0 * r0 * zeta) * betainv * rinv * r0inv2);
    T c26 = k * rinv2 * r0inv * (-g2 * c12 - g1 * c13 + g2 * c13 * c2 * rinv);
    T ddfdtdxx = dfdt * c21;
    T ddfdtdxv = dfdt * (c22 - h * c21);
    T ddfdtdvx = ddfdtdxv;
    T c34 = (-beta * (eta * eta) * (g2 * g2) - eta * k * h8 - h6 * ksq - 2 * beta * eta * r0 * g1 * g2 +
             ((g2 * g2) - 3 * g1 * g3) * beta * k * r0 - beta * (g1 * g1) * (r0 * r0)) * betainv * rinv2 +
            (eta * (g2 * g2)) * rinv / g1 + (k * h8) * betainv * rinv / g1;
    T ddfdtdvv = dfdt * (c34 - 2 * h * c22 + hsq * c21);
    T ddfdtdk = dfdt * (1 / k - betainv * r0inv - c17 * betainv * rinv * r0inv - c1 * (g1 * c2 - g0 * r) * rinv2 * r0inv / g1);
    T ddfdtdk2 = -(g2 * k - r0) * (beta * r0 * (g3 - g1 * g2) - beta * eta * (g2 * g2) + k * h3) * betainv * rinv2 * r0inv;
    T ddfdtdh = dfdt * (g0 * rinv / g1 - c2 * rinv2 - (2 * k * r0inv - beta) * c22 - eta * c21);
    T dgdtxx = c25;
    T dgdtxv = c26 - h * c25;
    T dgdtvx = c26 - h * c25;
    h2 = H2(gamma, beta, sqb);
    T c33 = d * k * rinv3 * r0inv * k * (h * g2 - r0 * g3) +
            k * (-eta * k * g1 * (g2 * g2) - g1 * g2 * g3 * ksq - r0 * eta * beta * g1 * (g2 * g2) - r0 * k * g1 * h2 -
                 beta * (g2 * g2) * g0 * (r0 * r0)) * betainv * rinv2 * r0inv;
    T dgdtvv = c33 - 2 * h * c26 + hsq * c25;
    T dgdtk = rinv * r0inv * (-k * (c13 - g2 * r0) * betainv * r0inv + c13 - k * c13 * c17 * betainv * rinv * r0inv +
                              k * c1 * c12 * rinv * r0inv - k * c1 * c2 * c13 * rinv2 * r0inv);
    T g22 = g2 * g2;
    T dgdtk2 = k * betainv * rinv2 * r0inv *
               (-beta * (eta * eta) * (g22 * g22) + eta * g2 * (g1 * g22 + (g1 * g1) * g3 - 5 * g2 * g3) * k + g2 * g3 * h3 * ksq +
                2 * eta * r0 * beta * g22 * (g3 - g1 * g2) + (4 * g3 - g0 * g3 - g1 * g2) * (g3 - g1 * g2) * r0 * k +
                beta * (2 * g1 * g3 * g2 - (g1 * g1) * g22 - (g3 * g3)) * (r0 * r0));
    T dgdth = g1 * k * rinv * r0inv + k * c12 * rinv2 * r0inv - k * c2 * c13 * rinv3 * r0inv - (2 * k * r0inv - beta) * c26 - eta * c25;
    for (int j = 0; j < 3; ++j) {
      DJ(3 + j, j) = dfdt;
      DJ(3 + j, 3 + j) = dgdtm1;
      for (int i = 0; i < 3; ++i) {
        DJ(3 + j, i) += (ddfdtdxx * x0[i] + ddfdtdxv * v0[i]) * x0[j] + (dgdtxx * x0[i] + dgdtxv * v0[i]) * v0[j];
        DJ(3 + j, 3 + i) += (ddfdtdvx * x0[i] + ddfdtdvv * v0[i]) * x0[j] + (dgdtvx * x0[i] + dgdtvv * v0[i]) * v0[j];
      }
      DJ(3 + j, 6) = ddfdtdk * x0[j] + dgdtk * v0[j];
      DJ(3 + j, 7) = ddfdtdh * x0[j] + dgdth * v0[j];
      jac_mass[3 + j] = G * G * r0inv * rinv * (ddfdtdk2 * x0[j] + dgdtk2 * v0[j]);
    }
  } else {  // :1021-1117
    T d = (h + eta * g2 + 2 * k * g3) * betainv;
    T c1 = d - r0 * g3;
    T c2 = eta * g0 + g1 * zeta;
    T c3 = d * k + g1 * (r0 * r0);
    T c14 = r0 * g2 - k * h1;
    T c15 = eta * h1 + h2 * r0;
    T c16 = eta * h2 + g1 * gamma * r0 / sqb;
    T c17 = r0 - r - g2 * k;
    T c19 = 4 * eta * h1 + 3 * h2 * r0;
    T c23 = h2 * k - r0 * g1;
    T h6 = H6(gamma, beta);
    T h3 = H3(gamma, beta, sqb);
    T h5 = H5(gamma, beta, sqb);
    T h8 = -2 * h3 + 3 * h5;
    T g22 = g2 * g2;
    T r0inv4 = (r0inv * r0inv) * (r0inv * r0inv);  // r0inv^4
    T dfm1dxx = k * rinv3 * betainv * r0inv4 *
                (k * h1 * (r * r) * r0 * (beta - 2 * k * r0inv) + beta * c3 * (r * c23 + c14 * c2) + c14 * r * (k * (r - g2 * k) + g0 * r0 * zeta));
    T dfm1dxv = k * rinv2 * r0inv * (k * (g2 * h2 + g1 * h1) - 2 * g1 * g2 * r0 + g2 * c14 * c2 * rinv);
    T dfm1dvx = dfm1dxv;
    T dfm1dvv = k * r0inv * rinv2 * betainv *
                (2 * eta * k * (g2 * g3 - g1 * h1) + (3 * g3 * h2 - 4 * h1 * g2) * ksq + beta * g2 * r0 * (3 * h1 * k - g2 * r0) +
                 c14 * rinv * (-beta * g22 * (eta * eta) + eta * k * (2 * g0 * g3 - h2) - h6 * ksq + (-2 * eta * g1 * g2 + k * (h1 - 2 * g1 * g3)) * beta * r0 -
                               beta * (g1 * g1) * (r0 * r0)));
    T dfm1dh = (g1 * k - h2 * ksq * r0inv - k * c14 * c2 * rinv * r0inv) * rinv2;
    T dfm1dk = rinv * r0inv *
               (4 * h1 * ksq * betainv * r0inv - k * h1 - 2 * g2 * k * betainv + c14 - k * c14 * c17 * betainv * rinv * r0inv +
                k * (g1 * r0 - k * h2) * c1 * rinv * r0inv - k * c14 * c1 * c2 * rinv2 * r0inv);
    T dfm1dk2 = betainv * r0inv * rinv2 *
                (r * (2 * eta * k * (g1 * h1 - g3 * g2) + (4 * g2 * h1 - 3 * g3 * h2) * ksq - eta * r0 * beta * g1 * h1 +
                      (g3 * h2 - 4 * g2 * h1) * beta * k * r0 + g2 * h1 * (beta * beta) * (r0 * r0)) -
                 c14 * (-(eta * eta) * beta * g22 - k * eta * h8 - ksq * h6 - eta * r0 * beta * (g1 * g2 + g0 * g3) +
                        2 * (h1 - g1 * g3) * beta * k * r0 - (g2 - beta * g1 * g3) * beta * (r0 * r0)));
    T rr0 = rinv * r0inv;
    T dgmhdxx = k * rinv * r0inv *
                (h2 + k * c19 * betainv * r0inv2 - c16 * c3 * rinv * r0inv2 + c2 * c3 * c15 * (rr0 * rr0) -
                 c15 * (k * (g2 * k + r) - g0 * r0 * zeta) * betainv * rinv * r0inv2);
    T dgmhdxv = k * rinv2 * (h1 * r - g2 * c16 - g1 * c15 + g2 * c2 * c15 * rinv);
    T dgmhdvx = dgmhdxv;
    T dgmhdvv = k * betainv * rinv2 *
                (2 * (eta * eta) * (g1 * h1 - g2 * g3) + eta * k * (4 * g2 * h1 - 3 * h2 * g3) + r0 * eta * (4 * g0 * h1 - 2 * g1 * g3) +
                 3 * r0 * k * ((g1 + beta * g3) * h1 - g3 * g2) + (g0 * h8 - beta * g1 * (g22 + g1 * g3)) * (r0 * r0) -
                 c15 * rinv * (beta * g22 * (eta * eta) + eta * k * h8 + h6 * ksq + (2 * eta * g1 * g2 - k * (g22 - 3 * g1 * g3)) * beta * r0 +
                               beta * (g1 * g1) * (r0 * r0)));
    T dgmhdk = rinv * (k * c1 * c16 * rinv * r0inv + c15 - k * c15 * c17 * betainv * rinv * r0inv - k * c19 * betainv * r0inv -
                       k * c1 * c2 * c15 * rinv2 * r0inv);
    T h7 = beta * g1 * g22 - g0 * h8;
    T dgmhdk2 = betainv * rinv2 *
                (r * (2 * (eta * eta) * (g3 * g2 - g1 * h1) + eta * k * (3 * g3 * h2 - 4 * g2 * h1) +
                      r0 * eta * (beta * g3 * (g1 * g2 + g0 * g3) - 2 * g0 * h6) + (-h6 * (g1 + beta * g3) + g2 * (2 * g3 - h2)) * r0 * k +
                      (h7 - (beta * beta) * g1 * (g3 * g3)) * (r0 * r0)) -
                 c15 * (-beta * (eta * eta) * g22 + eta * k * (-h2 + 2 * g0 * g3) - h6 * ksq - r0 * eta * beta * (h2 + 2 * g0 * g3) +
                        2 * beta * (2 * h1 - g22) * r0 * k + beta * (beta * g1 * g3 - g2) * (r0 * r0)));
    T dgmhdh = k * rinv3 * (r * c16 - c2 * c15);
    for (int j = 0; j < 3; ++j) {
      DJ(j, j) = fm1;
      DJ(j, 3 + j) = gmh;
      for (int i = 0; i < 3; ++i) {
        DJ(j, i) += (dfm1dxx * x0[i] + dfm1dxv * v0[i]) * x0[j] + (dgmhdxx * x0[i] + dgmhdxv * v0[i]) * v0[j];
        DJ(j, 3 + i) += (dfm1dvx * x0[i] + dfm1dvv * v0[i]) * x0[j] + (dgmhdvx * x0[i] + dgmhdvv * v0[i]) * v0[j];
      }
      DJ(j, 6) = dfm1dk * x0[j] + dgmhdk * v0[j];
      DJ(j, 7) = dfm1dh * x0[j] + dgmhdh * v0[j];
      jac_mass[j] = G * G * rinv * r0inv * (dfm1dk2 * x0[j] + dgmhdk2 * v0[j]);
    }
    T ddfdtdxx = dfdt * (eta * g1 * rinv - 2 - g0 * c3 * rinv * r0inv / g1 + c2 * c3 * r0inv * rinv2 - k * (k * g2 - r0) * betainv * rinv * r0inv) * r0inv2;
    T ddfdtdxv = -dfdt * (g0 * g2 / g1 + (r0 * g1 + eta * g2) * rinv) * rinv;
    T ddfdtdvx = ddfdtdxv;
    T ddfdtdvv = -k * rinv3 * r0inv * betainv *
                 ((beta * eta * g22 + k * h8) * (r0 * g0 + k * g2) +
                  g1 * (-h6 * ksq + (-2 * eta * g1 * g2 + (h1 - 2 * g1 * g3) * k) * beta * r0 - beta * (g1 * g1) * (r0 * r0)));
    T ddfdtdk = dfdt * (1 / k + c1 * (r0 - g2 * k) * r0inv * rinv2 / g1 - betainv * r0inv * (1 + c17 * rinv));
    T ddfdtdk2 = (r0 - g2 * k) * betainv * r0inv * rinv2 * (-eta * beta * g22 + h3 * k + (g3 - g1 * g2) * beta * r0);
    T ddfdtdh = dfdt * (r0 - g2 * k) * rinv2 / g1;
    T dgdxx = rinv2 * r0inv3 * ((eta * g2 + g1 * r0) * k * c3 * rinv + g2 * k * (k * (g2 * k - r) - g0 * r0 * zeta) * betainv);
    T dgdxv = k * g2 * rinv3 * (r * g1 + r0 * g1 + eta * g2);
    T dgdvx = dgdxv;
    T dgdvv = k * betainv * rinv3 *
              ((eta * eta) * beta * (g22 * g2) - eta * k * g2 * h3 + 3 * r0 * eta * beta * g1 * g22 + r0 * k * (-g0 * h6 + 3 * beta * g1 * g2 * g3) +
               beta * g2 * (g0 * g2 + (g1 * g1)) * (r0 * r0));
    T dgdk = rinv * r0inv * (-r0 * g2 + g2 * k * (r + r0 - g2 * k) * betainv * rinv - k * g1 * c1 * rinv + k * g2 * c1 * c2 * rinv2);
    T dgdk2 = betainv * rinv2 *
              (-beta * (eta * eta) * (g22 * g2) + eta * k * g2 * h3 + eta * r0 * beta * g2 * (g3 - 2 * g1 * g2) + (h6 - beta * (g22 * g2)) * r0 * k +
               beta * g1 * (g3 - g1 * g2) * (r0 * r0));
    T dgdh = k * rinv3 * (g2 * c2 - r * g1);
    for (int j = 0; j < 3; ++j) {
      DJ(3 + j, j) = dfdt;
      DJ(3 + j, 3 + j) = dgdtm1;
      for (int i = 0; i < 3; ++i) {
        DJ(3 + j, i) += (ddfdtdxx * x0[i] + ddfdtdxv * v0[i]) * x0[j] + (dgdxx * x0[i] + dgdxv * v0[i]) * v0[j];
        DJ(3 + j, 3 + i) += (ddfdtdvx * x0[i] + ddfdtdvv * v0[i]) * x0[j] + (dgdvx * x0[i] + dgdvv * v0[i]) * v0[j];
      }
      DJ(3 + j, 6) = ddfdtdk * x0[j] + dgdk * v0[j];
      DJ(3 + j, 7) = ddfdtdh * x0[j] + dgdh * v0[j];
      jac_mass[3 + j] = G * G * rinv * r0inv * (ddfdtdk2 * x0[j] + dgdk2 * v0[j]);
    }
  }
#undef DJ
}

#define SX(k_, i_) s.x[(k_) + 3 * (i_)]
#define SV(k_, i_) s.v[(k_) + 3 * (i_)]
#define SXE(k_, i_) s.xerror[(k_) + 3 * (i_)]
#define SVE(k_, i_) s.verror[(k_) + 3 * (i_)]
#define SA(k_, i_) s.a[(k_) + 3 * (i_)]

inline bool& b2_identity() { static thread_local bool f = false; return f; }
// ---- ahl21.jl:706-760 kepler_driftij_gamma! (grad) ---------------------------
template <class T> inline void kepler_driftij_gamma(State<T>& s, Derivs<T>& d, int i, int j, T h, bool drift_first) {
  for (int k = 0; k < 3; ++k) { s.x0[k] = SX(k, i) - SX(k, j); s.v0[k] = SV(k, i) - SV(k, j); }
  T gm = T(GNEWT) * (s.m[i] + s.m[j]);
  if (gm == T(0)) {  // quirk B-2: returns before clearing jac_ij (the caller re-applies the previous pair's operator)
    if (b2_identity()) { Derivs<T>::z(d.jac_ij); for (auto& q : d.dqdt_ij) q = T(0); }   // experiment switch (tests): apply the identity instead
    return;
  }
  Derivs<T>::z(d.jac_ij);
  for (int k = 0; k < 6; ++k) s.delxv[k] = T(0);
  Derivs<T>::z(d.jac_kepler); Derivs<T>::z(d.jac_mass);
  KepParams<T> P = jac_delxv_gamma(s, gm, h, drift_first);
  compute_jacobian_gamma(P, s.x0, s.v0, d.jac_kepler.data(), d.jac_mass.data(), drift_first);
  T mijinv = T(1) / (s.m[i] + s.m[j]);
  T mi = s.m[i] * mijinv, mj = s.m[j] * mijinv;
  for (int k = 0; k < 3; ++k) {
    comp_sum(SX(k, i), SXE(k, i), mj * s.delxv[k]);
    comp_sum(SX(k, j), SXE(k, j), -mi * s.delxv[k]);
  }
  for (int k = 0; k < 3; ++k) {
    comp_sum(SV(k, i), SVE(k, i), mj * s.delxv[3 + k]);
    comp_sum(SV(k, j), SVE(k, j), -mi * s.delxv[3 + k]);
  }
#define JIJ(r_, c_) d.jac_ij[(r_) + 14 * (c_)]
#define JK(r_, c_) d.jac_kepler[(r_) + 6 * (c_)]
  for (int l = 0; l < 6; ++l)
    for (int k = 0; k < 6; ++k) {
      JIJ(k, l) += mj * JK(k, l);
      JIJ(k, 7 + l) -= mj * JK(k, l);
      JIJ(7 + k, l) -= mi * JK(k, l);
      JIJ(7 + k, 7 + l) += mi * JK(k, l);
    }
  for (int k = 0; k < 6; ++k) {
    JIJ(k, 6) = d.jac_mass[k] * s.m[j];
    JIJ(k, 13) = mi * s.delxv[k] * mijinv + T(GNEWT) * mj * JK(k, 6);
    JIJ(7 + k, 6) = -mj * s.delxv[k] * mijinv - T(GNEWT) * mi * JK(k, 6);
    JIJ(7 + k, 13) = -d.jac_mass[k] * s.m[i];
  }
  for (int k = 0; k < 6; ++k) {
    d.dqdt_ij[k] = mj * JK(k, 7);
    d.dqdt_ij[7 + k] = -mi * JK(k, 7);
  }
#undef JIJ
#undef JK
}

// ---- ahl21_no_grad.jl:192-212 kepler_driftij_gamma! (no grad) ----------------
template <class T> inline void kepler_driftij_gamma_nograd(State<T>& s, int i, int j, T h, bool drift_first) {
  for (int k = 0; k < 3; ++k) { s.x0[k] = SX(k, i) - SX(k, j); s.v0[k] = SV(k, i) - SV(k, j); }
  T gm = T(GNEWT) * (s.m[i] + s.m[j]);
  if (gm == T(0)) return;
  jac_delxv_gamma(s, gm, h, drift_first);
  T mijinv = T(1) / (s.m[i] + s.m[j]);
  T mi = s.m[i] * mijinv, mj = s.m[j] * mijinv;
  for (int k = 0; k < 3; ++k) {
    comp_sum(SX(k, i), SXE(k, i), mj * s.delxv[k]);
    comp_sum(SX(k, j), SXE(k, j), -mi * s.delxv[k]);
    comp_sum(SV(k, i), SVE(k, i), mj * s.delxv[3 + k]);
    comp_sum(SV(k, j), SVE(k, j), -mi * s.delxv[3 + k]);
  }
}

// ---- ahl21.jl:318-331 drift_grad! -------------------------------------------
template <class T> inline void drift_grad(State<T>& s, T h) {
  const int M = s.M;
  for (int i = 0; i < s.n; ++i) {
    int indi = 7 * i;
    for (int j = 0; j < 3; ++j) comp_sum(SX(j, i), SXE(j, i), h * SV(j, i));
    for (int k = 0; k < M; ++k)
      for (int j = 0; j < 3; ++j)
        comp_sum(s.jac_step[indi + j + (size_t)M * k], s.jac_error[indi + j + (size_t)M * k], h * s.jac_step[indi + 3 + j + (size_t)M * k]);
  }
}
// ---- ahl21_no_grad.jl:24-29 drift! ------------------------------------------
template <class T> inline void drift(State<T>& s, T h) {
  for (int i = 0; i < s.n; ++i)
    for (int j = 0; j < 3; ++j) comp_sum(SX(j, i), SXE(j, i), h * SV(j, i));
}

// ---- ahl21.jl:337-386 kickfast! (grad) --------------------------------------
template <class T> inline void kickfast(State<T>& s, Derivs<T>& d, T h) {
  const int n = s.n, M = s.M;
  for (int k = 0; k < 3; ++k) s.rij[k] = T(0);
  Derivs<T>::z(d.jac_kick);
#define JKK(r_, c_) d.jac_kick[(r_) + (size_t)M * (c_)]
  for (int i = 0; i < n - 1; ++i) {
    int indi = 7 * i;
    for (int j = i + 1; j < n; ++j) {
      int indj = 7 * j;
      if (!s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) s.rij[k] = SX(k, i) - SX(k, j);
      T r2inv = T(1) / dot3(s.rij);
      T r3inv = r2inv * m_sqrt(r2inv);
      T fac2 = h * T(GNEWT) * r3inv;
      for (int k = 0; k < 3; ++k) {
        T fac = fac2 * s.rij[k];
        comp_sum(SV(k, i), SVE(k, i), -s.m[j] * fac);
        comp_sum(SV(k, j), SVE(k, j), s.m[i] * fac);
        d.dqdt_kick[indi + 3 + k] -= s.m[j] * fac / h;
        d.dqdt_kick[indj + 3 + k] += s.m[i] * fac / h;
        JKK(indi + 3 + k, indj + 6) -= fac;
        JKK(indj + 3 + k, indi + 6) += fac;
        fac *= 3 * r2inv;
        for (int p = 0; p < 3; ++p) {
          JKK(indi + 3 + k, indi + p) += fac * s.m[j] * s.rij[p];
          JKK(indi + 3 + k, indj + p) -= fac * s.m[j] * s.rij[p];
          JKK(indj + 3 + k, indj + p) += fac * s.m[i] * s.rij[p];
          JKK(indj + 3 + k, indi + p) -= fac * s.m[i] * s.rij[p];
        }
        JKK(indi + 3 + k, indi + k) -= fac2 * s.m[j];
        JKK(indi + 3 + k, indj + k) += fac2 * s.m[j];
        JKK(indj + 3 + k, indj + k) -= fac2 * s.m[i];
        JKK(indj + 3 + k, indi + k) += fac2 * s.m[i];
      }
    }
  }
#undef JKK
}
// ---- ahl21_no_grad.jl:35-56 kickfast! (no grad) -----------------------------
template <class T> inline void kickfast_nograd(State<T>& s, T h) {
  const int n = s.n;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      if (!s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) s.rij[k] = SX(k, i) - SX(k, j);
      T r2inv = T(1) / dot3(s.rij);
      T r3inv = r2inv * m_sqrt(r2inv);
      T fac2 = h * T(GNEWT) * r3inv;
      for (int k = 0; k < 3; ++k) {
        T fac = fac2 * s.rij[k];
        comp_sum(SV(k, i), SVE(k, i), -s.m[j] * fac);
        comp_sum(SV(k, j), SVE(k, j), s.m[i] * fac);
      }
    }
}

#define DADQ(k_, i_, p_, di_) d.dadq[(k_) + 3 * ((i_) + n * ((p_) + 4 * (di_)))]
#define JPH(r_, c_) d.jac_phi[(r_) + (size_t)M * (c_)]

// ---- ahl21.jl:392-552 phic! (grad) ------------------------------------------
template <class T> inline void phic(State<T>& s, Derivs<T>& d, T h) {
  const int n = s.n, M = s.M;
  std::fill(s.a.begin(), s.a.end(), T(0));
  for (int k = 0; k < 3; ++k) { s.rij[k] = T(0); s.aij[k] = T(0); }
  Derivs<T>::z(d.dadq); Derivs<T>::z(d.dotdadq); Derivs<T>::z(d.jac_phi);
  const T G = T(GNEWT);
  T coeff = h * h * h / 36 * G;
  for (int i = 0; i < n - 1; ++i) {
    int indi = 7 * i;
    for (int j = i + 1; j < n; ++j) {
      if (!s.pair[i + n * j]) continue;
      int indj = 7 * j;
      for (int k = 0; k < 3; ++k) s.rij[k] = SX(k, i) - SX(k, j);
      T r2inv = T(1) / dot3(s.rij);
      T r3inv = r2inv * m_sqrt(r2inv);
      for (int k = 0; k < 3; ++k) {
        T fac = G * s.rij[k] * r3inv;
        T facv = fac * 2 * h / 3;
        comp_sum(SV(k, i), SVE(k, i), -s.m[j] * facv);
        comp_sum(SV(k, j), SVE(k, j), s.m[i] * facv);
        d.dqdt_phi[indi + 3 + k] -= 1 / h * s.m[j] * facv;
        d.dqdt_phi[indj + 3 + k] += 1 / h * s.m[i] * facv;
        SA(k, i) -= s.m[j] * fac;
        SA(k, j) += s.m[i] * fac;
        JPH(indi + 3 + k, indj + 6) -= facv;
        JPH(indj + 3 + k, indi + 6) += facv;
        facv *= 3 * r2inv;
        for (int p = 0; p < 3; ++p) {
          JPH(indi + 3 + k, indi + p) += facv * s.m[j] * s.rij[p];
          JPH(indi + 3 + k, indj + p) -= facv * s.m[j] * s.rij[p];
          JPH(indj + 3 + k, indj + p) += facv * s.m[i] * s.rij[p];
          JPH(indj + 3 + k, indi + p) -= facv * s.m[i] * s.rij[p];
        }
        facv = 2 * h / 3 * G * r3inv;
        JPH(indi + 3 + k, indi + k) -= facv * s.m[j];
        JPH(indi + 3 + k, indj + k) += facv * s.m[j];
        JPH(indj + 3 + k, indj + k) -= facv * s.m[i];
        JPH(indj + 3 + k, indi + k) += facv * s.m[i];
        DADQ(k, i, 3, j) -= fac;
        DADQ(k, j, 3, i) += fac;
        fac *= 3 * r2inv;
        for (int p = 0; p < 3; ++p) {
          DADQ(k, i, p, i) += fac * s.m[j] * s.rij[p];
          DADQ(k, i, p, j) -= fac * s.m[j] * s.rij[p];
          DADQ(k, j, p, j) += fac * s.m[i] * s.rij[p];
          DADQ(k, j, p, i) -= fac * s.m[i] * s.rij[p];
        }
        fac = G * r3inv;
        DADQ(k, i, k, i) -= fac * s.m[j];
        DADQ(k, i, k, j) += fac * s.m[j];
        DADQ(k, j, k, j) -= fac * s.m[i];
        DADQ(k, j, k, i) += fac * s.m[i];
      }
    }
  }
  for (int i = 0; i < n - 1; ++i) {
    int indi = 7 * i;
    for (int j = i + 1; j < n; ++j) {
      if (!s.pair[i + n * j]) continue;
      int indj = 7 * j;
      for (int k = 0; k < 3; ++k) { s.aij[k] = SA(k, i) - SA(k, j); s.rij[k] = SX(k, i) - SX(k, j); }
      Derivs<T>::z(d.dotdadq);
      for (int di = 0; di < n; ++di)
        for (int p = 0; p < 4; ++p)
          for (int k = 0; k < 3; ++k) d.dotdadq[p + 4 * di] += s.rij[k] * (DADQ(k, i, p, di) - DADQ(k, j, p, di));
      T r2 = dot3(s.rij);
      T r1 = m_sqrt(r2);
      T ardot = dot3(s.aij, s.rij);
      T fac1 = coeff / (r2 * r2 * r1);
      T fac2 = 3 * ardot;
      for (int k = 0; k < 3; ++k) {
        T fac = fac1 * (s.rij[k] * fac2 - r2 * s.aij[k]);
        comp_sum(SV(k, i), SVE(k, i), s.m[j] * fac);
        comp_sum(SV(k, j), SVE(k, j), -s.m[i] * fac);
        d.dqdt_phi[indi + 3 + k] += 3 / h * s.m[j] * fac;
        d.dqdt_phi[indj + 3 + k] -= 3 / h * s.m[i] * fac;
        JPH(indi + 3 + k, indj + 6) += fac;
        JPH(indj + 3 + k, indi + 6) -= fac;
        fac *= 5 / r2;
        for (int p = 0; p < 3; ++p) {
          JPH(indi + 3 + k, indi + p) -= fac * s.m[j] * s.rij[p];
          JPH(indi + 3 + k, indj + p) += fac * s.m[j] * s.rij[p];
          JPH(indj + 3 + k, indj + p) -= fac * s.m[i] * s.rij[p];
          JPH(indj + 3 + k, indi + p) += fac * s.m[i] * s.rij[p];
        }
        fac = fac1 * fac2;
        JPH(indi + 3 + k, indi + k) += fac * s.m[j];
        JPH(indi + 3 + k, indj + k) -= fac * s.m[j];
        JPH(indj + 3 + k, indj + k) += fac * s.m[i];
        JPH(indj + 3 + k, indi + k) -= fac * s.m[i];
        fac = -2 * fac1 * s.aij[k];
        for (int p = 0; p < 3; ++p) {
          T fac3 = fac * s.rij[p] + fac1 * 3 * s.rij[k] * s.aij[p];
          JPH(indi + 3 + k, indi + p) += s.m[j] * fac3;
          JPH(indi + 3 + k, indj + p) -= s.m[j] * fac3;
          JPH(indj + 3 + k, indj + p) += s.m[i] * fac3;
          JPH(indj + 3 + k, indi + p) -= s.m[i] * fac3;
        }
        fac = -fac1 * r2;
        for (int di = 0; di < n; ++di) {
          int indd = 7 * di;
          for (int p = 0; p < 3; ++p) {
            JPH(indi + 3 + k, indd + p) += fac * s.m[j] * (DADQ(k, i, p, di) - DADQ(k, j, p, di));
            JPH(indj + 3 + k, indd + p) -= fac * s.m[i] * (DADQ(k, i, p, di) - DADQ(k, j, p, di));
          }
          JPH(indi + 3 + k, indd + 6) += fac * s.m[j] * (DADQ(k, i, 3, di) - DADQ(k, j, 3, di));
          JPH(indj + 3 + k, indd + 6) -= fac * s.m[i] * (DADQ(k, i, 3, di) - DADQ(k, j, 3, di));
        }
        fac = 3 * fac1 * s.rij[k];
        for (int di = 0; di < n; ++di) {
          int indd = 7 * di;
          for (int p = 0; p < 3; ++p) {
            JPH(indi + 3 + k, indd + p) += fac * s.m[j] * d.dotdadq[p + 4 * di];
            JPH(indj + 3 + k, indd + p) -= fac * s.m[i] * d.dotdadq[p + 4 * di];
          }
          JPH(indi + 3 + k, indd + 6) += fac * s.m[j] * d.dotdadq[3 + 4 * di];
          JPH(indj + 3 + k, indd + 6) -= fac * s.m[i] * d.dotdadq[3 + 4 * di];
        }
      }
    }
  }
}
// ---- ahl21_no_grad.jl:62-104 phic! (no grad) --------------------------------
template <class T> inline void phic_nograd(State<T>& s, T h) {
  const int n = s.n;
  std::fill(s.a.begin(), s.a.end(), T(0));
  const T G = T(GNEWT);
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      if (!s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) s.rij[k] = SX(k, i) - SX(k, j);
      T r2inv = T(1) / dot3(s.rij);
      T r3inv = r2inv * m_sqrt(r2inv);
      for (int k = 0; k < 3; ++k) {
        T fac = G * s.rij[k] * r3inv;
        T facv = fac * 2 * h / 3;
        comp_sum(SV(k, i), SVE(k, i), -s.m[j] * facv);
        comp_sum(SV(k, j), SVE(k, j), s.m[i] * facv);
        SA(k, i) -= s.m[j] * fac;
        SA(k, j) += s.m[i] * fac;
      }
    }
  T coeff = h * h * h / 36 * G;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      if (!s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) { s.aij[k] = SA(k, i) - SA(k, j); s.rij[k] = SX(k, i) - SX(k, j); }
      T r2 = dot3(s.rij);
      T r1 = m_sqrt(r2);
      T ardot = dot3(s.aij, s.rij);
      T fac1 = coeff / (r2 * r2 * r1);
      T fac2 = 3 * ardot;
      for (int k = 0; k < 3; ++k) {
        T fac = fac1 * (s.rij[k] * fac2 - r2 * s.aij[k]);
        comp_sum(SV(k, i), SVE(k, i), s.m[j] * fac);
        comp_sum(SV(k, j), SVE(k, j), -s.m[i] * fac);
      }
    }
}

// ---- ahl21.jl:558-700 phisalpha! (grad) -------------------------------------
template <class T> inline void phisalpha(State<T>& s, Derivs<T>& d, T h, T alpha) {
  const int n = s.n, M = s.M;
  std::fill(s.a.begin(), s.a.end(), T(0));
  Derivs<T>::z(d.dadq); Derivs<T>::z(d.dotdadq);
  for (int k = 0; k < 3; ++k) { s.rij[k] = T(0); s.aij[k] = T(0); }
  const T G = T(GNEWT);
  T coeff = alpha * (h * h * h) / 96 * 2 * G;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      if (s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) s.rij[k] = SX(k, i) - SX(k, j);
      T r2 = dot3(s.rij);
      T r3 = r2 * m_sqrt(r2);
      T fac2 = G / r3;
      for (int k = 0; k < 3; ++k) {
        T fac = fac2 * s.rij[k];
        SA(k, i) -= s.m[j] * fac;
        SA(k, j) += s.m[i] * fac;
        DADQ(k, i, 3, j) -= fac;
        DADQ(k, j, 3, i) += fac;
        fac *= 3 / r2;
        for (int p = 0; p < 3; ++p) {
          DADQ(k, i, p, i) += fac * s.m[j] * s.rij[p];
          DADQ(k, i, p, j) -= fac * s.m[j] * s.rij[p];
          DADQ(k, j, p, j) += fac * s.m[i] * s.rij[p];
          DADQ(k, j, p, i) -= fac * s.m[i] * s.rij[p];
        }
        DADQ(k, i, k, i) -= fac2 * s.m[j];
        DADQ(k, i, k, j) += fac2 * s.m[j];
        DADQ(k, j, k, j) -= fac2 * s.m[i];
        DADQ(k, j, k, i) += fac2 * s.m[i];
      }
    }
  for (int i = 0; i < n - 1; ++i) {
    int indi = 7 * i;
    for (int j = i + 1; j < n; ++j) {
      if (s.pair[i + n * j]) continue;
      int indj = 7 * j;
      for (int k = 0; k < 3; ++k) { s.aij[k] = SA(k, i) - SA(k, j); s.rij[k] = SX(k, i) - SX(k, j); }
      Derivs<T>::z(d.dotdadq);
      for (int di = 0; di < n; ++di)
        for (int p = 0; p < 4; ++p)
          for (int k = 0; k < 3; ++k) d.dotdadq[p + 4 * di] += s.rij[k] * (DADQ(k, i, p, di) - DADQ(k, j, p, di));
      T r2 = dot3(s.rij);
      T r1 = m_sqrt(r2);
      T ardot = dot3(s.aij, s.rij);
      T fac1 = coeff / (r2 * r2 * r1);
      T fac2 = (2 * G * (s.m[i] + s.m[j]) / r1 + 3 * ardot);
      for (int k = 0; k < 3; ++k) {
        T fac = fac1 * (s.rij[k] * fac2 - r2 * s.aij[k]);
        comp_sum(SV(k, i), SVE(k, i), s.m[j] * fac);
        comp_sum(SV(k, j), SVE(k, j), -s.m[i] * fac);
        d.dqdt_phi[indi + 3 + k] += 3 / h * s.m[j] * fac;
        d.dqdt_phi[indj + 3 + k] -= 3 / h * s.m[i] * fac;
        JPH(indi + 3 + k, indj + 6) += fac;
        JPH(indj + 3 + k, indi + 6) -= fac;
        fac *= 5 / r2;
        for (int p = 0; p < 3; ++p) {
          JPH(indi + 3 + k, indi + p) -= fac * s.m[j] * s.rij[p];
          JPH(indi + 3 + k, indj + p) += fac * s.m[j] * s.rij[p];
          JPH(indj + 3 + k, indj + p) -= fac * s.m[i] * s.rij[p];
          JPH(indj + 3 + k, indi + p) += fac * s.m[i] * s.rij[p];
        }
        fac = 2 * G * fac1 * s.rij[k] / r1;
        JPH(indi + 3 + k, indi + 6) += fac * s.m[j];
        JPH(indi + 3 + k, indj + 6) += fac * s.m[j];
        JPH(indj + 3 + k, indj + 6) -= fac * s.m[i];
        JPH(indj + 3 + k, indi + 6) -= fac * s.m[i];
        fac = fac1 * fac2;
        JPH(indi + 3 + k, indi + k) += fac * s.m[j];
        JPH(indi + 3 + k, indj + k) -= fac * s.m[j];
        JPH(indj + 3 + k, indj + k) += fac * s.m[i];
        JPH(indj + 3 + k, indi + k) -= fac * s.m[i];
        fac = -2 * fac1 * (s.rij[k] * G * (s.m[i] + s.m[j]) / (r2 * r1) + s.aij[k]);
        for (int p = 0; p < 3; ++p) {
          T fac3 = fac * s.rij[p] + fac1 * 3 * s.rij[k] * s.aij[p];
          JPH(indi + 3 + k, indi + p) += s.m[j] * fac3;
          JPH(indi + 3 + k, indj + p) -= s.m[j] * fac3;
          JPH(indj + 3 + k, indj + p) += s.m[i] * fac3;
          JPH(indj + 3 + k, indi + p) -= s.m[i] * fac3;
        }
        fac = -fac1 * r2;
        for (int di = 0; di < n; ++di) {
          int indd = 7 * di;
          for (int p = 0; p < 3; ++p) JPH(indi + 3 + k, indd + p) += fac * s.m[j] * (DADQ(k, i, p, di) - DADQ(k, j, p, di));
          for (int p = 0; p < 3; ++p) JPH(indj + 3 + k, indd + p) -= fac * s.m[i] * (DADQ(k, i, p, di) - DADQ(k, j, p, di));
          JPH(indi + 3 + k, indd + 6) += fac * s.m[j] * (DADQ(k, i, 3, di) - DADQ(k, j, 3, di));
          JPH(indj + 3 + k, indd + 6) -= fac * s.m[i] * (DADQ(k, i, 3, di) - DADQ(k, j, 3, di));
        }
        fac = 3 * fac1 * s.rij[k];
        for (int di = 0; di < n; ++di) {
          int indd = 7 * di;
          for (int p = 0; p < 3; ++p) JPH(indi + 3 + k, indd + p) += fac * s.m[j] * d.dotdadq[p + 4 * di];
          for (int p = 0; p < 3; ++p) JPH(indj + 3 + k, indd + p) -= fac * s.m[i] * d.dotdadq[p + 4 * di];
          JPH(indi + 3 + k, indd + 6) += fac * s.m[j] * d.dotdadq[3 + 4 * di];
          JPH(indj + 3 + k, indd + 6) -= fac * s.m[i] * d.dotdadq[3 + 4 * di];
        }
      }
    }
  }
}
// ---- ahl21_no_grad.jl:110-158 phisalpha! (no grad) --------------------------
template <class T> inline void phisalpha_nograd(State<T>& s, T h, T alpha) {
  const int n = s.n;
  std::fill(s.a.begin(), s.a.end(), T(0));
  const T G = T(GNEWT);
  T coeff = alpha * (h * h * h) / 96 * 2 * G;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      if (s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) s.rij[k] = SX(k, i) - SX(k, j);
      T r2 = dot3(s.rij);
      T r3 = r2 * m_sqrt(r2);
      T fac2 = G / r3;
      for (int k = 0; k < 3; ++k) {
        T fac = fac2 * s.rij[k];
        SA(k, i) -= s.m[j] * fac;
        SA(k, j) += s.m[i] * fac;
      }
    }
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      if (s.pair[i + n * j]) continue;
      for (int k = 0; k < 3; ++k) { s.aij[k] = SA(k, i) - SA(k, j); s.rij[k] = SX(k, i) - SX(k, j); }
      T r2 = dot3(s.rij);
      T r1 = m_sqrt(r2);
      T ardot = dot3(s.aij, s.rij);
      T fac1 = coeff / (r2 * r2 * r1);
      T fac2 = (2 * G * (s.m[i] + s.m[j]) / r1 + 3 * ardot);
      for (int k = 0; k < 3; ++k) {
        T fac = fac1 * (s.rij[k] * fac2 - r2 * s.aij[k]);
        comp_sum(SV(k, i), SVE(k, i), s.m[j] * fac);
        comp_sum(SV(k, j), SVE(k, j), -s.m[i] * fac);
      }
    }
}
#undef DADQ
#undef JPH

// ---- utils.jl:51-100 copy_submatrix! / ypoc_submatrix! ----------------------
template <class T> inline void copy_submatrix(State<T>& s, Derivs<T>& d, int indi, int indj) {
  const int M = s.M;
  for (int k2 = 0; k2 < M; ++k2)
    for (int k1 = 0; k1 < 7; ++k1) {
      d.jac_tmp1[k1 + 14 * k2] = s.jac_step[indi + k1 + (size_t)M * k2];
      d.jac_err1[k1 + 14 * k2] = s.jac_error[indi + k1 + (size_t)M * k2];
      d.jac_tmp1[7 + k1 + 14 * k2] = s.jac_step[indj + k1 + (size_t)M * k2];
      d.jac_err1[7 + k1 + 14 * k2] = s.jac_error[indj + k1 + (size_t)M * k2];
    }
  for (int k1 = 0; k1 < 7; ++k1) { d.dqdt_tmp1[k1] = s.dqdt[indi + k1]; d.dqdt_tmp1[7 + k1] = s.dqdt[indj + k1]; }
}
template <class T> inline void ypoc_submatrix(State<T>& s, Derivs<T>& d, int indi, int indj) {
  const int M = s.M;
  for (int k2 = 0; k2 < M; ++k2)
    for (int k1 = 0; k1 < 7; ++k1) {
      s.jac_step[indi + k1 + (size_t)M * k2] = d.jac_tmp1[k1 + 14 * k2];
      s.jac_error[indi + k1 + (size_t)M * k2] = d.jac_err1[k1 + 14 * k2];
      s.jac_step[indj + k1 + (size_t)M * k2] = d.jac_tmp1[7 + k1 + 14 * k2];
      s.jac_error[indj + k1 + (size_t)M * k2] = d.jac_err1[7 + k1 + 14 * k2];
    }
  for (int k1 = 0; k1 < 7; ++k1) { s.dqdt[indi + k1] = d.dqdt_ij[k1]; s.dqdt[indj + k1] = d.dqdt_ij[7 + k1]; }
}

// pair update shared by the two sweeps: ahl21.jl:29-43 / :60-76
template <class T> inline void pair_update(State<T>& s, Derivs<T>& d, int i, int j, T h2, bool drift_first) {
  const int M = s.M;
  kepler_driftij_gamma(s, d, i, j, h2, drift_first);
  copy_submatrix(s, d, 7 * i, 7 * j);
  gemm(d.jac_tmp2.data(), d.jac_ij.data(), d.jac_tmp1.data(), 14, 14, M);
  comp_sum_matrix(d.jac_tmp1.data(), d.jac_err1.data(), d.jac_tmp2.data(), (size_t)14 * M);
  for (int k = 0; k < 14; ++k) d.dqdt_ij[k] *= T(0.5);
  gemv(d.tmp14.data(), d.jac_ij.data(), d.dqdt_tmp1.data(), 14, 14);
  for (int k = 0; k < 14; ++k) d.dqdt_ij[k] += d.dqdt_tmp1[k] + d.tmp14[k];
  ypoc_submatrix(s, d, 7 * i, 7 * j);
}

// ---- ahl21.jl:5-95  ahl21!(s,d,h) -------------------------------------------
template <class T> inline void ahl21_grad(State<T>& s, Derivs<T>& d, T h) {
  const int n = s.n, M = s.M;
  const T half = T(0.5), two = T(2);
  T h2 = half * h, h6 = h / 6;
  d.zero_out();
  std::fill(s.dqdt.begin(), s.dqdt.end(), T(0));
  kickfast(s, d, h6);
  for (int k = 0; k < M; ++k) d.dqdt_kick[k] /= 6;
  gemv(d.tmp7n.data(), d.jac_kick.data(), s.dqdt.data(), M, M);
  for (int k = 0; k < M; ++k) s.dqdt[k] += d.dqdt_kick[k] + d.tmp7n[k];
  gemm(d.jac_copy.data(), d.jac_kick.data(), s.jac_step.data(), M, M, M);
  drift_grad(s, h2);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) s.dqdt[7 * i + k] = half * SV(k, i) + h2 * s.dqdt[7 * i + 3 + k];
  comp_sum_matrix(s.jac_step.data(), s.jac_error.data(), d.jac_copy.data(), (size_t)M * M);
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j)
      if (!s.pair[i + n * j]) pair_update(s, d, i, j, h2, true);
  phic(s, d, h);
  phisalpha(s, d, h, two);
  gemm(d.jac_copy.data(), d.jac_phi.data(), s.jac_step.data(), M, M, M);
  gemv(d.tmp7n.data(), d.jac_phi.data(), s.dqdt.data(), M, M);
  for (int k = 0; k < M; ++k) s.dqdt[k] += d.dqdt_phi[k] + d.tmp7n[k];
  comp_sum_matrix(s.jac_step.data(), s.jac_error.data(), d.jac_copy.data(), (size_t)M * M);
  for (int i = n - 2; i >= 0; --i)
    for (int j = n - 1; j >= i + 1; --j)
      if (!s.pair[i + n * j]) pair_update(s, d, i, j, h2, false);
  drift_grad(s, h2);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) s.dqdt[7 * i + k] += half * SV(k, i) + h2 * s.dqdt[7 * i + 3 + k];
  std::fill(d.dqdt_kick.begin(), d.dqdt_kick.end(), T(0));
  kickfast(s, d, h6);
  for (int k = 0; k < M; ++k) d.dqdt_kick[k] /= 6;
  gemv(d.tmp7n.data(), d.jac_kick.data(), s.dqdt.data(), M, M);
  for (int k = 0; k < M; ++k) s.dqdt[k] += d.dqdt_kick[k] + d.tmp7n[k];
  gemm(d.jac_copy.data(), d.jac_kick.data(), s.jac_step.data(), M, M, M);
  comp_sum_matrix(s.jac_step.data(), s.jac_error.data(), d.jac_copy.data(), (size_t)M * M);
}

// ---- ahl21_no_grad.jl:6-18 ahl21!(s,h) --------------------------------------
template <class T> inline void ahl21_nograd(State<T>& s, T h) {
  const int n = s.n;
  T h2 = T(0.5) * h, h6 = h / 6;
  kickfast_nograd(s, h6);
  drift(s, h2);
  for (int i = 0; i < n - 1; ++i)  // drift_kepler! :164-172
    for (int j = i + 1; j < n; ++j)
      if (!s.pair[i + n * j]) kepler_driftij_gamma_nograd(s, i, j, h2, true);
  phic_nograd(s, h);
  phisalpha_nograd(s, h, T(2));
  for (int i = n - 2; i >= 0; --i)  // kepler_drift! :178-186
    for (int j = n - 1; j >= i + 1; --j)
      if (!s.pair[i + n * j]) kepler_driftij_gamma_nograd(s, i, j, h2, false);
  drift(s, h2);
  kickfast_nograd(s, h6);
}

// ---- Integrator.jl:249-259 check_step ----------------------------------------
template <class T> inline T check_step(T t0, T tmax) {
  if (m_abs(tmax) > m_abs(t0)) return jl_sign(tmax);
  if (jl_sign(tmax) != jl_sign(t0)) return jl_sign(tmax);
  return -1 * jl_sign(tmax);
}
// Julia round(Int64, x): ties to even
inline long jl_round(double x) { return (long)std::nearbyint(x); }
inline long jl_round(quad x) { return (long)nearbyintq(x); }

// ---- Integrator.jl:159-197 (intr)(s,time;grad) -------------------------------
template <class T> inline void integrate_to(State<T>& s, T hstep, T time, bool grad) {
  T t0 = s.t;
  long nsteps = std::labs(jl_round((time - t0) / hstep));
  T h = hstep * check_step(t0, time);
  T tmax = t0 + (h * T((double)nsteps));
  Derivs<T>* d = grad ? new Derivs<T>(s.n) : nullptr;
  for (long i = 0; i < nsteps; ++i) { if (grad) ahl21_grad(s, *d, h); else ahl21_nograd(s, h); }
  if (tmax != time) {
    T hf = time - tmax;
    if (grad) ahl21_grad(s, *d, hf); else ahl21_nograd(s, hf);
  }
  s.t = time;
  delete d;
}
// ---- Integrator.jl:211-234 (intr)(s,N;grad) ----------------------------------
template <class T> inline void integrate_nsteps(State<T>& s, T hstep, long N, bool grad) {
  T s2 = T(0);
  T h = hstep;
  if (N < 0) { h = -h; N = -N; }
  Derivs<T>* d = grad ? new Derivs<T>(s.n) : nullptr;
  for (long i = 0; i < N; ++i) {
    if (grad) ahl21_grad(s, *d, h); else ahl21_nograd(s, h);
    comp_sum(s.t, s2, h);
  }
  delete d;
}

// ---- Transits.jl:14-56, 68-110  TransitTiming / TransitParameters ------------
// tt[i + n*k]; dtdq0[i + n*(k + ntt*(q + 7*p))]  (Julia tt[i,k], dtdq0[i,k,q,p]);
// ttbv[c + 3*(i + n*k)]; dtbvdq0[c + 3*(i + n*(k + ntt*(q + 7*p)))].
template <class T> struct TransitOut {
  int n = 0, ntt = 0, ti = 0, ntbv = 1;  // ntbv = 1: TransitTiming, 3: TransitParameters
  std::vector<T> tt, dtdq0, dtdelements, dtbvdq, gsave;
  std::vector<long> count;
  std::vector<int> occs;
  State<T> s_prior;
  TransitOut(int n_, int ntt_, int ti_, int ntbv_) : n(n_), ntt(ntt_), ti(ti_), ntbv(ntbv_), s_prior(n_) {
    tt.assign((size_t)ntbv * n * ntt, T(0));
    dtdq0.assign((size_t)ntbv * n * ntt * 7 * n, T(0));
    dtdelements.assign((size_t)ntbv * n * ntt * 7 * n, T(0));
    dtbvdq.assign((size_t)ntbv * 7 * n, T(0));
    gsave.assign(n, T(0));
    count.assign(n, 0);
    for (int i = 0; i < n; ++i) if (i != ti) occs.push_back(i);
  }
};
// Transits.jl:44-45: ntt = maximum(ceil(|tmax/P_i|)+3) over bodies with finite tmax/P_i
inline int ntt_from_periods(double tmax, const double* periods, int n) {
  int ntt = 0; bool any = false;
  for (int i = 0; i < n; ++i) {
    double q = tmax / periods[i];
    if (!std::isfinite(q)) continue;
    int v = (int)std::ceil(std::fabs(q)) + 3;
    if (!any || v > ntt) ntt = v;
    any = true;
  }
  return ntt;
}

// ---- timing.jl:141-153 g!, gd!, calc_bsky2, calc_vsky -----------------------
template <class T> inline T gsky(int i, int j, const State<T>& s) {
  return (SX(0, j) - SX(0, i)) * (SV(0, j) - SV(0, i)) + (SX(1, j) - SX(1, i)) * (SV(1, j) - SV(1, i));
}
template <class T> inline T gdot(int i, int j, const State<T>& s) {
  const std::vector<T>& q = s.dqdt;
  return ((SX(0, j) - SX(0, i)) * (q[7 * j + 3] - q[7 * i + 3]) + (SX(1, j) - SX(1, i)) * (q[7 * j + 4] - q[7 * i + 4]) +
          (SV(0, j) - SV(0, i)) * (q[7 * j + 0] - q[7 * i + 0]) + (SV(1, j) - SV(1, i)) * (q[7 * j + 1] - q[7 * i + 1]));
}
template <class T> inline T calc_bsky2(const State<T>& s, int i, int j) {
  T a = SX(0, j) - SX(0, i), b = SX(1, j) - SX(1, i);
  return a * a + b * b;
}
template <class T> inline T calc_vsky(const State<T>& s, int i, int j) {
  T a = SV(0, j) - SV(0, i), b = SV(1, j) - SV(1, i);
  return m_sqrt(a * a + b * b);
}
// ---- timing.jl:155-194 dtbvdq! ----------------------------------------------
// dtbvdq[c + ntbv*(k + 7*p)]
template <class T> inline void dtbvdq(int i, int j, const State<T>& s, T* out, int ntbv, T* vsky_out, T* bsky2_out) {
  const int n = s.n, M = s.M;
  T gd = gdot(i, j, s);
  for (int q = 0; q < ntbv * 7 * n; ++q) out[q] = T(0);
  int indj = 7 * j, indi = 7 * i;
#define JS(r_, c_) s.jac_step[(r_) + (size_t)M * (c_)]
  for (int p = 0; p < n; ++p) {
    int indp = 7 * p;
    for (int k = 0; k < 7; ++k) {
      out[0 + ntbv * (k + 7 * p)] =
          -((JS(indj, indp + k) - JS(indi, indp + k)) * (SV(0, j) - SV(0, i)) + (JS(indj + 1, indp + k) - JS(indi + 1, indp + k)) * (SV(1, j) - SV(1, i)) +
            (JS(indj + 3, indp + k) - JS(indi + 3, indp + k)) * (SX(0, j) - SX(0, i)) + (JS(indj + 4, indp + k) - JS(indi + 4, indp + k)) * (SX(1, j) - SX(1, i))) / gd;
    }
  }
  if (ntbv == 3) {
    T vsky = calc_vsky(s, i, j);
    T bsky2 = calc_bsky2(s, i, j);
    const std::vector<T>& q = s.dqdt;
    T dvdt = ((SV(0, j) - SV(0, i)) * (q[7 * j + 3] - q[7 * i + 3]) + (SV(1, j) - SV(1, i)) * (q[7 * j + 4] - q[7 * i + 4])) / vsky;
    for (int p = 0; p < n; ++p) {
      int indp = 7 * p;
      for (int k = 0; k < 7; ++k) {
        out[1 + 3 * (k + 7 * p)] = ((JS(indj + 3, indp + k) - JS(indi + 3, indp + k)) * (SV(0, j) - SV(0, i)) +
                                    (JS(indj + 4, indp + k) - JS(indi + 4, indp + k)) * (SV(1, j) - SV(1, i))) / vsky + dvdt * out[0 + 3 * (k + 7 * p)];
        out[2 + 3 * (k + 7 * p)] = 2 * ((JS(indj, indp + k) - JS(indi, indp + k)) * (SX(0, j) - SX(0, i)) +
                                        (JS(indj + 1, indp + k) - JS(indi + 1, indp + k)) * (SX(1, j) - SX(1, i)));
      }
    }
    *vsky_out = vsky; *bsky2_out = bsky2;
  }
#undef JS
}

inline long& itmax_hits() { static thread_local long c = 0; return c; }
// ---- timing.jl:31-110 findtransit! ------------------------------------------
// i = transited body (tt.ti), j = occultor.  hstat receives the Newton iteration count.
template <class T> inline void findtransit(int i, int j, T dt0, State<T>& s, Derivs<T>& d, TransitOut<T>& tt, bool grad, long* newton_iters) {
  std::fill(s.dqdt.begin(), s.dqdt.end(), T(0));
  T dt = T(1), gd = T(0), gs = T(0), stmp = T(0);
  int iter = 0;
  T tt1 = dt0 + 1, tt2 = dt0 + 2;
  const int ITMAX = 20;
  while (true) {
    tt2 = tt1;
    tt1 = dt0;
    set_state(s, tt.s_prior);
    d.zero_out();
    ahl21_grad(s, d, dt0);
    gs = gsky(i, j, s);
    gd = gdot(i, j, s);
    dt = -gs / gd;
    comp_sum(dt0, stmp, dt);
    iter += 1;
    if (iter >= ITMAX || dt0 == tt1 || dt0 == tt2) break;
  }
  if (newton_iters) *newton_iters += iter;
  if (iter >= ITMAX) itmax_hits() += 1;   // instrumentation only (tests compare how often the 20-iteration cap fires)
  if (grad) {
    set_state(s, tt.s_prior);
    d.zero_out();
    ahl21_grad(s, d, dt0);
  }
  const int n = s.n, ntt = tt.ntt;
  long cnt = tt.count[j] - 1;  // 0-based slot
  if (tt.ntbv == 3) {
    tt.tt[0 + 3 * (j + n * cnt)] = s.t + dt0;
    if (grad) {
      T vsky, bsky2;
      dtbvdq(i, j, s, tt.dtbvdq.data(), 3, &vsky, &bsky2);
      tt.tt[1 + 3 * (j + n * cnt)] = vsky;
      tt.tt[2 + 3 * (j + n * cnt)] = bsky2;
      for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 7; ++k)
          for (int p = 0; p < n; ++p)
            tt.dtdq0[c + 3 * (j + n * (cnt + (size_t)ntt * (k + 7 * p)))] = tt.dtbvdq[c + 3 * (k + 7 * p)];
      return;
    }
    tt.tt[1 + 3 * (j + n * cnt)] = calc_vsky(s, i, j);
    tt.tt[2 + 3 * (j + n * cnt)] = calc_bsky2(s, i, j);
    return;
  }
  tt.tt[j + n * cnt] = s.t + dt0;
  if (grad) {
    dtbvdq(i, j, s, tt.dtbvdq.data(), 1, (T*)nullptr, (T*)nullptr);
    for (int k = 0; k < 7; ++k)
      for (int p = 0; p < n; ++p) tt.dtdq0[j + n * (cnt + (size_t)ntt * (k + 7 * p))] = tt.dtbvdq[k + 7 * p];
  }
}

// ---- timing.jl:3-29 detect_transits! ----------------------------------------
template <class T> inline void detect_transits(State<T>& s, Derivs<T>& d, TransitOut<T>& tt, T intr_h, bool grad, long* newton_iters) {
  const T rstar = T(1e12);
  set_state(tt.s_prior, s);
  for (int i : tt.occs) {
    T gi = gsky(i, tt.ti, s);
    T ri = m_sqrt(SX(0, i) * SX(0, i) + SX(1, i) * SX(1, i) + SX(2, i) * SX(2, i));
    if (gi > T(0) && tt.gsave[i] < T(0) && -SX(2, i) > T(0.25) * ri && ri < rstar) {
      tt.count[i] += 1;
      if (tt.count[i] <= tt.ntt) {
        T dt0 = -gi * intr_h / (gi - tt.gsave[i]);
        set_state(s, tt.s_prior);
        findtransit(tt.ti, i, dt0, s, d, tt, grad, newton_iters);
      }
    }
    tt.gsave[i] = gi;
    set_state(s, tt.s_prior);
  }
}

// ---- timing.jl:112-138 calc_dtdelements! ------------------------------------
template <class T> inline void calc_dtdelements(const State<T>& s, TransitOut<T>& tt) {
  const int n = s.n, M = s.M, ntt = tt.ntt, C = tt.ntbv;
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < n; ++i)
      for (long j = 0; j < tt.count[i]; ++j) {
        if (j >= ntt) continue;
        for (int k = 0; k < n; ++k)
          for (int l = 0; l < 7; ++l) {
            T acc = T(0);
            for (int p = 0; p < n; ++p)
              for (int q = 0; q < 7; ++q)
                acc += tt.dtdq0[c + C * (i + n * (j + (size_t)ntt * (q + 7 * p)))] * s.jac_init[(7 * p + q) + (size_t)M * (7 * k + l)];
            tt.dtdelements[c + C * (i + n * (j + (size_t)ntt * (l + 7 * k)))] = acc;
          }
      }
}

// ---- Transits.jl:140-170 (intr)(s,tt,d;grad) --------------------------------
template <class T> inline void integrate_transits(State<T>& s, TransitOut<T>& tt, T intr_h, T intr_tmax, bool grad, long* newton_iters = nullptr) {
  Derivs<T> d(s.n);
  T t0 = s.t;
  long nsteps = std::labs(jl_round(intr_tmax / intr_h));
  T h = intr_h * check_step(t0, intr_tmax + t0);
  for (int i : tt.occs) tt.gsave[i] = gsky(i, tt.ti, s);
  long istep = 0;
  for (long it = 0; it < nsteps; ++it) {
    if (grad) ahl21_grad(s, d, h); else ahl21_nograd(s, h);
    istep += 1;
    s.t = t0 + (T((double)istep) * h);
    detect_transits(s, d, tt, intr_h, grad, newton_iters);
  }
  if (grad) calc_dtdelements(s, tt);
}

#undef SX
#undef SV
#undef SXE
#undef SVE
#undef SA
}  // namespace nbgo
