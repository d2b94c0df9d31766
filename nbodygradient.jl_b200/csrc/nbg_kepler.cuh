// Two-body Kepler-minus-drift solve in the universal variable gamma, and its analytic Jacobian.
// Device code for sm_100a; one thread solves one pair of one system (lanes = systems).
//
// Replaces, for this path, the reference functions
//   jac_delxv_gamma!         src/integrator/ahl21/ahl21.jl:766-890
//   compute_jacobian_gamma!  src/integrator/ahl21/ahl21.jl:896-1139 (debug=false)
//   G3,H1,H2,H3,H5,H6, cubic1  src/utils.jl:103-396
// Same mathematics and branch structure (elliptic/hyperbolic, drift_first or not, series below
// |gamma|=0.5, repeat-terminated Newton), written for the GPU: FP64 throughout, FMA contraction
// allowed, reciprocals shared between expressions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace nbg {

constexpr double kYear = 365.242;
constexpr double kG = 39.4845 / (kYear * kYear);  // NbodyGradient.jl:14-15
constexpr double kThird = 1.0 / 3.0;

__device__ __forceinline__ double sgn(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

// utils.jl:16-23 (Kahan, scalar form)
__device__ __forceinline__ void ksum(double& val, double& err, double a) {
  err = __dadd_rn(err, a);
  double tmp = __dadd_rn(val, err);
  err = __dadd_rn(__dsub_rn(val, tmp), err);
  val = tmp;
}
// utils.jl:36-46 (Kahan, matrix form)
__device__ __forceinline__ void ksum_m(double& val, double& err, double a) {
  err = __dadd_rn(err, a);
  double tmp = __dadd_rn(val, err);
  err = __dadd_rn(err, __dsub_rn(val, tmp));
  val = tmp;
}

// utils.jl:103-122
__device__ __forceinline__ double cubic1(double a, double b, double c) {
  double a3 = a * kThird;
  double Q = a3 * a3 - b * kThird;
  double R = a3 * a3 * a3 + 0.5 * (-a3 * b + c);
  double R2 = R * R, Q3 = Q * Q * Q;
  if (R2 < Q3) return -c / b;
  double A = -sgn(R) * cbrt(fabs(R) + sqrt(R2 - Q3));
  double B = (A == 0.0) ? 0.0 : Q / A;
  return A + B - a3;
}

// ---- G3 / H-function series (utils.jl:137-165, 180-209, 225-254, 271-303, 320-349, 364-396) ----
// Each is sum_n c_n x2^n with x2 = -sign(beta) gamma^2.  The reference adds terms until the partial sum repeats; the
// coefficients are fixed rationals, so for |gamma| < 0.5 the same sum is a short polynomial, evaluated here by Horner
// with precomputed coefficients (no divisions, no data-dependent trip count; agrees with the term-by-term sum to ~1 ulp).
// For |gamma| >= 0.5 the closed forms are used (the reference does so for gamma >= 0.5 and runs the slowly converging
// series for gamma <= -0.5; both equal the closed form).
#include "nbg_series_coeffs.inc"
template <int NC> __device__ __forceinline__ double horner(const double (&c)[NC], double x) {
  double s = c[0];
#pragma unroll
  for (int q = 1; q < NC; ++q) s = fma(s, x, c[q]);
  return s;
}
__device__ __forceinline__ double series_g3(double x2) { constexpr double c[] = NBG_SER_G3; return horner(c, x2); }
__device__ __forceinline__ double series_h1(double x2) { constexpr double c[] = NBG_SER_H1; return horner(c, x2); }
__device__ __forceinline__ double series_h2(double x2) { constexpr double c[] = NBG_SER_H2; return horner(c, x2); }
__device__ __forceinline__ double series_h3(double x2) { constexpr double c[] = NBG_SER_H3; return horner(c, x2); }
__device__ __forceinline__ double series_h5(double x2) { constexpr double c[] = NBG_SER_H5; return horner(c, x2); }
__device__ __forceinline__ double series_h6(double x2) { constexpr double c[] = NBG_SER_H6; return horner(c, x2); }

// sin/cos (elliptic) or sinh/cosh (hyperbolic) of the half-angle xx = gamma/2 (ahl21.jl:817-821, :836-840).
// |xx| <= 0.5 (every step size an integrator would use): Taylor polynomials; otherwise libdevice.
__device__ __forceinline__ void trig_pair(bool ell, double xx, double& sx, double& cx) {
  if (fabs(xx) <= 0.5) {
    constexpr double sc[] = NBG_SIN_C;
    constexpr double cc[] = NBG_COS_C;
    const double x2 = xx * xx;
    // sin x = x + x^3 P(x^2), cos x = 1 + x^2 C(x^2); the hyperbolic series have the same coefficients with all signs
    // positive, i.e. sinh x = x - x^3 P(-x^2), cosh x = 1 - x^2 C(-x^2)
    const double y = ell ? x2 : -x2;
    const double ps = horner(sc, y), pc = horner(cc, y);
    sx = ell ? fma(xx * x2, ps, xx) : fma(-(xx * x2), ps, xx);
    cx = ell ? fma(x2, pc, 1.0) : fma(-x2, pc, 1.0);
  } else if (ell) {
    sincos(xx, &sx, &cx);
  } else {
    sx = sinh(xx);
    cx = exp(-xx) + sx;
  }
}

__device__ __forceinline__ double G3f(double gamma, double beta, double sqb) {
  if (fabs(gamma) < 0.5) { double x2 = -sgn(beta) * (gamma * gamma); return series_g3(x2) * (-x2 * gamma / (6.0 * beta * sqb)); }
  return (beta >= 0.0) ? (gamma - sin(gamma)) / (sqb * beta) : (gamma - sinh(gamma)) / (sqb * beta);
}
__device__ __forceinline__ double H1f(double gamma, double beta) {
  if (fabs(gamma) < 0.5) { double x2 = -sgn(beta) * (gamma * gamma); return series_h1(x2) * ((x2 * x2) / (12.0 * (beta * beta))); }
  if (beta >= 0.0) { double s = sin(0.5 * gamma); return (4.0 * (s * s) - gamma * sin(gamma)) / (beta * beta); }
  double s = sinh(0.5 * gamma);
  return (-4.0 * (s * s) + gamma * sinh(gamma)) / (beta * beta);
}
__device__ __forceinline__ double H2f(double gamma, double beta, double sqb) {
  if (fabs(gamma) < 0.5) { double x2 = -sgn(beta) * (gamma * gamma); return series_h2(x2) * (-x2 * gamma / (3.0 * beta * sqb)); }
  return (beta >= 0.0) ? (sin(gamma) - gamma * cos(gamma)) / (sqb * beta) : (sinh(gamma) - gamma * cosh(gamma)) / (sqb * beta);
}
__device__ __forceinline__ double H3f(double gamma, double beta, double sqb) {
  if (fabs(gamma) < 0.5) { double x2 = -sgn(beta) * (gamma * gamma); return series_h3(x2) * (-(x2 * x2) * gamma / (beta * sqb)); }
  return (beta >= 0.0) ? (4.0 * sin(gamma) - sin(gamma) * cos(gamma) - 3.0 * gamma) / (beta * sqb)
                       : (4.0 * sinh(gamma) - sinh(gamma) * cosh(gamma) - 3.0 * gamma) / (beta * sqb);
}
__device__ __forceinline__ double H5f(double gamma, double beta, double sqb) {
  if (fabs(gamma) < 0.5) { double x2 = -sgn(beta) * (gamma * gamma); return series_h5(x2) * (-(x2 * x2) * gamma / (beta * sqb)); }
  return (beta >= 0.0) ? (3.0 * sin(gamma) - 2.0 * gamma - gamma * cos(gamma)) / (beta * sqb)
                       : (3.0 * sinh(gamma) - 2.0 * gamma - gamma * cosh(gamma)) / (beta * sqb);
}
__device__ __forceinline__ double H6f(double gamma, double beta) {
  if (fabs(gamma) < 0.5) { double x2 = -sgn(beta) * (gamma * gamma); return series_h6(x2) * (-(x2 * x2 * x2) / (beta * beta)); }
  return (beta >= 0.0) ? (9.0 - 8.0 * cos(gamma) - cos(2.0 * gamma) - 6.0 * gamma * sin(gamma)) / (2.0 * (beta * beta))
                       : (9.0 - 8.0 * cosh(gamma) - cosh(2.0 * gamma) + 6.0 * gamma * sinh(gamma)) / (2.0 * (beta * beta));
}

// Result of one pair solve.  dx[0..2] = delta position, dx[3..5] = delta velocity (relative coordinates);
// jk[r][c]: d(delxv[r]) / d(x0, v0, k, h)[c]  (6 x 8); jm[r]: cancellation-safe mass derivative (6).
struct KepJac { double jk[6][8]; double jm[6]; };
// the 22 scalars jac_delxv_gamma! hands to compute_jacobian_gamma! (ahl21.jl:888)
struct KepScal { double gamma, g0, g1, g2, g3, h1, h2, dfdt, fm1, gmh, dgdtm1, r0, r, r0inv, rinv, k, h, beta, betainv, eta, sqb, zeta; };

// jac_delxv_gamma! (ahl21.jl:766-890).  k = G (m_i + m_j) != 0.
// NOT a template and never inlined: the grad and no-grad paths execute the very same instructions, which is what
// makes x, v bit-identical between them (the reference asserts this: test/test_integrator.jl:184-185).
__device__ __noinline__ void kepler_solve(const double* __restrict__ x0, const double* __restrict__ v0, double k, double h, bool drift_first,
                                          double* __restrict__ delxv, KepScal* __restrict__ P) {
  const double rt0 = x0[0] - h * v0[0], rt1 = x0[1] - h * v0[1], rt2 = x0[2] - h * v0[2];
  const double r0 = drift_first ? sqrt(rt0 * rt0 + rt1 * rt1 + rt2 * rt2) : sqrt(x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2]);
  const double r0inv = 1.0 / r0;
  const double beta = 2.0 * k * r0inv - (v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2]);
  const double betainv = 1.0 / beta;
  const double signb = sgn(beta);
  const double sqb = sqrt(signb * beta);
  const double zeta = k - r0 * beta;
  const double eta = drift_first ? (rt0 * v0[0] + rt1 * v0[1] + rt2 * v0[2]) : (x0[0] * v0[0] + x0[1] * v0[1] + x0[2] * v0[2]);
  double gamma;
  if (zeta != 0.0) {
    double zinv = 6.0 / zeta;
    gamma = cubic1(0.5 * eta * sqb * zinv, r0 * signb * beta * zinv, -h * signb * beta * sqb * zinv);
  } else if (eta != 0.0) {
    double reta = r0 / eta;
    double disc = reta * reta + 2.0 * h / eta;
    gamma = disc > 0.0 ? sqb * (-reta + sqrt(disc)) : h * r0inv * sqb;
  } else {
    gamma = h * r0inv * sqb;
  }
  double gamma1 = 2.0 * gamma, gamma2 = 3.0 * gamma;
  const double c2n = -2.0 * zeta, c3n = 2.0 * eta * signb * sqb, c4n = -sqb * h * beta;
  const double d1 = 2.0 * signb * zeta, d3 = r0 * beta;
  const bool ell = beta > 0.0;
  double sx, cx;
  for (int iter = 0; iter < 20; ++iter) {
    gamma2 = gamma1;
    gamma1 = gamma;
    trig_pair(ell, 0.5 * gamma, sx, cx);
    gamma -= (k * gamma + c2n * sx * cx + c3n * (sx * sx) + c4n) / (d1 * (sx * sx) + c3n * sx * cx + d3);
    if (gamma == gamma2 || gamma == gamma1) break;
  }
  trig_pair(ell, 0.5 * gamma, sx, cx);
  const double g1 = 2.0 * sx * cx / sqb;
  const double g2 = 2.0 * signb * (sx * sx) * betainv;
  const double g0 = 1.0 - beta * g2;
  const double g3 = G3f(gamma, beta, sqb);
  double h1 = 0.0, h2 = 0.0;
  const double r = r0 * g0 + eta * g1 + k * g2;
  const double rinv = 1.0 / r;
  const double dfdt = -k * g1 * rinv * r0inv;
  double fm1, gmh, dgdtm1;
  if (drift_first) {
    fm1 = -k * r0inv * g2;
    gmh = k * r0inv * (h * g2 - r0 * g3);
    dgdtm1 = k * r0inv * rinv * (h * g1 - r0 * g2);
  } else {
    h1 = H1f(gamma, beta);
    h2 = H2f(gamma, beta, sqb);
    fm1 = k * rinv * (g2 - k * r0inv * h1);
    gmh = k * rinv * (r0 * h2 + eta * h1);
    dgdtm1 = -k * rinv * g2;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    delxv[j] = fm1 * x0[j] + gmh * v0[j];
    delxv[3 + j] = dfdt * x0[j] + dgdtm1 * v0[j];
  }
  P->gamma = gamma; P->g0 = g0; P->g1 = g1; P->g2 = g2; P->g3 = g3; P->h1 = h1; P->h2 = h2; P->dfdt = dfdt; P->fm1 = fm1; P->gmh = gmh;
  P->dgdtm1 = dgdtm1; P->r0 = r0; P->r = r; P->r0inv = r0inv; P->rinv = rinv; P->k = k; P->h = h; P->beta = beta; P->betainv = betainv;
  P->eta = eta; P->sqb = sqb; P->zeta = zeta;
}

// The 22 scalars of one pair solve from its inputs and the CONVERGED gamma: everything kepler_solve computes except the cubic guess and
// the Newton iteration.  pair_op_kernel recomputes them from 10 doubles (x0, v0, gamma, k, m_i, m_j) instead of reading all 22 from the
// scalar stream: 12 instead of 32 doubles per pair section cross HBM twice (the kernel is HBM-bound).  The values can differ from the
// trajectory thread's in the last bit (they only feed the Jacobian records).
__device__ __forceinline__ void kepler_scalars(const double* __restrict__ x0, const double* __restrict__ v0, double k, double h, bool drift_first,
                                               double gamma, KepScal* __restrict__ P) {
  const double rt0 = x0[0] - h * v0[0], rt1 = x0[1] - h * v0[1], rt2 = x0[2] - h * v0[2];
  const double r0 = drift_first ? sqrt(rt0 * rt0 + rt1 * rt1 + rt2 * rt2) : sqrt(x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2]);
  const double r0inv = 1.0 / r0;
  const double beta = 2.0 * k * r0inv - (v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2]);
  const double betainv = 1.0 / beta;
  const double signb = sgn(beta);
  const double sqb = sqrt(signb * beta);
  const double zeta = k - r0 * beta;
  const double eta = drift_first ? (rt0 * v0[0] + rt1 * v0[1] + rt2 * v0[2]) : (x0[0] * v0[0] + x0[1] * v0[1] + x0[2] * v0[2]);
  double sx, cx;
  trig_pair(beta > 0.0, 0.5 * gamma, sx, cx);
  const double g1 = 2.0 * sx * cx / sqb;
  const double g2 = 2.0 * signb * (sx * sx) * betainv;
  const double g0 = 1.0 - beta * g2;
  const double g3 = G3f(gamma, beta, sqb);
  double h1 = 0.0, h2 = 0.0;
  const double r = r0 * g0 + eta * g1 + k * g2;
  const double rinv = 1.0 / r;
  const double dfdt = -k * g1 * rinv * r0inv;
  double fm1, gmh, dgdtm1;
  if (drift_first) {
    fm1 = -k * r0inv * g2;
    gmh = k * r0inv * (h * g2 - r0 * g3);
    dgdtm1 = k * r0inv * rinv * (h * g1 - r0 * g2);
  } else {
    h1 = H1f(gamma, beta);
    h2 = H2f(gamma, beta, sqb);
    fm1 = k * rinv * (g2 - k * r0inv * h1);
    gmh = k * rinv * (r0 * h2 + eta * h1);
    dgdtm1 = -k * rinv * g2;
  }
  P->gamma = gamma; P->g0 = g0; P->g1 = g1; P->g2 = g2; P->g3 = g3; P->h1 = h1; P->h2 = h2; P->dfdt = dfdt; P->fm1 = fm1; P->gmh = gmh;
  P->dgdtm1 = dgdtm1; P->r0 = r0; P->r = r; P->r0inv = r0inv; P->rinv = rinv; P->k = k; P->h = h; P->beta = beta; P->betainv = betainv;
  P->eta = eta; P->sqb = sqb; P->zeta = zeta;
}

// compute_jacobian_gamma! (ahl21.jl:896-1139, debug = false)
// _inl: inlined into pair_op_kernel (everything in registers); the __noinline__ wrapper below is what the one-thread-per-
// system kernels call twice per pair (keeps their instruction footprint small).
__device__ __forceinline__ void kepler_jacobian_inl(const KepScal* __restrict__ P, const double* __restrict__ x0, const double* __restrict__ v0,
                                                    bool drift_first, KepJac* __restrict__ J) {
  const double gamma = P->gamma, g0 = P->g0, g1 = P->g1, g2 = P->g2, g3 = P->g3, h1 = P->h1, h2 = P->h2, dfdt = P->dfdt, fm1 = P->fm1,
               gmh = P->gmh, dgdtm1 = P->dgdtm1, r0 = P->r0, r = P->r, r0inv = P->r0inv, rinv = P->rinv, k = P->k, h = P->h, beta = P->beta,
               betainv = P->betainv, eta = P->eta, sqb = P->sqb, zeta = P->zeta;
  const double r0inv2 = r0inv * r0inv, r0inv3 = r0inv2 * r0inv;
  const double rinv2 = rinv * rinv, rinv3 = rinv2 * rinv;
  const double hsq = h * h, ksq = k * k;
  const double r0sq = r0 * r0;
  const double g1inv = 1.0 / g1, g22 = g2 * g2;
  const double h6 = H6f(gamma, beta);
  const double h3 = H3f(gamma, beta, sqb);
  const double h5 = H5f(gamma, beta, sqb);
  const double h8 = -2.0 * h3 + 3.0 * h5;
  const double d = (h + eta * g2 + 2.0 * k * g3) * betainv;
  const double c1 = d - r0 * g3;
  const double c2 = eta * g0 + g1 * zeta;
  const double c3 = d * k + g1 * r0sq;
  const double c17 = r0 - r - g2 * k;
  // coefficient sets: {a}dxx etc multiply (x0[i], v0[i]) x (x0[j], v0[j]) outer products
  double dfm1dxx, dfm1dxv, dfm1dvv, dfm1dh, dfm1dk, dfm1dk2;
  double dgmhdxx, dgmhdxv, dgmhdvv, dgmhdh, dgmhdk, dgmhdk2;
  double ddfdtdxx, ddfdtdxv, ddfdtdvv, ddfdtdk, ddfdtdk2, ddfdtdh;
  double dgdxx, dgdxv, dgdvv, dgdk, dgdk2, dgdh;
  double mass_x_scale;   // prefactor of jac_mass[0..2]
  double mass_x_sign;    // sign of the dgmhdk2 term in jac_mass[0..2]
  if (drift_first) {
    const double g2inv = 1.0 / g2;
    const double c13 = g1 * h - g2 * r0;
    const double c9 = 2.0 * g2 * h - 3.0 * g3 * r0;
    const double c10 = k * (r0inv2 * r0inv2) * (-g2 * r0 * h + k * c9 * betainv - c3 * c13 * rinv);
    const double tkb = 2.0 * k * r0inv - beta;
    const double c24 = r0inv3 * (r0 * tkb * betainv - g1 * c3 * rinv * g2inv);
    dfm1dxx = fm1 * c24;
    dfm1dxv = -fm1 * (g1 * rinv + h * c24);
    dfm1dvv = fm1 * rinv * (-r0 * g2 + k * h6 * betainv * g2inv + h * (2.0 * g1 + h * r * c24));
    dfm1dh = fm1 * (g1 * rinv * (g2inv + tkb) - eta * c24);
    dfm1dk = fm1 * (1.0 / k + g1 * c1 * rinv * r0inv * g2inv - 2.0 * betainv * r0inv);
    const double h4 = -H1f(gamma, beta) * beta;
    dfm1dk2 = (r0 * h4 + k * h6);
    dgmhdxx = c10;
    dgmhdxv = -g2 * k * c13 * rinv * r0inv - h * c10;
    dgmhdvv = 2.0 * g2 * h * k * c13 * rinv * r0inv + hsq * c10 +
              k * betainv * rinv * r0inv * (r0sq * h8 - beta * h * r0 * g22 + (h * k + eta * r0) * h6);
    dgmhdh = g2 * k * r0inv + k * c13 * rinv * r0inv + g2 * k * tkb * c13 * rinv * r0inv - eta * c10;
    dgmhdk = r0inv * (k * c1 * c13 * rinv * r0inv + g2 * h - g3 * r0 - k * c9 * betainv * r0inv);
    dgmhdk2 = (h6 * g3 * ksq + eta * r0 * (h6 + g2 * h4) + r0sq * g0 * h5 + k * eta * g2 * h6 + (g1 * h6 + g3 * h4) * k * r0);
    mass_x_scale = (kG * r0inv) * (kG * r0inv) * betainv * rinv;
    mass_x_sign = -1.0;
    const double c12 = g0 * h - g1 * r0;
    const double c20 = k * (g2 * k + r) - g0 * r0 * zeta;
    const double c21 = (g2 * k - r0) * (beta * c3 - k * g1 * r) * betainv * rinv2 * r0inv3 * g1inv + eta * g1 * rinv * r0inv2 - 2.0 * r0inv2;
    const double c22 = rinv * (-g1 - g0 * g2 * g1inv + g2 * c2 * rinv);
    const double c25 = k * rinv * r0inv2 *
                       (-g2 + k * (c13 - g2 * r0) * betainv * r0inv2 - c13 * r0inv - c12 * c3 * rinv * r0inv2 + c13 * c2 * c3 * rinv2 * r0inv2 -
                        c13 * c20 * betainv * rinv * r0inv2);
    const double c26 = k * rinv2 * r0inv * (-g2 * c12 - g1 * c13 + g2 * c13 * c2 * rinv);
    ddfdtdxx = dfdt * c21;
    ddfdtdxv = dfdt * (c22 - h * c21);
    const double c34 = (-beta * (eta * eta) * g22 - eta * k * h8 - h6 * ksq - 2.0 * beta * eta * r0 * g1 * g2 + (g22 - 3.0 * g1 * g3) * beta * k * r0 -
                        beta * (g1 * g1) * r0sq) * betainv * rinv2 +
                       (eta * g22) * rinv * g1inv + (k * h8) * betainv * rinv * g1inv;
    ddfdtdvv = dfdt * (c34 - 2.0 * h * c22 + hsq * c21);
    ddfdtdk = dfdt * (1.0 / k - betainv * r0inv - c17 * betainv * rinv * r0inv - c1 * (g1 * c2 - g0 * r) * rinv2 * r0inv * g1inv);
    ddfdtdk2 = -(g2 * k - r0) * (beta * r0 * (g3 - g1 * g2) - beta * eta * g22 + k * h3) * betainv * rinv2 * r0inv;
    ddfdtdh = dfdt * (g0 * rinv * g1inv - c2 * rinv2 - tkb * c22 - eta * c21);
    dgdxx = c25;
    dgdxv = c26 - h * c25;
    const double h2b = H2f(gamma, beta, sqb);
    const double c33 = d * k * rinv3 * r0inv * k * (h * g2 - r0 * g3) +
                       k * (-eta * k * g1 * g22 - g1 * g2 * g3 * ksq - r0 * eta * beta * g1 * g22 - r0 * k * g1 * h2b - beta * g22 * g0 * r0sq) * betainv *
                           rinv2 * r0inv;
    dgdvv = c33 - 2.0 * h * c26 + hsq * c25;
    dgdk = rinv * r0inv *
           (-k * (c13 - g2 * r0) * betainv * r0inv + c13 - k * c13 * c17 * betainv * rinv * r0inv + k * c1 * c12 * rinv * r0inv -
            k * c1 * c2 * c13 * rinv2 * r0inv);
    dgdk2 = k * betainv * rinv2 * r0inv *
            (-beta * (eta * eta) * (g22 * g22) + eta * g2 * (g1 * g22 + (g1 * g1) * g3 - 5.0 * g2 * g3) * k + g2 * g3 * h3 * ksq +
             2.0 * eta * r0 * beta * g22 * (g3 - g1 * g2) + (4.0 * g3 - g0 * g3 - g1 * g2) * (g3 - g1 * g2) * r0 * k +
             beta * (2.0 * g1 * g3 * g2 - (g1 * g1) * g22 - (g3 * g3)) * r0sq);
    dgdh = g1 * k * rinv * r0inv + k * c12 * rinv2 * r0inv - k * c2 * c13 * rinv3 * r0inv - tkb * c26 - eta * c25;
  } else {
    const double c14 = r0 * g2 - k * h1;
    const double c15 = eta * h1 + h2 * r0;
    const double c16 = eta * h2 + g1 * gamma * r0 / sqb;
    const double c19 = 4.0 * eta * h1 + 3.0 * h2 * r0;
    const double c23 = h2 * k - r0 * g1;
    const double c20 = k * (g2 * k + r) - g0 * r0 * zeta;
    const double r0inv4 = r0inv2 * r0inv2;
    dfm1dxx = k * rinv3 * betainv * r0inv4 *
              (k * h1 * (r * r) * r0 * (beta - 2.0 * k * r0inv) + beta * c3 * (r * c23 + c14 * c2) + c14 * r * (k * (r - g2 * k) + g0 * r0 * zeta));
    dfm1dxv = k * rinv2 * r0inv * (k * (g2 * h2 + g1 * h1) - 2.0 * g1 * g2 * r0 + g2 * c14 * c2 * rinv);
    dfm1dvv = k * r0inv * rinv2 * betainv *
              (2.0 * eta * k * (g2 * g3 - g1 * h1) + (3.0 * g3 * h2 - 4.0 * h1 * g2) * ksq + beta * g2 * r0 * (3.0 * h1 * k - g2 * r0) +
               c14 * rinv * (-beta * g22 * (eta * eta) + eta * k * (2.0 * g0 * g3 - h2) - h6 * ksq + (-2.0 * eta * g1 * g2 + k * (h1 - 2.0 * g1 * g3)) * beta * r0 -
                             beta * (g1 * g1) * r0sq));
    dfm1dh = (g1 * k - h2 * ksq * r0inv - k * c14 * c2 * rinv * r0inv) * rinv2;
    dfm1dk = rinv * r0inv *
             (4.0 * h1 * ksq * betainv * r0inv - k * h1 - 2.0 * g2 * k * betainv + c14 - k * c14 * c17 * betainv * rinv * r0inv +
              k * (g1 * r0 - k * h2) * c1 * rinv * r0inv - k * c14 * c1 * c2 * rinv2 * r0inv);
    dfm1dk2 = betainv * r0inv * rinv2 *
              (r * (2.0 * eta * k * (g1 * h1 - g3 * g2) + (4.0 * g2 * h1 - 3.0 * g3 * h2) * ksq - eta * r0 * beta * g1 * h1 +
                    (g3 * h2 - 4.0 * g2 * h1) * beta * k * r0 + g2 * h1 * (beta * beta) * r0sq) -
               c14 * (-(eta * eta) * beta * g22 - k * eta * h8 - ksq * h6 - eta * r0 * beta * (g1 * g2 + g0 * g3) + 2.0 * (h1 - g1 * g3) * beta * k * r0 -
                      (g2 - beta * g1 * g3) * beta * r0sq));
    const double rr0 = rinv * r0inv;
    dgmhdxx = k * rinv * r0inv *
              (h2 + k * c19 * betainv * r0inv2 - c16 * c3 * rinv * r0inv2 + c2 * c3 * c15 * (rr0 * rr0) - c15 * c20 * betainv * rinv * r0inv2);
    dgmhdxv = k * rinv2 * (h1 * r - g2 * c16 - g1 * c15 + g2 * c2 * c15 * rinv);
    dgmhdvv = k * betainv * rinv2 *
              (2.0 * (eta * eta) * (g1 * h1 - g2 * g3) + eta * k * (4.0 * g2 * h1 - 3.0 * h2 * g3) + r0 * eta * (4.0 * g0 * h1 - 2.0 * g1 * g3) +
               3.0 * r0 * k * ((g1 + beta * g3) * h1 - g3 * g2) + (g0 * h8 - beta * g1 * (g22 + g1 * g3)) * r0sq -
               c15 * rinv * (beta * g22 * (eta * eta) + eta * k * h8 + h6 * ksq + (2.0 * eta * g1 * g2 - k * (g22 - 3.0 * g1 * g3)) * beta * r0 +
                             beta * (g1 * g1) * r0sq));
    dgmhdk = rinv * (k * c1 * c16 * rinv * r0inv + c15 - k * c15 * c17 * betainv * rinv * r0inv - k * c19 * betainv * r0inv -
                     k * c1 * c2 * c15 * rinv2 * r0inv);
    const double h7 = beta * g1 * g22 - g0 * h8;
    dgmhdk2 = betainv * rinv2 *
              (r * (2.0 * (eta * eta) * (g3 * g2 - g1 * h1) + eta * k * (3.0 * g3 * h2 - 4.0 * g2 * h1) +
                    r0 * eta * (beta * g3 * (g1 * g2 + g0 * g3) - 2.0 * g0 * h6) + (-h6 * (g1 + beta * g3) + g2 * (2.0 * g3 - h2)) * r0 * k +
                    (h7 - (beta * beta) * g1 * (g3 * g3)) * r0sq) -
               c15 * (-beta * (eta * eta) * g22 + eta * k * (-h2 + 2.0 * g0 * g3) - h6 * ksq - r0 * eta * beta * (h2 + 2.0 * g0 * g3) +
                      2.0 * beta * (2.0 * h1 - g22) * r0 * k + beta * (beta * g1 * g3 - g2) * r0sq));
    dgmhdh = k * rinv3 * (r * c16 - c2 * c15);
    mass_x_scale = kG * kG * rinv * r0inv;
    mass_x_sign = 1.0;
    ddfdtdxx = dfdt * (eta * g1 * rinv - 2.0 - g0 * c3 * rinv * r0inv * g1inv + c2 * c3 * r0inv * rinv2 - k * (k * g2 - r0) * betainv * rinv * r0inv) * r0inv2;
    ddfdtdxv = -dfdt * (g0 * g2 * g1inv + (r0 * g1 + eta * g2) * rinv) * rinv;
    ddfdtdvv = -k * rinv3 * r0inv * betainv *
               ((beta * eta * g22 + k * h8) * (r0 * g0 + k * g2) +
                g1 * (-h6 * ksq + (-2.0 * eta * g1 * g2 + (h1 - 2.0 * g1 * g3) * k) * beta * r0 - beta * (g1 * g1) * r0sq));
    ddfdtdk = dfdt * (1.0 / k + c1 * (r0 - g2 * k) * r0inv * rinv2 * g1inv - betainv * r0inv * (1.0 + c17 * rinv));
    ddfdtdk2 = (r0 - g2 * k) * betainv * r0inv * rinv2 * (-eta * beta * g22 + h3 * k + (g3 - g1 * g2) * beta * r0);
    ddfdtdh = dfdt * (r0 - g2 * k) * rinv2 * g1inv;
    dgdxx = rinv2 * r0inv3 * ((eta * g2 + g1 * r0) * k * c3 * rinv + g2 * k * (k * (g2 * k - r) - g0 * r0 * zeta) * betainv);
    dgdxv = k * g2 * rinv3 * (r * g1 + r0 * g1 + eta * g2);
    dgdvv = k * betainv * rinv3 *
            ((eta * eta) * beta * (g22 * g2) - eta * k * g2 * h3 + 3.0 * r0 * eta * beta * g1 * g22 + r0 * k * (-g0 * h6 + 3.0 * beta * g1 * g2 * g3) +
             beta * g2 * (g0 * g2 + (g1 * g1)) * r0sq);
    dgdk = rinv * r0inv * (-r0 * g2 + g2 * k * (r + r0 - g2 * k) * betainv * rinv - k * g1 * c1 * rinv + k * g2 * c1 * c2 * rinv2);
    dgdk2 = betainv * rinv2 *
            (-beta * (eta * eta) * (g22 * g2) + eta * k * g2 * h3 + eta * r0 * beta * g2 * (g3 - 2.0 * g1 * g2) + (h6 - beta * (g22 * g2)) * r0 * k +
             beta * g1 * (g3 - g1 * g2) * r0sq);
    dgdh = k * rinv3 * (g2 * c2 - r * g1);
  }
  const double mass_v_scale = kG * kG * r0inv * rinv;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      J->jk[j][i] = (dfm1dxx * x0[i] + dfm1dxv * v0[i]) * x0[j] + (dgmhdxx * x0[i] + dgmhdxv * v0[i]) * v0[j] + (i == j ? fm1 : 0.0);
      J->jk[j][3 + i] = (dfm1dxv * x0[i] + dfm1dvv * v0[i]) * x0[j] + (dgmhdxv * x0[i] + dgmhdvv * v0[i]) * v0[j] + (i == j ? gmh : 0.0);
      J->jk[3 + j][i] = (ddfdtdxx * x0[i] + ddfdtdxv * v0[i]) * x0[j] + (dgdxx * x0[i] + dgdxv * v0[i]) * v0[j] + (i == j ? dfdt : 0.0);
      J->jk[3 + j][3 + i] = (ddfdtdxv * x0[i] + ddfdtdvv * v0[i]) * x0[j] + (dgdxv * x0[i] + dgdvv * v0[i]) * v0[j] + (i == j ? dgdtm1 : 0.0);
    }
    J->jk[j][6] = dfm1dk * x0[j] + dgmhdk * v0[j];
    J->jk[j][7] = dfm1dh * x0[j] + dgmhdh * v0[j];
    J->jk[3 + j][6] = ddfdtdk * x0[j] + dgdk * v0[j];
    J->jk[3 + j][7] = ddfdtdh * x0[j] + dgdh * v0[j];
    J->jm[j] = mass_x_scale * (dfm1dk2 * x0[j] + mass_x_sign * dgmhdk2 * v0[j]);
    J->jm[3 + j] = mass_v_scale * (ddfdtdk2 * x0[j] + dgdk2 * v0[j]);
  }
}
__device__ __noinline__ void kepler_jacobian(const KepScal* __restrict__ P, const double* __restrict__ x0, const double* __restrict__ v0,
                                             bool drift_first, KepJac* __restrict__ J) {
  kepler_jacobian_inl(P, x0, v0, drift_first, J);
}

}  // namespace nbg
