// Initial-condition layer on the device: orbital elements -> Cartesian x, v and jac_init =
// d(x, v, m) / d(elements, m), one thread per planetary system.
//
// Replaces, for this path,
//   init_nbody(ic::ElementsIC)   src/ics/init_nbody.jl:13-27     (kepcalc :50-105, d_dm :120-162, amatrix :176-188, Sigma m :203-229)
//   kepler_init(t, m, elements, jac_init)   src/ics/kepler_init.jl:66-210
//   ekepler                      src/ics/kepler.jl:1-41
// The hierarchy matrix epsilon (src/ics/setup_hierarchy.jl) is the same for every system of a batch and comes from the host,
// together with what the reference's kepcalc loop derives from it alone: the elements row of each Keplerian.
// jac_kepler (6n x 7n, init_nbody.jl:60) is block sparse -- Keplerian k only depends on the six elements of its own row and
// on the masses of its members -- so only the 6 x 7 block of each Keplerian is kept (42 doubles instead of 2,688 per system).
#pragma once
#include "nbg_kepler.cuh"

namespace nbg {

constexpr int ICN = 16;  // max bodies

struct IcsHierarchy {
  double eps[ICN * ICN];   // eps[i + n*j], Julia column-major (setup_hierarchy.jl)
  int row[ICN];            // elements row (0-based) of Keplerian k = 0..n-2   (kepcalc's i+1+b bookkeeping, init_nbody.jl:66-103)
};

// kepler.jl:1-41
__device__ __forceinline__ double ics_ekepler(double m, double ecc) {
  if (m == 0.0) return 0.0;
  const double pi2 = 6.283185307179586;
  double ms = fmod(m, pi2);
  if (ms != 0.0 && ms < 0.0) ms += pi2;  // Julia mod(): sign of the divisor
  double de0 = ecc * 0.85 * sgn(ms);
  double de1 = 2.0 * de0, de2 = 3.0 * de0;
  for (int iter = 0; iter < 20; ++iter) {
    de2 = de1;
    de1 = de0;
    const double f3 = ecc * cos(de0 + ms), f2 = ecc * sin(de0 + ms);
    de0 = (f2 - de1 * f3) / (1.0 - f3);
    if (de0 == de1 || de0 == de2) break;
  }
  return de0 + m;
}

__device__ __forceinline__ void m3mul(const double* A, const double* B, double* C) {  // 3x3 column-major a[r + 3c]
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r + 3 * c] = A[r] * B[3 * c] + A[r + 3] * B[1 + 3 * c] + A[r + 6] * B[2 + 3 * c];
}
__device__ __forceinline__ void m3vec(const double* A, const double* x, double f, double* y) {
#pragma unroll
  for (int r = 0; r < 3; ++r) y[r] = f * (A[r] * x[0] + A[r + 3] * x[1] + A[r + 6] * x[2]);
}

// kepler_init.jl:66-210.  el = (P, t0, ecosw, esinw, I, Omega); jac[r][c]: rows x(3), v(3); columns the six elements and the mass.
__device__ __noinline__ void ics_kepler_init(double time, double mass, const double* __restrict__ el, double* __restrict__ xo, double* __restrict__ vo,
                                             double (*__restrict__ jac)[7]) {
  const double pi = 3.141592653589793;
  const double period = el[0], n = 2.0 * pi / period, t0 = el[1];
  const double semi = cbrt(kG * mass * (period * period) / 4.0 / (pi * pi));
  const double dsemidp = 2.0 * kThird * semi / period, dsemidm = kThird * semi / mass;
  const double ecosw = el[2], esinw = el[3];
  const double ecc = sqrt(esinw * esinw + ecosw * ecosw);
  const double deccdecos = ecc != 0.0 ? ecosw / ecc : 0.0, deccdesin = ecc != 0.0 ? esinw / ecc : 0.0;
  const double sq = sqrt(1.0 - ecc * ecc);
  const double den1 = esinw - ecosw - ecc;
  double tp;
  if (ecc == 0.0) tp = t0 - 3.0 * period / 4.0;
  else tp = t0 - sq / n * ecosw / (1.0 - esinw) - 2.0 / n * atan2(sqrt(1.0 - ecc) * (esinw + ecosw + ecc), sqrt(1.0 + ecc) * den1);
  const double dtpdp = (tp - t0) / period;
  const double fac = sqrt((1.0 - ecc) / (1.0 + ecc));
  const double den2 = 1.0 / (den1 * den1);
  const double theta = fac * (esinw + ecosw + ecc) / den1;
  const double epc = ecc + ecosw;
  const double dthetadecc = (epc * epc + 2.0 * (1.0 - ecc * ecc) * esinw - esinw * esinw) / (sq * (1.0 + ecc)) * den2;
  const double dthetadecos = 2.0 * fac * esinw * den2, dthetadesin = -2.0 * fac * (ecosw + ecc) * den2;
  const double omes = 1.0 - esinw, t2 = 2.0 / n / (1.0 + theta * theta);
  const double dtpdecc = ecc / sq / n * ecosw / omes - t2 * dthetadecc;
  const double dtpdecos = dtpdecc * deccdecos - sq / n / omes - t2 * dthetadecos;
  const double dtpdesin = dtpdecc * deccdesin - sq / n * ecosw / (omes * omes) - t2 * dthetadesin;
  const double m = n * (time - tp), dmdp = -m / period, dmdtp = -n;
  const double ekep = ics_ekepler(m, ecc);
  double sinekep, cosekep;
  sincos(ekep, &sinekep, &cosekep);
  const double r = semi * (1.0 - ecc * cosekep), denom = semi / r;
  const double dekepdecos = sinekep * denom * deccdecos, dekepdesin = sinekep * denom * deccdesin, dekepdm = denom;
  double sincap, coscap, sininc, cosinc;
  sincos(el[5], &sincap, &coscap);
  sincos(el[4], &sininc, &cosinc);
  const double cosw = ecc != 0.0 ? ecosw / ecc : 1.0, sinw = ecc != 0.0 ? esinw / ecc : 0.0;
  const double P1[9] = {cosw, sinw, 0, -sinw, cosw, 0, 0, 0, 1};
  const double P2[9] = {1, 0, 0, 0, cosinc, sininc, 0, -sininc, cosinc};
  const double P3[9] = {coscap, sincap, 0, -sincap, coscap, 0, 0, 0, 1};
  const double Mi[9] = {0, 0, 0, 0, -sininc, cosinc, 0, -cosinc, -sininc};
  const double Mc[9] = {-sincap, coscap, 0, -coscap, -sincap, 0, 0, 0, 0};
  double P32[9], P321[9], P3i[9], P3i1[9], Pc2[9], Pc21[9];
  m3mul(P3, P2, P32); m3mul(P32, P1, P321);
  m3mul(P3, Mi, P3i); m3mul(P3i, P1, P3i1);
  m3mul(Mc, P2, Pc2); m3mul(Pc2, P1, Pc21);
  const double xplane[3] = {semi * (cosekep - ecc), semi * (sq * sinekep), 0.0};
  const double vplane[3] = {-sinekep, sq * cosekep, 0.0};
  const double xrot[3] = {-xplane[1], xplane[0], 0.0}, vrot[3] = {-vplane[1], vplane[0], 0.0};
  double x[3], dxdekep[3], dxdecc[3], t3[3], dxdecos[3], dxdesin[3], dxdinc[3], dxdcom[3];
  m3vec(P321, xplane, 1.0, x);
  m3vec(P321, vplane, semi, dxdekep);
  const double e1[3] = {cosekep, sinekep / sq, 0.0};
  m3vec(P321, e1, -semi / ecc, dxdecc);
  m3vec(P32, xplane, 1.0 / ecc, t3);
#pragma unroll
  for (int k = 0; k < 3; ++k) dxdecos[k] = dxdecc[k] * deccdecos + t3[k];
  m3vec(P32, xrot, 1.0 / ecc, t3);
#pragma unroll
  for (int k = 0; k < 3; ++k) dxdesin[k] = dxdecc[k] * deccdesin + t3[k];
  m3vec(P3i1, xplane, 1.0, dxdinc);
  m3vec(Pc21, xplane, 1.0, dxdcom);
  const double vs = n * semi * denom;
  double v[3], dvdekep[3], dvdecc[3], dvdecos[3], dvdesin[3], dvdinc[3], dvdcom[3];
  m3vec(P321, vplane, vs, v);
  const double e2[3] = {-cosekep, -sq * sinekep, 0.0};
  m3vec(P321, e2, vs, t3);
#pragma unroll
  for (int k = 0; k < 3; ++k) dvdekep[k] = -v[k] * ecc * sinekep * denom + t3[k];
  const double e3[3] = {0.0, -ecc / sq * cosekep, 0.0};
  m3vec(P321, e3, vs, t3);
#pragma unroll
  for (int k = 0; k < 3; ++k) dvdecc[k] = -v[k] / ecc + v[k] * cosekep * denom + t3[k];
  m3vec(P32, vplane, vs / ecc, t3);
#pragma unroll
  for (int k = 0; k < 3; ++k) dvdecos[k] = dvdecc[k] * deccdecos + t3[k];
  m3vec(P32, vrot, vs / ecc, t3);
#pragma unroll
  for (int k = 0; k < 3; ++k) dvdesin[k] = dvdecc[k] * deccdesin + t3[k];
  m3vec(P3i1, vplane, vs, dvdinc);
  m3vec(Pc21, vplane, vs, dvdcom);
  const double c1 = dekepdm * (dmdp + dmdtp * dtpdp), c2 = dekepdm * dmdtp;  // dtpdt0 = 1
  const double c3 = dekepdm * dmdtp * dtpdecos + dekepdecos, c4 = dekepdm * dmdtp * dtpdesin + dekepdesin;
  const bool ez = ecc == 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double dxda = x[k] / semi, dvda = v[k] / semi;
    jac[k][0] = dxda * dsemidp + dxdekep[k] * c1;
    jac[k][1] = dxdekep[k] * c2;
    jac[k][2] = ez ? 0.0 : dxdecos[k] + dxdekep[k] * c3;
    jac[k][3] = ez ? 0.0 : dxdesin[k] + dxdekep[k] * c4;
    jac[k][4] = dxdinc[k];
    jac[k][5] = dxdcom[k];
    jac[k][6] = dxda * dsemidm;
    jac[3 + k][0] = -v[k] / period + dvda * dsemidp + dvdekep[k] * c1;
    jac[3 + k][1] = dvdekep[k] * c2;
    jac[3 + k][2] = ez ? 0.0 : dvdecos[k] + dvdekep[k] * c3;
    jac[3 + k][3] = ez ? 0.0 : dvdesin[k] + dvdekep[k] * c4;
    jac[3 + k][4] = dvdinc[k];
    jac[3 + k][5] = dvdcom[k];
    jac[3 + k][6] = dvda * dsemidm;
    xo[k] = x[k];
    vo[k] = v[k];
  }
}

// n x n column-major helpers on thread-local storage
__device__ __forceinline__ void ics_inverse(double* a, double* inv, int n) {  // Gauss-Jordan with partial pivoting (Julia inv())
  for (int q = 0; q < n * n; ++q) inv[q] = 0.0;
  for (int i = 0; i < n; ++i) inv[i + n * i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(a[c + n * c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(a[r + n * c]) > best) { best = fabs(a[r + n * c]); piv = r; }
    if (piv != c)
      for (int k = 0; k < n; ++k) {
        double t = a[c + n * k]; a[c + n * k] = a[piv + n * k]; a[piv + n * k] = t;
        t = inv[c + n * k]; inv[c + n * k] = inv[piv + n * k]; inv[piv + n * k] = t;
      }
    const double pinv = 1.0 / a[c + n * c];
    for (int k = 0; k < n; ++k) { a[c + n * k] *= pinv; inv[c + n * k] *= pinv; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[r + n * c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; ++k) { a[r + n * k] -= f * a[c + n * k]; inv[r + n * k] -= f * inv[c + n * k]; }
    }
  }
}

// One thread per system.  elements: [sys][7][n] (Julia elements[i, c], column-major, system slowest);
// x, v, m: SoA [q][ld]; jac_init: [sys][col][row] (M x M, Julia column-major).
__global__ void __launch_bounds__(64) ics_kernel(const double* __restrict__ elements, IcsHierarchy H, int n, long nsys, size_t ld, double t0,
                                                 double* __restrict__ X, double* __restrict__ V, double* __restrict__ Mm, double* __restrict__ jac_init, int write_xv) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  const int M = 7 * n;
  const double* el = elements + (size_t)sys * 7 * n;
  double m[ICN];
  for (int i = 0; i < n; ++i) m[i] = el[i];
  // amatrix (init_nbody.jl:176-188) and Sigma m (:203-229)
  double A[ICN * ICN], Ainv[ICN * ICN], SM[ICN * ICN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int l = 0; l < n; ++l) s += (H.eps[i + n * j] == H.eps[i + n * l]) ? m[l] : 0.0;
      SM[i + n * j] = s;
      A[i + n * j] = (H.eps[i + n * j] * m[j]) / s;
    }
  ics_inverse(A, Ainv, n);  // destroys A
  // kepcalc (init_nbody.jl:50-105): one Keplerian per row of eps but the last
  double rk[ICN][3], rdk[ICN][3], jk[ICN][6][7];
  for (int k = 0; k < n - 1; ++k) {
    double mu = 0.0;
    for (int j = 0; j < n; ++j) mu += H.eps[k + n * j] != 0.0 ? m[j] : 0.0;
    double e6[6];
    for (int c = 0; c < 6; ++c) e6[c] = el[(1 + c) * n + H.row[k]];
    ics_kepler_init(t0, mu, e6, rk[k], rdk[k], jk[k]);
  }
  for (int c = 0; c < 3; ++c) { rk[n - 1][c] = 0.0; rdk[n - 1][c] = 0.0; }
  // x = A^-1 r, v = A^-1 rdot   (init_nbody.jl:20-24)
  if (write_xv) {
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        double sx = 0.0, sv = 0.0;
        for (int l = 0; l < n; ++l) { sx += Ainv[i + n * l] * rk[l][c]; sv += Ainv[i + n * l] * rdk[l][c]; }
        X[(size_t)(3 * i + c) * ld + sys] = sx;
        V[(size_t)(3 * i + c) * ld + sys] = sv;
      }
    for (int i = 0; i < n; ++i) Mm[(size_t)i * ld + sys] = m[i];
  }
  if (!jac_init) return;
  // d_dm (init_nbody.jl:120-162): jac_init = blockdiag(A^-1) jac_kepler + d(A^-1)/dm (r, rdot)
  double* J = jac_init + (size_t)sys * M * M;
  for (int q = 0; q < M * M; ++q) J[q] = 0.0;
  for (int ii = 0; ii < n; ++ii) {
    // element columns: Keplerian k fills the columns of body k + 1 (init_nbody.jl:90-94, whatever elements row it read)
    for (int k = 0; k < n - 1; ++k) {
      const double a = Ainv[ii + n * k];
      const int body = k + 1;
      for (int e = 0; e < 6; ++e)
        for (int r = 0; r < 6; ++r) J[(size_t)(7 * body + e) * M + 7 * ii + r] = a * jk[k][r][e];
    }
    // mass columns, Kepler part: members of Keplerian k feel its mass derivative
    for (int j = 0; j < n; ++j)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < n - 1; ++k) s += H.eps[k + n * j] != 0.0 ? Ainv[ii + n * k] * jk[k][r][6] : 0.0;
        J[(size_t)(7 * j + 6) * M + 7 * ii + r] = s;
      }
    J[(size_t)(7 * ii + 6) * M + 7 * ii + 6] = 1.0;
  }
  // mass columns, d(A^-1)/dm_k = - A^-1 (dA/dm_k) A^-1
  double dA[ICN * ICN], T1[ICN * ICN];
  for (int k = 0; k < n; ++k) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        const double e = H.eps[i + n * j], sm = SM[i + n * j];
        dA[i + n * j] = ((k == j ? 1.0 : 0.0) * e) / sm - ((e == H.eps[i + n * k] ? 1.0 : 0.0) * e * m[j] / (sm * sm));
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += Ainv[i + n * l] * dA[l + n * j];
        T1[i + n * j] = -s;
      }
    for (int ii = 0; ii < n; ++ii) {
      double dx[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
      for (int l = 0; l < n; ++l) {
        double d = 0.0;  // dAinvdm_k[ii][l]
        for (int q = 0; q < n; ++q) d += T1[ii + n * q] * Ainv[q + n * l];
        for (int c = 0; c < 3; ++c) { dx[c] += d * rk[l][c]; dv[c] += d * rdk[l][c]; }
      }
      for (int c = 0; c < 3; ++c) {
        J[(size_t)(7 * k + 6) * M + 7 * ii + c] += dx[c];
        J[(size_t)(7 * k + 6) * M + 7 * ii + 3 + c] += dv[c];
      }
    }
  }
}

// ---- Cartesian state -> orbital elements -------------------------------------------------------------------
// get_orbital_elements(s, ic)   src/outputs/elements.jl:108-137  (get_relative_positions :25-35, get_relative_masses :38-48,
//                               hvec :55-59, calc_Omega :61-65, calc_omega :67-82, convert_to_elements :84-106)
// One thread per system; x, v, m: SoA [q][ld] (the resident state).  out[sys][body][11] = (m, P, t0 = 0, ecosw, esinw, I, Omega, a, e,
// omega, tp), the fields of the reference's Elements; body 0 carries only its mass.
__device__ __forceinline__ void ics_convert_to_elements(const double* x, const double* v, double Gmm, double* o) {
  const double R = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const double V = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double hx = x[1] * v[2] - x[2] * v[1], hy = -(x[0] * v[2] - x[2] * v[0]);
  const double hz = x[0] * v[1] - x[1] * v[0];
  if (hz >= 0.0) hy *= -1.0; else hx *= -1.0;
  const double h = sqrt(hx * hx + hy * hy + hz * hz);
  const double xv = x[0] * v[0] + x[1] * v[1] + x[2] * v[2];
  const double Rdot = sgn(xv) * sqrt(V * V - (h / R) * (h / R));
  const double a = 1.0 / ((2.0 / R) - (V * V) / Gmm);
  const double e = sqrt(1.0 - (h * h / (Gmm * a)));
  const double I = acos(hz / h);
  double Om = 0.0, wpf = 0.0;
  if (I != 0.0) {
    const double si = sin(I);
    Om = atan2(hx / (h * si), hy / (h * si));
    const double swpf = x[2] / (R * si);
    const double cwpf = ((x[0] / R) + sin(Om) * swpf * cos(I)) / cos(Om);
    wpf = atan2(swpf, cwpf);
  }
  const double sinf = a * Rdot * (1.0 - e * e) / (h * e), cosf = (a * (1.0 - e * e) / R - 1.0) / e;
  const double w = wpf - atan2(sinf, cosf);
  const double P = 6.283185307179586 * sqrt(a * a * a / Gmm);
  const double n = 6.283185307179586 / P;
  const double ecw = e * cos(w), esw = e * sin(w);
  const double tp = fmod(-sqrt(1.0 - e * e) * ecw / (n * (1.0 - esw)) -
                             (2.0 / n) * atan2(sqrt(1.0 - e) * (esw + ecw + e), sqrt(1.0 + e) * (esw - ecw - e)), P);
  o[0] = P; o[1] = 0.0; o[2] = ecw; o[3] = esw; o[4] = I; o[5] = Om; o[6] = a; o[7] = e; o[8] = w; o[9] = tp;
}

__global__ void __launch_bounds__(64) elements_out_kernel(const double* __restrict__ X, const double* __restrict__ V, const double* __restrict__ Mm,
                                                          IcsHierarchy H, int n, long nsys, size_t ld, double* __restrict__ out) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  double m[ICN];
  for (int i = 0; i < n; ++i) m[i] = Mm[(size_t)i * ld + sys];
  double* o = out + (size_t)sys * n * 11;
  for (int q = 0; q < 11 * n; ++q) o[q] = 0.0;
  o[0] = m[0];
  int i = 1, b = 0;
  while (i < n) {  // the Keplerian bookkeeping of get_orbital_elements (elements.jl:118-135)
    if (H.eps[(i - 1) + 0] == 0.0) b += 1;
    const int q = i - 1 + b;
    if (q >= 0 && q < n - 1) {
      // row q of amat x (init_nbody.jl:176-188), and G sum |eps| m (elements.jl:38-48)
      double xr[3] = {0, 0, 0}, vr[3] = {0, 0, 0}, mu = 0.0;
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += (H.eps[q + n * j] == H.eps[q + n * l]) ? m[l] : 0.0;
        const double aqj = (H.eps[q + n * j] * m[j]) / s;
        for (int k = 0; k < 3; ++k) { xr[k] += aqj * X[(size_t)(3 * j + k) * ld + sys]; vr[k] += aqj * V[(size_t)(3 * j + k) * ld + sys]; }
        mu += fabs(H.eps[q + n * j]) * m[j];
      }
      o[11 * i] = m[i];
      ics_convert_to_elements(xr, vr, kG * mu, o + 11 * i + 1);
    }
    if (b > 0) b -= 2; else if (b < 0) i += 1;
    i += 1;
  }
}

}  // namespace nbg
