// ORACLE — TEST INFRASTRUCTURE ONLY (see nbg_oracle.hpp header).
//
// IC layer restatement: orbital elements -> Cartesian (x, v) and jac_init.
// Follows src/ics/init_nbody.jl, src/ics/kepler_init.jl:66-210, src/ics/kepler.jl:1-41,
// src/ics/setup_hierarchy.jl (fully-nested hierarchies as produced by
// ElementsIC(t0, N::Int, elements) -> hierarchy([N,1,...,1]); an explicit
// epsilon matrix may be passed for anything else).
#pragma once
#include "nbg_oracle.hpp"

namespace nbgo {

// small dense helpers (column-major n x n)
template <class T> inline std::vector<T> mat_inv(const std::vector<T>& A, int n) {
  // Gauss-Jordan with partial pivoting (stand-in for LAPACK getrf/getri behind Julia's inv()).
  std::vector<T> a(A), inv((size_t)n * n, T(0));
  for (int i = 0; i < n; ++i) inv[i + (size_t)n * i] = T(1);
  for (int c = 0; c < n; ++c) {
    int piv = c;
    T best = m_abs(a[c + (size_t)n * c]);
    for (int r = c + 1; r < n; ++r)
      if (m_abs(a[r + (size_t)n * c]) > best) { best = m_abs(a[r + (size_t)n * c]); piv = r; }
    if (piv != c)
      for (int k = 0; k < n; ++k) { std::swap(a[c + (size_t)n * k], a[piv + (size_t)n * k]); std::swap(inv[c + (size_t)n * k], inv[piv + (size_t)n * k]); }
    T pinv = T(1) / a[c + (size_t)n * c];
    for (int k = 0; k < n; ++k) { a[c + (size_t)n * k] *= pinv; inv[c + (size_t)n * k] *= pinv; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      T f = a[r + (size_t)n * c];
      if (f == T(0)) continue;
      for (int k = 0; k < n; ++k) { a[r + (size_t)n * k] -= f * a[c + (size_t)n * k]; inv[r + (size_t)n * k] -= f * inv[c + (size_t)n * k]; }
    }
  }
  return inv;
}
template <class T> inline std::vector<T> mat_mul(const std::vector<T>& A, const std::vector<T>& B, int n) {
  std::vector<T> C((size_t)n * n, T(0));
  gemm(C.data(), A.data(), B.data(), n, n, n);
  return C;
}

// setup_hierarchy.jl:9-29 + nlevel for bins = [1,1,...]: row i has -1 for bodies 0..i, +1 for body i+1; last row all -1.
inline std::vector<double> nested_hierarchy(int n) {
  std::vector<double> e((size_t)n * n, 0.0);
  for (int i = 0; i < n - 1; ++i) {
    for (int j = 0; j <= i; ++j) e[i + (size_t)n * j] = -1.0;
    e[i + (size_t)n * (i + 1)] = 1.0;
  }
  for (int j = 0; j < n; ++j) e[(n - 1) + (size_t)n * j] = -1.0;
  return e;
}

template <class T> struct ElementsIC {
  int n = 0;
  std::vector<T> elements;  // n x 7 column-major: elements[i + n*c]; c = m,P,t0,ecosw,esinw,I,Omega
  std::vector<T> eps, amat, m;
  T t0 = T(0);
};

// init_nbody.jl:203-229 delta / Sigma m
template <class T> inline T kdelta(T a, T b) { return a == b ? T(1) : T(0); }
template <class T> inline T sum_m(const std::vector<T>& m, int i, int j, const std::vector<T>& eps, int n) {
  T s = T(0);
  for (int l = 0; l < n; ++l) s += m[l] * kdelta(eps[i + (size_t)n * j], eps[i + (size_t)n * l]);
  return s;
}
// init_nbody.jl:176-188 amatrix
template <class T> inline void amatrix(ElementsIC<T>& ic) {
  const int n = ic.n;
  ic.amat.assign((size_t)n * n, T(0));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) ic.amat[i + (size_t)n * j] = (ic.eps[i + (size_t)n * j] * ic.m[j]) / sum_m(ic.m, i, j, ic.eps, n);
}
// InitialConditions.jl:142-162 ElementsIC(t0, H::Matrix, elements)
template <class T> inline ElementsIC<T> make_elements_ic(T t0, int n, const T* elements_colmajor, const double* eps_or_null) {
  ElementsIC<T> ic;
  ic.n = n; ic.t0 = t0;
  ic.elements.assign(elements_colmajor, elements_colmajor + (size_t)n * 7);
  std::vector<double> e = eps_or_null ? std::vector<double>(eps_or_null, eps_or_null + (size_t)n * n) : nested_hierarchy(n);
  ic.eps.resize((size_t)n * n);
  for (size_t q = 0; q < e.size(); ++q) ic.eps[q] = T(e[q]);
  ic.m.resize(n);
  for (int i = 0; i < n; ++i) ic.m[i] = ic.elements[i];
  amatrix(ic);
  return ic;
}

// kepler.jl:1-41 ekepler
template <class T> inline T ekepler(T m, T ecc) {
  if (m == T(0)) return T(0);
  T pi2 = T(2 * PI);
  T ms = m_fmod(m, pi2);  // Julia mod(): result takes the sign of the divisor
  if (ms != T(0) && ((ms < T(0)) != (pi2 < T(0)))) ms += pi2;
  T de0 = ecc * T(0.85) * jl_sign(ms);
  T de1 = 2 * de0, de2 = 3 * de0;
  int iter = 0;
  while (true) {
    de2 = de1;
    de1 = de0;
    T f3 = ecc * m_cos(de0 + ms);
    T f2 = ecc * m_sin(de0 + ms);
    de0 = (f2 - de1 * f3) / (1 - f3);
    iter += 1;
    if (iter >= 20 || de0 == de1 || de0 == de2) break;
  }
  return de0 + m;
}

template <class T> struct M3 { T a[9]; T& operator()(int r, int c) { return a[r + 3 * c]; } T operator()(int r, int c) const { return a[r + 3 * c]; } };
template <class T> inline M3<T> mm(const M3<T>& A, const M3<T>& B) {
  M3<T> C;
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) { T s = T(0); for (int k = 0; k < 3; ++k) s += A(r, k) * B(k, c); C(r, c) = s; }
  return C;
}
template <class T> inline M3<T> ms(const M3<T>& A, T f) { M3<T> C; for (int q = 0; q < 9; ++q) C.a[q] = A.a[q] * f; return C; }
template <class T> inline M3<T> md(const M3<T>& A, T f) { M3<T> C; for (int q = 0; q < 9; ++q) C.a[q] = A.a[q] / f; return C; }
template <class T> struct V3 { T a[3]; T& operator[](int k) { return a[k]; } T operator[](int k) const { return a[k]; } };
template <class T> inline V3<T> mv(const M3<T>& A, const V3<T>& x) {
  V3<T> y;
  for (int r = 0; r < 3; ++r) { T s = T(0); for (int k = 0; k < 3; ++k) s += A(r, k) * x[k]; y[r] = s; }
  return y;
}
template <class T> inline V3<T> vs(const V3<T>& x, T f) { return V3<T>{{x[0] * f, x[1] * f, x[2] * f}}; }
template <class T> inline V3<T> vd(const V3<T>& x, T f) { return V3<T>{{x[0] / f, x[1] / f, x[2] / f}}; }
template <class T> inline V3<T> va(const V3<T>& x, const V3<T>& y) { return V3<T>{{x[0] + y[0], x[1] + y[1], x[2] + y[2]}}; }
template <class T> inline V3<T> vneg(const V3<T>& x) { return V3<T>{{-x[0], -x[1], -x[2]}}; }

// kepler_init.jl:66-210: elements (P,t0,ecosw,esinw,I,Omega) of one Keplerian -> x, v and the
// 7x7 Jacobian jac (column-major jac[r+7*c]) of (x,v,m) w.r.t. (P,t0,ecosw,esinw,I,Omega,m).
template <class T> inline void kepler_init(T time, T mass, const T* el, T* xo, T* vo, T* jac) {
  const T G = T(GNEWT), third = T(THIRD), pi = T(PI);
  T period = el[0];
  T n = T(2 * PI) / period;
  T t0 = el[1];
  T semi = m_cbrt(G * mass * (period * period) / 4 / T(PI * PI));
  T dsemidp = 2 * third * semi / period;
  T dsemidm = third * semi / mass;
  T ecosomega = el[2], esinomega = el[3];
  T ecc = m_sqrt(esinomega * esinomega + ecosomega * ecosomega);
  T deccdecos = ecc != T(0) ? ecosomega / ecc : T(0);
  T deccdesin = ecc != T(0) ? esinomega / ecc : T(0);
  T sqrt1mecc2 = m_sqrt(T(1) - ecc * ecc);
  T den1 = esinomega - ecosomega - ecc;
  T tp;
  if (ecc == T(0)) tp = t0 - 3 * period / 4;
  else tp = (t0 - sqrt1mecc2 / n * ecosomega / (T(1) - esinomega) -
             2 / n * m_atan2(m_sqrt(T(1) - ecc) * (esinomega + ecosomega + ecc), m_sqrt(T(1) + ecc) * den1));
  T dtpdp = (tp - t0) / period;
  T fac = m_sqrt((T(1) - ecc) / (T(1) + ecc));
  T den2 = T(1) / (den1 * den1);
  T theta = fac * (esinomega + ecosomega + ecc) / den1;
  T epc = ecc + ecosomega;
  T dthetadecc = (epc * epc + 2 * (T(1) - ecc * ecc) * esinomega - esinomega * esinomega) / (sqrt1mecc2 * (T(1) + ecc)) * den2;
  T dthetadecos = 2 * fac * esinomega * den2;
  T dthetadesin = -2 * fac * (ecosomega + ecc) * den2;
  T omes = T(1) - esinomega;
  T dtpdecc = ecc / sqrt1mecc2 / n * ecosomega / omes - 2 / n / (T(1) + theta * theta) * dthetadecc;
  T dtpdecos = dtpdecc * deccdecos - sqrt1mecc2 / n / omes - 2 / n / (T(1) + theta * theta) * dthetadecos;
  T dtpdesin = dtpdecc * deccdesin - sqrt1mecc2 / n * ecosomega / (omes * omes) - 2 / n / (T(1) + theta * theta) * dthetadesin;
  T dtpdt0 = T(1);
  T m = n * (time - tp);
  T dmdp = -m / period;
  T dmdtp = -n;
  T ekep = ekepler(m, ecc);
  T cosekep = m_cos(ekep), sinekep = m_sin(ekep);
  T r = semi * (T(1) - ecc * cosekep);
  T denom = semi / r;
  T dekepdecos = sinekep * denom * deccdecos;
  T dekepdesin = sinekep * denom * deccdesin;
  T dekepdm = denom;
  T inc = el[4], capomega = el[5];
  T coscap = m_cos(capomega), sincap = m_sin(capomega);
  T cosomega = ecc != T(0) ? ecosomega / ecc : T(1);
  T sinomega = ecc != T(0) ? esinomega / ecc : T(0);
  T cosinc = m_cos(inc), sininc = m_sin(inc);
  const T Z = T(0), O = T(1);
  M3<T> P1{{cosomega, sinomega, Z, -sinomega, cosomega, Z, Z, Z, O}};
  M3<T> P2{{O, Z, Z, Z, cosinc, sininc, Z, -sininc, cosinc}};
  M3<T> P3{{coscap, sincap, Z, -sincap, coscap, Z, Z, Z, O}};
  M3<T> P321 = mm(mm(P3, P2), P1);
  V3<T> xplane{{semi * (cosekep - ecc), semi * (sqrt1mecc2 * sinekep), semi * Z}};
  V3<T> vplane{{-sinekep, sqrt1mecc2 * cosekep, Z}};
  V3<T> x = mv(P321, xplane);
  V3<T> dxda = vd(x, semi);
  V3<T> dxdekep = mv(ms(P321, semi), vplane);
  M3<T> P32 = mm(P3, P2);
  V3<T> dxdecc = mv(md(ms(ms(P321, T(-1)), semi), ecc), V3<T>{{cosekep, sinekep / sqrt1mecc2, Z}});
  V3<T> dxdecos = va(vs(dxdecc, deccdecos), mv(md(P32, ecc), xplane));
  V3<T> dxdesin = va(vs(dxdecc, deccdesin), mv(md(P32, ecc), V3<T>{{-xplane[1], xplane[0], Z}}));
  M3<T> Mi{{Z, Z, Z, Z, -sininc, cosinc, Z, -cosinc, -sininc}};
  M3<T> Mc{{-sincap, coscap, Z, -coscap, -sincap, Z, Z, Z, Z}};
  V3<T> dxdinc = mv(mm(mm(P3, Mi), P1), xplane);
  V3<T> dxdcom = mv(mm(mm(Mc, P2), P1), xplane);
  M3<T> Pv = ms(ms(ms(P321, n), semi), denom);
  V3<T> v = mv(Pv, vplane);
  V3<T> dvda = vd(v, semi);
  V3<T> dvdp = vd(vneg(v), period);
  V3<T> dvdekep = va(vs(vs(vs(vneg(v), ecc), sinekep), denom), mv(Pv, V3<T>{{-cosekep, -sqrt1mecc2 * sinekep, Z}}));
  V3<T> dvdecc = va(va(vd(vneg(v), ecc), vs(vs(v, cosekep), denom)), mv(Pv, V3<T>{{Z, -ecc / sqrt1mecc2 * cosekep, Z}}));
  M3<T> P32v = md(ms(ms(ms(P32, n), semi), denom), ecc);
  V3<T> dvdecos = va(vs(dvdecc, deccdecos), mv(P32v, vplane));
  V3<T> dvdesin = va(vs(dvdecc, deccdesin), mv(P32v, V3<T>{{-vplane[1], vplane[0], Z}}));
  V3<T> dvdinc = mv(ms(ms(ms(mm(mm(P3, Mi), P1), n), semi), denom), vplane);
  V3<T> dvdcom = mv(ms(ms(ms(mm(mm(Mc, P2), P1), n), semi), denom), vplane);
  for (int q = 0; q < 49; ++q) jac[q] = T(0);
#define JC(r_, c_) jac[(r_) + 7 * (c_)]
  T c1 = dekepdm * (dmdp + dmdtp * dtpdp);
  T c2 = dekepdm * dmdtp * dtpdt0;
  T c3 = dekepdm * dmdtp * dtpdecos + dekepdecos;
  T c4 = dekepdm * dmdtp * dtpdesin + dekepdesin;
  for (int k = 0; k < 3; ++k) {
    JC(k, 0) = dxda[k] * dsemidp + dxdekep[k] * dekepdm * (dmdp + dmdtp * dtpdp);
    JC(k, 1) = dxdekep[k] * dekepdm * dmdtp * dtpdt0;
    JC(k, 2) = ecc != T(0) ? dxdecos[k] + dxdekep[k] * c3 : T(0);
    JC(k, 3) = ecc != T(0) ? dxdesin[k] + dxdekep[k] * c4 : T(0);
    JC(k, 4) = dxdinc[k];
    JC(k, 5) = dxdcom[k];
    JC(k, 6) = dxda[k] * dsemidm;
    JC(3 + k, 0) = dvdp[k] + dvda[k] * dsemidp + dvdekep[k] * dekepdm * (dmdp + dmdtp * dtpdp);
    JC(3 + k, 1) = dvdekep[k] * dekepdm * dmdtp * dtpdt0;
    JC(3 + k, 2) = ecc != T(0) ? dvdecos[k] + dvdekep[k] * c3 : T(0);
    JC(3 + k, 3) = ecc != T(0) ? dvdesin[k] + dvdekep[k] * c4 : T(0);
    JC(3 + k, 4) = dvdinc[k];
    JC(3 + k, 5) = dvdcom[k];
    JC(3 + k, 6) = dvda[k] * dsemidm;
  }
  (void)c1; (void)c2; (void)pi;
  JC(6, 6) = T(1);
#undef JC
  for (int k = 0; k < 3; ++k) { xo[k] = x[k]; vo[k] = v[k]; }
}

// init_nbody.jl:13-27 + kepcalc :50-105 + d_dm :120-162.
// Outputs x[k+3*i], v[k+3*i], jac_init[r + M*c] (M = 7n).
template <class T> inline void init_nbody(const ElementsIC<T>& ic, std::vector<T>& x, std::vector<T>& v, std::vector<T>& jac_init) {
  const int n = ic.n, M = 7 * n;
  std::vector<T> rk((size_t)n * 3, T(0)), rdk((size_t)n * 3, T(0));  // rk[i + n*k]
  std::vector<T> jac_kepler((size_t)6 * n * M, T(0));               // (6n x 7n): jk[r + 6n*c]
  T j21[49];
  int i = 1, b = 0;  // 1-based as in the reference loop
  while (i < n) {
    T mu = T(0);
    for (int j = 0; j < n; ++j) if (ic.eps[(i - 1) + (size_t)n * j] != T(0)) mu += ic.m[j];
    if (ic.eps[(i - 1) + 0] == T(0)) b += 1;
    int row = i + b;  // 0-based row of elements = (i+1+b)-1
    T el[6];
    for (int c = 0; c < 6; ++c) el[c] = ic.elements[row + (size_t)n * (1 + c)];
    T r3[3], v3[3];
    kepler_init(ic.t0, mu, el, r3, v3, j21);
    for (int k = 0; k < 3; ++k) { rk[(i - 1) + (size_t)n * k] = r3[k]; rdk[(i - 1) + (size_t)n * k] = v3[k]; }
    for (int j = 0; j < 6; ++j)
      for (int k = 0; k < 6; ++k) jac_kepler[((i - 1) * 6 + j) + (size_t)6 * n * (i * 7 + k)] = j21[j + 7 * k];
    for (int j = 0; j < n; ++j)
      if (ic.eps[(i - 1) + (size_t)n * j] != T(0))
        for (int k = 0; k < 6; ++k) jac_kepler[((i - 1) * 6 + k) + (size_t)6 * n * (j * 7 + 6)] = j21[k + 7 * 6];
    if (b > 0) b -= 2; else if (b < 0) b = 0;
    i += 1;
  }
  // d_dm
  jac_init.assign((size_t)M * M, T(0));
  std::vector<T> Ainv = mat_inv(ic.amat, n);
  std::vector<std::vector<T>> dAinvdm(n);
  for (int k = 0; k < n; ++k) {
    std::vector<T> dA((size_t)n * n, T(0));
    for (int ii = 0; ii < n; ++ii)
      for (int j = 0; j < n; ++j) {
        T sm = sum_m(ic.m, ii, j, ic.eps, n);
        T e = ic.eps[ii + (size_t)n * j];
        dA[ii + (size_t)n * j] = ((kdelta(T(k), T(j)) * e) / sm) - ((kdelta(e, ic.eps[ii + (size_t)n * k])) * e * ic.m[j] / (sm * sm));
      }
    std::vector<T> t1 = mat_mul(Ainv, dA, n);
    for (auto& q : t1) q = -q;
    dAinvdm[k] = mat_mul(t1, Ainv, n);
  }
  for (int ii = 0; ii < n; ++ii) {
    for (int k = 0; k < n; ++k)
      for (int j = 0; j < 3; ++j)
        for (int l = 0; l < M; ++l) {
          jac_init[(7 * ii + j) + (size_t)M * l] += Ainv[ii + (size_t)n * k] * jac_kepler[(6 * k + j) + (size_t)6 * n * l];
          jac_init[(7 * ii + 3 + j) + (size_t)M * l] += Ainv[ii + (size_t)n * k] * jac_kepler[(6 * k + 3 + j) + (size_t)6 * n * l];
        }
    for (int k = 0; k < n; ++k)
      for (int c = 0; c < 3; ++c) {
        T dx = T(0), dv = T(0);
        for (int l = 0; l < n; ++l) { dx += dAinvdm[k][ii + (size_t)n * l] * rk[l + (size_t)n * c]; dv += dAinvdm[k][ii + (size_t)n * l] * rdk[l + (size_t)n * c]; }
        jac_init[(7 * ii + c) + (size_t)M * (7 * k + 6)] += dx;
        jac_init[(7 * ii + 3 + c) + (size_t)M * (7 * k + 6)] += dv;
      }
    jac_init[(7 * ii + 6) + (size_t)M * (7 * ii + 6)] = T(1);
  }
  x.assign((size_t)3 * n, T(0)); v.assign((size_t)3 * n, T(0));
  for (int ii = 0; ii < n; ++ii)
    for (int c = 0; c < 3; ++c) {
      T sx = T(0), sv = T(0);
      for (int l = 0; l < n; ++l) { sx += Ainv[ii + (size_t)n * l] * rk[l + (size_t)n * c]; sv += Ainv[ii + (size_t)n * l] * rdk[l + (size_t)n * c]; }
      x[c + 3 * ii] = sx; v[c + 3 * ii] = sv;
    }
}

// Integrator.jl:82-103 State(ic)
template <class T> inline State<T> make_state(const ElementsIC<T>& ic) {
  State<T> s(ic.n);
  init_nbody(ic, s.x, s.v, s.jac_init);
  s.m = ic.m;
  s.t = ic.t0;
  return s;
}


// ---- src/outputs/elements.jl: Cartesian state -> orbital elements ------------------------------------------------------------
// get_relative_positions :25-35, get_relative_masses :38-48, hvec :55-59, calc_Omega :61-65, calc_omega :67-82,
// convert_to_elements :84-106, get_orbital_elements :108-137.
// out[body][11] = (m, P, t0 = 0, ecosw, esinw, I, Omega, a, e, omega, tp) -- the fields of Elements; body 0 carries only its mass.
template <class T> inline void convert_to_elements(const T* x, const T* v, T Gmm, T* out10) {
  const T R = m_sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  const T V = m_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  // hvec: cross(r, rdot) with the reference's sign conventions (:21, :55-59)
  T hx = x[1] * v[2] - x[2] * v[1], hy = -(x[0] * v[2] - x[2] * v[0]), hz = x[0] * v[1] - x[1] * v[0];
  if (hz >= T(0)) hy *= -1; else hx *= -1;
  const T h = m_sqrt(hx * hx + hy * hy + hz * hz);
  const T xv = x[0] * v[0] + x[1] * v[1] + x[2] * v[2];
  const T Rdot = jl_sign(xv) * m_sqrt(V * V - (h / R) * (h / R));
  const T a = T(1) / ((T(2) / R) - (V * V) / Gmm);
  const T e = m_sqrt(T(1) - (h * h / (Gmm * a)));
  const T I = m_acos(hz / h);
  T Om = T(0);
  if (I != T(0)) { const T sO = hx / (h * m_sin(I)), cO = hy / (h * m_sin(I)); Om = m_atan2(sO, cO); }
  T wpf = T(0);
  if (I != T(0)) {
    const T swpf = x[2] / (R * m_sin(I));
    const T cwpf = ((x[0] / R) + m_sin(Om) * swpf * m_cos(I)) / m_cos(Om);
    wpf = m_atan2(swpf, cwpf);
  }
  const T sinf = a * Rdot * (T(1) - e * e) / (h * e), cosf = (a * (T(1) - e * e) / R - T(1)) / e;
  const T w = wpf - m_atan2(sinf, cosf);
  const T P = T(2 * PI) * m_sqrt(a * a * a / Gmm);
  const T n = T(2 * PI) / P;
  const T ecw = e * m_cos(w), esw = e * m_sin(w);
  // Julia's % is the truncated remainder (rem), as fmod
  const T tp = m_fmod(-m_sqrt(T(1) - e * e) * ecw / (n * (T(1) - esw)) -
                          (T(2) / n) * m_atan2(m_sqrt(T(1) - e) * (esw + ecw + e), m_sqrt(T(1) + e) * (esw - ecw - e)), P);
  out10[0] = P; out10[1] = T(0); out10[2] = ecw; out10[3] = esw; out10[4] = I; out10[5] = Om; out10[6] = a; out10[7] = e; out10[8] = w; out10[9] = tp;
}
template <class T> inline void get_orbital_elements(const ElementsIC<T>& ic, const T* x, const T* v, T* out /* n x 11, row-major [body][field] */) {
  const int n = ic.n;
  for (int q = 0; q < 11 * n; ++q) out[q] = T(0);
  std::vector<T> X(3 * (size_t)n, T(0)), Vv(3 * (size_t)n, T(0)), mu(n > 1 ? n - 1 : 0, T(0));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k)
      for (int j = 0; j < n; ++j) {
        X[k + 3 * (size_t)i] += ic.amat[i + (size_t)n * j] * x[k + 3 * j];
        Vv[k + 3 * (size_t)i] += ic.amat[i + (size_t)n * j] * v[k + 3 * j];
      }
  for (int i = 0; i < n - 1; ++i) {
    for (int j = 0; j < n; ++j) mu[i] += m_abs(ic.eps[i + (size_t)n * j]) * ic.m[j];
    mu[i] *= T(GNEWT);
  }
  out[0] = ic.m[0];
  int i = 1, b = 0;
  while (i < n) {
    if (ic.eps[(i - 1) + 0] == T(0)) b += 1;
    const int q = i - 1 + b;   // 0-based index of X[i+b]
    out[11 * i] = ic.m[i];
    convert_to_elements(&X[3 * (size_t)q], &Vv[3 * (size_t)q], mu[q], out + 11 * i + 1);
    if (b > 0) b -= 2; else if (b < 0) i += 1;
    i += 1;
  }
}

}  // namespace nbgo
