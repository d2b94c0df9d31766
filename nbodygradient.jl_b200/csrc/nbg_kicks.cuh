// Fast-kick pairs: the parts of the AHL21 map that act only on pairs flagged in s.pair (all-false unless the caller
// sets it by hand: src/integrator/Integrator.jl:91, test/test_kickfast.jl:24-31, test/test_phic.jl:25-31).
//
// Replaces, for this path,
//   kickfast!(s,d,h)   src/integrator/ahl21/ahl21.jl:337-386   (no-grad: ahl21_no_grad.jl:35-56)
//   phic!(s,d,h)       src/integrator/ahl21/ahl21.jl:392-552   (no-grad: ahl21_no_grad.jl:62-104)
// and phisalpha! restricted to the NON-flagged pairs (ahl21.jl:566,599: `if ~s.pair[i,j]`).
// One thread per system as everywhere in the trajectory stage.  x, v (Kahan) and dq/dh are advanced here; the Jacobians
// jac_kick and jac_phi are described by compact per-pair records (PF doubles, three sets per step: first kick, phic+phisalpha,
// second kick) which phi_dense_kernel expands into dense 3N x 4N operators for the Jacobian kernel.  A record carries, besides
// the phisalpha fields, a direct-kick coefficient kappa and the pair's class (1 = flagged):
//   dv_i = sum_{j != i} { m_j [ Rm_p (dx_i - dx_j) + S_p (da_i^c - da_j^c) + us_p r_ij (dm_i + dm_j) ] + F_ij dm_j
//                         - kappa_p [ m_j T_p (dx_i - dx_j) + gam_ij dm_j ] },       c = class of pair p,
//   da_i^c = - sum_{l != i, class(i,l) = c} { m_l T_il (dx_i - dx_l) + gam_il dm_l }.
// kickfast!: kappa = h, everything else zero.  phic!: kappa = 2h/3, fac1 = (h^3/36) G / r^5, no 2 G (m_i+m_j)/r term (us = 0).
#pragma once
#include "nbg_step.cuh"

namespace nbg {

constexpr int PF_KAPPA = 22;
constexpr int PF_CLASS = 23;

__device__ __forceinline__ bool kicked(const KMask& kmask, int p) { return kmask.bit(p); }

// kickfast!(s,d,hk) over the flagged pairs.  first: the call at the start of the step (dq/dh is still zero there, ahl21.jl:9-14).
template <bool GRAD, int EMIT>
__device__ __forceinline__ void kick_section(Body& b, double* dq, int n, double hk, const KMask& kmask, bool first, const Emit& em, size_t pf_base) {
  double dvacc[3 * NMAX];
  if (GRAD) for (int q = 0; q < 3 * n; ++q) dvacc[q] = 0.0;
  int p = 0;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j, ++p) {
      double rec[PF];
      if (EMIT) {
#pragma unroll
        for (int f = 0; f < PF; ++f) rec[f] = 0.0;
      }
      if (kicked(kmask, p)) {
        double r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = b.x[3 * i + k] - b.x[3 * j + k];
        const double r2inv = 1.0 / (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        const double r3inv = r2inv * sqrt(r2inv);
        const double fac2 = hk * kG * r3inv;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double fac = fac2 * r[k];
          ksum(b.v[3 * i + k], b.ve[3 * i + k], -b.m[j] * fac);
          ksum(b.v[3 * j + k], b.ve[3 * j + k], b.m[i] * fac);
          if (GRAD) {  // dqdt_kick / 6 with hk = h/6: the acceleration over 6 (ahl21.jl:364-365, :11-14)
            dvacc[3 * i + k] -= b.m[j] * fac / hk / 6.0;
            dvacc[3 * j + k] += b.m[i] * fac / hk / 6.0;
          }
        }
        if (GRAD && !first) {  // jac_kick * dqdt: - hk m_j T (dq_x_i - dq_x_j)   (mass entries of dqdt are zero)
          double w[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) w[k] = dq[6 * i + k] - dq[6 * j + k];
          const double g3 = kG * r3inv;
          const double f3 = 3.0 * g3 * r2inv * (r[0] * w[0] + r[1] * w[1] + r[2] * w[2]);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double gw = g3 * w[k] - f3 * r[k];
            dvacc[3 * i + k] -= hk * b.m[j] * gw;
            dvacc[3 * j + k] += hk * b.m[i] * gw;
          }
        }
        if (EMIT) {
#pragma unroll
          for (int k = 0; k < 3; ++k) rec[PF_R + k] = r[k];
          rec[PF_G3] = kG * r3inv;
          rec[PF_G5] = 3.0 * kG * r3inv * r2inv;
          rec[PF_MI] = b.m[i];
          rec[PF_MJ] = b.m[j];
          rec[PF_KAPPA] = hk;
          rec[PF_CLASS] = 1.0;
        }
      }
      if (EMIT) em.put_record<PF>(pf_base + (size_t)p * PF, rec);
    }
  if (GRAD)
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) dq[6 * i + 3 + k] = first ? dvacc[3 * i + k] : dq[6 * i + 3 + k] + dvacc[3 * i + k];
}

// phic!(s,d,h) over the flagged pairs followed by phisalpha!(s,d,h,2) over the others; both add into the same jac_phi /
// dqdt_phi in the reference (ahl21.jl:46-54), so one dense operator describes them.
template <bool GRAD, int EMIT>
__device__ __forceinline__ void phi_kicked_section(Body& b, double* dq, int n, double h, const KMask& kmask, const Emit& em, size_t pf_base) {
  double a[3 * NMAX], da[3 * NMAX], dvacc[3 * NMAX];
  if (GRAD) for (int q = 0; q < 3 * n; ++q) dvacc[q] = 0.0;
#pragma unroll 1
  for (int cls = 1; cls >= 0; --cls) {  // 1: phic! on flagged pairs, then 0: phisalpha! on the rest
    const double coeff = cls ? (h * h * h) / 36.0 * kG : 2.0 * (h * h * h) / 96.0 * 2.0 * kG;  // ahl21.jl:399, :564
    const double kappa = cls ? 2.0 * h / 3.0 : 0.0;
    for (int q = 0; q < 3 * n; ++q) { a[q] = 0.0; da[q] = 0.0; }
    // accelerations of this class (+ for phic! the direct kick 2h/3 a, ahl21.jl:409-416) and their dq/dh
    int p = 0;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j, ++p) {
        if ((int)kicked(kmask, p) != cls) continue;
        double r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = b.x[3 * i + k] - b.x[3 * j + k];
        const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        double g3;
        if (cls) { const double r2inv = 1.0 / r2; g3 = kG * (r2inv * sqrt(r2inv)); }
        else g3 = kG / (r2 * sqrt(r2));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double fac = g3 * r[k];
          if (cls) {
            const double facv = fac * 2.0 * h / 3.0;
            ksum(b.v[3 * i + k], b.ve[3 * i + k], -b.m[j] * facv);
            ksum(b.v[3 * j + k], b.ve[3 * j + k], b.m[i] * facv);
            if (GRAD) { dvacc[3 * i + k] -= 1.0 / h * b.m[j] * facv; dvacc[3 * j + k] += 1.0 / h * b.m[i] * facv; }
          }
          a[3 * i + k] -= b.m[j] * fac;
          a[3 * j + k] += b.m[i] * fac;
        }
        if (GRAD) {
          double w[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) w[k] = dq[6 * i + k] - dq[6 * j + k];
          const double f3 = 3.0 * g3 / r2 * (r[0] * w[0] + r[1] * w[1] + r[2] * w[2]);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double gw = g3 * w[k] - f3 * r[k];
            da[3 * i + k] -= b.m[j] * gw;
            da[3 * j + k] += b.m[i] * gw;
            if (cls) { dvacc[3 * i + k] -= kappa * b.m[j] * gw; dvacc[3 * j + k] += kappa * b.m[i] * gw; }
          }
        }
      }
    // force-gradient term of this class
    p = 0;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j, ++p) {
        if ((int)kicked(kmask, p) != cls) continue;
        double r[3], aij[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { aij[k] = a[3 * i + k] - a[3 * j + k]; r[k] = b.x[3 * i + k] - b.x[3 * j + k]; }
        const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        const double r1 = sqrt(r2);
        const double ardot = aij[0] * r[0] + aij[1] * r[1] + aij[2] * r[2];
        const double fac1 = coeff / (r2 * r2 * r1);
        const double gmu = cls ? 0.0 : kG * (b.m[i] + b.m[j]);
        const double fac2 = cls ? 3.0 * ardot : (2.0 * gmu / r1 + 3.0 * ardot);
        double F[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          F[k] = fac1 * (r[k] * fac2 - r2 * aij[k]);
          ksum(b.v[3 * i + k], b.ve[3 * i + k], b.m[j] * F[k]);
          ksum(b.v[3 * j + k], b.ve[3 * j + k], -b.m[i] * F[k]);
        }
        if (GRAD || EMIT) {
          double Rm[3][3];
          const double r2inv = 1.0 / r2;
          const double gr3 = gmu / (r2 * r1);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double f5 = -5.0 * F[k] * r2inv;
            const double fd = -2.0 * fac1 * (r[k] * gr3 + aij[k]);
            const double f3r = 3.0 * fac1 * r[k];
#pragma unroll
            for (int q = 0; q < 3; ++q) Rm[k][q] = f5 * r[q] + fd * r[q] + f3r * aij[q] + (k == q ? fac1 * fac2 : 0.0);
          }
          if (GRAD) {
            double w[3], wa[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { w[k] = dq[6 * i + k] - dq[6 * j + k]; wa[k] = da[3 * i + k] - da[3 * j + k]; }
            const double rwa = r[0] * wa[0] + r[1] * wa[1] + r[2] * wa[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const double dF = Rm[k][0] * w[0] + Rm[k][1] * w[1] + Rm[k][2] * w[2] + fac1 * (3.0 * r[k] * rwa - r2 * wa[k]);
              dvacc[3 * i + k] += b.m[j] * (3.0 / h * F[k] + dF);
              dvacc[3 * j + k] -= b.m[i] * (3.0 / h * F[k] + dF);
            }
          }
          if (EMIT) {
            double rec[PF];
#pragma unroll
            for (int k = 0; k < 3; ++k) { rec[PF_R + k] = r[k]; rec[PF_F + k] = F[k]; }
            const double g3 = kG / (r2 * r1);
            rec[PF_G3] = g3;
            rec[PF_G5] = 3.0 * g3 * r2inv;
            rec[PF_MI] = b.m[i];
            rec[PF_MJ] = b.m[j];
            rec[PF_FAC1] = fac1;
            rec[PF_R2] = r2;
            rec[PF_US] = cls ? 0.0 : 2.0 * kG * fac1 / r1;
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
              for (int q = 0; q < 3; ++q) rec[PF_RM + 3 * k + q] = Rm[k][q];
            rec[PF_KAPPA] = kappa;
            rec[PF_CLASS] = (double)cls;
            em.put_record<PF>(pf_base + (size_t)p * PF, rec);
          }
        }
      }
  }
  if (GRAD) for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) dq[6 * i + 3 + k] += dvacc[3 * i + k];
}

// dq: d(state)/dh, 6 entries per body (x then v); mass entries are identically zero and not stored.
// em.base points at this step's region of the operator stream (used when EMIT).
// kmask: bit p set = pair p (rx_pair_index order) is a fast-kick pair (s.pair[i,j]); 0 = the reference default.
// KICKS = false compiles the fast-kick branches out (the default path pays nothing for them).
template <bool GRAD, int EMIT, bool KICKS = false>
__device__ void ahl21_step(Body& b, double* dq, int n, double h, const Emit& em, const KMask& kmask = KMask{}) {
  const bool kicks = KICKS && kmask.any();
  const double h2 = 0.5 * h;
  const int P = npairs(n);
  // fill!(s.dqdt,0); kickfast!; drift_grad!/drift!; dqdt[x] = v/2 + h2 dqdt[v]   (ahl21.jl:8-21)
  if (kicks) kick_section<GRAD, EMIT>(b, dq, n, h / 6.0, kmask, true, em, phi_rec_offset(n, 0));
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ksum(b.x[3 * i + k], b.xe[3 * i + k], h2 * b.v[3 * i + k]);
      if (GRAD) {
        if (kicks) dq[6 * i + k] = 0.5 * b.v[3 * i + k] + h2 * dq[6 * i + 3 + k];
        else { dq[6 * i + k] = 0.5 * b.v[3 * i + k]; dq[6 * i + 3 + k] = 0.0; }
      }
    }
  }
  int rec = 0;
  for (int i = 0; i < n - 1; ++i) {
    BodyRegs bi, bj, bn;
    load_body<GRAD>(b, dq, i, bi);
    load_body<GRAD>(b, dq, i + 1, bn);
    for (int j = i + 1; j < n; ++j, ++rec) {
      bj = bn;
      if (j + 1 < n) load_body<GRAD>(b, dq, j + 1, bn);  // in flight while pair (i, j) is solved
      if (!kicks || !kicked(kmask, rec)) pair_section<GRAD, EMIT>(bi, bj, h2, true, em, (size_t)rec * KF);
      store_body<GRAD>(b, dq, j, bj);
    }
    store_body<GRAD>(b, dq, i, bi);
  }
  if (kicks) phi_kicked_section<GRAD, EMIT>(b, dq, n, h, kmask, em, phi_rec_offset(n, 1));
  else phisalpha_section<GRAD, EMIT>(b, dq, n, h, em, (size_t)2 * P * KF);
  for (int i = n - 2; i >= 0; --i) {
    BodyRegs bi, bj, bn;
    load_body<GRAD>(b, dq, i, bi);
    load_body<GRAD>(b, dq, n - 1, bn);
    for (int j = n - 1; j >= i + 1; --j, ++rec) {
      bj = bn;
      if (j - 1 >= i + 1) load_body<GRAD>(b, dq, j - 1, bn);
      if (!kicks || !kicked(kmask, i * n - i * (i + 1) / 2 + (j - i - 1))) pair_section<GRAD, EMIT>(bi, bj, h2, false, em, (size_t)rec * KF);
      store_body<GRAD>(b, dq, j, bj);
    }
    store_body<GRAD>(b, dq, i, bi);
  }
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      ksum(b.x[3 * i + k], b.xe[3 * i + k], h2 * b.v[3 * i + k]);
      if (GRAD) dq[6 * i + k] += 0.5 * b.v[3 * i + k] + h2 * dq[6 * i + 3 + k];
    }
  }
  if (kicks) kick_section<GRAD, EMIT>(b, dq, n, h / 6.0, kmask, false, em, phi_rec_offset(n, 2));
}

// timing.jl:141-150  g!, gd!   (i = transited body, j = occultor)
__device__ __forceinline__ double gsky(const Body& b, int i, int j) {
  return (b.x[3 * j] - b.x[3 * i]) * (b.v[3 * j] - b.v[3 * i]) + (b.x[3 * j + 1] - b.x[3 * i + 1]) * (b.v[3 * j + 1] - b.v[3 * i + 1]);
}
__device__ __forceinline__ double gdot(const Body& b, const double* dq, int i, int j) {
  return ((b.x[3 * j] - b.x[3 * i]) * (dq[6 * j + 3] - dq[6 * i + 3]) + (b.x[3 * j + 1] - b.x[3 * i + 1]) * (dq[6 * j + 4] - dq[6 * i + 4]) +
          (b.v[3 * j] - b.v[3 * i]) * (dq[6 * j] - dq[6 * i]) + (b.v[3 * j + 1] - b.v[3 * i + 1]) * (dq[6 * j + 1] - dq[6 * i + 1]));
}


// d g / d t along the exact flow, from the instantaneous Newtonian accelerations of bodies i and j: the derivative used by the
// gradient-free pre-iterations of the transit Newton solve (transit_kernel).  The reference's own derivative (gd!, timing.jl:147-150:
// the step-size derivative of the AHL21 MAP) differs from it by the integrator's truncation error only.
__device__ __forceinline__ double gdot_flow(const Body& b, int n, int i, int j) {
  double a[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const int who[2] = {i, j};
#pragma unroll
  for (int w = 0; w < 2; ++w)
    for (int l = 0; l < n; ++l) {
      if (l == who[w]) continue;
      const double r0 = b.x[3 * who[w]] - b.x[3 * l], r1 = b.x[3 * who[w] + 1] - b.x[3 * l + 1], r2 = b.x[3 * who[w] + 2] - b.x[3 * l + 2];
      const double d2 = r0 * r0 + r1 * r1 + r2 * r2;
      const double f = kG * b.m[l] / (d2 * sqrt(d2));
      a[w][0] -= f * r0;
      a[w][1] -= f * r1;
    }
  const double dx = b.x[3 * j] - b.x[3 * i], dy = b.x[3 * j + 1] - b.x[3 * i + 1];
  const double dvx = b.v[3 * j] - b.v[3 * i], dvy = b.v[3 * j + 1] - b.v[3 * i + 1];
  return dvx * dvx + dvy * dvy + dx * (a[1][0] - a[0][0]) + dy * (a[1][1] - a[0][1]);
}

}  // namespace nbg
