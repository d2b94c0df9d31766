// Jacobian propagation: applies the per-step operator stream written by the trajectory kernel to the
// 7N x 7N matrix jac_step of each system, one thread per COLUMN, with the reference's Kahan-compensated
// accumulation (jac_step + jac_error).  Generic-N version: the matrix of one system lives in shared memory
// for the whole chunk of steps.
//
// Replaces, for this path, the Jacobian half of ahl21!(s,d::Derivatives,h) (src/integrator/ahl21/ahl21.jl:5-95):
//   drift_grad!                         ahl21.jl:318-331      J[x rows] (+)= h2 J[v rows]
//   comp_sum_matrix!(jac_step,...)      ahl21.jl:23,54,93     (with jac_kick == 0 these are pure Kahan folds)
//   copy_submatrix! / mul!(jac_ij, .) / comp_sum_matrix! / ypoc_submatrix!   ahl21.jl:31-35, 64-68; utils.jl:36-100
//   mul!(jac_copy, jac_phi, jac_step)   ahl21.jl:48           (factored form, see nbg_step.cuh)
// and dtbvdq! (src/transits/timing.jl:155-194) for the transit-time gradient.
//
// Every update is a LEFT multiplication, so column c never needs another column: no inter-thread traffic
// beyond broadcasting the operator records.  Rows 7i+6 (masses) stay unit rows forever (rows 7/14 of jac_ij
// are zero, ahl21.jl:735-750; jac_phi and the drift never touch them), so only 6N rows are stored and a mass
// COLUMN of an operator contributes only to the thread that owns column 7p+6.
#pragma once
#include "nbg_kicks.cuh"

namespace nbg {

struct JacSmem {
  double* Jv;   // [6N][M]
  double* Je;   // [6N][M]
  double* da;   // [3N][M] scratch for phisalpha
  double* phi;  // [P][PF] staged phisalpha records (may be null -> read from global)
  double* rec;  // [2][KF] staged Kepler record (double buffered)
};

// operator-stream reader: field f of this system/slot at ((f/4)*stride + idx)*4 + f%4  (see Emit)
struct Src {
  const double* base;
  size_t stride;
  size_t idx;
  __device__ __forceinline__ double get(size_t f) const { return __ldg(base + ((f >> 2) * stride + idx) * 4 + (f & 3)); }
};

// Kahan fold of every stored entry of this thread's column: comp_sum_matrix! with a zero addend (utils.jl:36-46).
__device__ __forceinline__ void fold_column(const JacSmem& S, int n, int M, int c) {
  for (int row = 0; row < 6 * n; ++row) {
    double val = S.Jv[row * M + c], err = S.Je[row * M + c];
    ksum_m(val, err, 0.0);
    S.Jv[row * M + c] = val;
    S.Je[row * M + c] = err;
  }
}
// drift_grad! Jacobian part (ahl21.jl:326-328): scalar-form Kahan
__device__ __forceinline__ void drift_column(const JacSmem& S, int n, int M, int c, double h2) {
  for (int b = 0; b < n; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int rx = (6 * b + k) * M + c, rv = (6 * b + 3 + k) * M + c;
      double val = S.Jv[rx], err = S.Je[rx];
      ksum(val, err, h2 * S.Jv[rv]);
      S.Jv[rx] = val;
      S.Je[rx] = err;
    }
}

// rows(i) u rows(j) of column c  (+)=  jac_ij * same rows      (ahl21.jl:31-35)
__device__ __forceinline__ void kepler_pair_column(const JacSmem& S, const double* __restrict__ R, int i, int j, int M, int c) {
  double d[6], w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) d[r] = S.Jv[(6 * i + r) * M + c] - S.Jv[(6 * j + r) * M + c];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += R[kf_k(r, k)] * d[k];
    w[r] = s;
  }
  const double mi = R[KF_MI], mj = R[KF_MJ];
  double ai[6], aj[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) { ai[r] = mj * w[r]; aj[r] = -mi * w[r]; }
  if (c == 7 * i + 6) {
#pragma unroll
    for (int r = 0; r < 6; ++r) { ai[r] += R[kf_ci7(r)]; aj[r] += R[kf_cj7(r)]; }
  }
  if (c == 7 * j + 6) {
#pragma unroll
    for (int r = 0; r < 6; ++r) { ai[r] += R[kf_ci14(r)]; aj[r] += R[kf_cj14(r)]; }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int ri = (6 * i + r) * M + c, rj = (6 * j + r) * M + c;
    double val = S.Jv[ri], err = S.Je[ri];
    ksum_m(val, err, ai[r]);
    S.Jv[ri] = val; S.Je[ri] = err;
    val = S.Jv[rj]; err = S.Je[rj];
    ksum_m(val, err, aj[r]);
    S.Jv[rj] = val; S.Je[rj] = err;
  }
}

// jac_step (+)= jac_phi * jac_step for column c, then the Kahan add over the whole column (ahl21.jl:48,54).
// PHI(p, f): field f of phisalpha record p.
template <class PhiGet>
__device__ __forceinline__ void phisalpha_column(const JacSmem& S, PhiGet PHI, int n, int M, int c) {
  // pass 1: da_b for every body
  for (int q = 0; q < 3 * n; ++q) S.da[q * M + c] = 0.0;
  int p = 0;
  for (int i = 0; i < n - 1; ++i) {
    double dai[3] = {0.0, 0.0, 0.0};
    double xi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) xi[k] = S.Jv[(6 * i + k) * M + c];
    for (int j = i + 1; j < n; ++j, ++p) {
      const double r0 = PHI(p, PF_R), r1 = PHI(p, PF_R + 1), r2v = PHI(p, PF_R + 2);
      const double g3 = PHI(p, PF_G3), mi = PHI(p, PF_MI), mj = PHI(p, PF_MJ);
      double w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) w[k] = xi[k] - S.Jv[(6 * j + k) * M + c];
      const double rw = r0 * w[0] + r1 * w[1] + r2v * w[2];
      const double f3 = PHI(p, PF_G5) * rw;
      double gw[3] = {g3 * w[0] - f3 * r0, g3 * w[1] - f3 * r1, g3 * w[2] - f3 * r2v};
      // mass columns: da_i -= gam_ij dm_j ; da_j += gam_ij dm_i     (gam = G r / r^3)
      const double dmj = (c == 7 * j + 6) ? 1.0 : 0.0, dmi = (c == 7 * i + 6) ? 1.0 : 0.0;
      const double ga[3] = {g3 * r0, g3 * r1, g3 * r2v};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dai[k] -= mj * gw[k] + ga[k] * dmj;
        S.da[(3 * j + k) * M + c] += mi * gw[k] + ga[k] * dmi;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) S.da[(3 * i + k) * M + c] += dai[k];
  }
  // pass 2: dv_b = sum over partners; each ordered (b, d) pair evaluated from the record of the unordered pair
  for (int b = 0; b < n; ++b) {
    double dv[3] = {0.0, 0.0, 0.0};
    double xb[3], ab[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { xb[k] = S.Jv[(6 * b + k) * M + c]; ab[k] = S.da[(3 * b + k) * M + c]; }
    for (int d = 0; d < n; ++d) {
      if (d == b) continue;
      const int i = b < d ? b : d, j = b < d ? d : b;
      const int pp = i * n - i * (i + 1) / 2 + (j - i - 1);
      // orient everything as (i,j): w = dx_i - dx_j, wa = da_i - da_j
      const double sg = (b == i) ? 1.0 : -1.0;
      double w[3], wa[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        w[k] = sg * (xb[k] - S.Jv[(6 * d + k) * M + c]);
        wa[k] = sg * (ab[k] - S.da[(3 * d + k) * M + c]);
      }
      const double r0 = PHI(pp, PF_R), r1 = PHI(pp, PF_R + 1), r2v = PHI(pp, PF_R + 2);
      const double fac1 = PHI(pp, PF_FAC1), r2 = PHI(pp, PF_R2), us = PHI(pp, PF_US);
      const double rwa = r0 * wa[0] + r1 * wa[1] + r2v * wa[2];
      const double dmsum = ((c == 7 * i + 6) ? 1.0 : 0.0) + ((c == 7 * j + 6) ? 1.0 : 0.0);
      const double rr[3] = {r0, r1, r2v};
      // body b receives  +m_j dF (b == i)  or  -m_i dF (b == j), plus F dm of the partner
      const double mpart = (b == i) ? PHI(pp, PF_MJ) : PHI(pp, PF_MI);
      const double dmpart = (c == 7 * d + 6) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double dF = PHI(pp, PF_RM + 3 * k) * w[0] + PHI(pp, PF_RM + 3 * k + 1) * w[1] + PHI(pp, PF_RM + 3 * k + 2) * w[2] +
                    fac1 * (3.0 * rr[k] * rwa - r2 * wa[k]) + us * rr[k] * dmsum;
        dv[k] += sg * (mpart * dF + PHI(pp, PF_F + k) * dmpart);
      }
    }
    // comp_sum_matrix!(jac_step, jac_error, jac_copy): v rows of this body (v rows are never read by jac_phi)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int rv = (6 * b + 3 + k) * M + c;
      double val = S.Jv[rv], err = S.Je[rv];
      ksum_m(val, err, dv[k]);
      S.Jv[rv] = val; S.Je[rv] = err;
    }
  }
  // ... and the x rows, whose addend is zero (pure fold) -- only after every dv has been formed from the unfolded values
  for (int b = 0; b < n; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int rx = (6 * b + k) * M + c;
      double val = S.Jv[rx], err = S.Je[rx];
      ksum_m(val, err, 0.0);
      S.Jv[rx] = val; S.Je[rx] = err;
    }
}

// One full AHL21 Jacobian step on the smem-resident matrix.  All threads of the block call this
// (threads with c >= M only help staging records).  tid/nthr: thread index / count within the system's group.
__device__ __forceinline__ void jac_apply_step(const JacSmem& S, const Src& src, int n, int M, int c, double h2, int tid, int nthr) {
  const int P = npairs(n);
  const bool act = c < M;
  // prefetch record 0 and the phisalpha block while drifting
  double pre[2];
  const int nf = (KF + nthr - 1) / nthr;  // fields per thread (<= 2 for nthr >= 32)
#pragma unroll 2
  for (int q = 0; q < 2; ++q) { int f = tid + q * nthr; pre[q] = (q < nf && f < KF) ? src.get(f) : 0.0; }
  if (S.phi) {
    const size_t pb = (size_t)2 * P * KF;
    for (int f = tid; f < P * PF; f += nthr) S.phi[f] = src.get(pb + f);
  }
  if (act) {
    drift_column(S, n, M, c, h2);
    fold_column(S, n, M, c);
  }
  int buf = 0;
#pragma unroll 2
  for (int q = 0; q < 2; ++q) { int f = tid + q * nthr; if (q < nf && f < KF) S.rec[buf * KF + f] = pre[q]; }
  __syncthreads();
  int rec = 0;
  // ascending sweep
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j, ++rec) {
      const bool more = rec + 1 < 2 * P;
#pragma unroll 2
      for (int q = 0; q < 2; ++q) { int f = tid + q * nthr; pre[q] = (more && q < nf && f < KF) ? src.get((size_t)(rec + 1) * KF + f) : 0.0; }
      if (act) kepler_pair_column(S, S.rec + buf * KF, i, j, M, c);
      buf ^= 1;
#pragma unroll 2
      for (int q = 0; q < 2; ++q) { int f = tid + q * nthr; if (q < nf && f < KF) S.rec[buf * KF + f] = pre[q]; }
      __syncthreads();
    }
  // phisalpha
  if (act) {
    if (S.phi) {
      const double* ph = S.phi;
      phisalpha_column(S, [ph](int p, int f) { return ph[p * PF + f]; }, n, M, c);
    } else {
      const size_t pb = (size_t)2 * P * KF;
      phisalpha_column(S, [&src, pb](int p, int f) { return src.get(pb + (size_t)p * PF + f); }, n, M, c);
    }
  }
  // descending sweep
  for (int i = n - 2; i >= 0; --i)
    for (int j = n - 1; j >= i + 1; --j, ++rec) {
      const bool more = rec + 1 < 2 * P;
#pragma unroll 2
      for (int q = 0; q < 2; ++q) { int f = tid + q * nthr; pre[q] = (more && q < nf && f < KF) ? src.get((size_t)(rec + 1) * KF + f) : 0.0; }
      if (act) kepler_pair_column(S, S.rec + buf * KF, i, j, M, c);
      buf ^= 1;
#pragma unroll 2
      for (int q = 0; q < 2; ++q) { int f = tid + q * nthr; if (q < nf && f < KF) S.rec[buf * KF + f] = pre[q]; }
      __syncthreads();
    }
  if (act) {
    drift_column(S, n, M, c, h2);
    fold_column(S, n, M, c);
  }
}

}  // namespace nbg
