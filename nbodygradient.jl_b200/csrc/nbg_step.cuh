// One AHL21 step of one planetary system by ONE thread (lanes of a warp = different systems):
// positions/velocities with Kahan compensation, dq/dh, and -- when EMIT -- the stream of
// per-pair linear operators that the Jacobian kernel (nbg_jacobian.cuh) applies to the 7N x 7N matrix.
//
// Replaces, for this path,
//   ahl21!(s,d::Derivatives,h)    src/integrator/ahl21/ahl21.jl:5-95   (x, v, dqdt part; Jacobian part is streamed)
//   ahl21!(s,h)                   src/integrator/ahl21/ahl21_no_grad.jl:6-18 (GRAD=false)
//   drift!/drift_grad!            ahl21_no_grad.jl:24-29, ahl21.jl:318-331
//   kepler_driftij_gamma!         ahl21.jl:706-760, ahl21_no_grad.jl:192-212
//   phisalpha!                    ahl21.jl:558-700, ahl21_no_grad.jl:110-158
// kickfast!/phic! act only on pairs flagged in s.pair, which is all-false unless set by hand (Integrator.jl:91); they live in
// nbg_kicks.cuh together with ahl21_step, which composes the pieces of this file.
//
// Why x, v, dqdt never need the Jacobian: every Jacobian update in the reference is a left
// multiplication of jac_step by an operator built from x, v, m only, and dqdt obeys the same recurrence
// as one extra column.  So this thread advances the state and describes the operators; a thread group
// per system applies them (column per thread) in a separate kernel.
#pragma once
#include "nbg_kepler.cuh"

namespace nbg {

constexpr int NMAX = 16;  // max bodies per system
constexpr int KF = 64;    // doubles per Kepler-pair operator record
constexpr int PF = 24;    // doubles per phisalpha-pair operator record
// Fast-kick pairs (s.pair): bit p = pair index i*n - i(i+1)/2 + (j-i-1), i < j; N <= 16 -> 120 pairs in 128 bits.
struct KMask {
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  __host__ __device__ __forceinline__ bool any() const { return (w[0] | w[1] | w[2] | w[3]) != 0u; }
  __host__ __device__ __forceinline__ bool bit(int p) const { return (w[(p >> 5) & 3] >> (p & 31)) & 1u; }
  __host__ __device__ __forceinline__ void set(int p) { w[(p >> 5) & 3] |= 1u << (p & 31); }
};
constexpr int SCF = 12;   // doubles per pair section in the scalar stream of the split path: x0, v0, gamma, k, m_i, m_j (+2 pad); see kepler_scalars

// Kepler record fields.  The 6x6 block jac_kepler is stored as four 3x3 blocks ordered for the Jacobian kernel's
// x-rows / v-rows thread halves: [Kxx, Kxv | Kvv, Kvx], then the mass-column terms split the same way.
__host__ __device__ constexpr int kf_k(int r, int c) {  // jac_kepler[r][c], r,c < 6
  return r < 3 ? (c < 3 ? 3 * r + c : 9 + 3 * r + (c - 3)) : (c >= 3 ? 18 + 3 * (r - 3) + (c - 3) : 27 + 3 * (r - 3) + c);
}
constexpr int KF_MI = 36;    // m_i/(m_i+m_j)
constexpr int KF_MJ = 37;    // m_j/(m_i+m_j)
// mass-column terms of jac_ij (ahl21.jl:743-750), row r < 6 of body i or j, column m_i ("7") or m_j ("14")
__host__ __device__ constexpr int kf_ci7(int r) { return r < 3 ? 38 + r : 50 + (r - 3); }    // rows i, col m_i
__host__ __device__ constexpr int kf_cj7(int r) { return r < 3 ? 41 + r : 53 + (r - 3); }    // rows j, col m_i
__host__ __device__ constexpr int kf_ci14(int r) { return r < 3 ? 44 + r : 56 + (r - 3); }   // rows i, col m_j
__host__ __device__ constexpr int kf_cj14(int r) { return r < 3 ? 47 + r : 59 + (r - 3); }   // rows j, col m_j
// phisalpha record fields, grouped by 4 (one 32-byte sector each) in the order the dense-operator kernel needs them
constexpr int PF_R = 0;      // 3: r_ij
constexpr int PF_G3 = 3;     // G / r^3
constexpr int PF_G5 = 4;     // 3 G / r^5
constexpr int PF_MI = 5;     // m_i
constexpr int PF_MJ = 6;     // m_j
constexpr int PF_FAC1 = 7;   // coeff / r^5
constexpr int PF_R2 = 8;     // r^2
constexpr int PF_US = 9;     // 2 G fac1 / r   (mass-sum derivative is US * r_ij)
constexpr int PF_F = 10;     // 3: F_ij
constexpr int PF_RM = 13;    // 9: dF/dr [k][p]  (index 3*k + p)

__host__ __device__ inline int npairs(int n) { return n * (n - 1) / 2; }
// doubles per system per step in the operator stream: [2P Kepler records | P phisalpha records | dense phisalpha operator]
// dense operator (written by phi_dense_kernel, read by the register-resident Jacobian kernel): element
// ((3 i + k) N + d) 4 + p = d v_i[k] / d x_d[p] for p < 3, d v_i[k] / d m_d for p = 3.
// With fast-kick pairs (s.pair not all-false) a step carries THREE sets of compact records and dense operators: first
// kickfast!, phic!+phisalpha!, second kickfast! (nbg_kicks.cuh).
__host__ __device__ inline size_t phi_dense_fields(int n) { return (size_t)12 * n * n; }
__host__ __device__ inline int phi_sets(bool kicked) { return kicked ? 3 : 1; }
__host__ __device__ inline size_t phi_rec_offset(int n, int set) { return (size_t)npairs(n) * (2 * KF + set * PF); }
__host__ __device__ inline size_t phi_dense_offset(int n, bool kicked = false, int set = 0) {
  return (size_t)npairs(n) * (2 * KF + phi_sets(kicked) * PF) + set * phi_dense_fields(n);
}
__host__ __device__ inline size_t step_fields(int n, bool kicked = false) { return phi_dense_offset(n, kicked, phi_sets(kicked)); }

struct Body {
  double x[3 * NMAX], v[3 * NMAX], xe[3 * NMAX], ve[3 * NMAX], m[NMAX];
};

// Operator-stream addressing.  Systems (or queued transits) are tiled by 32 -- one warp of the trajectory kernel -- and
// the block of one step of one tile is contiguous: [step][tile][group of 4 fields][32 systems][4 doubles].  A warp writes
// 1 KB runs; the Jacobian kernel reads its system's 32-byte sectors 1 KB apart inside one 32*sf*8-byte region (the
// earlier [group][all systems] layout put every sector of a system's step in a different 2 MB page: the Jacobian kernel
// ran 1.36x slower on 65,536 systems than on 16,384).
constexpr int TILE = 32;
__host__ __device__ inline size_t tile_offset(size_t sf, size_t ntiles, size_t step, size_t item) {
  return (step * ntiles + item / TILE) * sf * TILE;  // in doubles, from the start of the stream
}

// Operator-stream writer for this thread's system/slot (idx) out of `stride` systems/slots.
struct Emit {
  double* base;
  size_t stride;
  size_t idx;
  double* sbase = nullptr;  // scalar records of the split path (EMIT == 2): SCF doubles per pair section, same tiling
  // fields are packed in groups of 4 per system (32-byte sectors): element f at ((f/4)*stride + idx)*4 + f%4
  __device__ __forceinline__ void put(size_t f, double val) const { base[((f >> 2) * stride + idx) * 4 + (f & 3)] = val; }
  // a whole record of NF (multiple of 4) doubles starting at field f0 (multiple of 4), as 16-byte stores
  template <int NF> __device__ __forceinline__ void put_record(size_t f0, const double (&rec)[NF], bool scalars = false) const {
#pragma unroll
    for (int g = 0; g < NF / 4; ++g) {
      double2* dst = reinterpret_cast<double2*>((scalars ? sbase : base) + (((f0 >> 2) + g) * stride + idx) * 4);
      dst[0] = make_double2(rec[4 * g], rec[4 * g + 1]);
      dst[1] = make_double2(rec[4 * g + 2], rec[4 * g + 3]);
    }
  }
};

// One body's state held in registers while it takes part in a run of pair updates.  Body arrays live in thread-local
// memory with run-time indices; ncu (profiles/r01_t2_traj) showed >50% of the trajectory kernel's stall samples waiting
// on those dependent local loads at the top of every pair.  The sweeps therefore keep body i in registers across its
// whole j loop and fetch body j+1 while pair (i, j) is being solved.
struct BodyRegs {
  double x[3], v[3], xe[3], ve[3], m, dq[6];
};
template <bool GRAD> __device__ __forceinline__ void load_body(const Body& b, const double* dq, int i, BodyRegs& r) {
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.x[k] = b.x[3 * i + k]; r.v[k] = b.v[3 * i + k]; r.xe[k] = b.xe[3 * i + k]; r.ve[k] = b.ve[3 * i + k]; }
  r.m = b.m[i];
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < 6; ++k) r.dq[k] = dq[6 * i + k];
  }
}
template <bool GRAD> __device__ __forceinline__ void store_body(Body& b, double* dq, int i, const BodyRegs& r) {
#pragma unroll
  for (int k = 0; k < 3; ++k) { b.x[3 * i + k] = r.x[k]; b.v[3 * i + k] = r.v[k]; b.xe[3 * i + k] = r.xe[k]; b.ve[3 * i + k] = r.ve[k]; }
  if (GRAD) {
#pragma unroll
    for (int k = 0; k < 6; ++k) dq[6 * i + k] = r.dq[k];
  }
}

// Packs / unpacks what pair_op_kernel needs to rebuild a pair section's operator record (split path, trajectory kernel -> pair_op_kernel).
__device__ __forceinline__ void scal_pack(double (&rec)[SCF], const double* x0, const double* v0, const KepScal& P, double mi, double mj) {
  rec[0] = x0[0]; rec[1] = x0[1]; rec[2] = x0[2]; rec[3] = v0[0]; rec[4] = v0[1]; rec[5] = v0[2];
  rec[6] = P.gamma; rec[7] = P.k; rec[8] = mi; rec[9] = mj; rec[10] = 0.0; rec[11] = 0.0;
}
__device__ __forceinline__ void scal_unpack(const double (&rec)[SCF], double* x0, double* v0, double& gamma, double& k, double& mi, double& mj) {
  x0[0] = rec[0]; x0[1] = rec[1]; x0[2] = rec[2]; v0[0] = rec[3]; v0[1] = rec[4]; v0[2] = rec[5];
  gamma = rec[6]; k = rec[7]; mi = rec[8]; mj = rec[9];
}
// The Kepler operator record of one pair section from the pair's Jacobian (jac_ij without its identity, ahl21.jl:735-758):
// the 6x6 block on relative coordinates, the mass fractions, and the four rank-one mass columns.
__device__ __forceinline__ void kepler_record(double (&rec)[KF], const KepJac& J, const double* dl, double bim, double bjm) {
  const double mijinv = 1.0 / (bim + bjm);
  const double mi = bim * mijinv, mj = bjm * mijinv;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) rec[kf_k(r, c)] = J.jk[r][c];
  rec[KF_MI] = mi;
  rec[KF_MJ] = mj;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    rec[kf_ci7(r)] = J.jm[r] * bjm;
    rec[kf_cj7(r)] = -mj * dl[r] * mijinv - kG * mi * J.jk[r][6];
    rec[kf_ci14(r)] = mi * dl[r] * mijinv + kG * mj * J.jk[r][6];
    rec[kf_cj14(r)] = -J.jm[r] * bim;
  }
  rec[62] = 0.0; rec[63] = 0.0;
}

// EMIT: 0 = nothing, 1 = full operator records (needs GRAD), 2 = split path: only the inputs of kepler_jacobian are
// written (SCF doubles per section) and pair_op_kernel builds the records; GRAD = false then (dq/dh restarts from zero at
// every step, ahl21.jl:9, so only the LAST step of an integration needs it: that one runs with GRAD = true, EMIT = 1).
template <bool GRAD, int EMIT>
__device__ __forceinline__ void pair_section(BodyRegs& bi, BodyRegs& bj, double h2, bool drift_first, const Emit& em, size_t rec_base) {
  double x0[3], v0[3], dl[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) { x0[k] = bi.x[k] - bj.x[k]; v0[k] = bi.v[k] - bj.v[k]; }
  const double msum = bi.m + bj.m;
  const double gm = kG * msum;
  KepJac J;
  if (gm == 0.0) {
    // Two massless bodies: no interaction.  (The reference returns before touching anything, ahl21.jl:713,
    // and then re-applies the previous pair's stale jac_ij; here the pair is the identity.)
    if (EMIT == 1) {
      double rec[KF];
#pragma unroll
      for (int f = 0; f < KF; ++f) rec[f] = 0.0;
      em.put_record<KF>(rec_base, rec);
    }
    if (EMIT == 2) {
      double rec[SCF];
#pragma unroll
      for (int f = 0; f < SCF; ++f) rec[f] = 0.0;   // k = 0 marks "no interaction"
      em.put_record<SCF>(rec_base / KF * SCF, rec, true);
    }
    return;
  }
  KepScal P;
  kepler_solve(x0, v0, gm, h2, drift_first, dl, &P);
  if (GRAD) kepler_jacobian(&P, x0, v0, drift_first, &J);
  const double mijinv = 1.0 / msum;
  const double mi = bi.m * mijinv, mj = bj.m * mijinv;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ksum(bi.x[k], bi.xe[k], mj * dl[k]);
    ksum(bj.x[k], bj.xe[k], -mi * dl[k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ksum(bi.v[k], bi.ve[k], mj * dl[3 + k]);
    ksum(bj.v[k], bj.ve[k], -mi * dl[3 + k]);
  }
  if (GRAD) {
    // dqdt_ij = dqdt_ij/2 + dqdt_old + jac_ij * dqdt_old   (ahl21.jl:38-40); mass entries of dqdt are identically 0
    double dd[6], w[6];
#pragma unroll
    for (int l = 0; l < 6; ++l) dd[l] = bi.dq[l] - bj.dq[l];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < 6; ++l) s += J.jk[r][l] * dd[l];
      w[r] = s;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      bi.dq[r] = 0.5 * (mj * J.jk[r][7]) + bi.dq[r] + mj * w[r];
      bj.dq[r] = 0.5 * (-mi * J.jk[r][7]) + bj.dq[r] - mi * w[r];
    }
  }
  if (EMIT == 1) {
    double rec[KF];
    kepler_record(rec, J, dl, bi.m, bj.m);
    em.put_record<KF>(rec_base, rec);
  }
  if (EMIT == 2) {
    double rec[SCF];
    scal_pack(rec, x0, v0, P, bi.m, bj.m);
    em.put_record<SCF>(rec_base / KF * SCF, rec, true);
  }
}

// phisalpha!(s,d,h,alpha=2): velocity kick m_j F_ij / -m_i F_ij per pair, with
//   F_ij = fac1 (r fac2 - r^2 a_ij),  fac1 = coeff / r^5,  fac2 = 2 G (m_i+m_j)/r + 3 a_ij.r,  a_i = -sum_d G m_d r_id / r_id^3.
// Its Jacobian is applied in factored form (same linear operator as the reference's dense jac_phi):
//   da_i = - sum_d m_d Gam_id (dx_i - dx_d) - sum_d gam_id dm_d,      Gam = G (I/r^3 - 3 r r^T / r^5),  gam = G r / r^3
//   dF_ij = Rm_ij (dx_i - dx_j) + fac1 (3 r r^T - r^2 I) (da_i - da_j) + US r (dm_i + dm_j)
//   dv_i += m_j dF_ij + F_ij dm_j ;  dv_j -= m_i dF_ij + F_ij dm_i.
// The pieces of phisalpha! that determine x, v are non-template and never inlined, so the grad and no-grad
// paths run identical instructions (x, v bit-identical between them, as the reference asserts).
__device__ __noinline__ void phis_accel(const Body& b, int n, double* __restrict__ a) {
  for (int q = 0; q < 3 * n; ++q) a[q] = 0.0;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j) {
      double r[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = b.x[3 * i + k] - b.x[3 * j + k];
      const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      const double r3 = r2 * sqrt(r2);
      const double fac2 = kG / r3;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double fac = fac2 * r[k];
        a[3 * i + k] -= b.m[j] * fac;
        a[3 * j + k] += b.m[i] * fac;
      }
    }
}
// out = {F0, F1, F2, fac1, fac2, r2, r1}
__device__ __noinline__ void phis_force(const double* __restrict__ r, const double* __restrict__ aij, double gmu, double coeff,
                                        double* __restrict__ out) {
  const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double r1 = sqrt(r2);
  const double ardot = aij[0] * r[0] + aij[1] * r[1] + aij[2] * r[2];
  const double fac1 = coeff / (r2 * r2 * r1);
  const double fac2 = 2.0 * gmu / r1 + 3.0 * ardot;
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = fac1 * (r[k] * fac2 - r2 * aij[k]);
  out[3] = fac1; out[4] = fac2; out[5] = r2; out[6] = r1;
}

template <bool GRAD, int EMIT>
__device__ __forceinline__ void phisalpha_section(Body& b, double* dq, int n, double h, const Emit& em, size_t rec_base) {
  double a[3 * NMAX], da[3 * NMAX];
  const double coeff = 2.0 * (h * h * h) / 96.0 * 2.0 * kG;  // alpha = 2  (ahl21.jl:564)
  phis_accel(b, n, a);
  if (GRAD) {
    for (int q = 0; q < 3 * n; ++q) da[q] = 0.0;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double r[3], w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { r[k] = b.x[3 * i + k] - b.x[3 * j + k]; w[k] = dq[6 * i + k] - dq[6 * j + k]; }
        const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        const double fac2 = kG / (r2 * sqrt(r2));
        const double rw = r[0] * w[0] + r[1] * w[1] + r[2] * w[2];
        const double f3 = 3.0 * fac2 / r2 * rw;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double gw = fac2 * w[k] - f3 * r[k];
          da[3 * i + k] -= b.m[j] * gw;
          da[3 * j + k] += b.m[i] * gw;
        }
      }
  }
  double dvacc[3 * NMAX];
  if (GRAD) for (int q = 0; q < 3 * n; ++q) dvacc[q] = 0.0;
  int p = 0;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j, ++p) {
      double r[3], aij[3], fo[7];
#pragma unroll
      for (int k = 0; k < 3; ++k) { aij[k] = a[3 * i + k] - a[3 * j + k]; r[k] = b.x[3 * i + k] - b.x[3 * j + k]; }
      const double gmu = kG * (b.m[i] + b.m[j]);
      phis_force(r, aij, gmu, coeff, fo);
      const double F[3] = {fo[0], fo[1], fo[2]};
      const double fac1 = fo[3], fac2 = fo[4], r2 = fo[5], r1 = fo[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
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
          double w[3], wa[3], dF[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) { w[k] = dq[6 * i + k] - dq[6 * j + k]; wa[k] = da[3 * i + k] - da[3 * j + k]; }
          const double rwa = r[0] * wa[0] + r[1] * wa[1] + r[2] * wa[2];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            dF[k] = Rm[k][0] * w[0] + Rm[k][1] * w[1] + Rm[k][2] * w[2] + fac1 * (3.0 * r[k] * rwa - r2 * wa[k]);
            // dqdt_phi (explicit h-dependence, coeff ~ h^3) + jac_phi * dqdt   (ahl21.jl:632-633, :51-52)
            dvacc[3 * i + k] += b.m[j] * (3.0 / h * F[k] + dF[k]);
            dvacc[3 * j + k] -= b.m[i] * (3.0 / h * F[k] + dF[k]);
          }
        }
        if (EMIT) {
          double rec[PF];
#pragma unroll
          for (int k = 0; k < 3; ++k) { rec[PF_R + k] = r[k]; rec[PF_F + k] = F[k]; }
          const double g3 = kG / (r2 * r1);
          rec[PF_G3] = g3;
          rec[PF_FAC1] = fac1;
          rec[PF_R2] = r2;
          rec[PF_US] = 2.0 * kG * fac1 / r1;
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < 3; ++q) rec[PF_RM + 3 * k + q] = Rm[k][q];
          rec[PF_MI] = b.m[i];
          rec[PF_MJ] = b.m[j];
          rec[PF_G5] = 3.0 * g3 * r2inv;
          rec[22] = 0.0; rec[23] = 0.0;
          em.put_record<PF>(rec_base + (size_t)p * PF, rec);
        }
      }
    }
  if (GRAD) for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) dq[6 * i + 3 + k] += dvacc[3 * i + k];
}

}  // namespace nbg
