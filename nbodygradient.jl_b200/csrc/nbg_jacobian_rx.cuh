// Register-resident Jacobian propagation for N <= NBG_RX_MAX_BODIES = 16 bodies (the production path; nbg_jacobian.cuh is the generic
// shared-memory version used for N = 15, 16, where 12 N doubles of resident state per thread no longer fit in 255 registers
// and a double-buffered operator block no longer fits in 227 KB of shared memory).
//
// Why this shape (measured on B200, profiles/microbench/r01_smem_fp64_probe.txt): per SM and per cycle the machine
// issues 2 warp-DFMA, delivers ONE broadcast double from shared memory (LDS.64/.128 uniform or two addresses split by
// half-warp), half a distinct LDS.64, and 1 SHFL.  A column-per-thread update with jac_step in shared memory needs
// ~100 LDS/STS per pair per warp and is bound by that pipe (v1: 54% smem wavefronts, 21% FP64 pipe).  Here jac_step
// and jac_error never leave registers:
//   * each COLUMN of jac_step is owned by two lanes 16 apart in a warp: lane half 0 holds the x rows of every body,
//     half 1 the v rows (3N values + 3N Kahan errors each).  A warp covers 16 columns; a system uses ceil(7N/16) warps.
//   * a pair update needs d = J_i - J_j (x part and v part): each half forms its 3 values and swaps them with its
//     partner by one SHFL.xor 16 per double; then w = A d_mine + B d_other with (A,B) = (Kxx,Kxv) for the x half and
//     (Kvv,Kvx) for the v half.  The record stores the four 3x3 blocks in exactly that order, so a half-warp reads 18
//     contiguous doubles with LDS.128 (two addresses by half-warp = same cost as a uniform broadcast).
//   * per pair and thread: 3 DADD + 18 DFMA + 6 DMUL + 24 Kahan DADD, 10 LDS.128, 6 SHFL.32 -> FP64 pipe ~ LSU pipe.
//   * the whole operator block of one step (34 KB at N = 8) is fetched with cp.async into a double-buffered shared
//     memory ring, one __syncthreads per STEP (not per pair); pair indices are compile-time (fully unrolled), so all
//     register indexing is static.
// Replaces the same reference code as nbg_jacobian.cuh (ahl21.jl:5-95 Jacobian half, timing.jl:155-194).
#pragma once
#include <cuda_pipeline.h>
#include <type_traits>
#include "nbg_jacobian.cuh"

namespace nbg {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NBG_RX_MAX_BODIES = 16;  // jac_step + jac_error in registers; the operator block in shared memory is double-buffered for N <= 14, single for 15, 16
__host__ __device__ constexpr int rx_warps(int n) { return (7 * n + 15) / 16; }

__device__ __forceinline__ double shx(double v) { return __shfl_xor_sync(FULL, v, 16); }

// compile-time loop: f(std::integral_constant<int, B>{}), ..., f(std::integral_constant<int, E-1>{})
template <int B, int E, class F> __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

template <int N> struct RxState {
  double jv[N][3];
  double je[N][3];
};

// async copy of one step's operator block (4-packed stream layout) into shared memory: the first g0 groups (Kepler
// records) and g1 groups starting at group `skip` (dense phisalpha operator), packed back to back
__device__ __forceinline__ void rx_fetch(double* dst, const double* base, size_t stride, size_t idx, int g0, int skip, int g1, int tid, int nthr) {
  for (int g = tid; g < g0 + g1; g += nthr) {
    const int gg = g < g0 ? g : g - g0 + skip;
    const double* src = base + ((size_t)gg * stride + idx) * 4;
    __pipeline_memcpy_async(dst + 4 * g, src, 16);
    __pipeline_memcpy_async(dst + 4 * g + 2, src + 2, 16);
  }
  __pipeline_commit();
}

// ---- compile-time POSITIONS, run-time BODIES ------------------------------------------------------------------
// Fully unrolling the 2 x N(N-1)/2 pair updates makes ~12k instructions of straight-line code per step and the kernel
// stalls on instruction fetch (ncu r01_rx2: no_instruction 2.6 of 4.2 stall cycles per issue).  Instead the register
// arrays are indexed by POSITION and rotated by one position between groups of pairs, so that the body being paired
// with all later bodies always sits at position 0: the code holds N-1 pair bodies per sweep instead of N(N-1)/2, and
// the group loop is a real loop.  A rotation is 3N (or 6N) register moves, issued in the shadow of the FP64 pipe.
template <int N, int K> __device__ __forceinline__ void rot_left_by(double (&a)[N][3]) {  // position p <- position p + K (mod N)
  constexpr int KK = ((K % N) + N) % N;
  if constexpr (KK != 0) {
    double t[N][3];
#pragma unroll
    for (int p = 0; p < N; ++p) { t[p][0] = a[(p + KK) % N][0]; t[p][1] = a[(p + KK) % N][1]; t[p][2] = a[(p + KK) % N][2]; }
#pragma unroll
    for (int p = 0; p < N; ++p) { a[p][0] = t[p][0]; a[p][1] = t[p][1]; a[p][2] = t[p][2]; }
  }
}

// pair update between positions PA and PB; ci/cj = mass columns (7*body+6) of the bodies at those positions
template <int N, int PA, int PB>
__device__ __forceinline__ void rx_pair(RxState<N>& S, const double* __restrict__ R, int half, int c, int ci, int cj) {
  double md[3], od[3], w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) md[k] = S.jv[PA][k] - S.jv[PB][k];
#pragma unroll
  for (int k = 0; k < 3; ++k) od[k] = shx(md[k]);
  const double* __restrict__ Kb = R + 18 * half;  // [A (3x3) | B (3x3)] of this half
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // one 6-term chain, own-half terms first (they do not wait for the shuffle): 3 FP64 instructions fewer per pair than two
    // 3-term chains plus an add, and measured 4% faster (the three k chains and the Kahan updates supply the parallelism)
    double s = Kb[3 * k] * md[0];
    s = fma(Kb[3 * k + 1], md[1], s);
    s = fma(Kb[3 * k + 2], md[2], s);
    s = fma(Kb[9 + 3 * k], od[0], s);
    s = fma(Kb[9 + 3 * k + 1], od[1], s);
    w[k] = fma(Kb[9 + 3 * k + 2], od[2], s);
  }
  const double2 mm = *reinterpret_cast<const double2*>(R + KF_MI);
  // comp_sum_matrix! (utils.jl:36-46) with the scaling by the mass fractions folded into its first addition:
  // err + m w is formed by one FMA (one rounding instead of two)
  double ei[3], ej[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { ei[k] = fma(mm.y, w[k], S.je[PA][k]); ej[k] = fma(-mm.x, w[k], S.je[PB][k]); }
  if (c == ci || c == cj) {  // mass columns of the two bodies: rank-one terms of jac_ij (ahl21.jl:743-750)
    const double* mb = R + 38 + 12 * half + (c == ci ? 0 : 6);
#pragma unroll
    for (int k = 0; k < 3; ++k) { ei[k] += mb[k]; ej[k] += mb[3 + k]; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double ti = __dadd_rn(S.jv[PA][k], ei[k]);
    S.je[PA][k] = __dadd_rn(ei[k], __dsub_rn(S.jv[PA][k], ti));
    S.jv[PA][k] = ti;
    const double tj = __dadd_rn(S.jv[PB][k], ej[k]);
    S.je[PB][k] = __dadd_rn(ej[k], __dsub_rn(S.jv[PB][k], tj));
    S.jv[PB][k] = tj;
  }
}

template <int N> __device__ __forceinline__ void rx_drift(RxState<N>& S, double h2, int half) {
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double ov = shx(S.jv[b][k]);  // x half receives the v row
      if (half == 0) ksum(S.jv[b][k], S.je[b][k], h2 * ov);
    }
}
template <int N> __device__ __forceinline__ void rx_fold(RxState<N>& S) {
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) ksum_m(S.jv[b][k], S.je[b][k], 0.0);
}

// ---- phisalpha as a DENSE operator -----------------------------------------------------------------------------
// jac_phi (ahl21.jl:635-693) is nonzero only in its v rows x {x, m} columns.  Applying it pair by pair in factored form
// costs ~85 FP64 instructions per pair and thread (r01 profile: 45% of the kernel's FP64 issue for 5% of the canonical
// flops); as a dense 3N x 3N block it is 9N^2/2 DFMA per thread.  The trajectory kernel streams compact per-pair records
// (PF doubles); phi_dense_kernel (one thread per system, step and body i; lanes = systems) expands them into the dense
// blocks between the trajectory and the Jacobian kernel:
//   a_i = - sum_d G m_d r_id / r^3,   T_p = G (I/r^3 - 3 r r^T / r^5),   S_p = fac1 (3 r r^T - r^2 I)      (p = pair)
//   Ax[i][d] = m_d T_id (d != i),  Ax[i][i] = - sum_l m_l T_il;   Am[i][d] = - G r_id / r^3 (d != i),  Am[i][i] = 0
//   Phi_x[i][d] = sum_{j != i} m_j S_ij (Ax[i][d] - Ax[j][d])  +  (d == i ? sum_j m_j Rm_ij : - m_d Rm_id)
//   Phi_m[i][d] = sum_{j != i} m_j S_ij (Am[i][d] - Am[j][d])  +  (d == i ? sum_j m_j us_ij r_ij : m_d us_id r_id + F_id)
// with r_ij, F_ij oriented from body i (sign flips for i > j); same linear operator as the reference's jac_phi.
__device__ __forceinline__ int rx_pair_index(int n, int a, int b) { return a * n - a * (a + 1) / 2 + (b - a - 1); }  // a < b

// blk: this step's (or this queued transit's) operator block; stride/idx as in Emit/Src; i: body whose v rows are formed
template <int N> __device__ __forceinline__ void phi_dense_rows(double* __restrict__ blk, size_t stride, size_t idx, int i) {
  constexpr int P = N * (N - 1) / 2;
  const double* __restrict__ PH = blk + ((size_t)(2 * P * KF / 4) * stride + idx) * 4;  // group 0 of record 0
  double* __restrict__ OUT = blk + ((size_t)(P * (2 * KF + PF) / 4) * stride + idx) * 4;
  const size_t gs = stride * 4;  // doubles between consecutive groups of one system
  auto grp = [&](int p, int g) { return reinterpret_cast<const double2*>(PH + (size_t)(p * (PF / 4) + g) * gs); };
  // T (symmetric 3x3 as xx xy xz yy yz zz) and gam = G (x_a - x_b) / r^3 of the pair {a, b}, oriented from a
  struct TG { double T[6], gam[3], ma, mb; };
  auto load_tg = [&](int a, int b) {
    const int p = rx_pair_index(N, a < b ? a : b, a < b ? b : a);
    const double2 u = __ldg(grp(p, 0)), v = __ldg(grp(p, 0) + 1), e = __ldg(grp(p, 1)), f = __ldg(grp(p, 1) + 1);
    const double r0 = u.x, r1 = u.y, r2 = v.x, g3 = v.y, g5 = e.x, mi = e.y, mj = f.x;
    const double sg = a < b ? 1.0 : -1.0;
    TG o;
    o.T[0] = g3 - g5 * r0 * r0; o.T[1] = -g5 * r0 * r1; o.T[2] = -g5 * r0 * r2;
    o.T[3] = g3 - g5 * r1 * r1; o.T[4] = -g5 * r1 * r2; o.T[5] = g3 - g5 * r2 * r2;
    o.gam[0] = sg * g3 * r0; o.gam[1] = sg * g3 * r1; o.gam[2] = sg * g3 * r2;
    o.ma = a < b ? mi : mj; o.mb = a < b ? mj : mi;
    return o;
  };
#pragma unroll 1
  for (int d = 0; d < N; ++d) {
    // Ax[d][d] = - sum_l m_l T_dl; m_d
    double diag[6] = {0, 0, 0, 0, 0, 0}, md = 0.0;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      if (l == d) continue;
      const TG t = load_tg(d, l);
      md = t.ma;
#pragma unroll
      for (int q = 0; q < 6; ++q) diag[q] = fma(-t.mb, t.T[q], diag[q]);
    }
    // A[b][d]: x part (6) and mass part (3)
    auto Aget = [&](int b, double (&ax)[6], double (&am)[3]) {
      if (b == d) {
#pragma unroll
        for (int q = 0; q < 6; ++q) ax[q] = diag[q];
        am[0] = 0.0; am[1] = 0.0; am[2] = 0.0;
      } else {
        const TG t = load_tg(d, b);
#pragma unroll
        for (int q = 0; q < 6; ++q) ax[q] = md * t.T[q];
        am[0] = t.gam[0]; am[1] = t.gam[1]; am[2] = t.gam[2];
      }
    };
    double axi[6], ami[3];
    Aget(i, axi, ami);
    double acc[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[k][q] = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const int p = rx_pair_index(N, i < j ? i : j, i < j ? j : i);
      const double2 a = __ldg(grp(p, 0)), b = __ldg(grp(p, 0) + 1), e = __ldg(grp(p, 1)), f = __ldg(grp(p, 1) + 1), g = __ldg(grp(p, 2));
      const double r[3] = {a.x, a.y, b.x};
      const double mj = i < j ? f.x : e.y, fac1 = f.y, rsq = g.x, us = g.y;
      const double sg = i < j ? 1.0 : -1.0;
      double axj[6], amj[3];
      Aget(j, axj, amj);
      double u[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) u[q] = axi[q] - axj[q];
      // symmetric storage [0 1 2; 1 3 4; 2 4 5]: column q of the difference; column 3 = mass part
      const double col[4][3] = {{u[0], u[1], u[2]}, {u[1], u[3], u[4]}, {u[2], u[4], u[5]}, {ami[0] - amj[0], ami[1] - amj[1], ami[2] - amj[2]}};
      double ev[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double ru = 3.0 * (r[0] * col[q][0] + r[1] * col[q][1] + r[2] * col[q][2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ev[q][k] = fac1 * (r[k] * ru - rsq * col[q][k]);
      }
      if (d == i || d == j) {
        const double2 g1 = __ldg(grp(p, 2) + 1), h0 = __ldg(grp(p, 3)), h1 = __ldg(grp(p, 3) + 1), k0 = __ldg(grp(p, 4)), k1 = __ldg(grp(p, 4) + 1),
                      l0 = __ldg(grp(p, 5));
        const double F[3] = {g1.x, g1.y, h0.x};
        const double Rm[9] = {h0.y, h1.x, h1.y, k0.x, k0.y, k1.x, k1.y, l0.x, l0.y};
        const double sr = d == i ? 1.0 : -1.0;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int k = 0; k < 3; ++k) ev[q][k] = fma(sr, Rm[3 * k + q], ev[q][k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ev[3][k] = fma(sg * us, r[k], ev[3][k]);
        if (d == j) {
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[k][3] = fma(sg, F[k], acc[k][3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k][q] = fma(mj, ev[q][k], acc[k][q]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double2* o = reinterpret_cast<double2*>(OUT + (size_t)((3 * i + k) * N + d) * gs);
      o[0] = make_double2(acc[k][0], acc[k][1]);
      o[1] = make_double2(acc[k][2], acc[k][3]);
    }
  }
}

// phi_dense_rows with the per-pair tensors cached in shared memory.  In phi_dense_rows a thread rebuilds T and gam of a pair from the
// record 15 times per column block d (120 times per step: ~40 % of its FP64 instructions and most of its L1 traffic, 97 KB per
// system-step for 5.4 KB of records).  Here the block (32 systems x N bodies) first builds them once per pair, then the per-body sums
// A_dd = - sum_l m_l T_dl, and the row loop reads both from shared memory (lane = system: conflict-free).  Same arithmetic in the same
// order as phi_dense_rows: results are bit-identical.
//   tg[(p * TGF + q) * 32 + lane]: q < 6: T, 6..8: gam oriented from the lower-index body, 9: its mass, 10: the other mass
//   dg[(d * 7 + q) * 32 + lane]:   q < 6: A_dd, 6: m_d
// FULL: fields 11..16 also cache r_ij (3), fac1, r^2 and us of the pair, so that the row loop reads nothing but shared memory except
// the F / dF/dr terms of the 2 column blocks d = i, j (N <= 10: 214 KB at N = 10)
__host__ __device__ constexpr int phi_tgf(bool full) { return full ? 17 : 11; }
__host__ __device__ constexpr size_t phi_cache_bytes(int n, bool full) { return (size_t)(n * (n - 1) / 2 * phi_tgf(full) + n * 7) * 32 * 8; }

template <int N, bool FULL>
__device__ __forceinline__ void phi_dense_rows_cached(double* __restrict__ blk, size_t stride, size_t idx, int i, bool live, double* __restrict__ tg,
                                                       double* __restrict__ dg, int lane) {
  constexpr int P = N * (N - 1) / 2, TGF = phi_tgf(FULL);
  const double* __restrict__ PH = blk + ((size_t)(2 * P * KF / 4) * stride + idx) * 4;  // group 0 of record 0
  double* __restrict__ OUT = blk + ((size_t)(P * (2 * KF + PF) / 4) * stride + idx) * 4;
  const size_t gs = stride * 4;  // doubles between consecutive groups of one system
  auto grp = [&](int p, int g) { return reinterpret_cast<const double2*>(PH + (size_t)(p * (PF / 4) + g) * gs); };
  // phase 1: T, gam, masses of every pair, once
  for (int p = i; p < P; p += N) {
    const double2 u = __ldg(grp(p, 0)), v = __ldg(grp(p, 0) + 1), e = __ldg(grp(p, 1)), f = __ldg(grp(p, 1) + 1);
    const double r0 = u.x, r1 = u.y, r2 = v.x, g3 = v.y, g5 = e.x;
    double* o = tg + (size_t)p * TGF * 32 + lane;
    o[0 * 32] = g3 - g5 * r0 * r0; o[1 * 32] = -g5 * r0 * r1; o[2 * 32] = -g5 * r0 * r2;
    o[3 * 32] = g3 - g5 * r1 * r1; o[4 * 32] = -g5 * r1 * r2; o[5 * 32] = g3 - g5 * r2 * r2;
    o[6 * 32] = g3 * r0; o[7 * 32] = g3 * r1; o[8 * 32] = g3 * r2;
    o[9 * 32] = e.y; o[10 * 32] = f.x;
    if (FULL) {
      const double2 g = __ldg(grp(p, 2));
      o[11 * 32] = r0; o[12 * 32] = r1; o[13 * 32] = r2; o[14 * 32] = f.y; o[15 * 32] = g.x; o[16 * 32] = g.y;
    }
  }
  __syncthreads();
  {  // A_dd and m_d of body d = i
    const int d = i;
    double diag[6] = {0, 0, 0, 0, 0, 0}, md = 0.0;
#pragma unroll
    for (int l = 0; l < N; ++l) {
      if (l == d) continue;
      const double* t = tg + (size_t)rx_pair_index(N, d < l ? d : l, d < l ? l : d) * TGF * 32 + lane;
      md = t[(d < l ? 9 : 10) * 32];
      const double ml = t[(d < l ? 10 : 9) * 32];
#pragma unroll
      for (int q = 0; q < 6; ++q) diag[q] = fma(-ml, t[q * 32], diag[q]);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) dg[(d * 7 + q) * 32 + lane] = diag[q];
    dg[(d * 7 + 6) * 32 + lane] = md;
  }
  __syncthreads();
#pragma unroll 1
  for (int d = 0; d < N; ++d) {
    double diag[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) diag[q] = dg[(d * 7 + q) * 32 + lane];
    const double md = dg[(d * 7 + 6) * 32 + lane];
    // A[b][d]: x part (6) and mass part (3)
    auto Aget = [&](int b, double (&ax)[6], double (&am)[3]) {
      if (b == d) {
#pragma unroll
        for (int q = 0; q < 6; ++q) ax[q] = diag[q];
        am[0] = 0.0; am[1] = 0.0; am[2] = 0.0;
      } else {
        const double* t = tg + (size_t)rx_pair_index(N, d < b ? d : b, d < b ? b : d) * TGF * 32 + lane;
        const double sg = d < b ? 1.0 : -1.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) ax[q] = md * t[q * 32];
        am[0] = sg * t[6 * 32]; am[1] = sg * t[7 * 32]; am[2] = sg * t[8 * 32];
      }
    };
    double axi[6], ami[3];
    Aget(i, axi, ami);
    double acc[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[k][q] = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const int p = rx_pair_index(N, i < j ? i : j, i < j ? j : i);
      double r[3], mj, fac1, rsq, us;
      if (FULL) {
        const double* t = tg + (size_t)p * TGF * 32 + lane;
        r[0] = t[11 * 32]; r[1] = t[12 * 32]; r[2] = t[13 * 32];
        mj = t[(i < j ? 10 : 9) * 32]; fac1 = t[14 * 32]; rsq = t[15 * 32]; us = t[16 * 32];
      } else {
        const double2 a = __ldg(grp(p, 0)), b = __ldg(grp(p, 0) + 1), e = __ldg(grp(p, 1)), f = __ldg(grp(p, 1) + 1), g = __ldg(grp(p, 2));
        r[0] = a.x; r[1] = a.y; r[2] = b.x;
        mj = i < j ? f.x : e.y; fac1 = f.y; rsq = g.x; us = g.y;
      }
      const double sg = i < j ? 1.0 : -1.0;
      double axj[6], amj[3];
      Aget(j, axj, amj);
      double u[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) u[q] = axi[q] - axj[q];
      const double col[4][3] = {{u[0], u[1], u[2]}, {u[1], u[3], u[4]}, {u[2], u[4], u[5]}, {ami[0] - amj[0], ami[1] - amj[1], ami[2] - amj[2]}};
      double ev[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double ru = 3.0 * (r[0] * col[q][0] + r[1] * col[q][1] + r[2] * col[q][2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ev[q][k] = fac1 * (r[k] * ru - rsq * col[q][k]);
      }
      if (d == i || d == j) {
        const double2 g1 = __ldg(grp(p, 2) + 1), h0 = __ldg(grp(p, 3)), h1 = __ldg(grp(p, 3) + 1), k0 = __ldg(grp(p, 4)), k1 = __ldg(grp(p, 4) + 1),
                      l0 = __ldg(grp(p, 5));
        const double F[3] = {g1.x, g1.y, h0.x};
        const double Rm[9] = {h0.y, h1.x, h1.y, k0.x, k0.y, k1.x, k1.y, l0.x, l0.y};
        const double sr = d == i ? 1.0 : -1.0;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int k = 0; k < 3; ++k) ev[q][k] = fma(sr, Rm[3 * k + q], ev[q][k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ev[3][k] = fma(sg * us, r[k], ev[3][k]);
        if (d == j) {
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[k][3] = fma(sg, F[k], acc[k][3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k][q] = fma(mj, ev[q][k], acc[k][q]);
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double2* o = reinterpret_cast<double2*>(OUT + (size_t)((3 * i + k) * N + d) * gs);
        o[0] = make_double2(acc[k][0], acc[k][1]);
        o[1] = make_double2(acc[k][2], acc[k][3]);
      }
    }
  }
}

// The same expansion with fast-kick pairs (nbg_kicks.cuh): records carry a class (the accelerations a^c only sum pairs of the
// same class) and a direct-kick coefficient kappa.  rec_off / out_off: field offsets of the record set and of its dense block.
template <int N> __device__ __forceinline__ void phi_dense_rows_kicked(double* __restrict__ blk, size_t stride, size_t idx, int i, size_t rec_off,
                                                                         size_t out_off) {
  const double* __restrict__ PH = blk + ((rec_off / 4) * stride + idx) * 4;
  double* __restrict__ OUT = blk + ((out_off / 4) * stride + idx) * 4;
  const size_t gs = stride * 4;
  auto grp = [&](int p, int g) { return reinterpret_cast<const double2*>(PH + (size_t)(p * (PF / 4) + g) * gs); };
  struct TG { double T[6], gam[3], ma, mb; int cls; };
  auto load_tg = [&](int a, int b) {
    const int p = rx_pair_index(N, a < b ? a : b, a < b ? b : a);
    const double2 u = __ldg(grp(p, 0)), v = __ldg(grp(p, 0) + 1), e = __ldg(grp(p, 1)), f = __ldg(grp(p, 1) + 1), z = __ldg(grp(p, 5) + 1);
    const double r0 = u.x, r1 = u.y, r2 = v.x, g3 = v.y, g5 = e.x, mi = e.y, mj = f.x;
    const double sg = a < b ? 1.0 : -1.0;
    TG o;
    o.T[0] = g3 - g5 * r0 * r0; o.T[1] = -g5 * r0 * r1; o.T[2] = -g5 * r0 * r2;
    o.T[3] = g3 - g5 * r1 * r1; o.T[4] = -g5 * r1 * r2; o.T[5] = g3 - g5 * r2 * r2;
    o.gam[0] = sg * g3 * r0; o.gam[1] = sg * g3 * r1; o.gam[2] = sg * g3 * r2;
    o.ma = a < b ? mi : mj; o.mb = a < b ? mj : mi;
    o.cls = z.y != 0.0 ? 1 : 0;
    return o;
  };
#pragma unroll 1
  for (int d = 0; d < N; ++d) {
    double diag[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
#pragma unroll
    for (int l = 0; l < N; ++l) {
      if (l == d) continue;
      const TG t = load_tg(d, l);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        diag[0][q] = fma(t.cls ? 0.0 : -t.mb, t.T[q], diag[0][q]);
        diag[1][q] = fma(t.cls ? -t.mb : 0.0, t.T[q], diag[1][q]);
      }
    }
    // A^c[b][d]: x part (6) and mass part (3) of d a_b^c / d (x_d, m_d)
    auto Aget = [&](int b, int c, double (&ax)[6], double (&am)[3]) {
      if (b == d) {
#pragma unroll
        for (int q = 0; q < 6; ++q) ax[q] = c ? diag[1][q] : diag[0][q];
        am[0] = 0.0; am[1] = 0.0; am[2] = 0.0;
      } else {
        const TG t = load_tg(d, b);
        const double on = t.cls == c ? 1.0 : 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) ax[q] = on * t.ma * t.T[q];
        am[0] = on * t.gam[0]; am[1] = on * t.gam[1]; am[2] = on * t.gam[2];
      }
    };
    double acc[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[k][q] = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const int p = rx_pair_index(N, i < j ? i : j, i < j ? j : i);
      const double2 a = __ldg(grp(p, 0)), b = __ldg(grp(p, 0) + 1), e = __ldg(grp(p, 1)), f = __ldg(grp(p, 1) + 1), g = __ldg(grp(p, 2)),
                    z = __ldg(grp(p, 5) + 1);
      const double r[3] = {a.x, a.y, b.x};
      const double g3 = b.y, g5 = e.x, mj = i < j ? f.x : e.y, fac1 = f.y, rsq = g.x, us = g.y, kappa = z.x;
      const int c = z.y != 0.0 ? 1 : 0;
      const double sg = i < j ? 1.0 : -1.0;
      double axi[6], ami[3], axj[6], amj[3];
      Aget(i, c, axi, ami);
      Aget(j, c, axj, amj);
      double u[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) u[q] = axi[q] - axj[q];
      const double col[4][3] = {{u[0], u[1], u[2]}, {u[1], u[3], u[4]}, {u[2], u[4], u[5]}, {ami[0] - amj[0], ami[1] - amj[1], ami[2] - amj[2]}};
      double ev[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double ru = 3.0 * (r[0] * col[q][0] + r[1] * col[q][1] + r[2] * col[q][2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ev[q][k] = fac1 * (r[k] * ru - rsq * col[q][k]);
      }
      if (d == i || d == j) {
        const double2 g1 = __ldg(grp(p, 2) + 1), h0 = __ldg(grp(p, 3)), h1 = __ldg(grp(p, 3) + 1), k0 = __ldg(grp(p, 4)), k1 = __ldg(grp(p, 4) + 1),
                      l0 = __ldg(grp(p, 5));
        const double F[3] = {g1.x, g1.y, h0.x};
        const double Rm[9] = {h0.y, h1.x, h1.y, k0.x, k0.y, k1.x, k1.y, l0.x, l0.y};
        const double T[9] = {g3 - g5 * r[0] * r[0], -g5 * r[0] * r[1], -g5 * r[0] * r[2], -g5 * r[1] * r[0], g3 - g5 * r[1] * r[1], -g5 * r[1] * r[2],
                             -g5 * r[2] * r[0], -g5 * r[2] * r[1], g3 - g5 * r[2] * r[2]};
        const double sr = d == i ? 1.0 : -1.0;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int k = 0; k < 3; ++k) ev[q][k] = fma(sr, Rm[3 * k + q] - kappa * T[3 * k + q], ev[q][k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) ev[3][k] = fma(sg * us, r[k], ev[3][k]);
        if (d == j) {
#pragma unroll
          for (int k = 0; k < 3; ++k) acc[k][3] += sg * (F[k] - kappa * g3 * r[k]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k][q] = fma(mj, ev[q][k], acc[k][q]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double2* o = reinterpret_cast<double2*>(OUT + (size_t)((3 * i + k) * N + d) * gs);
      o[0] = make_double2(acc[k][0], acc[k][1]);
      o[1] = make_double2(acc[k][2], acc[k][3]);
    }
  }
}

// jac_step (+)= jac_phi * jac_step with the dense operator W (layout of phi_dense_fields) staged in shared memory;
// jv/je at offset OFF (position p holds body (OFF + p) mod N).  The x half multiplies the x rows of bodies 0..NA-1, the
// v half (which receives them by shuffle) those of bodies NA..N-1, each into all 3N outputs; the v half adds the two.
// HOLD: instead of adding, park dv in hold[(3 b + k) * hs] (first kickfast!: jac_kick * jac_step is formed BEFORE drift_grad!
// and added after it, ahl21.jl:16-23); rx_fold_add applies it.
template <int N, int OFF, bool HOLD = false>
__device__ __forceinline__ void rx_phisalpha_dense(RxState<N>& S, const double* __restrict__ W, int half, int c, double* hold = nullptr, int hs = 0) {
  constexpr int NA = (N + 1) / 2;
  double in[NA][3];
  static_for<0, NA>([&](auto Qc) {
    constexpr int q = decltype(Qc)::value;
    constexpr int pa = (q - OFF + 2 * N) % N;              // position of body q (x half's input)
    constexpr int bb = NA + q;                             // body whose x rows the v half receives
    if constexpr (bb < N) {
      constexpr int pb = (bb - OFF + 2 * N) % N;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double got = shx(S.jv[pb][k]);
        in[q][k] = half ? got : S.jv[pa][k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) in[q][k] = half ? 0.0 : S.jv[pa][k];
    }
  });
  const int dm = (c % 7 == 6 && c < 7 * N) ? c / 7 : -1;   // this column is the mass column of body dm
  const int d0 = half ? (NA < N ? NA : 0) : 0;             // first body of this half's inputs (odd N: last v-half slot is padding)
  static_for<0, N>([&](auto Bc) {
    constexpr int b = decltype(Bc)::value;
    constexpr int pos = (b - OFF + 2 * N) % N;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* __restrict__ row = W + ((3 * b + k) * N + d0) * 4;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int q = 0; q < NA; ++q) {
        // odd N: the v half's last input slot is zero padding; point it at a valid weight instead of reading past the row
        const int qq = ((N & 1) && q == NA - 1 && half) ? q - 1 : q;
        const double2 w01 = *reinterpret_cast<const double2*>(row + 4 * qq);
        const double w2 = row[4 * qq + 2];
        s0 = fma(w01.x, in[q][0], s0);
        s1 = fma(w01.y, in[q][1], s1);
        s0 = fma(w2, in[q][2], s0);
      }
      const double part = s0 + s1;
      const double other = shx(part);
      double dv = part + other;
      if (dm >= 0) dv += W[((3 * b + k) * N + dm) * 4 + 3];
      // comp_sum_matrix!(jac_step, jac_error, jac_phi * jac_step): v rows get dv, x rows a zero addend (fold)
      if (HOLD) hold[(3 * b + k) * hs] = dv;
      else ksum_m(S.jv[pos][k], S.je[pos][k], half == 1 ? dv : 0.0);
    }
  });
}
template <int N> __device__ __forceinline__ void rx_fold_add(RxState<N>& S, const double* hold, int hs, int half) {  // offset 0
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) ksum_m(S.jv[b][k], S.je[b][k], half == 1 ? hold[(3 * b + k) * hs] : 0.0);
}

// ---- sweeps over pairs in blocks of U pivot bodies ---------------------------------------------------------------
// "offset r" = position p holds body (r + p) mod N.  A block of U consecutive pivots runs with the lowest pivot at
// position 0 (pivot u at static position u, its partner at static position u + T); between blocks the arrays rotate by U.
// U = 1 minimises code, U = N is the full unroll; U trades register moves against instruction-cache footprint.
template <int N, int U, bool SYNC = true> struct RxSweep {
  static constexpr int NB = (N - 1 + U - 1) / U;      // blocks covering pivots 0 .. N-2
  static constexpr int A1 = (NB * U) % N;             // offset after an ascending pass
  static constexpr int KTOP = (N - 2) / U;            // top block of the descending sweep
  // ascending pairs (i < j in the reference order); F(PA, PB, record pointer, body_i, body_j)
  template <class Arr, class F> static __device__ __forceinline__ void asc(Arr&& rotate, const double* R, int stride, F&& f) {
#pragma unroll 1
    for (int g0 = 0; g0 < N - 1; g0 += U) {
      static_for<0, U>([&](auto Uc) {
        constexpr int u = decltype(Uc)::value;
        const int g = g0 + u;
        if (g < N - 1) {
          const int tmax = N - 1 - g;
          static_for<1, N - u>([&](auto Tc) {
            constexpr int T = decltype(Tc)::value;
            if (T <= tmax) f(std::integral_constant<int, u>{}, std::integral_constant<int, u + T>{}, R + (T - 1) * stride, g, g + T);
          });
          R += tmax * stride;
        }
      });
      rotate(std::integral_constant<int, U>{});
      // keep the warps of the block within one group of each other: the instruction stream is long and straight, and
      // warps at unrelated program counters thrash the instruction cache (ncu r01_rx4: no_instruction 1.0 per issue)
      if (SYNC) __syncthreads();
    }
  }
  // descending pairs (i = N-2 .. 0, j = N-1 .. i+1); entered at offset KTOP*U, leaves at offset 0
  template <class Arr, class F> static __device__ __forceinline__ void desc(Arr&& rotate, const double* R, int stride, F&& f) {
#pragma unroll 1
    for (int kb = KTOP; kb >= 0; --kb) {
      static_for<0, U>([&](auto Uc) {
        constexpr int u = U - 1 - decltype(Uc)::value;  // U-1 down to 0
        const int i = kb * U + u;
        if (i <= N - 2) {
          const int tmax = N - 1 - i;
          static_for<1, N - u>([&](auto Vc) {
            constexpr int T = N - u - decltype(Vc)::value;  // N-1-u down to 1
            if (T <= tmax) f(std::integral_constant<int, u>{}, std::integral_constant<int, u + T>{}, R + (tmax - T) * stride, i, i + T);
          });
          R += tmax * stride;
        }
      });
      if (kb > 0) rotate(std::integral_constant<int, -U>{});
      if (SYNC) __syncthreads();
    }
  }
};

// one AHL21 Jacobian step from a staged operator block [2P Kepler records | dense operators]; offset 0 on entry and exit.
// KICK: fast-kick pairs present (kmask, nbg_kicks.cuh): three dense operators (kickfast!, phic!+phisalpha!, kickfast!), the
// flagged pairs are skipped in the Kepler sweeps; hold/hs: per-thread scratch for the first kick (see rx_phisalpha_dense).
template <int N, int U, bool SYNC = true, bool KICK = false>
__device__ __forceinline__ void rx_step(RxState<N>& S, const double* __restrict__ blk, double h2, int half, int c, const KMask& kmask = KMask{},
                                        double* hold = nullptr, int hs = 0) {
  constexpr int P = N * (N - 1) / 2, DF = 12 * N * N;
  using SW = RxSweep<N, U, SYNC>;
  auto rotate = [&](auto Kc) { rot_left_by<N, decltype(Kc)::value>(S.jv); rot_left_by<N, decltype(Kc)::value>(S.je); };
  auto pair = [&](auto PA, auto PB, const double* R, int bi, int bj) {
    if (!KICK || !kmask.bit(rx_pair_index(N, bi, bj))) rx_pair<N, decltype(PA)::value, decltype(PB)::value>(S, R, half, c, 7 * bi + 6, 7 * bj + 6);
  };
  const double* __restrict__ D = blk + 2 * P * KF;
  if (KICK) rx_phisalpha_dense<N, 0, true>(S, D, half, c, hold, hs);
  rx_drift<N>(S, h2, half);
  if (KICK) rx_fold_add<N>(S, hold, hs, half);
  else rx_fold<N>(S);
  SW::asc(rotate, blk, KF, pair);                                          // offset 0 -> A1
  rx_phisalpha_dense<N, SW::A1>(S, D + (KICK ? DF : 0), half, c);         // A1 -> A1
  rotate(std::integral_constant<int, SW::KTOP * U - SW::A1>{});            // A1 -> KTOP*U
  SW::desc(rotate, blk + P * KF, KF, pair);                                // -> 0
  rx_drift<N>(S, h2, half);
  if (KICK) rx_phisalpha_dense<N, 0>(S, D + 2 * DF, half, c);
  else rx_fold<N>(S);
}

}  // namespace nbg
