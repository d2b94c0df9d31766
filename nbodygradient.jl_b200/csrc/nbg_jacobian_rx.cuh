// Register-resident Jacobian propagation for N <= 8 bodies (the production path; nbg_jacobian.cuh is the generic
// shared-memory version used for larger N).
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
#include "nbg_jacobian.cuh"

namespace nbg {

constexpr unsigned FULL = 0xffffffffu;
__host__ __device__ constexpr int rx_warps(int n) { return (7 * n + 15) / 16; }

__device__ __forceinline__ double shx(double v) { return __shfl_xor_sync(FULL, v, 16); }

template <int N> struct RxState {
  double jv[N][3];
  double je[N][3];
};

// async copy of one step's operator block (sf doubles, 4-packed layout) into shared memory
__device__ __forceinline__ void rx_fetch(double* dst, const double* base, size_t stride, size_t idx, int ngroups, int tid, int nthr) {
  for (int g = tid; g < ngroups; g += nthr) {
    const double* src = base + ((size_t)g * stride + idx) * 4;
    __pipeline_memcpy_async(dst + 4 * g, src, 16);
    __pipeline_memcpy_async(dst + 4 * g + 2, src + 2, 16);
  }
  __pipeline_commit();
}

template <int N, int I, int J>
__device__ __forceinline__ void rx_pair(RxState<N>& S, const double* __restrict__ R, int half, int c) {
  double md[3], od[3], w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) md[k] = S.jv[I][k] - S.jv[J][k];
#pragma unroll
  for (int k = 0; k < 3; ++k) od[k] = shx(md[k]);
  const double* __restrict__ Kb = R + 18 * half;  // [A (3x3) | B (3x3)] of this half
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = Kb[3 * k] * md[0];
    s = fma(Kb[3 * k + 1], md[1], s);
    s = fma(Kb[3 * k + 2], md[2], s);
    s = fma(Kb[9 + 3 * k], od[0], s);
    s = fma(Kb[9 + 3 * k + 1], od[1], s);
    s = fma(Kb[9 + 3 * k + 2], od[2], s);
    w[k] = s;
  }
  const double2 mm = *reinterpret_cast<const double2*>(R + KF_MI);
  double ai[3], aj[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { ai[k] = mm.y * w[k]; aj[k] = -mm.x * w[k]; }
  if (c == 7 * I + 6 || c == 7 * J + 6) {
    const double* mb = R + 38 + 12 * half;
    if (c == 7 * I + 6) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { ai[k] += mb[k]; aj[k] += mb[3 + k]; }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) { ai[k] += mb[6 + k]; aj[k] += mb[9 + k]; }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ksum_m(S.jv[I][k], S.je[I][k], ai[k]);
    ksum_m(S.jv[J][k], S.je[J][k], aj[k]);
  }
}

// compile-time pair loops
template <int N, int I, int J> struct RxAsc {
  static __device__ __forceinline__ void run(RxState<N>& S, const double* R, int half, int c) {
    rx_pair<N, I, J>(S, R, half, c);
    if constexpr (J + 1 < N) RxAsc<N, I, J + 1>::run(S, R + KF, half, c);
    else if constexpr (I + 2 < N) RxAsc<N, I + 1, I + 2>::run(S, R + KF, half, c);
  }
};
template <int N, int I, int J> struct RxDesc {
  static __device__ __forceinline__ void run(RxState<N>& S, const double* R, int half, int c) {
    rx_pair<N, I, J>(S, R, half, c);
    if constexpr (J - 1 > I) RxDesc<N, I, J - 1>::run(S, R + KF, half, c);
    else if constexpr (I - 1 >= 0) RxDesc<N, I - 1, N - 1>::run(S, R + KF, half, c);
  }
};

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

// phisalpha Jacobian in factored form (see nbg_step.cuh).  The x half computes; the v half's aux registers hold the
// per-body da accumulators during pass 1/2, the x half's aux registers hold the dv accumulators during pass 2.
// Pair loops are compile-time recursions (all register indexing static).
template <int N, int I, int J> struct RxPhi1 {  // pass 1: x half forms Gam_ij (dx_i - dx_j) (+ mass term) and ships it; v half accumulates da
  static __device__ __forceinline__ void run(const RxState<N>& S, double (&aux)[N][3], const double* __restrict__ R, int half, int c) {
    const double r0 = R[PF_R], r1 = R[PF_R + 1], r2v = R[PF_R + 2], g3 = R[PF_G3];
    const double mi = R[PF_MI], mj = R[PF_MJ];
    double w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = S.jv[I][k] - S.jv[J][k];
    const double rw = r0 * w[0] + r1 * w[1] + r2v * w[2];
    const double f3 = R[PF_G5] * rw;
    const double dmj = (c == 7 * J + 6) ? 1.0 : 0.0, dmi = (c == 7 * I + 6) ? 1.0 : 0.0;
    const double rr[3] = {r0, r1, r2v};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double gw = g3 * w[k] - f3 * rr[k];
      const double ga = g3 * rr[k];
      const double ti_ = shx(mj * gw + ga * dmj);  // contribution to -da_i
      const double tj_ = shx(mi * gw + ga * dmi);  // contribution to +da_j
      if (half == 1) { aux[I][k] -= ti_; aux[J][k] += tj_; }
    }
    if constexpr (J + 1 < N) RxPhi1<N, I, J + 1>::run(S, aux, R + PF, half, c);
    else if constexpr (I + 2 < N) RxPhi1<N, I + 1, I + 2>::run(S, aux, R + PF, half, c);
  }
};
template <int N, int I, int J> struct RxPhi2 {  // pass 2: v half ships da_i - da_j, x half forms dF and accumulates dv in its aux
  static __device__ __forceinline__ void run(const RxState<N>& S, double (&aux)[N][3], const double* __restrict__ R, int half, int c) {
    const double r0 = R[PF_R], r1 = R[PF_R + 1], r2v = R[PF_R + 2], fac1 = R[PF_FAC1], r2 = R[PF_R2], us = R[PF_US];
    const double mi = R[PF_MI], mj = R[PF_MJ];
    double w[3], wa[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { w[k] = S.jv[I][k] - S.jv[J][k]; wa[k] = shx(aux[I][k] - aux[J][k]); }
    if (half == 0) {
      const double rwa = r0 * wa[0] + r1 * wa[1] + r2v * wa[2];
      const double dmi = (c == 7 * I + 6) ? 1.0 : 0.0, dmj = (c == 7 * J + 6) ? 1.0 : 0.0;
      const double rr[3] = {r0, r1, r2v};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double dF = R[PF_RM + 3 * k] * w[0] + R[PF_RM + 3 * k + 1] * w[1] + R[PF_RM + 3 * k + 2] * w[2] +
                          fac1 * (3.0 * rr[k] * rwa - r2 * wa[k]) + us * rr[k] * (dmi + dmj);
        const double F = R[PF_F + k];
        aux[I][k] += mj * dF + F * dmj;
        aux[J][k] -= mi * dF + F * dmi;
      }
    }
    if constexpr (J + 1 < N) RxPhi2<N, I, J + 1>::run(S, aux, R + PF, half, c);
    else if constexpr (I + 2 < N) RxPhi2<N, I + 1, I + 2>::run(S, aux, R + PF, half, c);
  }
};

template <int N> __device__ __forceinline__ void rx_phisalpha(RxState<N>& S, const double* __restrict__ PH, int half, int c) {
  double aux[N][3];
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) aux[b][k] = 0.0;
  RxPhi1<N, 0, 1>::run(S, aux, PH, half, c);
  RxPhi2<N, 0, 1>::run(S, aux, PH, half, c);
  // comp_sum_matrix!(jac_step, jac_error, jac_phi * jac_step): v rows get dv, x rows a zero addend (fold)
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dv = shx(aux[b][k]);  // v half receives the x half's accumulator
      ksum_m(S.jv[b][k], S.je[b][k], half == 1 ? dv : 0.0);
    }
}

// one AHL21 Jacobian step from a staged operator block
template <int N> __device__ __forceinline__ void rx_step(RxState<N>& S, const double* __restrict__ blk, double h2, int half, int c) {
  constexpr int P = N * (N - 1) / 2;
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
  RxAsc<N, 0, 1>::run(S, blk, half, c);
  rx_phisalpha<N>(S, blk + 2 * P * KF, half, c);
  RxDesc<N, N - 2, N - 1>::run(S, blk + P * KF, half, c);
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
}

}  // namespace nbg
