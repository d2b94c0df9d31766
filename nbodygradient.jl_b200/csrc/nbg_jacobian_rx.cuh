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
#include <type_traits>
#include "nbg_jacobian.cuh"

namespace nbg {

constexpr unsigned FULL = 0xffffffffu;
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

// async copy of one step's operator block (sf doubles, 4-packed layout) into shared memory
__device__ __forceinline__ void rx_fetch(double* dst, const double* base, size_t stride, size_t idx, int ngroups, int tid, int nthr) {
  for (int g = tid; g < ngroups; g += nthr) {
    const double* src = base + ((size_t)g * stride + idx) * 4;
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
template <int N> __device__ __forceinline__ void rot_left(double (&a)[N][3]) {
  const double t0 = a[0][0], t1 = a[0][1], t2 = a[0][2];
#pragma unroll
  for (int p = 0; p < N - 1; ++p) { a[p][0] = a[p + 1][0]; a[p][1] = a[p + 1][1]; a[p][2] = a[p + 1][2]; }
  a[N - 1][0] = t0; a[N - 1][1] = t1; a[N - 1][2] = t2;
}
template <int N> __device__ __forceinline__ void rot_right(double (&a)[N][3]) {
  const double t0 = a[N - 1][0], t1 = a[N - 1][1], t2 = a[N - 1][2];
#pragma unroll
  for (int p = N - 1; p > 0; --p) { a[p][0] = a[p - 1][0]; a[p][1] = a[p - 1][1]; a[p][2] = a[p - 1][2]; }
  a[0][0] = t0; a[0][1] = t1; a[0][2] = t2;
}

// pair update between positions 0 and T; ci/cj = mass columns (7*body+6) of the bodies at those positions
template <int N, int T>
__device__ __forceinline__ void rx_pair(RxState<N>& S, const double* __restrict__ R, int half, int c, int ci, int cj) {
  double md[3], od[3], w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) md[k] = S.jv[0][k] - S.jv[T][k];
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
  if (c == ci || c == cj) {
    const double* mb = R + 38 + 12 * half + (c == ci ? 0 : 6);
#pragma unroll
    for (int k = 0; k < 3; ++k) { ai[k] += mb[k]; aj[k] += mb[3 + k]; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ksum_m(S.jv[0][k], S.je[0][k], ai[k]);
    ksum_m(S.jv[T][k], S.je[T][k], aj[k]);
  }
}

// group of the ascending sweep: body g (position 0) with bodies g+1 .. N-1 (positions 1 .. N-1-g), ascending
template <int N> __device__ __forceinline__ void rx_group_asc(RxState<N>& S, const double* __restrict__ R, int g, int half, int c) {
  const int tmax = N - 1 - g;
  static_for<1, N>([&](auto Tc) {
    constexpr int T = decltype(Tc)::value;
    if (T <= tmax) rx_pair<N, T>(S, R + (T - 1) * KF, half, c, 7 * g + 6, 7 * (g + T) + 6);
  });
}
// group of the descending sweep: body i (position 0) with bodies N-1 .. i+1 (positions N-1-i .. 1), descending
template <int N> __device__ __forceinline__ void rx_group_desc(RxState<N>& S, const double* __restrict__ R, int i, int half, int c) {
  const int tmax = N - 1 - i;
  static_for<1, N>([&](auto Uc) {
    constexpr int T = N - decltype(Uc)::value;  // N-1 down to 1
    if (T <= tmax) rx_pair<N, T>(S, R + (tmax - T) * KF, half, c, 7 * i + 6, 7 * (i + T) + 6);
  });
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

// phisalpha Jacobian in factored form (see nbg_step.cuh).  The x half computes; the v half's aux registers hold the
// per-body da accumulators during pass 1/2, the x half's aux registers hold the dv accumulators during pass 2.
// Pass 1, positions 0 and T: x half forms Gam_ij (dx_i - dx_j) (+ mass term) and ships it; v half accumulates da.
template <int N, int T>
__device__ __forceinline__ void rx_phi1(const RxState<N>& S, double (&aux)[N][3], const double* __restrict__ R, int half, int c, int ci, int cj) {
  const double r0 = R[PF_R], r1 = R[PF_R + 1], r2v = R[PF_R + 2], g3 = R[PF_G3];
  const double mi = R[PF_MI], mj = R[PF_MJ];
  double w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = S.jv[0][k] - S.jv[T][k];
  const double rw = r0 * w[0] + r1 * w[1] + r2v * w[2];
  const double f3 = R[PF_G5] * rw;
  const double dmj = (c == cj) ? 1.0 : 0.0, dmi = (c == ci) ? 1.0 : 0.0;
  const double rr[3] = {r0, r1, r2v};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double gw = g3 * w[k] - f3 * rr[k];
    const double ga = g3 * rr[k];
    const double ti_ = shx(mj * gw + ga * dmj);  // contribution to -da_i
    const double tj_ = shx(mi * gw + ga * dmi);  // contribution to +da_j
    if (half == 1) { aux[0][k] -= ti_; aux[T][k] += tj_; }
  }
}
// Pass 2, positions 0 and T: v half ships da_i - da_j, x half forms dF and accumulates dv in its aux.
template <int N, int T>
__device__ __forceinline__ void rx_phi2(const RxState<N>& S, double (&aux)[N][3], const double* __restrict__ R, int half, int c, int ci, int cj) {
  const double r0 = R[PF_R], r1 = R[PF_R + 1], r2v = R[PF_R + 2], fac1 = R[PF_FAC1], r2 = R[PF_R2], us = R[PF_US];
  const double mi = R[PF_MI], mj = R[PF_MJ];
  double w[3], wa[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { w[k] = S.jv[0][k] - S.jv[T][k]; wa[k] = shx(aux[0][k] - aux[T][k]); }
  if (half == 0) {
    const double rwa = r0 * wa[0] + r1 * wa[1] + r2v * wa[2];
    const double dmi = (c == ci) ? 1.0 : 0.0, dmj = (c == cj) ? 1.0 : 0.0;
    const double rr[3] = {r0, r1, r2v};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dF = R[PF_RM + 3 * k] * w[0] + R[PF_RM + 3 * k + 1] * w[1] + R[PF_RM + 3 * k + 2] * w[2] +
                        fac1 * (3.0 * rr[k] * rwa - r2 * wa[k]) + us * rr[k] * (dmi + dmj);
      const double F = R[PF_F + k];
      aux[0][k] += mj * dF + F * dmj;
      aux[T][k] -= mi * dF + F * dmi;
    }
  }
}

// On entry and exit the arrangement is "rotated left by N-1" ([N-1, 0, 1, ..., N-2]); each pass makes N left rotations.
template <int N> __device__ __forceinline__ void rx_phisalpha(RxState<N>& S, const double* __restrict__ PH, int half, int c) {
  double aux[N][3];
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) aux[b][k] = 0.0;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const double* R = PH;
#pragma unroll 1
    for (int g = 0; g < N; ++g) {
      rot_left<N>(S.jv);
      rot_left<N>(aux);
      const int tmax = N - 1 - g;
      if (pass == 0) {
        static_for<1, N>([&](auto Tc) {
          constexpr int T = decltype(Tc)::value;
          if (T <= tmax) rx_phi1<N, T>(S, aux, R + (T - 1) * PF, half, c, 7 * g + 6, 7 * (g + T) + 6);
        });
      } else {
        static_for<1, N>([&](auto Tc) {
          constexpr int T = decltype(Tc)::value;
          if (T <= tmax) rx_phi2<N, T>(S, aux, R + (T - 1) * PF, half, c, 7 * g + 6, 7 * (g + T) + 6);
        });
      }
      R += tmax * PF;
    }
  }
  // arrangement is [N-1, 0, ..., N-2] again, for jv, je (never moved) and aux alike
  // comp_sum_matrix!(jac_step, jac_error, jac_phi * jac_step): v rows get dv, x rows a zero addend (fold)
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dv = shx(aux[b][k]);  // v half receives the x half's accumulator
      ksum_m(S.jv[b][k], S.je[b][k], half == 1 ? dv : 0.0);
    }
}

// one AHL21 Jacobian step from a staged operator block; arrangement is the identity on entry and exit
template <int N> __device__ __forceinline__ void rx_step(RxState<N>& S, const double* __restrict__ blk, double h2, int half, int c) {
  constexpr int P = N * (N - 1) / 2;
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
  {
    const double* R = blk;
#pragma unroll 1
    for (int g = 0; g < N - 1; ++g) {
      rx_group_asc<N>(S, R, g, half, c);
      R += (N - 1 - g) * KF;
      rot_left<N>(S.jv);
      rot_left<N>(S.je);
    }
  }
  rx_phisalpha<N>(S, blk + 2 * P * KF, half, c);
  {
    const double* R = blk + P * KF;
#pragma unroll 1
    for (int i = N - 2; i >= 0; --i) {
      rot_right<N>(S.jv);
      rot_right<N>(S.je);
      rx_group_desc<N>(S, R, i, half, c);
      R += (N - 1 - i) * KF;
    }
  }
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
}

}  // namespace nbg
