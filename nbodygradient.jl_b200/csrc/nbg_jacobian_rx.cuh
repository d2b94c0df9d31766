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
    // two 3-term chains (own half / partner half) instead of one 6-term chain: shorter dependent latency; the
    // partner-half chain also waits on the shuffle, the own-half chain does not
    double s = Kb[3 * k] * md[0];
    s = fma(Kb[3 * k + 1], md[1], s);
    s = fma(Kb[3 * k + 2], md[2], s);
    double u = Kb[9 + 3 * k] * od[0];
    u = fma(Kb[9 + 3 * k + 1], od[1], u);
    u = fma(Kb[9 + 3 * k + 2], od[2], u);
    w[k] = s + u;
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
    ksum_m(S.jv[PA][k], S.je[PA][k], ai[k]);
    ksum_m(S.jv[PB][k], S.je[PB][k], aj[k]);
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

// phisalpha Jacobian in factored form (see nbg_step.cuh).  The x half computes; the v half's aux registers hold the
// per-body da accumulators during pass 1/2, the x half's aux registers hold the dv accumulators during pass 2.
// Pass 1, positions 0 and T: x half forms Gam_ij (dx_i - dx_j) (+ mass term) and ships it; v half accumulates da.
template <int N, int PA, int PB>
__device__ __forceinline__ void rx_phi1(const RxState<N>& S, double (&aux)[N][3], const double* __restrict__ R, int half, int c, int ci, int cj) {
  const double r0 = R[PF_R], r1 = R[PF_R + 1], r2v = R[PF_R + 2], g3 = R[PF_G3];
  const double mi = R[PF_MI], mj = R[PF_MJ];
  double w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = S.jv[PA][k] - S.jv[PB][k];
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
    if (half == 1) { aux[PA][k] -= ti_; aux[PB][k] += tj_; }
  }
}
// Pass 2, positions 0 and T: v half ships da_i - da_j, x half forms dF and accumulates dv in its aux.
template <int N, int PA, int PB>
__device__ __forceinline__ void rx_phi2(const RxState<N>& S, double (&aux)[N][3], const double* __restrict__ R, int half, int c, int ci, int cj) {
  const double r0 = R[PF_R], r1 = R[PF_R + 1], r2v = R[PF_R + 2], fac1 = R[PF_FAC1], r2 = R[PF_R2], us = R[PF_US];
  const double mi = R[PF_MI], mj = R[PF_MJ];
  double w[3], wa[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { w[k] = S.jv[PA][k] - S.jv[PB][k]; wa[k] = shx(aux[PA][k] - aux[PB][k]); }
  if (half == 0) {
    const double rwa = r0 * wa[0] + r1 * wa[1] + r2v * wa[2];
    const double dmi = (c == ci) ? 1.0 : 0.0, dmj = (c == cj) ? 1.0 : 0.0;
    const double rr[3] = {r0, r1, r2v};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dF = R[PF_RM + 3 * k] * w[0] + R[PF_RM + 3 * k + 1] * w[1] + R[PF_RM + 3 * k + 2] * w[2] +
                        fac1 * (3.0 * rr[k] * rwa - r2 * wa[k]) + us * rr[k] * (dmi + dmj);
      const double F = R[PF_F + k];
      aux[PA][k] += mj * dF + F * dmj;
      aux[PB][k] -= mi * dF + F * dmi;
    }
  }
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

// phisalpha: jv and aux enter at offset A1 (je stays there throughout), leave at offset A1
template <int N, int U> __device__ __forceinline__ void rx_phisalpha(RxState<N>& S, const double* __restrict__ PH, int half, int c) {
  using SW = RxSweep<N, U>;
  double aux[N][3];
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) aux[b][k] = 0.0;
  auto rotate = [&](auto Kc) { rot_left_by<N, decltype(Kc)::value>(S.jv); rot_left_by<N, decltype(Kc)::value>(aux); };
  rotate(std::integral_constant<int, N - SW::A1>{});  // to offset 0
  SW::asc(rotate, PH, PF, [&](auto PA, auto PB, const double* R, int bi, int bj) {
    rx_phi1<N, decltype(PA)::value, decltype(PB)::value>(S, aux, R, half, c, 7 * bi + 6, 7 * bj + 6);
  });
  rotate(std::integral_constant<int, N - SW::A1>{});
  SW::asc(rotate, PH, PF, [&](auto PA, auto PB, const double* R, int bi, int bj) {
    rx_phi2<N, decltype(PA)::value, decltype(PB)::value>(S, aux, R, half, c, 7 * bi + 6, 7 * bj + 6);
  });
  // jv, je, aux all at offset A1
  // comp_sum_matrix!(jac_step, jac_error, jac_phi * jac_step): v rows get dv, x rows a zero addend (fold)
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dv = shx(aux[b][k]);  // v half receives the x half's accumulator
      ksum_m(S.jv[b][k], S.je[b][k], half == 1 ? dv : 0.0);
    }
}

// one AHL21 Jacobian step from a staged operator block; offset 0 (identity) on entry and exit
template <int N, int U> __device__ __forceinline__ void rx_step(RxState<N>& S, const double* __restrict__ blk, double h2, int half, int c) {
  constexpr int P = N * (N - 1) / 2;
  using SW = RxSweep<N, U>;
  auto rotate = [&](auto Kc) { rot_left_by<N, decltype(Kc)::value>(S.jv); rot_left_by<N, decltype(Kc)::value>(S.je); };
  auto pair = [&](auto PA, auto PB, const double* R, int bi, int bj) {
    rx_pair<N, decltype(PA)::value, decltype(PB)::value>(S, R, half, c, 7 * bi + 6, 7 * bj + 6);
  };
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
  SW::asc(rotate, blk, KF, pair);                                          // offset 0 -> A1
  rx_phisalpha<N, U>(S, blk + 2 * P * KF, half, c);                        // A1 -> A1
  rotate(std::integral_constant<int, SW::KTOP * U - SW::A1>{});            // A1 -> KTOP*U
  SW::desc(rotate, blk + P * KF, KF, pair);                                // -> 0
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
}

}  // namespace nbg
