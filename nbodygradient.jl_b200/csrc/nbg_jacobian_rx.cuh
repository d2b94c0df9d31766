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

// ---- phisalpha as a DENSE operator -----------------------------------------------------------------------------
// jac_phi (ahl21.jl:635-693) is nonzero only in its v rows x {x, m} columns.  Applying it pair by pair in factored form
// costs ~85 FP64 instructions per pair and thread (r01 profile: 45% of the kernel's FP64 issue for 5% of the canonical
// flops); as a dense 3N x 3N block it is 9N^2/2 DFMA per thread.  The trajectory kernel still streams the compact
// per-pair records (PF doubles); the block assembles the dense blocks cooperatively in shared memory once per step:
//   a_i = - sum_d G m_d r_id / r^3,   T_p = G (I/r^3 - 3 r r^T / r^5),   S_p = fac1 (3 r r^T - r^2 I)      (p = pair)
//   Ax[i][d] = m_d T_id (d != i),  Ax[i][i] = - sum_l Ax[i][l];   Am[i][d] = - G r_id / r^3 (d != i),  Am[i][i] = 0
//   Phi_x[i][d] = sum_{j != i} m_j S_ij (Ax[i][d] - Ax[j][d])  +  (d == i ? sum_j m_j Rm_ij : - m_d Rm_id)
//   Phi_m[i][d] = sum_{j != i} m_j S_ij (Am[i][d] - Am[j][d])  +  (d == i ? sum_j m_j us_ij r_ij : m_d us_id r_id + F_id)
// with r_ij, F_ij oriented from body i (sign flips for i > j); same linear operator as the reference's jac_phi.
template <int N> struct RxPhi {
  static constexpr int NA = (N + 1) / 2;        // bodies whose x rows the x half multiplies; the v half takes the rest
  static constexpr int KIN = 3 * NA;            // inputs per half (zero padded when N is odd)
  static constexpr int PHX = 0;                 // [half][3N rows][KIN]
  static constexpr int PHM = PHX + 2 * 3 * N * KIN;  // [3N rows][N]
  static constexpr int AX = PHM + 3 * N * N;    // [i][d][6]  symmetric 3x3: xx xy xz yy yz zz
  static constexpr int AM = AX + 6 * N * N;     // [i][d][3]
  static constexpr int SIZE = AM + 3 * N * N;   // doubles of scratch
  static constexpr bool ALIAS = SIZE <= (N * (N - 1) / 2) * KF;  // fits in the (dead) ascending-sweep records of the current buffer
};
__device__ __forceinline__ int rx_pair_index(int n, int a, int b) { return a * n - a * (a + 1) / 2 + (b - a - 1); }  // a < b

template <int N> __device__ __forceinline__ void rx_phi_assemble(const double* __restrict__ PH, double* __restrict__ W, int tid, int nthr) {
  using L = RxPhi<N>;
  // stage A: off-diagonal blocks of Ax, Am
  for (int t = tid; t < N * N; t += nthr) {
    const int i = t / N, d = t % N;
    if (i != d) {
      const double* R = PH + rx_pair_index(N, i < d ? i : d, i < d ? d : i) * PF;
      const double sg = i < d ? 1.0 : -1.0;
      const double r0 = R[PF_R], r1 = R[PF_R + 1], r2 = R[PF_R + 2], g3 = R[PF_G3], g5 = R[PF_G5];
      const double md = i < d ? R[PF_MJ] : R[PF_MI];
      double* a = W + L::AX + 6 * t;
      a[0] = md * (g3 - g5 * r0 * r0); a[1] = md * (-g5 * r0 * r1); a[2] = md * (-g5 * r0 * r2);
      a[3] = md * (g3 - g5 * r1 * r1); a[4] = md * (-g5 * r1 * r2); a[5] = md * (g3 - g5 * r2 * r2);
      double* am = W + L::AM + 3 * t;
      am[0] = -sg * g3 * r0; am[1] = -sg * g3 * r1; am[2] = -sg * g3 * r2;
    } else {
      double* am = W + L::AM + 3 * t;
      am[0] = 0.0; am[1] = 0.0; am[2] = 0.0;
    }
  }
  if (N & 1)  // zero padding of the v half's unused inputs
    for (int t = tid; t < 3 * N * 3; t += nthr) W[L::PHX + (3 * N + t / 3) * L::KIN + (L::KIN - 3) + t % 3] = 0.0;
  __syncthreads();
  // stage B: diagonal blocks
  for (int t = tid; t < 6 * N; t += nthr) {
    const int i = t / 6, q = t % 6;
    double s = 0.0;
    for (int l = 0; l < N; ++l)
      if (l != i) s -= W[L::AX + 6 * (i * N + l) + q];
    W[L::AX + 6 * (i * N + i) + q] = s;
  }
  __syncthreads();
  // stage C: one task per (i, d, input column): p < 3 an x column of body d, p == 3 its mass column
  for (int t = tid; t < N * N * 4; t += nthr) {
    const int p = t & 3, d = (t >> 2) % N, i = (t >> 2) / N;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const double* R = PH + rx_pair_index(N, i < j ? i : j, i < j ? j : i) * PF;
      const double sg = i < j ? 1.0 : -1.0;
      const double mj = i < j ? R[PF_MJ] : R[PF_MI];
      const double r0 = R[PF_R], r1 = R[PF_R + 1], r2 = R[PF_R + 2], fac1 = R[PF_FAC1], rsq = R[PF_R2];
      double u0, u1, u2;  // column p of A[i][d] - A[j][d]
      if (p < 3) {
        const double* ai = W + L::AX + 6 * (i * N + d);
        const double* aj = W + L::AX + 6 * (j * N + d);
        // symmetric storage: column p of [0 1 2; 1 3 4; 2 4 5]
        const int q0 = p, q1 = p == 0 ? 1 : (p == 1 ? 3 : 4), q2 = p == 0 ? 2 : (p == 1 ? 4 : 5);
        u0 = ai[q0] - aj[q0]; u1 = ai[q1] - aj[q1]; u2 = ai[q2] - aj[q2];
      } else {
        const double* ai = W + L::AM + 3 * (i * N + d);
        const double* aj = W + L::AM + 3 * (j * N + d);
        u0 = ai[0] - aj[0]; u1 = ai[1] - aj[1]; u2 = ai[2] - aj[2];
      }
      const double ru = 3.0 * (r0 * u0 + r1 * u1 + r2 * u2);
      double e0 = fac1 * (r0 * ru - rsq * u0), e1 = fac1 * (r1 * ru - rsq * u1), e2 = fac1 * (r2 * ru - rsq * u2);
      if (p < 3) {
        if (d == i) { e0 += R[PF_RM + p]; e1 += R[PF_RM + 3 + p]; e2 += R[PF_RM + 6 + p]; }
        else if (d == j) { e0 -= R[PF_RM + p]; e1 -= R[PF_RM + 3 + p]; e2 -= R[PF_RM + 6 + p]; }
        acc0 = fma(mj, e0, acc0); acc1 = fma(mj, e1, acc1); acc2 = fma(mj, e2, acc2);
      } else {
        if (d == i || d == j) { const double us = sg * R[PF_US]; e0 = fma(us, r0, e0); e1 = fma(us, r1, e1); e2 = fma(us, r2, e2); }
        acc0 = fma(mj, e0, acc0); acc1 = fma(mj, e1, acc1); acc2 = fma(mj, e2, acc2);
        if (d == j) { acc0 = fma(sg, R[PF_F], acc0); acc1 = fma(sg, R[PF_F + 1], acc1); acc2 = fma(sg, R[PF_F + 2], acc2); }
      }
    }
    if (p < 3) {
      const int hf = d >= L::NA ? 1 : 0, col = 3 * (d - hf * L::NA) + p;
      double* o = W + L::PHX + (hf * 3 * N + 3 * i) * L::KIN + col;
      o[0] = acc0; o[L::KIN] = acc1; o[2 * L::KIN] = acc2;
    } else {
      double* o = W + L::PHM + (3 * i) * N + d;
      o[0] = acc0; o[N] = acc1; o[2 * N] = acc2;
    }
  }
  __syncthreads();
}

// jac_step (+)= jac_phi * jac_step with the dense blocks of rx_phi_assemble; jv/je at offset OFF (position p holds
// body (OFF + p) mod N).  Each half multiplies its KIN inputs into all 3N outputs, the v half adds the two partial sums.
template <int N, int OFF> __device__ __forceinline__ void rx_phisalpha_dense(RxState<N>& S, const double* __restrict__ W, int half, int c) {
  using L = RxPhi<N>;
  double in[L::KIN];
  static_for<0, L::NA>([&](auto Qc) {
    constexpr int q = decltype(Qc)::value;
    constexpr int pa = (q - OFF + 2 * N) % N;              // position of body q (x half's input)
    constexpr int bb = L::NA + q;                          // body whose x rows the v half receives
    if constexpr (bb < N) {
      constexpr int pb = (bb - OFF + 2 * N) % N;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double got = shx(S.jv[pb][k]);
        in[3 * q + k] = half ? got : S.jv[pa][k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) in[3 * q + k] = half ? 0.0 : S.jv[pa][k];
    }
  });
  // mass column of body dm (or -1): the v half adds Phi_m[:, dm]
  const int dm = (c % 7 == 6 && c < 7 * N) ? c / 7 : -1;
  const double* __restrict__ Wh = W + L::PHX + half * 3 * N * L::KIN;
  static_for<0, N>([&](auto Bc) {
    constexpr int b = decltype(Bc)::value;
    constexpr int pos = (b - OFF + 2 * N) % N;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* __restrict__ row = Wh + (3 * b + k) * L::KIN;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int t = 0; t + 1 < L::KIN; t += 2) { s0 = fma(row[t], in[t], s0); s1 = fma(row[t + 1], in[t + 1], s1); }
      if (L::KIN & 1) s0 = fma(row[L::KIN - 1], in[L::KIN - 1], s0);
      double part = s0 + s1;
      const double other = shx(part);
      double dv = part + other;
      if (dm >= 0) dv += W[L::PHM + (3 * b + k) * N + dm];
      // comp_sum_matrix!(jac_step, jac_error, jac_phi * jac_step): v rows get dv, x rows a zero addend (fold)
      ksum_m(S.jv[pos][k], S.je[pos][k], half == 1 ? dv : 0.0);
    }
  });
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

// one AHL21 Jacobian step from a staged operator block; offset 0 (identity) on entry and exit.
// scr: scratch for the dense phisalpha blocks (the block's own ascending records when RxPhi<N>::ALIAS).
template <int N, int U>
__device__ __forceinline__ void rx_step(RxState<N>& S, double* __restrict__ blk, double* __restrict__ scr, double h2, int half, int c, int tid, int nthr) {
  constexpr int P = N * (N - 1) / 2;
  using SW = RxSweep<N, U>;
  auto rotate = [&](auto Kc) { rot_left_by<N, decltype(Kc)::value>(S.jv); rot_left_by<N, decltype(Kc)::value>(S.je); };
  auto pair = [&](auto PA, auto PB, const double* R, int bi, int bj) {
    rx_pair<N, decltype(PA)::value, decltype(PB)::value>(S, R, half, c, 7 * bi + 6, 7 * bj + 6);
  };
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
  SW::asc(rotate, blk, KF, pair);                                          // offset 0 -> A1; ends with a block barrier
  rx_phi_assemble<N>(blk + 2 * P * KF, scr, tid, nthr);
  rx_phisalpha_dense<N, SW::A1>(S, scr, half, c);                          // A1 -> A1
  rotate(std::integral_constant<int, SW::KTOP * U - SW::A1>{});            // A1 -> KTOP*U
  SW::desc(rotate, blk + P * KF, KF, pair);                                // -> 0
  rx_drift<N>(S, h2, half);
  rx_fold<N>(S);
}

}  // namespace nbg
