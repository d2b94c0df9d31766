// Jacobian propagation with the pair update on the FP64 tensor-core path (DMMA, mma.sync.m8n8k4.f64).
//
// Why a second register-resident kernel (nbg_jacobian_rx.cuh is the first): in the column-per-lane-pair layout the 6x6 Kepler block
// K of a pair reaches the FMAs as BROADCAST shared-memory operands -- 10 LDS.128 and 6 SHFL per pair and warp, 26 cycles of the
// shared-memory/shuffle pipe next to 22 cycles of FP64 issue, and with two warps per scheduler the two do not overlap (ncu r01h:
// FP64 pipe 48 %, shared-memory pipe 59 %).  DMMA issues at exactly the DFMA rate (profiles/microbench/r01_dmma_probe.txt), so it
// does not add arithmetic throughput; what it changes is operand delivery: the B fragment is ONE element per lane, so K is fetched
// with two LDS.64 (distinct addresses) per pair and warp, and the x/v contraction happens inside the instruction -- no shuffles.
//
// Layout.  A warp owns two tiles of 8 columns of jac_step.  Lane l = 4 g + t holds, for column 8 T + g and every body b, the pair
// (x_t, v_t) -- t = 0, 1, 2; lanes with t = 3 are padding -- and the same for jac_error.  With W^T = D^T K^T (D = J_i - J_j, 6 x 8):
//   A fragment (8 x 4, lane holds A[g][t])       = D^T : k-step 0 takes the x components, k-step 1 the v components: the lane's own
//                                                   two registers, no data movement;
//   B fragment (4 x 8, lane holds B[t][g'])       = K[o(g')][t] resp. K[o(g')][3 + t] with o(g') = 3 (g' & 1) + (g' >> 1);
//   C fragment (8 x 8, lane holds C[g][2t, 2t+1]) = (W[x_t], W[v_t]) of column g: exactly the storage layout, so the Kahan update
//                                                   of J_i and J_j is elementwise in registers.
// The padding (6 of 8 outputs, 3 of 4 k slots) makes a pair cost 4 DMMA = 32 DFMA-equivalents per warp instead of 18 DFMA per thread,
// i.e. more FP64 pipe time for the product, but the pipe no longer waits for operands.  drift_grad! becomes lane-local (x_t and v_t
// sit in one lane).  The dense phisalpha operator is applied the same way: A = x components of body d, B = the weights of two
// output bodies interleaved (even n -> body 2 bp, odd n -> body 2 bp + 1), so C again lands in the owning lane.
// Replaces the same reference code as nbg_jacobian_rx.cuh (ahl21.jl:5-95 Jacobian half, timing.jl:155-194).
//
// STATUS: experiment, off by default (NBG_JAC_MMA=1: two tiles per warp, 250 registers; =2: one tile per warp, 128 registers, 14
// warps per SM).  Parity-green at 1e-11 on the N = 8 tests (r01l), but SLOWER than jac_rx_kernel on B200: 798 ms / 757 ms against
// 675 ms per 3 bench steps.  ncu (profiles/r01l_jac_mma_kernel.txt): the top stall is the fixed-latency wait behind each DMMA, the
// FP64 pipe is the only busy unit (shared-memory pipe 18 %), i.e. the padded product (4 DMMA = 32 DFMA-equivalents per pair and warp
// against 18 DFMA per thread) costs more pipe time than the operand delivery it saves.  Kept as measured evidence for DESIGN.md 4.
#pragma once
#include "nbg_jacobian_rx.cuh"

namespace nbg {

__host__ __device__ constexpr int mma_tiles(int n) { return (7 * n + 7) / 8; }
__host__ __device__ constexpr int mma_warps(int n, int tpw = 2) { return (mma_tiles(n) + tpw - 1) / tpw; }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int N, int TPW> struct MmaState {  // TPW: 8-column tiles per warp (2: fewer operand loads per column; 1: twice the warps per SM)
  double jv[TPW][N][2];  // [tile][body][0: x_t, 1: v_t]
  double je[TPW][N][2];
};

struct MmaLane {
  int g, t;        // column within the tile, component
  int c[2];        // global column of each tile
  int dm[2];       // body whose mass column c[T] is, else -1
  int koff0, koff1;  // offsets of this lane's B elements inside a Kepler record (k-step 0: x inputs, 1: v inputs)
  int woff;        // offset of this lane's B element inside the dense phisalpha operator
  bool bvalid;     // g < 6 && t < 3: the lane holds a real B element
  bool odd;        // g & 1
};

__device__ __forceinline__ MmaLane mma_lane(int lane, int warp, int n, int tpw) {
  MmaLane L;
  L.g = lane >> 2; L.t = lane & 3;
  const int M = 7 * n;
#pragma unroll
  for (int T = 0; T < 2; ++T) {
    L.c[T] = (tpw * warp + (T < tpw ? T : 0)) * 8 + L.g;
    L.dm[T] = (L.c[T] < M && L.c[T] % 7 == 6) ? L.c[T] / 7 : -1;
  }
  L.odd = (L.g & 1) != 0;
  L.bvalid = L.g < 6 && L.t < 3;
  const int o = 3 * (L.g & 1) + (L.g >> 1), tt = L.t < 3 ? L.t : 0, oo = o < 6 ? o : 0;
  L.koff0 = kf_k(oo, tt);
  L.koff1 = kf_k(oo, 3 + tt);
  L.woff = ((3 * (L.g & 1) + (L.g >> 1)) * n) * 4 + tt;
  return L;
}

// pair update between bodies PA < PB (comp_sum_matrix! of jac_ij * rows(i,j), ahl21.jl:31-35, 64-68)
template <int N, int TPW, int PA, int PB>
__device__ __forceinline__ void mma_pair(MmaState<N, TPW>& S, const double* __restrict__ R, const MmaLane& L) {
  const double b0 = L.bvalid ? R[L.koff0] : 0.0;
  const double b1 = L.bvalid ? R[L.koff1] : 0.0;
  const double2 mm = *reinterpret_cast<const double2*>(R + KF_MI);
  constexpr int ci = 7 * PA + 6, cj = 7 * PB + 6;
#pragma unroll
  for (int T = 0; T < TPW; ++T) {
    const double d0 = S.jv[T][PA][0] - S.jv[T][PB][0];
    const double d1 = S.jv[T][PA][1] - S.jv[T][PB][1];
    double w0 = 0.0, w1 = 0.0;
    dmma884(w0, w1, d0, b0);
    dmma884(w0, w1, d1, b1);
    // comp_sum_matrix! (utils.jl:36-46) with the scaling by the mass fractions folded into its first addition
    double ei0 = fma(mm.y, w0, S.je[T][PA][0]), ei1 = fma(mm.y, w1, S.je[T][PA][1]);
    double ej0 = fma(-mm.x, w0, S.je[T][PB][0]), ej1 = fma(-mm.x, w1, S.je[T][PB][1]);
    if ((L.c[T] == ci || L.c[T] == cj) && L.t < 3) {  // mass columns of the two bodies: rank-one terms of jac_ij (ahl21.jl:743-750)
      const double* mb = R + 38 + (L.c[T] == ci ? 0 : 6) + L.t;
      ei0 += mb[0]; ej0 += mb[3]; ei1 += mb[12]; ej1 += mb[15];
    }
    const double ti0 = __dadd_rn(S.jv[T][PA][0], ei0);
    S.je[T][PA][0] = __dadd_rn(ei0, __dsub_rn(S.jv[T][PA][0], ti0));
    S.jv[T][PA][0] = ti0;
    const double ti1 = __dadd_rn(S.jv[T][PA][1], ei1);
    S.je[T][PA][1] = __dadd_rn(ei1, __dsub_rn(S.jv[T][PA][1], ti1));
    S.jv[T][PA][1] = ti1;
    const double tj0 = __dadd_rn(S.jv[T][PB][0], ej0);
    S.je[T][PB][0] = __dadd_rn(ej0, __dsub_rn(S.jv[T][PB][0], tj0));
    S.jv[T][PB][0] = tj0;
    const double tj1 = __dadd_rn(S.jv[T][PB][1], ej1);
    S.je[T][PB][1] = __dadd_rn(ej1, __dsub_rn(S.jv[T][PB][1], tj1));
    S.jv[T][PB][1] = tj1;
  }
}

// drift_grad! (ahl21.jl:318-331): x rows += h/2 * v rows, Kahan; lane-local in this layout
template <int N, int TPW> __device__ __forceinline__ void mma_drift(MmaState<N, TPW>& S, double h2) {
#pragma unroll
  for (int T = 0; T < TPW; ++T)
#pragma unroll
    for (int b = 0; b < N; ++b) ksum(S.jv[T][b][0], S.je[T][b][0], h2 * S.jv[T][b][1]);
}
// comp_sum_matrix! with a zero addend (jac_kick = 0, ahl21.jl:23,93): folds jac_error into jac_step
template <int N, int TPW> __device__ __forceinline__ void mma_fold(MmaState<N, TPW>& S) {
#pragma unroll
  for (int T = 0; T < TPW; ++T)
#pragma unroll
    for (int b = 0; b < N; ++b) { ksum_m(S.jv[T][b][0], S.je[T][b][0], 0.0); ksum_m(S.jv[T][b][1], S.je[T][b][1], 0.0); }
}

// jac_step (+)= jac_phi * jac_step with the dense operator W (layout of phi_dense_fields) in shared memory
template <int N, int TPW> __device__ __forceinline__ void mma_phisalpha(MmaState<N, TPW>& S, const double* __restrict__ W, const MmaLane& L) {
  constexpr int NP = (N + 1) / 2;
  double dv[TPW][2 * NP];  // [tile][body]: this lane's component t of the v-row increment
  static_for<0, NP>([&](auto BPc) {
    constexpr int bp = decltype(BPc)::value;
    const bool valid = L.bvalid && (2 * bp + 1 < N || !L.odd);
    double c[TPW][2];
#pragma unroll
    for (int T = 0; T < TPW; ++T) { c[T][0] = 0.0; c[T][1] = 0.0; }
    static_for<0, N>([&](auto Dc) {
      constexpr int d = decltype(Dc)::value;
      const double bw = valid ? W[L.woff + (6 * bp * N + d) * 4] : 0.0;
#pragma unroll
      for (int T = 0; T < TPW; ++T) dmma884(c[T][0], c[T][1], S.jv[T][d][0], bw);
    });
#pragma unroll
    for (int T = 0; T < TPW; ++T) { dv[T][2 * bp] = c[T][0]; dv[T][2 * bp + 1] = c[T][1]; }
  });
#pragma unroll
  for (int T = 0; T < TPW; ++T)
#pragma unroll
    for (int b = 0; b < N; ++b) {
      double a = dv[T][b];
      if (L.dm[T] >= 0 && L.t < 3) a += W[((3 * b + L.t) * N + L.dm[T]) * 4 + 3];
      // comp_sum_matrix!(jac_step, jac_error, jac_phi * jac_step): v rows get the increment, x rows a zero addend (fold)
      ksum_m(S.jv[T][b][1], S.je[T][b][1], a);
      ksum_m(S.jv[T][b][0], S.je[T][b][0], 0.0);
    }
}

// one AHL21 Jacobian step (no fast-kick pairs) from a staged operator block [2P Kepler records | dense phisalpha operator]
template <int N, int TPW> __device__ __forceinline__ void mma_step(MmaState<N, TPW>& S, const double* __restrict__ blk, double h2, const MmaLane& L) {
  constexpr int P = N * (N - 1) / 2;
  using SW = RxSweep<N, N, false>;  // full unroll: positions are bodies, no rotation
  auto rotate = [&](auto Kc) { static_assert(decltype(Kc)::value % N == 0, "full unroll never rotates"); };
  auto pair = [&](auto PA, auto PB, const double* R, int, int) { mma_pair<N, TPW, decltype(PA)::value, decltype(PB)::value>(S, R, L); };
  mma_drift<N, TPW>(S, h2);
  mma_fold<N, TPW>(S);
  SW::asc(rotate, blk, KF, pair);
  mma_phisalpha<N, TPW>(S, blk + 2 * P * KF, L);
  rotate(std::integral_constant<int, SW::KTOP * N - SW::A1>{});
  SW::desc(rotate, blk + P * KF, KF, pair);
  mma_drift<N, TPW>(S, h2);
  mma_fold<N, TPW>(S);
}

}  // namespace nbg
