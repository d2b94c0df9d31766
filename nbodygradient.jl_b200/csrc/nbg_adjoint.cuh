// Transit-time gradients through the ADJOINT of the transit sub-step.
//
// The reference (findtransit!, timing.jl:75-110) takes one more full gradient step of size dt0 from the state before the transit --
// jac_step' = T jac_step, all 7N columns -- and then reads a handful of rows of jac_step' (dtbvdq!, timing.jl:155-194):
//     dt/dq0[c] = - w^T (T J)[:, c] / gdot,      w = (dv_x, dv_y on the x rows, dx, dy on the v rows) of occultor minus transited body,
// after which it restores jac_step (set_state!(s, s_prior)).  T J is never needed as a matrix: w^T (T J) = (T^T w)^T J.  One thread per
// queued transit applies the TRANSPOSED operators of the sub-step, in reverse order, to the 7N-vector w (a few thousand flops instead of
// a full Jacobian step of 1.5 Mflop), and the Jacobian kernel, which holds J in registers at that moment, finishes with one dot product
// per column.  The Jacobian kernel therefore never applies a transit step, never saves / restores its matrix (196 KB of HBM traffic per
// transit before), and no longer reads the transits' operator stream.
//
// Forward order of one AHL21 Jacobian step (rx_step, nbg_jacobian_rx.cuh; ahl21.jl:5-95), every factor a LEFT multiplication of J:
//   [K0: first kickfast!, formed BEFORE the drift and added after it]  D(h/2)  asc pairs  PHI (phic!+phisalpha!)  desc pairs  D(h/2)  [K2]
// z = (zx[3N], zv[3N], zm[N]) lives on (x rows, v rows, mass rows); mass rows of J are unit rows, so zm adds to the mass columns.
//   D^T:     zv += h2 zx
//   pair^T:  u = mj' z_i - mi' z_j (6-vector, mass fractions of the record), g = K^T u:  z_i += g, z_j -= g,
//            zm_i += ci7 . z_i + cj7 . z_j,  zm_j += ci14 . z_i + cj14 . z_j        (old z_i, z_j)
//   W^T:     zx_d += sum_ik W[i,k,d,0..2] zv_ik,   zm_d += sum_ik W[i,k,d,3] zv_ik       (dense 3N x 4N operator of phi_dense_kernel)
//   K0 with quirk B-3 (x' = x + h2 v, v' = v + W0 x_old):  zx' = zx + W0x^T zv, zv' = zv + h2 zx, zm' = zm + W0m^T zv  (simultaneous)
// Rounding: the same bilinear form summed in another order; agreement with the reference's forward evaluation is at the 1e-15 level
// relative to the terms (parity tests at 1e-11 unchanged); the Jacobian kernels evaluate the final dot product in difference form
// (rows of J minus the rows of the transited body) -- see rx_transit_out in nbg_b200.cu.
#pragma once
#include "nbg_jacobian_rx.cuh"

namespace nbg {

// z layout per transit: [comp][7N] = [zx (3N) | zv (3N) | zm (N)], comp < C
__host__ __device__ inline size_t zfields(int n, int C) { return (size_t)C * 7 * n; }

template <int NB>  // NB = array bound (NMAX); n = run-time body count
struct AdjVec {
  double zx[3 * NB], zv[3 * NB], zm[NB];
};

// transposed Kepler-pair operator on C vectors at once (the record is loaded once)
template <int NC>
__device__ __forceinline__ void adj_pair(AdjVec<NMAX> (&Z)[NC], const Src& S, size_t rec0, int i, int j) {
  double K[36], mc[24];
#pragma unroll
  for (int g = 0; g < 9; ++g) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(S.base + (((rec0 >> 2) + g) * S.stride + S.idx) * 4));
    const double2 b = __ldg(reinterpret_cast<const double2*>(S.base + (((rec0 >> 2) + g) * S.stride + S.idx) * 4) + 1);
    K[4 * g] = a.x; K[4 * g + 1] = a.y; K[4 * g + 2] = b.x; K[4 * g + 3] = b.y;
  }
  const double mif = S.get(rec0 + KF_MI), mjf = S.get(rec0 + KF_MJ);
  if (mif == 0.0 && mjf == 0.0) return;  // two massless bodies: identity (zero record)
#pragma unroll
  for (int f = 0; f < 24; ++f) mc[f] = S.get(rec0 + 38 + f);
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    AdjVec<NMAX>& z = Z[q];
    double zi[6], zj[6], u[6], g6[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) { zi[k] = z.zx[3 * i + k]; zi[3 + k] = z.zv[3 * i + k]; zj[k] = z.zx[3 * j + k]; zj[3 + k] = z.zv[3 * j + k]; }
#pragma unroll
    for (int r = 0; r < 6; ++r) u[r] = mjf * zi[r] - mif * zj[r];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) s = fma(K[kf_k(r, c)], u[r], s);
      g6[c] = s;
    }
    // mass columns (ahl21.jl:743-750): rows i / j, columns m_i ("7") and m_j ("14")
    double a7 = 0.0, a14 = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      a7 = fma(mc[kf_ci7(r) - 38], zi[r], fma(mc[kf_cj7(r) - 38], zj[r], a7));
      a14 = fma(mc[kf_ci14(r) - 38], zi[r], fma(mc[kf_cj14(r) - 38], zj[r], a14));
    }
    z.zm[i] += a7;
    z.zm[j] += a14;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      z.zx[3 * i + k] = zi[k] + g6[k]; z.zv[3 * i + k] = zi[3 + k] + g6[3 + k];
      z.zx[3 * j + k] = zj[k] - g6[k]; z.zv[3 * j + k] = zj[3 + k] - g6[3 + k];
    }
  }
}

// transposed dense v-row operator W (phi_dense_fields layout at field offset d0): zx += Wx^T zv, zm += Wm^T zv.
// simultaneous = true: the first kickfast! of a step with fast-kick pairs (see the header): the drift transposes against the OLD zx.
template <int NC>
__device__ __forceinline__ void adj_dense(AdjVec<NMAX> (&Z)[NC], const Src& S, size_t d0, int n, double h2_simultaneous, bool simultaneous) {
  if (simultaneous) {
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      // zv' = zv + h2 zx (old zx) must not see the W^T update of zx, and zx' = zx + W^T zv (old zv) must not see the drift: stash old zv
      // in place by applying the W^T update first from old zv, then the drift from the stashed old zx
      AdjVec<NMAX>& z = Z[q];
      double oldx[3 * NMAX];
      for (int r = 0; r < 3 * n; ++r) oldx[r] = z.zx[r];
      for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
          const double zvk = z.zv[3 * i + k];
          for (int d = 0; d < n; ++d) {
            const size_t f = d0 + (size_t)((3 * i + k) * n + d) * 4;
            z.zx[3 * d] = fma(S.get(f), zvk, z.zx[3 * d]); z.zx[3 * d + 1] = fma(S.get(f + 1), zvk, z.zx[3 * d + 1]);
            z.zx[3 * d + 2] = fma(S.get(f + 2), zvk, z.zx[3 * d + 2]); z.zm[d] = fma(S.get(f + 3), zvk, z.zm[d]);
          }
        }
      for (int r = 0; r < 3 * n; ++r) z.zv[r] = fma(h2_simultaneous, oldx[r], z.zv[r]);
    }
    return;
  }
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k)
      for (int d = 0; d < n; ++d) {
        const size_t f = d0 + (size_t)((3 * i + k) * n + d) * 4;
        const double2 a = __ldg(reinterpret_cast<const double2*>(S.base + ((f >> 2) * S.stride + S.idx) * 4));
        const double2 b = __ldg(reinterpret_cast<const double2*>(S.base + ((f >> 2) * S.stride + S.idx) * 4) + 1);
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          AdjVec<NMAX>& z = Z[q];
          const double zvk = z.zv[3 * i + k];
          z.zx[3 * d] = fma(a.x, zvk, z.zx[3 * d]); z.zx[3 * d + 1] = fma(a.y, zvk, z.zx[3 * d + 1]);
          z.zx[3 * d + 2] = fma(b.x, zvk, z.zx[3 * d + 2]); z.zm[d] = fma(b.y, zvk, z.zm[d]);
        }
      }
}

// z = T^T w for the NC output components of one queued transit.  blk: the transit's operator block in the tiled stream.
template <int NC>
__device__ __forceinline__ void adjoint_step(AdjVec<NMAX> (&Z)[NC], const Src& S, int n, double h2, const KMask& kmask) {
  const int P = npairs(n);
  const bool kicks = kmask.any();
  auto drift_t = [&]() {
#pragma unroll
    for (int q = 0; q < NC; ++q)
      for (int r = 0; r < 3 * n; ++r) Z[q].zv[r] = fma(h2, Z[q].zx[r], Z[q].zv[r]);
  };
  // reverse of: [K0] D asc PHI desc D [K2]
  if (kicks) adj_dense<NC>(Z, S, phi_dense_offset(n, true, 2), n, 0.0, false);   // second kickfast!
  drift_t();
  {  // descending sweep reversed: the forward order is i = n-2..0, j = n-1..i+1 with records P, P+1, ...
    int rec = 2 * P - 1;
    for (int i = 0; i <= n - 2; ++i)
      for (int j = i + 1; j <= n - 1; ++j, --rec)
        if (!kicks || !kmask.bit(rx_pair_index(n, i, j))) adj_pair<NC>(Z, S, (size_t)rec * KF, i, j);
  }
  adj_dense<NC>(Z, S, phi_dense_offset(n, kicks, kicks ? 1 : 0), n, 0.0, false);   // phic! + phisalpha!
  {  // ascending sweep reversed
    int rec = P - 1;
    for (int i = n - 2; i >= 0; --i)
      for (int j = n - 1; j >= i + 1; --j, --rec)
        if (!kicks || !kmask.bit(rx_pair_index(n, i, j))) adj_pair<NC>(Z, S, (size_t)rec * KF, i, j);
  }
  if (kicks) adj_dense<NC>(Z, S, phi_dense_offset(n, true, 0), n, h2, true);       // first kickfast! + drift (quirk B-3)
  else drift_t();
}

}  // namespace nbg
