// libnbgrad_b200.so — kernels and C ABI (include/nbgrad.h).  sm_100a only; no CPU fallback.
//
// Pipeline per chunk of S steps:
//   traj_kernel       one THREAD per system: x, v (Kahan); detects transits (detect_transits!, timing.jl:3-29) and queues them with a
//                     snapshot of the state; main-loop steps leave only the scalars of each Kepler solve (split path)
//   pair_op_kernel    one thread per (system, step, pair section): compute_jacobian_gamma! -> Kepler operator records
//   phi_dense_kernel  dense phisalpha operator of every step
//   transit_kernel    one THREAD per queued transit: findtransit! Newton iterations (timing.jl:31-73), Jacobian-free because x, v,
//                     dqdt never read jac_step; then the one final step (timing.jl:75-80) whose operator records are written
//   jac_rx_kernel     one block per system, jac_step + jac_error in REGISTERS for the whole chunk (two lanes per column), operator
//                     block of a step staged in shared memory; at each queued transit one dot product per column with the adjoint
//                     vectors of the transit sub-step (transit_adjoint_kernel, nbg_adjoint.cuh) gives dtbvdq! (timing.jl:155-194) or
//                     the fused chi^2 gradient.  All N = 2..16; with fast-kick pairs the KICK variant (three dense operators per step).
// The host loop (run_steps) reads the number of queued transits back after the trajectory kernel of every chunk: a queue that is too
// small is grown and the chunk re-run from a saved trajectory state (no transit is ever dropped), and the per-transit buffers are sized
// from the actual count.  One-shot calls stream every chunk's transit rows to the caller's host arrays while the next chunk computes.
// Data layout in HBM: trajectory state SoA with the system index fastest (lanes = systems); operator stream tiled by 32 systems;
// jac_step/jac_error [sys][row 6N][col 7N].
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "../../include/nbgrad.h"
#include "nbg_jacobian_rx.cuh"
#include "nbg_adjoint.cuh"
#ifdef NBG_EXPERIMENTS
#include "nbg_jacobian_mma.cuh"   // DMMA Jacobian kernel: measured 12-18 % slower, kept as evidence
#endif
#include "nbg_ics.cuh"

using namespace nbg;

#ifndef NBG_U8
#define NBG_U8 4   // pivot bodies per block of the pair sweeps of the N = 8 Jacobian kernel (8 = full unroll); -DNBG_U8=8 / 2 for A/B builds
#endif

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(NBG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct TrajArrays {
  double *x, *v, *xe, *ve, *m, *dq, *gsave, *t;
  int32_t* count;
  uint32_t* status;
  size_t ld;  // padded system count (leading dimension of every SoA array)
  // CartesianOutput-style sampling (Outputs.jl:26-49): x, v BEFORE every samp_stride-th step, [sample][3n][ld]; null = off
  double *samp_x = nullptr, *samp_v = nullptr;
  long samp_stride = 1;
};

struct EventQueue {
  int32_t* n;  // number queued this chunk (may exceed cap: the host then grows the queue and re-runs the chunk)
  int32_t cap;
  int32_t *sys, *step, *body, *k;
  double *dt0, *t;   // initial guess / time of the prior state
  double* snap;      // [12N][cap]  x, v, xe, ve
  double* hdr;       // [HDR][cap]  dx, dy, dvx, dvy, 1/gdot, 1/vsky, dvdt, dt0_final, chi^2 weight, chi^2 term  (written by transit_kernel)
  double* stream;    // [step_fields][nq] operator stream of the final step (sized from the actual count): read by transit_adjoint_kernel
  double* z;         // [nq][C][7N] adjoint vectors of the transit sub-step (nbg_adjoint.cuh): d out_c / d q0 = z_c^T jac_step
};
constexpr int HDR = 10;

struct TransitOut {
  double* tt;      // dense: [sys][RT][C]; event rows: [slot][C]; null = not wanted
  double* dtdq0;   // dense: [sys][RT][M][C]; event rows: [slot][M][C]; null = not wanted (fused chi^2)
  const int32_t* ntt_body;  // device [N]
  const int32_t* off;       // device [N]
  int32_t RT, C;            // C = 1 (TransitTiming) or 3 (TransitParameters)
  int32_t ev_rows;          // 1: the output row of a transit is its queue slot (per-chunk compact buffers, scattered by the host)
  // fused transit-time likelihood (nbg_transit_chi2_fused): observations and per-system accumulators; null = off
  const double* tobs; const double* sigma; int32_t per_system;
  double* chi2;    // [sys]
  double* gq;      // [sys][M]
};
__device__ __forceinline__ size_t out_rec(const TransitOut& O, long sys, int body, int k, int slot) {
  return O.ev_rows ? (size_t)slot : (size_t)sys * O.RT + O.off[body] + k;
}

// ------------------------------------------------------------------------------------------------------------------
// Launch bounds: the light (split-path) instantiation is capped at 128 registers so that 4 blocks fit an SM: a thread owns
// a system for the whole chunk, so 65,536 systems must all be resident at once (148 SMs x 512 threads) or the kernel pays
// a nearly empty second wave.
template <bool GRAD, int EMIT, bool KICKS = false>
__global__ void __launch_bounds__(128, (EMIT == 2 && !GRAD) ? 4 : 1) traj_kernel(TrajArrays T, int n, long nsys, double h, int nsteps, double* stream, double* scal, int detect, int ti,
                                                   double t0, long istep0, double h_intr, const int32_t* ntt_body, EventQueue Q,
                                                   int32_t* evlist, uint32_t* evmask, int time_mode_kahan, double* tkahan_err, KMask kmask) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  const size_t ld = T.ld;
  Body b;
  double dq[6 * NMAX];
  for (int q = 0; q < 3 * n; ++q) {
    b.x[q] = T.x[q * ld + sys]; b.v[q] = T.v[q * ld + sys]; b.xe[q] = T.xe[q * ld + sys]; b.ve[q] = T.ve[q * ld + sys];
  }
  for (int q = 0; q < n; ++q) b.m[q] = T.m[q * ld + sys];
  for (int q = 0; q < 6 * n; ++q) dq[q] = GRAD ? T.dq[q * ld + sys] : 0.0;
  double gs[NMAX];
  int32_t cnt[NMAX];
  if (detect) for (int i = 0; i < n; ++i) { gs[i] = T.gsave[i * ld + sys]; cnt[i] = T.count[i * ld + sys]; }
  double tnow = T.t[sys], terr = tkahan_err ? tkahan_err[sys] : 0.0;
  uint32_t st = 0;
  const size_t sf = step_fields(n, kmask.any());
  for (int s = 0; s < nsteps; ++s) {
    if (T.samp_x && (istep0 + s) % T.samp_stride == 0) {  // o.states[i] = deepcopy(s) before the step (Outputs.jl:40)
      const size_t k = (size_t)((istep0 + s) / T.samp_stride);
      for (int q = 0; q < 3 * n; ++q) { T.samp_x[(k * 3 * n + q) * ld + sys] = b.x[q]; T.samp_v[(k * 3 * n + q) * ld + sys] = b.v[q]; }
    }
    Emit em{EMIT ? stream + tile_offset(sf, ld / TILE, (size_t)s, (size_t)sys) : nullptr, TILE, (size_t)(sys % TILE),
            EMIT == 2 ? scal + tile_offset((size_t)2 * npairs(n) * SCF, ld / TILE, (size_t)s, (size_t)sys) : nullptr};
    ahl21_step<GRAD, EMIT, KICKS>(b, dq, n, h, em, kmask);
    if (time_mode_kahan) ksum(tnow, terr, h);                      // (intr)(s,N): Integrator.jl:229
    else tnow = t0 + ((double)(istep0 + s + 1) * h);               // Transits.jl:161
    if (detect) {
      uint32_t mask = 0;
      for (int i = 0; i < n; ++i) {
        int32_t slot = -1;
        if (i != ti) {
          const double gi = gsky(b, i, ti);
          const double ri = sqrt(b.x[3 * i] * b.x[3 * i] + b.x[3 * i + 1] * b.x[3 * i + 1] + b.x[3 * i + 2] * b.x[3 * i + 2]);
          if (gi > 0.0 && gs[i] < 0.0 && -b.x[3 * i + 2] > 0.25 * ri && ri < 1e12) {
            cnt[i] += 1;
            if (cnt[i] <= ntt_body[i]) {
              slot = atomicAdd(Q.n, 1);
              if (slot < Q.cap) {
                Q.sys[slot] = (int32_t)sys; Q.step[slot] = s; Q.body[slot] = i; Q.k[slot] = cnt[i] - 1;
                Q.dt0[slot] = -gi * h_intr / (gi - gs[i]);
                Q.t[slot] = tnow;
                for (int q = 0; q < 3 * n; ++q) {
                  Q.snap[(size_t)(q)*Q.cap + slot] = b.x[q];
                  Q.snap[(size_t)(3 * n + q) * Q.cap + slot] = b.v[q];
                  Q.snap[(size_t)(6 * n + q) * Q.cap + slot] = b.xe[q];
                  Q.snap[(size_t)(9 * n + q) * Q.cap + slot] = b.ve[q];
                }
              } else {
                slot = -1;   // the host sees *Q.n > cap, grows the queue and re-runs the chunk (run_steps)
              }
            } else {
              st |= NBG_ST_NTT_OVERFLOW;
            }
          }
          gs[i] = gi;
        }
        if (evlist) evlist[((size_t)s * n + i) * ld + sys] = slot;
        if (slot >= 0) mask |= 1u << i;
      }
      if (evmask) evmask[(size_t)s * ld + sys] = mask;
    }
  }
  bool finite = true;
  for (int q = 0; q < 3 * n; ++q) {
    T.x[q * ld + sys] = b.x[q]; T.v[q * ld + sys] = b.v[q]; T.xe[q * ld + sys] = b.xe[q]; T.ve[q * ld + sys] = b.ve[q];
    finite = finite && isfinite(b.x[q]) && isfinite(b.v[q]);
  }
  if (GRAD) for (int q = 0; q < 6 * n; ++q) T.dq[q * ld + sys] = dq[q];
  if (detect) for (int i = 0; i < n; ++i) { T.gsave[i * ld + sys] = gs[i]; T.count[i * ld + sys] = cnt[i]; }
  T.t[sys] = tnow;
  if (tkahan_err) tkahan_err[sys] = terr;
  if (!finite) st |= NBG_ST_NONFINITE;
  if (st) atomicOr(&T.status[sys], st);
}

// initial gsave: Transits.jl:146-149
__global__ void gsave_init_kernel(TrajArrays T, int n, long nsys, int ti) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  const size_t ld = T.ld;
  for (int i = 0; i < n; ++i) {
    double g = 0.0;
    if (i != ti)
      g = (T.x[(3 * ti) * ld + sys] - T.x[(3 * i) * ld + sys]) * (T.v[(3 * ti) * ld + sys] - T.v[(3 * i) * ld + sys]) +
          (T.x[(3 * ti + 1) * ld + sys] - T.x[(3 * i + 1) * ld + sys]) * (T.v[(3 * ti + 1) * ld + sys] - T.v[(3 * i + 1) * ld + sys]);
    T.gsave[i * ld + sys] = g;
    T.count[i * ld + sys] = 0;
  }
}

// One gradient-free AHL21 step followed by the Newton correction -g / (dg/dt along the flow): kept out of line so that the
// reference-form iterations of transit_kernel compile exactly as they do without it.
template <bool KICKS>
__device__ __noinline__ double transit_pre_iteration(const Body& b0, int n, double dt0, int ti, int j, const KMask& kmask) {
  Body b = b0;
  Emit none{nullptr, 0, 0};
  ahl21_step<false, 0, KICKS>(b, nullptr, n, dt0, none, kmask);
  return dt0 - gsky(b, ti, j) / gdot_flow(b, n, ti, j);
}

// ------------------------------------------------------------------------------------------------------------------
// findtransit! (timing.jl:31-110).  One thread per queued transit.
// (capping the registers for 3 / 4 blocks per SM was measured: 167 -> 198 / 206 ms per 3 bench steps, the spills cost more than the warps hide)
template <bool GRAD, bool KICKS = false>
__global__ void __launch_bounds__(128) transit_kernel(TrajArrays T, int n, EventQueue Q, int ti, TransitOut O, unsigned long long* counters, KMask kmask, int npre) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = min(*Q.n, Q.cap);
  if (e >= nq) return;
  const int sys = Q.sys[e], j = Q.body[e];
  const size_t ld = T.ld;
  Body b0, b;
  for (int q = 0; q < 3 * n; ++q) {
    b0.x[q] = Q.snap[(size_t)q * Q.cap + e];
    b0.v[q] = Q.snap[(size_t)(3 * n + q) * Q.cap + e];
    b0.xe[q] = Q.snap[(size_t)(6 * n + q) * Q.cap + e];
    b0.ve[q] = Q.snap[(size_t)(9 * n + q) * Q.cap + e];
  }
  for (int q = 0; q < n; ++q) b0.m[q] = T.m[q * ld + sys];
  double dq[6 * NMAX];
  double dt0 = Q.dt0[e], stmp = 0.0;
  Emit none{nullptr, 0, 0};
  // Better starting guess for the reference's Newton iteration: `npre` gradient-free iterations (no pair Jacobians, no dq/dh;
  // derivative of g along the exact flow).  The reference-form iterations below then start next to their fixed point and
  // stop after one or two passes instead of three or four; the converged dt0 is the same fixed point.
#pragma unroll 1
  for (int pre = 0; pre < npre; ++pre) dt0 = transit_pre_iteration<KICKS>(b0, n, dt0, ti, j, kmask);
  double tt1 = dt0 + 1.0, tt2 = dt0 + 2.0;
  int iter = 0;
  // The final step at the converged dt0 (timing.jl:75-80) is always re-run with record emission.  (When the loop ends with
  // dt0 == tt1 it repeats the last iteration bit for bit and could be skipped by emitting records inside the loop; measured on
  // B200 that is slower -- 72 vs 65 ms per bench step for this kernel with full records, 69 with only the 14 KB of scalars that
  // pair_op_kernel needs -- the iterations, already at 255 registers with spills, get slower by more than the saved step.)
  Emit em{GRAD ? Q.stream + tile_offset(step_fields(n, kmask.any()), 0, 0, (size_t)e) : nullptr, TILE, (size_t)(e % TILE)};
  while (true) {
    tt2 = tt1;
    tt1 = dt0;
    b = b0;
    ahl21_step<true, 0, KICKS>(b, dq, n, dt0, none, kmask);
    const double gs = gsky(b, ti, j);
    const double gd = gdot(b, dq, ti, j);
    const double dt = -gs / gd;
    ksum(dt0, stmp, dt);
    iter += 1;
    if (iter >= 20 || dt0 == tt1 || dt0 == tt2) break;
  }
  uint32_t st = (iter >= 20) ? NBG_ST_TRANSIT_ITMAX : 0u;
  const bool redo = GRAD;
  if (redo) {
    b = b0;
    ahl21_step<true, 1, KICKS>(b, dq, n, dt0, em, kmask);
  }
  const double dx = b.x[3 * j] - b.x[3 * ti], dy = b.x[3 * j + 1] - b.x[3 * ti + 1];
  const double dvx = b.v[3 * j] - b.v[3 * ti], dvy = b.v[3 * j + 1] - b.v[3 * ti + 1];
  const double vsky = sqrt(dvx * dvx + dvy * dvy), bsky2 = dx * dx + dy * dy;
  const size_t rec = out_rec(O, sys, j, Q.k[e], e);
  const double ttv = Q.t[e] + dt0;
  if (O.tt) {
    O.tt[rec * O.C] = ttv;
    if (O.C == 3) { O.tt[rec * 3 + 1] = vsky; O.tt[rec * 3 + 2] = bsky2; }
  }
  // fused likelihood: residual of this transit against its observation slot; masked slots (sigma <= 0, non-finite t_obs) weigh nothing
  double chiw = 0.0, chit = 0.0;
  if (O.chi2) {
    const size_t ob = (O.per_system ? (size_t)sys * O.RT : 0) + O.off[j] + Q.k[e];
    const double sg = O.sigma[ob], to = O.tobs[ob];
    if (sg > 0.0 && isfinite(to)) {
      const double r = (ttv - to) / sg;
      chit = r * r;
      chiw = 2.0 * r / sg;
    }
    if (!GRAD) atomicAdd(&O.chi2[sys], chit);   // no Jacobian kernel follows: sum here
  }
  if (GRAD) {
    const double gd = gdot(b, dq, ti, j);
    const double dvdt = (dvx * (dq[6 * j + 3] - dq[6 * ti + 3]) + dvy * (dq[6 * j + 4] - dq[6 * ti + 4])) / vsky;
    double* H = Q.hdr;
    const size_t cap = Q.cap;
    H[0 * cap + e] = dx; H[1 * cap + e] = dy; H[2 * cap + e] = dvx; H[3 * cap + e] = dvy;
    H[4 * cap + e] = 1.0 / gd; H[5 * cap + e] = 1.0 / vsky; H[6 * cap + e] = dvdt; H[7 * cap + e] = dt0;
    H[8 * cap + e] = chiw; H[9 * cap + e] = chit;
  }
  if (st) atomicOr(&T.status[sys], st);
  atomicAdd(&counters[1], (unsigned long long)iter);
  atomicAdd(&counters[3], 1ull);
  if (redo) atomicAdd(&counters[2], 1ull);
  if (GRAD) atomicAdd(&counters[4], 1ull);
}

// ------------------------------------------------------------------------------------------------------------------
// z = T^T w of every queued transit of the chunk (nbg_adjoint.cuh): one thread per transit, NC = 1 (TransitTiming) or 3
// (TransitParameters: time, v_sky, b_sky^2).  Reads the operator block of the transit's final step (written by transit_kernel and
// phi_dense_kernel) and the header (sky-plane separations, 1/gdot, ...); writes z scaled so that the Jacobian kernel's dot product
// IS the output:  out_c[col] = z_c^T J[:, col] (+ zm of the body whose mass column it is).
template <int NC>
__global__ void __launch_bounds__(64) transit_adjoint_kernel(EventQueue Q, int nq, int n, int ti, KMask kmask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nq) return;
  const size_t cap = Q.cap;
  const int occ = Q.body[e];
  const double dx = Q.hdr[0 * cap + e], dy = Q.hdr[1 * cap + e], dvx = Q.hdr[2 * cap + e], dvy = Q.hdr[3 * cap + e];
  const double gdinv = Q.hdr[4 * cap + e], vskyinv = Q.hdr[5 * cap + e], dvdt = Q.hdr[6 * cap + e], dt0 = Q.hdr[7 * cap + e];
  AdjVec<NMAX> Z[NC];
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    for (int r = 0; r < 3 * n; ++r) { Z[q].zx[r] = 0.0; Z[q].zv[r] = 0.0; }
    for (int r = 0; r < n; ++r) Z[q].zm[r] = 0.0;
  }
  // dtbvdq! (timing.jl:155-194): rows x0, x1, v0, v1 of the occultor minus those of the transited body
  //   time:    -(dJx0 dvx + dJx1 dvy + dJv0 dx + dJv1 dy) / gdot
  //   v_sky:   (dJv0 dvx + dJv1 dvy) / vsky + dvdt * d time        b_sky^2:  2 (dJx0 dx + dJx1 dy)
  Z[0].zx[3 * occ] = dvx; Z[0].zx[3 * occ + 1] = dvy; Z[0].zv[3 * occ] = dx; Z[0].zv[3 * occ + 1] = dy;
  Z[0].zx[3 * ti] = -dvx; Z[0].zx[3 * ti + 1] = -dvy; Z[0].zv[3 * ti] = -dx; Z[0].zv[3 * ti + 1] = -dy;
  if constexpr (NC == 3) {
    Z[NC - 2].zv[3 * occ] = dvx * vskyinv; Z[NC - 2].zv[3 * occ + 1] = dvy * vskyinv;
    Z[NC - 2].zv[3 * ti] = -dvx * vskyinv; Z[NC - 2].zv[3 * ti + 1] = -dvy * vskyinv;
    Z[NC - 1].zx[3 * occ] = 2.0 * dx; Z[NC - 1].zx[3 * occ + 1] = 2.0 * dy;
    Z[NC - 1].zx[3 * ti] = -2.0 * dx; Z[NC - 1].zx[3 * ti + 1] = -2.0 * dy;
  }
  const Src S{Q.stream + tile_offset(step_fields(n, kmask.any()), 0, 0, (size_t)e), TILE, (size_t)(e % TILE)};
  adjoint_step<NC>(Z, S, n, 0.5 * dt0, kmask);
  double* out = Q.z + (size_t)e * NC * 7 * n;
  for (int r = 0; r < 3 * n; ++r) { out[r] = -gdinv * Z[0].zx[r]; out[3 * n + r] = -gdinv * Z[0].zv[r]; }
  for (int r = 0; r < n; ++r) out[6 * n + r] = -gdinv * Z[0].zm[r];
  if constexpr (NC == 3) {
    double* o1 = out + 7 * n;
    double* o2 = out + 14 * n;
    for (int r = 0; r < 3 * n; ++r) {
      o1[r] = fma(dvdt, out[r], Z[NC - 2].zx[r]); o1[3 * n + r] = fma(dvdt, out[3 * n + r], Z[NC - 2].zv[r]);
      o2[r] = Z[NC - 1].zx[r]; o2[3 * n + r] = Z[NC - 1].zv[r];
    }
    for (int r = 0; r < n; ++r) { o1[6 * n + r] = fma(dvdt, out[6 * n + r], Z[NC - 2].zm[r]); o2[6 * n + r] = Z[NC - 1].zm[r]; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Jacobian kernel (generic N; only with NBG_FORCE_GENERIC_JAC=1 since r2): one block per system, blockDim = 32*ceil(M/32) threads, thread c owns column c.
__global__ void jac_kernel(double* __restrict__ Jv_g, double* __restrict__ Je_g, int n, size_t ld, const double* stream,
                           int nsteps, double h, const int32_t* evlist, EventQueue Q, int ti, TransitOut O, int stage_phi) {
  extern __shared__ double sm[];
  const int M = 7 * n, R6 = 6 * n;
  const long sys = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x, c = tid;
  JacSmem S;
  S.Jv = sm;
  S.Je = S.Jv + (size_t)R6 * M;
  S.da = S.Je + (size_t)R6 * M;
  S.rec = S.da + (size_t)3 * n * M;
  S.phi = stage_phi ? S.rec + 2 * KF : nullptr;
  const size_t jsz = (size_t)R6 * M;
  for (size_t q = tid; q < jsz; q += nthr) { S.Jv[q] = Jv_g[sys * jsz + q]; S.Je[q] = Je_g[sys * jsz + q]; }
  __syncthreads();
  const size_t sf = step_fields(n);
  double gacc = 0.0, cacc = 0.0;   // fused chi^2: gradient entry of column c / the system's chi^2 terms (thread 0), this chunk
  for (int s = 0; s < nsteps; ++s) {
    Src src{stream + tile_offset(sf, ld / TILE, (size_t)s, (size_t)sys), TILE, (size_t)(sys % TILE)};
    jac_apply_step(S, src, n, M, c, 0.5 * h, tid, nthr);
    if (evlist) {
      for (int i = 0; i < n; ++i) {
        const int32_t slot = evlist[((size_t)s * n + i) * ld + sys];
        if (slot < 0 || c >= M) continue;
        // transit of body i found after this step: its outputs are z^T jac_step with the adjoint vectors of the sub-step (nbg_adjoint.cuh)
        const size_t cap = Q.cap;
        const size_t rec = out_rec(O, sys, i, Q.k[slot], slot);
        for (int comp = 0; comp < O.C; ++comp) {
          const double* __restrict__ z = Q.z + ((size_t)slot * O.C + comp) * M;
          double a = 0.0;
          for (int b = 0; b < n; ++b)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              // rows of body b minus those of the transited body (see rx_transit_out: the x / v parts of z sum to zero over the bodies)
              a = fma(__ldg(z + 3 * b + k), S.Jv[(6 * b + k) * M + c] - S.Jv[(6 * ti + k) * M + c], a);
              a = fma(__ldg(z + 3 * n + 3 * b + k), S.Jv[(6 * b + 3 + k) * M + c] - S.Jv[(6 * ti + 3 + k) * M + c], a);
            }
          if (c % 7 == 6) a += __ldg(z + 6 * n + c / 7);
          if (comp == 0 && O.gq) { gacc = fma(Q.hdr[8 * cap + slot], a, gacc); if (c == 0) cacc += Q.hdr[9 * cap + slot]; }
          if (O.dtdq0) O.dtdq0[(rec * M + c) * O.C + comp] = a;
        }
      }
    }
  }
  __syncthreads();
  for (size_t q = tid; q < jsz; q += nthr) { Jv_g[sys * jsz + q] = S.Jv[q]; Je_g[sys * jsz + q] = S.Je[q]; }
  if (O.gq && evlist) {
    if (c < M) O.gq[(size_t)sys * M + c] += gacc;
    if (c == 0) O.chi2[sys] += cacc;
  }
}

// Register-resident Jacobian kernel (N <= NBG_RX_MAX_BODIES = 16): one block per system, rx_warps(N) warps, see nbg_jacobian_rx.cuh.
// register budget: 48 + 48 doubles of resident state (jac_step + jac_error halves) at N = 8 plus temporaries needs ~246 registers -> 2 blocks
// of 4 warps per SM; N = 9: 2 blocks of 4 warps at 255 registers; N = 10: 2 blocks of 5 warps at 168 registers; N = 11..14: one block of 5-7
// warps per SM (255 registers, 136-219 KB of operator ring)
template <int N> __host__ __device__ constexpr int rx_minblocks() { return N >= 6 ? 3 : (N == 5 ? 3 : 6); }

// ---- outputs of one queued transit: out_comp[c] = z_comp^T J[:, c] with the adjoint vectors of the transit sub-step (nbg_adjoint.cuh) ----
// Accuracy, two points:
//  * sum_b z_b J_b is evaluated as sum_b z_b (J_b - J_ti).  The x / v parts of z sum to zero over the bodies (every operator of the step
//    is translation invariant: pair operators add +g / -g, the drift acts per body, the columns of the force-gradient operator sum to
//    zero), and the rows of J share a large common mode in the columns of far bodies and masses (barycentre shifts).  The reference
//    forms J'_occ - J'_ti BEFORE multiplying (timing.jl:163-170); multiplying first would lose |J| / |J_occ - J_ti| in relative accuracy.
//    (A compensated dot product -- TwoProduct / TwoSum -- was measured as well: the full-length deviation of dtdq0 from the __float128
//    run went from 6.3e-11 to 5.7e-11, i.e. the summation is not what limits it, and the kernel lost 7 %; not kept.)
// Registers: the step loop of jac_rx_kernel sits at the 255-register limit and its speed depends on what else ptxas has to fit around it
// (A/B on one box, ms per 64-step window at 65,536 systems, profiles/r02d_ab.jsonl: dot product unrolled in registers with 24 loads in
// flight 240.6; as a non-inlined function on a local-memory copy of the rows 242.3; staged through shared memory with a rolled loop 226.3
// (222.0 on another box, where the r01 kernel that applied a full step per transit took 226.5 and this loop with a compensated sum 237.0).  So: both operands in shared memory -- this step's operator buffer is
// free by now -- and a rolled loop over the rows.  scratch: >= 3 * 7N + R * NT doubles, R rows per pass (all 3N rows in one pass for N >= 4).
template <int N, int NT, int SB>
__device__ __forceinline__ void rx_transit_out(const RxState<N>& S, const EventQueue& Q, const TransitOut& O, long sys, int body, int slot, int ti, int half,
                                               int c, bool valid, int tid, double* __restrict__ scratch, double* __restrict__ acc) {
  constexpr int M = 7 * N, ZMAX = 3 * M, RFIT = (SB - ZMAX) / NT, R = RFIT < 3 * N ? RFIT : 3 * N, NP = (3 * N + R - 1) / R;
  static_assert(R >= 1, "operator buffer too small for the transit dot product");
  double* const zs = scratch;         // [C][7N]
  double* const ex = scratch + ZMAX;  // [R][NT]: this thread's rows of J minus those of the transited body
  double jt[3] = {0.0, 0.0, 0.0};   // rows of the transited body (ti is a run-time index: selected arithmetically so that jv stays in registers;
#pragma unroll                      //  a chain of ?: on the register array becomes an indexed load and moves the matrix to local memory)
  for (int b = 0; b < N; ++b) {
    const double on = b == ti ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) jt[k] = fma(on, S.jv[b][k], jt[k]);
  }
  __syncthreads();  // everyone is done with this step's operators (and with the previous transit's scratch)
  for (int q = tid; q < O.C * M; q += NT) zs[q] = __ldg(Q.z + (size_t)slot * O.C * M + q);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  const bool three = O.C == 3;
  static_for<0, NP>([&](auto Pc) {
    constexpr int pass = decltype(Pc)::value;
    if (pass > 0) __syncthreads();
    static_for<pass * R, (pass + 1) * R < 3 * N ? (pass + 1) * R : 3 * N>([&](auto Rc) {
      constexpr int r = decltype(Rc)::value;
      ex[(r - pass * R) * NT + tid] = S.jv[r / 3][r % 3] - jt[r % 3];
    });
    __syncthreads();
    constexpr int nr = (pass + 1) * R < 3 * N ? R : 3 * N - pass * R;
    const double* __restrict__ zr = zs + 3 * N * half + pass * R;
#pragma unroll 1
    for (int r = 0; r < nr; ++r) {
      const double d = ex[r * NT + tid];
      a0 = fma(zr[r], d, a0);
      if (three) { a1 = fma(zr[M + r], d, a1); a2 = fma(zr[2 * M + r], d, a2); }
    }
  });
  const size_t rec = out_rec(O, sys, body, Q.k[slot], slot);
  const bool mass = valid && c % 7 == 6;   // mass rows of jac_step are unit rows: column 7p+6 also receives zm[p]
  const int zm = 6 * N + c / 7;
  double r0 = a0, r1 = a1, r2 = a2;
  r0 += shx(r0);
  if (mass) r0 += zs[zm];
  if (O.gq) {  // fused chi^2: d chi2 / d q0[c] += w_transit * d tt / d q0[c]; the chi^2 terms are summed by thread 0 in event order
    acc[tid] = fma(Q.hdr[8 * (size_t)Q.cap + slot], r0, acc[tid]);
    if (tid == 0) acc[NT] += Q.hdr[9 * (size_t)Q.cap + slot];
  }
  if (three) {
    r1 += shx(r1);
    r2 += shx(r2);
    if (mass) { r1 += zs[M + zm]; r2 += zs[2 * M + zm]; }
  }
  if (O.dtdq0 && valid && half == 0) {
    if (!three) O.dtdq0[rec * M + c] = r0;
    else { double* o = O.dtdq0 + (rec * M + c) * 3; o[0] = r0; o[1] = r1; o[2] = r2; }
  }
}

// DBUF = false (N = 15, 16): ONE operator buffer (129 / 147 KB; two do not fit in 227 KB): the next step's block is fetched after the step
// instead of under it -- a few microseconds per step exposed, against the shared-memory kernel that is 4-5 x slower at these sizes.
// shared memory of the fast-kick variant: operator buffer(s) with three dense operators, + the per-thread scratch of the first kick if it fits
__host__ __device__ constexpr size_t rx_kick_staged(int n) { return (size_t)n * (n - 1) * KF + (size_t)3 * 12 * n * n; }
__host__ __device__ constexpr bool rx_kick_dbuf(int n) { return (2 * rx_kick_staged(n) + (size_t)3 * n * rx_warps(n) * 32 + rx_warps(n) * 32 + 1) * 8 <= (size_t)227 * 1024; }
__host__ __device__ constexpr bool rx_kick_hold_fits(int n, bool dbuf) {
  return ((dbuf ? 2 : 1) * rx_kick_staged(n) + (size_t)3 * n * rx_warps(n) * 32 + rx_warps(n) * 32 + 1) * 8 <= (size_t)227 * 1024;
}
template <int N, int U, bool SYNC = true, int MB = rx_minblocks<N>(), bool KICK = false, bool DBUF = true>
__global__ void __launch_bounds__(rx_warps(N) * 32, MB)
    jac_rx_kernel(double* __restrict__ Jv_g, double* __restrict__ Je_g, size_t ld, const double* __restrict__ stream, int nsteps, double h,
                  const int32_t* __restrict__ evlist, const uint32_t* __restrict__ evmask, EventQueue Q, int ti, TransitOut O, KMask kmask, long nsys) {
  extern __shared__ __align__(16) double smrx[];
  constexpr int NS = KICK ? 3 : 1;  // dense operators per step (nbg_kicks.cuh)
  constexpr int M = 7 * N, P = N * (N - 1) / 2, SFS = P * (2 * KF + NS * PF) + NS * 12 * N * N /* stream */,
                SB = 2 * P * KF + NS * 12 * N * N /* staged */, G0 = 2 * P * KF / 4, GSKIP = P * (2 * KF + NS * PF) / 4, G1 = NS * 3 * N * N,
                NT = rx_warps(N) * 32;
  const int tid = (int)threadIdx.x;
  constexpr int NBUF = DBUF ? 2 : 1;
  // KICK only: 3N doubles per thread for the first kickfast! (rx_phisalpha_dense HOLD): in shared memory with stride NT, or, where the
  // operator buffer leaves no room for them (N = 15, 16), in local memory
  constexpr bool HOLD_LOCAL = KICK && !rx_kick_hold_fits(N, DBUF);
  double hold_l[HOLD_LOCAL ? 3 * N : 1];
  double* const hold = HOLD_LOCAL ? hold_l : smrx + NBUF * SB + threadIdx.x;
  constexpr int HS = HOLD_LOCAL ? 1 : NT;
  double* const acc = smrx + NBUF * SB + (KICK && !HOLD_LOCAL ? 3 * N * NT : 0);  // fused chi^2: NT gradient accumulators + the chi^2 sum of this chunk
  const long sys = blockIdx.x;
  if (sys >= nsys) return;
  acc[tid] = 0.0;
  if (tid == 0) acc[NT] = 0.0;
  const int lane = tid & 31, warp = tid >> 5, half = lane >> 4, c = warp * 16 + (lane & 15);
  const bool valid = c < M;
  RxState<N> S;
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const size_t q = ((size_t)sys * 6 * N + 6 * b + 3 * half + k) * M + c;
      S.jv[b][k] = valid ? Jv_g[q] : 0.0;
      S.je[b][k] = valid ? Je_g[q] : 0.0;
    }
  double* const buf0 = smrx;
  double* const buf1 = smrx + (DBUF ? SB : 0);
  const size_t ntiles = ld / TILE;
  if (nsteps > 0) rx_fetch(buf0, stream + tile_offset(SFS, ntiles, 0, (size_t)sys), TILE, (size_t)(sys % TILE), G0, GSKIP, G1, tid, NT);
  for (int s = 0; s < nsteps; ++s) {
    double* const cur = (s & 1) ? buf1 : buf0;
    uint32_t pend = evmask ? evmask[(size_t)s * ld + sys] : 0u;   // bodies with a queued transit at the end of step s
    if (!DBUF && s > 0) {   // single buffer: everyone is done with step s-1 (and its transits); fetch step s now
      __syncthreads();
      rx_fetch(buf0, stream + tile_offset(SFS, ntiles, (size_t)s, (size_t)sys), TILE, (size_t)(sys % TILE), G0, GSKIP, G1, tid, NT);
    }
    __pipeline_wait_prior(0);
    __syncthreads();  // step s operators visible; everyone is done with the other buffer
    if (DBUF && s + 1 < nsteps)
      rx_fetch((s & 1) ? buf0 : buf1, stream + tile_offset(SFS, ntiles, (size_t)(s + 1), (size_t)sys), TILE, (size_t)(sys % TILE), G0, GSKIP, G1, tid, NT);
    rx_step<N, U, SYNC, KICK>(S, cur, 0.5 * h, half, c, kmask, hold, HS);
#ifdef NBG_EXPERIMENTS
    if (kmask.w[3] == 0x80000000u) pend = 0u;  // NBG_RX_UNROLL=99, timing experiment only: no transit outputs (the step loop alone)
#endif
    while (pend != 0u) {  // uniform across the block
      const int body = __ffs(pend) - 1;
      pend &= pend - 1u;
      const int32_t slot = evlist[((size_t)s * N + body) * ld + sys];
      rx_transit_out<N, NT, SB>(S, Q, O, sys, body, slot, ti, half, c, valid, tid, cur, acc);
    }
  }
  if (valid) {
#pragma unroll
    for (int b = 0; b < N; ++b)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const size_t q = ((size_t)sys * 6 * N + 6 * b + 3 * half + k) * M + c;
        Jv_g[q] = S.jv[b][k];
        Je_g[q] = S.je[b][k];
      }
  }
  if (O.gq && evlist) {
    if (valid && half == 0) O.gq[(size_t)sys * M + c] += acc[tid];
    if (tid == 0) O.chi2[sys] += acc[NT];
  }
}

#ifdef NBG_EXPERIMENTS
// DMMA Jacobian kernel (nbg_jacobian_mma.cuh): one block per system, mma_warps(N) warps, two 8-column tiles per warp; the same
// operator staging and transit handling as jac_rx_kernel.  No fast-kick pairs (those run jac_rx_kernel<KICK>).
template <int N, int MB, int TPW>
__global__ void __launch_bounds__(mma_warps(N, TPW) * 32, MB)
    jac_mma_kernel(double* __restrict__ Jv_g, double* __restrict__ Je_g, size_t ld, const double* __restrict__ stream, int nsteps, double h,
                   const int32_t* __restrict__ evlist, const uint32_t* __restrict__ evmask, EventQueue Q, int ti, TransitOut O, long nsys) {
  extern __shared__ __align__(16) double smrx[];
  constexpr int M = 7 * N, P = N * (N - 1) / 2, SFS = P * (2 * KF + PF) + 12 * N * N /* stream */, SB = 2 * P * KF + 12 * N * N /* staged */,
                G0 = 2 * P * KF / 4, GSKIP = P * (2 * KF + PF) / 4, G1 = 3 * N * N, NT = mma_warps(N, TPW) * 32;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long sys = blockIdx.x;
  if (sys >= nsys) return;
  const MmaLane L = mma_lane(lane, warp, N, TPW);
  const bool val[2] = {L.c[0] < M && L.t < 3, TPW > 1 && L.c[1] < M && L.t < 3};
  MmaState<N, TPW> S;
#pragma unroll
  for (int T = 0; T < TPW; ++T)
#pragma unroll
    for (int b = 0; b < N; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const size_t q = ((size_t)sys * 6 * N + 6 * b + 3 * e + L.t) * M + L.c[T];
        S.jv[T][b][e] = val[T] ? Jv_g[q] : 0.0;
        S.je[T][b][e] = val[T] ? Je_g[q] : 0.0;
      }
  double* const buf0 = smrx;
  double* const buf1 = smrx + SB;
  const size_t ntiles = ld / TILE;
  if (nsteps > 0) rx_fetch(buf0, stream + tile_offset(SFS, ntiles, 0, (size_t)sys), TILE, (size_t)(sys % TILE), G0, GSKIP, G1, tid, NT);
  for (int s = 0; s < nsteps; ++s) {
    double* const cur = (s & 1) ? buf1 : buf0;
    uint32_t pend = evmask ? evmask[(size_t)s * ld + sys] : 0u;
    __pipeline_wait_prior(0);
    __syncthreads();  // step s operators visible; everyone is done with the other buffer
    if (s + 1 < nsteps)
      rx_fetch((s & 1) ? buf0 : buf1, stream + tile_offset(SFS, ntiles, (size_t)(s + 1), (size_t)sys), TILE, (size_t)(sys % TILE), G0, GSKIP, G1, tid, NT);
    mma_step<N, TPW>(S, cur, 0.5 * h, L);
    while (pend != 0u) {
      const int body = __ffs(pend) - 1;
      pend &= pend - 1u;
      const int32_t slot = evlist[((size_t)s * N + body) * ld + sys];
      // out_comp[c] = z_comp^T J[:, c] (nbg_adjoint.cuh): lane t of a column group holds rows x_t and v_t of every body
      const size_t rec = out_rec(O, sys, body, Q.k[slot], slot);
      for (int comp = 0; comp < O.C; ++comp) {
        const double* __restrict__ z = Q.z + ((size_t)slot * O.C + comp) * M;
#pragma unroll
        for (int T = 0; T < TPW; ++T) {
          double a = 0.0, jt0 = 0.0, jt1 = 0.0;
#pragma unroll
          for (int b = 0; b < N; ++b) { const double on = b == ti ? 1.0 : 0.0; jt0 = fma(on, S.jv[T][b][0], jt0); jt1 = fma(on, S.jv[T][b][1], jt1); }
          if (L.t < 3) {
#pragma unroll
            for (int b = 0; b < N; ++b) {   // rows of body b minus those of the transited body (see rx_transit_out)
              a = fma(__ldg(z + 3 * b + L.t), S.jv[T][b][0] - jt0, a);
              a = fma(__ldg(z + 3 * N + 3 * b + L.t), S.jv[T][b][1] - jt1, a);
            }
          }
          a += __shfl_xor_sync(FULL, a, 1);
          a += __shfl_xor_sync(FULL, a, 2);
          if (val[T] && L.c[T] % 7 == 6) a += __ldg(z + 6 * N + L.c[T] / 7);
          if (O.dtdq0 && val[T] && L.t == 0) O.dtdq0[(rec * M + L.c[T]) * O.C + comp] = a;
        }
      }
    }
  }
#pragma unroll
  for (int T = 0; T < TPW; ++T)
    if (val[T]) {
#pragma unroll
      for (int b = 0; b < N; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const size_t q = ((size_t)sys * 6 * N + 6 * b + 3 * e + L.t) * M + L.c[T];
          Jv_g[q] = S.jv[T][b][e];
          Je_g[q] = S.je[T][b][e];
        }
    }
}

template <int N, int MB, int TPW>
int launch_jac_mma(cudaStream_t st, long nsys, double* Jv, double* Je, size_t ld, const double* stream, int nsteps, double h,
                   const int32_t* evlist, const uint32_t* evmask, const EventQueue& Q, int ti, const TransitOut& O) {
  constexpr int P = N * (N - 1) / 2, SB = 2 * P * KF + 12 * N * N;
  const size_t smem = (size_t)2 * SB * 8;
  if (cudaFuncSetAttribute(jac_mma_kernel<N, MB, TPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  cudaFuncSetAttribute(jac_mma_kernel<N, MB, TPW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  jac_mma_kernel<N, MB, TPW><<<(unsigned)nsys, mma_warps(N, TPW) * 32, smem, st>>>(Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, nsys);
  return 0;
}
#endif  // NBG_EXPERIMENTS

// Split path, second stage: the Kepler operator records of the main steps.  One thread per (system, step, pair section);
// the 32 lanes of a warp are the systems of one tile, so the section index (hence drift_first) is uniform in a warp and
// every load/store is a 1 KB run.  Reads the SCF doubles the trajectory kernel left, rebuilds the 22 scalars (kepler_scalars), runs compute_jacobian_gamma!
// (ahl21.jl:896-1139) in registers and writes the KF-double record the Jacobian kernel consumes.
__global__ void __launch_bounds__(128, 2) pair_op_kernel(const double* __restrict__ scal, double* __restrict__ stream, int n, size_t ntiles, long nsys, double h2) {
  const int P = npairs(n);
  const int sec = blockIdx.z * blockDim.y + threadIdx.y;
  const long sys = (long)blockIdx.x * TILE + threadIdx.x;
  if (sec >= 2 * P || sys >= nsys) return;
  const double* __restrict__ src = scal + tile_offset((size_t)2 * P * SCF, ntiles, blockIdx.y, (size_t)sys) + ((size_t)sec * (SCF / 4) * TILE + threadIdx.x) * 4;
  double in[SCF];
#pragma unroll
  for (int g = 0; g < SCF / 4; ++g) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(src + (size_t)g * TILE * 4));
    const double2 b = __ldg(reinterpret_cast<const double2*>(src + (size_t)g * TILE * 4) + 1);
    in[4 * g] = a.x; in[4 * g + 1] = a.y; in[4 * g + 2] = b.x; in[4 * g + 3] = b.y;
  }
  double x0[3], v0[3], gamma, kk, bim, bjm;
  scal_unpack(in, x0, v0, gamma, kk, bim, bjm);
  KepScal S;
  S.k = kk;
  double rec[KF];
  if (kk != 0.0) kepler_scalars(x0, v0, kk, h2, sec < P, gamma, &S);
  if (S.k == 0.0) {
#pragma unroll
    for (int f = 0; f < KF; ++f) rec[f] = 0.0;
  } else {
    KepJac J;
    kepler_jacobian_inl(&S, x0, v0, sec < P, &J);
    double dl[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) { dl[j] = S.fm1 * x0[j] + S.gmh * v0[j]; dl[3 + j] = S.dfdt * x0[j] + S.dgdtm1 * v0[j]; }
    kepler_record(rec, J, dl, bim, bjm);
  }
  Emit em{stream + tile_offset(step_fields(n), ntiles, blockIdx.y, (size_t)sys), TILE, (size_t)threadIdx.x};
  em.put_record<KF>((size_t)sec * KF, rec);
}

// Dense phisalpha operator (see nbg_jacobian_rx.cuh): one thread per (system or queued transit, step, body i); the 32
// lanes of a warp are consecutive systems, so every record read and every output write is a coalesced run of sectors.
template <int N>
__global__ void __launch_bounds__(32 * N, 512 / (32 * N)) phi_dense_kernel(double* __restrict__ base, size_t ntiles, long nitems,
                                                                         const int32_t* __restrict__ nitems_dev, int kicked) {
  const long idx = (long)blockIdx.x * 32 + threadIdx.x;
  const long nv = nitems_dev ? min((long)*nitems_dev, nitems) : nitems;
  if (idx >= nv) return;
  double* blk = base + tile_offset(step_fields(N, kicked != 0), ntiles, blockIdx.y, (size_t)idx);
  if (!kicked) phi_dense_rows<N>(blk, TILE, (size_t)threadIdx.x, (int)threadIdx.y);
  else
    phi_dense_rows_kicked<N>(blk, TILE, (size_t)threadIdx.x, (int)threadIdx.y, phi_rec_offset(N, blockIdx.z), phi_dense_offset(N, true, blockIdx.z));
}
// the same (no fast-kick pairs) with the per-pair tensors cached in shared memory (phi_dense_rows_cached): FULL for N <= 10
// (one block per SM at N = 8), the T / gam cache alone for N = 11, 12 (207 KB at N = 12)
template <int N, bool FULL>
__global__ void __launch_bounds__(32 * N, FULL ? 1 : 512 / (32 * N)) phi_dense_cached_kernel(double* __restrict__ base, size_t ntiles, long nitems,
                                                                                           const int32_t* __restrict__ nitems_dev) {
  extern __shared__ double sm_phi[];
  const long idx = (long)blockIdx.x * 32 + threadIdx.x;
  const long nv = nitems_dev ? min((long)*nitems_dev, nitems) : nitems;
  if ((long)blockIdx.x * 32 >= nv) return;  // uniform: the whole tile is empty
  double* blk = base + tile_offset(step_fields(N, false), ntiles, blockIdx.y, (size_t)idx);
  phi_dense_rows_cached<N, FULL>(blk, TILE, (size_t)threadIdx.x, (int)threadIdx.y, idx < nv, sm_phi,
                                 sm_phi + (size_t)(N * (N - 1) / 2) * phi_tgf(FULL) * 32, (int)threadIdx.x);
}
template <int N, bool FULL> int launch_phi_dense_cached(cudaStream_t st, const dim3& grid, const dim3& block, double* base, size_t ntiles, long nitems,
                                                        const int32_t* nitems_dev) {
  const size_t smem = phi_cache_bytes(N, FULL);
  if (cudaFuncSetAttribute(phi_dense_cached_kernel<N, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  phi_dense_cached_kernel<N, FULL><<<grid, block, smem, st>>>(base, ntiles, nitems, nitems_dev);
  return 0;
}
// main steps: ntiles = ld / 32, nsteps steps; queued transits: ntiles = 0 (one "step"), nitems_dev = device count of queued transits
int launch_phi_dense(cudaStream_t st, int n, double* base, size_t ntiles, long nitems, const int32_t* nitems_dev, int nsteps, bool kicked = false,
                     int cached = 2) {
  if (nitems <= 0 || nsteps <= 0) return 0;
  const dim3 grid((unsigned)((nitems + 31) / 32), (unsigned)nsteps, kicked ? 3u : 1u), block(32, n);
  const int kf = kicked ? 1 : 0;
  if (!kicked && cached) switch (n) {
    case 2: return cached == 2 ? launch_phi_dense_cached<2, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<2, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 3: return cached == 2 ? launch_phi_dense_cached<3, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<3, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 4: return cached == 2 ? launch_phi_dense_cached<4, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<4, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 5: return cached == 2 ? launch_phi_dense_cached<5, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<5, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 6: return cached == 2 ? launch_phi_dense_cached<6, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<6, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 7: return cached == 2 ? launch_phi_dense_cached<7, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<7, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 8: return cached == 2 ? launch_phi_dense_cached<8, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<8, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 9: return cached == 2 ? launch_phi_dense_cached<9, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<9, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 10: return cached == 2 ? launch_phi_dense_cached<10, true>(st, grid, block, base, ntiles, nitems, nitems_dev) : launch_phi_dense_cached<10, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 11: return launch_phi_dense_cached<11, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    case 12: return launch_phi_dense_cached<12, false>(st, grid, block, base, ntiles, nitems, nitems_dev);
    default: break;  // 13, 14: the cache does not fit in shared memory
  }
  switch (n) {
    case 2: phi_dense_kernel<2><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 3: phi_dense_kernel<3><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 4: phi_dense_kernel<4><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 5: phi_dense_kernel<5><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 6: phi_dense_kernel<6><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 7: phi_dense_kernel<7><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 8: phi_dense_kernel<8><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 9: phi_dense_kernel<9><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 10: phi_dense_kernel<10><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 11: phi_dense_kernel<11><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 12: phi_dense_kernel<12><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 13: phi_dense_kernel<13><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 14: phi_dense_kernel<14><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;
    case 15: phi_dense_kernel<15><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;   // N = 15, 16: only the queued transits
    case 16: phi_dense_kernel<16><<<grid, block, 0, st>>>(base, ntiles, nitems, nitems_dev, kf); break;   // (their adjoint needs the dense operator)
    default: return -1;
  }
  return 0;
}

template <int N, int U, bool SYNC = true, int MB = rx_minblocks<N>(), bool KICK = false, bool DBUF = true>
int launch_jac_rx(cudaStream_t st, long nsys, double* Jv, double* Je, size_t ld, const double* stream, int nsteps, double h, const int32_t* evlist,
                  const uint32_t* evmask, const EventQueue& Q, int ti, const TransitOut& O, const KMask& kmask = KMask{}) {
  constexpr int P = N * (N - 1) / 2, NS = KICK ? 3 : 1, SB = 2 * P * KF + NS * 12 * N * N;
  const size_t smem = ((size_t)(DBUF ? 2 : 1) * SB + (KICK && rx_kick_hold_fits(N, DBUF) ? (size_t)3 * N * rx_warps(N) * 32 : 0) + rx_warps(N) * 32 + 1) * 8;
  // per launch, not once: function attributes are per device, and plans of one process may live on different devices
  if (cudaFuncSetAttribute(jac_rx_kernel<N, U, SYNC, MB, KICK, DBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  cudaFuncSetAttribute(jac_rx_kernel<N, U, SYNC, MB, KICK, DBUF>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  jac_rx_kernel<N, U, SYNC, MB, KICK, DBUF><<<(unsigned)nsys, rx_warps(N) * 32, smem, st>>>(Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask, nsys);
  return 0;
}
// fast-kick pairs: one generic variant per N (pivot blocks of 1, one block per SM)
int launch_jac_rx_kicked(int n, cudaStream_t st, long nsys, double* Jv, double* Je, size_t ld, const double* stream, int nsteps, double h,
                         const int32_t* evlist, const uint32_t* evmask, const EventQueue& Q, int ti, const TransitOut& O, const KMask& kmask) {
  switch (n) {
    case 2: return launch_jac_rx<2, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 3: return launch_jac_rx<3, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 4: return launch_jac_rx<4, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 5: return launch_jac_rx<5, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 6: return launch_jac_rx<6, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 7: return launch_jac_rx<7, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 8: return launch_jac_rx<8, 1, true, 1, true>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    // N > 8: two operator buffers while they fit (N <= 11), then one; N = 15, 16 keep the first kick's scratch in local memory
    case 9: return launch_jac_rx<9, 1, true, 1, true, rx_kick_dbuf(9)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 10: return launch_jac_rx<10, 1, true, 1, true, rx_kick_dbuf(10)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 11: return launch_jac_rx<11, 1, true, 1, true, rx_kick_dbuf(11)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 12: return launch_jac_rx<12, 1, true, 1, true, rx_kick_dbuf(12)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 13: return launch_jac_rx<13, 1, true, 1, true, rx_kick_dbuf(13)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 14: return launch_jac_rx<14, 1, true, 1, true, rx_kick_dbuf(14)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 15: return launch_jac_rx<15, 1, true, 1, true, rx_kick_dbuf(15)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
    case 16: return launch_jac_rx<16, 1, true, 1, true, rx_kick_dbuf(16)>(st, nsys, Jv, Je, ld, stream, nsteps, h, evlist, evmask, Q, ti, O, kmask);
  }
  return -1;
}

// dtdelements = dtdq0 . jac_init  (calc_dtdelements!, timing.jl:112-138).  One block per (system, record tile).
__global__ void dtdelements_kernel(const double* __restrict__ dtdq0, const double* __restrict__ jac_init, double* __restrict__ out,
                                   const int32_t* __restrict__ count, const int32_t* ntt_body, const int32_t* off, int n, size_t ld, int RT, int C,
                                   long sys0) {
  extern __shared__ double ji[];  // M x M column-major with the column stride padded to an odd MP: ji[col*MP + row] (thread = col: no bank conflicts)
  const int M = 7 * n, MP = M | 1;
  const long sys = sys0 + blockIdx.x;
  for (int q = threadIdx.x; q < M * M; q += blockDim.x) ji[(q / M) * MP + (q % M)] = jac_init[(size_t)sys * M * M + q];
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    const int nk = min(count[i * ld + sys], ntt_body[i]);
    for (int idx = threadIdx.x; idx < nk * M * C; idx += blockDim.x) {
      const int comp = idx % C, col = (idx / C) % M, k = idx / (C * M);
      const size_t rec = (size_t)sys * RT + off[i] + k;
      double acc = 0.0;
      for (int row = 0; row < M; ++row) acc += dtdq0[(rec * M + row) * C + comp] * ji[col * MP + row];
      out[(rec * M + col) * C + comp] = acc;
    }
  }
}

// The same product for the event rows of one chunk (one-shot calls stream their outputs chunk by chunk): out[e] = rows[e] . jac_init[sys_e].
// One block per queued transit, thread = output column; ascending-row accumulation as in dtdelements_kernel.
__global__ void dtde_rows_kernel(const double* __restrict__ rows, const double* __restrict__ jac_init, double* __restrict__ out,
                                 const int32_t* __restrict__ qsys, int nq, int M, int C) {
  extern __shared__ double rw[];  // this transit's dtdq0 row, M * C doubles
  const int e = blockIdx.x;
  if (e >= nq) return;
  for (int q = threadIdx.x; q < M * C; q += blockDim.x) rw[q] = rows[(size_t)e * M * C + q];
  __syncthreads();
  const double* __restrict__ ji = jac_init + (size_t)qsys[e] * M * M;
  for (int idx = threadIdx.x; idx < M * C; idx += blockDim.x) {
    const int comp = idx % C, col = idx / C;
    double acc = 0.0;
    for (int row = 0; row < M; ++row) acc += rw[row * C + comp] * __ldg(ji + (size_t)col * M + row);
    out[(size_t)e * M * C + idx] = acc;
  }
}

// Fused transit-time likelihood: chi^2 = sum ((tt - t_obs) / sigma)^2 over the stored transits of one system and
// its gradients with respect to the initial Cartesian coordinates (dtdq0) and to the orbital elements (dtdelements), reduced on
// the device: 1 + 2M doubles per system leave the GPU instead of the full dtdq0 / dtdelements arrays.  One block per system, one
// thread per column; observations with sigma <= 0 (or a non-finite t_obs) are skipped.
__global__ void chi2_kernel(const double* __restrict__ tt, const double* __restrict__ dtdq0, const double* __restrict__ dtde,
                            const int32_t* __restrict__ count, const int32_t* __restrict__ ntt_body, const int32_t* __restrict__ off, int n, size_t ld,
                            int RT, const double* __restrict__ tobs, const double* __restrict__ sigma, int per_system, double* __restrict__ chi2,
                            double* __restrict__ gq, double* __restrict__ ge) {
  const int M = 7 * n, c = threadIdx.x;
  const long sys = blockIdx.x;
  const double* to = tobs + (per_system ? (size_t)sys * RT : 0);
  const double* sg = sigma + (per_system ? (size_t)sys * RT : 0);
  double acc = 0.0, aq = 0.0, ae = 0.0;
  for (int i = 0; i < n; ++i) {
    const int nk = min(count[i * ld + sys], ntt_body[i]);
    for (int k = 0; k < nk; ++k) {
      const int slot = off[i] + k;
      const double s = sg[slot], t = to[slot];
      if (!(s > 0.0) || !isfinite(t)) continue;
      const size_t rec = (size_t)sys * RT + slot;
      const double r = (tt[rec] - t) / s;
      acc = fma(r, r, acc);
      const double w = 2.0 * r / s;
      if (c < M) {
        if (dtdq0 && gq) aq = fma(w, dtdq0[rec * M + c], aq);
        if (dtde && ge) ae = fma(w, dtde[rec * M + c], ae);
      }
    }
  }
  if (c == 0) chi2[sys] = acc;
  if (c < M) {
    if (gq) gq[(size_t)sys * M + c] = aq;
    if (ge) ge[(size_t)sys * M + c] = ae;
  }
}

// ---- layout conversion kernels (host AoS <-> device SoA) ----
__global__ void pack_xvm_kernel(const double* x, const double* v, const double* m, const double* xe, const double* ve, const double* dqdt,
                                TrajArrays T, int n, long nsys, double t0) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  const size_t ld = T.ld;
  for (int q = 0; q < 3 * n; ++q) {
    T.x[q * ld + sys] = x[sys * 3 * n + q];
    T.v[q * ld + sys] = v[sys * 3 * n + q];
    T.xe[q * ld + sys] = xe ? xe[sys * 3 * n + q] : 0.0;
    T.ve[q * ld + sys] = ve ? ve[sys * 3 * n + q] : 0.0;
  }
  for (int q = 0; q < n; ++q) T.m[q * ld + sys] = m[sys * n + q];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 6; ++k) T.dq[(6 * i + k) * ld + sys] = dqdt ? dqdt[sys * 7 * n + 7 * i + k] : 0.0;
  T.t[sys] = t0;
  T.status[sys] = 0;
}
__global__ void unpack_xv_kernel(TrajArrays T, int n, long nsys, double* x, double* v, double* xe, double* ve, double* dqdt) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  const size_t ld = T.ld;
  for (int q = 0; q < 3 * n; ++q) {
    if (x) x[sys * 3 * n + q] = T.x[q * ld + sys];
    if (v) v[sys * 3 * n + q] = T.v[q * ld + sys];
    if (xe) xe[sys * 3 * n + q] = T.xe[q * ld + sys];
    if (ve) ve[sys * 3 * n + q] = T.ve[q * ld + sys];
  }
  if (dqdt)
    for (int i = 0; i < n; ++i) {
      for (int k = 0; k < 6; ++k) dqdt[sys * 7 * n + 7 * i + k] = T.dq[(6 * i + k) * ld + sys];
      dqdt[sys * 7 * n + 7 * i + 6] = 0.0;
    }
}
// device [sys][6N][M] <-> Julia [sys][col][row] (M x M), identity mass rows
__global__ void jac_to_julia_kernel(const double* J6, double* out, int n, int is_error) {
  const int M = 7 * n;
  const long sys = blockIdx.x;
  for (int q = threadIdx.x; q < M * M; q += blockDim.x) {
    const int col = q / M, row = q % M, b = row / 7, k = row % 7;
    double val;
    if (k == 6) val = (!is_error && row == col) ? 1.0 : 0.0;
    else val = J6[((size_t)sys * 6 * n + 6 * b + k) * M + col];
    out[(size_t)sys * M * M + q] = val;
  }
}
__global__ void jac_from_julia_kernel(const double* in, double* J6, int n, int is_error) {
  const int M = 7 * n;
  const long sys = blockIdx.x;
  for (int q = threadIdx.x; q < 6 * n * M; q += blockDim.x) {
    const int r6 = q / M, col = q % M, b = r6 / 6, k = r6 % 6;
    double val;
    if (in) val = in[(size_t)sys * M * M + (size_t)col * M + 7 * b + k];
    else val = (!is_error && (7 * b + k) == col) ? 1.0 : 0.0;
    J6[(size_t)sys * 6 * n * M + q] = val;
  }
}
// [sample][3n][ld] -> [sample][sys][body][3]
__global__ void unpack_samples_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, long nsys, size_t ld, long nsamp) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long k = blockIdx.y;
  if (sys >= nsys || k >= nsamp) return;
  for (int q = 0; q < 3 * n; ++q) dst[((size_t)k * nsys + sys) * 3 * n + q] = src[((size_t)k * 3 * n + q) * ld + sys];
}
__global__ void count_out_kernel(const int32_t* count, size_t ld, int n, long nsys, int64_t* out) {
  const long sys = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sys >= nsys) return;
  for (int i = 0; i < n; ++i) out[sys * n + i] = count[i * ld + sys];
}

// FP64 pipe peak: 8 independent DFMA chains per thread (roofline denominator; MEASURED_PEAKS.json has no FP64 entry)
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t b) {
    if (b <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    if (cudaMalloc(&p, b) != cudaSuccess) { cudaGetLastError(); return -1; }
    bytes = b;
    return 0;
  }
  // per-chunk buffers sized from a measured count: grow with headroom so that a slowly rising count does not reallocate every chunk
  int ensure_grow(size_t b) { return b <= bytes ? 0 : ensure(b + b / 4 + 4096); }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return (T*)p; }
};
struct PinBuf {  // pinned host memory
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t b) {
    if (b <= bytes) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    b += b / 4 + 4096;
    if (cudaHostAlloc(&p, b, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
    bytes = b;
    return 0;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

}  // namespace

struct nbg_plan {
  // multi-device plan (nbg_plan_create_multi): the parent only holds its children, one per device slice, and the first system of each
  std::vector<nbg_plan*> kids;
  std::vector<long> kid_lo;
  int n = 0, device = 0;
  long nsys = 0;
  size_t ld = 0;
  int64_t stream_budget = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // uploads that overlap the stepping (jac_init) and the per-chunk output copies
  cudaEvent_t copy_done = nullptr;
  cudaStream_t aux_stream = nullptr;   // operator kernels of the main steps, concurrent with the transit refinement
  cudaEvent_t ev_traj = nullptr, ev_ops = nullptr, ev_ops2 = nullptr, ev_chunk = nullptr;
  cudaStream_t aux2_stream = nullptr;  // NBG_OVERLAP=2: the dense phisalpha operators on a third stream
  TrajArrays T{};
  DevBuf bx, bv, bxe, bve, bm, bdq, bgs, bt, bterr, bcount, bstatus;
  DevBuf bbackup;  // trajectory state at the start of a chunk: a chunk whose transit queue overflowed is re-run from it
  DevBuf bJv, bJe, bstream, bscal, bevlist, bevmask;
  DevBuf qn, qsys, qstep, qbody, qk, qdt0, qt, qsnap, qhdr, qstream, qz;
  DevBuf btt, bdtdq0, bdtde, bjinit, bntt, boff, bcounters, belem;
  DevBuf bevtt, bevd, beve;            // event-row outputs of one chunk (one-shot calls)
  DevBuf bchi2, bgq, btobs, bsigma;    // fused likelihood
  DevBuf stage[8];  // staging for host<->device conversions
  PinBuf hqn;       // queued-transit count of the chunk, read back after the trajectory kernel
  PinBuf hstage[3]; // pinned staging of three chunks' transit rows in flight (copy / scatter / free)
  cudaEvent_t ev_copied[3] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> tev;  // timing events, created once and reused by every call (Timer)
  bool has_state = false, jac_valid = false, force_generic_jac = false, split_traj = true, overlap = true, overlap3 = false;
  double* samp_jac = nullptr;  // nbg_integrate_sampled_jac: jac_step before every samp_stride-th step, [sample][sys][7n][7n] (Julia layout); null = off
  KMask kmask;  // fast-kick pairs (s.pair), bit = pair index i*n - i(i+1)/2 + (j-i-1), i < j
  int rx_unroll = 38;
  int phi_cached = 2;  // NBG_PHI_CACHED: 0 = phi_dense_kernel without the shared-memory cache of the per-pair tensors, 1 = T / gam cached, 2 = all pair fields
  int jac_mma = 0;     // NBG_JAC_MMA (NBG_EXPERIMENTS builds only): DMMA Jacobian kernel for N = 8, measured 12-18 % slower than jac_rx_kernel
  int newton_pre = 2;  // gradient-free pre-iterations of the transit Newton solve (NBG_NEWTON_PRE)
  unsigned scatter_threads = 4;  // NBG_SCATTER_THREADS: host threads that scatter a chunk's transit rows into the caller's arrays
  long queue_cap0 = 0; // NBG_QUEUE_CAP0: initial transit-queue capacity (tests force it tiny to exercise the re-run)
  int32_t ntt_body[NBG_MAX_BODIES] = {0}, off[NBG_MAX_BODIES] = {0};
  int RT = 0, C = 1;
  bool have_transit = false, have_dtde = false, transit_grad = false;
  bool jinit_resident = false;  // bjinit holds jac_init computed on the device by nbg_set_state_elements
  unsigned long long counters_host[8] = {0};
  double timings[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long launches = 0;
  long chunk_retries = 0;      // chunks re-run because the transit queue was too small
  int64_t generation = 0;      // bumped whenever the resident state changes (callers that cache residency compare it)
  // Output mode of the transit driver for the current call.
  //  dense  : tt / dtdq0 / dtdelements as full device arrays [sys][RT]... (nbg_transit_timing_resident + nbg_transit_fetch)
  //  rows   : per-chunk compact rows, one per queued transit, copied to pinned staging on the copy stream and scattered into the
  //           caller's host arrays while the following chunks compute (nbg_transit_timing, host buffers known up front)
  //  fused  : chi^2 and its gradient accumulated in the Jacobian kernel; no per-transit gradient rows exist anywhere
  struct Sink { double* tt = nullptr; double* dtdq0 = nullptr; double* dtde = nullptr; bool want_dtde = false, active = false; };
  Sink sink;
  bool fused = false, fused_tt = false;
  int fused_per_system = 0;
  struct Job { int buf; int nq; };
  std::deque<Job> jobs;        // chunks whose rows are in pinned staging, not yet scattered
  long chunk_seq = 0;
  bool trace = false;          // NBG_TRACE=1: host wall-clock of the stages of the one-shot calls on stderr
};
static double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace {

int alloc_state(nbg_plan* p) {
  const size_t ld = p->ld, n = p->n;
  int bad = 0;
  bad |= p->bx.ensure(3 * n * ld * 8); bad |= p->bv.ensure(3 * n * ld * 8); bad |= p->bxe.ensure(3 * n * ld * 8); bad |= p->bve.ensure(3 * n * ld * 8);
  bad |= p->bm.ensure(n * ld * 8); bad |= p->bdq.ensure(6 * n * ld * 8); bad |= p->bgs.ensure(n * ld * 8); bad |= p->bt.ensure(ld * 8);
  bad |= p->bterr.ensure(ld * 8); bad |= p->bcount.ensure(n * ld * 4); bad |= p->bstatus.ensure(ld * 4);
  bad |= p->bcounters.ensure(8 * 8); bad |= p->bntt.ensure(NBG_MAX_BODIES * 4); bad |= p->boff.ensure(NBG_MAX_BODIES * 4);
  bad |= p->qn.ensure(4);
  bad |= p->hqn.ensure(64);
  if (bad) return -1;
  p->T = TrajArrays{p->bx.as<double>(), p->bv.as<double>(), p->bxe.as<double>(), p->bve.as<double>(), p->bm.as<double>(), p->bdq.as<double>(),
                    p->bgs.as<double>(), p->bt.as<double>(), p->bcount.as<int32_t>(), p->bstatus.as<uint32_t>(), ld};
  return 0;
}

size_t jac_smem_bytes(int n, bool stage_phi) {
  const size_t M = 7 * n, R6 = 6 * n;
  size_t d = 2 * R6 * M + 3 * (size_t)n * M + 2 * KF + (stage_phi ? (size_t)npairs(n) * PF : 0);
  return d * 8;
}

// Per-kernel device times of one call: pairs of events from the plan's pool (created on first use, destroyed with the plan).
struct Timer {
  nbg_plan* p;
  cudaStream_t s;
  size_t first, used = 0;
  std::vector<int> kinds;
  cudaStream_t cur = nullptr;
  explicit Timer(nbg_plan* plan) : p(plan), s(plan->stream), first(0) {}
  cudaEvent_t ev(size_t k) {
    while (p->tev.size() <= k) {
      cudaEvent_t e = nullptr;
      cudaEventCreate(&e);
      p->tev.push_back(e);
    }
    return p->tev[k];
  }
  // the first pair brackets the whole call
  void start() { cudaEventRecord(ev(0), s); used = 2; }
  void stop() { cudaEventRecord(ev(1), s); }
  void begin(int kind, cudaStream_t on = nullptr) {
    cur = on ? on : s;
    cudaEventRecord(ev(used), cur);
    kinds.push_back(kind);
    used += 2;
  }
  void end() { cudaEventRecord(ev(used - 1), cur); }
  void collect(double* ms8) {
    for (size_t q = 0; q < kinds.size(); ++q) {
      float t = 0;
      if (cudaEventElapsedTime(&t, ev(2 + 2 * q), ev(3 + 2 * q)) == cudaSuccess) ms8[kinds[q]] += t;
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, ev(0), ev(1));
    ms8[4] = tot;
    cudaGetLastError();
  }
};

double check_step(double t0, double tmax) {  // Integrator.jl:249-259
  auto sg = [](double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); };
  if (std::fabs(tmax) > std::fabs(t0)) return sg(tmax);
  if (sg(tmax) != sg(t0)) return sg(tmax);
  return -1 * sg(tmax);
}

inline long round32(long v) { return (v + 31) / 32 * 32; }

// ---- per-chunk output rows -> caller's arrays ----------------------------------------------------------------------------------
// Staging layout of one chunk with nq queued transits (nqp = nq rounded up to 32):
//   int32 sys[nqp], body[nqp], k[nqp] | double tt[nq][C] | dtdq0[nq][M][C] | dtdelements[nq][M][C]
struct StageLayout {
  size_t o_sys, o_body, o_k, o_tt, o_d, o_e, total;
  StageLayout(long nq, size_t M, size_t C, bool grad, bool dtde) {
    const size_t nqp = (size_t)round32(nq);
    o_sys = 0; o_body = nqp * 4; o_k = 2 * nqp * 4;
    o_tt = (3 * nqp * 4 + 7) / 8 * 8;
    o_d = o_tt + (size_t)nq * C * 8;
    o_e = o_d + (grad ? (size_t)nq * M * C * 8 : 0);
    total = o_e + (dtde ? (size_t)nq * M * C * 8 : 0);
  }
};
// the host half of the streaming: rows of a finished chunk into tt[sys][off[i]+k], dtdq0[sys][off[i]+k][...]
int scatter_job(nbg_plan* p, const nbg_plan::Job& j) {
  CK(cudaEventSynchronize(p->ev_copied[j.buf]));
  const size_t M = 7 * (size_t)p->n, C = p->C, RT = p->RT, row = M * C;
  const bool grad = p->sink.dtdq0 != nullptr, dtde = p->sink.want_dtde && p->sink.dtde != nullptr;
  const StageLayout L(j.nq, M, C, p->transit_grad, p->sink.want_dtde);
  const char* base = (const char*)p->hstage[j.buf].p;
  const int32_t* sys = (const int32_t*)(base + L.o_sys);
  const int32_t* body = (const int32_t*)(base + L.o_body);
  const int32_t* k = (const int32_t*)(base + L.o_k);
  const double* tt = (const double*)(base + L.o_tt);
  const double* d = (const double*)(base + L.o_d);
  const double* e = (const double*)(base + L.o_e);
  auto rows = [&](long q0, long q1) {
    for (long q = q0; q < q1; ++q) {
      const size_t rec = (size_t)sys[q] * RT + p->off[body[q]] + k[q];
      if (p->sink.tt) std::memcpy(p->sink.tt + rec * C, tt + (size_t)q * C, C * 8);
      if (grad && p->transit_grad) std::memcpy(p->sink.dtdq0 + rec * row, d + (size_t)q * row, row * 8);
      if (dtde) std::memcpy(p->sink.dtde + rec * row, e + (size_t)q * row, row * 8);
    }
  };
  // rows land in disjoint places: a large chunk (65,536 systems: ~90 k rows, 80 MB) is scattered by a few threads so that the host keeps
  // up with the device and the last chunks' rows, whose scatter nothing overlaps, cost little
  const int nth = j.nq >= 16384 ? (int)std::min<unsigned>(p->scatter_threads, std::max(1u, std::thread::hardware_concurrency())) : 1;
  if (nth <= 1) {
    rows(0, j.nq);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nth; ++t) th.emplace_back(rows, (long)j.nq * t / nth, (long)j.nq * (t + 1) / nth);
    for (auto& t : th) t.join();
  }
  return 0;
}
// scatter finished chunks until at most `keep` remain in flight
int drain_jobs(nbg_plan* p, size_t keep) {
  while (p->jobs.size() > keep) {
    if (int r = scatter_job(p, p->jobs.front())) return r;
    p->jobs.pop_front();
  }
  return 0;
}

// core driver: runs `nsteps` steps of size h from the resident state in chunks.
// detect: transit detection on; grad: propagate Jacobian; kahan_time: s.t accumulates h with Kahan
int run_steps(nbg_plan* p, double h, long nsteps, bool grad, bool detect, int ti, double t0, double h_intr, bool kahan_time, Timer& tm,
              double rate_hint) {
  if (nsteps <= 0) return 0;
  const int n = p->n;
  const long nsys = p->nsys;
  const size_t ld = p->ld;
  const size_t M = 7 * (size_t)n, C = p->C;
  const bool kicks = p->kmask.any();
  const size_t sf = step_fields(n, kicks);
  const bool rows = detect && p->sink.active;            // per-chunk event rows, streamed to the host
  const bool fused = detect && grad && p->fused;        // chi^2 gradient accumulated in the Jacobian kernel
  // chunk length from the stream budget
  size_t per_step = sf * ld * 8;
  const size_t per_step_scal = p->split_traj ? (size_t)2 * npairs(n) * SCF * ld * 8 : 0;
  long S = 1;
  if (grad) {
    S = (long)std::max<int64_t>(1, std::min<int64_t>(p->stream_budget / (int64_t)(per_step + per_step_scal), 64));
    S = std::min(S, nsteps);
    // equal chunks: 64 steps under a budget of 9 run as 8 x 8, not 7 x 9 + 1 (a one-step chunk pays a full Jacobian load/store)
    const long nchunks = (nsteps + S - 1) / S;
    S = (nsteps + nchunks - 1) / nchunks;
    if (p->bstream.ensure((size_t)S * per_step)) return fail(NBG_ERR_NOMEM, "operator stream allocation failed");
    if (per_step_scal && p->bscal.ensure((size_t)S * per_step_scal)) return fail(NBG_ERR_NOMEM, "scalar stream allocation failed");
  } else {
    S = std::min<long>(nsteps, 256);
  }
  EventQueue Q{};
  int32_t* evlist = nullptr;
  uint32_t* evmask = nullptr;
  // queue arrays that the trajectory kernel fills (sized by capacity); the per-transit operator stream and output rows are sized
  // from the measured count after the trajectory kernel
  auto alloc_queue = [&](long want) -> int {
    long cap = std::min<long>(std::max<long>(want, 32), (long)nsys * S * (n - 1));
    cap = round32(std::max<long>(cap, 32));
    int bad = 0;
    bad |= p->qsys.ensure(cap * 4); bad |= p->qstep.ensure(cap * 4); bad |= p->qbody.ensure(cap * 4); bad |= p->qk.ensure(cap * 4);
    bad |= p->qdt0.ensure(cap * 8); bad |= p->qt.ensure(cap * 8); bad |= p->qsnap.ensure((size_t)12 * n * cap * 8);
    bad |= p->qhdr.ensure((size_t)HDR * cap * 8);
    if (bad) return fail(NBG_ERR_NOMEM, "transit queue allocation failed");
    Q = EventQueue{p->qn.as<int32_t>(), (int32_t)cap, p->qsys.as<int32_t>(), p->qstep.as<int32_t>(), p->qbody.as<int32_t>(), p->qk.as<int32_t>(),
                   p->qdt0.as<double>(), p->qt.as<double>(), p->qsnap.as<double>(), p->qhdr.as<double>(), p->qstream.as<double>(), p->qz.as<double>()};
    return 0;
  };
  // trajectory state saved at the start of every chunk with detection: x, v, xe, ve, gsave, t, terr, count
  const size_t bk_d = (size_t)(12 * n + n + 2) * ld;  // doubles
  struct Seg { DevBuf* b; size_t bytes; };
  const Seg segs[8] = {{&p->bx, 3 * (size_t)n * ld * 8}, {&p->bv, 3 * (size_t)n * ld * 8}, {&p->bxe, 3 * (size_t)n * ld * 8}, {&p->bve, 3 * (size_t)n * ld * 8},
                       {&p->bgs, (size_t)n * ld * 8},    {&p->bt, ld * 8},                 {&p->bterr, ld * 8},               {&p->bcount, (size_t)n * ld * 4}};
  auto backup = [&](bool restore) -> int {
    char* bk = p->bbackup.as<char>();
    size_t o = 0;
    for (const Seg& sg : segs) {
      if (restore) CK(cudaMemcpyAsync(sg.b->p, bk + o, sg.bytes, cudaMemcpyDeviceToDevice, p->stream));
      else CK(cudaMemcpyAsync(bk + o, sg.b->p, sg.bytes, cudaMemcpyDeviceToDevice, p->stream));
      o += sg.bytes;
    }
    return 0;
  };
  if (detect) {
    long cap0 = p->queue_cap0;
    if (cap0 <= 0) {  // heuristic from the caller's slot counts: twice the mean transit rate; wrong guesses only cost a re-run of the chunk
      const double rate = std::min<double>(n - 1, rate_hint * 2.0 + 0.05);
      cap0 = (long)std::ceil((double)nsys * (double)S * rate) + 4096;
    }
    if (int r = alloc_queue(cap0)) return r;
    int bad = 0;
    bad |= p->bevlist.ensure((size_t)S * n * ld * 4);
    bad |= p->bevmask.ensure((size_t)S * ld * 4);
    bad |= p->bbackup.ensure(bk_d * 8 + (size_t)n * ld * 4);
    if (bad) return fail(NBG_ERR_NOMEM, "event list allocation failed");
    evlist = p->bevlist.as<int32_t>();
    evmask = p->bevmask.as<uint32_t>();
  }
  const bool use_rx = n <= NBG_RX_MAX_BODIES && (!p->force_generic_jac || kicks);
  const int tpb = 128;
  const unsigned gridA = (unsigned)((nsys + tpb - 1) / tpb);
  const int tps = 32 * ((7 * n + 31) / 32);
  bool stage_phi = true;
  size_t smem = jac_smem_bytes(n, true);
  if (smem > 227 * 1024) { stage_phi = false; smem = jac_smem_bytes(n, false); }
  if (grad && !use_rx) {
    if (smem > 227 * 1024) return fail(NBG_ERR_UNSUPPORTED, "jac_step does not fit in shared memory for this nbody");
    CK(cudaFuncSetAttribute(jac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  unsigned long long* dcount = p->bcounters.as<unsigned long long>();
  bool waited_jinit = false;
  int prev_buf = -1;  // staging buffer whose device-to-host copy may still read the event-row buffers
  long done = 0;
  while (done < nsteps) {
    int s = (int)std::min<long>(S, nsteps - done);
    if (p->samp_jac && grad) {
      // the saved State of the reference carries jac_step (Outputs.jl:40 deep-copies it): chunks end on sample steps, and the matrix of
      // every sample step is copied out (in stream order, after the previous chunk's Jacobian kernel) before the chunk that starts there
      const long st = p->T.samp_stride;
      if (done % st == 0) {
        jac_to_julia_kernel<<<(unsigned)nsys, 256, 0, p->stream>>>(p->bJv.as<double>(), p->samp_jac + (size_t)(done / st) * nsys * M * M, n, 0);
        p->launches++;
      }
      s = (int)std::min<long>(s, st - done % st);
    }
    double* tkerr = kahan_time ? p->bterr.as<double>() : nullptr;
    int s_split = 0;  // steps of this chunk whose Kepler records come from pair_op_kernel (split path)
    long nq = 0;
    for (int attempt = 0;; ++attempt) {
      if (detect) {
        if (int r = backup(attempt > 0)) return r;
        CK(cudaMemsetAsync(p->qn.p, 0, 4, p->stream));
      }
      tm.begin(0);
      if (grad && p->split_traj && !kicks) {
        // dq/dh restarts from zero every step (ahl21.jl:9), so only the last step of the integration needs it: that step
        // also evaluates the pair Jacobians in the trajectory thread (GRAD = true).  Every step's operator records come from
        // pair_op_kernel, so jac_step does not depend on how the integration is cut into chunks or calls.
        const bool last_chunk = done + s == nsteps;
        const int s_light = last_chunk ? s - 1 : s;
        s_split = s;
        if (s_light > 0)
          traj_kernel<false, 2><<<gridA, tpb, 0, p->stream>>>(p->T, n, nsys, h, s_light, p->bstream.as<double>(), p->bscal.as<double>(), detect, ti, t0,
                                                            done, h_intr, p->bntt.as<int32_t>(), Q, evlist, evmask, kahan_time, tkerr, p->kmask);
        if (s_light < s) {
          const size_t o = (size_t)s_light;
          traj_kernel<true, 2><<<gridA, tpb, 0, p->stream>>>(p->T, n, nsys, h, s - s_light, p->bstream.as<double>() + o * sf * ld,
                                                           p->bscal.as<double>() + o * 2 * npairs(n) * SCF * ld, detect, ti, t0, done + s_light, h_intr,
                                                           p->bntt.as<int32_t>(), Q, evlist ? evlist + o * n * ld : nullptr,
                                                           evmask ? evmask + o * ld : nullptr, kahan_time, tkerr, p->kmask);
          if (s_light > 0) p->launches++;
        }
      } else if (grad && kicks) {
        traj_kernel<true, 1, true><<<gridA, tpb, 0, p->stream>>>(p->T, n, nsys, h, s, p->bstream.as<double>(), nullptr, detect, ti, t0, done, h_intr,
                                                               p->bntt.as<int32_t>(), Q, evlist, evmask, kahan_time, tkerr, p->kmask);
      } else if (grad) {
        traj_kernel<true, 1><<<gridA, tpb, 0, p->stream>>>(p->T, n, nsys, h, s, p->bstream.as<double>(), nullptr, detect, ti, t0, done, h_intr,
                                                         p->bntt.as<int32_t>(), Q, evlist, evmask, kahan_time, tkerr, p->kmask);
      } else if (kicks) {
        traj_kernel<false, 0, true><<<gridA, tpb, 0, p->stream>>>(p->T, n, nsys, h, s, nullptr, nullptr, detect, ti, t0, done, h_intr,
                                                                p->bntt.as<int32_t>(), Q, evlist, evmask, kahan_time, tkerr, p->kmask);
      } else {
        traj_kernel<false, 0><<<gridA, tpb, 0, p->stream>>>(p->T, n, nsys, h, s, nullptr, nullptr, detect, ti, t0, done, h_intr, p->bntt.as<int32_t>(),
                                                           Q, evlist, evmask, kahan_time, tkerr, p->kmask);
      }
      tm.end();
      p->launches++;
      if (!detect) break;
      // While the GPU finishes the previous chunk and runs this trajectory kernel, the host scatters the rows of the chunk before
      // the previous one (its copy finished long ago) into the caller's arrays.
      if (attempt == 0 && rows) if (int r = drain_jobs(p, 1)) return r;
      int32_t* hq = (int32_t*)p->hqn.p;
      CK(cudaMemcpyAsync(hq, Q.n, 4, cudaMemcpyDeviceToHost, p->stream));
      CK(cudaStreamSynchronize(p->stream));
      nq = *hq;
      if (nq <= Q.cap) break;
      // more transits than the queue holds: nothing of this chunk has been consumed yet, so grow the queue to the measured count
      // and run the trajectory kernel again from the saved state (deterministic: the same nq transits are found)
      if (attempt >= 2) return fail(NBG_ERR_CUDA, "transit queue overflow persists after the re-run");
      p->chunk_retries++;
      if (int r = alloc_queue(nq + 32)) return r;
    }
    // The operator kernels of the main steps (pair_op, phi_dense) depend only on the trajectory kernel; they run on the aux
    // stream next to the transit refinement (latency-bound, few threads) and join before the Jacobian kernel.
    const bool fork = grad && (s_split > 0 || use_rx) && p->overlap;
    cudaStream_t aux = p->overlap ? p->aux_stream : p->stream;
    if (fork) {
      CK(cudaEventRecord(p->ev_traj, p->stream));
      CK(cudaStreamWaitEvent(p->aux_stream, p->ev_traj, 0));
    }
    if (s_split > 0) {
      tm.begin(6, aux);
      const int py = 4;
      const dim3 grid((unsigned)(ld / TILE), (unsigned)s_split, (unsigned)((2 * npairs(n) + py - 1) / py)), block(TILE, py);
      pair_op_kernel<<<grid, block, 0, aux>>>(p->bscal.as<double>(), p->bstream.as<double>(), n, ld / TILE, nsys, 0.5 * h);
      tm.end();
      p->launches++;
    }
    const bool fork3 = fork && p->overlap3;
    if (grad && use_rx) {
      cudaStream_t aux2 = fork3 ? p->aux2_stream : aux;
      if (fork3) CK(cudaStreamWaitEvent(p->aux2_stream, p->ev_traj, 0));
      tm.begin(5, aux2);
      if (launch_phi_dense(aux2, n, p->bstream.as<double>(), ld / TILE, nsys, nullptr, s, kicks, p->phi_cached)) return fail(NBG_ERR_CUDA, "phi_dense launch failed");
      tm.end();
      p->launches++;
      if (fork3) CK(cudaEventRecord(p->ev_ops2, aux2));
    }
    if (fork) CK(cudaEventRecord(p->ev_ops, aux));
    TransitOut O{};
    int buf = -1;
    if (detect) {
      // per-transit buffers from the measured count
      const size_t nqp = (size_t)round32(std::max<long>(nq, 1));
      if (grad && (p->qstream.ensure_grow(sf * nqp * 8) || p->qz.ensure_grow(nqp * C * M * 8)))
        return fail(NBG_ERR_NOMEM, "transit operator stream allocation failed");
      Q.stream = p->qstream.as<double>();
      Q.z = p->qz.as<double>();
      O = TransitOut{p->btt.as<double>(), grad ? p->bdtdq0.as<double>() : nullptr, p->bntt.as<int32_t>(), p->boff.as<int32_t>(), p->RT, p->C, 0,
                     nullptr, nullptr, 0, nullptr, nullptr};
      if (rows) {
        int bad = p->bevtt.ensure_grow(nqp * C * 8);
        if (grad) bad |= p->bevd.ensure_grow(nqp * M * C * 8);
        if (grad && p->sink.want_dtde) bad |= p->beve.ensure_grow(nqp * M * C * 8);
        buf = (int)(p->chunk_seq % 3);
        const StageLayout L(nq, M, C, grad, grad && p->sink.want_dtde);
        bad |= p->hstage[buf].ensure(L.total + 64);
        if (bad) return fail(NBG_ERR_NOMEM, "transit row buffers allocation failed");
        O.tt = p->bevtt.as<double>();
        O.dtdq0 = grad ? p->bevd.as<double>() : nullptr;
        O.ev_rows = 1;
        // the previous chunk's rows may still be on their way to the host
        if (prev_buf >= 0) CK(cudaStreamWaitEvent(p->stream, p->ev_copied[prev_buf], 0));
        if (nq > 0) {  // who the rows belong to: copied now, before the next trajectory kernel reuses the queue
          char* hb = (char*)p->hstage[buf].p;
          CK(cudaMemcpyAsync(hb + L.o_sys, Q.sys, (size_t)nq * 4, cudaMemcpyDeviceToHost, p->stream));
          CK(cudaMemcpyAsync(hb + L.o_body, Q.body, (size_t)nq * 4, cudaMemcpyDeviceToHost, p->stream));
          CK(cudaMemcpyAsync(hb + L.o_k, Q.k, (size_t)nq * 4, cudaMemcpyDeviceToHost, p->stream));
        }
      } else if (fused) {
        O.tt = p->fused_tt ? p->btt.as<double>() : nullptr;
        O.dtdq0 = nullptr;
        O.tobs = p->btobs.as<double>(); O.sigma = p->bsigma.as<double>(); O.per_system = p->fused_per_system;
        O.chi2 = p->bchi2.as<double>(); O.gq = p->bgq.as<double>();
      } else if (p->fused) {  // chi^2 only (grad = 0): summed by the transit kernel
        O.tt = p->fused_tt ? p->btt.as<double>() : nullptr;
        O.tobs = p->btobs.as<double>(); O.sigma = p->bsigma.as<double>(); O.per_system = p->fused_per_system;
        O.chi2 = p->bchi2.as<double>();
      }
      if (nq > 0) {
        tm.begin(1);
        const unsigned gridT = (unsigned)((nq + tpb - 1) / tpb);
        if (grad && kicks) transit_kernel<true, true><<<gridT, tpb, 0, p->stream>>>(p->T, n, Q, ti, O, dcount, p->kmask, p->newton_pre);
        else if (grad) transit_kernel<true><<<gridT, tpb, 0, p->stream>>>(p->T, n, Q, ti, O, dcount, p->kmask, p->newton_pre);
        else if (kicks) transit_kernel<false, true><<<gridT, tpb, 0, p->stream>>>(p->T, n, Q, ti, O, dcount, p->kmask, p->newton_pre);
        else transit_kernel<false><<<gridT, tpb, 0, p->stream>>>(p->T, n, Q, ti, O, dcount, p->kmask, p->newton_pre);
        tm.end();
        p->launches++;
        if (grad) {
          // dense phisalpha operator of every transit's final step, then the adjoint vectors z = T^T w of the sub-step (nbg_adjoint.cuh):
          // the Jacobian kernel turns them into dt/dq0 (or the chi^2 gradient) with one dot product per column
          tm.begin(5);
          if (launch_phi_dense(p->stream, n, Q.stream, 0, nq, nullptr, 1, kicks, p->phi_cached)) return fail(NBG_ERR_CUDA, "phi_dense launch failed");
          tm.end();
          tm.begin(7);
          const unsigned gridZ = (unsigned)((nq + 63) / 64);
          if (C == 3) transit_adjoint_kernel<3><<<gridZ, 64, 0, p->stream>>>(Q, (int)nq, n, ti, p->kmask);
          else transit_adjoint_kernel<1><<<gridZ, 64, 0, p->stream>>>(Q, (int)nq, n, ti, p->kmask);
          tm.end();
          p->launches += 2;
        }
      }
    }
    if (fork) CK(cudaStreamWaitEvent(p->stream, p->ev_ops, 0));
    if (fork3 && grad && use_rx) CK(cudaStreamWaitEvent(p->stream, p->ev_ops2, 0));
    if (grad) {
      p->counters_host[6] = (unsigned long long)S;
      p->counters_host[7] += 1;
      tm.begin(2);
      if (use_rx) {
        const int32_t* evl = detect ? evlist : nullptr;
        const uint32_t* evm = detect ? evmask : nullptr;
        double *Jv = p->bJv.as<double>(), *Je = p->bJe.as<double>();
        const double* strm = p->bstream.as<double>();
        cudaStream_t st = p->stream;
        int rc = 0;
        if (kicks) rc = launch_jac_rx_kicked(n, st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O, p->kmask);
        else switch (n) {
          case 2: rc = launch_jac_rx<2, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 3: rc = launch_jac_rx<3, 3>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 4: rc = launch_jac_rx<4, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 5: rc = launch_jac_rx<5, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 6: rc = launch_jac_rx<6, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 7: rc = launch_jac_rx<7, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 9: rc = launch_jac_rx<9, 1, true, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 10:  // two blocks of 5 warps per SM at 168 registers (spills ~45 doubles): measured 1.25x faster than one block at 255
            rc = launch_jac_rx<10, 1, true, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O);
            break;
          case 11: rc = launch_jac_rx<11, 1, true, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 12: rc = launch_jac_rx<12, 1, true, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 13: rc = launch_jac_rx<13, 1, true, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 14: rc = launch_jac_rx<14, 1, true, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          case 15: rc = launch_jac_rx<15, 1, true, 1, false, false>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;   // single operator buffer
          case 16: rc = launch_jac_rx<16, 1, true, 1, false, false>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break;
          default:
#ifdef NBG_EXPERIMENTS
            // measured and rejected: the DMMA kernel, pivot blocks of 2 with a barrier per group (22), pivot blocks of 2 at
            // 3 blocks/SM (other values)
            if (p->jac_mma == 1) { rc = launch_jac_mma<8, 2, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break; }
            if (p->jac_mma == 2) { rc = launch_jac_mma<8, 2, 1>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break; }
            if (p->rx_unroll == 22) { rc = launch_jac_rx<8, 2, true, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break; }
            if (p->rx_unroll == 23) { rc = launch_jac_rx<8, 4, false, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break; }
            if (p->rx_unroll == 24) { rc = launch_jac_rx<8, 8, true, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break; }
            if (p->rx_unroll == 99) { KMask notr; notr.w[3] = 0x80000000u; rc = launch_jac_rx<8, 8, false, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O, notr); break; }
            if (p->rx_unroll != 38) { rc = launch_jac_rx<8, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O); break; }
#endif
            // NBG_U8 = 4 pivot bodies per block of the pair sweeps, no per-group barrier, 2 blocks/SM at 250 registers.  The full unroll
            // (U = 8: 110 KB of straight-line code per step, the r01 choice) is at the edge of the instruction cache: the SAME kernel source
            // took 222, 226 or 238 ms per bench window depending on what else changed in the file, U = 4 (60 KB) 219-223 ms in every
            // build (profiles/r02f_ab.jsonl, r02g_ab.jsonl); U = 2 is 243 ms (rotation moves).
            rc = launch_jac_rx<8, NBG_U8, false, 2>(st, nsys, Jv, Je, ld, strm, s, h, evl, evm, Q, ti, O);
            break;
        }
        if (rc) return fail(NBG_ERR_CUDA, "jac_rx_kernel attribute setup failed");
      } else {
        jac_kernel<<<(unsigned)nsys, tps, smem, p->stream>>>(p->bJv.as<double>(), p->bJe.as<double>(), n, ld, p->bstream.as<double>(), s, h,
                                                            detect ? evlist : nullptr, Q, ti, O, stage_phi ? 1 : 0);
      }
      tm.end();
      p->launches++;
    }
    if (rows) {
      // this chunk's rows: dtdelements of the rows, then everything to pinned staging on the copy stream while the next chunk computes
      const StageLayout L(nq, M, C, grad, grad && p->sink.want_dtde);
      if (nq > 0 && grad && p->sink.want_dtde) {
        if (!waited_jinit) { CK(cudaStreamWaitEvent(p->stream, p->copy_done, 0)); waited_jinit = true; }  // jac_init: uploaded / computed on the copy stream
        dtde_rows_kernel<<<(unsigned)nq, 64, M * C * 8, p->stream>>>(p->bevd.as<double>(), p->bjinit.as<double>(), p->beve.as<double>(), Q.sys, (int)nq,
                                                                      (int)M, (int)C);
        p->launches++;
      }
      CK(cudaEventRecord(p->ev_chunk, p->stream));
      CK(cudaStreamWaitEvent(p->copy_stream, p->ev_chunk, 0));
      if (nq > 0) {
        char* hb = (char*)p->hstage[buf].p;
        CK(cudaMemcpyAsync(hb + L.o_tt, p->bevtt.p, (size_t)nq * C * 8, cudaMemcpyDeviceToHost, p->copy_stream));
        if (grad) CK(cudaMemcpyAsync(hb + L.o_d, p->bevd.p, (size_t)nq * M * C * 8, cudaMemcpyDeviceToHost, p->copy_stream));
        if (grad && p->sink.want_dtde) CK(cudaMemcpyAsync(hb + L.o_e, p->beve.p, (size_t)nq * M * C * 8, cudaMemcpyDeviceToHost, p->copy_stream));
      }
      CK(cudaEventRecord(p->ev_copied[buf], p->copy_stream));
      p->jobs.push_back(nbg_plan::Job{buf, (int)nq});
      prev_buf = buf;
      p->chunk_seq++;
    }
    CK(cudaGetLastError());
    done += s;
    p->counters_host[0] += (unsigned long long)nsys * s;
    if (grad) p->counters_host[5] += (unsigned long long)nsys * s;
  }
  if (rows) if (int r = drain_jobs(p, 0)) return r;
  return 0;
}

int ensure_jac(nbg_plan* p) {
  const size_t jsz = (size_t)6 * p->n * 7 * p->n;
  if (p->bJv.ensure((size_t)p->nsys * jsz * 8) || p->bJe.ensure((size_t)p->nsys * jsz * 8)) return fail(NBG_ERR_NOMEM, "jac_step allocation failed");
  return 0;
}

// s.pair (Integrator.jl:91): Julia column-major N x N Bool, entry [i,j] at i + n*j; the reference only reads i < j
int pair_mask(const uint8_t* pair, int n, KMask* mask) {
  *mask = KMask{};
  if (!pair) return 0;
  for (int i = 0; i < n - 1; ++i)
    for (int j = i + 1; j < n; ++j)
      if (pair[i + n * j]) mask->set(i * n - i * (i + 1) / 2 + (j - i - 1));
  return 0;
}

// hierarchy matrix for the device IC / elements kernels: eps (Julia column-major n x n) or NULL = fully nested, plus the elements row of
// each Keplerian
int make_hierarchy(IcsHierarchy* Hp, const double* eps, int ni) {
  IcsHierarchy& H = *Hp;
  const size_t n = (size_t)ni;
  std::memset(&H, 0, sizeof(H));
  if (eps) {
    for (size_t q = 0; q < n * n; ++q) H.eps[q] = eps[q];
  } else {  // fully nested: ElementsIC(t0, N::Int, elements) -> hierarchy([N, ones(N-1)...])  (setup_hierarchy.jl:9-29)
    for (size_t i = 0; i + 1 < n; ++i) {
      for (size_t j = 0; j <= i; ++j) H.eps[i + n * j] = -1.0;
      H.eps[i + n * (i + 1)] = 1.0;
    }
    for (size_t j = 0; j < n; ++j) H.eps[(n - 1) + n * j] = -1.0;
  }
  {  // elements row of each Keplerian: the i+1+b bookkeeping of kepcalc (init_nbody.jl:66-103), a function of eps alone
    int i = 1, b = 0;
    while (i < (int)n) {
      if (H.eps[(i - 1) + 0] == 0.0) b += 1;
      const int row = i + b;
      if (row < 0 || row >= (int)n) return fail(NBG_ERR_ARG, "hierarchy matrix does not map Keplerians to element rows");
      H.row[i - 1] = row;
      if (b > 0) b -= 2; else if (b < 0) b = 0;
      i += 1;
    }
  }
  return 0;
}

// ---- multi-device plans -----------------------------------------------------------------------------------------------------------
// f(kid, first system of its slice, systems in its slice) runs on one host thread per child plan (= per device slice); the first
// failure (code and message) is reported to the caller's thread.
template <class F> int for_kids(nbg_plan* p, F&& f) {
  const size_t K = p->kids.size();
  std::vector<int> rc(K, 0);
  std::vector<std::string> err(K);
  std::vector<std::thread> th;
  for (size_t k = 0; k < K; ++k)
    th.emplace_back([&, k]() {
      rc[k] = f(p->kids[k], p->kid_lo[k], p->kids[k]->nsys);
      if (rc[k]) err[k] = g_err;
    });
  for (auto& t : th) t.join();
  for (size_t k = 0; k < K; ++k)
    if (rc[k]) return fail(rc[k], "device slice " + std::to_string(k) + ": " + err[k]);
  return 0;
}
template <class T> T* at(T* base, size_t off) { return base ? base + off : nullptr; }

}  // namespace

extern "C" {

int32_t nbg_version(void) { return 200; }
const char* nbg_last_error(void) { return g_err.c_str(); }
#ifndef NBG_SRC_HASH
#define NBG_SRC_HASH "unknown"
#endif
const char* nbg_source_hash(void) { return NBG_SRC_HASH; }
int32_t nbg_build_flags(void) {
#ifdef NBG_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}
int32_t nbg_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
  return c;
}

int32_t nbg_plan_create(nbg_plan** out, int32_t nbody, int64_t nsys, int32_t device, int64_t stream_budget_bytes) {
  if (!out) return fail(NBG_ERR_ARG, "plan pointer is NULL");
  *out = nullptr;
  if (nbody < 2 || nbody > NBG_MAX_BODIES) return fail(NBG_ERR_ARG, "nbody must be in 2..16");
  if (nsys < 1) return fail(NBG_ERR_ARG, "nsys must be >= 1");
  int ndev = nbg_device_count();
  if (ndev == 0) return fail(NBG_ERR_NO_DEVICE, "no CUDA device: libnbgrad_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(NBG_ERR_ARG, "device index out of range");
  CK(cudaSetDevice(device));
  nbg_plan* p = new nbg_plan();
  p->n = nbody; p->nsys = nsys; p->device = device;
  p->ld = (size_t)((nsys + 31) / 32 * 32);
  struct Guard { nbg_plan* p; ~Guard() { if (p) nbg_plan_destroy(p); } } guard{p};  // releases everything on an early return
  CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&p->copy_done, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&p->ev_traj, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&p->ev_ops, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&p->ev_ops2, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&p->ev_chunk, cudaEventDisableTiming));
  for (auto& e : p->ev_copied) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&p->aux2_stream, cudaStreamNonBlocking));
  if (stream_budget_bytes <= 0) {
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot));
    stream_budget_bytes = (int64_t)(fr / 4);
  }
  p->stream_budget = stream_budget_bytes;
  if (const char* e = getenv("NBG_FORCE_GENERIC_JAC")) p->force_generic_jac = (e[0] == '1');
  if (const char* e = getenv("NBG_RX_UNROLL")) p->rx_unroll = atoi(e);
  if (const char* e = getenv("NBG_JAC_MMA")) p->jac_mma = atoi(e);
  if (const char* e = getenv("NBG_PHI_CACHED")) p->phi_cached = atoi(e);
  if (const char* e = getenv("NBG_SPLIT_TRAJ")) p->split_traj = (e[0] != '0');
  if (const char* e = getenv("NBG_OVERLAP")) { p->overlap = (e[0] != '0'); p->overlap3 = (e[0] == '2'); }   // 0: operator kernels on the main stream (clean per-kernel times)
  if (const char* e = getenv("NBG_NEWTON_PRE")) p->newton_pre = std::max(0, std::min(8, atoi(e)));
  if (const char* e = getenv("NBG_TRACE")) p->trace = (e[0] == '1');
  if (const char* e = getenv("NBG_QUEUE_CAP0")) p->queue_cap0 = std::max(0L, atol(e));
  if (const char* e = getenv("NBG_SCATTER_THREADS")) p->scatter_threads = (unsigned)std::max(1, std::min(64, atoi(e)));
  if (alloc_state(p)) return fail(NBG_ERR_NOMEM, "state allocation failed");
  CK(cudaMemsetAsync(p->bcounters.p, 0, 64, p->stream));
  guard.p = nullptr;
  *out = p;
  return NBG_OK;
}

// One plan over several devices: contiguous slices of the batch, one child plan + one host thread per entry of
// `devices`; every call on the parent runs on all slices concurrently and reads / writes the slices of the caller's arrays.  The
// same device may be listed more than once (slices then share it).
int32_t nbg_plan_create_multi(nbg_plan** out, int32_t nbody, int64_t nsys, const int32_t* devices, int32_t ndev, int64_t stream_budget_bytes) {
  if (!out) return fail(NBG_ERR_ARG, "plan pointer is NULL");
  *out = nullptr;
  if (!devices || ndev < 1) return fail(NBG_ERR_ARG, "devices[ndev] is required");
  if (nsys < ndev) return fail(NBG_ERR_ARG, "fewer systems than device slices");
  if (ndev == 1) return nbg_plan_create(out, nbody, nsys, devices[0], stream_budget_bytes);
  nbg_plan* p = new nbg_plan();
  p->n = nbody; p->nsys = nsys; p->device = devices[0];
  struct Guard { nbg_plan* p; ~Guard() { if (p) nbg_plan_destroy(p); } } guard{p};
  for (int k = 0; k < ndev; ++k) {
    const long lo = (long)(nsys * k / ndev), hi = (long)(nsys * (k + 1) / ndev);
    int share = 0;
    for (int q = 0; q < ndev; ++q) share += devices[q] == devices[k];
    int64_t budget = stream_budget_bytes > 0 ? stream_budget_bytes / share : 0;
    if (budget == 0 && share > 1) {  // the default (1/4 of the free memory) divided among the slices that share the device
      int nd = nbg_device_count();
      if (devices[k] < 0 || devices[k] >= nd) return fail(nd ? NBG_ERR_ARG : NBG_ERR_NO_DEVICE, nd ? "device index out of range" : "no CUDA device: libnbgrad_b200 has no CPU fallback");
      CK(cudaSetDevice(devices[k]));
      size_t fr = 0, tot = 0;
      CK(cudaMemGetInfo(&fr, &tot));
      budget = (int64_t)(tot / 4 / share);
      budget = std::min<int64_t>(budget, (int64_t)(fr / 2));
    }
    nbg_plan* kid = nullptr;
    if (int r = nbg_plan_create(&kid, nbody, hi - lo, devices[k], budget)) return r;
    p->kids.push_back(kid);
    p->kid_lo.push_back(lo);
  }
  guard.p = nullptr;
  *out = p;
  return NBG_OK;
}

int32_t nbg_plan_destroy(nbg_plan* p) {
  if (!p) return NBG_OK;
  if (!p->kids.empty()) {
    for (nbg_plan* k : p->kids) nbg_plan_destroy(k);
    delete p;
    return NBG_OK;
  }
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
  if (p->aux_stream) cudaStreamSynchronize(p->aux_stream);
  if (p->aux2_stream) cudaStreamSynchronize(p->aux2_stream);
  DevBuf* all[] = {&p->bx, &p->bv, &p->bxe, &p->bve, &p->bm, &p->bdq, &p->bgs, &p->bt, &p->bterr, &p->bcount, &p->bstatus, &p->bbackup, &p->bJv, &p->bJe,
                   &p->qz, &p->bstream, &p->bscal, &p->bevlist, &p->bevmask, &p->qn, &p->qsys, &p->qstep, &p->qbody, &p->qk, &p->qdt0, &p->qt, &p->qsnap,
                   &p->qhdr, &p->qstream, &p->btt, &p->bdtdq0, &p->bdtde, &p->bjinit, &p->bntt, &p->boff, &p->bcounters, &p->belem, &p->bevtt, &p->bevd,
                   &p->beve, &p->bchi2, &p->bgq, &p->btobs, &p->bsigma};
  for (auto* b : all) b->release();
  for (auto& b : p->stage) b.release();
  for (auto& b : p->hstage) b.release();
  p->hqn.release();
  if (p->stream) cudaStreamDestroy(p->stream);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->copy_done) cudaEventDestroy(p->copy_done);
  if (p->aux_stream) cudaStreamDestroy(p->aux_stream);
  if (p->aux2_stream) cudaStreamDestroy(p->aux2_stream);
  if (p->ev_ops2) cudaEventDestroy(p->ev_ops2);
  if (p->ev_traj) cudaEventDestroy(p->ev_traj);
  if (p->ev_ops) cudaEventDestroy(p->ev_ops);
  if (p->ev_chunk) cudaEventDestroy(p->ev_chunk);
  for (cudaEvent_t e : p->ev_copied) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : p->tev) cudaEventDestroy(e);
  p->tev.clear();
  cudaGetLastError();
  delete p;
  return NBG_OK;
}

int32_t nbg_set_pair(nbg_plan* p, const uint8_t* pair) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) {
    for (nbg_plan* k : p->kids) if (int r = nbg_set_pair(k, pair)) return r;
    return NBG_OK;
  }
  KMask mask;
  if (int r = pair_mask(pair, p->n, &mask)) return r;
  p->kmask = mask;
  return NBG_OK;
}

int32_t nbg_set_state(nbg_plan* p, const double* x, const double* v, const double* m, double t0, const double* xerror, const double* verror,
                      const double* jac_step, const double* jac_error, const double* dqdt) {
  if (!p || !x || !v || !m) return fail(NBG_ERR_ARG, "plan, x, v, m are required");
  if (!p->kids.empty()) {
    const size_t n = p->n, M = 7 * n;
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_set_state(k, x + lo * 3 * n, v + lo * 3 * n, m + lo * n, t0, at(xerror, lo * 3 * n), at(verror, lo * 3 * n), at(jac_step, lo * M * M),
                           at(jac_error, lo * M * M), at(dqdt, lo * M));
    });
  }
  CK(cudaSetDevice(p->device));
  p->generation++;
  const size_t n = p->n, nsys = p->nsys, M = 7 * n;
  const double* src[6] = {x, v, m, xerror, verror, dqdt};
  const size_t cnt[6] = {3 * n, 3 * n, n, 3 * n, 3 * n, M};
  double* dev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int q = 0; q < 6; ++q) {
    if (!src[q]) continue;
    if (p->stage[q].ensure(cnt[q] * nsys * 8)) return fail(NBG_ERR_NOMEM, "staging allocation failed");
    CK(cudaMemcpyAsync(p->stage[q].p, src[q], cnt[q] * nsys * 8, cudaMemcpyHostToDevice, p->stream));
    dev[q] = p->stage[q].as<double>();
  }
  const int tpb = 128;
  pack_xvm_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->stream>>>(dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], p->T, (int)n, (long)nsys, t0);
  p->launches++;
  CK(cudaMemsetAsync(p->bterr.p, 0, p->ld * 8, p->stream));
  // jac_step / jac_error: lazily created as identity / zero unless supplied
  p->jac_valid = false;
  if (jac_step || jac_error) {
    if (int r = ensure_jac(p)) return r;
    const double* js[2] = {jac_step, jac_error};
    double* dst[2] = {p->bJv.as<double>(), p->bJe.as<double>()};
    for (int q = 0; q < 2; ++q) {
      double* d = nullptr;
      if (js[q]) {
        if (p->stage[6].ensure(M * M * nsys * 8)) return fail(NBG_ERR_NOMEM, "staging allocation failed");
        CK(cudaMemcpyAsync(p->stage[6].p, js[q], M * M * nsys * 8, cudaMemcpyHostToDevice, p->stream));
        d = p->stage[6].as<double>();
      }
      jac_from_julia_kernel<<<(unsigned)nsys, 256, 0, p->stream>>>(d, dst[q], (int)n, q);
      p->launches++;
      CK(cudaStreamSynchronize(p->stream));
    }
    p->jac_valid = true;
  }
  CK(cudaStreamSynchronize(p->stream));
  p->has_state = true;
  p->have_transit = false;
  p->jinit_resident = false;
  return NBG_OK;
}

// State(ic::ElementsIC) on the device: init_nbody (src/ics/init_nbody.jl:13-27) for every system of the batch.
int32_t nbg_set_state_elements(nbg_plan* p, const double* elements, const double* eps, double t0, int32_t want_jac_init) {
  if (!p || !elements) return fail(NBG_ERR_ARG, "plan and elements are required");
  if (!p->kids.empty())
    return for_kids(p, [&](nbg_plan* k, long lo, long) { return nbg_set_state_elements(k, elements + (size_t)lo * 7 * p->n, eps, t0, want_jac_init); });
  CK(cudaSetDevice(p->device));
  p->generation++;
  const size_t n = p->n, nsys = p->nsys, M = 7 * n;
  IcsHierarchy H;
  if (int r = make_hierarchy(&H, eps, (int)n)) return r;
  CK(cudaStreamSynchronize(p->copy_stream));  // a previous jac_init kernel may still read the elements buffer
  if (p->belem.ensure(7 * n * nsys * 8)) return fail(NBG_ERR_NOMEM, "elements allocation failed");
  CK(cudaMemcpyAsync(p->belem.p, elements, 7 * n * nsys * 8, cudaMemcpyHostToDevice, p->stream));
  if (want_jac_init && p->bjinit.ensure(nsys * M * M * 8)) return fail(NBG_ERR_NOMEM, "jac_init allocation failed");
  // errors, dq/dh, status, time: the State(ic) defaults (Integrator.jl:82-103)
  CK(cudaMemsetAsync(p->bxe.p, 0, 3 * n * p->ld * 8, p->stream));
  CK(cudaMemsetAsync(p->bve.p, 0, 3 * n * p->ld * 8, p->stream));
  CK(cudaMemsetAsync(p->bdq.p, 0, 6 * n * p->ld * 8, p->stream));
  CK(cudaMemsetAsync(p->bstatus.p, 0, p->ld * 4, p->stream));
  CK(cudaMemsetAsync(p->bterr.p, 0, p->ld * 8, p->stream));
  std::vector<double> tv(p->ld, t0);
  CK(cudaMemcpyAsync(p->bt.p, tv.data(), p->ld * 8, cudaMemcpyHostToDevice, p->stream));
  const int tpb = 64;
  // x, v, m now; jac_init (only needed by the dtdelements kernel at the end of a transit-timing call) on the copy stream, so that
  // it overlaps the stepping -- nbg_transit_timing_resident / nbg_get_jac_init wait for copy_done
  ics_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->stream>>>(p->belem.as<double>(), H, (int)n, (long)nsys, p->ld, t0, p->T.x, p->T.v, p->T.m,
                                                                         nullptr, 1);
  p->launches++;
  CK(cudaStreamSynchronize(p->stream));
  if (want_jac_init) {
    ics_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->copy_stream>>>(p->belem.as<double>(), H, (int)n, (long)nsys, p->ld, t0, p->T.x, p->T.v,
                                                                                p->T.m, p->bjinit.as<double>(), 0);
    p->launches++;
    CK(cudaEventRecord(p->copy_done, p->copy_stream));
  }
  CK(cudaGetLastError());
  p->jac_valid = false;
  p->has_state = true;
  p->have_transit = false;
  p->jinit_resident = want_jac_init != 0;
  return NBG_OK;
}

int32_t nbg_get_jac_init(nbg_plan* p, double* jac_init) {
  if (!p || !jac_init) return fail(NBG_ERR_ARG, "NULL argument");
  if (!p->kids.empty()) {
    const size_t M = 7 * (size_t)p->n;
    return for_kids(p, [&](nbg_plan* k, long lo, long) { return nbg_get_jac_init(k, jac_init + (size_t)lo * M * M); });
  }
  if (!p->jinit_resident) return fail(NBG_ERR_ARG, "no device-computed jac_init (call nbg_set_state_elements with want_jac_init)");
  CK(cudaSetDevice(p->device));
  const size_t M = 7 * (size_t)p->n;
  CK(cudaStreamWaitEvent(p->stream, p->copy_done, 0));
  CK(cudaMemcpyAsync(jac_init, p->bjinit.p, p->nsys * M * M * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return NBG_OK;
}

static int make_jac_identity(nbg_plan* p) {
  if (p->jac_valid) return 0;
  if (int r = ensure_jac(p)) return r;
  jac_from_julia_kernel<<<(unsigned)p->nsys, 256, 0, p->stream>>>(nullptr, p->bJv.as<double>(), p->n, 0);
  jac_from_julia_kernel<<<(unsigned)p->nsys, 256, 0, p->stream>>>(nullptr, p->bJe.as<double>(), p->n, 1);
  p->launches += 2;
  p->jac_valid = true;
  return 0;
}

int32_t nbg_get_state(nbg_plan* p, double* x, double* v, double* xerror, double* verror, double* jac_step, double* jac_error, double* dqdt, double* t,
                      uint32_t* status) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) {
    const size_t n = p->n, M = 7 * n;
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_get_state(k, at(x, lo * 3 * n), at(v, lo * 3 * n), at(xerror, lo * 3 * n), at(verror, lo * 3 * n), at(jac_step, lo * M * M),
                           at(jac_error, lo * M * M), at(dqdt, lo * M), at(t, lo), at(status, lo));
    });
  }
  if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
  CK(cudaSetDevice(p->device));
  const size_t n = p->n, nsys = p->nsys, M = 7 * n;
  double* outs[5] = {x, v, xerror, verror, dqdt};
  const size_t cnt[5] = {3 * n, 3 * n, 3 * n, 3 * n, M};
  double* dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int q = 0; q < 5; ++q)
    if (outs[q]) {
      if (p->stage[q].ensure(cnt[q] * nsys * 8)) return fail(NBG_ERR_NOMEM, "staging allocation failed");
      dev[q] = p->stage[q].as<double>();
    }
  const int tpb = 128;
  unpack_xv_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->stream>>>(p->T, (int)n, (long)nsys, dev[0], dev[1], dev[2], dev[3], dev[4]);
  p->launches++;
  for (int q = 0; q < 5; ++q)
    if (outs[q]) CK(cudaMemcpyAsync(outs[q], dev[q], cnt[q] * nsys * 8, cudaMemcpyDeviceToHost, p->stream));
  if (jac_step || jac_error) {
    if (int r = make_jac_identity(p)) return r;
    double* js[2] = {jac_step, jac_error};
    const double* srcs[2] = {p->bJv.as<double>(), p->bJe.as<double>()};
    for (int q = 0; q < 2; ++q)
      if (js[q]) {
        if (p->stage[6].ensure(M * M * nsys * 8)) return fail(NBG_ERR_NOMEM, "staging allocation failed");
        jac_to_julia_kernel<<<(unsigned)nsys, 256, 0, p->stream>>>(srcs[q], p->stage[6].as<double>(), (int)n, q);
        p->launches++;
        CK(cudaMemcpyAsync(js[q], p->stage[6].p, M * M * nsys * 8, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaStreamSynchronize(p->stream));
      }
  }
  if (t) CK(cudaMemcpyAsync(t, p->bt.p, nsys * 8, cudaMemcpyDeviceToHost, p->stream));
  if (status) CK(cudaMemcpyAsync(status, p->bstatus.p, nsys * 4, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return NBG_OK;
}

static void finish_timings(nbg_plan* p, Timer& tm) {
  for (double& q : p->timings) q = 0;
  tm.collect(p->timings);
  const double tot = p->timings[4];
  // pair_op / phi_dense run on the aux stream concurrently with the transit kernel, so the per-kernel times can add up to more
  // than the total; "other" is what is left of the total, never negative
  p->timings[3] = std::max(0.0, tot - p->timings[0] - p->timings[1] - p->timings[2] - p->timings[5] - p->timings[6] - p->timings[7]);
  unsigned long long dc[8];
  cudaMemcpy(dc, p->bcounters.p, 64, cudaMemcpyDeviceToHost);
  for (int q = 1; q <= 4; ++q) p->counters_host[q] = dc[q];
}

int32_t nbg_integrate_resident(nbg_plan* p, double h, int64_t nsteps, double h_last, int32_t grad, int32_t time_mode, double t_final) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) return for_kids(p, [&](nbg_plan* k, long, long) { return nbg_integrate_resident(k, h, nsteps, h_last, grad, time_mode, t_final); });
  if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
  if (nsteps < 0) return fail(NBG_ERR_ARG, "nsteps must be >= 0");
  CK(cudaSetDevice(p->device));
  p->generation++;
  if (grad) if (int r = make_jac_identity(p)) return r;
  if (time_mode == 0) CK(cudaMemsetAsync(p->bterr.p, 0, p->ld * 8, p->stream));  // s2 = zero(T): Integrator.jl:212
  Timer tm(p);
  tm.start();
  if (int r = run_steps(p, h, (long)nsteps, grad != 0, false, 0, 0.0, h, time_mode == 0, tm, 0.0)) return r;
  if (h_last != 0.0)
    if (int r = run_steps(p, h_last, 1, grad != 0, false, 0, 0.0, h_last, time_mode == 0, tm, 0.0)) return r;
  if (time_mode == 1) {
    std::vector<double> tv(p->nsys, t_final);
    CK(cudaMemcpyAsync(p->bt.p, tv.data(), p->nsys * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
  }
  tm.stop();
  CK(cudaStreamSynchronize(p->stream));
  finish_timings(p, tm);
  CK(cudaGetLastError());
  return NBG_OK;
}

// (intr)(s, o::CartesianOutput) (Outputs.jl:26-49) without the per-step host round trip: positions and velocities before every
// `stride`-th step are collected on the device and copied out once.
int32_t nbg_integrate_sampled_jac(nbg_plan* p, double h, int64_t nsteps, int64_t stride, int32_t grad, double* x_samples, double* v_samples,
                                  double* jac_samples) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (nsteps < 1 || stride < 1 || !x_samples || !v_samples) return fail(NBG_ERR_ARG, "nsteps >= 1, stride >= 1 and both sample buffers are required");
  if (jac_samples && !grad) return fail(NBG_ERR_ARG, "jac_step samples need grad = true");
  if (!p->kids.empty()) {  // a slice's samples are [k][its systems]: collected per slice, then interleaved into [k][all systems]
    const size_t n3 = 3 * (size_t)p->n, MM = (size_t)49 * p->n * p->n, nsamp = (size_t)((nsteps + stride - 1) / stride), B = (size_t)p->nsys;
    return for_kids(p, [&](nbg_plan* k, long lo, long cnt) {
      std::vector<double> xs(nsamp * cnt * n3), vs(nsamp * cnt * n3), js(jac_samples ? nsamp * cnt * MM : 0);
      if (int r = nbg_integrate_sampled_jac(k, h, nsteps, stride, grad, xs.data(), vs.data(), jac_samples ? js.data() : nullptr)) return r;
      for (size_t q = 0; q < nsamp; ++q) {
        std::memcpy(x_samples + (q * B + lo) * n3, xs.data() + q * cnt * n3, cnt * n3 * 8);
        std::memcpy(v_samples + (q * B + lo) * n3, vs.data() + q * cnt * n3, cnt * n3 * 8);
        if (jac_samples) std::memcpy(jac_samples + (q * B + lo) * MM, js.data() + q * cnt * MM, cnt * MM * 8);
      }
      return 0;
    });
  }
  if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
  CK(cudaSetDevice(p->device));
  p->generation++;
  const size_t n = p->n, nsys = p->nsys, ld = p->ld, MM = 49 * n * n;
  const long nsamp = (long)((nsteps + stride - 1) / stride);
  DevBuf sx, sv, ox, sj;
  auto release = [&]() { sx.release(); sv.release(); ox.release(); sj.release(); };
  if (sx.ensure((size_t)nsamp * 3 * n * ld * 8) || sv.ensure((size_t)nsamp * 3 * n * ld * 8) || ox.ensure((size_t)nsamp * 3 * n * nsys * 8) ||
      (jac_samples && sj.ensure((size_t)nsamp * nsys * MM * 8))) {
    release();
    return fail(NBG_ERR_NOMEM, "sample buffers do not fit (reduce nsteps / stride or the batch)");
  }
  double t0 = 0;
  CK(cudaMemcpy(&t0, p->bt.p, 8, cudaMemcpyDeviceToHost));
  if (grad) if (int r = make_jac_identity(p)) { release(); return r; }
  p->T.samp_x = sx.as<double>(); p->T.samp_v = sv.as<double>(); p->T.samp_stride = (long)stride;
  p->samp_jac = jac_samples ? sj.as<double>() : nullptr;
  Timer tm(p);
  tm.start();
  // s.t[1] = t0 + h i (Outputs.jl:43): same time bookkeeping as the transit driver
  const int rc = run_steps(p, h, (long)nsteps, grad != 0, false, 0, t0, h, false, tm, 0.0);
  p->T.samp_x = nullptr; p->T.samp_v = nullptr; p->samp_jac = nullptr;
  tm.stop();
  cudaStreamSynchronize(p->stream);
  finish_timings(p, tm);
  if (rc) { release(); return rc; }
  const dim3 grid((unsigned)((nsys + 127) / 128), (unsigned)nsamp);
  double* outs[2] = {x_samples, v_samples};
  const double* srcs[2] = {sx.as<double>(), sv.as<double>()};
  for (int q = 0; q < 2; ++q) {
    unpack_samples_kernel<<<grid, 128, 0, p->stream>>>(srcs[q], ox.as<double>(), (int)n, (long)nsys, ld, nsamp);
    p->launches++;
    cudaMemcpyAsync(outs[q], ox.p, (size_t)nsamp * 3 * n * nsys * 8, cudaMemcpyDeviceToHost, p->stream);
    cudaStreamSynchronize(p->stream);
  }
  if (jac_samples) {
    cudaMemcpyAsync(jac_samples, sj.p, (size_t)nsamp * nsys * MM * 8, cudaMemcpyDeviceToHost, p->stream);
    cudaStreamSynchronize(p->stream);
  }
  const cudaError_t err = cudaGetLastError();
  release();
  if (err != cudaSuccess) return fail(NBG_ERR_CUDA, cudaGetErrorString(err));
  return NBG_OK;
}
int32_t nbg_integrate_sampled(nbg_plan* p, double h, int64_t nsteps, int64_t stride, int32_t grad, double* x_samples, double* v_samples) {
  return nbg_integrate_sampled_jac(p, h, nsteps, stride, grad, x_samples, v_samples, nullptr);
}

// get_orbital_elements(s, ic) (src/outputs/elements.jl:108-137) for the resident state of every system: the adjacent step after the path
// for RV / astrometry consumers.  eps: hierarchy matrix (Julia column-major n x n) or NULL = fully nested.
int32_t nbg_orbital_elements(nbg_plan* p, const double* eps, double* elements_out) {
  if (!p || !elements_out) return fail(NBG_ERR_ARG, "NULL argument");
  if (!p->kids.empty()) return for_kids(p, [&](nbg_plan* k, long lo, long) { return nbg_orbital_elements(k, eps, elements_out + (size_t)lo * p->n * 11); });
  if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
  CK(cudaSetDevice(p->device));
  IcsHierarchy H;
  if (int r = make_hierarchy(&H, eps, p->n)) return r;
  const size_t nsys = p->nsys, bytes = nsys * p->n * 11 * 8;
  if (p->stage[6].ensure(bytes)) return fail(NBG_ERR_NOMEM, "staging allocation failed");
  const int tpb = 64;
  elements_out_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->stream>>>(p->T.x, p->T.v, p->T.m, H, p->n, (long)nsys, p->ld, p->stage[6].as<double>());
  p->launches++;
  CK(cudaMemcpyAsync(elements_out, p->stage[6].p, bytes, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  CK(cudaGetLastError());
  return NBG_OK;
}

int32_t nbg_integrate(nbg_plan* p, const double* x0, const double* v0, const double* m, const uint8_t* pair, double t0, double h, int64_t nsteps,
                      double h_last, int32_t grad, double* x, double* v, double* xerror, double* verror, double* jac_step, double* jac_error,
                      double* dqdt, uint32_t* status) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) {  // every slice runs its whole upload / integrate / download sequence on its own thread
    if (!x0 || !v0 || !m) return fail(NBG_ERR_ARG, "plan, x, v, m are required");
    const size_t n = p->n, M = 7 * n;
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_integrate(k, x0 + lo * 3 * n, v0 + lo * 3 * n, m + lo * n, pair, t0, h, nsteps, h_last, grad, at(x, lo * 3 * n), at(v, lo * 3 * n),
                           at(xerror, lo * 3 * n), at(verror, lo * 3 * n), at(jac_step, lo * M * M), at(jac_error, lo * M * M), at(dqdt, lo * M),
                           at(status, lo));
    });
  }
  if (int r = nbg_set_pair(p, pair)) return r;
  if (int r = nbg_set_state(p, x0, v0, m, t0, nullptr, nullptr, nullptr, nullptr, nullptr)) return r;
  if (int r = nbg_integrate_resident(p, h, nsteps, h_last, grad, 0, 0.0)) return r;
  return nbg_get_state(p, x, v, xerror, verror, grad ? jac_step : nullptr, grad ? jac_error : nullptr, grad ? dqdt : nullptr, nullptr, status);
}

// Shared body of the transit drivers.  out_mode 0: dense device arrays (nbg_transit_fetch / nbg_transit_chi2 read them afterwards);
// 1: per-chunk rows streamed to p->sink; 2: fused chi^2 (p->btobs / bsigma / bchi2 / bgq prepared by the caller).
static int transit_run(nbg_plan* p, double h, double tmax, int32_t ti, const int32_t* ntt_body, int32_t mode, int32_t grad, const double* jac_init,
                       int out_mode) {
  if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
  if (!ntt_body) return fail(NBG_ERR_ARG, "ntt_body is required");
  if (ti < 0 || ti >= p->n) return fail(NBG_ERR_ARG, "ti out of range");
  if (mode != 0 && mode != 1) return fail(NBG_ERR_ARG, "mode must be 0 (TransitTiming) or 1 (TransitParameters)");
  if (h == 0.0) return fail(NBG_ERR_ARG, "h must be non-zero");
  CK(cudaSetDevice(p->device));
  p->generation++;
  const int n = p->n;
  const size_t nsys = p->nsys, M = 7 * n;
  int RT = 0;
  for (int i = 0; i < n; ++i) {
    if (ntt_body[i] < 0) return fail(NBG_ERR_ARG, "ntt_body must be >= 0");
    p->ntt_body[i] = ntt_body[i]; p->off[i] = RT; RT += ntt_body[i];
  }
  p->RT = RT; p->C = mode == 1 ? 3 : 1;
  const size_t C = p->C;
  const bool dense = out_mode == 0;
  if (dense || (out_mode == 2 && p->fused_tt)) {
    if (p->btt.ensure(std::max<size_t>(8, nsys * RT * C * 8))) return fail(NBG_ERR_NOMEM, "tt allocation failed");
    CK(cudaMemsetAsync(p->btt.p, 0, nsys * RT * C * 8, p->stream));
  }
  if (grad) {
    if (dense) {
      if (p->bdtdq0.ensure(std::max<size_t>(8, nsys * RT * M * C * 8)))
        return fail(NBG_ERR_NOMEM, "dtdq0 does not fit on the device as a dense array: use nbg_transit_timing (streams rows to host buffers) or nbg_transit_chi2_fused");
      CK(cudaMemsetAsync(p->bdtdq0.p, 0, nsys * RT * M * C * 8, p->stream));
    }
    if (int r = make_jac_identity(p)) return r;
  }
  CK(cudaMemcpyAsync(p->bntt.p, p->ntt_body, n * 4, cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemcpyAsync(p->boff.p, p->off, n * 4, cudaMemcpyHostToDevice, p->stream));
  // t0 is s.t[1] of the resident state (identical for all systems of a batch)
  double t0 = 0;
  CK(cudaMemcpyAsync(&t0, p->bt.p, 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  const long nsteps = std::labs((long)std::nearbyint(tmax / h));         // Transits.jl:143
  const double hs = h * check_step(t0, tmax + t0);                       // Transits.jl:144
  Timer tm(p);
  tm.start();
  const int tpb = 128;
  gsave_init_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->stream>>>(p->T, n, (long)nsys, ti);
  p->launches++;
  // jac_init (M^2 doubles per system, the bulk of the input bytes) is uploaded on a second stream while the steps run
  const bool want_dtde = out_mode != 2 && grad && (jac_init || p->jinit_resident);
  if (want_dtde) {
    if (jac_init && p->bjinit.ensure(nsys * M * M * 8)) return fail(NBG_ERR_NOMEM, "jac_init allocation failed");
    if (dense && p->bdtde.ensure(std::max<size_t>(8, nsys * RT * M * C * 8))) return fail(NBG_ERR_NOMEM, "dtdelements allocation failed");
    if (jac_init) {
      CK(cudaMemcpyAsync(p->bjinit.p, jac_init, nsys * M * M * 8, cudaMemcpyHostToDevice, p->copy_stream));
      p->jinit_resident = false;
    }  // else: the ics_kernel launched by nbg_set_state_elements on the copy stream produces it
    if (dense) CK(cudaMemsetAsync(p->bdtde.p, 0, nsys * RT * M * C * 8, p->copy_stream));
    CK(cudaEventRecord(p->copy_done, p->copy_stream));
  }
  const double rate = nsteps > 0 ? (double)RT / (double)nsteps : 1.0;
  p->sink.want_dtde = want_dtde;
  p->sink.active = out_mode == 1;
  p->fused = out_mode == 2;
  p->transit_grad = grad != 0;
  p->jobs.clear();
  p->chunk_seq = 0;
  const double wr0 = wall_ms();
  const int rr = run_steps(p, hs, nsteps, grad != 0, true, ti, t0, h, false, tm, rate);
  p->sink.active = false;
  p->fused = false;
  if (rr) { cudaStreamSynchronize(p->copy_stream); cudaStreamSynchronize(p->stream); p->jobs.clear(); return rr; }
  p->have_transit = dense;
  p->have_dtde = false;
  if (dense && want_dtde) {
    CK(cudaStreamWaitEvent(p->stream, p->copy_done, 0));
    CK(cudaFuncSetAttribute(dtdelements_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(M * (M | 1) * 8)));
    dtdelements_kernel<<<(unsigned)nsys, 256, M * (M | 1) * 8, p->stream>>>(p->bdtdq0.as<double>(), p->bjinit.as<double>(), p->bdtde.as<double>(),
                                                                     p->bcount.as<int32_t>(), p->bntt.as<int32_t>(), p->boff.as<int32_t>(), n, p->ld, RT,
                                                                     (int)C, 0);
    p->launches++;
    p->have_dtde = true;
  }
  tm.stop();
  const double wr1 = wall_ms();
  CK(cudaStreamSynchronize(p->stream));
  CK(cudaStreamSynchronize(p->copy_stream));
  const double wr2 = wall_ms();
  finish_timings(p, tm);
  if (p->trace) fprintf(stderr, "[nbg trace] transit run: host loop %.2f ms, then waited %.2f ms, timing readback %.2f ms, chunk re-runs so far %ld\n", wr1 - wr0, wr2 - wr1, wall_ms() - wr2, p->chunk_retries);
  CK(cudaGetLastError());
  return NBG_OK;
}

int32_t nbg_transit_timing_resident(nbg_plan* p, double h, double tmax, int32_t ti, const int32_t* ntt_body, int32_t mode, int32_t grad,
                                    const double* jac_init) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) {
    const size_t M = 7 * (size_t)p->n;
    return for_kids(p, [&](nbg_plan* k, long lo, long) { return nbg_transit_timing_resident(k, h, tmax, ti, ntt_body, mode, grad, at(jac_init, lo * M * M)); });
  }
  return transit_run(p, h, tmax, ti, ntt_body, mode, grad, jac_init, 0);
}

int32_t nbg_transit_fetch(nbg_plan* p, double* tt, int64_t* count, double* dtdq0, double* dtdelements) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) {
    const nbg_plan* k0 = p->kids[0];
    const size_t M = 7 * (size_t)p->n, C = k0->C, RT = k0->RT;
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_transit_fetch(k, at(tt, lo * RT * C), at(count, (size_t)lo * p->n), at(dtdq0, lo * RT * M * C), at(dtdelements, lo * RT * M * C));
    });
  }
  CK(cudaSetDevice(p->device));
  const size_t nsys = p->nsys, M = 7 * p->n, C = p->C, RT = p->RT;
  if ((tt || dtdq0 || dtdelements) && !p->have_transit) return fail(NBG_ERR_ARG, "no dense transit results (run nbg_transit_timing_resident first)");
  if (tt) CK(cudaMemcpyAsync(tt, p->btt.p, nsys * RT * C * 8, cudaMemcpyDeviceToHost, p->stream));
  if (count) {
    if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
    if (p->stage[7].ensure(nsys * p->n * 8)) return fail(NBG_ERR_NOMEM, "staging allocation failed");
    const int tpb = 128;
    count_out_kernel<<<(unsigned)((nsys + tpb - 1) / tpb), tpb, 0, p->stream>>>(p->bcount.as<int32_t>(), p->ld, p->n, (long)nsys, p->stage[7].as<int64_t>());
    p->launches++;
    CK(cudaMemcpyAsync(count, p->stage[7].p, nsys * p->n * 8, cudaMemcpyDeviceToHost, p->stream));
  }
  if (dtdq0 && p->transit_grad) CK(cudaMemcpyAsync(dtdq0, p->bdtdq0.p, nsys * RT * M * C * 8, cudaMemcpyDeviceToHost, p->stream));
  if (dtdelements && p->have_dtde) CK(cudaMemcpyAsync(dtdelements, p->bdtde.p, nsys * RT * M * C * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return NBG_OK;
}

int32_t nbg_transit_chi2(nbg_plan* p, const double* t_obs, const double* sigma, int32_t per_system, double* chi2, double* grad_q0,
                         double* grad_elements) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!t_obs || !sigma || !chi2) return fail(NBG_ERR_ARG, "t_obs, sigma and chi2 are required");
  if (!p->kids.empty()) {
    const size_t M = 7 * (size_t)p->n, RT = p->kids[0]->RT;
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_transit_chi2(k, per_system ? t_obs + lo * RT : t_obs, per_system ? sigma + lo * RT : sigma, per_system, chi2 + lo, at(grad_q0, lo * M),
                              at(grad_elements, lo * M));
    });
  }
  if (!p->have_transit) return fail(NBG_ERR_ARG, "no dense transit results (run nbg_transit_timing_resident first)");
  if (p->C != 1) return fail(NBG_ERR_UNSUPPORTED, "chi^2 is defined for TransitTiming (mode 0) results");
  if (grad_q0 && !p->transit_grad) return fail(NBG_ERR_ARG, "the last transit call ran with grad = 0");
  if (grad_elements && !p->have_dtde) return fail(NBG_ERR_ARG, "no dtdelements (jac_init was not given)");
  CK(cudaSetDevice(p->device));
  const size_t nsys = p->nsys, M = 7 * (size_t)p->n, RT = p->RT, nobs = (per_system ? nsys : 1) * RT;
  if (p->stage[0].ensure(std::max<size_t>(8, nobs * 8)) || p->stage[1].ensure(std::max<size_t>(8, nobs * 8)) || p->stage[2].ensure(nsys * 8) ||
      p->stage[3].ensure(nsys * M * 8) || p->stage[4].ensure(nsys * M * 8))
    return fail(NBG_ERR_NOMEM, "staging allocation failed");
  CK(cudaMemcpyAsync(p->stage[0].p, t_obs, nobs * 8, cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemcpyAsync(p->stage[1].p, sigma, nobs * 8, cudaMemcpyHostToDevice, p->stream));
  const int threads = 32 * (int)((M + 31) / 32);
  chi2_kernel<<<(unsigned)nsys, threads, 0, p->stream>>>(p->btt.as<double>(), p->transit_grad ? p->bdtdq0.as<double>() : nullptr,
                                                          p->have_dtde ? p->bdtde.as<double>() : nullptr, p->bcount.as<int32_t>(), p->bntt.as<int32_t>(),
                                                          p->boff.as<int32_t>(), p->n, p->ld, (int)RT, p->stage[0].as<double>(), p->stage[1].as<double>(),
                                                          per_system ? 1 : 0, p->stage[2].as<double>(), grad_q0 ? p->stage[3].as<double>() : nullptr,
                                                          grad_elements ? p->stage[4].as<double>() : nullptr);
  p->launches++;
  CK(cudaMemcpyAsync(chi2, p->stage[2].p, nsys * 8, cudaMemcpyDeviceToHost, p->stream));
  if (grad_q0) CK(cudaMemcpyAsync(grad_q0, p->stage[3].p, nsys * M * 8, cudaMemcpyDeviceToHost, p->stream));
  if (grad_elements) CK(cudaMemcpyAsync(grad_elements, p->stage[4].p, nsys * M * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  CK(cudaGetLastError());
  return NBG_OK;
}

// Fused transit-time likelihood (SURVEY 8(f) f2 as specified): the whole transit-timing run from the resident state with chi^2 and its
// gradient accumulated where dt/dq0 is produced (the transit branch of the Jacobian kernel); no dtdq0 / dtdelements row is ever stored.
int32_t nbg_transit_chi2_fused(nbg_plan* p, double h, double tmax, int32_t ti, const int32_t* ntt_body, const double* t_obs, const double* sigma,
                               int32_t per_system, int32_t seed_jac_init, int32_t grad, double* chi2, double* grad_out, int64_t* count, double* tt) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!t_obs || !sigma || !chi2 || !ntt_body) return fail(NBG_ERR_ARG, "ntt_body, t_obs, sigma and chi2 are required");
  if (grad && !grad_out) return fail(NBG_ERR_ARG, "grad_out is required with grad = 1");
  if (!p->kids.empty()) {
    size_t RT = 0;
    for (int i = 0; i < p->n; ++i) RT += (size_t)std::max(0, ntt_body[i]);
    const size_t M = 7 * (size_t)p->n;
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_transit_chi2_fused(k, h, tmax, ti, ntt_body, per_system ? t_obs + lo * RT : t_obs, per_system ? sigma + lo * RT : sigma, per_system,
                                    seed_jac_init, grad, chi2 + lo, at(grad_out, lo * M), at(count, (size_t)lo * p->n), at(tt, lo * RT));
    });
  }
  if (!p->has_state) return fail(NBG_ERR_ARG, "no state set");
  CK(cudaSetDevice(p->device));
  const size_t nsys = p->nsys, n = p->n, M = 7 * n;
  size_t RT = 0;
  for (size_t i = 0; i < n; ++i) RT += (size_t)std::max(0, ntt_body[i]);
  const size_t nobs = (per_system ? nsys : 1) * RT;
  if (p->btobs.ensure(std::max<size_t>(8, nobs * 8)) || p->bsigma.ensure(std::max<size_t>(8, nobs * 8)) || p->bchi2.ensure(nsys * 8) ||
      p->bgq.ensure(nsys * M * 8))
    return fail(NBG_ERR_NOMEM, "chi^2 buffers allocation failed");
  CK(cudaMemcpyAsync(p->btobs.p, t_obs, nobs * 8, cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemcpyAsync(p->bsigma.p, sigma, nobs * 8, cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemsetAsync(p->bchi2.p, 0, nsys * 8, p->stream));
  CK(cudaMemsetAsync(p->bgq.p, 0, nsys * M * 8, p->stream));
  if (grad && seed_jac_init) {
    // jac_step = jac_init instead of the identity: every row the Jacobian kernel forms is then a derivative with respect to the
    // orbital elements, so the accumulated gradient is d chi2 / d elements with no dtdelements pass
    if (!p->jinit_resident) return fail(NBG_ERR_ARG, "seed_jac_init needs the device-computed jac_init (nbg_set_state_elements with want_jac_init)");
    if (int r = ensure_jac(p)) return r;
    CK(cudaStreamWaitEvent(p->stream, p->copy_done, 0));
    jac_from_julia_kernel<<<(unsigned)nsys, 256, 0, p->stream>>>(p->bjinit.as<double>(), p->bJv.as<double>(), (int)n, 0);
    jac_from_julia_kernel<<<(unsigned)nsys, 256, 0, p->stream>>>(nullptr, p->bJe.as<double>(), (int)n, 1);
    p->launches += 2;
    p->jac_valid = true;
  }
  p->fused_per_system = per_system ? 1 : 0;
  p->fused_tt = tt != nullptr;
  if (int r = transit_run(p, h, tmax, ti, ntt_body, 0, grad, nullptr, 2)) return r;
  CK(cudaMemcpyAsync(chi2, p->bchi2.p, nsys * 8, cudaMemcpyDeviceToHost, p->stream));
  if (grad) CK(cudaMemcpyAsync(grad_out, p->bgq.p, nsys * M * 8, cudaMemcpyDeviceToHost, p->stream));
  if (tt) CK(cudaMemcpyAsync(tt, p->btt.p, nsys * RT * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  if (count) return nbg_transit_fetch(p, nullptr, count, nullptr, nullptr);
  return NBG_OK;
}

int32_t nbg_transit_timing(nbg_plan* p, const double* x0, const double* v0, const double* m, const uint8_t* pair, double t0, double h, double tmax,
                           int32_t ti, const int32_t* ntt_body, int32_t mode, int32_t grad, const double* jac_init, double* tt, int64_t* count,
                           double* dtdq0, double* dtdelements, double* x, double* v, double* xerror, double* verror, double* jac_step,
                           double* jac_error, double* dqdt, double* t, uint32_t* status) {
  if (!p) return fail(NBG_ERR_ARG, "plan is NULL");
  if (!p->kids.empty()) {  // every slice runs the whole one-shot sequence on its own thread: uploads, steps and output streaming of the devices overlap
    if (!x0 || !v0 || !m || !ntt_body) return fail(NBG_ERR_ARG, "x, v, m and ntt_body are required");
    const size_t n = p->n, M = 7 * n, C = mode == 1 ? 3 : 1;
    size_t RT = 0;
    for (size_t i = 0; i < n; ++i) RT += (size_t)std::max(0, ntt_body[i]);
    return for_kids(p, [&](nbg_plan* k, long lo, long) {
      return nbg_transit_timing(k, x0 + lo * 3 * n, v0 + lo * 3 * n, m + lo * n, pair, t0, h, tmax, ti, ntt_body, mode, grad, at(jac_init, lo * M * M),
                                at(tt, lo * RT * C), at(count, lo * n), at(dtdq0, lo * RT * M * C), at(dtdelements, lo * RT * M * C), at(x, lo * 3 * n),
                                at(v, lo * 3 * n), at(xerror, lo * 3 * n), at(verror, lo * 3 * n), at(jac_step, lo * M * M), at(jac_error, lo * M * M),
                                at(dqdt, lo * M), at(t, lo), at(status, lo));
    });
  }
  const double w0 = wall_ms();
  if (int r = nbg_set_pair(p, pair)) return r;
  if (int r = nbg_set_state(p, x0, v0, m, t0, nullptr, nullptr, nullptr, nullptr, nullptr)) return r;
  const double w1 = wall_ms();
  // The host destinations are known before the run: every chunk's transit rows (tt, dtdq0, dtdelements) are copied to pinned staging on
  // the copy stream and scattered into the caller's arrays while the following chunks compute, so the device never holds more than
  // one chunk of outputs -- the full-length BASELINE configuration (2 x 83 GB of gradients at 65,536 systems) runs in one call.
  p->sink = nbg_plan::Sink{tt, grad ? dtdq0 : nullptr, grad ? dtdelements : nullptr, false, false};
  const int rr = transit_run(p, h, tmax, ti, ntt_body, mode, grad, jac_init, 1);
  p->sink = nbg_plan::Sink{};
  if (rr) return rr;
  const double w2 = wall_ms();
  if (count) if (int r = nbg_transit_fetch(p, nullptr, count, nullptr, nullptr)) return r;
  const double w3 = wall_ms();
  const int rg = nbg_get_state(p, x, v, xerror, verror, grad ? jac_step : nullptr, grad ? jac_error : nullptr, grad ? dqdt : nullptr, t, status);
  if (p->trace)
    fprintf(stderr, "[nbg trace] transit_timing: set_state %.2f ms, run %.2f ms (device total %.2f), count %.2f ms, get_state %.2f ms\n", w1 - w0,
            w2 - w1, p->timings[4], w3 - w2, wall_ms() - w3);
  return rg;
}

int32_t nbg_counters(nbg_plan* p, int64_t* c8) {
  if (!p || !c8) return fail(NBG_ERR_ARG, "NULL argument");
  if (!p->kids.empty()) {
    for (int q = 0; q < 8; ++q) c8[q] = 0;
    for (nbg_plan* k : p->kids) {
      int64_t c[8];
      nbg_counters(k, c);
      for (int q = 0; q < 8; ++q) c8[q] = q == 6 ? std::max(c8[q], c[q]) : c8[q] + c[q];
    }
    return NBG_OK;
  }
  for (int q = 0; q < 8; ++q) c8[q] = (int64_t)p->counters_host[q];
  c8[4] = p->launches;
  c8[5] = (int64_t)p->counters_host[5];   // the Jacobian kernel applies main-loop steps only: transit outputs come from the adjoint vectors
  return NBG_OK;
}
int32_t nbg_counters_reset(nbg_plan* p) {
  if (!p) return fail(NBG_ERR_ARG, "NULL argument");
  if (!p->kids.empty()) {
    for (nbg_plan* k : p->kids) if (int r = nbg_counters_reset(k)) return r;
    return NBG_OK;
  }
  CK(cudaSetDevice(p->device));
  for (auto& q : p->counters_host) q = 0;
  p->launches = 0;
  p->chunk_retries = 0;
  CK(cudaMemsetAsync(p->bcounters.p, 0, 64, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return NBG_OK;
}
int32_t nbg_last_timings(nbg_plan* p, double* ms8) {
  if (!p || !ms8) return fail(NBG_ERR_ARG, "NULL argument");
  if (!p->kids.empty()) {  // slices run concurrently: the slowest slice is the call
    for (int q = 0; q < 8; ++q) ms8[q] = 0;
    for (nbg_plan* k : p->kids) for (int q = 0; q < 8; ++q) ms8[q] = std::max(ms8[q], k->timings[q]);
    return NBG_OK;
  }
  for (int q = 0; q < 8; ++q) ms8[q] = p->timings[q];
  return NBG_OK;
}
int64_t nbg_cuda_stream(nbg_plan* p) { return !p ? 0 : (int64_t)(intptr_t)(p->kids.empty() ? p->stream : p->kids[0]->stream); }
int64_t nbg_chunk_retries(nbg_plan* p) {
  if (!p) return 0;
  long r = p->chunk_retries;
  for (nbg_plan* k : p->kids) r += k->chunk_retries;
  return r;
}
int64_t nbg_state_generation(nbg_plan* p) {
  if (!p) return 0;
  int64_t g = p->generation;
  for (nbg_plan* k : p->kids) g += k->generation;
  return g;
}
int32_t nbg_plan_devices(nbg_plan* p, int32_t* devices, int32_t cap) {
  if (!p) return 0;
  if (p->kids.empty()) { if (devices && cap > 0) devices[0] = p->device; return 1; }
  for (size_t k = 0; k < p->kids.size() && (int)k < cap; ++k) if (devices) devices[k] = p->kids[k]->device;
  return (int32_t)p->kids.size();
}

int32_t nbg_fp64_peak(int32_t device, double* tflops, double* ms) {
  const int ndev = nbg_device_count();
  if (ndev == 0) return fail(NBG_ERR_NO_DEVICE, "no CUDA device: libnbgrad_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(NBG_ERR_ARG, "device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  struct Res {  // released on every exit path
    double* out = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Res() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (out) cudaFree(out); }
  } R;
  CK(cudaMalloc(&R.out, (size_t)blocks * threads * 8));
  CK(cudaEventCreate(&R.e0));
  CK(cudaEventCreate(&R.e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(R.e0));
    dfma_peak_kernel<<<blocks, threads>>>(R.out, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(R.e1));
    CK(cudaEventSynchronize(R.e1));
    float t = 0;
    CK(cudaEventElapsedTime(&t, R.e0, R.e1));
    if (rep > 0 && t < best) best = t;
  }
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return NBG_OK;
}

}  // extern "C"
