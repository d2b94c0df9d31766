/* nbgrad.h — C ABI of libnbgrad_b200.so: batched AHL21 step + forward-mode Jacobian + transit timing on B200 (sm_100a).
 *
 * The reference (ericagol/NbodyGradient.jl, pure Julia) has no FFI of its own; its only seam is the
 * Function-valued field Integrator.scheme called once per system per step
 * (src/integrator/Integrator.jl:17-22,178-180; src/transits/Transits.jl:156-158).  That granularity cannot
 * cross PCIe, so this ABI sits one level up and replaces the callable-Integrator drivers for a BATCH of
 * independent systems (batch of 1 = the reference call):
 *
 *   nbg_integrate        <-> (intr::Integrator)(s, time; grad) / (s, N; grad) / (s; grad)
 *                            src/integrator/Integrator.jl:159-197, :211-234, :247
 *   nbg_transit_timing   <-> (intr::Integrator)(s, tt::TransitTiming|TransitParameters; grad)
 *                            src/transits/Transits.jl:140-180 + detect_transits!/findtransit!/dtbvdq!/
 *                            calc_dtdelements!  src/transits/timing.jl:3-194
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary; every function returns 0 on success or a negative NBG_ERR_*;
 *    nbg_last_error() gives a thread-local message.  Numerical events (iteration caps, non-finite state,
 *    transit-slot overflow) are reported per system in `status` and results are still written.
 *  - all floating point data is IEEE double; arrays are Julia's column-major arrays with the SYSTEM index
 *    slowest:   x[sys][body][3]  (Julia x[dim,body]),  m[sys][body],
 *               jac_step[sys][col][row]  (Julia jac_step[row,col], M = 7*nbody),  dqdt[sys][M].
 *    Row/column index of body i (0-based) in jac_step/dqdt: 7*i+{0,1,2} = x, +{3,4,5} = v, +6 = m
 *    (src/integrator/ahl21/ahl21.jl:19-21).
 *  - the caller owns every host pointer; the library owns device memory inside the plan and keeps no host
 *    pointer after a call returns.  Calls on one plan must be serialised by the caller; different plans may be
 *    used from different threads.  Calls block until results are in the host buffers.
 *  - s.pair (Integrator.jl:91; all-false by default, nothing in src/ sets it): `pair` is Julia's N x N Bool matrix
 *    (entry [i,j] at i + N*j, one byte each; only i < j is read, as in the reference) or NULL = all-false.  Flagged pairs
 *    get kickfast!/phic! instead of Kepler drifts (ahl21.jl:337-552); any nbody <= 16 (128-bit pair mask).
 *    One pair matrix applies to every system of the batch.
 *  - there is no CPU fallback: without a CUDA device every compute call returns NBG_ERR_NO_DEVICE.
 */
#ifndef NBGRAD_H
#define NBGRAD_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NBG_OK 0
#define NBG_ERR_ARG (-1)
#define NBG_ERR_NO_DEVICE (-2)
#define NBG_ERR_CUDA (-3)
#define NBG_ERR_UNSUPPORTED (-4)
#define NBG_ERR_NOMEM (-5)

#define NBG_MAX_BODIES 16

/* per-system status bits */
#define NBG_ST_NONFINITE 1u       /* x, v not finite at the end */
#define NBG_ST_TRANSIT_ITMAX 2u   /* findtransit! Newton hit its 20-iteration cap (timing.jl:49,70) */
#define NBG_ST_EVENT_OVERFLOW 4u  /* reserved (v1: transits dropped on queue overflow).  Never set since v2: a chunk whose transit queue
                                     overflows is re-run with a queue of the measured size (nbg_chunk_retries counts the re-runs). */
#define NBG_ST_NTT_OVERFLOW 8u    /* a body had more transits than its ntt capacity (counted, not stored: timing.jl:18-19) */

typedef struct nbg_plan nbg_plan;

int32_t nbg_version(void);
const char* nbg_last_error(void);
int32_t nbg_device_count(void); /* number of CUDA devices, 0 if none / no driver */
const char* nbg_source_hash(void); /* hash of the sources the library was compiled from (nbgrad/build.py); "unknown" for a hand build */
int32_t nbg_build_flags(void);  /* bit 0: built with -DNBG_EXPERIMENTS (the rejected kernel variants of DESIGN.md 5 are selectable) */

/* A plan holds device buffers for `nsys` systems of `nbody` bodies on CUDA device `device`.
 * stream_budget_bytes bounds the operator-stream buffer (0 = default: 1/4 of free device memory). */
int32_t nbg_plan_create(nbg_plan** plan, int32_t nbody, int64_t nsys, int32_t device, int64_t stream_budget_bytes);
/* One plan over several devices (systems are independent, no exchange step): the batch is cut into ndev contiguous
 * slices, slice k lives on CUDA device devices[k] with its own child plan, stream and host thread; every call on the returned plan
 * runs on all slices concurrently and reads / writes the slices of the caller's arrays in place (the host-side "gather" is implicit in
 * the disjoint slices).  Results are bit-identical to a single-device plan.  A device may be listed more than once. */
int32_t nbg_plan_create_multi(nbg_plan** plan, int32_t nbody, int64_t nsys, const int32_t* devices, int32_t ndev, int64_t stream_budget_bytes);
int32_t nbg_plan_devices(nbg_plan* plan, int32_t* devices, int32_t cap); /* returns the number of slices, fills devices[0..min(cap, slices)) */
int32_t nbg_plan_destroy(nbg_plan* plan);
/* Bumped by every call that changes the resident state (set_state*, integrate*, transit*): a caller that skips an upload because
 * "the plan already holds my state" must compare this token with the one it saw when it put the state there. */
int64_t nbg_state_generation(nbg_plan* plan);

/* ---- resident state (State{T}, Integrator.jl:49-72) ------------------------------------------------------
 * nbg_set_state uploads x, v, m (required) and optionally xerror, verror, jac_step, jac_error, dqdt
 * (NULL = the State(ic) defaults: zeros, jac_step = I; Integrator.jl:82-103), and sets s.t = t0 for all systems.
 * nbg_get_state downloads whatever is non-NULL. */
/* s.pair for the resident-state calls (the one-shot calls take it as an argument); sticky until changed. */
int32_t nbg_set_pair(nbg_plan* plan, const uint8_t* pair);
int32_t nbg_set_state(nbg_plan* plan, const double* x, const double* v, const double* m, double t0, const double* xerror,
                      const double* verror, const double* jac_step, const double* jac_error, const double* dqdt);
int32_t nbg_get_state(nbg_plan* plan, double* x, double* v, double* xerror, double* verror, double* jac_step, double* jac_error,
                      double* dqdt, double* t, uint32_t* status);

/* State(ic::ElementsIC) computed on the device (init_nbody, src/ics/init_nbody.jl:13-27; kepler_init, src/ics/kepler_init.jl:66-210):
 * elements[sys][c][i] is Julia's elements[i,c] (n x 7 column-major: m, P, t0, ecosw, esinw, I, Omega) with the system index slowest;
 * eps is the n x n hierarchy matrix ic.epsilon (Julia column-major, one for the whole batch) or NULL = fully nested
 * (ElementsIC(t0, N::Int, elements)).  Sets x, v, m and the State(ic) defaults, and -- if want_jac_init -- keeps
 * jac_init = d(x,v,m)/d(elements,m) resident: nbg_transit_timing_resident with jac_init == NULL then uses it for dtdelements.
 * nbg_get_jac_init downloads it ([sys][col][row]). */
int32_t nbg_set_state_elements(nbg_plan* plan, const double* elements, const double* eps, double t0, int32_t want_jac_init);
int32_t nbg_get_jac_init(nbg_plan* plan, double* jac_init);

/* ---- plain integration on the resident state -----------------------------------------------------------------
 * nsteps steps of size h, then (if h_last != 0) one step of size h_last — exactly what (intr)(s,time) does with
 * nsteps = |round((time-t0)/h)| and h_last = time - (t0 + h*nsteps) (Integrator.jl:159-197).
 * time_mode 0: s.t += Kahan sum of h per step ((intr)(s,N), Integrator.jl:229); 1: s.t = t_final. */
int32_t nbg_integrate_resident(nbg_plan* plan, double h, int64_t nsteps, double h_last, int32_t grad, int32_t time_mode, double t_final);

/* (intr)(s, o::CartesianOutput) (src/outputs/Outputs.jl:26-49): nsteps steps of size h from the resident state (s.t = t0 + h i);
 * x, v BEFORE every `stride`-th step (Outputs.jl:40 saves the state before the step) are returned as
 * x_samples[k][sys][body][3], k = 0 .. ceil(nsteps/stride)-1.
 * The reference keeps the whole State per sample, jac_step included (Outputs.jl:40 deepcopy): nbg_integrate_sampled_jac also returns
 * jac_samples[k][sys][7n x 7n, Julia column-major] = jac_step before step k*stride (k = 0: the resident jac_step, i.e. the identity for a
 * fresh State -- the reference does not reset it either); needs grad = 1; the chunks of the
 * device pipeline then end on sample steps.  jac_samples = NULL: identical to nbg_integrate_sampled. */
int32_t nbg_integrate_sampled(nbg_plan* plan, double h, int64_t nsteps, int64_t stride, int32_t grad, double* x_samples, double* v_samples);
int32_t nbg_integrate_sampled_jac(nbg_plan* plan, double h, int64_t nsteps, int64_t stride, int32_t grad, double* x_samples, double* v_samples,
                                  double* jac_samples);

/* get_orbital_elements(s, ic) (src/outputs/elements.jl:108-137) of the resident state of every system, on the device:
 * elements_out[sys][body][11] = (m, P, t0 = 0, ecosw, esinw, I, Omega, a, e, omega, tp), the fields of the reference's Elements (body 0 carries
 * only its mass).  eps = ic.epsilon (Julia column-major n x n) or NULL = fully nested.  Called between integration calls it gives the
 * elements at sampled steps without moving x, v off the device. */
int32_t nbg_orbital_elements(nbg_plan* plan, const double* eps, double* elements_out);

/* One-shot form with HOST buffers: set_state + integrate + get_state. */
int32_t nbg_integrate(nbg_plan* plan, const double* x0, const double* v0, const double* m, const uint8_t* pair, double t0, double h,
                      int64_t nsteps, double h_last, int32_t grad, double* x, double* v, double* xerror, double* verror,
                      double* jac_step, double* jac_error, double* dqdt, uint32_t* status);

/* ---- transit timing ---------------------------------------------------------------------------------------------
 * Runs nsteps = |round(tmax/h)| steps from the resident state with transit detection for every body except `ti`
 * (0-based; the reference default ti=1 is 0 here) and Newton refinement of each transit (timing.jl:3-110).
 *
 * Output layout ("ragged by body"): body i owns ntt_body[i] slots; off[i] = sum_{b<i} ntt_body[b], RT = sum ntt_body.
 *   tt[sys][off[i]+k]             k-th transit time of body i            (Julia tt.tt[i,k])
 *   count[sys][i]                 number of transits detected (may exceed ntt_body[i]; extra ones are not stored)
 *   dtdq0[sys][off[i]+k][7*p+q]   d tt / d (q-th coordinate of body p)   (Julia tt.dtdq0[i,k,q,p])
 *   dtdelements[...] same shape    dtdq0 . jac_init                       (Julia tt.dtdelements[i,k,l,k'])
 * With ntt_body[i] = ntt for all i this is the reference's dense TransitTiming with a permuted index order; unfilled
 * slots are 0 (Transits.jl:46-48).  jac_init is [sys][col][row] (Julia column-major); NULL = use the device-computed one of
 * nbg_set_state_elements if there is one, else skip dtdelements.
 * mode 0 = TransitTiming; mode 1 = TransitParameters: tt/dtdq0/dtdelements get a leading component axis of 3
 * (time, v_sky, b_sky^2): ttbv[sys][off[i]+k][3], dtbvdq0[sys][off[i]+k][7*p+q][3].
 * grad = 0: times only (dtdq0/dtdelements untouched, no Jacobian propagated).
 * The *_resident form keeps DENSE outputs on the device (nbg_transit_fetch copies them out, nbg_transit_chi2 reduces them); it fails
 * with NBG_ERR_NOMEM when nsys * RT * 7N doubles do not fit.  The one-shot form nbg_transit_timing takes host buffers for everything
 * and never holds more than one chunk of outputs on the device: after every chunk of steps the rows of the transits found in it
 * (tt, dtdq0, dtdelements: one row per transit) go to pinned staging on a copy stream and are scattered into the caller's arrays
 * while the next chunks compute.  That is the call for the full-length configurations (65,536 systems x 1600 d: 2 x 83 GB of
 * gradients).  nbg_transit_fetch with only `count` non-NULL works after either form.
 * No transit is ever dropped: the number queued in a chunk is read back after the trajectory kernel, and a queue that was too small
 * is grown and the chunk re-run from its saved start state before anything consumed it. */
int32_t nbg_transit_timing_resident(nbg_plan* plan, double h, double tmax, int32_t ti, const int32_t* ntt_body, int32_t mode,
                                    int32_t grad, const double* jac_init_host_or_null);
int32_t nbg_transit_fetch(nbg_plan* plan, double* tt, int64_t* count, double* dtdq0, double* dtdelements);
/* Fused transit-time likelihood on the results of the last nbg_transit_timing_resident call (mode 0): what an optimiser / HMC
 * caller computes from tt, dtdq0, dtdelements (docs/src/gradients.md), reduced on the device so that 1 + 2M doubles per system are
 * copied out instead of the full arrays:
 *   chi2[sys] = sum ((tt - t_obs) / sigma)^2,   grad_q0[sys][7p+q] = sum 2 (tt - t_obs) / sigma^2 * dtdq0[..][7p+q],   grad_elements likewise
 * over the stored transits.  t_obs, sigma use tt's slot layout ([off[i]+k]); per_system = 0: one table [RT] for the whole batch,
 * 1: [sys][RT].  Slots with sigma <= 0 or a non-finite t_obs are skipped.  grad_q0 / grad_elements may be NULL. */
int32_t nbg_transit_chi2(nbg_plan* plan, const double* t_obs, const double* sigma, int32_t per_system, double* chi2, double* grad_q0,
                         double* grad_elements);
/* The same run with the likelihood FUSED into the Jacobian kernel: chi2[sys] and grad[sys][7p+q] are accumulated where
 * d tt / d q0 is produced, so no dtdq0 / dtdelements row is ever stored and 1 + M doubles per system leave the device.  Runs from the
 * resident state (nbg_set_state / nbg_set_state_elements).  seed_jac_init = 1 (needs nbg_set_state_elements(..., want_jac_init = 1))
 * starts jac_step from jac_init instead of the identity, which makes every derivative one with respect to the orbital elements:
 * grad is then d chi2 / d elements (Julia dtdelements index order) and the resident jac_step afterwards is d state / d elements;
 * seed_jac_init = 0 gives d chi2 / d q0.  grad = 0: chi2 only (no Jacobian is propagated).  count [sys][N] and tt [sys][RT] are
 * optional outputs (NULL = not wanted); t_obs / sigma / per_system as in nbg_transit_chi2. */
int32_t nbg_transit_chi2_fused(nbg_plan* plan, double h, double tmax, int32_t ti, const int32_t* ntt_body, const double* t_obs,
                               const double* sigma, int32_t per_system, int32_t seed_jac_init, int32_t grad, double* chi2, double* grad_out,
                               int64_t* count, double* tt);
int32_t nbg_transit_timing(nbg_plan* plan, const double* x0, const double* v0, const double* m, const uint8_t* pair, double t0,
                           double h, double tmax, int32_t ti, const int32_t* ntt_body, int32_t mode, int32_t grad,
                           const double* jac_init, double* tt, int64_t* count, double* dtdq0, double* dtdelements, double* x,
                           double* v, double* xerror, double* verror, double* jac_step, double* jac_error, double* dqdt,
                           double* t, uint32_t* status);

/* ---- instrumentation ---------------------------------------------------------------------------------------------
 * Counters accumulated since plan creation / nbg_counters_reset:
 *  c[0] main-loop system-steps, c[1] findtransit Newton iterations in the reference's form (full gradient steps; the gradient-free
 *  pre-iterations are not counted), c[2] extra final steps at the converged time (only taken when the last iteration was not already
 *  at it),
 *  c[3] transits stored, c[4] kernel launches, c[5] Jacobian system-steps applied (main-loop steps; the transit sub-steps are not applied to
 *  the matrix since v2: their outputs come from adjoint vectors, csrc/nbg_adjoint.cuh),
 *  c[6] steps per chunk of the last call (operator-stream budget), c[7] Jacobian-kernel launches.
 * For a multi-device plan the counters are summed over the slices and the timings are the maximum over the slices.
 * nbg_last_timings: device milliseconds (CUDA events on the plan's stream) spent in the last compute call in the
 *  trajectory kernel [0], transit-refinement kernel [1], Jacobian kernel [2], everything else [3]; [4] = total;
 *  [5] = dense phisalpha-operator kernel; [6] = pair-operator kernel (split path); [7] = transit adjoint kernel. */
int32_t nbg_counters(nbg_plan* plan, int64_t* c8);
int32_t nbg_counters_reset(nbg_plan* plan);
int32_t nbg_last_timings(nbg_plan* plan, double* ms8);
int64_t nbg_cuda_stream(nbg_plan* plan); /* cudaStream_t of the plan (of its first slice), for callers that time with their own events */
int64_t nbg_chunk_retries(nbg_plan* plan); /* chunks re-run because their transit queue was too small, since plan creation / nbg_counters_reset */
/* FP64 (DFMA) pipe peak of `device`, measured with 8 independent FMA chains per thread: the roofline denominator
 * for this path (the driver-written MEASURED_PEAKS.json carries HBM and bf16 figures only). */
int32_t nbg_fp64_peak(int32_t device, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif
