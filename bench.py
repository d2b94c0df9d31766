#!/usr/bin/env python
"""bench.py — system-steps/s with grad on the TRAPPIST-1 batch (BASELINE.json configs[1]).

Workload (config.workload): TRAPPIST-1, 8 bodies, AHL21 h = 0.06 d, grad = true, TransitTiming, a batch of
`--nsys` (default 65,536) perturbed systems per GPU (counter-based Philox RNG, seed 20211582).
One bench "step" = one pass of the hot path over the batch for a WINDOW of `--window` AHL21 steps (default 64:
3.84 d of the 1600 d integration, continuing from the resident state of the previous step), including transit
detection, findtransit! Newton refinement and the transit-time gradients.  The metric is a rate, so the window
only bounds the run time; `--window 26667` is the full 1600 d configuration.

  value : whole-job main-loop system-steps/s with inputs resident in HBM (nbg_transit_timing_resident)
  e2e   : same metric through the one-shot C-ABI call with HOST (pinned) buffers: H2D of x, v, m, jac_init and D2H
          of tt, count, dtdq0, dtdelements, x, v inside the timed region (rows streamed chunk by chunk)
  --impl reference : the reference's CPU path (the oracle restatement; Julia cannot run here) on all host cores.
  --full           : the FULL-LENGTH configuration in ONE library call: 26,667 steps (1600 d), every transit's tt / dtdq0 /
                     dtdelements delivered to host arrays (--full-output arrays; needs ~2.6 MB of host RAM per system) or reduced
                     to chi^2 + gradient on the device (--full-output chi2); 3 sampled systems are checked against the CPU oracle.
  --single-process : with --gpus N (no torchrun): ONE process, one multi-device plan (nbg_plan_create_multi), one host thread per GPU
                     inside the library.
Per-GPU batch: 65,536 (BASELINE cfg 2) on one GPU, 131,072 with N > 1 (cfg 5: 1,048,576 systems on 8 GPUs); --nsys overrides.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "nbodygradient.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, T0, SEED = 0.06, 7257.0, 20211582
NBODY = 8


def f_grad(n):
    """Canonical flops per grad step (SURVEY.md 8(d) / BASELINE.md)."""
    return 3191 * n * n * (n - 1) + 168 * n ** 3 + 798 * n * n + 2400 * n * (n - 1)


def f_jac(n):
    """Exactly countable Jacobian-propagation part of F_grad (pair GEMMs + Kahan, phisalpha update, drifts, folds)."""
    return 2744 * n * n * (n - 1) + 392 * n * n * (n - 1) + 168 * n ** 3 + 210 * n * n + 588 * n * n


def make_elements(nsys, rank=0):
    """cfg 2 ensemble: system b gets elements*(1+1e-4 xi) on m, P; +1e-4 xi on ecosw, esinw, t0; system 0 unperturbed."""
    import nbgrad as nb
    el = nb.trappist1_elements()
    rng = np.random.Generator(np.random.Philox(key=SEED + rank))
    elb = np.broadcast_to(el, (nsys, NBODY, 7)).copy()
    xi = rng.standard_normal((nsys, NBODY - 1, 5))
    if rank == 0:
        xi[0] = 0.0
    elb[:, 1:, 0] *= 1 + 1e-4 * xi[..., 0]
    elb[:, 1:, 1] *= 1 + 1e-4 * xi[..., 1]
    elb[:, 1:, 2] += 1e-4 * xi[..., 2]
    elb[:, 1:, 3] += 1e-4 * xi[..., 3]
    elb[:, 1:, 4] += 1e-4 * xi[..., 4]
    return elb


def make_batch(nsys, rank=0):
    """The ensemble with its host-side initial conditions (x, v, jac_init from the numpy IC layer)."""
    import nbgrad as nb
    elb = make_elements(nsys, rank)
    x, v, jac_init = nb.init_nbody_elements(elb, T0)
    return elb, x, v, jac_init


def ntt_window(window, elb, pad=2):
    pmin = elb[:, 1:, 1].min(axis=0)
    ntt = np.zeros(NBODY, dtype=np.int32)
    ntt[1:] = np.ceil(window * H / pmin).astype(np.int32) + pad
    return ntt


def load_profile(name):
    """A committed ncu-derived profile (profiles/<name>) is only used if it was taken from the sources this library was built from."""
    from nbgrad.build import source_hash
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None, "no profiles/%s" % name
    pj = json.load(open(path))
    if pj.get("source_hash") != source_hash():
        return None, "profiles/%s is stale: taken at source hash %s, library sources are %s" % (name, pj.get("source_hash"), source_hash())
    return pj, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


CPU_BUILD = "oracle -O3 build"


def cpu_run(nsys, window, nthreads, x, v, m, jac_init):
    """Reference CPU path (oracle, -O3 build) on `nsys` systems for `window` steps, one system per thread."""
    from oracle.binding import Oracle, build
    try:
        build(fast_native=True)  # rebuild the timing build for this host's ISA
    except Exception:
        pass
    o = Oracle(fast=True, blas=not os.environ.get("NBGRAD_REF_NO_BLAS"))
    global CPU_BUILD
    CPU_BUILD = "oracle -O3 build, " + ("dense products (the reference's mul! calls) through %s" % o.blas if o.blas else "built-in loop nests for the dense products (%s)" % ("NBGRAD_REF_NO_BLAS set" if os.environ.get("NBGRAD_REF_NO_BLAS") else "no OpenBLAS found"))
    tmax = window * H
    ntt = int(np.ceil(tmax / 1.5) + 3)
    jcm = np.ascontiguousarray(jac_init[:nsys].transpose(0, 2, 1))
    t = time.perf_counter()
    r = o.batch_transit_timing(x[:nsys], v[:nsys], m[:nsys], T0, H, tmax, ntt, grad=True, jac_init_cm=jcm, nthreads=nthreads)
    dt = time.perf_counter() - t
    return nsys * window / dt, dt, r


def workload_name(ngpus):
    """the same string in both arms (own and --impl reference)"""
    return "TRAPPIST-1 N=8 h=0.06 grad=true TransitTiming (BASELINE cfg %s)" % ("2" if ngpus == 1 else "5 shard: 131,072 systems per GPU = 1,048,576 on 8")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsys = cores * args.ref_systems_per_core
    elb, x, v, jac_init = make_batch(max(nsys, 8))
    m = np.ascontiguousarray(elb[:, :, 0])
    window = args.ref_window
    for _ in range(args.warmup):
        cpu_run(min(nsys, cores), max(4, window // 8), cores, x, v, m, jac_init)
    ts = []
    for _ in range(args.steps):
        rate, dt, _ = cpu_run(nsys, window, cores, x, v, m, jac_init)
        ts.append(dt)
    T = float(np.sum(ts))
    value = nsys * window * args.steps / T
    out = {"impl": "reference", "metric": "system-steps/s w/ grad (TRAPPIST-1 batch)", "value": value, "unit": "system-steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.gpus), "batch_per_gpu": args.nsys if args.nsys > 0 else (65536 if args.gpus == 1 else 131072),
                      "sample_batch": nsys, "window_steps": window,
                      "parallelism": "CPU: one system per host thread, %d threads" % cores},
           "cpu_baseline": {"value": value, "unit": "system-steps/s", "cores": cores, "kind": "port",
                            "sample": "%d systems x %d steps per step, %s, one system per thread" % (nsys, window, CPU_BUILD)},
           "e2e": {"value": value, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def roofline_block(L, local, NB, cnt, ksum, ndev=1):
    """Roofline of the dominant kernel + the whole path, from the counters / device timings of the timed region.  ndev > 1 (one process
    driving several GPUs): the counters are sums over the devices while the timings are the slowest device's, so work is per device."""
    cnt = np.asarray(cnt, dtype=np.float64).copy()
    for q in (0, 1, 2, 3, 5, 7):
        cnt[q] /= ndev
    tfl, pms = C.c_double(0), C.c_double(0)
    L.nbg_fp64_peak(C.c_int32(local), C.byref(tfl), C.byref(pms))
    peak = tfl.value
    names = ["traj_kernel", "transit_kernel", "jac_rx_kernel"]
    dom = 2 if ksum[2] >= max(ksum[0] + ksum[6], ksum[1]) else int(np.argmax(ksum[:2]))
    jac_steps = float(cnt[5])                       # Jacobian system-steps applied to the matrix (main-loop steps; transit outputs use adjoint vectors)
    step_equiv = float(cnt[0] + cnt[1] + cnt[2])    # trajectory step-equivalents: main steps + reference-form Newton iterations + final steps
    f_scalar = f_grad(NB) - f_jac(NB)
    flops = {0: f_scalar * float(cnt[0]), 1: f_scalar * float(cnt[1] + cnt[2]), 2: f_jac(NB) * jac_steps}
    dom_ms = ksum[dom] + (ksum[6] if dom == 0 else 0.0)
    achieved = flops[dom] / (dom_ms * 1e-3) / 1e12
    # whole path: F_jac for every step that propagates the Jacobian, F_scalar for every trajectory step-equivalent (the Jacobian-free
    # Newton iterations are NOT credited with Jacobian flops)
    path_achieved = (f_jac(NB) * jac_steps + f_scalar * step_equiv) / (ksum[4] * 1e-3) / 1e12
    # the reference applies one more full Jacobian step per transit (findtransit!, timing.jl:75-80); here those outputs come from adjoint
    # vectors, so that work is not executed -- credited only in the "reference-equivalent" figure
    path_ref_equiv = (f_jac(NB) * (jac_steps + float(cnt[3])) + f_scalar * step_equiv) / (ksum[4] * 1e-3) / 1e12
    prof, why = load_profile("r02_jac_rx_profile.json")
    traffic, pipe = None, None
    if prof and dom == 2:
        traffic = prof["dram_bytes_per_jacobian_step"] * jac_steps / max(1, int(cnt[7]))
        pipe = {"fp64_pipe_cycles_active_pct": prof.get("fp64_pipe_cycles_active_pct"), "smem_lsu_wavefronts_pct": prof.get("smem_lsu_wavefronts_pct"),
                "issue_active_pct": prof.get("issue_active_pct"), "source": prof.get("source")}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    # algorithmic HBM bytes per main system-step: scalars, compact phisalpha records, Kepler records, dense phisalpha operator, each
    # written once and read once
    P = NB * (NB - 1) // 2
    stream_bytes = 2 * (2 * P * 12 + P * 24 + 2 * P * 64 + 12 * NB * NB) * 8.0 * (float(cnt[0]) + float(cnt[2]))
    hbm_gbs = stream_bytes / (ksum[4] * 1e-3) / 1e9
    return {"bound": "fp64", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "traffic": traffic, "traffic_note": why or "dram read+write bytes per launch: ncu-measured bytes per Jacobian step (profiles/r02_jac_rx_profile.json, same sources) x steps per launch",
            "executed": pipe, "executed_note": "frac is CANONICAL flops (SURVEY 8d) over the measured DFMA peak; the kernel executes ~0.29 FP64 thread-instructions per canonical flop "
                                               "(6x6 block + rank-one structure of jac_ij), so the hardware figure is the ncu FP64-pipe activity in `executed`",
            "flops": "canonical (SURVEY 8d): F_jac(%d)=%d per Jacobian step, F_scalar=%d per trajectory step" % (NB, f_jac(NB), f_scalar),
            "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry; nominal 37.2)",
            "path": {"achieved": path_achieved, "frac": path_achieved / peak if peak else None,
                     "note": "whole path: (F_jac x Jacobian steps + F_scalar x trajectory step-equivalents) / total device time",
                     "reference_equivalent_tflops": path_ref_equiv,
                     "reference_equivalent_note": "as above plus F_jac for the per-transit Jacobian step the reference takes and this path replaces by adjoint vectors"},
            "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks.get("hbm_gbs"), "note": "operator / scalar streams written once and read once, algorithmic bytes over total device time"}}


def oracle_check(x, v, m, jac_init, systems, tmax, tt_gpu, off, ntt_body, nthreads):
    """tt of `systems` (rows of the batch) from the CPU oracle over tmax, compared with the GPU's tt[sys][off[i]+k] rows.  Checker only."""
    from oracle.binding import Oracle
    o = Oracle(fast=False)
    sel = np.asarray(systems)
    ntt = int(ntt_body.max())
    r = o.batch_transit_timing(x[sel], v[sel], m[sel], T0, H, tmax, ntt, grad=False, nthreads=nthreads, want_grad_arrays=False)
    worst, ntr = 0.0, 0
    for q, b in enumerate(sel):
        for i in range(1, NBODY):
            nk = int(min(r["count"][q, i], ntt_body[i]))
            ref = r["tt"][q, :nk, i]
            got = tt_gpu[b, off[i]:off[i] + nk]
            if nk:
                worst = max(worst, float(np.max(np.abs(got - ref) / np.abs(ref))))
                ntr += nk
    return worst, ntr


def run_full(args, L, nb, _lib, torch, local):
    """The full-length configuration in ONE library call."""
    ptr = _lib.ptr
    nsys, window = args.nsys, int(round(1600.0 / H))
    elb, x, v, jac_init = make_batch(nsys, 0)
    m = np.ascontiguousarray(elb[:, :, 0])
    ntt = ntt_window(window, elb, pad=3)
    off = np.concatenate([[0], np.cumsum(ntt)[:-1]])
    RT, M = int(ntt.sum()), 7 * NBODY
    devices = np.arange(args.gpus, dtype=np.int32) if args.single_process else np.array([local], dtype=np.int32)
    plan = C.c_void_p()
    _lib.check(L.nbg_plan_create_multi(C.byref(plan), C.c_int32(NBODY), C.c_int64(nsys), ptr(devices), C.c_int32(len(devices)), C.c_int64(int(args.stream_budget_gb * 1e9))))
    tt = np.zeros((nsys, RT))
    cnt = np.zeros((nsys, NBODY), dtype=np.int64)
    status = np.zeros(nsys, dtype=np.uint32)
    xo, vo = np.zeros((nsys, NBODY, 3)), np.zeros((nsys, NBODY, 3))
    sampler = ClockSampler(local)
    el_t = np.ascontiguousarray(elb.transpose(0, 2, 1))
    # warm-up: three short windows through the same entry point (allocations, first-launch costs)
    ntw = ntt_window(64, elb)
    for _ in range(max(3, args.warmup)):
        if args.full_output == "arrays":
            RTw = int(ntw.sum())
            ttw, dw, ew = np.zeros((nsys, RTw)), np.zeros((nsys, RTw, M)), np.zeros((nsys, RTw, M))
            _lib.check(L.nbg_transit_timing(plan, ptr(x), ptr(v), ptr(m), None, C.c_double(T0), C.c_double(H), C.c_double(64 * H), C.c_int32(0), ptr(ntw),
                                            C.c_int32(0), C.c_int32(1), ptr(np.ascontiguousarray(jac_init.transpose(0, 2, 1))), ptr(ttw), ptr(cnt), ptr(dw), ptr(ew),
                                            ptr(xo), ptr(vo), None, None, None, None, None, None, ptr(status)))
            del ttw, dw, ew
        else:
            _lib.check(L.nbg_set_state_elements(plan, ptr(el_t), None, C.c_double(T0), C.c_int32(1)))
            tob, sig = np.full(int(ntw.sum()), T0 + 1.0), np.full(int(ntw.sum()), 1e-3)
            chi, g = np.zeros(nsys), np.zeros((nsys, M))
            _lib.check(L.nbg_transit_chi2_fused(plan, C.c_double(H), C.c_double(64 * H), C.c_int32(0), ptr(ntw), ptr(tob), ptr(sig), C.c_int32(0), C.c_int32(1),
                                                C.c_int32(1), ptr(chi), ptr(g), ptr(cnt), None))
    _lib.check(L.nbg_counters_reset(plan))
    torch.cuda.synchronize()
    sampler.start()
    t_start = time.perf_counter()
    if args.full_output == "arrays":
        # pageable destinations (2 x nsys x RT x 56 doubles): the library stages every chunk's rows through pinned memory and scatters
        d, e = np.zeros((nsys, RT, M)), np.zeros((nsys, RT, M))
        jcm = np.ascontiguousarray(jac_init.transpose(0, 2, 1))
        h2d = x.nbytes + v.nbytes + m.nbytes + jcm.nbytes
        _lib.check(L.nbg_transit_timing(plan, ptr(x), ptr(v), ptr(m), None, C.c_double(T0), C.c_double(H), C.c_double(1600.0), C.c_int32(0), ptr(ntt),
                                        C.c_int32(0), C.c_int32(1), ptr(jcm), ptr(tt), ptr(cnt), ptr(d), ptr(e), ptr(xo), ptr(vo), None, None, None, None,
                                        None, None, ptr(status)))
        d2h = None   # counted from the transits actually delivered, below
    else:
        # elements in, chi^2 + d chi2 / d elements out: the optimiser's call; tt comes back too (for the oracle check)
        tob, sig = np.full(RT, T0 + 800.0), np.full(RT, 1e-3)
        chi, g = np.zeros(nsys), np.zeros((nsys, M))
        h2d = el_t.nbytes + tob.nbytes + sig.nbytes
        _lib.check(L.nbg_set_state_elements(plan, ptr(el_t), None, C.c_double(T0), C.c_int32(1)))
        _lib.check(L.nbg_transit_chi2_fused(plan, C.c_double(H), C.c_double(1600.0), C.c_int32(0), ptr(ntt), ptr(tob), ptr(sig), C.c_int32(0), C.c_int32(1),
                                            C.c_int32(1), ptr(chi), ptr(g), ptr(cnt), ptr(tt)))
        _lib.check(L.nbg_get_state(plan, ptr(xo), ptr(vo), None, None, None, None, None, None, ptr(status)))
        d2h = chi.nbytes + g.nbytes + cnt.nbytes + tt.nbytes + xo.nbytes + vo.nbytes
    wall = time.perf_counter() - t_start
    clocks = sampler.stop()
    kt = np.zeros(8); c8 = np.zeros(8, dtype=np.int64)
    L.nbg_last_timings(plan, ptr(kt)); L.nbg_counters(plan, ptr(c8))
    retries = int(L.nbg_chunk_retries(plan))
    stored = int(np.minimum(cnt, ntt[None, :]).sum())
    if d2h is None:
        d2h = stored * (1 + 2 * M) * 8 + 12 * stored + cnt.nbytes + xo.nbytes + vo.nbytes
        filled = int(np.count_nonzero(tt)), int(np.count_nonzero(np.any(d != 0, axis=2))), int(np.count_nonzero(np.any(e != 0, axis=2)))
    else:
        filled = (int(np.count_nonzero(tt)),)
    # 3 sampled PERTURBED systems against the CPU oracle at full length (all their ~2,770 transit times each)
    sel = [1, nsys // 2 + 1, nsys - 1] if nsys > 3 else list(range(nsys))
    worst, ntr = oracle_check(x, v, m, jac_init, sel, 1600.0, tt, off, ntt, nthreads=min(len(sel), os.cpu_count() or 1))
    main_steps = nsys * window
    out = {"metric": "system-steps/s w/ grad (TRAPPIST-1 batch)", "value": main_steps / (kt[4] * 1e-3), "unit": "system-steps/s", "n_gpus": len(devices),
           "steps": 1, "warmup": max(3, args.warmup), "ms_per_step": float(kt[4]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "full_length": True,
           "config": {"workload": "TRAPPIST-1 N=8 h=0.06 grad=true TransitTiming, FULL LENGTH 1600 d = %d steps in one library call (BASELINE cfg 2)" % window,
                      "batch_per_gpu": nsys // len(devices), "window_steps": window, "chunk_steps": int(c8[6]), "jac_launches": int(c8[7]),
                      "output": args.full_output, "ntt_body": ntt.tolist(),
                      "l2": "inputs larger than L2 (operator streams of a chunk are GBs; L2 = 126 MB)", "parallelism": "single process, %d device slice(s)" % len(devices)},
           "clocks": clocks, "gpu_launches": int(c8[4]),
           "e2e": {"value": main_steps / wall, "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "timing": "host wall clock around the one blocking library call (%.1f s)" % wall},
           "rates": {"transits_stored": stored, "transits_per_system": stored / nsys, "newton_iters_per_transit": float(c8[1]) / max(1, int(c8[3])),
                     "jacobian_steps_per_s": float(c8[5]) / (kt[4] * 1e-3)},
           "kernel_ms": {"traj_kernel": float(kt[0]), "transit_kernel": float(kt[1]), "jac_rx_kernel": float(kt[2]), "phi_dense_kernel": float(kt[5]),
                         "pair_op_kernel": float(kt[6]), "transit_adjoint_kernel": float(kt[7]), "other": float(kt[3]), "total": float(kt[4])},
           "status_bits": {"nonfinite": int((status & 1 != 0).sum()), "transit_itmax": int((status & 2 != 0).sum()),
                           "event_overflow": int((status & 4 != 0).sum()), "ntt_overflow": int((status & 8 != 0).sum())},
           "chunk_reruns": retries, "outputs_filled": filled,
           "full_check": {"systems": sel, "transits_compared": ntr, "max_rel_dev_tt_vs_cpu_oracle": worst, "tolerance": 1e-11, "ok": bool(worst < 1e-11)}}
    if True:
        out["roofline"] = roofline_block(L, local, NBODY, c8, kt, len(devices))
    L.nbg_plan_destroy(plan)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nsys", type=int, default=0, help="systems per GPU (0 = 65,536 on one GPU, 131,072 per GPU on several: BASELINE cfg 2 / cfg 5)")
    ap.add_argument("--window", type=int, default=64)
    ap.add_argument("--ref-window", type=int, default=4096, help="steps per system of the CPU-baseline sample")
    ap.add_argument("--ref-systems-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-input", choices=["elements", "cartesian"], default="cartesian",
                    help="what crosses PCIe on the way in: orbital elements (init_nbody on the device) or x, v, m + host-computed jac_init")
    ap.add_argument("--e2e-output", choices=["arrays", "chi2"], default="arrays",
                    help="with --e2e-input elements: copy out tt/dtdq0/dtdelements, or only the fused chi^2 and its gradient")
    ap.add_argument("--stream-budget-gb", type=float, default=0.0, help="HBM for the operator streams of a chunk (0 = the library default, 1/4 of free memory)")
    ap.add_argument("--full", action="store_true", help="full-length run (1600 d) in one library call; see the module docstring")
    ap.add_argument("--full-output", choices=["arrays", "chi2"], default="chi2")
    ap.add_argument("--single-process", action="store_true", help="--gpus N in ONE process through a multi-device plan (not under torchrun)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import nbgrad as nb
    from nbgrad import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    single = args.single_process and world == 1 and args.gpus > 1
    if args.nsys <= 0:
        args.nsys = 65536 if (world == 1 and not single) else 131072
    dist = None
    # rank 0 prints exactly ONE JSON line on stdout.  NCCL writes its version banner (and anything NCCL_DEBUG asks for) to the C-level
    # stdout, so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a saved copy of the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    L = _lib.lib()
    if args.full:
        if single:
            args.nsys *= args.gpus
        sys.stdout = os.fdopen(real_stdout, "w")
        return run_full(args, L, nb, _lib, torch, local)

    ngpu = args.gpus if single else 1           # GPUs driven by THIS process
    nsys, window = args.nsys * ngpu, args.window
    if single:
        # one process feeding N GPUs: the initial conditions come from the device IC layer (elements in); the numpy IC layer would
        # spend minutes on a million systems
        elb = make_elements(nsys, rank)
        x = v = jac_init = None
        args.no_cpu_baseline = True
        single_ic_from_device = args.e2e_input == "cartesian" and not args.no_e2e   # x, v, jac_init for the one-shot call: fetched below
    else:
        elb, x, v, jac_init = make_batch(nsys, rank)
    m = np.ascontiguousarray(elb[:, :, 0])
    ntt = ntt_window(window, elb)
    off = np.concatenate([[0], np.cumsum(ntt)[:-1]])
    RT, M = int(ntt.sum()), 7 * NBODY
    devices = np.arange(ngpu, dtype=np.int32) if single else np.array([local], dtype=np.int32)

    def new_plan():
        pl = C.c_void_p()
        _lib.check(L.nbg_plan_create_multi(C.byref(pl), C.c_int32(NBODY), C.c_int64(nsys), _lib.ptr(devices), C.c_int32(len(devices)),
                                           C.c_int64(int(args.stream_budget_gb * 1e9))))
        return pl

    plan = new_plan()
    stream = torch.cuda.ExternalStream(L.nbg_cuda_stream(plan), device=torch.device("cuda", local))
    ptr = _lib.ptr

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    # ---- device-resident arm ----
    if single:
        el_dev = np.ascontiguousarray(elb.transpose(0, 2, 1))
        _lib.check(L.nbg_set_state_elements(plan, ptr(el_dev), None, C.c_double(T0), C.c_int32(1 if single_ic_from_device else 0)))
        if single_ic_from_device:   # the host copies of x, v, jac_init that the end-to-end arm uploads again every step (untimed here)
            x, v, jcm = np.empty((nsys, NBODY, 3)), np.empty((nsys, NBODY, 3)), np.empty((nsys, M, M))
            _lib.check(L.nbg_get_state(plan, ptr(x), ptr(v), None, None, None, None, None, None, None))
            _lib.check(L.nbg_get_jac_init(plan, ptr(jcm)))
            jac_init = jcm.transpose(0, 2, 1)   # [row, col] view; pinned below in the ABI's [col][row] order again
    else:
        _lib.check(L.nbg_set_state(plan, ptr(x), ptr(v), ptr(m), C.c_double(T0), None, None, None, None, None))
    tmaxw = window * H

    def step_resident():
        _lib.check(L.nbg_transit_timing_resident(plan, C.c_double(H), C.c_double(tmaxw), C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), None))

    def barrier():
        if dist is not None:
            dist.barrier()
        for d in devices:
            torch.cuda.synchronize(int(d))

    for _ in range(args.warmup):
        step_resident()
    _lib.check(L.nbg_counters_reset(plan))
    kt = np.zeros(8)
    ksum = np.zeros(8)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        L.nbg_last_timings(plan, ptr(kt))
        ksum += kt
    e1.record(stream)
    barrier()
    wall_res = time.perf_counter() - w0
    clocks = sampler.stop()
    # one GPU per process: CUDA events on the plan's stream.  Single process over several GPUs: the calls block until every slice is
    # done, so the host clock between the two barriers is the time of the slowest device.
    ms = wall_res * 1e3 if single else e0.elapsed_time(e1)
    tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    cnt = np.zeros(8, dtype=np.int64)
    L.nbg_counters(plan, ptr(cnt))
    status = np.zeros(nsys, dtype=np.uint32)
    _lib.check(L.nbg_get_state(plan, None, None, None, None, None, None, None, None, ptr(status)))
    retries = int(L.nbg_chunk_retries(plan))
    main_steps = nsys * window * args.steps
    value = world * main_steps / (ms_max * 1e-3)

    # ---- end-to-end arm: the one-shot C-ABI call with HOST (pinned) buffers ----
    # Every byte of input and output crosses PCIe inside the timed region, every step; the library streams each chunk's transit rows
    # to pinned staging and scatters them into the caller's arrays while the next chunk computes.
    e2e = None
    e2e_tt = None
    if not args.no_e2e:
        _lib.check(L.nbg_plan_destroy(plan))
        plan = new_plan()
        zero = np.zeros(1)
        B = dict(x=pin(x if x is not None else zero)[1], v=pin(v if v is not None else zero)[1], m=pin(m)[1],
                 j=pin(jac_init.transpose(0, 2, 1) if jac_init is not None else zero)[1], el=pin(elb.transpose(0, 2, 1))[1],
                 tobs=pin(np.full(RT, T0 + 1.0))[1], sig=pin(np.full(RT, 1e-3))[1], chi2=pin(np.zeros(nsys))[1], g=pin(np.zeros((nsys, M)))[1],
                 tt=pin(np.zeros((nsys, RT)))[1], c=pin(np.zeros((nsys, NBODY), dtype=np.int64))[1], d=pin(np.zeros((nsys, RT, M)))[1],
                 e=pin(np.zeros((nsys, RT, M)))[1], xo=pin(np.zeros((nsys, NBODY, 3)))[1], vo=pin(np.zeros((nsys, NBODY, 3)))[1])
        fused = args.e2e_output == "chi2" and args.e2e_input == "elements"
        h2d = sum(B[k].nbytes for k in (("x", "v", "m", "j") if args.e2e_input == "cartesian" else (("el", "tobs", "sig") if fused else ("el",))))
        d2h_fixed = sum(B[k].nbytes for k in (("chi2", "g", "c") if fused else ("c", "xo", "vo")))

        def step_e2e():
            if args.e2e_input == "cartesian":   # x, v, m and the host-computed jac_init go up (1.6 GB of jac_init at 65,536 systems)
                _lib.check(L.nbg_transit_timing(plan, ptr(B["x"]), ptr(B["v"]), ptr(B["m"]), None, C.c_double(T0), C.c_double(H), C.c_double(tmaxw),
                                                C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), ptr(B["j"]), ptr(B["tt"]), ptr(B["c"]), ptr(B["d"]),
                                                ptr(B["e"]), ptr(B["xo"]), ptr(B["vo"]), None, None, None, None, None, None, None))
            elif fused:   # the optimiser's call: elements up, chi^2 + d chi2 / d elements down; no gradient array exists anywhere
                _lib.check(L.nbg_set_state_elements(plan, ptr(B["el"]), None, C.c_double(T0), C.c_int32(1)))
                _lib.check(L.nbg_transit_chi2_fused(plan, C.c_double(H), C.c_double(tmaxw), C.c_int32(0), ptr(ntt), ptr(B["tobs"]), ptr(B["sig"]), C.c_int32(0),
                                                    C.c_int32(1), C.c_int32(1), ptr(B["chi2"]), ptr(B["g"]), ptr(B["c"]), None))
            else:   # the reference's user-level sequence ElementsIC -> State -> intr(s, tt): orbital elements go up, init_nbody runs on the device
                _lib.check(L.nbg_set_state_elements(plan, ptr(B["el"]), None, C.c_double(T0), C.c_int32(1)))
                _lib.check(L.nbg_transit_timing_resident(plan, C.c_double(H), C.c_double(tmaxw), C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), None))
                _lib.check(L.nbg_transit_fetch(plan, ptr(B["tt"]), ptr(B["c"]), ptr(B["d"]), ptr(B["e"])))
                _lib.check(L.nbg_get_state(plan, ptr(B["xo"]), ptr(B["vo"]), None, None, None, None, None, None, None))

        step_e2e()   # warm-up (allocations inside the plan)
        step_e2e()
        barrier()
        t_e0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        wall = time.perf_counter() - t_e0
        t2 = torch.tensor([wall * 1e3], device="cuda", dtype=torch.float64)   # blocking host calls: wall clock is the end-to-end time
        if dist is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        stored = int(np.minimum(B["c"], ntt[None, :]).sum())
        if fused:
            d2h = d2h_fixed
        elif args.e2e_input == "cartesian":   # streamed rows: one row of (tt, dtdq0, dtdelements) + 12 bytes of (system, body, k) per transit
            d2h = d2h_fixed + stored * ((1 + 2 * M) * 8 + 12)
        else:
            d2h = d2h_fixed + B["tt"].nbytes + B["d"].nbytes + B["e"].nbytes
        e2e = {"value": world * main_steps / (float(t2.item()) * 1e-3), "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "input": args.e2e_input, "output": "chi2 (fused in the Jacobian kernel)" if fused else "arrays",
               "transits_per_step": stored, "chi2_sum": float(B["chi2"].sum()),
               "timing": "host wall clock around %d blocking library calls, max over ranks" % args.steps,
               "output_streaming": "every chunk's transit rows go to pinned staging on the copy stream and are scattered into the caller's arrays while the next chunk computes"}
        e2e_tt = B["tt"]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    out = {
        "metric": "system-steps/s w/ grad (TRAPPIST-1 batch)", "value": value, "unit": "system-steps/s", "n_gpus": world * ngpu, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(world * ngpu),
                   "batch_per_gpu": args.nsys, "window_steps": window,
                   "chunk_steps": int(cnt[6]), "jac_launches": int(cnt[7]),
                   "l2": "inputs larger than L2: operator stream %.1f GB + scalar stream %.1f GB per chunk, jac_step %.1f GB per GPU (L2 = 126 MB)" %
                         ((28 * 152 + 768) * 8 * args.nsys * int(cnt[6]) / 1e9, 56 * 12 * 8 * args.nsys * int(cnt[6]) / 1e9, args.nsys * 48 * 56 * 16 / 1e9),
                   "parallelism": ("one process, one multi-device plan (nbg_plan_create_multi), one host thread per GPU" if single else
                                   "one process per GPU (torchrun), systems sharded across GPUs, no collective"),
                   "timing": "host clock between two device-wide synchronisations around blocking calls" if single else "CUDA events on the plan's stream, max over ranks"},
        "clocks": clocks, "gpu_launches": int(cnt[4]),
        "rates": {"main_steps_per_s": value, "step_equivalents_per_s": world * float(cnt[0] + cnt[1] + cnt[2]) / (ms_max * 1e-3),
                  "jacobian_steps_per_s": world * float(cnt[5]) / (ms_max * 1e-3), "transits": int(cnt[3]),
                  "newton_iters_per_transit": float(cnt[1]) / max(1, int(cnt[3]))},
        "kernel_ms": {"traj_kernel": float(ksum[0]), "transit_kernel": float(ksum[1]), "jac_rx_kernel": float(ksum[2]), "phi_dense_kernel": float(ksum[5]),
                      "pair_op_kernel": float(ksum[6]), "transit_adjoint_kernel": float(ksum[7]), "other": float(ksum[3]), "total": float(ksum[4])},
        "roofline": roofline_block(L, local, NBODY, cnt, ksum, ngpu),
        "status_bits": {"nonfinite": int((status & 1 != 0).sum()), "transit_itmax": int((status & 2 != 0).sum()),
                        "event_overflow": int((status & 4 != 0).sum()), "ntt_overflow": int((status & 8 != 0).sum())},
        "chunk_reruns": retries,
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = cores * args.ref_systems_per_core
        rate, dt, r = cpu_run(ns, args.ref_window, cores, x, v, m, jac_init)
        out["cpu_baseline"] = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": "port",
                               "sample": "%d systems x %d steps (%.1f s), %s, one system per thread" % (ns, args.ref_window, dt, CPU_BUILD)}
        if e2e_tt is not None and args.e2e_input == "cartesian":
            # the CPU sample ran the first `ns` systems of this batch from the same state: their transits inside the bench window are the
            # same events the end-to-end arm returned -- compare them (the oracle as the checker of the timed arm)
            worst, ntr = 0.0, 0
            for b in range(min(ns, nsys)):
                for i in range(1, NBODY):
                    nk = int(min(e2e_tt.shape[1] - off[i], ntt[i], np.count_nonzero(e2e_tt[b, off[i]:off[i] + ntt[i]])))
                    ref = r["tt"][b, :nk, i]
                    if nk:
                        worst = max(worst, float(np.max(np.abs(e2e_tt[b, off[i]:off[i] + nk] - ref) / np.abs(ref))))
                        ntr += nk
            out["e2e"]["check"] = {"systems": int(min(ns, nsys)), "transits_compared": ntr, "max_rel_dev_tt_vs_cpu_oracle": worst, "tolerance": 1e-11,
                                   "ok": bool(ntr > 0 and worst < 1e-11)}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if plan is not None:
        L.nbg_plan_destroy(plan)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
