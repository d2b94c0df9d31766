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
          of tt, count, dtdq0, dtdelements, x, v inside the timed region
  --impl reference : the reference's CPU path (the oracle restatement; Julia cannot run here) on all host cores.
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


def make_batch(nsys, rank=0):
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
    x, v, jac_init = nb.init_nbody_elements(elb, T0)
    return elb, x, v, jac_init


def ntt_window(window, elb):
    pmin = elb[:, 1:, 1].min(axis=0)
    ntt = np.zeros(NBODY, dtype=np.int32)
    ntt[1:] = np.ceil(window * H / pmin).astype(np.int32) + 2
    return ntt


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


def cpu_run(nsys, window, nthreads, x, v, m, jac_init):
    """Reference CPU path (oracle, -O3 build) on `nsys` systems for `window` steps, one system per thread."""
    from oracle.binding import Oracle, build
    try:
        build(fast_native=True)  # rebuild the timing build for this host's ISA
    except Exception:
        pass
    o = Oracle(fast=True)
    tmax = window * H
    ntt = int(np.ceil(tmax / 1.5) + 3)
    jcm = np.ascontiguousarray(jac_init[:nsys].transpose(0, 2, 1))
    t = time.perf_counter()
    r = o.batch_transit_timing(x[:nsys], v[:nsys], m[:nsys], T0, H, tmax, ntt, grad=True, jac_init_cm=jcm, nthreads=nthreads)
    dt = time.perf_counter() - t
    return nsys * window / dt, dt, r


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
           "config": {"workload": "TRAPPIST-1 N=8 h=0.06 grad=true TransitTiming (cfg 2)", "batch": 65536, "sample_batch": nsys, "window_steps": window},
           "cpu_baseline": {"value": value, "unit": "system-steps/s", "cores": cores, "kind": "port",
                            "sample": "%d systems x %d steps per step, oracle -O3 build, one system per thread" % (nsys, window)},
           "e2e": {"value": value, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nsys", type=int, default=65536)
    ap.add_argument("--window", type=int, default=64)
    ap.add_argument("--ref-window", type=int, default=4096, help="steps per system of the CPU-baseline sample")
    ap.add_argument("--ref-systems-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-input", choices=["elements", "cartesian"], default="cartesian",
                    help="what crosses PCIe on the way in: orbital elements (init_nbody on the device) or x, v, m + host-computed jac_init")
    ap.add_argument("--e2e-output", choices=["arrays", "chi2"], default="arrays",
                    help="with --e2e-input elements: copy out tt/dtdq0/dtdelements, or only the fused chi^2 and its gradients")
    ap.add_argument("--stream-budget-gb", type=float, default=0.0, help="HBM for the operator streams of a chunk (0 = the library default, 1/4 of free memory)")
    ap.add_argument("--e2e-slices", type=int, default=1, help="slices (plans + host threads) of the end-to-end arm")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import nbgrad as nb
    from nbgrad import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    nsys, window = args.nsys, args.window
    elb, x, v, jac_init = make_batch(nsys, rank)
    m = np.ascontiguousarray(elb[:, :, 0])
    ntt = ntt_window(window, elb)
    RT, M = int(ntt.sum()), 7 * NBODY

    plan = C.c_void_p()
    _lib.check(L.nbg_plan_create(C.byref(plan), C.c_int32(NBODY), C.c_int64(nsys), C.c_int32(local), C.c_int64(int(args.stream_budget_gb * 1e9))))
    stream = torch.cuda.ExternalStream(L.nbg_cuda_stream(plan), device=torch.device("cuda", local))
    ptr = _lib.ptr

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    # ---- device-resident arm ----
    _lib.check(L.nbg_set_state(plan, ptr(x), ptr(v), ptr(m), C.c_double(T0), None, None, None, None, None))
    tmaxw = window * H

    def step_resident():
        _lib.check(L.nbg_transit_timing_resident(plan, C.c_double(H), C.c_double(tmaxw), C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), None))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    _lib.check(L.nbg_counters_reset(plan))
    kt = np.zeros(8)
    ksum = np.zeros(8)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        L.nbg_last_timings(plan, ptr(kt))
        ksum += kt
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    cnt = np.zeros(8, dtype=np.int64)
    L.nbg_counters(plan, ptr(cnt))
    status = np.zeros(nsys, dtype=np.uint32)
    _lib.check(L.nbg_get_state(plan, None, None, None, None, None, None, None, None, ptr(status)))
    main_steps = nsys * window * args.steps
    value = world * main_steps / (ms_max * 1e-3)

    # ---- end-to-end arm: the one-shot C-ABI call with HOST (pinned) buffers ----
    # The batch is cut into `--e2e-slices` slices, each with its own plan (= its own stream) driven by its own host thread --
    # "different plans may be used from different threads" (include/nbgrad.h) -- so one slice's H2D/D2H copies overlap the
    # other slices' kernels.  Every byte of input and output crosses PCIe inside the timed region, every step.
    e2e = None
    if not args.no_e2e:
        _lib.check(L.nbg_plan_destroy(plan))
        plan = None
        K = max(1, min(args.e2e_slices, nsys))
        free_b, _tot = torch.cuda.mem_get_info()
        bounds = [nb.shard_range(nsys, k, K) for k in range(K)]
        slices = []
        h2d = d2h = 0
        for lo, hi in bounds:
            ns = hi - lo
            pl = C.c_void_p()
            _lib.check(L.nbg_plan_create(C.byref(pl), C.c_int32(NBODY), C.c_int64(ns), C.c_int32(local),
                                         C.c_int64(int(args.stream_budget_gb * 1e9 / K) if args.stream_budget_gb > 0 else (0 if K == 1 else int(free_b // (5 * K))))))
            bufs = dict(x=pin(x[lo:hi])[1], v=pin(v[lo:hi])[1], m=pin(m[lo:hi])[1], j=pin(jac_init[lo:hi].transpose(0, 2, 1))[1],
                        el=pin(elb[lo:hi].transpose(0, 2, 1))[1], tobs=pin(np.full(RT, T0 + 1.0))[1], sig=pin(np.full(RT, 1e-3))[1],
                        chi2=pin(np.zeros(ns))[1], gq=pin(np.zeros((ns, M)))[1], ge=pin(np.zeros((ns, M)))[1],
                        tt=pin(np.zeros((ns, RT)))[1], c=pin(np.zeros((ns, NBODY), dtype=np.int64))[1], d=pin(np.zeros((ns, RT, M)))[1],
                        e=pin(np.zeros((ns, RT, M)))[1], xo=pin(np.zeros((ns, NBODY, 3)))[1], vo=pin(np.zeros((ns, NBODY, 3)))[1])
            h2d += sum(bufs[k].nbytes for k in (("x", "v", "m", "j") if args.e2e_input == "cartesian" else ("el",)))
            d2h += sum(bufs[k].nbytes for k in (("chi2", "gq", "ge") if (args.e2e_output == "chi2" and args.e2e_input == "elements")
                                                else ("tt", "c", "d", "e", "xo", "vo")))
            slices.append((pl, bufs))

        def step_e2e(pl, B):
            if args.e2e_input == "cartesian":   # x, v, m and the host-computed jac_init go up (1.6 GB of jac_init)
                _lib.check(L.nbg_transit_timing(pl, ptr(B["x"]), ptr(B["v"]), ptr(B["m"]), None, C.c_double(T0), C.c_double(H), C.c_double(tmaxw),
                                                C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), ptr(B["j"]), ptr(B["tt"]), ptr(B["c"]), ptr(B["d"]),
                                                ptr(B["e"]), ptr(B["xo"]), ptr(B["vo"]), None, None, None, None, None, None, None))
            else:   # the reference's user-level sequence ElementsIC -> State -> intr(s, tt): orbital elements go up, init_nbody runs on the device
                _lib.check(L.nbg_set_state_elements(pl, ptr(B["el"]), None, C.c_double(T0), C.c_int32(1)))
                _lib.check(L.nbg_transit_timing_resident(pl, C.c_double(H), C.c_double(tmaxw), C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), None))
                if args.e2e_output == "chi2":   # fused likelihood: chi^2 + gradients (1 + 2M doubles per system) instead of the arrays
                    _lib.check(L.nbg_transit_chi2(pl, ptr(B["tobs"]), ptr(B["sig"]), C.c_int32(0), ptr(B["chi2"]), ptr(B["gq"]), ptr(B["ge"])))
                else:
                    _lib.check(L.nbg_transit_fetch(pl, ptr(B["tt"]), ptr(B["c"]), ptr(B["d"]), ptr(B["e"])))
                    _lib.check(L.nbg_get_state(pl, ptr(B["xo"]), ptr(B["vo"]), None, None, None, None, None, None, None))

        gate = threading.Barrier(K + 1)
        errs = []

        def worker(pl, B):
            try:
                torch.cuda.set_device(local)
                step_e2e(pl, B)          # warm-up (allocations inside the plan)
                gate.wait()
                for _ in range(args.steps):
                    step_e2e(pl, B)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
                gate.abort()

        threads = [threading.Thread(target=worker, args=sl) for sl in slices]
        [t.start() for t in threads]
        try:
            gate.wait()
        except threading.BrokenBarrierError:
            pass
        barrier_t0 = time.perf_counter()
        [t.join() for t in threads]
        torch.cuda.synchronize()
        wall = time.perf_counter() - barrier_t0
        if errs:
            raise errs[0]
        t2 = torch.tensor([wall * 1e3], device="cuda", dtype=torch.float64)   # blocking host calls: wall clock is the end-to-end time
        if dist is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": world * main_steps / (float(t2.item()) * 1e-3), "unit": "system-steps/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "slices": K, "input": args.e2e_input, "output": args.e2e_output if args.e2e_input == "elements" else "arrays", "transits_checked": int(sum(B["c"].sum() for _, B in slices)), "chi2_sum": float(sum(B["chi2"].sum() for _, B in slices)),
               "timing": "host wall clock around %d blocking nbg_transit_timing calls per slice, max over ranks" % args.steps,
               "output_streaming": "the last chunk's Jacobian kernel runs in %s batch slices, each copied out while the next computes" % os.environ.get("NBG_OUT_SLICES", "8")}
        for pl, _B in slices:
            L.nbg_plan_destroy(pl)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    tfl, pms = C.c_double(0), C.c_double(0)
    L.nbg_fp64_peak(C.c_int32(local), C.byref(tfl), C.byref(pms))
    peak = tfl.value
    names = ["traj_kernel", "transit_kernel", "jac_kernel"]
    dom = 2 if ksum[2] >= max(ksum[0] + ksum[6], ksum[1]) else int(np.argmax(ksum[:2]))
    jac_steps = float(cnt[5])        # Jacobian system-steps applied (main + transit final steps)
    step_equiv = float(cnt[0] + cnt[1] + cnt[2])
    flops = {0: (f_grad(NBODY) - f_jac(NBODY)) * float(cnt[0]), 1: (f_grad(NBODY) - f_jac(NBODY)) * float(cnt[1] + cnt[2]),
             2: f_jac(NBODY) * jac_steps}
    dom_ms = ksum[dom] + (ksum[6] if dom == 0 else 0.0)
    achieved = flops[dom] / (dom_ms * 1e-3) / 1e12
    # measured DRAM traffic of the dominant kernel (ncu --set full, profiles/r01_traffic.json), scaled to this run's launches
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and dom == 2:
        tj = json.load(open(tpath))
        traffic = tj["jac_rx_kernel"]["dram_bytes_per_jacobian_step"] * jac_steps / max(1, int(cnt[7]))
    path_achieved = f_grad(NBODY) * step_equiv / (ksum[4] * 1e-3) / 1e12
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    # algorithmic HBM bytes per main system-step: scalars (written, read), compact phisalpha records (written, read), Kepler records
    # (written, read), dense phisalpha operator (written, read)
    stream_bytes = 2 * (56 * 32 + 28 * 24 + 56 * 64 + 768) * 8.0 * (float(cnt[0]) + float(cnt[2]))
    hbm_gbs = stream_bytes / (ksum[4] * 1e-3) / 1e9
    out = {
        "metric": "system-steps/s w/ grad (TRAPPIST-1 batch)", "value": value, "unit": "system-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "TRAPPIST-1 N=8 h=0.06 grad=true TransitTiming (BASELINE cfg 2)", "batch_per_gpu": nsys, "window_steps": window,
                   "chunk_steps": int(cnt[6]), "jac_launches": int(cnt[7]),
                   "l2": "inputs larger than L2: operator stream %.1f GB + scalar stream %.1f GB per chunk, jac_step %.1f GB (L2 = 126 MB)" %
                         ((28 * 152 + 768) * 8 * nsys * int(cnt[6]) / 1e9, 56 * 32 * 8 * nsys * int(cnt[6]) / 1e9, nsys * 48 * 56 * 16 / 1e9),
                   "parallelism": "systems sharded across GPUs, no collective"},
        "clocks": clocks, "gpu_launches": int(cnt[4]),
        "rates": {"main_steps_per_s": value, "step_equivalents_per_s": world * step_equiv / (ms_max * 1e-3),
                  "jacobian_steps_per_s": world * jac_steps / (ms_max * 1e-3), "transits": int(cnt[3]),
                  "newton_iters_per_transit": float(cnt[1]) / max(1, int(cnt[3]))},
        "kernel_ms": {names[k]: float(ksum[k]) for k in range(3)} | {"phi_dense_kernel": float(ksum[5]), "pair_op_kernel": float(ksum[6]), "other": float(ksum[3]), "total": float(ksum[4])},
        "roofline": {"bound": "fp64", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_note": "dram read+write bytes per launch: ncu-measured bytes per Jacobian step x steps per launch",
                     "flops": "canonical (SURVEY 8d): F_jac(8)=%d per Jacobian step, F_scalar(8)=%d per trajectory step" %
                     (f_jac(NBODY), f_grad(NBODY) - f_jac(NBODY)),
                     "peak_source": "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry; nominal 37.2)",
                     "path": {"achieved": path_achieved, "frac": path_achieved / peak if peak else None,
                              "note": "whole path: F_grad(8)=%d x step-equivalents / total device time" % f_grad(NBODY)},
                     "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks.get("hbm_gbs"), "note": "operator / scalar streams written once and read once, algorithmic bytes over total device time"}},
        "status_bits": {"nonfinite": int((status & 1 != 0).sum()), "transit_itmax": int((status & 2 != 0).sum()),
                        "event_overflow": int((status & 4 != 0).sum()), "ntt_overflow": int((status & 8 != 0).sum())},
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = cores * args.ref_systems_per_core
        rate, dt, r = cpu_run(ns, args.ref_window, cores, x, v, m, jac_init)
        out["cpu_baseline"] = {"value": rate, "unit": "system-steps/s", "cores": cores, "kind": "port",
                               "sample": "%d systems x %d steps (%.1f s), oracle -O3 build, one system per thread" % (ns, args.ref_window, dt)}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if plan is not None:
        L.nbg_plan_destroy(plan)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
