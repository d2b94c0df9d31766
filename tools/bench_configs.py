#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU, through the C ABI (not a bench.py line: bench.py
measures configs[1]; these are the parity-test configurations, timed for the record).

  cfg 3  outer solar system, N = 5, grad = false, h = 25 d, batch 16,384  -> system-steps/s, max |dE/E|, max |dL/L|
  cfg 4  N = 2..16, nested star + planets, h = 0.05 d, grad = true, (intr)(s, N) plain driver -> system-steps/s and
         canonical F_grad(N) TFLOP/s; with --generic, N = 9..14 also with NBG_FORCE_GENERIC_JAC=1 (shared-memory Jacobian kernel)

One JSON object per line on stdout.  Times are the device time of the call (CUDA events inside the library,
nbg_last_timings[4]); inputs are resident in HBM.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbodygradient.jl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def f_grad(n):
    return 3191 * n * n * (n - 1) + 168 * n ** 3 + 798 * n * n + 2400 * n * (n - 1)


def cfg4_elements(n, nsys, rng):
    el = np.zeros((n, 7)); el[0, 0] = 1.0
    for k in range(1, n):
        el[k] = [3e-5, 1.5 * 1.6 ** (k - 1), 0.1 * k, 0.01, 0.0, np.pi / 2, 0.0]
    elb = np.broadcast_to(el, (nsys, n, 7)).copy()
    elb[1:, 1:, 1] *= 1 + 1e-4 * rng.standard_normal((nsys - 1, n - 1))
    return elb


def device_ms(L, plan):
    from nbgrad import _lib
    kt = np.zeros(8)
    L.nbg_last_timings(plan, _lib.ptr(kt))
    return kt


def run_cfg4(L, n, nsys, steps, peak, generic=False, variant=None):
    from nbgrad import _lib
    rng = np.random.Generator(np.random.Philox(key=20211582 + n))
    elb = cfg4_elements(n, nsys, rng)
    if generic:
        os.environ["NBG_FORCE_GENERIC_JAC"] = "1"
    else:
        os.environ.pop("NBG_FORCE_GENERIC_JAC", None)
    if variant:
        os.environ["NBG_RX_UNROLL"] = str(variant)
    plan = C.c_void_p()
    _lib.check(L.nbg_plan_create(C.byref(plan), C.c_int32(n), C.c_int64(nsys), C.c_int32(0), C.c_int64(0)))
    os.environ.pop("NBG_FORCE_GENERIC_JAC", None)
    os.environ.pop("NBG_RX_UNROLL", None)
    el_t = np.ascontiguousarray(elb.transpose(0, 2, 1))
    _lib.check(L.nbg_set_state_elements(plan, _lib.ptr(el_t), None, C.c_double(0.0), C.c_int32(0)))
    h = 0.05
    _lib.check(L.nbg_integrate_resident(plan, C.c_double(h), C.c_int64(max(2, min(64, steps // 4))), C.c_double(0.0), C.c_int32(1), C.c_int32(0), C.c_double(0.0)))
    best = None
    for _ in range(2 if steps <= 256 else 1):
        _lib.check(L.nbg_integrate_resident(plan, C.c_double(h), C.c_int64(steps), C.c_double(0.0), C.c_int32(1), C.c_int32(0), C.c_double(0.0)))
        kt = device_ms(L, plan)
        if best is None or kt[4] < best[4]:
            best = kt.copy()
    status = np.zeros(nsys, dtype=np.uint32)
    _lib.check(L.nbg_get_state(plan, None, None, None, None, None, None, None, None, _lib.ptr(status)))
    L.nbg_plan_destroy(plan)
    rate = nsys * steps / (best[4] * 1e-3)
    tf = rate * f_grad(n) / 1e12
    return {"config": "cfg4", "nbody": n, "variant": variant, "batch": nsys, "steps": steps, "jacobian_kernel": "shared-memory (generic)" if generic else ("register-resident" if n <= 14 else "register-resident, single operator buffer"),
            "device_ms": float(best[4]), "system_steps_per_s": rate, "canonical_tflops": tf, "frac_fp64_peak": tf / peak if peak else None,
            "kernel_ms": {"traj": float(best[0]), "jac": float(best[2]), "phi_dense": float(best[5]), "pair_op": float(best[6])},
            "nonfinite": int((status & 1 != 0).sum())}


def run_cfg3(L, nsys, steps):
    from nbgrad import _lib
    from golden.outer_ss import outer_ss_cartesian, energy_angmom
    m, x, v = outer_ss_cartesian()
    rng = np.random.Generator(np.random.Philox(key=20211582))
    xb = np.broadcast_to(x, (nsys, 5, 3)).copy(); xb[1:] *= 1 + 1e-8 * rng.standard_normal((nsys - 1, 5, 3))
    vb = np.broadcast_to(v, (nsys, 5, 3)).copy()
    mb = np.broadcast_to(m, (nsys, 5)).copy()
    plan = C.c_void_p()
    _lib.check(L.nbg_plan_create(C.byref(plan), C.c_int32(5), C.c_int64(nsys), C.c_int32(0), C.c_int64(0)))
    _lib.check(L.nbg_set_state(plan, _lib.ptr(xb), _lib.ptr(vb), _lib.ptr(mb), C.c_double(0.0), None, None, None, None, None))
    h = 25.0
    _lib.check(L.nbg_integrate_resident(plan, C.c_double(h), C.c_int64(256), C.c_double(0.0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)))
    _lib.check(L.nbg_integrate_resident(plan, C.c_double(h), C.c_int64(steps), C.c_double(0.0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)))
    kt = device_ms(L, plan)
    xo, vo = np.zeros_like(xb), np.zeros_like(vb)
    _lib.check(L.nbg_get_state(plan, _lib.ptr(xo), _lib.ptr(vo), None, None, None, None, None, None, None))
    L.nbg_plan_destroy(plan)
    dE = dL = 0.0
    for b in range(0, nsys, max(1, nsys // 256)):
        E0, L0 = energy_angmom(m, xb[b], vb[b])
        E1, L1 = energy_angmom(m, xo[b], vo[b])
        dE = max(dE, abs(E1 / E0 - 1)); dL = max(dL, float(np.linalg.norm(L1 - L0) / np.linalg.norm(L0)))
    return {"config": "cfg3", "nbody": 5, "batch": nsys, "steps": steps + 256, "h": h, "grad": False, "device_ms": float(kt[4]),
            "system0_x": xo[0].tolist(), "system0_v": vo[0].tolist(),
            "system_steps_per_s": nsys * steps / (kt[4] * 1e-3), "max_abs_dE_over_E": dE, "max_abs_dL_over_L": dL,
            "note": "energy / angular momentum after %d steps of h = 25 d, sampled over 256 systems of the batch" % (steps + 256)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg3-steps", type=int, default=20000)
    ap.add_argument("--nmin", type=int, default=2)
    ap.add_argument("--nmax", type=int, default=16)
    ap.add_argument("--skip-cfg3", action="store_true")
    ap.add_argument("--skip-cfg4", action="store_true")
    ap.add_argument("--generic", action="store_true", help="N = 9..14: also time the shared-memory Jacobian kernel (NBG_FORCE_GENERIC_JAC=1)")
    ap.add_argument("--cfg4-steps", type=int, default=10000, help="steps of the N sweep (BASELINE cfg 4: exactly 10^4)")
    args = ap.parse_args()
    from nbgrad import _lib
    L = _lib.lib()
    tfl, pms = C.c_double(0), C.c_double(0)
    L.nbg_fp64_peak(C.c_int32(0), C.byref(tfl), C.byref(pms))
    print(json.dumps({"fp64_peak_tflops_measured": tfl.value}), flush=True)
    if not args.skip_cfg3:
        print(json.dumps(run_cfg3(L, 16384, args.cfg3_steps)), flush=True)
    for n in range(args.nmin, (args.nmax if not args.skip_cfg4 else args.nmin - 1) + 1):
        # batch: fills the GPU many times over (>= 55 waves of one block per SM) while keeping the 10^4-step run of the largest N short
        nsys = 65536 if n <= 8 else (32768 if n <= 10 else (16384 if n <= 12 else 8192))
        steps = args.cfg4_steps
        print(json.dumps(run_cfg4(L, n, nsys, steps, tfl.value)), flush=True)
        if 9 <= n <= 14 and args.generic:
            print(json.dumps(run_cfg4(L, n, nsys, min(steps, 64), tfl.value, generic=True)), flush=True)


if __name__ == "__main__":
    main()
