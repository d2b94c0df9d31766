#!/usr/bin/env python
"""cfg 3 (outer solar system, N = 5, grad = false, h = 25 d, 16,384 systems) over the full 10^6 steps: device time per 100,000-step call,
then the same length as ONE call on a fresh plan.  usage: tools/diag_cfg3.py [LIB]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbodygradient.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from golden.outer_ss import outer_ss_cartesian, energy_angmom  # noqa: E402

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nbodygradient.jl_b200", "csrc", "libnbgrad_b200.so")
L = C.CDLL(lib)
nsys = 16384
m, x, v = outer_ss_cartesian()
rng = np.random.Generator(np.random.Philox(key=20211582))
xb = np.broadcast_to(x, (nsys, 5, 3)).copy(); xb[1:] *= 1 + 1e-8 * rng.standard_normal((nsys - 1, 5, 3))
vb = np.broadcast_to(v, (nsys, 5, 3)).copy()
mb = np.broadcast_to(m, (nsys, 5)).copy()
ptr = lambda a: a.ctypes.data_as(C.c_void_p)


def fresh():
    plan = C.c_void_p()
    assert L.nbg_plan_create(C.byref(plan), C.c_int32(5), C.c_int64(nsys), C.c_int32(0), C.c_int64(0)) == 0
    assert L.nbg_set_state(plan, ptr(xb), ptr(vb), ptr(mb), C.c_double(0.0), None, None, None, None, None) == 0
    return plan


def run(plan, n):
    w = time.time()
    assert L.nbg_integrate_resident(plan, C.c_double(25.0), C.c_int64(n), C.c_double(0.0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)) == 0
    kt = np.zeros(8)
    L.nbg_last_timings(plan, ptr(kt))
    return float(kt[4]), float(kt[0]), (time.time() - w) * 1e3


def energy(plan):
    xo, vo = np.zeros_like(xb), np.zeros_like(vb)
    assert L.nbg_get_state(plan, ptr(xo), ptr(vo), None, None, None, None, None, None, None) == 0
    return max(abs(energy_angmom(m, xo[b], vo[b])[0] / energy_angmom(m, xb[b], vb[b])[0] - 1) for b in range(0, nsys, 64)), xo[0].tolist()


plan = fresh()
run(plan, 256)
for k in range(10):
    tot, traj, wall = run(plan, 100000)
    print(json.dumps({"segment": k, "steps": 100000, "device_ms": tot, "traj_kernel_ms": traj, "host_wall_ms": wall,
                      "system_steps_per_s": nsys * 100000 / (tot * 1e-3)}), flush=True)
dE, x0 = energy(plan)
print(json.dumps({"after_steps": 1000256, "max_abs_dE_over_E": dE, "system0_x": x0}), flush=True)
L.nbg_plan_destroy(plan)
plan = fresh()
run(plan, 256)
tot, traj, wall = run(plan, 1000000)
dE, x0 = energy(plan)
print(json.dumps({"one_call_steps": 1000000, "device_ms": tot, "traj_kernel_ms": traj, "host_wall_ms": wall,
                  "system_steps_per_s": nsys * 1e6 / (tot * 1e-3), "max_abs_dE_over_E": dE, "system0_x": x0}), flush=True)
L.nbg_plan_destroy(plan)
