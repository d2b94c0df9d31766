#!/usr/bin/env python
"""A/B of library builds on BASELINE cfg 3 (outer solar system, N = 5, grad = false, h = 25 d, 16,384 systems) through the symbols every
build has.  usage: tools/ab_cfg3.py LIB [steps]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbodygradient.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from golden.outer_ss import outer_ss_cartesian, energy_angmom  # noqa: E402

lib, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20000
L = C.CDLL(lib)
nsys = 16384
m, x, v = outer_ss_cartesian()
rng = np.random.Generator(np.random.Philox(key=20211582))
xb = np.broadcast_to(x, (nsys, 5, 3)).copy(); xb[1:] *= 1 + 1e-8 * rng.standard_normal((nsys - 1, 5, 3))
vb = np.broadcast_to(v, (nsys, 5, 3)).copy()
mb = np.broadcast_to(m, (nsys, 5)).copy()
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
plan = C.c_void_p()
assert L.nbg_plan_create(C.byref(plan), C.c_int32(5), C.c_int64(nsys), C.c_int32(0), C.c_int64(0)) == 0
assert L.nbg_set_state(plan, ptr(xb), ptr(vb), ptr(mb), C.c_double(0.0), None, None, None, None, None) == 0
out = {}
for tag, n in (("warm", 256), ("timed", steps)):
    assert L.nbg_integrate_resident(plan, C.c_double(25.0), C.c_int64(n), C.c_double(0.0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)) == 0
    kt = np.zeros(8)
    L.nbg_last_timings(plan, ptr(kt))
    out[tag] = float(kt[4])
xo, vo = np.zeros_like(xb), np.zeros_like(vb)
assert L.nbg_get_state(plan, ptr(xo), ptr(vo), None, None, None, None, None, None, None) == 0
dE = max(abs(energy_angmom(m, xo[b], vo[b])[0] / energy_angmom(m, xb[b], vb[b])[0] - 1) for b in range(0, nsys, 64))
print(json.dumps({"lib": lib, "steps": steps, "device_ms": out["timed"], "warm_ms_256_steps": out["warm"], "system_steps_per_s": nsys * steps / (out["timed"] * 1e-3),
                  "max_abs_dE_over_E": dE}))
L.nbg_plan_destroy(plan)
