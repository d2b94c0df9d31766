#!/usr/bin/env python
"""A/B timing of two builds of libnbgrad_b200.so on the bench workload (65,536 TRAPPIST-1 systems, 64-step windows) through the symbols
both have: prints the per-kernel device times of `reps` windows.  usage: tools/ab_time.py LIB [reps]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbodygradient.jl_b200")):
    sys.path.insert(0, p)
os.environ["NBGRAD_ALLOW_STALE"] = "1"
import bench  # noqa: E402

lib, reps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
L = C.CDLL(lib)
L.nbg_last_error.restype = C.c_char_p
nsys, window = 65536, 64
elb, x, v, jac = bench.make_batch(nsys)
m = np.ascontiguousarray(elb[:, :, 0])
ntt = bench.ntt_window(window, elb)
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
plan = C.c_void_p()
assert L.nbg_plan_create(C.byref(plan), C.c_int32(8), C.c_int64(nsys), C.c_int32(0), C.c_int64(0)) == 0
assert L.nbg_set_state(plan, ptr(x), ptr(v), ptr(m), C.c_double(bench.T0), None, None, None, None, None) == 0
tot = np.zeros(8)
for r in range(reps + 2):
    rc = L.nbg_transit_timing_resident(plan, C.c_double(bench.H), C.c_double(window * bench.H), C.c_int32(0), ptr(ntt), C.c_int32(0), C.c_int32(1), None)
    assert rc == 0, L.nbg_last_error()
    kt = np.zeros(8)
    L.nbg_last_timings(plan, ptr(kt))
    if r >= 2:
        tot += kt
print(json.dumps({"lib": lib, "reps": reps, "NBG_RX_UNROLL": os.environ.get("NBG_RX_UNROLL"), "ms_per_window": {k: round(float(t) / reps, 2) for k, t in
                  zip(("traj", "transit", "jac", "other", "total", "phi_dense", "pair_op", "adjoint"), tot)},
                  "system_steps_per_s": nsys * window * reps / (tot[4] * 1e-3)}))
L.nbg_plan_destroy(plan)
