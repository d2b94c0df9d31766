#!/usr/bin/env python
"""Generates tests/golden/cfg2_quad_system0.npz: BASELINE cfg 2 (TRAPPIST-1, N = 8, h = 0.06 d, tmax = 1600 d = 26,667 steps, grad = true,
TransitTiming) for the unperturbed system, evaluated by the oracle in __float128 (oracle/oracle_capi.cpp: nbgoq_transit_timing_grad) from
the same Float64 x, v, m, jac_init that the Float64 oracle and the GPU path receive.  The result is "the reference algorithm without
round-off": tests/test_gpu_parity.py::test_cfg2_full_length_vs_quad asserts |GPU - quad| <= 1.5 |oracle_f64 - quad| block by block.

Usage: python tools/gen_quad_golden.py [tmax] [out.npz]     (about an hour of one CPU core at tmax = 1600)
Stored: every transit time; every STRIDE-th stored transit row of dtdq0 / dtdelements; final x, v, jac_step.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.binding import Oracle, build  # noqa: E402

STRIDE = 3


def main():
    tmax = float(sys.argv[1]) if len(sys.argv) > 1 else 1600.0
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", "cfg2_quad_system0.npz")
    build()
    o = Oracle()
    el = np.loadtxt(os.path.join(ROOT, "tests", "golden", "elements.txt"), delimiter=",")
    n, t0, h = 8, 7257.0, 0.06
    ntt = o.ntt(tmax, el[1:, 1])
    x, v, jac = o.init_nbody(el, t0)
    t = time.time()
    r = o.quad_transit_timing_grad(x, v, el[:, 0], jac, t0, h, tmax, ntt)
    print("quad run: %.1f s, %d transits" % (time.time() - t, int(r["count"].sum())))
    rows = [(i, k) for i in range(n) for k in range(min(int(r["count"][i]), ntt))][::STRIDE]
    idx = np.array(rows, dtype=np.int32)
    np.savez_compressed(out, tmax=tmax, h=h, t0=t0, ntt=ntt, stride=STRIDE, tt=r["tt"], count=r["count"], rows=idx,
                        dtdq0_rows=np.stack([r["dtdq0"][i, k] for i, k in rows]), dtdelements_rows=np.stack([r["dtdelements"][i, k] for i, k in rows]),
                        x=r["x"], v=r["v"], jac_step_cm=r["jac_step_cm"])
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
