#!/usr/bin/env python
"""Diagnostic for test_transit_parameters_and_ti[1]: which entries of the v_sky gradient deviate from the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nbodygradient.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import nbgrad as nb
from oracle.binding import Oracle
o = Oracle()
elements = np.loadtxt(os.path.join(ROOT, "tests", "golden", "elements.txt"), delimiter=",")
T0 = 7257.93115525
ti = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, t0 = 3, T0 - 7300.0 - 0.5
el = elements[:N].copy(); el[1:, 2] -= 7300.0; el[:, 6] = 0; el[1, 0] *= 10; el[2, 0] *= 10
h, tmax = 0.04, 10.0
ic = nb.ElementsIC(t0, N, el)
s, tp = nb.State(ic), nb.TransitParameters(tmax, ic, ti)
nb.Integrator(h, tmax)(s, tp)
x, v, jac = o.init_nbody(el, t0)
so = o.new_state(x, v, el[:, 0], t0)
r = o.transit_timing(so, h, tmax, tp.ntt, ti=ti, grad=True, jac_init=jac, ntbv=3)
print("env", {k: v for k, v in os.environ.items() if k.startswith("NBG_")}, "ti", ti, "count", tp.count[0], r["count"])
for c in range(3):
    g, q = tp.dtbvdq0[0, c], r["dtdq0"][c]
    print("comp", c, "max-norm rel", np.max(np.abs(g - q)) / max(np.max(np.abs(q)), 1e-300))
    for i in range(N):
        for k in range(int(min(r["count"][i], tp.ntt))):
            d = np.abs(g[i, k] - q[i, k]); sc = np.max(np.abs(q[i, k]))
            if sc > 0 and d.max() / sc > 1e-11:
                qq, pp = np.unravel_index(np.argmax(d), d.shape)
                print("   body %d transit %d: row rel err %.3e at (q=%d, p=%d): gpu %.6e oracle %.6e; tt %.6f" % (i, k, d.max() / sc, qq, pp, g[i, k, qq, pp], q[i, k, qq, pp], tp.ttbv[0, 0, i, k]))
