#!/usr/bin/env python
"""Generates tests/golden/cfg2_quad_ensemble.npz: K perturbed TRAPPIST-1 systems (the cfg 2 ensemble recipe) integrated for TMAX days
with grad = true by the oracle in __float128 (nbgoq_transit_timing_grad) from the Float64 x, v, m, jac_init.  One trajectory is one
realisation of a random walk of rounding errors, so "the GPU's round-off is no worse than the reference's" can only be tested over an
ensemble: tests/test_gpu_parity.py::test_roundoff_no_worse_than_reference compares the RMS over the ensemble of |GPU - exact| with the
RMS of |Float64 oracle - exact|.

Usage: python tools/gen_quad_ensemble.py [K=16] [TMAX=100]     (K x ~2.3 min of CPU at TMAX = 100, spread over the cores)
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.binding import Oracle, build  # noqa: E402

T0, H, SEED, STRIDE = 7257.0, 0.06, 101, 3


def ensemble(K):
    el = np.loadtxt(os.path.join(ROOT, "tests", "golden", "elements.txt"), delimiter=",")
    rng = np.random.default_rng(SEED)
    elb = np.broadcast_to(el, (K, 8, 7)).copy()
    elb[:, 1:, 0] *= 1 + 1e-4 * rng.standard_normal((K, 7))
    elb[:, 1:, 1] *= 1 + 1e-4 * rng.standard_normal((K, 7))
    elb[:, 1:, 3:5] += 1e-4 * rng.standard_normal((K, 7, 2))
    return elb


def one(args):
    el, tmax = args
    o = Oracle()
    ntt = o.ntt(tmax, np.full(7, 1.5))
    x, v, jac = o.init_nbody(el, T0)
    r = o.quad_transit_timing_grad(x, v, el[:, 0], jac, T0, H, tmax, ntt)
    rows = [(i, k) for i in range(8) for k in range(min(int(r["count"][i]), ntt))][::STRIDE]
    return dict(tt=r["tt"], count=r["count"], rows=np.array(rows, dtype=np.int32), dtdq0_rows=np.stack([r["dtdq0"][i, k] for i, k in rows]),
                dtdelements_rows=np.stack([r["dtdelements"][i, k] for i, k in rows]), x=r["x"], v=r["v"], jac_step_cm=r["jac_step_cm"], ntt=ntt)


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    tmax = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
    build()
    elb = ensemble(K)
    t = time.time()
    with Pool(min(K, os.cpu_count() or 1)) as pool:
        res = pool.map(one, [(elb[k], tmax) for k in range(K)])
    print("%d quad runs: %.0f s" % (K, time.time() - t))
    nrow = min(len(r["rows"]) for r in res)
    out = os.path.join(ROOT, "tests", "golden", "cfg2_quad_ensemble.npz")
    np.savez_compressed(out, elements=elb, tmax=tmax, h=H, t0=T0, ntt=res[0]["ntt"], tt=np.stack([r["tt"] for r in res]), count=np.stack([r["count"] for r in res]),
                        rows=np.stack([r["rows"][:nrow] for r in res]), dtdq0_rows=np.stack([r["dtdq0_rows"][:nrow] for r in res]),
                        dtdelements_rows=np.stack([r["dtdelements_rows"][:nrow] for r in res]), x=np.stack([r["x"] for r in res]),
                        v=np.stack([r["v"] for r in res]), jac_step_cm=np.stack([r["jac_step_cm"] for r in res]))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
