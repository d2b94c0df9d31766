"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/nbgrad.h declares,
argument errors are reported, and without a CUDA device compute entry points refuse loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def nb():
    from nbgrad.build import build
    build()
    import nbgrad
    return nbgrad


def test_header_symbols_exported(nb):
    hdr = open(os.path.join(ROOT, "include", "nbgrad.h")).read()
    declared = set(re.findall(r"\b(nbg_[a-z_0-9]+)\s*\(", hdr)) - {"nbg_plan"}
    assert declared == set(nb.SYMBOLS), declared ^ set(nb.SYMBOLS)
    L = nb.lib()
    for s in declared:
        assert hasattr(L, s)
    assert L.nbg_version() >= 100


def test_header_cites_reference(nb):
    hdr = open(os.path.join(ROOT, "include", "nbgrad.h")).read()
    for cite in ("Integrator.jl:159-197", "Transits.jl:140-180", "timing.jl:3-194"):
        assert cite in hdr


def test_no_cpu_fallback(nb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = nb.lib()
    assert L.nbg_device_count() == 0
    p = C.c_void_p()
    rc = L.nbg_plan_create(C.byref(p), C.c_int32(3), C.c_int64(4), C.c_int32(0), C.c_int64(0))
    assert rc == -2  # NBG_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.nbg_last_error()
    ic = nb.ElementsIC(0.0, 3, nb.trappist1_elements())
    s = nb.State(ic)  # host-side IC layer works anywhere
    with pytest.raises(nb.NbgError):
        nb.Integrator(0.05, 1.0)(s, 2)


def test_argument_errors(nb):
    L = nb.lib()
    assert L.nbg_plan_create(None, C.c_int32(3), C.c_int64(4), C.c_int32(0), C.c_int64(0)) == -1
    p = C.c_void_p()
    assert L.nbg_plan_create(C.byref(p), C.c_int32(1), C.c_int64(4), C.c_int32(0), C.c_int64(0)) == -1
    assert b"nbody" in L.nbg_last_error()
    assert L.nbg_get_state(None, None, None, None, None, None, None, None, None, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nbodygradient.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), os.path.join(dp, f)


def test_host_ic_layer_matches_oracle(nb, oracle, elements):
    for n, t0 in ((8, 7257.0), (3, 7257.93115525)):
        x, v, j = nb.init_nbody_elements(elements[:n], t0)
        xo, vo, jo = oracle.init_nbody(elements[:n], t0)
        assert np.max(np.abs(x[0] - xo)) / np.max(np.abs(xo)) < 1e-14
        assert np.max(np.abs(v[0] - vo)) / np.max(np.abs(vo)) < 1e-14
        assert np.max(np.abs(j[0] - jo)) / np.max(np.abs(jo)) < 1e-12
    ic = nb.get_default_ICs("trappist-1", 7257.0)
    assert ic.nbody == 8 and np.array_equal(ic.elements[0], elements)
    tt = nb.TransitTiming(1600.0, ic)
    assert tt.ntt == 1062  # ceil(1600/1.5109)+3
    with pytest.raises(ValueError):
        nb.get_default_ICs("nope")
    with pytest.raises(ValueError):
        nb.Elements(m=1.0, P=1.0, ecosw=1.2)


def test_runtime_knobs_are_documented():
    # every environment variable the library reads is listed in INTEGRATION.md section 6 (and nothing stale is listed)
    src = ""
    csrc = os.path.join(ROOT, "nbodygradient.jl_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")):
            src += open(os.path.join(csrc, f)).read()
    read = set(re.findall(r'getenv\("(NBG_[A-Z0-9_]+)"\)', src))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    listed = set(re.findall(r"`(NBG_[A-Z0-9_]+)`", doc.split("## 6.")[1])) - {"NBG_JAC_MMA=1"}
    assert read and read == listed, read ^ listed
