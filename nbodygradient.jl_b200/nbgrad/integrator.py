"""Host-side mirror of the reference's driver API for the hot path, over the C ABI (include/nbgrad.h).

  State                 src/integrator/Integrator.jl:49-103
  Integrator            src/integrator/Integrator.jl:17-31, callable forms :159-247
  TransitTiming         src/transits/Transits.jl:14-56
  TransitParameters     src/transits/Transits.jl:68-110
  (intr)(s, tt; grad)   src/transits/Transits.jl:140-180

Same names, argument meaning and error behaviour; every object carries a leading batch axis B (B = 1 is the
reference's single-system call).  Index order is numpy's, i.e. the reverse of Julia's and 0-based:
Julia s.x[k,i] -> s.x[b,i,k];  jac_step[r,c] -> s.jac_step[b,r,c];  tt.tt[i,k] -> tt.tt[b,i,k];
tt.dtdq0[i,k,q,p] -> tt.dtdq0[b,i,k,q,p];  body indices (ti, occs) are 0-based.
All arithmetic of the path runs in libnbgrad_b200.so on the GPU; nothing here computes on the CPU.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check, ptr


def ahl21(*_a, **_k):
    """Placeholder for the reference's scheme function ahl21! (the only scheme of this build)."""
    raise _lib.NbgError("ahl21 is executed on the device through Integrator(...)")


import atexit
from collections import OrderedDict

_plans = OrderedDict()   # (n, nsys, devices, stream_budget) -> plan; least recently used first
MAX_PLANS = 4            # a plan keeps its device buffers (operator streams: up to 1/4 of the free memory): callers that vary the batch
                         # shape must not accumulate them


def _devices(device):
    """device=: an int (one GPU) or a sequence of ints (nbg_plan_create_multi: the batch is cut into one contiguous slice per entry)."""
    if isinstance(device, (int, np.integer)):
        return (int(device),)
    return tuple(int(d) for d in device)


def _plan(n, nsys, device=0, stream_budget=0):
    devs = _devices(device)
    key = (n, nsys, devs, stream_budget)
    if key in _plans:
        _plans.move_to_end(key)
        return _plans[key]
    while len(_plans) >= MAX_PLANS:
        _, old = _plans.popitem(last=False)
        _lib.lib().nbg_plan_destroy(old)
    p = C.c_void_p()
    if len(devs) == 1:
        check(_lib.lib().nbg_plan_create(C.byref(p), C.c_int32(n), C.c_int64(nsys), C.c_int32(devs[0]), C.c_int64(stream_budget)))
    else:
        d = np.asarray(devs, dtype=np.int32)
        check(_lib.lib().nbg_plan_create_multi(C.byref(p), C.c_int32(n), C.c_int64(nsys), ptr(d), C.c_int32(len(devs)), C.c_int64(stream_budget)))
    _plans[key] = p
    return p


def release_plans():
    for p in _plans.values():
        _lib.lib().nbg_plan_destroy(p)
    _plans.clear()


atexit.register(release_plans)


def check_step(t0, tmax):
    """Integrator.jl:249-259."""
    sg = lambda x: (x > 0) - (x < 0)
    if abs(tmax) > abs(t0):
        return sg(tmax)
    if sg(tmax) != sg(t0):
        return sg(tmax)
    return -1 * sg(tmax)


class State:
    """State(ic) — Integrator.jl:82-103."""

    def __init__(self, ic, on_device=False, device=0):
        """on_device=True: init_nbody (elements -> x, v, jac_init) runs in libnbgrad_b200 (nbg_set_state_elements) instead of
        the numpy IC layer; the state stays resident, so the next Integrator call skips the upload of x, v, m and jac_init."""
        self._resident_plan = None
        self._resident_copy = None
        if on_device:
            if not hasattr(ic, "elements"):
                raise TypeError("State(ic, on_device=True) needs an ElementsIC")
            x, v, jac_init = self._init_on_device(ic, device)
        else:
            x, v, jac_init = ic.init_nbody()
        B, n = x.shape[0], ic.nbody
        M = 7 * n
        self.n = n
        self.nsys = B
        self.x, self.v = np.ascontiguousarray(x), np.ascontiguousarray(v)
        self.m = ic.m  # by reference, as in the reference (quirk: Integrator.jl:101)
        self.t = np.full(B, float(ic.t0))
        self.jac_step = np.broadcast_to(np.eye(M), (B, M, M)).copy()
        self.jac_error = np.zeros((B, M, M))
        self.dqdt = np.zeros((B, M))
        self.dqdt_error = np.zeros((B, M))
        self.jac_init = jac_init if jac_init is not None else np.zeros((B, 0, 0))
        self.xerror = np.zeros((B, n, 3))
        self.verror = np.zeros((B, n, 3))
        self.pair = np.zeros((n, n), dtype=bool)
        self.status = np.zeros(B, dtype=np.uint32)
        self._fresh = True   # no integrator call has touched this State yet: jac_step = I, the error terms and dq/dh are zero

    def _init_on_device(self, ic, device):
        L = _lib.lib()
        B, n = ic.elements.shape[0], ic.nbody
        M = 7 * n
        plan = _plan(n, B, device, 0)
        self._resident_t0 = float(ic.t0)
        el = ic.elements.copy()
        el[:, :, 0] = ic.m                                               # masses live in ic.m (as in ic.init_nbody)
        el = np.ascontiguousarray(el.transpose(0, 2, 1))                # [sys][c][i] = Julia elements[i,c], system slowest
        eps = np.asfortranarray(np.asarray(ic.eps, dtype=np.float64))   # Julia column-major n x n
        check(L.nbg_set_state_elements(plan, ptr(el), ptr(eps), C.c_double(float(ic.t0)), C.c_int32(1 if ic.der else 0)))
        x, v = np.empty((B, n, 3)), np.empty((B, n, 3))
        check(L.nbg_get_state(plan, ptr(x), ptr(v), None, None, None, None, None, None, None))
        jac_init = None
        if ic.der:
            jcm = np.empty((B, M, M))
            check(L.nbg_get_jac_init(plan, ptr(jcm)))
            jac_init = jcm.transpose(0, 2, 1).copy()
        # The device already holds this state: the first Integrator call on it skips the upload -- but only if the plan (shared by every
        # State / Integrator of this shape) has not been given another state or stepped since: nbg_state_generation is bumped by every
        # call that changes the resident state, and the host copies are compared bit for bit (the arrays are public and may be edited).
        self._resident_plan = plan
        self._resident_gen = int(L.nbg_state_generation(plan))
        self._resident_copy = (x.copy(), v.copy(), np.array(ic.m, dtype=np.float64), None if jac_init is None else jac_init.copy())
        return x, v, jac_init

    def copy(self):
        import copy
        s = copy.copy(self)
        s._resident_plan = None
        s._resident_copy = None
        for k, val in self.__dict__.items():
            if isinstance(val, np.ndarray) and k != "m":
                setattr(s, k, val.copy())
        return s

    def __repr__(self):  # Base.show(::State): Integrator.jl:137-143
        f = lambda a: "finite" if np.all(np.isfinite(a)) else "infinite!"
        return "State{Float64}:\nPositions  : %s\nVelocities : %s\nJacobian   : %s" % (f(self.x), f(self.v), f(self.jac_step))

    # -- device transfer helpers
    def _upload(self, plan, with_jac):
        """Returns True if the state (and jac_init) was already resident from State(ic, on_device=True)."""
        L = _lib.lib()
        if self._resident_plan is not None:
            # still the state the device computed, and still the one the plan holds?
            cx, cv, cm, cj = self._resident_copy
            fresh = (self._resident_plan.value == plan.value and int(L.nbg_state_generation(plan)) == self._resident_gen and self._fresh
                     and not self.jac_error.any() and not self.xerror.any() and not self.verror.any() and not self.dqdt.any()
                     and np.array_equal(self.x, cx) and np.array_equal(self.v, cv) and np.array_equal(np.asarray(self.m), cm)
                     and (cj is None or np.array_equal(self.jac_init, cj)) and float(self.t[0]) == self._resident_t0)
            self._resident_plan = None
            self._resident_copy = None
            if fresh:
                pr = np.asfortranarray(self.pair.astype(np.uint8))
                check(L.nbg_set_pair(plan, ptr(pr) if pr.any() else None))
                return True
        pr = np.asfortranarray(self.pair.astype(np.uint8))  # Julia layout: [i,j] at i + n*j
        check(L.nbg_set_pair(plan, ptr(pr) if pr.any() else None))
        self._m_c = np.ascontiguousarray(self.m, dtype=np.float64)
        js = np.ascontiguousarray(self.jac_step.transpose(0, 2, 1)) if with_jac else None
        je = np.ascontiguousarray(self.jac_error.transpose(0, 2, 1)) if with_jac else None
        check(L.nbg_set_state(plan, ptr(self.x), ptr(self.v), ptr(self._m_c), C.c_double(float(self.t[0])), ptr(self.xerror), ptr(self.verror),
                              ptr(js), ptr(je), ptr(self.dqdt) if with_jac else None))
        return False

    def _download(self, plan, with_jac):
        L = _lib.lib()
        B, M = self.nsys, 7 * self.n
        js = np.empty((B, M, M)) if with_jac else None
        je = np.empty((B, M, M)) if with_jac else None
        check(L.nbg_get_state(plan, ptr(self.x), ptr(self.v), ptr(self.xerror), ptr(self.verror), ptr(js), ptr(je), ptr(self.dqdt) if with_jac else None,
                              ptr(self.t), ptr(self.status)))
        if with_jac:
            self.jac_step[...] = js.transpose(0, 2, 1)
            self.jac_error[...] = je.transpose(0, 2, 1)
        self._fresh = False


def dState(ic):
    """dState(ic) — Integrator.jl:106-110; the Derivatives scratch lives on the device, so only the State is returned."""
    return State(ic), None


class _TransitOutput:
    ncomp = 1

    def __init__(self, tmax, ic, ti=0, ntt=None):
        n = ic.nbody
        if ntt is None:
            if not hasattr(ic, "elements"):
                raise TypeError("TransitTiming(tmax, ic) needs an ElementsIC (Transits.jl:42-45); pass ntt= for other ICs")
            with np.errstate(divide="ignore", invalid="ignore"):
                q = float(tmax) / ic.elements[:, :, 1]
            fin = np.isfinite(q)
            ntt = int(np.max(np.ceil(np.abs(q[fin]))) + 3)  # Transits.jl:44-45
        B = ic.m.shape[0]
        self.n, self.nsys, self.ntt, self.ti = n, B, int(ntt), int(ti)
        self.occs = [i for i in range(n) if i != ti]
        self.count = np.zeros((B, n), dtype=np.int64)
        self._alloc(B, n, self.ntt)

    def zero_out(self):  # zero_out!(tt): Transits.jl:127-134
        for k, val in self.__dict__.items():
            if isinstance(val, np.ndarray):
                val[...] = 0


class TransitTiming(_TransitOutput):
    """TransitTiming(tmax, ic, ti) — tt[b,i,k], dtdq0[b,i,k,q,p], dtdelements[b,i,k,q,p], count[b,i]."""

    def _alloc(self, B, n, ntt):
        self.tt = np.zeros((B, n, ntt))
        self.dtdq0 = np.zeros((B, n, ntt, 7, n))
        self.dtdelements = np.zeros((B, n, ntt, 7, n))


class TransitParameters(_TransitOutput):
    """TransitParameters(tmax, ic, ti) — ttbv[b,c,i,k] (c = time, v_sky, b_sky^2), dtbvdq0[b,c,i,k,q,p], dtbvdelements."""
    ncomp = 3

    def _alloc(self, B, n, ntt):
        self.ttbv = np.zeros((B, 3, n, ntt))
        self.dtbvdq0 = np.zeros((B, 3, n, ntt, 7, n))
        self.dtbvdelements = np.zeros((B, 3, n, ntt, 7, n))


class ElementsOutput:
    """Orbital elements at sampled steps: get_orbital_elements (src/outputs/elements.jl:108-137) applied to the state before every
    `stride`-th step, as CartesianOutput does for x, v.  The conversion runs on the device on the resident state (nbg_orbital_elements);
    x, v never leave it.  o.elements[k, b, i, :] = (m, P, t0 = 0, ecosw, esinw, I, Omega, a, e, omega, tp) of body i; o.t[k]."""

    FIELDS = ("m", "P", "t0", "ecosw", "esinw", "I", "Omega", "a", "e", "omega", "tp")

    def __init__(self, nbody, nstep, stride=1, eps=None):
        self.nbody, self.nstep, self.stride = int(nbody), int(nstep), int(stride)
        self.eps = None if eps is None else np.asfortranarray(np.asarray(eps, dtype=np.float64))
        self.elements = self.t = None


class CartesianOutput:
    """CartesianOutput(nbody, nstep) -- Outputs.jl:7-17.  The reference deep-copies the whole State before every step; here x and v
    before every `stride`-th step are collected on the device (nbg_integrate_sampled): o.x[k, b, i, :], o.v[k, b, i, :], o.t[k].
    jac=True (needs grad=True) also keeps the saved States' jac_step: o.jac_step[k, b, :, :] (row, column as State.jac_step)."""

    def __init__(self, nbody, nstep, stride=1, jac=False):
        self.nbody, self.nstep, self.stride, self.jac = int(nbody), int(nstep), int(stride), bool(jac)
        self.x = self.v = self.t = self.jac_step = None


class Integrator:
    """Integrator(h, tmax) | Integrator(h, t0, tmax) | Integrator(scheme, h, t0, tmax) — Integrator.jl:17-31."""

    def __init__(self, *args, device=0, devices=None, stream_budget=0, keep_dense=False):
        """device / devices: one CUDA device index, or a list of them -- the batch is then cut into one contiguous slice per entry
        (nbg_plan_create_multi; one host thread per slice inside the library).  keep_dense: keep tt / dtdq0 / dtdelements of a transit
        call as dense arrays on the device (needed by chi2(); fails if they do not fit) instead of streaming rows chunk by chunk."""
        if len(args) and callable(args[0]):
            if args[0] is not ahl21:
                raise _lib.NbgError("NBG_ERR_UNSUPPORTED: only the ahl21 scheme is built")
            args = args[1:]
        if len(args) == 2:
            h, tmax = args
            t0 = 0.0
        elif len(args) == 3:
            h, t0, tmax = args
        else:
            raise TypeError("Integrator(h, tmax) | Integrator(h, t0, tmax) | Integrator(scheme, h, t0, tmax)")
        self.scheme, self.h, self.t0, self.tmax = ahl21, float(h), float(t0), float(tmax)
        self.device, self.stream_budget, self.keep_dense = (devices if devices is not None else device), stream_budget, keep_dense
        self.last_timings = None
        self._last_plan = self._last_tt = None
        self._last_dense = False

    def _p(self, s):
        return _plan(s.n, s.nsys, self.device, self.stream_budget)

    def _timings(self, plan):
        ms = np.zeros(8)
        check(_lib.lib().nbg_last_timings(plan, ptr(ms)))
        self.last_timings = dict(traj_ms=ms[0], transit_ms=ms[1], jac_ms=ms[2], other_ms=ms[3], total_ms=ms[4], phi_dense_ms=ms[5], pair_op_ms=ms[6])

    def __call__(self, s, arg=None, grad=True):
        if isinstance(arg, _TransitOutput):
            return self._transits(s, arg, grad)
        if isinstance(arg, CartesianOutput):
            return self._sampled(s, arg, grad)
        if isinstance(arg, ElementsOutput):
            return self._sampled_elements(s, arg, grad)
        if isinstance(arg, (int, np.integer)) and not isinstance(arg, bool):
            return self._nsteps(s, int(arg), grad)
        if arg is None:
            arg = float(s.t[0]) + self.tmax  # (intr)(s): Integrator.jl:247
        return self._to_time(s, float(arg), grad)

    # (intr)(s, time; grad) — Integrator.jl:159-197
    def _to_time(self, s, time, grad):
        t0 = float(s.t[0])
        nsteps = abs(int(np.rint((time - t0) / self.h)))
        h = self.h * check_step(t0, time)
        tmax = t0 + (h * nsteps)
        h_last = (time - tmax) if tmax != time else 0.0
        plan = self._p(s)
        s._upload(plan, grad)
        check(_lib.lib().nbg_integrate_resident(plan, C.c_double(h), C.c_int64(nsteps), C.c_double(h_last), C.c_int32(1 if grad else 0),
                                                C.c_int32(1), C.c_double(time)))
        s._download(plan, grad)
        self._timings(plan)

    # (intr)(s, N; grad) — Integrator.jl:211-234
    def _nsteps(self, s, N, grad):
        h = self.h
        if N < 0:
            h, N = -h, -N
        plan = self._p(s)
        s._upload(plan, grad)
        check(_lib.lib().nbg_integrate_resident(plan, C.c_double(h), C.c_int64(N), C.c_double(0.0), C.c_int32(1 if grad else 0), C.c_int32(0),
                                                C.c_double(0.0)))
        s._download(plan, grad)
        self._timings(plan)

    # (intr)(s, o::CartesianOutput) — Outputs.jl:26-49
    def _sampled(self, s, o, grad):
        t0 = float(s.t[0])
        h = self.h * check_step(t0, self.tmax)
        ns = (o.nstep + o.stride - 1) // o.stride
        plan = self._p(s)
        s._upload(plan, grad)
        o.x, o.v = np.zeros((ns, s.nsys, s.n, 3)), np.zeros((ns, s.nsys, s.n, 3))
        o.t = t0 + h * o.stride * np.arange(ns)
        if o.jac:
            if not grad:
                raise _lib.NbgError("NBG_ERR_ARG: CartesianOutput(jac=True) needs grad=True")
            M = 7 * s.n
            jcm = np.zeros((ns, s.nsys, M, M))   # per system column-major (Julia layout)
            check(_lib.lib().nbg_integrate_sampled_jac(plan, C.c_double(h), C.c_int64(o.nstep), C.c_int64(o.stride), C.c_int32(1), ptr(o.x), ptr(o.v), ptr(jcm)))
            o.jac_step = np.ascontiguousarray(jcm.transpose(0, 1, 3, 2))
        else:
            check(_lib.lib().nbg_integrate_sampled(plan, C.c_double(h), C.c_int64(o.nstep), C.c_int64(o.stride), C.c_int32(1 if grad else 0),
                                                   ptr(o.x), ptr(o.v)))
        s._download(plan, grad)
        self._timings(plan)

    # elements at sampled steps: the (intr)(s, o::CartesianOutput) loop with get_orbital_elements applied to each saved state
    def _sampled_elements(self, s, o, grad):
        L = _lib.lib()
        t0 = float(s.t[0])
        h = self.h * check_step(t0, self.tmax)
        ns = (o.nstep + o.stride - 1) // o.stride
        plan = self._p(s)
        s._upload(plan, grad)
        o.elements = np.zeros((ns, s.nsys, s.n, 11))
        o.t = t0 + h * o.stride * np.arange(ns)
        done = 0
        for k in range(ns):
            check(L.nbg_orbital_elements(plan, ptr(o.eps), ptr(o.elements[k])))            # the state BEFORE step k * stride
            nstep = min(o.stride, o.nstep - done)
            check(L.nbg_integrate_resident(plan, C.c_double(h), C.c_int64(nstep), C.c_double(0.0), C.c_int32(1 if grad else 0), C.c_int32(1),
                                           C.c_double(t0 + h * (done + nstep))))
            done += nstep
        s._download(plan, grad)
        self._timings(plan)

    def orbital_elements(self, s, eps=None):
        """get_orbital_elements(s, ic) of a State as it is now: [b, i, 11] (uploads the state, converts on the device)."""
        plan = self._p(s)
        s._upload(plan, False)
        out = np.zeros((s.nsys, s.n, 11))
        e = None if eps is None else np.asfortranarray(np.asarray(eps, dtype=np.float64))
        check(_lib.lib().nbg_orbital_elements(plan, ptr(e), ptr(out)))
        return out

    # (intr)(s, tt; grad) — Transits.jl:140-180
    def _transits(self, s, tt, grad):
        L = _lib.lib()
        plan = self._p(s)
        n, B, ntt, M = s.n, s.nsys, tt.ntt, 7 * s.n
        ntt_body = np.full(n, ntt, dtype=np.int32)
        mode = 1 if tt.ncomp == 3 else 0
        want_dtde = grad and s.jac_init.size
        Cn = tt.ncomp
        shp_t = (B, n, ntt) if Cn == 1 else (B, n, ntt, 3)
        shp_d = (B, n, ntt, n, 7) if Cn == 1 else (B, n, ntt, n, 7, 3)
        t_raw = np.zeros(shp_t)
        d_raw = np.zeros(shp_d) if grad else None
        e_raw = np.zeros(shp_d) if want_dtde else None
        one_shot = (not self.keep_dense and s._fresh and s._resident_plan is None and not s.xerror.any() and not s.verror.any()
                    and not s.dqdt.any() and not s.jac_error.any())
        if one_shot:
            # a fresh State(ic): the one-shot entry point knows the host destinations before it starts, so the library streams every
            # chunk's transit rows to them while the next chunk computes and never holds the full arrays on the device
            pr = np.asfortranarray(s.pair.astype(np.uint8))
            m_c = np.ascontiguousarray(s.m, dtype=np.float64)
            ji = np.ascontiguousarray(s.jac_init.transpose(0, 2, 1)) if want_dtde else None
            js = np.empty((B, M, M)) if grad else None
            je = np.empty((B, M, M)) if grad else None
            check(L.nbg_transit_timing(plan, ptr(s.x), ptr(s.v), ptr(m_c), ptr(pr) if pr.any() else None, C.c_double(float(s.t[0])),
                                       C.c_double(self.h), C.c_double(self.tmax), C.c_int32(tt.ti), ptr(ntt_body), C.c_int32(mode),
                                       C.c_int32(1 if grad else 0), ptr(ji), ptr(t_raw), ptr(tt.count), ptr(d_raw), ptr(e_raw), ptr(s.x), ptr(s.v),
                                       ptr(s.xerror), ptr(s.verror), ptr(js), ptr(je), ptr(s.dqdt) if grad else None, ptr(s.t), ptr(s.status)))
            if grad:
                s.jac_step[...] = js.transpose(0, 2, 1)
                s.jac_error[...] = je.transpose(0, 2, 1)
            s._fresh = False
        else:
            resident = s._upload(plan, grad)
            ji = None
            if want_dtde and not resident:   # resident: jac_init computed on the device is used (jac_init = NULL)
                ji = np.ascontiguousarray(s.jac_init.transpose(0, 2, 1))
            check(L.nbg_transit_timing_resident(plan, C.c_double(self.h), C.c_double(self.tmax), C.c_int32(tt.ti), ptr(ntt_body), C.c_int32(mode),
                                                C.c_int32(1 if grad else 0), ptr(ji)))
            check(L.nbg_transit_fetch(plan, ptr(t_raw), ptr(tt.count), ptr(d_raw), ptr(e_raw)))
            s._download(plan, grad)
        if Cn == 1:
            tt.tt[...] = t_raw
            if grad:
                tt.dtdq0[...] = d_raw.transpose(0, 1, 2, 4, 3)
                if e_raw is not None:
                    tt.dtdelements[...] = e_raw.transpose(0, 1, 2, 4, 3)
        else:
            tt.ttbv[...] = t_raw.transpose(0, 3, 1, 2)
            if grad:
                tt.dtbvdq0[...] = d_raw.transpose(0, 5, 1, 2, 4, 3)
                if e_raw is not None:
                    tt.dtbvdelements[...] = e_raw.transpose(0, 5, 1, 2, 4, 3)
        self._timings(plan)
        self._last_plan, self._last_tt, self._last_dense = plan, tt, not one_shot
        st = s.status
        if (st & _lib.ST_EVENT_OVERFLOW).any():   # cannot happen since v2 (chunks are re-run); never let incomplete results pass silently
            raise _lib.NbgError("transit queue overflow: results are incomplete")

    def chi2(self, t_obs, sigma):
        """Transit-time likelihood reduced on the device from the dense arrays of the LAST (s, tt::TransitTiming) call of this Integrator
        (nbg_transit_chi2; needs Integrator(..., keep_dense=True) or a non-fresh State): chi2[b] = sum ((tt - t_obs)/sigma)^2 and its
        gradients w.r.t. the initial Cartesian state [b,q,p] and the elements [b,q,p].
        t_obs, sigma: [n, ntt] (shared by the batch) or [B, n, ntt]; sigma <= 0 or NaN t_obs = no observation in that slot."""
        if self._last_plan is None or not self._last_dense:
            raise _lib.NbgError("chi2() reduces the dense device arrays of the last transit call: use Integrator(..., keep_dense=True), or chi2_fused()")
        plan, tt = self._last_plan, self._last_tt
        B, n, M = tt.nsys, tt.n, 7 * tt.n
        t_obs = np.ascontiguousarray(t_obs, dtype=np.float64); sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        per_system = 1 if t_obs.ndim == 3 else 0
        chi2, gq, ge = np.zeros(B), np.zeros((B, n, 7)), np.zeros((B, n, 7))
        check(_lib.lib().nbg_transit_chi2(plan, ptr(t_obs), ptr(sigma), C.c_int32(per_system), ptr(chi2), ptr(gq), ptr(ge)))
        return chi2, gq.transpose(0, 2, 1).copy(), ge.transpose(0, 2, 1).copy()


def _chi2_fused(self, s, tt, t_obs, sigma, wrt="q0", grad=True, want_tt=False):
    """The transit-timing run with the likelihood fused into the Jacobian kernel (nbg_transit_chi2_fused): no dtdq0 / dtdelements array
    exists anywhere, 1 + 7n doubles per system come back.  wrt = "q0": gradient w.r.t. the initial Cartesian state; "elements": w.r.t.
    the orbital elements -- needs State(ic, on_device=True) (jac_step is seeded with the device-computed jac_init).
    Returns chi2[b], grad[b,q,p] (None with grad=False); tt.count is filled, tt.tt too if want_tt."""
    L = _lib.lib()
    plan = self._p(s)
    n, B, ntt = s.n, s.nsys, tt.ntt
    if tt.ncomp != 1:
        raise _lib.NbgError("NBG_ERR_UNSUPPORTED: chi^2 is defined for TransitTiming")
    seed = wrt == "elements"
    resident = s._upload(plan, grad)
    if seed and not resident:
        raise _lib.NbgError("wrt='elements' needs a fresh State(ic, on_device=True) (the device-computed jac_init seeds jac_step)")
    ntt_body = np.full(n, ntt, dtype=np.int32)
    t_obs = np.ascontiguousarray(t_obs, dtype=np.float64); sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    per_system = 1 if t_obs.ndim == 3 else 0
    chi2 = np.zeros(B)
    g = np.zeros((B, n, 7)) if grad else None
    t_raw = np.zeros((B, n, ntt)) if want_tt else None
    check(L.nbg_transit_chi2_fused(plan, C.c_double(self.h), C.c_double(self.tmax), C.c_int32(tt.ti), ptr(ntt_body), ptr(t_obs), ptr(sigma),
                                   C.c_int32(per_system), C.c_int32(1 if seed else 0), C.c_int32(1 if grad else 0), ptr(chi2), ptr(g), ptr(tt.count),
                                   ptr(t_raw)))
    if want_tt:
        tt.tt[...] = t_raw
    s._download(plan, grad)
    self._timings(plan)
    self._last_plan, self._last_tt, self._last_dense = plan, tt, False
    return chi2, (g.transpose(0, 2, 1).copy() if grad else None)


Integrator.chi2_fused = _chi2_fused


def device_count():
    return int(_lib.lib().nbg_device_count())
