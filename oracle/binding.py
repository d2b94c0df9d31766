"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/libnbg_oracle*.so (the C++ restatement of the reference hot path,
see nbg_oracle.hpp).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(fast_native=False):
    """Compile the oracle (both builds).  fast_native rebuilds the timing build with -march=native."""
    args = ["make", "-C", _HERE, "-s"]
    if fast_native:
        subprocess.run(["rm", "-f", os.path.join(_HERE, "libnbg_oracle_fast.so")], check=False)
        args.append("FASTARCH=-march=native")
    subprocess.run(args, check=True)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def find_openblas():
    """(CDLL, dgemm symbol name, description) of an LP64 OpenBLAS that ships with the Python stack (scipy's), or None.
    The reference's mul! calls are OpenBLAS dgemm (Julia LinearAlgebra); the CPU-baseline timing build can use the same."""
    import glob
    try:
        import scipy
    except Exception:
        return None
    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    for path in sorted(glob.glob(os.path.join(root, "libscipy_openblas*.so"))):
        if "64_" in os.path.basename(path):
            continue   # ILP64 build: 64-bit integer arguments
        try:
            lib = C.CDLL(path)
            fn = getattr(lib, "scipy_dgemm_")
            try:
                lib.scipy_openblas_set_num_threads(C.c_int(1))   # the batch is threaded over systems
            except AttributeError:
                pass
            return lib, fn, "OpenBLAS dgemm (%s, 1 thread per call)" % os.path.basename(path)
        except (OSError, AttributeError):
            continue
    return None


class Oracle:
    def __init__(self, fast=False, blas=False):
        """fast: the -O3 timing build.  blas (timing build only): dense products through OpenBLAS dgemm, as the reference's mul! calls."""
        name = "libnbg_oracle_fast.so" if fast else "libnbg_oracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.nbgo_gnewt.restype = C.c_double
        self.GNEWT = self.lib.nbgo_gnewt()
        self.blas = None
        if blas:
            if not fast:
                raise ValueError("BLAS products are for the timing build only: the reference-semantics build keeps the k-ascending loops")
            found = find_openblas()
            if found is not None:
                self._blas_lib, fn, self.blas = found
                self.lib.nbgo_set_dgemm(C.cast(fn, C.c_void_p))
        elif fast and hasattr(self.lib, "nbgo_set_dgemm"):
            self.lib.nbgo_set_dgemm(None)

    # ---- IC layer ---------------------------------------------------------
    def init_nbody(self, elements, t0, eps=None):
        """elements: (n,7) rows = m,P,t0,ecosw,esinw,I,Omega.  Returns x(n,3), v(n,3), jac_init(M,M) [row,col]."""
        el = np.asfortranarray(np.asarray(elements, dtype=np.float64))
        n = el.shape[0]
        x = np.zeros((n, 3)); v = np.zeros((n, 3)); jac = np.zeros((7 * n, 7 * n))
        epsf = None if eps is None else np.asfortranarray(np.asarray(eps, dtype=np.float64))
        self.lib.nbgo_init_nbody(C.c_int(n), _ptr(el), C.c_double(t0), _ptr(epsf), _ptr(x), _ptr(v), _ptr(jac))
        return x, v, jac.T.copy()  # column-major -> [row, col]

    def orbital_elements(self, x, v, m, eps=None):
        """get_orbital_elements (src/outputs/elements.jl:108-137): returns (n, 11) rows (m, P, t0=0, ecosw, esinw, I, Omega, a, e, omega, tp)."""
        n = len(m)
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(n, 3)); v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(n, 3))
        m = np.ascontiguousarray(m, dtype=np.float64)
        epsf = None if eps is None else np.asfortranarray(np.asarray(eps, dtype=np.float64))
        out = np.zeros((n, 11))
        self.lib.nbgo_orbital_elements(C.c_int(n), _ptr(m), _ptr(epsf), _ptr(x), _ptr(v), _ptr(out))
        return out

    def ntt(self, tmax, periods):
        p = np.ascontiguousarray(periods, dtype=np.float64)
        return int(self.lib.nbgo_ntt(C.c_double(tmax), _ptr(p), C.c_int(len(p))))

    # ---- state helpers -------------------------------------------------------
    @staticmethod
    def new_state(x, v, m, t0=0.0):
        """State dict; x, v are (n,3) [body, dim]; jac_step (M,M) is stored column-major (jac[c, r])."""
        n = len(m)
        M = 7 * n
        return dict(n=n, x=np.array(x, dtype=np.float64).reshape(n, 3).copy(), v=np.array(v, dtype=np.float64).reshape(n, 3).copy(),
                    m=np.array(m, dtype=np.float64).copy(), xerr=np.zeros((n, 3)), verr=np.zeros((n, 3)),
                    jac_step_cm=np.eye(M), jac_err_cm=np.zeros((M, M)), dqdt=np.zeros(M), t=np.array([t0], dtype=np.float64), pair=None)

    def integrate(self, s, h, time=None, nsteps=None, grad=True):
        """(intr)(s,time) when time is given, else (intr)(s,N)."""
        n = s["n"]
        mode = 0 if time is not None else 1
        pair = None if s.get("pair") is None else np.asfortranarray(s["pair"].astype(np.uint8))
        self.lib.nbgo_integrate(C.c_int(n), _ptr(s["x"]), _ptr(s["v"]), _ptr(s["m"]), _ptr(pair), _ptr(s["xerr"]), _ptr(s["verr"]),
                                _ptr(s["jac_step_cm"]), _ptr(s["jac_err_cm"]), _ptr(s["dqdt"]), _ptr(s["t"]), C.c_double(h), C.c_int(mode),
                                C.c_double(0.0 if time is None else time), C.c_long(0 if nsteps is None else nsteps), C.c_int(1 if grad else 0))
        return s

    def transit_timing(self, s, h, tmax, ntt, ti=0, grad=True, jac_init=None, ntbv=1):
        """Returns dict(tt, count, dtdq0, dtdelements) in numpy index order tt[i,k] / ttbv[c,i,k];
        dtdq0[i,k,q,p] (or [c,i,k,q,p]) exactly as the reference's arrays (0-based)."""
        n = s["n"]
        M = 7 * n
        shp_tt = (ntt, n) if ntbv == 1 else (ntt, n, 3)
        shp_d = (n, 7, ntt, n) if ntbv == 1 else (n, 7, ntt, n, 3)
        tt = np.zeros(shp_tt); count = np.zeros(n, dtype=np.int64)
        dtdq0 = np.zeros(shp_d); dtde = np.zeros(shp_d)
        stats = np.zeros(4, dtype=np.int64)
        ji = None if jac_init is None else np.ascontiguousarray(np.asarray(jac_init).T)  # [row,col] -> column-major
        pair = None if s.get("pair") is None else np.asfortranarray(s["pair"].astype(np.uint8))
        self.lib.nbgo_transit_timing(C.c_int(n), _ptr(s["x"]), _ptr(s["v"]), _ptr(s["m"]), _ptr(pair), _ptr(ji), _ptr(s["t"]), C.c_double(h),
                                     C.c_double(tmax), C.c_int(ti), C.c_int(ntt), C.c_int(ntbv), C.c_int(1 if grad else 0), _ptr(tt),
                                     _ptr(count), _ptr(dtdq0), _ptr(dtde), _ptr(s["xerr"]), _ptr(s["verr"]), _ptr(s["jac_step_cm"]),
                                     _ptr(s["jac_err_cm"]), _ptr(s["dqdt"]), _ptr(stats))
        # column-major (i fastest) -> numpy [i,k,...]
        if ntbv == 1:
            out = dict(tt=tt.T.copy(), dtdq0=dtdq0.transpose(3, 2, 1, 0).copy(), dtdelements=dtde.transpose(3, 2, 1, 0).copy())
        else:
            out = dict(tt=tt.transpose(2, 1, 0).copy(), dtdq0=dtdq0.transpose(4, 3, 2, 1, 0).copy(), dtdelements=dtde.transpose(4, 3, 2, 1, 0).copy())
        out.update(count=count, newton_iters=int(stats[0]), kepler_calls=int(stats[1]), gamma_iters=int(stats[2]), itmax_transits=int(stats[3]))
        return out

    def batch_transit_timing(self, x, v, m, t0, h, tmax, ntt, ti=0, grad=True, jac_init_cm=None, nthreads=1, want_grad_arrays=True):
        """x, v: (B,n,3); m: (B,n); jac_init_cm: (B,M,M) column-major or None.  Returns dict like transit_timing with a leading batch axis
        (raw reference memory order: tt (B,ntt,n), dtdq0 (B,n,7,ntt,n))."""
        B, n = m.shape
        x = np.ascontiguousarray(x, dtype=np.float64).copy(); v = np.ascontiguousarray(v, dtype=np.float64).copy()
        m = np.ascontiguousarray(m, dtype=np.float64)
        tt = np.zeros((B, ntt, n)); count = np.zeros((B, n), dtype=np.int64)
        dtdq0 = np.zeros((B, n, 7, ntt, n)) if (grad and want_grad_arrays) else None
        dtde = np.zeros((B, n, 7, ntt, n)) if (grad and want_grad_arrays and jac_init_cm is not None) else None
        newton = C.c_long(0)
        itmax = np.zeros(B, dtype=np.int64)
        self.lib.nbgo_batch_transit_timing(C.c_long(B), C.c_int(n), _ptr(x), _ptr(v), _ptr(m), _ptr(jac_init_cm), C.c_double(t0), C.c_double(h),
                                           C.c_double(tmax), C.c_int(ti), C.c_int(ntt), C.c_int(1 if grad else 0), _ptr(tt), _ptr(count),
                                           _ptr(dtdq0), _ptr(dtde), C.c_int(nthreads), C.byref(newton), _ptr(itmax))
        return dict(tt=tt, count=count, dtdq0=dtdq0, dtdelements=dtde, x=x, v=v, newton_iters=newton.value, itmax_per_system=itmax)

    def batch_integrate(self, x, v, m, h, nsteps, grad=True, nthreads=1):
        B, n = m.shape
        M = 7 * n
        x = np.ascontiguousarray(x, dtype=np.float64).copy(); v = np.ascontiguousarray(v, dtype=np.float64).copy()
        m = np.ascontiguousarray(m, dtype=np.float64)
        xerr = np.zeros_like(x); verr = np.zeros_like(v)
        jac = np.ascontiguousarray(np.broadcast_to(np.eye(M), (B, M, M))).copy() if grad else None
        jerr = np.zeros((B, M, M)) if grad else None
        dqdt = np.zeros((B, M)) if grad else None
        self.lib.nbgo_batch_integrate(C.c_long(B), C.c_int(n), _ptr(x), _ptr(v), _ptr(m), _ptr(xerr), _ptr(verr), _ptr(jac), _ptr(jerr), _ptr(dqdt),
                                      C.c_double(h), C.c_long(nsteps), C.c_int(1 if grad else 0), C.c_int(nthreads))
        return dict(x=x, v=v, xerr=xerr, verr=verr, jac_step_cm=jac, jac_err_cm=jerr, dqdt=dqdt)

    # ---- unit pieces -----------------------------------------------------------
    def kepler_driftij(self, x, v, m, i, j, h, drift_first, grad=True):
        n = len(m)
        x = np.array(x, dtype=np.float64).reshape(n, 3).copy(); v = np.array(v, dtype=np.float64).reshape(n, 3).copy()
        m = np.ascontiguousarray(m, dtype=np.float64)
        jac = np.zeros((14, 14)); dq = np.zeros(14)
        self.lib.nbgo_kepler_driftij(C.c_int(n), _ptr(x), _ptr(v), _ptr(m), C.c_int(i), C.c_int(j), C.c_double(h), C.c_int(1 if drift_first else 0),
                                     C.c_int(1 if grad else 0), _ptr(jac), _ptr(dq))
        return x, v, jac.T.copy(), dq

    def kick_piece(self, which, x, v, m, pair, h, grad=True):
        """which: 'phisalpha' | 'phic' | 'kickfast'.  Returns x, v, jac[row,col] (identity removed), dq."""
        wid = dict(phisalpha=2, phic=3, kickfast=4)[which]
        n = len(m)
        M = 7 * n
        x = np.array(x, dtype=np.float64).reshape(n, 3).copy(); v = np.array(v, dtype=np.float64).reshape(n, 3).copy()
        m = np.ascontiguousarray(m, dtype=np.float64)
        pr = np.asfortranarray(np.asarray(pair).astype(np.uint8))
        jac = np.zeros((M, M)); dq = np.zeros(M)
        self.lib.nbgo_kick_piece(C.c_int(wid), C.c_int(n), _ptr(x), _ptr(v), _ptr(m), _ptr(pr), C.c_double(h), C.c_int(1 if grad else 0), _ptr(jac),
                                 _ptr(dq))
        return x, v, jac.T.copy(), dq

    # ---- __float128 finite differences ----------------------------------------------
    def fd_map(self, which, x, v, m, h, pair=None, nsteps=1, i=0, j=1, drift_first=True, dlnq=1e-20, want_jac=True, want_dqdt=True):
        mid = dict(ahl21=0, kepler_driftij=1, phisalpha=2, phic=3, kickfast=4)[which]
        n = len(m)
        M = 7 * n
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(n, 3)); v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(n, 3))
        m = np.ascontiguousarray(m, dtype=np.float64)
        pr = None if pair is None else np.asfortranarray(np.asarray(pair).astype(np.uint8))
        jac = np.zeros((M, M)) if want_jac else None
        dq = np.zeros(M) if want_dqdt else None
        self.lib.nbgoq_fd_map(C.c_int(mid), C.c_int(n), _ptr(x), _ptr(v), _ptr(m), _ptr(pr), C.c_double(h), C.c_long(nsteps), C.c_int(i), C.c_int(j),
                              C.c_int(1 if drift_first else 0), C.c_double(dlnq), _ptr(jac), _ptr(dq))
        return (None if jac is None else jac.T.copy()), dq

    def fd_transit_elements(self, elements, t0, h, tmax, ntt, ti=0, dq0=1e-10, ntbv=1):
        el = np.asfortranarray(np.asarray(elements, dtype=np.float64))
        n = el.shape[0]
        shp = (n, 7, ntt, n) if ntbv == 1 else (n, 7, ntt, n, 3)
        out = np.zeros(shp); count = np.zeros(n, dtype=np.int64)
        self.lib.nbgoq_fd_transit_elements(C.c_int(n), _ptr(el), C.c_double(t0), C.c_double(h), C.c_double(tmax), C.c_int(ti), C.c_int(ntt),
                                           C.c_int(ntbv), C.c_double(dq0), _ptr(out), _ptr(count))
        out = out.transpose(3, 2, 1, 0).copy() if ntbv == 1 else out.transpose(4, 3, 2, 1, 0).copy()
        return out, count

    def quad_transit_times(self, elements, t0, h, tmax, ntt, ti=0):
        el = np.asfortranarray(np.asarray(elements, dtype=np.float64))
        n = el.shape[0]
        tt = np.zeros((ntt, n)); count = np.zeros(n, dtype=np.int64)
        self.lib.nbgoq_transit_times(C.c_int(n), _ptr(el), C.c_double(t0), C.c_double(h), C.c_double(tmax), C.c_int(ti), C.c_int(ntt), _ptr(tt),
                                     _ptr(count))
        return tt.T.copy(), count

    def quad_transit_timing_grad(self, x, v, m, jac_init, t0, h, tmax, ntt, ti=0):
        """The full gradient path in __float128 from double inputs (nbgoq_transit_timing_grad); same return layout as transit_timing plus
        the final x, v, jac_step_cm, all rounded to double."""
        n = len(m)
        M = 7 * n
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(n, 3)); v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(n, 3))
        m = np.ascontiguousarray(m, dtype=np.float64)
        ji = np.ascontiguousarray(np.asarray(jac_init).T)
        tt = np.zeros((ntt, n)); count = np.zeros(n, dtype=np.int64)
        dtdq0 = np.zeros((n, 7, ntt, n)); dtde = np.zeros((n, 7, ntt, n))
        xo = np.zeros((n, 3)); vo = np.zeros((n, 3)); js = np.zeros((M, M))
        self.lib.nbgoq_transit_timing_grad(C.c_int(n), _ptr(x), _ptr(v), _ptr(m), _ptr(ji), C.c_double(t0), C.c_double(h), C.c_double(tmax), C.c_int(ti),
                                           C.c_int(ntt), _ptr(tt), _ptr(count), _ptr(dtdq0), _ptr(dtde), _ptr(xo), _ptr(vo), _ptr(js))
        return dict(tt=tt.T.copy(), count=count, dtdq0=dtdq0.transpose(3, 2, 1, 0).copy(), dtdelements=dtde.transpose(3, 2, 1, 0).copy(), x=xo, v=vo,
                    jac_step_cm=js)
