"""ctypes binding of csrc/libnbgrad_b200.so (include/nbgrad.h).  There is no fallback: if the shared library is
missing or cannot be loaded this module raises, and every compute call fails loudly without a CUDA device."""
import ctypes as C
import os

from .build import LIB

NBG_OK = 0
ERRORS = {-1: "NBG_ERR_ARG", -2: "NBG_ERR_NO_DEVICE", -3: "NBG_ERR_CUDA", -4: "NBG_ERR_UNSUPPORTED", -5: "NBG_ERR_NOMEM"}
ST_NONFINITE, ST_TRANSIT_ITMAX, ST_EVENT_OVERFLOW, ST_NTT_OVERFLOW = 1, 2, 4, 8

# every symbol include/nbgrad.h declares
SYMBOLS = ["nbg_version", "nbg_last_error", "nbg_device_count", "nbg_plan_create", "nbg_plan_destroy", "nbg_set_pair", "nbg_set_state", "nbg_set_state_elements", "nbg_get_jac_init", "nbg_get_state",
           "nbg_integrate_resident", "nbg_integrate_sampled", "nbg_integrate_sampled_jac", "nbg_integrate", "nbg_transit_timing_resident", "nbg_transit_fetch", "nbg_transit_chi2", "nbg_transit_timing",
           "nbg_counters", "nbg_counters_reset", "nbg_last_timings", "nbg_cuda_stream", "nbg_fp64_peak", "nbg_build_flags", "nbg_plan_create_multi", "nbg_plan_devices",
           "nbg_state_generation", "nbg_transit_chi2_fused", "nbg_chunk_retries", "nbg_orbital_elements", "nbg_source_hash"]


class NbgError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise NbgError("libnbgrad_b200.so is not built (%s): run `python __graft_entry__.py build`; there is no CPU fallback" % LIB)
        L = C.CDLL(LIB)
        for s in SYMBOLS:
            getattr(L, s)
        L.nbg_last_error.restype = C.c_char_p
        L.nbg_source_hash.restype = C.c_char_p
        from .build import source_hash
        if L.nbg_source_hash().decode() != source_hash() and os.environ.get("NBGRAD_ALLOW_STALE") != "1":
            raise NbgError("libnbgrad_b200.so was built from other sources (%s, sources are %s): run `python -m nbgrad.build --if-stale` "
                           "(or set NBGRAD_ALLOW_STALE=1)" % (L.nbg_source_hash().decode(), source_hash()))
        L.nbg_cuda_stream.restype = C.c_int64
        L.nbg_chunk_retries.restype = C.c_int64
        L.nbg_state_generation.restype = C.c_int64
        _lib = L
    return _lib


def check(rc):
    if rc != NBG_OK:
        raise NbgError("%s: %s" % (ERRORS.get(rc, rc), lib().nbg_last_error().decode()))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
