"""ctypes bridge to the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libref_matrix.so")


class OracleOpts(C.Structure):
    _fields_ = [("fixed_std_dev", C.c_double), ("free_std_dev", C.c_double), ("iteration_threshold", C.c_double),
                ("semi_major", C.c_double), ("inv_flattening", C.c_double), ("confidence_interval", C.c_double),
                ("max_iterations", C.c_uint32), ("scale_normals_to_unity", C.c_int32), ("use_ref", C.c_int32),
                ("threads", C.c_int32)]


class OracleResult(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("converged", C.c_int32), ("max_corr", C.c_double),
                ("max_corr_row", C.c_uint32), ("chi_squared", C.c_double), ("sigma_zero", C.c_double),
                ("dof", C.c_int64), ("measurement_params", C.c_uint32), ("unknown_params", C.c_uint32),
                ("outliers", C.c_uint32), ("global_pelzer", C.c_double), ("critical_value", C.c_double),
                ("used_ref", C.c_int32), ("seconds_prepare", C.c_double), ("seconds_solve", C.c_double),
                ("seconds_inverse", C.c_double)]


_lib = None


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_load_ref.argtypes = [C.c_char_p]
        L.oracle_adjust_simultaneous.argtypes = [C.POINTER(OracleOpts), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.POINTER(OracleResult)]
        L.oracle_adjust_phased.argtypes = [C.POINTER(OracleOpts), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_uint32),
                                           C.c_void_p, C.c_void_p, C.POINTER(OracleResult)]
        L.oracle_spd_inverse.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.oracle_geo_to_cart.argtypes = [C.c_double] * 5 + [C.c_void_p]
        L.oracle_cart_to_geo.argtypes = [C.c_double] * 5 + [C.c_void_p]
        if os.path.exists(REF_PATH):
            L.oracle_load_ref(REF_PATH.encode())
        _lib = L
    return _lib


def ref_loaded():
    return bool(lib().oracle_ref_loaded())


def default_opts(**kw):
    o = OracleOpts()
    lib().oracle_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def adjust_simultaneous(stn, msr, opts=None, want_normals=False, want_vcv=False):
    """Run the dense oracle.  ``stn`` / ``msr`` are mutated like the reference mutates its records."""
    L = lib()
    o = opts or default_opts()
    n = 3 * len(stn)
    est = np.zeros(n)
    rhs = np.zeros(n)
    corr = np.zeros(n)
    normals = np.zeros((n, n), order="F") if want_normals else None
    vcv = np.zeros((n, n), order="F") if want_vcv else None
    res = OracleResult()
    rc = L.oracle_adjust_simultaneous(C.byref(o), _ptr(stn), len(stn), _ptr(msr), len(msr), _ptr(est), _ptr(normals),
                                      _ptr(rhs), _ptr(corr), _ptr(vcv), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"oracle failed ({rc}): {L.oracle_last_error().decode()}")
    return dict(est=est.reshape(-1, 3), rhs=rhs, first_corr=corr, normals=normals, vcv=vcv, res=res)


def adjust_phased(stn, msr, inner_station_lists, opts=None, want_block=-1):
    """The reference's phased adjustment (forward pass, reverse pass, combination) over a chain of blocks given by
    their inner-station lists.  Returns estimates, every station's rigorous 3x3 variance block and, for block
    ``want_block``, its station list (inner then junction) and dense variance matrix."""
    L = lib()
    o = opts or default_opts()
    off = np.zeros(len(inner_station_lists) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(b) for b in inner_station_lists])
    isl = np.concatenate([np.asarray(b, dtype=np.uint32) for b in inner_station_lists])
    est = np.zeros((len(stn), 3))
    vcv = np.zeros((len(stn), 3, 3))
    res = OracleResult()
    bn = C.c_uint32(0)
    bst = bv = None
    if want_block >= 0:
        # capacity: a block never holds more stations than the network
        cap = len(stn)
        bst = np.zeros(cap, np.uint32)
        # size the dense matrix from the block's own station count: run is cheap to bound by inner + all later stations
        nmax = min(cap, len(inner_station_lists[want_block]) + cap)
        bv = np.zeros((3 * nmax) * (3 * nmax)) if nmax <= 4000 else None
        if bv is None:
            raise ValueError("block too large for the dense block-variance output")
    rc = L.oracle_adjust_phased(C.byref(o), _ptr(stn), len(stn), _ptr(msr), len(msr), len(inner_station_lists), _ptr(off),
                                _ptr(isl), _ptr(est), _ptr(vcv), want_block, C.byref(bn), _ptr(bst), _ptr(bv), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"phased oracle failed ({rc}): {L.oracle_last_error().decode()}")
    out = dict(est=est, vcv=vcv, res=res)
    if want_block >= 0:
        n = bn.value
        out["block_stations"] = bst[:n].copy()
        out["block_vcv"] = bv[:(3 * n) * (3 * n)].reshape(3 * n, 3 * n).copy()
    return out


def spd_inverse(a, use_ref=True):
    a = np.array(a, dtype=np.float64, order="F")
    rc = lib().oracle_spd_inverse(_ptr(a), a.shape[0], 1 if use_ref else 0)
    if rc != 0:
        raise np.linalg.LinAlgError("Matrix inversion failed, the matrix is singular.")
    return a


def geo_to_cart(lat, lon, h, a=6378137.0, invf=298.257222101):
    out = np.zeros(3)
    lib().oracle_geo_to_cart(lat, lon, h, a, invf, _ptr(out))
    return out


def cart_to_geo(x, y, z, a=6378137.0, invf=298.257222101):
    out = np.zeros(3)
    lib().oracle_cart_to_geo(x, y, z, a, invf, _ptr(out))
    return out
