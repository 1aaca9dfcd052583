"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/liboracle.so (the plain-C
restatement in oracle/bezier_oracle.c).  Used by tests and by bench.py's CPU
baseline; never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

from oracle import bezier_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "bezier_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    _lib = ctypes.CDLL(LIB)
    dp = ctypes.c_void_p
    _lib.oracle_temporal_separation.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_double, dp, dp, ctypes.c_longlong,
                                                ctypes.c_longlong, dp, ctypes.c_int]
    _lib.oracle_speed.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, dp, dp,
                                  ctypes.c_int]
    return _lib


def max_threads():
    return int(load().oracle_max_threads())


def temporal_separation(y, nVeh, dim, maxSep, E, pair_begin=0, npairs=None, nthreads=0, out=None):
    lib = load()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[1] - 1
    W = np.ascontiguousarray(O.prod_weights(n))
    T = np.ascontiguousarray(O.elev_matrix(2 * n, E))
    P = nVeh * (nVeh - 1) // 2
    if npairs is None:
        npairs = P - pair_begin
    L = 2 * n + E + 1
    if out is None:
        out = np.empty(npairs * L)
    rc = lib.oracle_temporal_separation(y.ctypes.data, nVeh, dim, n, E, float(maxSep), W.ctypes.data,
                                        T.ctypes.data, pair_begin, npairs, out.ctypes.data, nthreads)
    assert rc == 0
    return out


def speed(y, nVeh, dim, tf, E, alpha, beta, nthreads=0):
    lib = load()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.shape[1] - 1
    W = np.ascontiguousarray(O.prod_weights(n))
    T = np.ascontiguousarray(O.elev_matrix(2 * n, E))
    E1 = np.ascontiguousarray(O.elev_matrix(n - 1, 1))
    out = np.empty(nVeh * (2 * n + E + 1))
    rc = lib.oracle_speed(y.ctypes.data, nVeh, dim, n, E, float(tf), float(alpha), float(beta),
                          W.ctypes.data, T.ctypes.data, E1.ctypes.data, out.ctypes.data, nthreads)
    assert rc == 0
    return out
