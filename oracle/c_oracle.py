"""TEST INFRASTRUCTURE ONLY -- ctypes loader of oracle/libmde_oracle.so (the C restatement)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load(build=True):
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmde_oracle.so")
        if not os.path.exists(path) and build:
            subprocess.check_call(["make", "-s", "-C", _HERE])
        lib = C.CDLL(path)
        lib.mde_oracle_eval.restype = C.c_int
        lib.mde_oracle_eval.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.POINTER(C.c_double), C.c_void_p]
        _LIB = lib
    return _LIB


def average_distortion(X, edges, spec, want_grad=True, p_total=None):
    """Same contract as oracle.mde_oracle.average_distortion (float64 result)."""
    lib = load()
    X = np.ascontiguousarray(X, dtype=np.float32)
    e = np.ascontiguousarray(edges, dtype=np.int64)
    n, m = X.shape
    p = e.shape[0]
    par0 = np.ascontiguousarray(np.broadcast_to(np.asarray(spec.par0, dtype=np.float32), (p,)))
    par1 = None if spec.par1 is None else np.ascontiguousarray(spec.par1, dtype=np.float32)
    att = np.array(spec.att, dtype=np.float64)
    rep = np.array(spec.rep, dtype=np.float64)
    grad = np.empty((n, m), dtype=np.float64) if want_grad else None
    val = C.c_double()
    lib.mde_oracle_eval(n, m, p, e.ctypes.data, spec.fn_att, spec.fn_rep, att.ctypes.data, rep.ctypes.data,
                        int(spec.push_pull), par0.ctypes.data, None if par1 is None else par1.ctypes.data,
                        X.ctypes.data, int(p if p_total is None else p_total), C.byref(val),
                        None if grad is None else grad.ctypes.data)
    return val.value, grad
