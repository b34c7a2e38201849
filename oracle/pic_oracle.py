"""oracle/pic_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/_build/libpic_oracle.so (the C restatement in
oracle/pic_oracle.c; built by oracle/Makefile or __graft_entry__.build()).
PARITY UNPINNED for the PIC half -- see the header of pic_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpic_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pic_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_locate.argtypes = [_dp, C.c_int64, C.c_double, C.c_int, _ip, _dp]
        _lib.orc_deposit.argtypes = [_dp, _dp, C.c_double, C.c_int64, C.c_double, C.c_int,
                                     C.c_int, C.c_int, _dp]
        _lib.orc_poisson.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp, _dp]
        _lib.orc_poisson.restype = C.c_int
        _lib.orc_gather.argtypes = [_dp, _dp, C.c_int64, C.c_double, C.c_int, C.c_int, _dp]
        _lib.orc_stiffness.argtypes = [C.c_int, C.c_double, _dp]
        _lib.orc_integrate.argtypes = [_dp, _dp, _dp, C.c_double, C.c_int64, C.c_int, C.c_double,
                                       C.c_double, C.c_double, C.c_int, C.c_int,
                                       _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int]
        _lib.orc_integrate.restype = C.c_int
        _lib.orc_step.argtypes = [_dp, _dp, _dp, C.c_double, C.c_int64, C.c_int, C.c_double,
                                  C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]
        _lib.orc_step.restype = C.c_int
        _lib.orc_check_division.argtypes = [C.c_double, C.c_int, C.c_int64, C.c_uint64]
        _lib.orc_check_division.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def locate(x, L, nh):
    x = _f64(x)
    cell = np.empty(x.shape[0], dtype=np.int32)
    xi = np.empty(x.shape[0], dtype=np.float64)
    lib().orc_locate(_p(x), x.shape[0], L, nh, cell.ctypes.data_as(_ip), _p(xi))
    return cell, xi


def _weights(w, n):
    """-> (array or None, uniform value)"""
    if np.isscalar(w):
        return None, float(w)
    w = _f64(w).reshape(-1)
    assert w.shape[0] == n
    return w, 0.0


def deposit(x, w, L, nh, extended=True, nthreads=1):
    x = _f64(x)
    wa, wu = _weights(w, x.shape[0])
    rhs = np.empty(nh, dtype=np.float64)
    lib().orc_deposit(_p(x), _p(wa), wu, x.shape[0], L, nh, int(extended), nthreads, _p(rhs))
    return rhs


def poisson(rhs, nh, L, chi):
    rhs = _f64(rhs)
    phi = np.empty(nh, dtype=np.float64)
    W = C.c_double(0.0)
    rc = lib().orc_poisson(_p(rhs), nh, L, chi, _p(phi), C.byref(W))
    if rc:
        raise RuntimeError(f"orc_poisson failed ({rc})")
    return phi, W.value


def stiffness(nh, L):
    K = np.empty((nh, nh), dtype=np.float64)
    lib().orc_stiffness(nh, L, _p(K))
    return K


def gather(phi, x, L, nh, nthreads=1):
    x = _f64(x); phi = _f64(phi)
    a = np.empty_like(x)
    lib().orc_gather(_p(phi), _p(x), x.shape[0], L, nh, nthreads, _p(a))
    return a


def integrate(x0, v0, w, L, nh, chi, dt, nt, ns, extended=True, nthreads=1, outputs="XVAPWKM"):
    """One parameter sample.  Returns dict with X,V,A (ns,np), Phi (ns,nh), W,K,M (ns,)."""
    x0 = _f64(x0); v0 = _f64(v0)
    n = x0.shape[0]
    wa, wu = _weights(w, n)
    out = {}
    if "X" in outputs: out["X"] = np.zeros((ns, n))
    if "V" in outputs: out["V"] = np.zeros((ns, n))
    if "A" in outputs: out["A"] = np.zeros((ns, n))
    if "P" in outputs: out["Phi"] = np.zeros((ns, nh))
    out["W"] = np.zeros(ns); out["K"] = np.zeros(ns); out["M"] = np.zeros(ns)
    rc = lib().orc_integrate(_p(x0), _p(v0), _p(wa), wu, n, nh, L, chi, dt, nt, ns,
                             _p(out.get("X")), _p(out.get("V")), _p(out.get("A")),
                             _p(out.get("Phi")), _p(out["W"]), _p(out["K"]), _p(out["M"]),
                             int(extended), nthreads)
    if rc:
        raise RuntimeError(f"orc_integrate failed ({rc})")
    return out


def step(x, v, w, L, nh, chi, dt, extended=True, nthreads=1):
    """One drift-kick-drift step; returns (x1, v1, rhs, phi, a)."""
    x = _f64(x).copy(); v = _f64(v).copy()
    n = x.shape[0]
    wa, wu = _weights(w, n)
    rhs = np.empty(nh); phi = np.empty(nh); a = np.empty(n)
    rc = lib().orc_step(_p(x), _p(v), _p(wa), wu, n, nh, L, chi, dt, int(extended), nthreads,
                        _p(rhs), _p(phi), _p(a))
    if rc:
        raise RuntimeError(f"orc_step failed ({rc})")
    return x, v, rhs, phi, a


def check_division(L, nh, n=1_000_000, seed=1):
    return int(lib().orc_check_division(L, nh, n, seed))
