"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/clp_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package clp_b200 never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libclp_oracle.so")
_lib = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_ubyte_p = ctypes.POINTER(ctypes.c_ubyte)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, c_double_p,
                                 c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_set_option.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.orc_set_status.argtypes = [ctypes.c_void_p, c_ubyte_p]
        L.orc_dual.argtypes = [ctypes.c_void_p]
        L.orc_objective_value.restype = ctypes.c_double
        L.orc_objective_value.argtypes = [ctypes.c_void_p]
        L.orc_number_iterations.argtypes = [ctypes.c_void_p]
        L.orc_number_refactorizations.argtypes = [ctypes.c_void_p]
        L.orc_seconds_in_loop.restype = ctypes.c_double
        L.orc_seconds_in_loop.argtypes = [ctypes.c_void_p]
        L.orc_timed_seconds.restype = ctypes.c_double
        L.orc_timed_seconds.argtypes = [ctypes.c_void_p]
        L.orc_timed_iterations.argtypes = [ctypes.c_void_p]
        for f in ("orc_get_column_solution", "orc_get_row_activity", "orc_get_reduced_cost",
                  "orc_get_row_price"):
            getattr(L, f).argtypes = [ctypes.c_void_p, c_double_p]
        L.orc_get_status.argtypes = [ctypes.c_void_p, c_ubyte_p]
        L.orc_factorize.argtypes = [ctypes.c_void_p, c_int_p, c_int_p]
        L.orc_ftran.argtypes = [ctypes.c_void_p, c_double_p]
        L.orc_btran.argtypes = [ctypes.c_void_p, c_double_p]
        L.orc_replace_column.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.orc_transpose_times.argtypes = [ctypes.c_void_p, ctypes.c_double, c_double_p, c_double_p]
        L.orc_times.argtypes = [ctypes.c_void_p, ctypes.c_double, c_double_p, c_double_p]
        L.orc_dual_column.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, c_ubyte_p,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                      c_double_p, c_ubyte_p]
        L.orc_dual_column_bucketed.argtypes = L.orc_dual_column.argtypes
        L.orc_dse_update.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p,
                                     ctypes.c_int, ctypes.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _up(a):
    return a.ctypes.data_as(c_ubyte_p)


class OracleSimplex:
    """CPU restatement of ClpSimplex restricted to the dual path."""

    def __init__(self, lp):
        L = lib()
        self.lp = lp
        self.m, self.n = lp.m, lp.n
        self._keep = [np.ascontiguousarray(a, dtype=t) for a, t in (
            (lp.col_start, np.int32), (lp.row_index, np.int32), (lp.element, np.float64),
            (lp.col_lower, np.float64), (lp.col_upper, np.float64), (lp.objective, np.float64),
            (lp.row_lower, np.float64), (lp.row_upper, np.float64))]
        k = self._keep
        self.h = L.orc_create(lp.m, lp.n, _ip(k[0]), _ip(k[1]), _dp(k[2]), _dp(k[3]), _dp(k[4]),
                              _dp(k[5]), _dp(k[6]), _dp(k[7]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def set_option(self, key, value):
        lib().orc_set_option(self.h, key.encode(), float(value))

    def dual(self):
        return lib().orc_dual(self.h)

    @property
    def objective_value(self):
        return lib().orc_objective_value(self.h)

    @property
    def iterations(self):
        return lib().orc_number_iterations(self.h)

    @property
    def refactorizations(self):
        return lib().orc_number_refactorizations(self.h)

    @property
    def seconds(self):
        return lib().orc_seconds_in_loop(self.h)

    def timed_window(self):
        return lib().orc_timed_seconds(self.h), lib().orc_timed_iterations(self.h)

    def set_status(self, status):
        st = np.ascontiguousarray(status, dtype=np.uint8)
        lib().orc_set_status(self.h, _up(st))

    def _get(self, name, size):
        out = np.zeros(size)
        getattr(lib(), name)(self.h, _dp(out))
        return out

    def column_solution(self):
        return self._get("orc_get_column_solution", self.n)

    def row_activity(self):
        return self._get("orc_get_row_activity", self.m)

    def reduced_cost(self):
        return self._get("orc_get_reduced_cost", self.n)

    def row_price(self):
        return self._get("orc_get_row_price", self.m)

    def status(self):
        out = np.zeros(self.n + self.m, dtype=np.uint8)
        lib().orc_get_status(self.h, _up(out))
        return out

    # ---- kernel level
    def factorize(self, basic):
        basic = np.ascontiguousarray(basic, dtype=np.int32)
        pv = np.zeros(self.m, dtype=np.int32)
        rc = lib().orc_factorize(self.h, _ip(basic), _ip(pv))
        return rc, pv

    def ftran(self, v):
        v = np.array(v, dtype=np.float64)
        lib().orc_ftran(self.h, _dp(v))
        return v

    def btran(self, v):
        v = np.array(v, dtype=np.float64)
        lib().orc_btran(self.h, _dp(v))
        return v

    def replace_column(self, seq_in, pivot_row):
        return lib().orc_replace_column(self.h, int(seq_in), int(pivot_row))

    def transpose_times(self, scalar, pi):
        pi = np.ascontiguousarray(pi, dtype=np.float64)
        z = np.zeros(self.n)
        lib().orc_transpose_times(self.h, float(scalar), _dp(pi), _dp(z))
        return z

    def times(self, scalar, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.m)
        lib().orc_times(self.h, float(scalar), _dp(x), _dp(y))
        return y


def dual_column(alpha, dj, rng, stat, infeasibility, dual_tol=1e-7, acceptable=1e-7, bucketed=False):
    alpha = np.ascontiguousarray(alpha, dtype=np.float64)
    dj = np.ascontiguousarray(dj, dtype=np.float64)
    rng = np.ascontiguousarray(rng, dtype=np.float64)
    stat = np.ascontiguousarray(stat, dtype=np.uint8)
    theta = ctypes.c_double(0.0)
    flips = np.zeros(alpha.size, dtype=np.uint8)
    fn = lib().orc_dual_column_bucketed if bucketed else lib().orc_dual_column
    k = fn(alpha.size, _dp(alpha), _dp(dj), _dp(rng), _up(stat),
                              float(infeasibility), dual_tol, acceptable, ctypes.byref(theta),
                              _up(flips))
    return k, theta.value, flips


def dse_update(weights, alpha_col, tau, pivot_row, rho_norm2):
    w = np.array(weights, dtype=np.float64)
    a = np.ascontiguousarray(alpha_col, dtype=np.float64)
    t = np.ascontiguousarray(tau, dtype=np.float64)
    lib().orc_dse_update(w.size, _dp(w), _dp(a), _dp(t), int(pivot_row), float(rho_norm2))
    return w


def kkt_violations(lp, x, rowact, dj, tol=1e-5):
    """checkOptimalityConditions of test_racing_lp.cpp:36-116."""
    v = 0
    v += int(np.sum((x < lp.col_lower - tol) | (x > lp.col_upper + tol)))
    v += int(np.sum((rowact < lp.row_lower - tol) | (rowact > lp.row_upper + tol)))
    at_lb = (x - lp.col_lower) < tol
    at_ub = (lp.col_upper - x) < tol
    v += int(np.sum(at_lb & ~at_ub & (dj < -tol)))
    v += int(np.sum(at_ub & ~at_lb & (dj > tol)))
    v += int(np.sum(~at_lb & ~at_ub & (np.abs(dj) > tol)))
    A = lp.to_scipy()
    v += int(np.sum(np.abs(A @ x - rowact) > tol * (1 + np.abs(rowact))))
    return v
