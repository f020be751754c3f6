"""ctypes binding of oracle/libfm_oracle.so (the C oracle, fm_oracle.c).

TEST INFRASTRUCTURE ONLY (see fm_oracle.h): used by tests/, smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfm_oracle.so")


class FmoParams(C.Structure):
    _fields_ = [("task", C.c_int32), ("k", C.c_int32), ("k0", C.c_int32), ("k1", C.c_int32),
                ("n_slots", C.c_int64), ("reg0", C.c_double), ("regw", C.c_double),
                ("regv", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fm_oracle.c")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        pp = C.POINTER(FmoParams)
        L.fmo_mix64.restype = C.c_uint64
        L.fmo_mix64.argtypes = [C.c_uint64]
        L.fmo_predict_row.restype = C.c_double
        L.fmo_predict_row.argtypes = [pp, C.c_double, dp, dp, ip, dp, C.c_int64]
        L.fmo_predict.restype = None
        L.fmo_predict.argtypes = [pp, C.c_double, dp, dp, lp, ip, dp, C.c_int64, dp]
        L.fmo_predict_fast.restype = None
        L.fmo_predict_fast.argtypes = [pp, C.c_double, dp, dp, lp, ip, dp, C.c_int64, dp, C.c_int]
        L.fmo_loss_mult.restype = None
        L.fmo_loss_mult.argtypes = [C.c_int32, C.c_double, C.c_double, dp, dp]
        L.fmo_train_step.restype = C.c_double
        L.fmo_train_step.argtypes = [pp, dp, dp, dp, lp, ip, dp, dp, lp, C.c_int64, C.c_int64,
                                     C.c_double, C.c_int64, dp]
        L.fmo_gradient.restype = C.c_double
        L.fmo_gradient.argtypes = [pp, C.c_double, dp, dp, lp, ip, dp, dp, lp, C.c_int64, dp]
        L.fmo_update.restype = None
        L.fmo_update.argtypes = [pp, dp, dp, dp, dp, C.c_int64, C.c_double, C.c_int64]
        L.fmo_train_step_mt.restype = C.c_double
        L.fmo_train_step_mt.argtypes = [pp, dp, dp, dp, lp, ip, dp, dp, lp, C.c_int64, C.c_int64,
                                        C.c_double, C.c_int64, dp, C.c_int]
        L.fmo_sample_rows.restype = C.c_int64
        L.fmo_sample_rows.argtypes = [C.c_uint64, C.c_int64, C.c_double, C.c_int64, C.c_int64, lp]
        L.fmo_partition_rows.restype = C.c_int64
        L.fmo_partition_rows.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, lp]
        L.fmo_init_v.restype = None
        L.fmo_init_v.argtypes = [dp, C.c_int64, C.c_double, C.c_double, C.c_uint64]
        L.fmo_als_sweep.restype = C.c_double
        L.fmo_als_sweep.argtypes = [pp, dp, dp, dp, lp, ip, dp, dp, C.c_int64, C.c_int32, dp]
        L.fmo_max_threads.restype = C.c_int
        fp = C.POINTER(C.c_float)
        L.fmo_train_step_faithful_mt.restype = C.c_double
        L.fmo_train_step_faithful_mt.argtypes = L.fmo_train_step_mt.argtypes
        L.fmo_fast32_create.restype = C.c_void_p
        L.fmo_fast32_create.argtypes = [pp, C.c_int]
        L.fmo_fast32_destroy.restype = None
        L.fmo_fast32_destroy.argtypes = [C.c_void_p]
        L.fmo_fast32_set_model.restype = None
        L.fmo_fast32_set_model.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.fmo_fast32_get_model.restype = None
        L.fmo_fast32_get_model.argtypes = [C.c_void_p, dp, dp, dp]
        L.fmo_fast32_train_step.restype = C.c_double
        L.fmo_fast32_train_step.argtypes = [C.c_void_p, lp, ip, fp, fp, lp, C.c_int64, C.c_int64,
                                            C.c_double, C.c_int64]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _l(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class OracleFM:
    """fp64 FM model + SGD trainer on the C oracle.  Holds w0, w[n_slots], v[n_slots][k]."""

    def __init__(self, n_slots, k, task=0, k0=True, k1=True, reg=(0.0, 0.0, 0.0)):
        self.p = FmoParams(int(task), int(k), int(bool(k0)), int(bool(k1)), int(n_slots),
                           float(reg[0]), float(reg[1]), float(reg[2]))
        self.n_slots, self.k = int(n_slots), int(k)
        self.w0 = C.c_double(0.0)
        self.w = np.zeros(self.n_slots, dtype=np.float64)
        self.v = np.zeros((self.n_slots, self.k), dtype=np.float64)
        self._grad = None
        self._scratch = None

    def set_model(self, w0, w, v):
        self.w0 = C.c_double(float(w0))
        self.w = _c(w, np.float64).copy()
        self.v = _c(v, np.float64).reshape(self.n_slots, self.k).copy()

    def init_v(self, mean=0.0, stdev=0.01, seed=1):
        lib().fmo_init_v(_d(self.v), self.v.size, mean, stdev, seed)

    def predict(self, row_ptr, idx, val, fast=False, threads=1):
        row_ptr, idx, val = _c(row_ptr, np.int64), _c(idx, np.int32), _c(val, np.float64)
        n = len(row_ptr) - 1
        out = np.empty(n, dtype=np.float64)
        if fast:
            lib().fmo_predict_fast(C.byref(self.p), self.w0, _d(self.w), _d(self.v), _l(row_ptr),
                                   _i(idx), _d(val), n, _d(out), threads)
        else:
            lib().fmo_predict(C.byref(self.p), self.w0, _d(self.w), _d(self.v), _l(row_ptr),
                              _i(idx), _d(val), n, _d(out))
        return out

    def gradient(self, row_ptr, idx, val, label, row_ids):
        row_ptr, idx, val = _c(row_ptr, np.int64), _c(idx, np.int32), _c(val, np.float64)
        label, row_ids = _c(label, np.float64), _c(row_ids, np.int64)
        g = np.empty(self.n_slots * (self.k + 1) + 1, dtype=np.float64)
        loss = lib().fmo_gradient(C.byref(self.p), self.w0, _d(self.w), _d(self.v), _l(row_ptr),
                                  _i(idx), _d(val), _d(label), _l(row_ids), len(row_ids), _d(g))
        nk = self.n_slots * self.k
        return g[:nk].reshape(self.n_slots, self.k), g[nk:nk + self.n_slots], g[-1], loss

    def train_step(self, row_ptr, idx, val, label, row_ids, it, step_size, batch_count=None,
                   threads=1, faithful=False):
        """Returns the loss SUM over row_ids; updates the model in place.  faithful: the
        baseline variant whose forward makes k passes over the row (FMModel.scala:48-51)."""
        row_ptr, idx, val = _c(row_ptr, np.int64), _c(idx, np.int32), _c(val, np.float64)
        label, row_ids = _c(label, np.float64), _c(row_ids, np.int64)
        if batch_count is None:
            batch_count = len(row_ids)
        glen = self.n_slots * (self.k + 1) + 1
        if faithful:
            threads = max(threads, 1)
            if self._scratch is None or self._scratch.size < threads * glen:
                self._scratch = np.empty(threads * glen, dtype=np.float64)
            return lib().fmo_train_step_faithful_mt(
                C.byref(self.p), C.byref(self.w0), _d(self.w), _d(self.v), _l(row_ptr), _i(idx),
                _d(val), _d(label), _l(row_ids), len(row_ids), it, step_size, batch_count,
                _d(self._scratch), threads)
        if threads <= 1:
            if self._grad is None:
                self._grad = np.empty(glen, dtype=np.float64)
            return lib().fmo_train_step(C.byref(self.p), C.byref(self.w0), _d(self.w), _d(self.v),
                                        _l(row_ptr), _i(idx), _d(val), _d(label), _l(row_ids),
                                        len(row_ids), it, step_size, batch_count, _d(self._grad))
        if self._scratch is None or self._scratch.size < threads * glen:
            self._scratch = np.empty(threads * glen, dtype=np.float64)
        return lib().fmo_train_step_mt(C.byref(self.p), C.byref(self.w0), _d(self.w), _d(self.v),
                                       _l(row_ptr), _i(idx), _d(val), _d(label), _l(row_ids),
                                       len(row_ids), it, step_size, batch_count,
                                       _d(self._scratch), threads)


    def als_sweep(self, row_ptr, idx, val, label, ref_quirks=False, store_f32=False):
        """ALS.learn (fm/lib/ALS.scala:15-75): one sweep over w0, w, V in place.  Returns
        (rmse of the residuals after the sweep, residuals e = yhat - y)."""
        row_ptr, idx, val = _c(row_ptr, np.int64), _c(idx, np.int32), _c(val, np.float64)
        label = _c(label, np.float64)
        n = len(row_ptr) - 1
        e = np.empty(max(n, 1), dtype=np.float64)
        flags = (1 if ref_quirks else 0) | (2 if store_f32 else 0)
        rmse = lib().fmo_als_sweep(C.byref(self.p), C.byref(self.w0), _d(self.w), _d(self.v),
                                   _l(row_ptr), _i(idx), _d(val), _d(label), n, flags, _d(e))
        if rmse < 0:
            raise ValueError("fmo_als_sweep: a row stores the same feature twice (or out of memory)")
        return rmse, e[:n]


class OracleFast32:
    """Tuned CPU baseline (fp32, single pass, touched-only reduction): bench.py only."""

    def __init__(self, orc: "OracleFM", threads: int):
        self.h = lib().fmo_fast32_create(C.byref(orc.p), threads)
        if not self.h:
            raise MemoryError("fmo_fast32_create")
        self.orc = orc
        lib().fmo_fast32_set_model(self.h, orc.w0.value, _d(orc.w), _d(orc.v))

    def train_step(self, row_ptr, idx, val, label, row_ids, it, step_size, batch_count=None):
        row_ptr, idx = _c(row_ptr, np.int64), _c(idx, np.int32)
        label, row_ids = _c(label, np.float32), _c(row_ids, np.int64)
        fp = C.POINTER(C.c_float)
        vp = None if val is None else _c(val, np.float32).ctypes.data_as(fp)
        self._keep = (row_ptr, idx, label, row_ids, val)
        if batch_count is None:
            batch_count = len(row_ids)
        return lib().fmo_fast32_train_step(self.h, _l(row_ptr), _i(idx), vp,
                                           label.ctypes.data_as(fp), _l(row_ids), len(row_ids),
                                           it, step_size, batch_count)

    def get_model(self):
        n, k = self.orc.n_slots, self.orc.k
        w0 = C.c_double(0)
        w = np.empty(n, np.float64)
        v = np.empty((n, max(k, 1)), np.float64)
        lib().fmo_fast32_get_model(self.h, C.byref(w0), _d(w), _d(v))
        return w0.value, w, v[:, :k]

    def close(self):
        if self.h:
            lib().fmo_fast32_destroy(self.h)
            self.h = None


def sample_rows(seed, it, fraction, row_lo, row_hi):
    out = np.empty(max(row_hi - row_lo, 0), dtype=np.int64)
    n = lib().fmo_sample_rows(seed, it, fraction, row_lo, row_hi, _l(out))
    return out[:n].copy()


def partition_rows(seed, n_parts, part, row_lo, row_hi):
    out = np.empty(max(row_hi - row_lo, 0), dtype=np.int64)
    n = lib().fmo_partition_rows(seed, n_parts, part, row_lo, row_hi, _l(out))
    return out[:n].copy()


def n_parts_for(fraction):
    """P = round(1 / fraction), at least 1 (fraction is the fp32 value that crosses the ABI)."""
    f = float(np.float32(fraction))
    return 1 if f >= 1.0 else max(1, int(np.floor(1.0 / f + 0.5)))


def max_threads():
    return lib().fmo_max_threads()
