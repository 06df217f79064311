"""ctypes binding of the CPU oracle (oracle/bam_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs.  The product (bam_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from bam_b200.capi import (MESS_LEN, NENV, BamError, CProblem, Problem, Sizes, _dp, _ip, _ptr_d, _ptr_i, _vec,
                           declare_common)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bam_oracle.c")
    hdrs = [os.path.join(_HERE, "bam_oracle.h"), os.path.join(_HERE, "..", "include", "bam_gpu.h")]
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    lib = C.CDLL(build())
    H = C.c_void_p
    declare_common(lib, "oracle_")
    u64, i64, ci, cp, d = C.c_uint64, C.c_int64, C.c_int, C.c_char_p, C.c_double
    lib.oracle_logpost.argtypes = [H, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, ci, cp]
    lib.oracle_logpost_batch.argtypes = [H, ci, _dp, _dp, _ip, _ip, _dp, ci, cp]
    lib.oracle_apply_model.argtypes = [H, ci, _dp, _dp, _dp, _dp, _ip, cp]
    lib.oracle_calibration_run.argtypes = [H, _dp, _dp, _dp, _ip, cp]
    lib.oracle_calibration_run_states.argtypes = [H, _dp, _dp, _dp, _dp, _ip, cp]
    lib.oracle_predict.argtypes = [H, i64, _dp, _dp, ci, C.POINTER(_dp), _ip, ci, _ip, u64, ci, ci, _dp, _dp, ci, cp]
    lib.oracle_predict_states.argtypes = lib.oracle_predict.argtypes
    lib.oracle_envelope.argtypes = [_dp, i64, d, _dp]
    lib.oracle_envelope.restype = None
    lib.oracle_fit.argtypes = [H, ci, _dp, _dp, ci, ci, d, d, d, d, u64, ci, ci, _dp, C.POINTER(i64), _dp, _dp, i64, cp]
    lib.oracle_normal.argtypes = [u64, u64, C.c_uint32, C.c_uint32]
    lib.oracle_normal.restype = d
    lib.oracle_uniform.argtypes = [u64, u64, C.c_uint32, C.c_uint32]
    lib.oracle_uniform.restype = d
    lib.oracle_normal_quantile.argtypes = [d]
    lib.oracle_normal_quantile.restype = d
    lib.oracle_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    lib.oracle_philox4x32.restype = None
    lib.oracle_logpdf.argtypes = [ci, _dp, d, _dp, C.POINTER(ci), C.POINTER(ci)]
    lib.oracle_cdf.argtypes = [ci, _dp, d, _dp, C.POINTER(ci)]
    lib.oracle_textfile_compile.argtypes = [cp, C.POINTER(cp), ci, _ip, ci, _dp, ci, C.POINTER(ci), C.POINTER(ci)]
    _LIB = lib
    return lib


def _check(err, mess):
    if err > 0:
        raise BamError(err, mess.value.decode(errors="replace"))


class Oracle:
    def __init__(self, problem: Problem):
        self.lib = load()
        self.problem = problem
        self._cs = problem.c_struct()
        self.h = C.c_void_p()
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_create(C.byref(self.h), C.byref(self._cs), mess), mess)
        self.sizes = Sizes()
        self.lib.oracle_get_sizes(self.h, C.byref(self.sizes))
        self.P = self.sizes.nInfer
        self.nDpar = self.sizes.nDpar

    def close(self):
        if getattr(self, "h", None):
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def logpost_one(self, x, long_double=False):
        """-> dict(fx, prior, lkh, hyper, feas, isnull, dpar)"""
        xb = _vec(1, self.P, x)
        fx, pr, lk, hy = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        feas, isnull = C.c_int32(), C.c_int32()
        dpar = np.empty(max(self.nDpar, 1))
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_logpost(self.h, _ptr_d(xb), C.byref(fx), C.byref(pr), C.byref(lk), C.byref(hy),
                                       C.byref(feas), C.byref(isnull), _ptr_d(dpar), int(long_double), mess), mess)
        return dict(fx=fx.value, prior=pr.value, lkh=lk.value, hyper=hy.value, feas=feas.value,
                    isnull=isnull.value, dpar=dpar[: self.nDpar].copy())

    def logpost(self, x, nthreads=1, want_dpar=False):
        x = np.asarray(x, dtype=np.float64)
        K = 1 if x.ndim == 1 else x.shape[0]
        xb = _vec(K, self.P, x)
        fx = np.empty(K)
        feas = np.empty(K, dtype=np.int32)
        isnull = np.empty(K, dtype=np.int32)
        dpar = np.empty((K, max(self.nDpar, 1))) if want_dpar else None
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_logpost_batch(self.h, K, _ptr_d(xb), _ptr_d(fx), _ptr_i(feas), _ptr_i(isnull),
                                             _ptr_d(dpar) if want_dpar else C.cast(None, _dp), nthreads, mess), mess)
        if want_dpar:
            return fx, feas, isnull, dpar[:, : self.nDpar]
        return fx, feas, isnull

    def logpost_parts(self, x, long_double=False):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        r = [self.logpost_one(v, long_double) for v in x]
        g = lambda k: np.array([q[k] for q in r])  # noqa: E731
        return g("prior"), g("lkh"), g("hyper"), g("feas").astype(np.int32), g("isnull").astype(np.int32)

    def calibration_run(self, x):
        pr = self.problem
        x = np.ascontiguousarray(x, dtype=np.float64)
        Xtrue = np.empty((pr.nobs, pr.nX), order="F")
        Ysim = np.empty((pr.nobs, pr.nY), order="F")
        feas = C.c_int32(0)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_calibration_run(self.h, _ptr_d(x), _ptr_d(Xtrue), _ptr_d(Ysim), C.byref(feas), mess), mess)
        return Xtrue, Ysim, bool(feas.value)

    def calibration_run_states(self, x):
        pr = self.problem
        x = np.ascontiguousarray(x, dtype=np.float64)
        Xtrue = np.empty((pr.nobs, pr.nX), order="F")
        Ysim = np.empty((pr.nobs, pr.nY), order="F")
        state = np.empty((pr.nobs, self.sizes.nState), order="F")
        feas = C.c_int32(0)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_calibration_run_states(self.h, _ptr_d(x), _ptr_d(Xtrue), _ptr_d(Ysim), _ptr_d(state),
                                                      C.byref(feas), mess), mess)
        return Xtrue, Ysim, state, bool(feas.value)

    def apply_model(self, X, theta_fit):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        Xf = np.asfortranarray(X)
        n = Xf.shape[0]
        th = np.ascontiguousarray(theta_fit, dtype=np.float64)
        Y = np.empty((n, self.problem.nY), order="F")
        dpar = np.empty(64)
        feas = np.empty(n, dtype=np.int32)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_apply_model(self.h, n, _ptr_d(Xf), _ptr_d(th), _ptr_d(Y), _ptr_d(dpar), _ptr_i(feas),
                                           mess), mess)
        return Y, feas, dpar

    def predict(self, mc, mode, xspag, doParametric, doRemnant, seed=1, t0=0, t1=None, want_env=True,
                want_spag=False, nthreads=1, states=False):
        from bam_b200.capi import BamGpu

        nobs_pred = np.asarray(xspag[0]).shape[0]
        t1 = nobs_pred if t1 is None else t1
        mats, ptrs, nspag = BamGpu._xspag(xspag, nobs_pred)
        if mc is not None:
            mc = np.asfortranarray(np.asarray(mc, dtype=np.float64))
            Nrep = mc.shape[0]
        else:
            Nrep = 1
        mode = None if mode is None else np.ascontiguousarray(mode, dtype=np.float64)
        dr = np.asarray(doRemnant, dtype=np.int32)
        nT, nY = t1 - t0, self.problem.nY + (self.sizes.nState if states else 0)
        env = np.empty((nY, NENV, nT)) if want_env else None
        spag = np.empty((nY, nT, Nrep)) if want_spag else None
        mess = C.create_string_buffer(MESS_LEN)
        _check((self.lib.oracle_predict_states if states else self.lib.oracle_predict)(self.h, Nrep, _ptr_d(mc), _ptr_d(mode), nobs_pred, ptrs, _ptr_i(nspag),
                                       int(doParametric), _ptr_i(dr), seed, t0, t1,
                                       _ptr_d(env) if want_env else C.cast(None, _dp),
                                       _ptr_d(spag) if want_spag else C.cast(None, _dp), nthreads, mess), mess)
        return env, spag

    def fit(self, start, start_std, nAdapt, nCycles, MinMoveRate=0.1, MaxMoveRate=0.5, DownMult=0.9, UpMult=1.1,
            seed=1, burn=0, keep_every=1, chain_offset=0):
        start = np.asarray(start, dtype=np.float64)
        K = 1 if start.ndim == 1 else start.shape[0]
        s = _vec(K, self.P, start)
        sd = _vec(K, self.P, start_std)
        nIter = nAdapt * nCycles
        nKeep = max(0, (nIter - burn) // max(keep_every, 1))
        rows = np.empty((K, nKeep, self.P + 1 + self.nDpar))
        fx, fstd = np.empty((K, self.P)), np.empty((K, self.P))
        nk = C.c_int64(0)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.oracle_fit(self.h, K, _ptr_d(s), _ptr_d(sd), nAdapt, nCycles, MinMoveRate, MaxMoveRate, DownMult,
                                   UpMult, seed, burn, keep_every, _ptr_d(rows), C.byref(nk), _ptr_d(fx), _ptr_d(fstd),
                                   chain_offset, mess), mess)
        return rows, fx, fstd


def envelope(col, unfeas=-666.666):
    lib = load()
    col = np.ascontiguousarray(col, dtype=np.float64)
    out = np.empty(NENV)
    lib.oracle_envelope(_ptr_d(col), col.size, unfeas, _ptr_d(out))
    return out


def normal(seed, a, b, c):
    return load().oracle_normal(seed, a, b, c)


def uniform(seed, a, b, c):
    return load().oracle_uniform(seed, a, b, c)


def philox(c, k):
    out = (C.c_uint32 * 4)()
    load().oracle_philox4x32(c[0], c[1], c[2], c[3], k[0], k[1], out)
    return list(out)


def normal_quantile(p):
    return load().oracle_normal_quantile(p)
