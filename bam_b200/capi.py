"""ctypes binding of include/bam_gpu.h.

``Problem`` mirrors ``bamgpu_problem`` (the reference's INFER + MODEL module globals after
LoadBamObjects, src/BaM_tools.f90:73-101,133-420).  ``BamGpu`` wraps a ``bamgpu_t`` handle.
There is deliberately no CPU path here: if ``libbam_gpu.so`` is missing or no CUDA device is
present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

MESS_LEN = 250
PRIOR_NPAR = 4
UNDEF_RN = -999999999.0
NENV = 7

DIST = {
    "Gaussian": 1,
    "Uniform": 2,
    "Triangle": 3,
    "LogNormal": 4,
    "Exponential": 5,
    "FlatPrior": 6,
    "FlatPrior+": 7,
    "FlatPrior-": 8,
}
DIST_NPAR = {1: 2, 2: 2, 3: 3, 4: 2, 5: 2, 6: 0, 7: 0, 8: 0}
SIGMAFUNC = {"Constant": 1, "Linear": 2, "Proportional": 3, "Power": 4, "Exponential": 5, "Gaussian": 6}
SIGMAFUNC_NPAR = {1: 1, 2: 2, 3: 1, 4: 3, 5: 3, 6: 3}
PARTYPE = {"FIX": -1, "STD": 0, "VAR": 1}
SED_FORMULA = {"MPM": 1, "Camenen": 2, "Nielsen": 3, "Smart": 4}
SED_CSS = {"Constant": 1, "Soulsby": 2, "SoulsbyWhitehouse": 3}
SED_PPO = {"FIX": 1, "IN": 2, "PAR": 3}
SED_CO = {"FIX": 1, "PAR": 2}
SWOT_ID = {"SWOT_hS": 1, "SWOT_hSB": 2}
ORTHO_DISTO = {"None": 1, "Radial": 2}
VEG_GROWTH = {"Growth_Yin1": 1, "Growth_Yin2": 2, "Growth_Yin3": 3, "Growth_Yin1_temporal": 4}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class CProblem(C.Structure):
    _fields_ = [
        ("model_id", C.c_char_p),
        ("nobs", C.c_int32),
        ("nX", C.c_int32),
        ("nY", C.c_int32),
        ("ntheta", C.c_int32),
        ("X", _dp),
        ("Xu", _dp),
        ("Xb", _dp),
        ("Xbindx", _ip),
        ("Y", _dp),
        ("Yu", _dp),
        ("Yb", _dp),
        ("Ybindx", _ip),
        ("partype", _ip),
        ("nval", _ip),
        ("fix_value", _dp),
        ("var_indx", _ip),
        ("prior_theta_dist", _ip),
        ("prior_theta_par", _dp),
        ("sigma_func", _ip),
        ("nremnant", _ip),
        ("prior_gamma_dist", _ip),
        ("prior_gamma_par", _dp),
        ("priorM", _dp),
        ("xtra_i", _ip),
        ("n_xtra_i", C.c_int32),
        ("xtra_d", _dp),
        ("n_xtra_d", C.c_int32),
        ("xtra_text", C.c_char_p),
        ("mv", C.c_double),
        ("unfeas_flag", C.c_double),
    ]


class Sizes(C.Structure):
    _fields_ = [
        ("nFit", C.c_int32),
        ("nGamma", C.c_int32),
        ("nLatent", C.c_int32),
        ("nXbias", C.c_int32),
        ("nYbias", C.c_int32),
        ("nInfer", C.c_int32),
        ("nDpar", C.c_int32),
        ("nCombo", C.c_int32),
        ("nState", C.c_int32),
    ]


def _f64(a, shape=None):
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return np.asfortranarray(a)


def _i32(a, shape=None):
    a = np.asarray(a, dtype=np.int32)
    if shape is not None:
        a = a.reshape(shape)
    return np.asfortranarray(a)


def _ptr_d(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return C.cast(None, _dp)
    return a.ctypes.data_as(_dp)


def _ptr_i(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return C.cast(None, _ip)
    return a.ctypes.data_as(_ip)


class Problem:
    """Python-side owner of the arrays behind a ``bamgpu_problem``.

    All matrices are stored Fortran-ordered (column-major), as the reference passes them.
    ``priors_theta`` / ``priors_gamma`` are lists of ``(dist_name, [pars])`` for the fitted
    theta values (VAR parameters expanded, FIX skipped) and for the remnant parameters of every
    output in order.
    """

    def __init__(
        self,
        model_id: str,
        X,
        Y,
        Yu=None,
        Xu=None,
        Xb=None,
        Xbindx=None,
        Yb=None,
        Ybindx=None,
        partype: Sequence[int] = (),
        nval: Optional[Sequence[int]] = None,
        fix_value: Optional[Sequence[float]] = None,
        var_indx=None,
        priors_theta: Sequence = (),
        sigma_func: Sequence = (),
        priors_gamma: Sequence = (),
        priorM=None,
        xtra_i: Sequence[int] = (),
        xtra_d: Sequence[float] = (),
        xtra_text: Optional[str] = None,
        mv: float = -9999.0,
        unfeas_flag: float = -666.666,
    ):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        if Y.ndim == 1:
            Y = Y.reshape(-1, 1)
        n, nX = X.shape
        nY = Y.shape[1]
        self.model_id = model_id
        self.nobs, self.nX, self.nY = n, nX, nY
        self.X = _f64(X)
        self.Y = _f64(Y)
        z = lambda a, s, f: f(np.zeros(s) if a is None else a, s)  # noqa: E731
        self.Xu = z(Xu, (n, nX), _f64)
        self.Xb = z(Xb, (n, nX), _f64)
        self.Xbindx = z(Xbindx, (n, nX), _i32)
        self.Yu = z(Yu, (n, nY), _f64)
        self.Yb = z(Yb, (n, nY), _f64)
        self.Ybindx = z(Ybindx, (n, nY), _i32)
        self.partype = _i32(partype)
        self.ntheta = int(self.partype.size)
        self.nval = _i32(np.ones(self.ntheta) if nval is None else nval)
        self.fix_value = _f64(np.zeros(self.ntheta) if fix_value is None else fix_value)
        self.var_indx = None if var_indx is None else _i32(var_indx, (n, self.ntheta))
        self.sigma_func = _i32([SIGMAFUNC[s] if isinstance(s, str) else int(s) for s in sigma_func])
        self.nremnant = _i32([SIGMAFUNC_NPAR[int(s)] for s in self.sigma_func])
        self.prior_theta_dist, self.prior_theta_par = self._flatten_priors(priors_theta)
        self.prior_gamma_dist, self.prior_gamma_par = self._flatten_priors(priors_gamma)
        self.nFit = int(sum(int(v) for v, t in zip(self.nval, self.partype) if t != PARTYPE["FIX"]))
        if len(priors_theta) != self.nFit:
            raise ValueError(f"priors_theta has {len(priors_theta)} entries, expected nFit={self.nFit}")
        if len(priors_gamma) != int(self.nremnant.sum()):
            raise ValueError("priors_gamma does not match the sigma functions")
        k = self.nFit + int(self.nremnant.sum())
        self.priorM = None if priorM is None else _f64(priorM, (k, k))
        self.xtra_i = _i32(xtra_i)
        self.xtra_d = _f64(xtra_d)
        self.xtra_text = None if xtra_text is None else xtra_text.encode()
        self.mv = float(mv)
        self.unfeas_flag = float(unfeas_flag)
        self._model_id_b = model_id.encode()

    @staticmethod
    def _flatten_priors(priors):
        dist = np.zeros(len(priors), dtype=np.int32)
        par = np.zeros((len(priors), PRIOR_NPAR), dtype=np.float64)  # row per parameter, C order
        for i, (d, p) in enumerate(priors):
            dist[i] = DIST[d] if isinstance(d, str) else int(d)
            p = list(p)[: DIST_NPAR[int(dist[i])]]
            par[i, : len(p)] = p
        return dist, np.ascontiguousarray(par)

    def c_struct(self) -> CProblem:
        s = CProblem()
        s.model_id = self._model_id_b
        s.nobs, s.nX, s.nY, s.ntheta = self.nobs, self.nX, self.nY, self.ntheta
        s.X, s.Xu, s.Xb, s.Xbindx = _ptr_d(self.X), _ptr_d(self.Xu), _ptr_d(self.Xb), _ptr_i(self.Xbindx)
        s.Y, s.Yu, s.Yb, s.Ybindx = _ptr_d(self.Y), _ptr_d(self.Yu), _ptr_d(self.Yb), _ptr_i(self.Ybindx)
        s.partype, s.nval, s.fix_value = _ptr_i(self.partype), _ptr_i(self.nval), _ptr_d(self.fix_value)
        s.var_indx = _ptr_i(self.var_indx)
        s.prior_theta_dist, s.prior_theta_par = _ptr_i(self.prior_theta_dist), _ptr_d(self.prior_theta_par)
        s.sigma_func, s.nremnant = _ptr_i(self.sigma_func), _ptr_i(self.nremnant)
        s.prior_gamma_dist, s.prior_gamma_par = _ptr_i(self.prior_gamma_dist), _ptr_d(self.prior_gamma_par)
        s.priorM = _ptr_d(self.priorM)
        s.xtra_i, s.n_xtra_i = _ptr_i(self.xtra_i), int(self.xtra_i.size)
        s.xtra_d, s.n_xtra_d = _ptr_d(self.xtra_d), int(self.xtra_d.size)
        s.xtra_text = self.xtra_text
        s.mv, s.unfeas_flag = self.mv, self.unfeas_flag
        return s

    # -- (de)serialisation used by the golden fixtures ------------------------------------
    @classmethod
    def from_dict(cls, d: dict) -> "Problem":
        n, nX, nY = d["nobs"], d["nX"], d["nY"]
        m = lambda key, cols: None if d.get(key) is None else np.asarray(d[key], dtype=np.float64).reshape(cols, n).T  # noqa: E731
        mi = lambda key, cols: None if d.get(key) is None else np.asarray(d[key], dtype=np.int32).reshape(cols, n).T  # noqa: E731
        ntheta = len(d["partype"])
        return cls(
            d["model_id"],
            m("X", nX),
            m("Y", nY),
            Yu=m("Yu", nY),
            Xu=m("Xu", nX),
            Xb=m("Xb", nX),
            Xbindx=mi("Xbindx", nX),
            Yb=m("Yb", nY),
            Ybindx=mi("Ybindx", nY),
            partype=d["partype"],
            nval=d.get("nval"),
            fix_value=d.get("fix_value"),
            var_indx=mi("var_indx", ntheta),
            priors_theta=[tuple(p) for p in d["priors_theta"]],
            sigma_func=d["sigma_func"],
            priors_gamma=[tuple(p) for p in d["priors_gamma"]],
            priorM=d.get("priorM"),
            xtra_i=d.get("xtra_i", ()),
            xtra_d=d.get("xtra_d", ()),
            xtra_text=d.get("xtra_text"),
            mv=d.get("mv", -9999.0),
            unfeas_flag=d.get("unfeas_flag", -666.666),
        )


# ---------------------------------------------------------------------------------------------
# library loading
# ---------------------------------------------------------------------------------------------

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def library_path() -> str:
    return os.path.join(_REPO, "bam_b200", "libbam_gpu.so")


def declare_common(lib, prefix: str, handle_t=C.c_void_p):
    """argtypes shared by the CUDA library (prefix 'bamgpu_') and the oracle (prefix 'oracle_')."""
    f = getattr(lib, prefix + "create")
    f.argtypes = [C.POINTER(handle_t), C.POINTER(CProblem)] + ([C.c_int] if prefix == "bamgpu_" else []) + [C.c_char_p]
    f.restype = C.c_int
    getattr(lib, prefix + "destroy").argtypes = [handle_t]
    getattr(lib, prefix + "destroy").restype = None
    g = getattr(lib, prefix + "get_sizes")
    g.argtypes = [handle_t, C.POINTER(Sizes)]
    g.restype = C.c_int


def load_library(path: Optional[str] = None):
    """Load libbam_gpu.so; raises if it is not built (there is no fallback)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    declare_common(lib, "bamgpu_")
    u64, i64, ci, cp = C.c_uint64, C.c_int64, C.c_int, C.c_char_p
    lib.bamgpu_logpost.argtypes = [H, ci, _dp, _dp, _ip, _ip, _dp, cp]
    lib.bamgpu_logpost_parts.argtypes = [H, ci, _dp, _dp, _dp, _dp, _ip, _ip, cp]
    lib.bamgpu_calibration_run.argtypes = [H, C.POINTER(CProblem), _dp, _dp, _dp, _ip, cp]
    lib.bamgpu_calibration_run_states.argtypes = [H, C.POINTER(CProblem), _dp, _dp, _dp, _dp, _ip, cp]
    lib.bamgpu_chains_upload.argtypes = [H, ci, _dp, cp]
    lib.bamgpu_logpost_resident.argtypes = [H, cp]
    lib.bamgpu_chains_download.argtypes = [H, _dp, _ip, _ip, _dp, cp]
    lib.bamgpu_sync.argtypes = [H, cp]
    lib.bamgpu_fit.argtypes = [H, ci, _dp, _dp, ci, ci, C.c_double, C.c_double, C.c_double, C.c_double, u64, ci, ci,
                               _dp, C.POINTER(i64), cp]
    lib.bamgpu_moments.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, ci, cp]
    lib.bamgpu_moments_window.argtypes = [H, i64, ci, cp]
    lib.bamgpu_rhat.argtypes = [ci, ci, _dp, _dp, _dp, _dp, cp]
    lib.bamgpu_phase_times.argtypes = [H, _dp, _dp, ci]
    lib.bamgpu_comm_unique_id.argtypes = [cp, cp]
    lib.bamgpu_comm_init_rank.argtypes = [C.POINTER(H), ci, ci, cp, ci, cp]
    lib.bamgpu_comm_init_all.argtypes = [C.POINTER(H), ci, _ip, cp]
    lib.bamgpu_comm_destroy.argtypes = [H]
    lib.bamgpu_comm_destroy.restype = None
    lib.bamgpu_comm_info.argtypes = [H, _ip, _ip, _ip]
    lib.bamgpu_moments_allgather.argtypes = [H, H, _dp, _dp, _dp, _dp, _dp, cp]
    lib.bamgpu_envelope_allgather.argtypes = [H, H, ci, _dp, _dp, cp]
    lib.bamgpu_comm_allreduce_sum.argtypes = [H, C.c_void_p, C.c_size_t, ci, C.c_void_p, cp]
    lib.bamgpu_predict_sample_sharded.argtypes = [H, H, i64, i64, ci, _ip, u64, ci, ci, cp]
    lib.bamgpu_predict.argtypes = [H, i64, _dp, _dp, ci, C.POINTER(_dp), _ip, ci, _ip, u64, ci, ci, _dp, _dp, cp]
    lib.bamgpu_predict_states.argtypes = lib.bamgpu_predict.argtypes
    lib.bamgpu_predict_upload.argtypes = [H, i64, _dp, _dp, ci, C.POINTER(_dp), _ip, cp]
    lib.bamgpu_predict_resident.argtypes = [H, ci, _ip, u64, ci, ci, cp]
    lib.bamgpu_predict_download.argtypes = [H, _dp, cp]
    lib.bamgpu_prior_corr_to_M.argtypes = [ci, _dp, _dp, cp]
    lib.bamgpu_launch_count.argtypes = [H]
    lib.bamgpu_launch_count.restype = i64
    lib.bamgpu_last_kernel_ms.argtypes = [H]
    lib.bamgpu_last_kernel_ms.restype = C.c_double
    lib.bamgpu_fp64_peak.argtypes = [ci, _dp, cp]
    for name in ("step_begin", "step_end"):
        getattr(lib, "bamgpu_" + name).argtypes = [H]
    lib.bamgpu_step_times.argtypes = [H, _dp, ci]
    lib.bamgpu_kernel_times.argtypes = [H, _dp, ci]
    lib.bamgpu_flush_l2.argtypes = [H, C.c_size_t]
    lib.bamgpu_pin.argtypes = [C.c_void_p, C.c_size_t]
    lib.bamgpu_unpin.argtypes = [C.c_void_p]
    lib.bamgpu_set_chain_offset.argtypes = [H, i64]
    lib.bamgpu_set_option.argtypes = [H, cp, ci]
    lib.bamgpu_version.restype = cp
    for name in ("dist_id", "sigmafunc_id", "model_id"):
        getattr(lib, "bamgpu_" + name).argtypes = [cp]
    if path == library_path():
        _LIB = lib
    return lib


class BamError(RuntimeError):
    def __init__(self, err: int, mess: str):
        super().__init__(f"err={err}: {mess}")
        self.err = err
        self.mess = mess


def _check(err: int, mess) -> None:
    if err > 0:
        raise BamError(err, mess.value.decode(errors="replace"))


def _vec(K: int, P: int, x) -> np.ndarray:
    """K parameter vectors -> contiguous buffer with vector k at offset k*P (P x K column-major)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if x.shape != (K, P):
        raise ValueError(f"expected {K} vectors of length {P}, got {x.shape}")
    return np.ascontiguousarray(x)


class BamGpu:
    """A ``bamgpu_t`` handle: the GPU-resident problem + chain workspace."""

    def __init__(self, problem: Problem, device: int = 0, lib=None):
        self.lib = lib or load_library()
        self.problem = problem
        self._cs = problem.c_struct()
        self.h = C.c_void_p()
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_create(C.byref(self.h), C.byref(self._cs), device, mess), mess)
        self.sizes = Sizes()
        self.lib.bamgpu_get_sizes(self.h, C.byref(self.sizes))
        self.P = self.sizes.nInfer
        self.nDpar = self.sizes.nDpar
        self._K = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.bamgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- log-posterior -------------------------------------------------------------------
    def logpost(self, x, want_dpar: bool = False):
        x = np.asarray(x, dtype=np.float64)
        K = 1 if x.ndim == 1 else x.shape[0]
        xb = _vec(K, self.P, x)
        fx = np.empty(K)
        feas = np.empty(K, dtype=np.int32)
        isnull = np.empty(K, dtype=np.int32)
        dpar = np.empty((K, max(self.nDpar, 1))) if want_dpar else None
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_logpost(self.h, K, _ptr_d(xb), _ptr_d(fx), _ptr_i(feas), _ptr_i(isnull),
                                       _ptr_d(dpar) if want_dpar else C.cast(None, _dp), mess), mess)
        if want_dpar:
            return fx, feas, isnull, dpar[:, : self.nDpar]
        return fx, feas, isnull

    def logpost_parts(self, x):
        x = np.asarray(x, dtype=np.float64)
        K = 1 if x.ndim == 1 else x.shape[0]
        xb = _vec(K, self.P, x)
        prior, lkh, hyper = np.empty(K), np.empty(K), np.empty(K)
        feas = np.empty(K, dtype=np.int32)
        isnull = np.empty(K, dtype=np.int32)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_logpost_parts(self.h, K, _ptr_d(xb), _ptr_d(prior), _ptr_d(lkh), _ptr_d(hyper),
                                             _ptr_i(feas), _ptr_i(isnull), mess), mess)
        return prior, lkh, hyper, feas, isnull

    def chains_upload(self, x):
        x = np.asarray(x, dtype=np.float64)
        K = 1 if x.ndim == 1 else x.shape[0]
        xb = _vec(K, self.P, x)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_chains_upload(self.h, K, _ptr_d(xb), mess), mess)
        self._K = K

    def logpost_resident(self):
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_logpost_resident(self.h, mess), mess)

    def sync(self):
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_sync(self.h, mess), mess)

    def chains_download(self, want_dpar: bool = False):
        K = self._K
        fx = np.empty(K)
        feas = np.empty(K, dtype=np.int32)
        isnull = np.empty(K, dtype=np.int32)
        dpar = np.empty((K, max(self.nDpar, 1))) if want_dpar else None
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_chains_download(self.h, _ptr_d(fx), _ptr_i(feas), _ptr_i(isnull),
                                               _ptr_d(dpar) if want_dpar else C.cast(None, _dp), mess), mess)
        if want_dpar:
            return fx, feas, isnull, dpar[:, : self.nDpar]
        return fx, feas, isnull

    # -- sampler -------------------------------------------------------------------------
    def calibration_run(self, x):
        """GetSim_calibration: (Xtrue n x nX, Ysim n x nY, feas) for one parameter vector."""
        pr = self.problem
        x = np.ascontiguousarray(x, dtype=np.float64)
        Xtrue = np.empty((pr.nobs, pr.nX), order="F")
        Ysim = np.empty((pr.nobs, pr.nY), order="F")
        feas = C.c_int32(0)
        mess = C.create_string_buffer(MESS_LEN)
        cs = pr.c_struct()
        _check(self.lib.bamgpu_calibration_run(self.h, C.byref(cs), _ptr_d(x), _ptr_d(Xtrue), _ptr_d(Ysim),
                                               C.byref(feas), mess), mess)
        return Xtrue, Ysim, bool(feas.value)

    def calibration_run_states(self, x):
        """as calibration_run, plus the state variables (n x nState): (Xtrue, Ysim, state, feas)"""
        pr = self.problem
        x = np.ascontiguousarray(x, dtype=np.float64)
        Xtrue = np.empty((pr.nobs, pr.nX), order="F")
        Ysim = np.empty((pr.nobs, pr.nY), order="F")
        state = np.empty((pr.nobs, self.sizes.nState), order="F")
        feas = C.c_int32(0)
        mess = C.create_string_buffer(MESS_LEN)
        cs = pr.c_struct()
        _check(self.lib.bamgpu_calibration_run_states(self.h, C.byref(cs), _ptr_d(x), _ptr_d(Xtrue), _ptr_d(Ysim),
                                                      _ptr_d(state), C.byref(feas), mess), mess)
        return Xtrue, Ysim, state, bool(feas.value)

    def fit(self, start, start_std, nAdapt, nCycles, MinMoveRate=0.1, MaxMoveRate=0.5, DownMult=0.9, UpMult=1.1,
            seed=1, burn=0, keep_every=1, want_rows=True):
        start = np.asarray(start, dtype=np.float64)
        K = 1 if start.ndim == 1 else start.shape[0]
        s = _vec(K, self.P, start)
        sd = _vec(K, self.P, start_std)
        nIter = nAdapt * nCycles
        nKeep = max(0, (nIter - burn) // max(keep_every, 1))
        rowlen = self.P + 1 + self.nDpar
        rows = np.empty((K, nKeep, rowlen)) if want_rows else None
        nk = C.c_int64(0)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_fit(self.h, K, _ptr_d(s), _ptr_d(sd), nAdapt, nCycles, MinMoveRate, MaxMoveRate,
                                   DownMult, UpMult, seed, burn, keep_every,
                                   _ptr_d(rows) if want_rows else C.cast(None, _dp), C.byref(nk), mess), mess)
        self._K = K
        return rows

    def moments(self):
        K, P = self._K, self.P
        count, mean, m2 = np.empty(K), np.empty((K, P)), np.empty((K, P))
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_moments(self.h, count.ctypes.data, mean.ctypes.data, m2.ctypes.data, 0, mess), mess)
        return count, mean, m2

    def moments_window(self, first: int, every: int = 1):
        """moments over rows first, first+every, ... of the last fit (the cooked sample)"""
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_moments_window(self.h, first, every, mess), mess)

    def moments_device(self, count_ptr: int, mean_ptr: int, m2_ptr: int):
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_moments(self.h, count_ptr, mean_ptr, m2_ptr, 1, mess), mess)

    def rhat(self, count, mean, m2):
        count = np.ascontiguousarray(count, dtype=np.float64)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        m2 = np.ascontiguousarray(m2, dtype=np.float64)
        K, P = mean.shape
        out = np.empty(P)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_rhat(K, P, _ptr_d(count), _ptr_d(mean), _ptr_d(m2), _ptr_d(out), mess), mess)
        return out

    # -- prediction ----------------------------------------------------------------------
    @staticmethod
    def _xspag(xspag, nobs_pred):
        mats = []
        for m in xspag:
            m = np.asarray(m, dtype=np.float64)
            if m.ndim == 1:
                m = m.reshape(-1, 1)
            assert m.shape[0] == nobs_pred
            mats.append(np.asfortranarray(m))
        ptrs = (_dp * len(mats))(*[_ptr_d(m) for m in mats])
        nspag = np.asarray([m.shape[1] for m in mats], dtype=np.int32)
        return mats, ptrs, nspag

    def predict(self, mc, mode, xspag, doParametric, doRemnant, seed=1, t0=0, t1=None, want_env=True,
                want_spag=False, states=False):
        """states=True: the nState state variables follow the nY outputs in env / spag (bamgpu_predict_states)"""
        nobs_pred = np.asarray(xspag[0]).shape[0]
        t1 = nobs_pred if t1 is None else t1
        mats, ptrs, nspag = self._xspag(xspag, nobs_pred)
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
        _check((self.lib.bamgpu_predict_states if states else self.lib.bamgpu_predict)(self.h, Nrep, _ptr_d(mc), _ptr_d(mode), nobs_pred, ptrs, _ptr_i(nspag),
                                       int(doParametric), _ptr_i(dr), seed, t0, t1,
                                       _ptr_d(env) if want_env else C.cast(None, _dp),
                                       _ptr_d(spag) if want_spag else C.cast(None, _dp), mess), mess)
        return env, spag

    def predict_upload(self, mc, mode, xspag):
        nobs_pred = np.asarray(xspag[0]).shape[0]
        mats, ptrs, nspag = self._xspag(xspag, nobs_pred)
        mc = np.asfortranarray(np.asarray(mc, dtype=np.float64))
        mode = None if mode is None else np.ascontiguousarray(mode, dtype=np.float64)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_predict_upload(self.h, mc.shape[0], _ptr_d(mc), _ptr_d(mode), nobs_pred, ptrs,
                                              _ptr_i(nspag), mess), mess)
        self._pred_n = nobs_pred

    def predict_resident(self, doParametric, doRemnant, seed=1, t0=0, t1=None):
        t1 = self._pred_n if t1 is None else t1
        dr = np.asarray(doRemnant, dtype=np.int32)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_predict_resident(self.h, int(doParametric), _ptr_i(dr), seed, t0, t1, mess), mess)
        self._pred_nT = t1 - t0

    def predict_download(self):
        env = np.empty((self.problem.nY, NENV, self._pred_nT))
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_predict_download(self.h, _ptr_d(env), mess), mess)
        return env

    # -- measurement ---------------------------------------------------------------------
    def step_begin(self):
        self.lib.bamgpu_step_begin(self.h)

    def step_end(self):
        self.lib.bamgpu_step_end(self.h)

    def step_times(self, cap=8192):
        buf = np.empty(cap)
        n = self.lib.bamgpu_step_times(self.h, _ptr_d(buf), cap)
        return buf[:n].copy()

    def kernel_times(self, cap=8192):
        buf = np.empty(cap)
        n = self.lib.bamgpu_kernel_times(self.h, _ptr_d(buf), cap)
        return buf[:n].copy()

    def flush_l2(self, nbytes=256 << 20):
        self.lib.bamgpu_flush_l2(self.h, nbytes)

    def set_option(self, name: str, value: int):
        if self.lib.bamgpu_set_option(self.h, name.encode(), int(value)) != 0:
            raise ValueError(f"unknown option {name}")

    # -- multi-GPU exchanges (NCCL inside the library) ---------------------------------------
    def moments_allgather(self, comm: "Comm"):
        """R-hat over the chains of all ranks; returns (rhat[P], ms of the collective)"""
        out = np.empty(self.P)
        ms = np.zeros(1)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_moments_allgather(self.h, comm.c, _ptr_d(out), None, None, None, _ptr_d(ms), mess), mess)
        return out, float(ms[0])

    def envelope_allgather(self, comm: "Comm", nobs_pred: int, n_out: Optional[int] = None):
        """envelopes of the whole prediction from the per-rank slices; returns (env[nOut, 7, nobs_pred], ms)"""
        n_out = n_out or self.problem.nY
        env = np.empty((n_out, 7, nobs_pred))
        ms = np.zeros(1)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_envelope_allgather(self.h, comm.c, nobs_pred, _ptr_d(env), _ptr_d(ms), mess), mess)
        return env, float(ms[0])

    def predict_sample_sharded(self, comm: "Comm", rep_base: int, nrep_total: int, doParametric, doRemnant, seed=1, t0=0, t1=None):
        """after predict_upload of THIS rank's rows of the sample: envelopes of [t0, t1) over the rows of all ranks"""
        t1 = self._pred_n if t1 is None else t1
        dr = np.asarray(doRemnant, dtype=np.int32)
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_predict_sample_sharded(self.h, comm.c, rep_base, nrep_total, int(doParametric), _ptr_i(dr), seed, t0, t1, mess), mess)
        self._pred_nT = t1 - t0

    def set_chain_offset(self, off: int):
        self.lib.bamgpu_set_chain_offset(self.h, off)

    def phase_times(self, cap: int = 8192):
        a, b = np.zeros(cap), np.zeros(cap)
        n = self.lib.bamgpu_phase_times(self.h, _ptr_d(a), _ptr_d(b), cap)
        return a[:n].copy(), b[:n].copy()

    def launch_count(self) -> int:
        return int(self.lib.bamgpu_launch_count(self.h))

    def last_kernel_ms(self) -> float:
        return float(self.lib.bamgpu_last_kernel_ms(self.h))


class Comm:
    """One NCCL communicator rank (bamgpu_comm_t).  `unique_id()` on rank 0, broadcast the 128 bytes, `Comm(n, r, id, dev)`."""

    def __init__(self, nranks: int, rank: int, uid: bytes, device: int, lib=None):
        self.lib = lib or load_library()
        self.c = C.c_void_p()
        mess = C.create_string_buffer(MESS_LEN)
        _check(self.lib.bamgpu_comm_init_rank(C.byref(self.c), nranks, rank, uid, device, mess), mess)
        self.nranks, self.rank = nranks, rank

    @staticmethod
    def unique_id(lib=None) -> bytes:
        lib = lib or load_library()
        buf = C.create_string_buffer(128)
        mess = C.create_string_buffer(MESS_LEN)
        _check(lib.bamgpu_comm_unique_id(buf, mess), mess)
        return buf.raw

    def nccl_version(self) -> int:
        v = np.zeros(1, dtype=np.int32)
        self.lib.bamgpu_comm_info(self.c, None, None, v.ctypes.data_as(C.POINTER(C.c_int32)))
        return int(v[0])

    def close(self):
        if self.c:
            self.lib.bamgpu_comm_destroy(self.c)
            self.c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
