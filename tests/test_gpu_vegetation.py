"""GPU parity tests of the Vegetation model (row a14 of SURVEY.md section 8) against the CPU oracle.

Vegetation_model.f90:87-569: per-parameter-set growth pre-pass (closed form for Growth_Yin1_temporal, per-year
degree-day scan for Yin1/2/3), then per observation a bracketed Newton solve (znewton, un-vendored: declared
convention, see oracle/bam_oracle.c:veg_znewton).  Both sides polish the root, so parity is 1e-12 relative on the
log-posterior and on every output column; row feasibility (integer work) must be identical."""
import numpy as np
import pytest

from workloads import synth
from bam_b200.capi import VEG_GROWTH, BamGpu, Problem
from oracle.oracle_api import Oracle
from tests.helpers import load_fixture
from tests.test_gpu_logpost import check_parity
from tests.test_gpu_predict import cmp_env, cmp_spag

pytestmark = pytest.mark.gpu

CASES = [("Growth_Yin1_temporal", False, 5), ("Growth_Yin1_temporal", True, 5), ("Growth_Yin1_temporal", False, 4),
         ("Growth_Yin1", False, 5), ("Growth_Yin2", False, 5), ("Growth_Yin3", False, 4)]


@pytest.mark.parametrize("growth,var,nY", CASES)
def test_logpost_parity(growth, var, nY):
    prob, truth = synth.vegetation(nobs=700, growth=growth, var_gmax=var, nY=nY)
    x = synth.feasible_starts(truth, 64, rel=0.02, seed=5)
    x[3, 5] = 0.3     # chi > 0: infeasible parameter set (:183-186)
    x[4, 0] = -1.0    # B <= 0
    if growth.endswith("temporal"):
        x[5, 8] = x[5, 7] - 0.01  # t_i < t_0 (:397-401)
    else:
        x[5, 9] = x[5, 8] - 1.0   # Te <= Tb (:423)
        x[6, 9] = 1e6             # Te never reached within the year (:466-470)
    err = check_parity(prob, x)
    assert err <= 1e-12


def test_known_answer_fixture_on_device():
    """the reference's own SyntheticData.txt columns, through the device prediction path (maxpost run)"""
    d = load_fixture("vegetation_synthetic")
    X = np.column_stack([d["time"], d["stage"], d["Temp"], d["Temp"]])
    Y = np.column_stack([d["Q"], d["g"], d["cosg"], d["sing"], d["g0"]])
    nt = len(d["theta"])
    nw = d["newton"]
    prob = Problem("Vegetation", X, Y, partype=[0] * nt, priors_theta=[("FlatPrior", [])] * nt,
                   sigma_func=["Constant"] * 5, priors_gamma=[("FlatPrior", [])] * 5,
                   xtra_i=[VEG_GROWTH[d["growth"]], d["nYear"], int(nw[4])], xtra_d=nw[:4])
    g, o = BamGpu(prob), Oracle(prob)
    mode = np.concatenate([d["theta"], np.ones(5)])
    _, spag = g.predict(None, mode, [X[:, j] for j in range(4)], 0, [0] * 5, want_env=False, want_spag=True)
    Yh, feas, _ = o.apply_model(X, np.asarray(d["theta"]))
    assert feas.all()
    got = spag[:, :, 0].T  # [y][t][rep] -> n x nY
    assert np.max(np.abs(got - Yh) / np.maximum(np.abs(Yh), 1.0)) <= 1e-13
    tfrac = X[:, 0] - np.floor(X[:, 0])
    calm = (tfrac < 100 / 365.0) | (tfrac > 0.6)  # outside the season the file's Q is Q0 to its printed precision
    assert np.max(np.abs(got[calm, 0] - Y[calm, 0]) / Y[calm, 0]) <= 2e-9
    assert np.max(np.abs(got - Y) / np.maximum(np.abs(Y), 1.0)) <= 3e-7


@pytest.mark.parametrize("growth", ["Growth_Yin1_temporal", "Growth_Yin2"])
def test_prediction_parity(growth):
    prob, truth = synth.vegetation(nobs=400, growth=growth)
    rng = np.random.default_rng(8)
    Nrep = 300
    mc = truth * (1.0 + 0.02 * rng.standard_normal((Nrep, truth.size)))
    mc[7, 5] = 0.2  # infeasible replication -> a row of flags
    xs = [prob.X[:, j].copy() for j in range(prob.nX)]
    g, o = BamGpu(prob), Oracle(prob)
    for dop, dor in ((1, [0] * 5), (1, [1, 1, 0, 0, 1])):
        env, spag = g.predict(mc, truth, xs, dop, dor, seed=11, want_spag=True)
        oenv, ospag = o.predict(mc, truth, xs, dop, dor, seed=11, want_spag=True, nthreads=8)
        cmp_spag(spag[:1], ospag[:1])  # discharge
        cmp_env(env[:1], oenv[:1])
        # g, cos(pi g0), sin(pi g0), g0 are O(1) quantities (sin(pi g0) -> 0 at the peak of the season, where its
        # RELATIVE error is unbounded for any implementation): unit scale
        assert ((spag[1:] == -666.666) == (ospag[1:] == -666.666)).all()
        m = ospag[1:] != -666.666
        assert np.max(np.abs(spag[1:][m] - ospag[1:][m]) / np.maximum(np.abs(ospag[1:][m]), 1.0)) <= 1e-12
        me = oenv[1:] != -666.666
        assert ((env[1:] == -666.666) == ~me).all()
        assert np.max(np.abs(env[1:][me] - oenv[1:][me]) / np.maximum(np.abs(oenv[1:][me]), 1.0)) <= 1e-12


def test_sampler_runs_and_matches_oracle_for_a_short_run():
    prob, truth = synth.vegetation(nobs=158, growth="Growth_Yin1_temporal", var_gmax=True, nyear=5)
    K = 4
    start = np.tile(truth, (K, 1))
    sd = 0.02 * np.abs(start) + 1e-4
    g, o = BamGpu(prob), Oracle(prob)
    rows = g.fit(start, sd, nAdapt=5, nCycles=2, seed=21)
    orows = o.fit(start, sd, nAdapt=5, nCycles=2, seed=21)[0]
    assert rows.shape == orows.shape
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-11)


def test_unsupported_combinations_are_errors():
    prob, truth = synth.vegetation(nobs=200, growth="Growth_Yin2")
    bad = Problem("Vegetation", prob.X, prob.Y, Yu=prob.Yu, Xu=np.full(prob.X.shape, 0.01), partype=[0] * prob.ntheta,
                  priors_theta=[("FlatPrior", [])] * prob.ntheta, sigma_func=["Constant"] * 5,
                  priors_gamma=[("FlatPrior", [])] * 5, xtra_i=[VEG_GROWTH["Growth_Yin2"], 3, 100],
                  xtra_d=[5.0, 1000.0, 1e-5, 1e-10])
    with pytest.raises(RuntimeError):
        BamGpu(bad)
    empty = Problem("Vegetation", prob.X, prob.Y, Yu=prob.Yu, partype=[0] * (prob.ntheta + 1),
                    priors_theta=[("FlatPrior", [])] * (prob.ntheta + 1), sigma_func=["Constant"] * 5,
                    priors_gamma=[("FlatPrior", [])] * 5, xtra_i=[VEG_GROWTH["Growth_Yin2"], 4, 100],
                    xtra_d=[5.0, 1000.0, 1e-5, 1e-10])
    with pytest.raises(RuntimeError, match="empty year"):
        BamGpu(empty)
