"""GPU parity tests of the second-wave pointwise models against the CPU oracle:
SuspendedLoad, SWOT (hS / hSB), SGD, Recession_h, Orthorectification (no / radial distortion).
Same bar as the first wave: log-posterior 1e-12 relative, identical feasibility flags, prediction spaghetti and
envelopes 1e-12, short sampler runs trajectory-identical."""
import numpy as np
import pytest

from workloads import synth
from bam_b200.capi import BamGpu
from oracle.oracle_api import Oracle
from tests.test_gpu_logpost import check_parity
from tests.test_gpu_predict import cmp_env, cmp_spag

pytestmark = pytest.mark.gpu

CASES = [("SuspendedLoad", {}), ("SWOT", dict(variant="SWOT_hS")), ("SWOT", dict(variant="SWOT_hSB")), ("SGD", {}),
         ("Recession_h", {}), ("Orthorectification", dict(disto_npar=0)), ("Orthorectification", dict(disto_npar=2)),
         ("HydraulicControl_section", {}), ("Segmentation", dict(nS=3, option=1)), ("Segmentation", dict(nS=4, option=2)),
         ("Segmentation", dict(nS=1))]


@pytest.mark.parametrize("model,kw", CASES)
def test_logpost_prediction_sampler(model, kw):
    prob, truth = synth.second_wave(model, nobs=500, **kw)
    x = synth.feasible_starts(truth, 48, rel=0.01, seed=2)
    if model not in ("Recession_h", "Segmentation"):
        x[3, 0] = -abs(x[3, 0]) - 1.0  # K / theta1 <= 0 (or pixel size for Ortho below)
    if model == "Segmentation" and kw.get("nS", 3) > 1:
        nS = kw["nS"]
        x[3, nS] = 1e9 if kw["option"] == 1 else -1.0  # shift after the last time (:96) / negative duration (:94)
        x[4, nS] = 0.01                                # first segment holds fewer than nmin points (:102-104)
    if model == "Orthorectification":
        x[3] = truth
        x[3, 9] = -1e-6  # negative pixel size
    g, o = BamGpu(prob), Oracle(prob)
    assert g.nDpar == (20 if model == "Orthorectification" else 0)
    check_parity(prob, x)
    # prediction: propagate a small sample through the model on the calibration inputs
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.005 * rng.standard_normal((200, truth.size)))
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
    dor = [1] * prob.nY
    env, spag = g.predict(mc, truth, xs, 1, dor, seed=4, want_spag=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, dor, seed=4, want_spag=True, nthreads=8)
    if model == "Orthorectification":
        # pixel coordinates c = c0 - (f/pc) * (..)/(..) with c0 ~ 1000: a coordinate near 0 is a cancellation, so the
        # tolerance is relative to the image scale, not to |value|
        assert ((spag == -666.666) == (ospag == -666.666)).all()
        assert np.max(np.abs(spag - ospag) / np.maximum(np.abs(ospag), 1000.0)) <= 1e-12
        assert np.max(np.abs(env - oenv) / np.maximum(np.abs(oenv), 1000.0)) <= 1e-12
    else:
        cmp_spag(spag, ospag)
        cmp_env(env, oenv)
    # a short sampler run follows the oracle's trajectory
    s = np.tile(truth, (3, 1))
    sd = 0.01 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=4, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=4, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    # residual run
    Xt, Ys, feas = g.calibration_run(truth)
    oXt, oYs, ofeas = o.calibration_run(truth)
    assert feas == ofeas and np.allclose(Ys, oYs, rtol=1e-12, atol=1e-300)


def test_row_infeasibility_patterns():
    """per-row exits: SuspendedLoad h < t2 (:104), SWOT S < S0 (:137) and h <= h0 -> Q = 0 (:138), SGD h <= h0 (:100)"""
    for model, kw, col, val in (("SuspendedLoad", {}, 0, 0.0), ("SWOT", dict(variant="SWOT_hS"), 1, -1.0), ("SGD", {}, 0, 0.1),
                                ("HydraulicControl_section", {}, 0, 5.5)):  # stage above the bathymetry table (:84)
        prob, truth = synth.second_wave(model, nobs=60, **kw)
        X = np.array(prob.X)
        xs = [X[:, j].copy() for j in range(prob.nX)]
        xs[col][5] = val
        if model == "SWOT":
            xs[0][7] = 0.5  # below h0 = 1: discharge exactly zero, row feasible
        g, o = BamGpu(prob), Oracle(prob)
        mc = np.tile(truth, (40, 1))
        _, spag = g.predict(mc, truth, xs, 1, [0], want_env=False, want_spag=True)
        _, ospag = o.predict(mc, truth, xs, 1, [0], want_env=False, want_spag=True)
        cmp_spag(spag, ospag)
        assert (spag[0, 5] == -666.666).all()
        if model == "SWOT":
            assert (spag[0, 7] == 0.0).all()


def test_segmentation_and_hcs_prediction_uses_the_prediction_series():
    """whole-series feasibility of Segmentation is evaluated on the series the model is applied to (:96,102-104):
    a prediction grid that leaves a segment empty flags every replication"""
    prob, truth = synth.second_wave("Segmentation", nobs=200, nS=3, option=1)
    g, o = BamGpu(prob), Oracle(prob)
    mc = np.tile(truth, (50, 1))
    for grid in (np.linspace(0, 100, 40), np.linspace(0, 50, 40)):  # the second grid never reaches the 2nd shift
        env, spag = g.predict(mc, truth, [grid], 1, [0], want_spag=True)
        oenv, ospag = o.predict(mc, truth, [grid], 1, [0], want_spag=True)
        assert (spag == ospag).all() and (env == oenv).all()  # piecewise-constant outputs: exact
    assert (spag == -666.666).all() and (env == -666.666).all()


def test_baratinbac():
    """b-a-c parameterisation: per-parameter-set root finding for the activation stage of a REPLACED control
    (SolveContinuity cases I-VI), then the BaRatin discharge equation.  Derived parameters = (k_i, ktype_i)."""
    prob, truth = synth.baratin_bac(nobs=400)
    x = synth.feasible_starts(truth, 64, rel=0.01, seed=4)
    x[3, 1] = -1.0                  # a1 <= 0 (:266-268)
    x[4, 3] = 3.0                   # b2 above b3: the replacement root falls beyond its upper bound -> infeasible
    x[5, 5] = x[5, 2]               # c2 == c1: case I (explicit root)
    x[6, 3] = x[6, 0]               # b2 == b1: case II (explicit root)
    g, o = BamGpu(prob), Oracle(prob)
    assert g.nDpar == 6
    check_parity(prob, x)
    fx, feas, isnull, dpar = g.logpost(x, want_dpar=True)
    assert feas[0] == 1 and feas[3] == 0 and feas[4] == 0
    assert list(dpar[0, 3:]) == [0.0, 3.0, 0.0] or list(dpar[0, 3:]) == [0.0, 4.0, 0.0]  # ktype: addition, root case, addition
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.005 * rng.standard_normal((300, truth.size)))
    H = np.linspace(-1.5, 5.0, 120)
    env, spag = g.predict(mc, truth, [H], 1, [1], seed=4, want_spag=True)
    oenv, ospag = o.predict(mc, truth, [H], 1, [1], seed=4, want_spag=True, nthreads=8)
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)
    s = np.tile(truth, (3, 1))
    rows = g.fit(s, 0.01 * np.abs(s), nAdapt=4, nCycles=2, seed=5)
    orows = o.fit(s, 0.01 * np.abs(s), nAdapt=4, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)


def test_sfd():
    """twin-gauge rating curve: per-observation Newton for the transition stage, then one of two regimes"""
    prob, truth = synth.sfd(nobs=400)
    x = synth.feasible_starts(truth, 48, rel=0.005, seed=4)
    x[3, 0] = -1.0   # KB <= 0
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nState == 1
    check_parity(prob, x)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((200, truth.size)))
    h1 = np.linspace(0.5, 7.0, 80)
    h2 = np.full(80, 1.93)
    h2[5] = 9.0  # fall h1 - h2 - delta <= 0: infeasible row (:147)
    env, spag = g.predict(mc, truth, [h1, h2], 1, [1], seed=4, want_spag=True)
    oenv, ospag = o.predict(mc, truth, [h1, h2], 1, [1], seed=4, want_spag=True, nthreads=8)
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)
    assert (spag[0, 5] == -666.666).all()
    # both regimes occur along the rating curve
    _, det = g.predict(None, truth, [h1, h2], 0, [0], want_env=False, want_spag=True)
    qv = truth[0] * (h1 - truth[1]) ** truth[2] * np.sqrt(np.maximum(h1 - h2 - truth[6], 0) / truth[3])
    qc = truth[5] * (h1 - truth[4]) ** truth[7]
    ok = det[0, :, 0] != -666.666
    isv = np.isclose(det[0, ok, 0], qv[ok], rtol=1e-9)
    isc = np.isclose(det[0, ok, 0], qc[ok], rtol=1e-9)
    assert (isv | isc).all() and isv.any() and isc.any()
    # state spaghetti (BaM_Prediction :1071-1075): kappa follows the output, no structural-error noise on it; rows that
    # never solve for kappa keep undefRN, infeasible rows carry the unfeasible flag
    h1[9], h2[9] = truth[1] - 0.5, truth[1] - 1.5  # below h0 with a positive fall: Q = 0, kappa stays undefRN (:126,148)
    envs, spags = g.predict(mc, truth, [h1, h2], 1, [1], seed=4, want_spag=True, states=True)
    oenvs, ospags = o.predict(mc, truth, [h1, h2], 1, [1], seed=4, want_spag=True, nthreads=8, states=True)
    assert spags.shape == (2, 80, 200) and envs.shape == (2, 7, 80)
    cmp_spag(spags, ospags)
    cmp_env(envs, oenvs)
    assert (spags[1, 5] == -666.666).all() and (spags[1, 9] == -999999999.0).all()
    kap = spags[1][(spags[1] != -666.666) & (spags[1] != -999999999.0)]
    nsolved = int(((h1 - h2 - truth[6] > 0.02) & (h1 > truth[1] + 0.05) & (h2 > truth[4] + 0.05)).sum())
    assert kap.size >= nsolved * 190 and nsolved > 50 and (kap > 1.9).all() and (kap < 20.0).all()
    # the discharge switches regime exactly at kappa: rows with h1 <= kappa follow the variable-slope law
    _, dets = g.predict(None, truth, [h1, h2], 0, [0], want_env=False, want_spag=True, states=True)
    q, kp = dets[0, :, 0], dets[1, :, 0]
    okk = (kp != -666.666) & (kp != -999999999.0)
    assert np.allclose(q[okk & (h1 <= kp)], qv[okk & (h1 <= kp)], rtol=1e-9) and np.allclose(q[okk & (h1 > kp)], qc[okk & (h1 > kp)], rtol=1e-9)
    # residual run with the state column (GetSim_calibration)
    Xt, Ys, St, feas = g.calibration_run_states(truth)
    oXt, oYs, oSt, ofeas = o.calibration_run_states(truth)
    assert feas == ofeas and np.allclose(Ys, oYs, rtol=1e-12, atol=1e-300) and St.shape == (400, 1)
    assert ((St == -999999999.0) == (oSt == -999999999.0)).all() and ((St == -666.666) == (oSt == -666.666)).all()
    assert np.allclose(St, oSt, rtol=1e-12, atol=0)


@pytest.mark.parametrize("missing,shuffle", [(True, True), (False, True), (True, False)])
def test_mixture(missing, shuffle):
    """Mixture (Mixture_model.f90:74-135): rows interleaved across events exercise the pack() pairing of input and output
    rows; weights above one / negative / summing above one are infeasible parameter sets."""
    prob, truth = synth.mixture(nS=3, nEvt=6, nElt=9, missing=missing, shuffle=shuffle)
    nt = prob.ntheta
    x = synth.feasible_starts(truth, 40, rel=0.01, seed=2)
    x[3, 0] = 1.3                      # a weight above one: null under the Uniform(0,1) prior
    x[4, 0], x[4, 1] = 0.8, 0.7        # sum of an event's weights above one: infeasible model run (:112 / derived weight < 0)
    g, o = BamGpu(prob), Oracle(prob)
    assert g.nDpar == 6 and o.nDpar == 6
    check_parity(prob, x)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.005 * rng.standard_normal((150, truth.size)))
    mc[7, 0], mc[7, 1] = 0.8, 0.7      # an infeasible replication: its column of the spaghetti is flagged
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
    env, spag = g.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True, nthreads=8)
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)
    s = np.tile(truth, (3, 1))
    sd = 0.01 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=3, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    Xt, Ys, feas = g.calibration_run(truth)
    oXt, oYs, ofeas = o.calibration_run(truth)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-13, atol=0) and np.array_equal(Xt, oXt)
    assert nt == (3 * 6 + 9 if missing else 2 * 6)


def test_mixture_reference_workspace():
    """tests/BaM_Mixture: 148 weights + 2 remnant parameters, 351 rows (27 events x 13 elements)"""
    from tests.helpers import perturbed_vectors, problem_from_fixture

    prob, d = problem_from_fixture("mixture")
    x = perturbed_vectors(d["start"], 64, rel=0.02, seed=1)
    check_parity(prob, x)
    g, o = BamGpu(prob), Oracle(prob)
    mc = perturbed_vectors(d["start"], 300, rel=0.01, seed=2)
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
    env, spag = g.predict(mc, np.asarray(d["start"]), xs, 1, [1], seed=9, want_spag=True)
    oenv, ospag = o.predict(mc, np.asarray(d["start"]), xs, 1, [1], seed=9, want_spag=True, nthreads=8)
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)


def test_time_recursive_sfdtidal_qmec():
    """row f4: one sequential pass over the series per parameter vector (series_run), the pointwise kernels look the
    values up.  Log-posterior with gaps in the observations, prediction with the three state spaghetti, sampler, residual
    run -- all against the oracle."""
    prob, truth = synth.sfdtidal_qmec(nobs=500)
    x = synth.feasible_starts(truth, 40, rel=0.01, seed=2)
    x[3, 0] = -10.0   # A0 <= 0: infeasible vector (and null under its prior? no: Gaussian prior) -> likelihood infeasible
    x[4, 1] = 3.0     # bed above the water level: negative wetted area somewhere in the series (:131-133)
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nState == 3
    check_parity(prob, x, tol=1e-11)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((120, truth.size)))
    mc[5, 0] = -1.0  # an infeasible replication
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(2)]
    env, spag = g.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True, states=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True, nthreads=8, states=True)
    assert spag.shape == (4, 500, 120) and (spag[:, :, 5] == -666.666).all()
    cmp_spag(spag, ospag, tol=1e-10)
    cmp_env(env, oenv, tol=1e-10)
    # a sub-range of the inputs is a slice of the full prediction: the recursion always starts at the first input
    _, sfull = g.predict(mc, truth, xs, 1, [0], want_env=False, want_spag=True)
    _, s2 = g.predict(mc, truth, xs, 1, [0], want_env=False, want_spag=True, t0=100, t1=260)
    assert np.array_equal(s2[0], sfull[0, 100:260])
    s = np.tile(truth, (3, 1))
    sd = 0.005 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=3, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    Xt, Ys, St, feas = g.calibration_run_states(truth)
    oXt, oYs, oSt, ofeas = o.calibration_run_states(truth)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-11) and np.allclose(St, oSt, rtol=1e-10, atol=1e-12)


def test_sfdtidal_qmec2_reference_workspace():
    """tests/BaM_SFDTidal_Qmec2 (8000-step prefix, 128 gauged steps, 5 of 11 parameters fixed): the 8000-step recursion
    on the device against the oracle"""
    from tests.helpers import perturbed_vectors, problem_from_fixture

    prob, d = problem_from_fixture("sfdtidal_qmec2")
    x = perturbed_vectors(d["start"], 48, rel=0.01, seed=1)
    x[5, 0] = -3.0  # Be <= 0
    check_parity(prob, x, tol=1e-10)
    g, o = BamGpu(prob), Oracle(prob)
    x0 = np.asarray(d["start"])
    Xt, Ys, St, feas = g.calibration_run_states(x0)
    oXt, oYs, oSt, ofeas = o.calibration_run_states(x0)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-10, atol=1e-8) and np.allclose(St, oSt, rtol=1e-9, atol=1e-8)


def test_time_recursive_algae_biomass():
    """row f4: AlgaeBiomass on series_run (5 input series, 18 parameters, 2 outputs, 5 state variables): log-posterior with
    gaps in both outputs, prediction with the state spaghetti, sampler, residual run -- all against the oracle; input
    errors are refused at creation with the reference's message"""
    import copy

    from bam_b200.capi import BamError

    prob, truth = synth.algae_biomass(nobs=400)
    x = synth.feasible_starts(truth, 40, rel=0.01, seed=2)
    x[3, 0] = 2.0     # b0 > bmax: infeasible parameter set
    x[4, 4] = 40.0    # Topt >= Tmax
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nState == 5
    check_parity(prob, x, tol=1e-11)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((120, truth.size)))
    mc[5, 0] = 2.0  # an infeasible replication
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(5)]
    env, spag = g.predict(mc, truth, xs, 1, [1, 0], seed=4, want_spag=True, states=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, [1, 0], seed=4, want_spag=True, nthreads=8, states=True)
    assert spag.shape == (7, 400, 120) and (spag[:, :, 5] == -666.666).all()
    cmp_spag(spag, ospag, tol=1e-10)
    cmp_env(env, oenv, tol=1e-10)
    s = np.tile(truth, (3, 1))
    sd = 0.005 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=3, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    Xt, Ys, St, feas = g.calibration_run_states(truth)
    oXt, oYs, oSt, ofeas = o.calibration_run_states(truth)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-11) and np.allclose(St, oSt, rtol=1e-10, atol=1e-13)
    bad = copy.deepcopy(prob)
    bad.X[7, 1] = -1.0  # irradiance < 0
    with pytest.raises(BamError) as e:
        BamGpu(bad)
    assert "Irradiance<0" in str(e.value)


def _bias_problem():
    """Linear model, 2 inputs x 2 outputs, X and Y biases, missing Y values, no latent inputs."""
    from bam_b200.capi import Problem

    rng = np.random.default_rng(8)
    n = 333
    X = rng.uniform(1.0, 5.0, (n, 2))
    th = np.array([1.5, -0.7, 0.4, 2.0])
    Y = np.column_stack([X @ th[:2], X @ th[2:]]) + 0.1 * rng.standard_normal((n, 2))
    Y[5, 1] = -9999.0
    Y[300, 0] = -9999.0
    Xbindx = rng.integers(0, 3, (n, 2))
    Xb = np.where(Xbindx > 0, 0.02, 0.0)
    Ybindx = rng.integers(0, 3, (n, 2))
    Yb = np.where(Ybindx > 0, 0.05, 0.0)
    prob = Problem("Linear", X, Y, Yu=np.full((n, 2), 0.05), Xb=Xb, Xbindx=Xbindx, Yb=Yb, Ybindx=Ybindx, partype=[0] * 4,
                   priors_theta=[("Gaussian", [t, 1.0]) for t in th], sigma_func=["Linear", "Constant"],
                   priors_gamma=[("Uniform", [0, 10]), ("Uniform", [0, 10]), ("Uniform", [0, 10])])
    o = Oracle(prob)
    start = np.concatenate([th, [0.1, 0.02, 0.1], np.zeros(o.P - 7)])
    return prob, start


CHAIN_CASES = ["vegetation", "vegetation_var", "bias", "SWOT", "SuspendedLoad", "Orthorectification", "textfile", "sfdtidal",
               "mixture", "baratin_spd_yu0"]


@pytest.mark.parametrize("case", CHAIN_CASES)
def test_chain_parallel_generic_kernel(case):
    """logpost_chain (thread <-> chain, K >= 1024, no latent inputs): against the oracle, against the observation-parallel
    kernel, ragged last block, infeasible chains, and a chain's value independent of the chains it is launched with."""
    K = 1100
    tol = 1e-12
    if case == "vegetation":
        prob, truth = synth.vegetation(nobs=400, growth="Growth_Yin3", nY=5)
        x = synth.feasible_starts(truth, K, rel=0.002, seed=3)
    elif case == "vegetation_var":
        prob, truth = synth.vegetation(nobs=400, var_gmax=True)
        x = synth.feasible_starts(truth, K, rel=0.002, seed=3)
    elif case == "bias":
        prob, start = _bias_problem()
        rng = np.random.default_rng(6)
        x = np.tile(start, (K, 1)) * (1.0 + 0.01 * rng.standard_normal((K, start.size)))
        x[:, 7:] = 0.3 * rng.standard_normal((K, start.size - 7))
        x[9, 4] = -0.5  # gamma1 < 0 under a Uniform(0, 10) prior: null prior
    elif case == "textfile":
        from bam_b200.capi import Problem

        rng = np.random.default_rng(12)
        n = 450
        T, S = rng.uniform(5.0, 50.0, n), rng.uniform(0.5, 2.0, n)
        a, b, c = 2.0, 0.3, 1.4
        Yh = a * T**c + b * np.exp(-S) * np.log(T)
        Y = Yh * (1.0 + 0.05 * rng.standard_normal(n))
        text = "2\nT,S\n3\na,b,c\n1\na*T^c+b*exp(-S)*log(T)\n"
        prob = Problem("TextFile", np.column_stack([T, S]), Y, Yu=np.full(n, 0.5), partype=[0, 0, 0],
                       priors_theta=[("Gaussian", [0, 10])] * 3, sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 100])] * 2,
                       xtra_text=text)
        x = synth.feasible_starts(np.array([a, b, c, 0.5, 0.05]), K, rel=0.01, seed=3)
    elif case == "sfdtidal":
        prob, truth = synth.sfdtidal_qmec(nobs=300)
        x = synth.feasible_starts(truth, K, rel=0.01, seed=3)
        x[4, 1] = 3.0
        tol = 1e-11
    elif case == "mixture":
        prob, truth = synth.mixture()
        x = synth.feasible_starts(truth, K, rel=0.01, seed=3)
    elif case == "baratin_spd_yu0":
        prob, truth = synth.baratin_spd(nobs=700)
        prob.Yu[3, 0] = 0.0  # a zero Yu keeps the BaRatin fast path out: generic kernels
        x = synth.feasible_starts(truth, K, rel=0.01, seed=3, check=synth.spd_start_check)
    else:
        kw = dict(variant="SWOT_hSB") if case == "SWOT" else (dict(disto_npar=2) if case == "Orthorectification" else {})
        prob, truth = synth.second_wave(case, nobs=500, **kw)
        x = synth.feasible_starts(truth, K, rel=0.01, seed=3)
    if case not in ("bias", "mixture"):
        x[3, 0] = -abs(x[3, 0]) - 1.0
    check_parity(prob, x, tol=tol, floor=float(prob.nobs) * 1e-3)
    g, g2 = BamGpu(prob), BamGpu(prob)
    g2.set_option("force_generic", 1)
    a, b = g.logpost(x), g2.logpost(x)
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all()
    ok = (a[1] == 1) & (a[2] == 0)
    assert (np.abs(a[0][ok] - b[0][ok]) <= tol * np.maximum(np.abs(b[0][ok]), prob.nobs * 1e-3)).all()
    # batch independence: the tiling depends on the number of observations only
    sub = g.logpost(x[40:40 + 1024])
    assert (sub[0] == a[0][40:40 + 1024]).all()
    assert (g.logpost(x)[0] == a[0]).all()


def test_time_recursive_dynamic_vegetation():
    """row f4: DynamicVegetation on series_run (AlgaeBiomass budget on the clamped depth + a Newton solve per time step):
    log-posterior with gaps, prediction with the state spaghetti, sampler, residual run -- against the oracle; a Newton
    budget that cannot converge flags rows, which makes the vector infeasible on both sides."""
    import copy

    prob, truth = synth.dynamic_vegetation(nobs=400)
    x = synth.feasible_starts(truth, 40, rel=0.01, seed=2)
    x[3, 0] = 2.0     # b0 > bmax: infeasible parameter set
    x[4, 15] = 5.0    # b1 above every stage: depth 0 everywhere, Q = 0
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nState == 5
    check_parity(prob, x, tol=1e-11)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((120, truth.size)))
    mc[5, 0] = 2.0  # an infeasible replication
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(5)]
    env, spag = g.predict(mc, truth, xs, 1, [1, 1, 0], seed=4, want_spag=True, states=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, [1, 1, 0], seed=4, want_spag=True, nthreads=8, states=True)
    assert spag.shape == (8, 400, 120) and (spag[:, :, 5] == -666.666).all()
    cmp_spag(spag, ospag, tol=1e-10)
    cmp_env(env, oenv, tol=1e-10)
    s = np.tile(truth, (3, 1))
    sd = 0.005 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=3, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    Xt, Ys, St, feas = g.calibration_run_states(truth)
    oXt, oYs, oSt, ofeas = o.calibration_run_states(truth)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-10, atol=1e-13) and np.allclose(St, oSt, rtol=1e-10, atol=1e-13)
    tight = copy.deepcopy(prob)
    tight.xtra_i[0] = 2  # two evaluations: the bracket ends only
    fx, feas1, _ = BamGpu(tight).logpost(x[:8])[:3]
    ofx, ofeas1, _ = Oracle(tight).logpost(x[:8])[:3]
    assert (feas1 == ofeas1).all() and feas1[0] == 0 and feas1[4] == 1  # the all-dry vector needs no Newton solve


@pytest.mark.parametrize("model", ["SFDTidal_Qmec0", "SFDTidal_QmecMS"])
def test_time_recursive_sfdtidal_variants(model):
    """row f4: the original Qmec parameterisation and the steady Manning-Strickler version on series_run: log-posterior with
    gaps, prediction, sampler, residual run against the oracle; infeasible parameter sets and infeasible geometry."""
    prob, truth = synth.sfdtidal_variant(model, nobs=500)
    x = synth.feasible_starts(truth, 40, rel=0.01, seed=2)
    if model == "SFDTidal_Qmec0":
        x[3, 2] = 0.01   # dzeta >= 0
        x[4, 1] = 0.5    # effective depth so small that the wetted area turns negative at low tide
    else:
        x[3, 0] = -5.0   # A0 <= 0
        x[4, 1] = 3.0    # bed above the water level
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nState == 0
    check_parity(prob, x, tol=1e-11)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((120, truth.size)))
    mc[5, 0] = -1.0  # an infeasible replication
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(2)]
    env, spag = g.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True, nthreads=8)
    assert spag.shape == (1, 500, 120) and (spag[:, :, 5] == -666.666).all()
    cmp_spag(spag, ospag, tol=1e-10)
    cmp_env(env, oenv, tol=1e-10)
    s = np.tile(truth, (3, 1))
    sd = 0.005 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=3, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    Xt, Ys, feas = g.calibration_run(truth)
    oXt, oYs, ofeas = o.calibration_run(truth)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-11)


@pytest.mark.parametrize("model", ["SFDTidal", "SFDTidal2", "SFDTidal4", "SFDTidalJones"])
def test_sfdtidal_lagged_family(model):
    """row f4: the stage-fall-discharge curves that read the stages at a lagged time step (one device model id, four
    variants) on series_run: log-posterior with gaps, prediction, sampler, residual run against the oracle."""
    prob, truth = synth.sfdtidal_lag(model, nobs=500)
    x = synth.feasible_starts(truth, 40, rel=0.01, seed=2)
    x[3, 0] = -1.0                                  # K <= 0: infeasible parameter set
    x[4, 3 if model == "SFDTidal" else 2] = 10.0    # bed above every stage: Q = 0 everywhere, feasible
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nState == 0
    check_parity(prob, x, tol=1e-12)
    rng = np.random.default_rng(3)
    mc = truth * (1.0 + 0.003 * rng.standard_normal((120, truth.size)))
    mc[5, 0] = -1.0  # an infeasible replication
    xs = [np.asarray(prob.X)[:, j].copy() for j in range(prob.nX)]
    env, spag = g.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True)
    oenv, ospag = o.predict(mc, truth, xs, 1, [1], seed=4, want_spag=True, nthreads=8)
    assert spag.shape == (1, 500, 120) and (spag[:, :, 5] == -666.666).all()
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)
    _, s2 = g.predict(mc, truth, xs, 1, [0], want_env=False, want_spag=True, t0=100, t1=260)
    _, sfull = g.predict(mc, truth, xs, 1, [0], want_env=False, want_spag=True)
    assert np.array_equal(s2[0], sfull[0, 100:260])  # a sub-range still reads the lagged rows of the WHOLE series
    s = np.tile(truth, (3, 1))
    sd = 0.005 * np.abs(s) + 1e-12
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=5)
    orows = o.fit(s, sd, nAdapt=3, nCycles=2, seed=5)[0]
    assert np.allclose(rows, orows, rtol=1e-9, atol=1e-12)
    Xt, Ys, feas = g.calibration_run(truth)
    oXt, oYs, ofeas = o.calibration_run(truth)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-12)


@pytest.mark.parametrize("name", ["sfdtidal2", "sfdtidal4", "sfdtidal_jones"])
def test_lagged_sfdtidal_reference_workspaces(name):
    """tests/BaM_SFDTidal2, BaM_SFDTidal4, BaM_SFDTidal_Jones (the shipped workspaces whose parameter lists match the code):
    log-posterior around the configured starting point and the residual run on the device against the oracle"""
    from tests.helpers import perturbed_vectors, problem_from_fixture

    prob, d = problem_from_fixture(name)
    x = perturbed_vectors(d["start"], 48, rel=0.01, seed=1)
    x[5, 1] = -3.0  # K <= 0
    check_parity(prob, x, tol=1e-12)
    g, o = BamGpu(prob), Oracle(prob)
    x0 = np.asarray(d["start"])
    Xt, Ys, feas = g.calibration_run(x0)
    oXt, oYs, ofeas = o.calibration_run(x0)
    assert feas == ofeas == 1 and np.allclose(Ys, oYs, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name", ["hydraulic_control_section", "orthorectification", "regression_x2y1", "sediment", "textfilemodel",
                                  "baratin_qbias", "baratin_qbiashbias", "baratinbac", "baratinbac_spd", "suspendedload",
                                  "sediment_camenen_x2", "regression_catchments", "regression_x6", "segmentation"])
def test_remaining_shipped_workspaces(name):
    """the other shipped workspaces of models on the device, on the device against the oracle around their configured
    starting points"""
    from tests.helpers import perturbed_vectors, problem_from_fixture

    prob, d = problem_from_fixture(name)
    # (the b-a-c parameterisation with period-specific offsets is feasible in a narrow band around its starting point only)
    x = perturbed_vectors(d["start"], 48, rel=0.0005 if name == "baratinbac_spd" else 0.01, seed=1)
    check_parity(prob, x, tol=1e-12, floor=float(prob.nobs) * 1e-3)
