"""GPU parity on the edges of the domain: empty and ragged data sets, sizes that straddle the kernels' tiles, single
chains / single replications, sub-ranges of a prediction."""
import numpy as np
import pytest

from workloads import synth
from bam_b200.capi import BamGpu, Problem
from oracle.oracle_api import Oracle
from tests.helpers import perturbed_vectors, problem_from_fixture
from tests.test_gpu_logpost import check_parity
from tests.test_gpu_predict import cmp_env, cmp_spag

pytestmark = pytest.mark.gpu


def _baratin(n, seed=1, missing=None):
    prob, truth = synth.baratin_2c(nobs=max(n, 1), seed=seed)
    X, Y, Yu = np.asarray(prob.X)[:n], np.asarray(prob.Y)[:n].copy(), np.asarray(prob.Yu)[:n]
    if missing is not None:
        Y[missing] = -9999.0
    priors = [("Gaussian", [-0.5, 0.25]), ("Gaussian", [53, 10]), ("Gaussian", [1.5, 0.025]), ("Gaussian", [1.5, 0.5]),
              ("Gaussian", [144, 45]), ("Gaussian", [1.67, 0.025])]  # tests/BaM_BaRatin
    p = Problem("BaRatin", X.reshape(n, 1), Y.reshape(n, 1), Yu=Yu.reshape(n, 1), partype=[0] * 6, priors_theta=priors,
                sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000]), ("Uniform", [0, 1000])], xtra_i=[1, 0, 0, 1])
    return p, truth


def test_empty_data_set_is_the_prior():
    """nobs = 0: the posterior is the prior (the likelihood loop of GetLogLikelihood :3204-3227 runs zero times)"""
    prob, truth = _baratin(0)
    x = synth.feasible_starts(truth, 7, rel=0.02, seed=3)
    g, o = BamGpu(prob), Oracle(prob)
    fx, feas, isnull = g.logpost(x)
    ofx, ofeas, oisnull = o.logpost(x)
    pr, lk, hy, _, _ = g.logpost_parts(x)
    assert (feas == ofeas).all() and (isnull == oisnull).all() and feas.all()
    assert np.allclose(fx, ofx, rtol=1e-13) and (lk == 0.0).all() and (hy == 0.0).all() and np.allclose(pr, fx, rtol=0, atol=0)
    rows = g.fit(x[:3], 0.01 * np.abs(x[:3]), nAdapt=3, nCycles=2, seed=2)
    orows = o.fit(x[:3], 0.01 * np.abs(x[:3]), nAdapt=3, nCycles=2, seed=2)[0]
    assert np.allclose(rows, orows, rtol=1e-11, atol=1e-13)


def test_all_outputs_missing():
    """every Y is the missing-value code: no likelihood term at all (:3207), the model's feasibility still counts"""
    prob, truth = _baratin(40, missing=slice(None))
    x = synth.feasible_starts(truth, 9, rel=0.02, seed=4)
    x[2, 1] = -1.0  # a1 <= 0: infeasible parameter set even without data terms
    g, o = BamGpu(prob), Oracle(prob)
    fx, feas, isnull = g.logpost(x)
    ofx, ofeas, oisnull = o.logpost(x)
    assert (feas == ofeas).all() and (isnull == oisnull).all() and feas[2] == 0
    ok = feas == 1
    assert np.allclose(fx[ok], ofx[ok], rtol=1e-13)


@pytest.mark.parametrize("n", [1, 31, 33, 255, 257, 513])
@pytest.mark.parametrize("K", [1, 3, 129])
def test_sizes_across_tile_boundaries(n, K):
    """observation counts around the warp / block / tile sizes, chain counts around the fast-path threshold (K >= 128),
    a ragged set of missing outputs"""
    miss = np.arange(n) % 5 == 3 if n > 4 else None
    prob, truth = _baratin(n, seed=n, missing=miss)
    x = synth.feasible_starts(truth, K, rel=0.01, seed=n + K)
    check_parity(prob, x)


def test_latent_fast_path_ragged_sizes():
    """cfg-3 fast kernels at sizes that are not multiples of their tiles (4 chains x 256 threads x 16 observations)"""
    for n, K in ((1, 1), (255, 5), (4097, 3), (4096 + 300, 9)):
        prob, truth = synth.linear_latent(nobs=n, seed=n)
        rng = np.random.default_rng(n)
        x = np.tile(truth, (K, 1)) + 0.01 * rng.standard_normal((K, truth.size))
        check_parity(prob, x)


def test_prediction_single_replication_and_subranges():
    prob, d = problem_from_fixture("baratin")
    mc, mode, Hs = synth.baratin_spaghetti(Nrep=3000, nobs=97)
    g, o = BamGpu(prob), Oracle(prob)
    # Nrep = 1 (maxpost run): spaghetti only, no envelope (BaM_Message 16)
    _, s1 = g.predict(None, mode, [Hs], 0, [0], want_env=False, want_spag=True)
    _, o1 = o.predict(None, mode, [Hs], 0, [0], want_env=False, want_spag=True)
    assert s1.shape == (1, 97, 1)
    cmp_spag(s1, o1)
    # Nrep = 2: the smallest sample with an envelope
    e2, s2 = g.predict(mc[:2], mode, [Hs], 1, [1], seed=3, want_spag=True)
    oe2, os2 = o.predict(mc[:2], mode, [Hs], 1, [1], seed=3, want_spag=True)
    cmp_spag(s2, os2)
    cmp_env(e2, oe2)
    # a single input, and sub-ranges [t0, t1) that do not start on a multiple of four inputs (noise quads)
    for t0, t1 in ((0, 1), (5, 6), (3, 50), (94, 97)):
        e, s = g.predict(mc, mode, [Hs], 1, [1], seed=9, t0=t0, t1=t1, want_spag=True)
        oe, os_ = o.predict(mc, mode, [Hs], 1, [1], seed=9, t0=t0, t1=t1, want_spag=True, nthreads=8)
        cmp_spag(s, os_)
        cmp_env(e, oe)
    # the sub-range is a slice of the full prediction (noise keyed by the global input index)
    ef, sf = g.predict(mc, mode, [Hs], 1, [1], seed=9, want_spag=True)
    e, s = g.predict(mc, mode, [Hs], 1, [1], seed=9, t0=3, t1=50, want_spag=True)
    assert np.array_equal(s, sf[:, 3:50]) and np.array_equal(e, ef[:, :, 3:50])


def test_first_failing_prior_decides_the_flag():
    """GetPriorLogPdf leaves at the FIRST parameter (remnant parameters, then theta) that is infeasible (invalid prior
    parameters) or has zero density: a null gamma before an infeasible theta prior is "null", an infeasible gamma prior
    before a null theta is "infeasible".  The warp-parallel prior kernel reproduces the order."""
    prob0, truth = synth.baratin_2c(nobs=50)
    X, Y, Yu = np.asarray(prob0.X), np.asarray(prob0.Y), np.asarray(prob0.Yu)
    good = [("Gaussian", [-0.5, 0.25]), ("Gaussian", [53, 10]), ("Gaussian", [1.5, 0.025]), ("Uniform", [0.5, 2.5]),
            ("Gaussian", [144, 45]), ("Gaussian", [1.67, 0.025])]
    bad_theta = list(good)
    bad_theta[4] = ("Gaussian", [144, -1.0])  # invalid standard deviation: infeasible whatever the value
    x = synth.feasible_starts(truth, 6, rel=0.001, seed=1)
    x[1, 6] = -1.0   # gamma1 outside Uniform(0,1000): null, and gamma comes first
    x[2, 3] = 9.0    # k2 outside its Uniform prior: null at theta 4 (before the invalid prior of theta 5)
    for priors_theta, priors_gamma in ((bad_theta, [("Uniform", [0, 1000])] * 2),
                                       (good, [("Gaussian", [1.0, -2.0]), ("Uniform", [0, 1000])])):
        p = Problem("BaRatin", X, Y, Yu=Yu, partype=[0] * 6, priors_theta=priors_theta, sigma_func=["Linear"],
                    priors_gamma=priors_gamma, xtra_i=[1, 0, 0, 1])
        g, o = BamGpu(p), Oracle(p)
        fx, feas, isnull = g.logpost(x)
        ofx, ofeas, oisnull = o.logpost(x)
        assert (feas == ofeas).all() and (isnull == oisnull).all()
        assert (fx == -999999999.0).all()
    # first problem: vector 1 null (gamma first), vector 2 null (theta 4 before theta 5), the others infeasible
    p = Problem("BaRatin", X, Y, Yu=Yu, partype=[0] * 6, priors_theta=bad_theta, sigma_func=["Linear"],
                priors_gamma=[("Uniform", [0, 1000])] * 2, xtra_i=[1, 0, 0, 1])
    fx, feas, isnull = BamGpu(p).logpost(x)
    assert list(isnull) == [0, 1, 1, 0, 0, 0] and list(feas) == [0, 1, 1, 0, 0, 0]


def test_missing_output_rows_keep_their_feasibility_test():
    """ApplyModel takes feas = all(vfeas) over ALL rows (ModelLibrary_tools.f90:421-423): a row whose Y is missing still
    makes the vector infeasible when the model is infeasible there (Sediment: an input <= 0, :155), although it adds no
    likelihood term (:3207).  Only BaRatin / Linear rows -- which cannot be infeasible -- are dropped from the device."""
    import copy

    prob, truth = synth.sediment_camenen(nobs=60)
    ok_prob = copy.deepcopy(prob)
    ok_prob.Y[[5, 17], 0] = -9999.0
    bad_prob = copy.deepcopy(ok_prob)
    bad_prob.X[17, 0] = -0.1  # water depth <= 0 in a row with missing Y
    x = synth.feasible_starts(truth, 5, rel=0.01, seed=6)
    for p, expect in ((ok_prob, 1), (bad_prob, 0)):
        g, o = BamGpu(p), Oracle(p)
        fx, feas, isnull = g.logpost(x)
        ofx, ofeas, oisnull = o.logpost(x)
        assert (feas == ofeas).all() and (isnull == oisnull).all() and (feas == expect).all()
        if expect:
            assert np.allclose(fx, ofx, rtol=1e-12)


def test_derived_value_limit_is_checked_at_creation():
    """Segmentation with nS = 70 segments has 69 derived values: beyond the compiled limit (64), refused loudly"""
    from bam_b200.capi import BamError

    nS, n = 70, 400
    t = np.arange(n, dtype=np.float64)
    y = np.zeros(n)
    theta_pri = [("Gaussian", [0, 10])] * nS + [("Uniform", [0, float(n)])] * (nS - 1)
    p = Problem("Segmentation", t.reshape(n, 1), y.reshape(n, 1), Yu=np.full((n, 1), 0.1), partype=[0] * (2 * nS - 1),
                priors_theta=theta_pri, sigma_func=["Constant"], priors_gamma=[("Uniform", [0, 100])],
                xtra_i=[nS, 1, 1], xtra_d=[1.0])
    with pytest.raises(BamError) as e:
        BamGpu(p)
    assert "compiled limits" in str(e.value)
