"""GPU tests of the batched adaptive-Metropolis sampler (K3) and the moments / Gelman-Rubin (K5).

(c) of BASELINE.json's north_star: the sampler is stochastic -> statistical agreement.  Because the
device and the oracle share the RNG specification (Philox counters keyed by chain / iteration /
component), short runs can also be compared trajectory by trajectory."""
import numpy as np
import pytest

from workloads import synth
from bam_b200.capi import BamGpu
from oracle.oracle_api import Oracle
from tests.helpers import problem_from_fixture

pytestmark = pytest.mark.gpu


def start_std(start, K, factor=0.1):
    s = np.tile(np.asarray(start, dtype=np.float64), (K, 1))
    sd = np.where(s != 0, factor * np.abs(s), 0.1)  # Config_Read_MCMC default (:2459-2471)
    return s, sd


def test_short_trajectories_match_oracle():
    prob, d = problem_from_fixture("baratin")
    K = 6
    s, sd = start_std(d["start"], K)
    g, o = BamGpu(prob), Oracle(prob)
    rows = g.fit(s, sd, nAdapt=10, nCycles=4, seed=11)
    orow, ofx, ostd = o.fit(s, sd, nAdapt=10, nCycles=4, seed=11)
    assert rows.shape == orow.shape == (K, 40, 8 + 1 + 2)
    # identical accept/reject decisions -> identical states up to libm/libdevice rounding of fx
    assert np.allclose(rows[:, :, :8], orow[:, :, :8], rtol=1e-12, atol=0)
    assert np.allclose(rows[:, :, 8], orow[:, :, 8], rtol=1e-12)
    assert np.allclose(rows[:, :, 9:], orow[:, :, 9:], rtol=1e-12)
    # chains differ from each other (independent streams) and moved
    assert not np.allclose(rows[0, -1, :8], rows[1, -1, :8])
    fx, feas, isnull, dpar = g.chains_download(want_dpar=True)
    assert np.allclose(fx, orow[:, -1, 8], rtol=1e-12) and feas.all()
    cnt, mean, m2 = g.moments()
    assert (cnt == 40).all()
    assert np.allclose(mean, rows[:, :, :8].mean(axis=1), rtol=1e-11)
    assert np.allclose(m2, ((rows[:, :, :8] - mean[:, None, :]) ** 2).sum(axis=1), rtol=1e-8, atol=1e-14)


def test_burn_and_thinning_and_var_params():
    prob, d = problem_from_fixture("baratin_spd", with_copula=True)
    K = 4
    s, sd = start_std(d["start"], K, 0.05)
    g, o = BamGpu(prob), Oracle(prob)
    rows = g.fit(s, sd, nAdapt=5, nCycles=4, seed=3, burn=6, keep_every=3)
    orow, _, _ = o.fit(s, sd, nAdapt=5, nCycles=4, seed=3, burn=6, keep_every=3)
    assert rows.shape == orow.shape == (K, 4, 19 + 1 + 15)
    assert np.allclose(rows, orow, rtol=1e-11, atol=1e-13)


def test_posterior_statistics_and_rhat():
    """Statistical agreement: many short GPU chains vs. the oracle sampler on the same posterior."""
    prob, truth = synth.baratin_2c(nobs=60, seed=4)
    K = 256
    s, sd = start_std(truth, K, 0.05)
    g, o = BamGpu(prob), Oracle(prob)
    rows = g.fit(s, sd, nAdapt=50, nCycles=40, seed=5, burn=1000, keep_every=1)
    cnt, mean, m2 = g.moments()
    rhat = g.rhat(cnt, mean, m2)
    assert (rhat < 1.2).all(), rhat
    gm = rows[:, :, :8].reshape(-1, 8)
    orow, _, _ = o.fit(s[:24], sd[:24], nAdapt=50, nCycles=40, seed=99, burn=1000, keep_every=1)
    om = orow[:, :, :8].reshape(-1, 8)
    # posterior means agree within a few Monte-Carlo standard errors of the smaller (oracle) sample
    se = om.std(axis=0) / np.sqrt(24 * 1000 / 50.0)  # effective sample size: ~1 independent draw per 50 iterations
    assert (np.abs(gm.mean(axis=0) - om.mean(axis=0)) < 6 * se).all()
    for q in (0.16, 0.5, 0.84):
        assert (np.abs(np.quantile(gm, q, axis=0) - np.quantile(om, q, axis=0)) < 8 * se).all()


def test_infeasible_start_is_an_error():
    from bam_b200.capi import BamError

    prob, d = problem_from_fixture("baratin")
    s, sd = start_std(d["start"], 3)
    s[1, 3] = -2.0
    with pytest.raises(BamError) as e:
        BamGpu(prob).fit(s, sd, nAdapt=2, nCycles=1)
    assert "unfeasible starting point" in str(e.value)


def test_chain_offset_makes_sharding_invisible():
    """Chains sharded over ranks: rank r runs chains [r*K/2, (r+1)*K/2) with chain offset -> same rows."""
    prob, d = problem_from_fixture("baratin")
    K = 8
    s, sd = start_std(d["start"], K)
    g = BamGpu(prob)
    rows = g.fit(s, sd, nAdapt=6, nCycles=2, seed=21)
    g2 = BamGpu(prob)
    g2.lib.bamgpu_set_chain_offset(g2.h, 4)
    rows_b = g2.fit(s[4:], sd[4:], nAdapt=6, nCycles=2, seed=21)
    assert (rows_b == rows[4:]).all()


def test_persistent_and_launch_per_proposal_paths_agree():
    """Small data sets run the persistent warp-per-chain kernel, large ones one launch per proposal:
    same RNG streams and accept rule -> the same chains (up to the rounding of differently ordered sums)."""
    prob, d = problem_from_fixture("baratin_spd", with_copula=True)
    K = 5
    s, sd = start_std(d["start"], K, 0.05)
    g = BamGpu(prob)
    rows_p = g.fit(s, sd, nAdapt=6, nCycles=3, seed=13)
    cnt_p, mean_p, m2_p = g.moments()
    g2 = BamGpu(prob)
    g2.set_option("force_generic", 1)
    rows_g = g2.fit(s, sd, nAdapt=6, nCycles=3, seed=13)
    cnt_g, mean_g, m2_g = g2.moments()
    assert g.launch_count() < g2.launch_count() / 10
    assert np.allclose(rows_p, rows_g, rtol=1e-11, atol=1e-13)
    assert np.allclose(mean_p, mean_g, rtol=1e-11) and (cnt_p == cnt_g).all()
    assert np.allclose(m2_p, m2_g, rtol=1e-7, atol=1e-13)
    fx_p = g.chains_download()[0]
    assert np.allclose(fx_p, rows_p[:, -1, 19], rtol=1e-12)


def _latent_problem(n=40, seed=3):
    """Linear 2x2 with input errors on most cells, an X bias on the others (cfg 3 shape, small)."""
    from bam_b200.capi import Problem

    rng = np.random.default_rng(seed)
    prob0, truth0 = synth.linear_latent(nobs=n, seed=seed)
    Xu = np.array(prob0.Xu)
    Xu[::7, 1] = 0.0  # a few cells without random error
    Xbindx = np.zeros((n, 2), dtype=np.int32)
    Xbindx[::7, 1] = 1
    Xb = np.where(Xbindx > 0, 0.05, 0.0)
    prob = Problem("Linear", prob0.X, prob0.Y, Yu=prob0.Yu, Xu=Xu, Xb=Xb, Xbindx=Xbindx, partype=[0, 0, 0, 0],
                   priors_theta=[("Gaussian", [0, 10])] * 4, sigma_func=["Linear", "Linear"],
                   priors_gamma=[("Uniform", [0, 100])] * 4)
    lat = np.asarray(prob0.X).flatten(order="F")[(Xu > 0).flatten(order="F")]
    start = np.concatenate([truth0[:8], lat, [0.0]])
    sd = np.concatenate([0.05 * np.abs(truth0[:8]), Xu.flatten(order="F")[(Xu > 0).flatten(order="F")], [0.1]])
    return prob, start, sd


def test_latent_sweep_equals_full_posterior_oaat():
    """cfg 3: the one-launch sweep over all latent inputs uses the LOCAL posterior difference of one observation;
    the oracle runs the reference's one-at-a-time loop with a FULL posterior evaluation per component.  Same RNG
    streams -> same accept / reject decisions -> same chains (up to the rounding of the differently summed fx)."""
    prob, start, sd = _latent_problem()
    K = 5
    s = np.tile(start, (K, 1))
    sdev = np.tile(sd, (K, 1))
    g, o = BamGpu(prob), Oracle(prob)
    assert g.sizes.nLatent == start.size - 9 and g.sizes.nXbias == 1
    rows = g.fit(s, sdev, nAdapt=6, nCycles=3, seed=17)
    orow, _, ostd = o.fit(s, sdev, nAdapt=6, nCycles=3, seed=17)
    assert rows.shape == orow.shape
    P = start.size
    assert np.allclose(rows[:, :, :P], orow[:, :, :P], rtol=1e-10, atol=1e-12)
    assert np.allclose(rows[:, :, P], orow[:, :, P], rtol=1e-10)
    # the latent inputs really moved, and the launch-per-proposal path gives the same chains
    assert (np.abs(rows[:, -1, 8:P - 1] - s[:, 8:P - 1]) > 0).mean() > 0.5
    g2 = BamGpu(prob)
    g2.set_option("force_generic", 1)
    rows2 = g2.fit(s, sdev, nAdapt=6, nCycles=3, seed=17)
    assert np.allclose(rows, rows2, rtol=1e-10, atol=1e-12)
    assert g.launch_count() < g2.launch_count() / 5


def test_latent_fast_paths_follow_the_oracle():
    """cfg 3 shape (Linear 2x2, EVERY input cell latent, no biases): the specialised log-posterior kernel
    (logpost_linear_latent) and the specialised sweep (latent_sweep_linear, table-driven Box-Muller and log u) give the
    oracle's chains, and the generic sweep's."""
    prob, truth = synth.linear_latent(nobs=300)
    K = 6
    s = np.tile(truth, (K, 1))
    sd = np.tile(np.concatenate([0.01 * np.abs(truth[:8]), np.full(600, 0.1)]), (K, 1))
    g, o = BamGpu(prob), Oracle(prob)
    rows = g.fit(s, sd, nAdapt=5, nCycles=3, seed=23)
    orow, _, _ = o.fit(s, sd, nAdapt=5, nCycles=3, seed=23)
    P = truth.size
    assert np.allclose(rows[:, :, :P], orow[:, :, :P], rtol=1e-10, atol=1e-12)
    assert np.allclose(rows[:, :, P], orow[:, :, P], rtol=1e-10)
    assert (rows[:, -1, 8:P] != s[:, 8:]).mean() > 0.5
    g2 = BamGpu(prob)
    g2.set_option("no_fast_sweep", 1)
    rows2 = g2.fit(s, sd, nAdapt=5, nCycles=3, seed=23)
    assert np.allclose(rows, rows2, rtol=1e-10, atol=1e-12)
    # log-posterior of perturbed vectors: fast kernel against the oracle, flags included (a negative gamma1)
    rng = np.random.default_rng(2)
    x = s + 0.02 * rng.standard_normal(s.shape)
    x[2, 4] = -5.0  # gamma1 of output 1 outside its Uniform(0,100) prior -> null
    x[3, 5] = -50.0  # gamma2 so negative that sigma <= 0 on every row -> infeasible likelihood
    from tests.test_gpu_logpost import check_parity
    check_parity(prob, x)


def test_latent_sweep_textfile_and_moments():
    prob, truth = synth.textfile_latent(nobs=300)
    K = 8
    s = np.tile(truth, (K, 1))
    sd = np.concatenate([0.02 * np.abs(truth[:4]), np.full(300, 1.0)])
    g = BamGpu(prob)
    rows = g.fit(s, np.tile(sd, (K, 1)), nAdapt=20, nCycles=5, seed=5)
    assert np.isfinite(rows).all()
    cnt, mean, m2 = g.moments()
    assert (cnt == 100).all() and np.allclose(mean, rows[:, :, :304].mean(axis=1), rtol=1e-10, atol=1e-12)
    # latent inputs stay within a few input standard deviations of the observed inputs
    z = (rows[:, -1, 4:304] - np.asarray(prob.X)[:, 0]) / 1.0
    assert np.abs(z).max() < 6 and z.std() > 0.2


def test_partial_reevaluation_of_var_parameters():
    """cfg 2 shape (VAR k1 / k2 over 5 periods, 4096+ observations, >= 128 chains): a proposal on a VAR parameter
    only re-evaluates the observations of its period and re-uses the cached per-combination sums of the others.
    The chains must equal the oracle's, which evaluates the full posterior for every proposal."""
    prob, truth = synth.baratin_spd(nobs=6000, copula=True)
    K = 128
    s = synth.feasible_starts(truth, K, rel=0.005, seed=9, check=synth.spd_start_check)
    sd = 0.002 * np.abs(s) + 1e-6
    g, o = BamGpu(prob), Oracle(prob)
    rows = g.fit(s, sd, nAdapt=3, nCycles=2, seed=31)
    orow, _, _ = o.fit(s, sd, nAdapt=3, nCycles=2, seed=31)
    P = g.P
    assert np.allclose(rows[:, :, :P], orow[:, :, :P], rtol=1e-10, atol=1e-13)
    assert np.allclose(rows[:, :, P], orow[:, :, P], rtol=1e-11)
    moved = (rows[:, -1, :P] != s).mean()
    assert moved > 0.3
    # and equal to the generic launch-per-proposal path (no fast kernel, no partial evaluation)
    g2 = BamGpu(prob)
    g2.set_option("force_generic", 1)
    rows2 = g2.fit(s, sd, nAdapt=3, nCycles=2, seed=31)
    assert np.allclose(rows, rows2, rtol=1e-10, atol=1e-13)
    # the final state's log-posterior equals a fresh full evaluation
    fx, feas, isnull = g.chains_download()
    fx2, _, _ = BamGpu(prob).logpost(rows[:, -1, :P])
    assert np.allclose(fx, fx2, rtol=1e-13) and feas.all()


def test_launch_per_proposal_sampler_on_the_chain_parallel_kernel():
    """n > 4096 keeps a pointwise model out of the persistent warp sampler; with K >= 1024 chains every proposal is then
    evaluated by logpost_chain (thread = chain): same chains as with the observation-parallel kernel, and the first chains
    follow the oracle."""
    from workloads import synth

    prob, truth = synth.second_wave("SWOT", nobs=5000, variant="SWOT_hSB")
    K = 1024
    s = synth.feasible_starts(truth, K, rel=0.005, seed=4)
    sd = 0.002 * np.abs(s) + 1e-9
    g, g2 = BamGpu(prob), BamGpu(prob)
    g2.set_option("force_generic", 1)
    rows = g.fit(s, sd, nAdapt=3, nCycles=1, seed=6)
    rows2 = g2.fit(s, sd, nAdapt=3, nCycles=1, seed=6)
    assert rows.shape == rows2.shape and np.allclose(rows, rows2, rtol=1e-9, atol=1e-12)
    o = Oracle(prob)
    orows = o.fit(s[:4], sd[:4], nAdapt=3, nCycles=1, seed=6)[0]
    assert np.allclose(rows[:4], orows, rtol=1e-9, atol=1e-12)
