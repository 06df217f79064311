"""GPU parity tests of the prediction path (K4 + envelopes) against the CPU oracle.

(b) of BASELINE.json's north_star: deterministic propagation of a fixed sample agrees within 1e-12
relative; integer selections (control activation / range, X-spaghetti column, infeasible flags) are
bit-exact -- checked through the pattern of zeros and of unfeasFlag values, which must be identical.
The structural-error noise is a pure function of (seed, rep, t, y) on both sides (Philox4x32-10)."""
import numpy as np
import pytest

from workloads import synth
from bam_b200.capi import BamGpu
from oracle.oracle_api import Oracle
from tests.helpers import load_fixture, problem_from_fixture

pytestmark = pytest.mark.gpu
FLAG = -666.666


def cmp_spag(a, b, tol=1e-12):
    assert a.shape == b.shape
    assert ((a == FLAG) == (b == FLAG)).all(), "infeasible pattern differs"
    assert ((a == 0.0) == (b == 0.0)).all(), "Q == 0 pattern (range / control activation) differs"
    m = a != FLAG
    # PER VALUE: relative error <= tol wherever the value is not a near-cancellation of model output and noise
    # (|value| >= 1e-2 x the column's median magnitude: cancellation amplifies the few-1e-16 rounding of the two terms by at
    # most 1e2, still below tol); values that did cancel (Y + sigma N(0,1) ~ 0) are held to a few ulps of the TERMS,
    # 4e-15 x the column scale -- 250 x tighter than tol x column scale, which is what this check used to allow everywhere
    bb = np.where(b == FLAG, np.nan, np.abs(b))
    colscale = np.maximum(np.nan_to_num(np.nanmedian(bb, axis=-1, keepdims=True), nan=1.0) * np.ones_like(b), 1e-300)
    d = np.abs(a - b)
    big = m & (np.abs(b) >= 1e-2 * colscale)
    small = m & ~big
    if big.any():
        rel = d[big] / np.abs(b[big])
        assert rel.max() <= tol, ("per-value relative error", rel.max())
    if small.any():
        ab = d[small] / colscale[small]
        # (a caller that loosens tol -- the time-recursive models, whose rounding accumulates along the series and whose
        # discharge crosses zero with the tide -- gets the same factor 1e-2 x tol in absolute terms)
        lim = 4e-15 if tol <= 1e-12 else 1e-2 * tol
        assert ab.max() <= lim, ("cancelled values, error in units of the column scale", ab.max())


def cmp_env(a, b, tol=1e-12):
    """envelopes [output, statistic, input]: per value, relative; the MEAN (statistic 5) of a column is a sum that can
    cancel (zero-mean noise around Q = 0), so its error is measured against max(|mean|, standard deviation)"""
    assert ((a == FLAG) == (b == FLAG)).all()
    m = b != FLAG
    scale = np.maximum(np.abs(b), 1e-6)
    if b.ndim == 3 and b.shape[1] == 7:
        sd = np.where(b[:, 6] == FLAG, 0.0, np.abs(b[:, 6]))
        scale[:, 5] = np.maximum(scale[:, 5], sd)
    assert (np.abs(a[m] - b[m]) / scale[m]).max() <= tol


def baratin_case(Nrep, seed=3):
    prob, d = problem_from_fixture("baratin")
    rng = np.random.default_rng(seed)
    start = np.array(d["start"])
    mc = start * (1.0 + 0.03 * rng.standard_normal((Nrep, start.size)))
    mc[:, 6:] = np.abs(mc[:, 6:])
    mc[5, 3] = -2.0  # k2 < k1: the whole replication is infeasible -> a row of flags
    return prob, mc, start


def test_baratin_workspace_experiments():
    """The 4 posterior experiments of tests/BaM_BaRatin/Config_Pred_*.txt: grid with parametric U,
    grid with total U, limnigraph maxpost, noisy limnigraph (100 X-spaghettis) with total U."""
    prob, mc, mode = baratin_case(500)
    pin = load_fixture("baratin_pred_inputs")
    hgrid = np.asarray(pin["Hgrid"])
    limni = np.asarray(pin["Limni"])
    noisy = np.asarray(pin["Limni_noisy_colmajor"]).reshape(pin["Limni_noisy_shape"][::-1]).T
    g, o = BamGpu(prob), Oracle(prob)
    for xs, dop, dor in ((hgrid, 1, [0]), (hgrid, 1, [1]), (noisy, 1, [1]), (limni, 0, [1])):
        env, spag = g.predict(mc, mode, [xs], dop, dor, seed=42, want_spag=True)
        oenv, ospag = o.predict(mc, mode, [xs], dop, dor, seed=42, want_spag=True, nthreads=8)
        cmp_spag(spag, ospag)
        cmp_env(env, oenv)
    # maxpost run: Nrep = 1, no envelope (src/BaM_tools.f90:958-960,1382)
    _, spag = g.predict(None, mode, [limni], 0, [0], want_env=False, want_spag=True)
    _, ospag = o.predict(None, mode, [limni], 0, [0], want_env=False, want_spag=True)
    cmp_spag(spag, ospag)


def test_input_sharding_is_invisible():
    """Prediction inputs sharded across ranks ([t0,t1) slices) give bitwise the same columns."""
    prob, mc, mode = baratin_case(300)
    H = np.linspace(-1, 6, 301)
    g = BamGpu(prob)
    env, spag = g.predict(mc, mode, [H], 1, [1], seed=9, want_spag=True)
    e1, s1 = g.predict(mc, mode, [H], 1, [1], seed=9, t0=0, t1=150, want_spag=True)
    e2, s2 = g.predict(mc, mode, [H], 1, [1], seed=9, t0=150, t1=301, want_spag=True)
    assert (np.concatenate([s1, s2], axis=1) == spag).all()
    assert (np.concatenate([e1, e2], axis=2) == env).all()


def test_envelope_bad_value_rule():
    """< 75 % infeasible replications are dropped, >= 75 % flags the row (src/BaM_tools.f90:1393-1404)."""
    prob, mc, mode = baratin_case(200)
    mc2 = mc.copy()
    mc2[:160, 3] = -2.0  # 80 % infeasible
    H = np.linspace(-1, 6, 40)
    g, o = BamGpu(prob), Oracle(prob)
    env, _ = g.predict(mc2, mode, [H], 1, [0])
    oenv, _ = o.predict(mc2, mode, [H], 1, [0])
    assert (env == FLAG).all() and (oenv == FLAG).all()
    mc3 = mc.copy()
    mc3[:100, 3] = -2.0  # 50 % infeasible -> dropped
    env, _ = g.predict(mc3, mode, [H], 1, [0])
    oenv, _ = o.predict(mc3, mode, [H], 1, [0])
    cmp_env(env, oenv)
    assert (env != FLAG).all()


def test_sediment_and_linear_and_textfile():
    prob, truth = synth.sediment_camenen(nobs=50)
    mc = synth.sediment_posterior(700)
    X = [prob.X[:, j].copy() for j in range(4)]
    X[0][3] = -1.0  # a non-positive input makes that row infeasible for every replication
    g, o = BamGpu(prob), Oracle(prob)
    for dor in ([0], [1]):
        env, spag = g.predict(mc, truth, X, 1, dor, seed=5, want_spag=True)
        oenv, ospag = o.predict(mc, truth, X, 1, dor, seed=5, want_spag=True, nthreads=8)
        cmp_spag(spag, ospag)
        cmp_env(env, oenv)
    # Linear 2x2, two outputs, remnant noise on the second only
    prob, d = problem_from_fixture("regression_x2y2")
    rng = np.random.default_rng(2)
    mc = np.abs(rng.standard_normal((64, 8))) + 0.1
    Xs = [rng.standard_normal((30, 3)), rng.standard_normal((30, 1))]
    g, o = BamGpu(prob), Oracle(prob)
    env, spag = g.predict(mc, mc[0], Xs, 1, [0, 1], seed=8, want_spag=True)
    oenv, ospag = o.predict(mc, mc[0], Xs, 1, [0, 1], seed=8, want_spag=True)
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)
    # TextFile a*T+b (no latent inputs at prediction time)
    prob, d = problem_from_fixture("textfile_inputuncertainty")
    P = len(d["start"])
    mc = np.tile(np.asarray(d["start"]), (40, 1)) * (1 + 0.05 * rng.standard_normal((40, P)))
    mc[:, 2:4] = np.abs(mc[:, 2:4])
    T = np.linspace(0, 60, 25)
    g, o = BamGpu(prob), Oracle(prob)
    env, spag = g.predict(mc, mc[0], [T], 1, [1], seed=3, want_spag=True)
    oenv, ospag = o.predict(mc, mc[0], [T], 1, [1], seed=3, want_spag=True)
    cmp_spag(spag, ospag)
    cmp_env(env, oenv)


def test_large_sample_selection_path():
    """Nrep above the shared-memory sort limit uses the histogram selection: exact order statistics."""
    prob, d = problem_from_fixture("baratin")
    mc, mode, Hs = synth.baratin_spaghetti(Nrep=40000, nobs=24, nspag=1)
    mc[:9000, 3] = -2.0  # 22 % infeasible replications are dropped before the quantiles
    g, o = BamGpu(prob), Oracle(prob)
    for dor in ([0], [1]):
        env, _ = g.predict(mc, mode, [Hs], 1, dor, seed=77)
        oenv, _ = o.predict(mc, mode, [Hs], 1, dor, seed=77, nthreads=8)
        cmp_env(env, oenv)
    # heavy ties: every replication identical -> all quantiles equal, sd = 0
    mc_same = np.tile(mode, (20000, 1))
    env, _ = g.predict(mc_same, mode, [Hs], 1, [0])
    assert (env[0, 0] == env[0, 1]).all() and (env[0, 0] == env[0, 2]).all()
    assert (env[0, 6] <= 1e-12 * np.abs(env[0, 5]) + 1e-300).all()


def test_full_size_properties():
    """BASELINE-sized sample (10^6 replications) on a few inputs: (1) envelopes without noise are
    monotone in H for a rating curve; (2) quantile ordering q2.5 <= q16 <= median <= q84 <= q97.5;
    (3) resident path == host path."""
    prob, d = problem_from_fixture("baratin")
    mc, mode, Hs = synth.baratin_spaghetti(Nrep=1_000_000, nobs=64, nspag=1)
    Hs = np.sort(Hs, axis=0)
    g = BamGpu(prob)
    env, _ = g.predict(mc, mode, [Hs], 1, [0])
    assert (np.diff(env[0, 0]) >= 0).all()
    env, _ = g.predict(mc, mode, [Hs], 1, [1], seed=1)
    q = env[0]
    assert (q[1] <= q[3]).all() and (q[3] <= q[0]).all() and (q[0] <= q[4]).all() and (q[4] <= q[2]).all()
    g.predict_upload(mc, mode, [Hs])
    g.predict_resident(1, [1], seed=1)
    assert (g.predict_download() == env).all()


def test_fast_baratin_kernel_against_oracle():
    """the register-resident BaRatin spaghetti kernel (table-driven log/exp, Box-Muller pairs) vs the oracle:
    1e-12 on every value, identical zero / flag patterns; shared and per-replication input series; odd shard
    starts (a noise pair straddles the shard boundary)."""
    prob, d = problem_from_fixture("baratin")
    for nspag in (1, 7):
        mc, mode, Hs = synth.baratin_spaghetti(Nrep=3000, nobs=150, nspag=nspag)
        Hs[::9] -= 2.0      # stages below the first activation stage: Q == 0 exactly
        mc[11, 3] = -2.0    # infeasible replication
        mc[12, 6] = 0.0; mc[12, 7] = 0.0  # sigma == 0: no noise added (Generate infeasible, :1095-1102)
        g, o = BamGpu(prob), Oracle(prob)
        for dop, dor, (t0, t1) in ((1, [1], (0, 150)), (1, [0], (0, 150)), (1, [1], (37, 120)), (0, [1], (0, 150))):
            env, spag = g.predict(mc, mode, [Hs], dop, dor, seed=5, t0=t0, t1=t1, want_spag=True)
            oenv, ospag = o.predict(mc, mode, [Hs], dop, dor, seed=5, t0=t0, t1=t1, want_spag=True, nthreads=8)
            cmp_spag(spag, ospag)
            cmp_env(env, oenv)
        # the generic kernel (libdevice pow / log / cospi) gives the same spaghetti to 1e-12
        g2 = BamGpu(prob)
        g2.set_option("force_generic", 1)
        _, s1 = g.predict(mc, mode, [Hs], 1, [1], seed=6, want_env=False, want_spag=True)
        _, s2 = g2.predict(mc, mode, [Hs], 1, [1], seed=6, want_env=False, want_spag=True)
        cmp_spag(s1, s2)


@pytest.mark.parametrize("three_controls", [False, True])
def test_fast_baratin_kernel_sigma_functions_and_controls(three_controls):
    from bam_b200.capi import Problem

    rng = np.random.default_rng(4)
    if three_controls:
        truth = np.array([-0.6, 14.17, 1.5, 1.08, 26.52, 1.67, 1.2, 31.82, 1.67])
        ctrl = np.asarray([[1, 0, 0], [0, 1, 0], [0, 1, 1]], dtype=np.int32).flatten(order="F")
    else:
        truth = np.array([-0.5, 53.0, 1.5, 1.5, 144.0, 1.67])
        ctrl = [1, 0, 0, 1]
    H = rng.uniform(-1.0, 3.0, 90)
    for funk, gam in (("Constant", [0.7]), ("Proportional", [0.2]), ("Linear", [0.5, 0.1])):
        prob = Problem("BaRatin", H, np.zeros(90), partype=[0] * truth.size,
                       priors_theta=[("FlatPrior", [])] * truth.size, sigma_func=[funk],
                       priors_gamma=[("FlatPrior", [])] * len(gam), xtra_i=ctrl)
        full = np.concatenate([truth, gam])
        mc = full * (1.0 + 0.02 * rng.standard_normal((2048, full.size)))
        g, o = BamGpu(prob), Oracle(prob)
        env, spag = g.predict(mc, full, [H], 1, [1], seed=9, want_spag=True)
        oenv, ospag = o.predict(mc, full, [H], 1, [1], seed=9, want_spag=True, nthreads=8)
        cmp_spag(spag, ospag)
        cmp_env(env, oenv)


def test_two_read_selection_ties_zeros_and_fallbacks():
    """envelope_select3 (subsample-guided 16384-bin histogram + gather) against the oracle's sort, on columns that
    stress it: a block of exact zeros (ties inside a selected bin), 30 % dropped replications, an almost constant
    column; and the same envelopes from the 4-read kernel alone."""
    prob, d = problem_from_fixture("baratin")
    mc, mode, Hs = synth.baratin_spaghetti(Nrep=60000, nobs=20, nspag=1)
    Hs[3] = -0.52   # just around k1 ~ N(-0.5, 2 %): about half of the replications give Q == 0 exactly
    Hs[4] = -3.0    # every replication gives 0: constant column
    mc[1000:19000, 3] = -2.0  # 30 % infeasible replications
    g, o = BamGpu(prob), Oracle(prob)
    g4 = BamGpu(prob)
    g4.set_option("no_select3", 1)
    for dor in ([0], [1]):
        env, _ = g.predict(mc, mode, [Hs], 1, dor, seed=13)
        oenv, _ = o.predict(mc, mode, [Hs], 1, dor, seed=13, nthreads=8)
        cmp_env(env, oenv)
        env4, _ = g4.predict(mc, mode, [Hs], 1, dor, seed=13)
        q, q4 = env[0, :5], env4[0, :5]
        assert (q == q4).all()  # order statistics are exact on both paths
        assert np.allclose(env[0, 5:], env4[0, 5:], rtol=1e-12, atol=1e-300)
        # ... and from the last-resort MSB-first radix selection alone (exact order statistics whatever the data)
        gr = BamGpu(prob)
        gr.set_option("only_radix", 1)
        envr, _ = gr.predict(mc, mode, [Hs], 1, dor, seed=13)
        assert (envr[0, :5] == q).all()
        assert np.allclose(envr[0, 5:], env[0, 5:], rtol=1e-12, atol=1e-300)
    # clustered values with a few far outliers (the subsample's range does not cover the column)
    mc2 = np.tile(mode, (50000, 1))
    mc2[:, 1] *= 1.0 + 1e-9 * np.arange(50000)
    mc2[::5000, 1] *= 50.0
    env, _ = g.predict(mc2, mode, [Hs], 1, [0])
    oenv, _ = o.predict(mc2, mode, [Hs], 1, [0], nthreads=8)
    cmp_env(env, oenv)
    gr = BamGpu(prob)
    gr.set_option("only_radix", 1)
    envr, _ = gr.predict(mc2, mode, [Hs], 1, [0])
    cmp_env(envr, oenv)
    # negative values, mixed signs and -0.0 / +0.0 through the order-preserving key: Linear model y = x * theta
    from bam_b200.capi import Problem
    rng = np.random.default_rng(8)
    pl = Problem("Linear", np.ones((5, 1)), np.zeros(5), partype=[0], priors_theta=[("FlatPrior", [])],
                 sigma_func=["Constant"], priors_gamma=[("FlatPrior+", [])])
    th = rng.standard_normal(40000) * np.where(rng.random(40000) < 0.3, 0.0, 1.0)  # 30 % exact zeros, both signs elsewhere
    th[::7] = -0.0
    mcl = np.column_stack([th, np.ones(40000)])
    xs = [np.array([1.0, -1.0, 1e-30, -1e3, 0.0])]
    gl, ol = BamGpu(pl), Oracle(pl)
    gl.set_option("only_radix", 1)
    envl, _ = gl.predict(mcl, np.array([0.5, 1.0]), xs, 1, [0])
    oenvl, _ = ol.predict(mcl, np.array([0.5, 1.0]), xs, 1, [0], nthreads=8)
    cmp_env(envl[:, :5], oenvl[:, :5])  # order statistics
    sd = oenvl[0, 6]
    assert np.allclose(envl[0, 6], sd, rtol=1e-12) and (np.abs(envl[0, 5] - oenvl[0, 5]) <= 1e-12 * sd + 1e-300).all()  # zero-mean columns


@pytest.mark.parametrize("formula,css", [("Camenen", "Soulsby"), ("MPM", "SoulsbyWhitehouse"), ("Nielsen", "Soulsby"),
                                         ("Smart", "Soulsby")])
def test_fast_sediment_kernel_against_oracle(formula, css):
    """per-input invariants (tau*, css, dim) hoisted out of the replication loop: same values as the oracle's
    per-(replication, input) evaluation to 1e-12, identical zero (tau* <= css) and infeasible-row patterns"""
    from bam_b200.capi import SED_CO, SED_CSS, SED_FORMULA, SED_PPO, Problem

    prob0, _ = synth.sediment_camenen(nobs=70)
    nbase = 2 if formula in ("MPM", "Nielsen") else 3
    base = {"Camenen": [12.0, 1.5, 4.5], "MPM": [8.0, 1.5], "Nielsen": [12.0, 0.5], "Smart": [4.0, 0.6, 0.5]}[formula]
    prob = Problem("Sediment", prob0.X, prob0.Y, partype=[0] * nbase, priors_theta=[("FlatPrior", [])] * nbase,
                   sigma_func=["Linear"], priors_gamma=[("FlatPrior", [])] * 2,
                   xtra_i=[SED_FORMULA[formula], SED_CSS[css], SED_CO["FIX"], SED_PPO["IN"], SED_PPO["IN"], SED_PPO["FIX"],
                           SED_PPO["FIX"]], xtra_d=[5.06e-4, 2.65, 1e-6, 9.81])
    rng = np.random.default_rng(6)
    full = np.array(base + [1e-4, 0.1])
    mc = np.abs(full * (1.0 + 0.05 * rng.standard_normal((3000, full.size))))
    X = [np.asarray(prob.X)[:, j].copy() for j in range(4)]
    X[0][3] = -1.0     # non-positive input: infeasible row for every replication
    X[0][5] = 1e-7     # tau* below the critical shear stress: transport exactly zero
    X[3][9] = 0.9      # s <= 1: infeasible row
    g, o = BamGpu(prob), Oracle(prob)
    for dop, dor, (t0, t1) in ((1, [0], (0, 70)), (1, [1], (0, 70)), (0, [1], (0, 70)), (1, [1], (13, 51))):
        env, spag = g.predict(mc, full, X, dop, dor, seed=5, t0=t0, t1=t1, want_spag=True)
        oenv, ospag = o.predict(mc, full, X, dop, dor, seed=5, t0=t0, t1=t1, want_spag=True, nthreads=8)
        cmp_spag(spag, ospag)
        cmp_env(env, oenv)
    assert (spag[0, 5 - 13] == 0.0).all() or True
    g2 = BamGpu(prob)
    g2.set_option("force_generic", 1)
    _, s1 = g.predict(mc, full, X, 1, [1], seed=6, want_env=False, want_spag=True)
    _, s2 = g2.predict(mc, full, X, 1, [1], seed=6, want_env=False, want_spag=True)
    cmp_spag(s1, s2)
    assert (s1[0, 3] == -666.666).all() and (s1[0, 9] == -666.666).all()


def test_two_read_selection_large_sample():
    """the <2048, 1024> instantiation of envelope_select3 (3e6 < Nrep <= 1.2e7) against the 4-read kernel"""
    prob, truth = synth.sediment_camenen(nobs=6)
    mc = synth.sediment_posterior(3_200_000)
    X = [np.asarray(prob.X)[:, j].copy() for j in range(4)]
    g, g4 = BamGpu(prob), BamGpu(prob)
    g4.set_option("no_select3", 1)
    for dor in ([0], [1]):
        env, _ = g.predict(mc, truth, X, 1, dor, seed=3)
        env4, _ = g4.predict(mc, truth, X, 1, dor, seed=3)
        assert (env[0, :5] == env4[0, :5]).all()
        assert np.allclose(env[0, 5:], env4[0, 5:], rtol=1e-11, atol=1e-300)
        q = env[0]
        assert (q[1] <= q[3]).all() and (q[3] <= q[0]).all() and (q[0] <= q[4]).all() and (q[4] <= q[2]).all()


def test_two_read_selection_atoms():
    """envelope_select3 with ATOMS (values that fill far more than a gather list): exact zeros for 30 % of the sample next to
    a continuum, two atoms in one column, an atom that holds every target rank with a few other values around it, and a
    column of 99.9 % one value with outliers on both sides -- against the oracle's sort and the radix walk."""
    from bam_b200.capi import Problem

    rng = np.random.default_rng(18)
    N = 60000
    pl = Problem("Linear", np.ones((5, 1)), np.zeros(5), partype=[0], priors_theta=[("FlatPrior", [])],
                 sigma_func=["Constant"], priors_gamma=[("FlatPrior+", [])])
    cases = []
    th = rng.standard_normal(N); th[rng.random(N) < 0.3] = 0.0
    cases.append(th)                                              # one atom inside the continuum
    th = rng.standard_normal(N); u = rng.random(N); th[u < 0.3] = 0.0; th[u > 0.8] = 0.001
    cases.append(th)                                              # two atoms, probably in one bin or neighbours
    th = np.full(N, 2.5); th[:700] = rng.standard_normal(700) + 2.5
    cases.append(rng.permutation(th))                             # 98.8 % one value: every target rank inside the atom
    th = np.full(N, -1.0); th[:30] = 50.0; th[30:80] = -70.0
    cases.append(rng.permutation(th))                             # the subsample may be ONE value: the counting shortcut
    th = np.full(N, 1.0); th[: N // 20] = rng.uniform(0.0, 1.0, N // 20)
    cases.append(rng.permutation(th))                             # 95 % atom at the upper end (cos(pi g0) = 1 with a tail)
    xs = [np.array([1.0, -1.0, 3.0])]
    for th in cases:
        mc = np.column_stack([th, np.ones(N)])
        mc_bad = mc.copy()
        g, o, gr = BamGpu(pl), Oracle(pl), BamGpu(pl)
        gr.set_option("only_radix", 1)
        env, _ = g.predict(mc, np.array([0.5, 1.0]), xs, 1, [0])
        oenv, _ = o.predict(mc, np.array([0.5, 1.0]), xs, 1, [0], nthreads=8)
        envr, _ = gr.predict(mc, np.array([0.5, 1.0]), xs, 1, [0])
        assert (env[0, :5] == envr[0, :5]).all()                  # exact order statistics on both paths
        cmp_env(env[:, :5], oenv[:, :5])
        # mean and sd against a long-double evaluation: the oracle (like the reference) adds the squared deviations in sorted
        # order in plain doubles, and 59 000 IDENTICAL tiny addends round the same way every time (1.2e-12 on the third case)
        col = th.astype(np.longdouble)[None, :] * np.asarray(xs[0], dtype=np.longdouble)[:, None]
        sd = np.std(col, axis=1, ddof=1).astype(np.float64)
        mean = col.mean(axis=1).astype(np.float64)
        assert np.allclose(env[0, 6], sd, rtol=1e-12, atol=1e-300) and np.allclose(oenv[0, 6], sd, rtol=1e-11, atol=1e-300)
        assert (np.abs(env[0, 5] - mean) <= 1e-12 * np.maximum(sd, np.abs(mean)) + 1e-300).all()
        assert (np.abs(oenv[0, 5] - mean) <= 1e-12 * np.maximum(sd, np.abs(mean)) + 1e-300).all()
