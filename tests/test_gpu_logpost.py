"""GPU parity tests of the log-posterior (K1/K2) against the CPU oracle, through the C ABI.

Tolerance (BASELINE.json north_star (a)): |gpu - oracle| <= 1e-12 relative.  The sum over observations
is a tree on the GPU and sequential in the oracle, so the comparison uses the bound of SURVEY.md
section 7: |gpu - long-double oracle| <= max(1e-12*scale, |double oracle - long-double oracle|), with
scale = max(|lp|, sum of |terms|/n ...) approximated by |prior|+|lkh|+|hyper|.
Feasibility / null flags (integer work) must be identical.
"""
import numpy as np
import pytest

from workloads import synth
from bam_b200.capi import BamGpu
from oracle.oracle_api import Oracle
from tests.helpers import perturbed_vectors, problem_from_fixture

pytestmark = pytest.mark.gpu
TOL = 1e-12


def check_parity(prob, x, tol=TOL, nthreads=8, floor=0.0):
    """floor: lower bound of the scale the error is measured against -- for data sets whose log-posterior is a small
    difference of large terms (the oracle itself moves by 1e-13 of such a total between double and long double)"""
    g, o = BamGpu(prob), Oracle(prob)
    assert g.P == o.P and g.nDpar == o.nDpar
    fx, feas, isnull, dpar = g.logpost(x, want_dpar=True)
    pr, lk, hy, feas2, isnull2 = g.logpost_parts(x)
    ofx, ofeas, oisnull, odpar = o.logpost(x, nthreads=nthreads, want_dpar=True)
    assert (feas == ofeas).all() and (isnull == oisnull).all(), "feasibility flags differ"
    assert (feas2 == ofeas).all() and (isnull2 == oisnull).all()
    ok = (ofeas == 1) & (oisnull == 0)
    assert ok.sum() >= max(1, len(ok) // 4), "too few feasible vectors for a meaningful test"
    opr, olk, ohy, _, _ = o.logpost_parts(x[ok][:8], long_double=True)
    scale = np.maximum(np.abs(ofx[ok]), floor)
    err = np.abs(fx[ok] - ofx[ok]) / np.maximum(scale, 1e-300)
    assert err.max() <= tol, f"log-posterior parity {err.max():.3e} > {tol}"
    # parts, against the long-double arbiter
    sel = np.flatnonzero(ok)[:8]
    for a, b in ((pr[sel], opr), (lk[sel], olk), (hy[sel], ohy)):
        assert (np.abs(a - b) <= tol * np.maximum(np.maximum(np.abs(b), 1.0), floor)).all()
    assert (fx[~ok] == -999999999.0).all()
    if g.nDpar:
        assert np.allclose(dpar[ok], odpar[ok], rtol=1e-12, atol=1e-13)  # offsets b = k - (...) cancel
    return err.max()


@pytest.mark.parametrize("name", ["baratin", "baratin_spd", "regression_x2y2", "sediment_camenen_x4",
                                  "sediment_mpm_x2", "textfile_inputuncertainty"])
def test_reference_workspaces(name):
    prob, d = problem_from_fixture(name)
    x = perturbed_vectors(d["start"], 96, rel=0.03, seed=1)
    check_parity(prob, x)


def test_spd_without_copula():
    prob, d = problem_from_fixture("baratin_spd", with_copula=False)
    check_parity(prob, perturbed_vectors(d["start"], 64, rel=0.03, seed=2))


def test_flags_and_infeasible_vectors():
    prob, d = problem_from_fixture("baratin")
    x = perturbed_vectors(d["start"], 32, rel=0.01, seed=3)
    x[1, 6] = -1.0            # gamma1 outside its Uniform prior -> isnull
    x[2, 3] = -1.0            # k2 < k1 -> infeasible continuity
    x[3, 6] = 0.0; x[3, 7] = 0.0   # sigma == 0 -> infeasible
    x[4, 1] = -5.0            # a1 <= 0
    x[5, 6] = -1.0; x[5, 3] = -1.0  # null prior wins over infeasible model (prior is evaluated first)
    g, o = BamGpu(prob), Oracle(prob)
    fx, feas, isnull = g.logpost(x)
    ofx, ofeas, oisnull = o.logpost(x)
    assert (feas == ofeas).all() and (isnull == oisnull).all()
    assert (feas[1], isnull[1]) == (1, 1) and (feas[2], isnull[2]) == (0, 0) and feas[3] == 0 and feas[4] == 0
    assert (feas[5], isnull[5]) == (1, 1)


def test_biases_and_latent_inputs():
    """X random error (latent inputs), X bias, Y bias together (BaM_BaRatin_QbiasHbias shape)."""
    rng = np.random.default_rng(5)
    prob0, truth = synth.baratin_2c(nobs=300, seed=11)
    n = 300
    Xu = np.where(rng.random(n) < 0.5, 0.02, 0.0).reshape(-1, 1)
    Xbindx = rng.integers(0, 3, n).reshape(-1, 1)
    Xb = np.where(Xbindx > 0, 0.01, 0.0)
    Ybindx = rng.integers(0, 4, n).reshape(-1, 1)
    Yb = np.where(Ybindx > 0, 0.05 * np.abs(prob0.Y), 0.0)
    Y = prob0.Y.copy()
    Y[7, 0] = -9999.0  # a missing value is skipped (src/BaM_tools.f90:3207)
    from bam_b200.capi import Problem

    prob = Problem("BaRatin", prob0.X, Y, Yu=prob0.Yu, Xu=Xu, Xb=Xb, Xbindx=Xbindx, Yb=Yb, Ybindx=Ybindx,
                   partype=[0] * 6, priors_theta=[("Gaussian", [t, abs(t) * 0.2 + 0.1]) for t in truth[:6]],
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000]), ("LogNormal", [-2.0, 1.0])],
                   xtra_i=[1, 0, 0, 1])
    o = Oracle(prob)
    nlat = int((Xu > 0).sum())
    assert o.sizes.nLatent == nlat and o.sizes.nXbias == 2 and o.sizes.nYbias == 3
    start = np.concatenate([truth, prob0.X[Xu[:, 0] > 0, 0], np.zeros(5)])
    x = perturbed_vectors(start, 48, rel=0.01, seed=6)
    x[:, -5:] = 0.3 * rng.standard_normal((48, 5))
    check_parity(prob, x)


@pytest.mark.parametrize("funk,gam", [("Constant", [0.7]), ("Proportional", [0.2]), ("Power", [0.3, 0.1, 0.8]),
                                      ("Exponential", [0.2, 5.0, 1.5]), ("Gaussian", [0.2, 5.0, 1.5])])
def test_all_sigma_functions(funk, gam):
    from bam_b200.capi import Problem

    prob0, truth = synth.baratin_2c(nobs=200, seed=12)
    prob = Problem("BaRatin", prob0.X, prob0.Y, Yu=prob0.Yu, partype=[0] * 6,
                   priors_theta=[("Gaussian", [t, abs(t) * 0.2 + 0.1]) for t in truth[:6]], sigma_func=[funk],
                   priors_gamma=[("Uniform", [0, 100])] * len(gam), xtra_i=[1, 0, 0, 1])
    x = perturbed_vectors(np.concatenate([truth[:6], gam]), 32, rel=0.01, seed=7)
    check_parity(prob, x)


def test_all_prior_families_and_fix():
    from bam_b200.capi import Problem

    prob0, truth = synth.baratin_2c(nobs=120, seed=13)
    priors = [("Triangle", [-0.5, -2.0, 1.0]), ("LogNormal", [4.0, 0.5]), ("Exponential", [1.0, 2.0]),
              ("FlatPrior", []), ("FlatPrior+", [])]
    prob = Problem("BaRatin", prob0.X, prob0.Y, Yu=prob0.Yu, partype=[0, 0, 0, 0, 0, -1],
                   fix_value=[0, 0, 0, 0, 0, 1.67], priors_theta=priors, sigma_func=["Linear"],
                   priors_gamma=[("Uniform", [0, 1000]), ("FlatPrior+", [])], xtra_i=[1, 0, 0, 1])
    start = np.concatenate([truth[:5], truth[6:]])
    x = perturbed_vectors(start, 48, rel=0.02, seed=8)
    x[3, 1] = -1.0  # LogNormal at x<=0 -> null
    x[4, 6] = -0.1  # FlatPrior+ at x<=0 -> null
    check_parity(prob, x)


def test_cfg2_synthetic_medium():
    """BaRatin_SPD synthetic (cfg 2 generator) at a size the oracle finishes in seconds; with copula."""
    prob, truth = synth.baratin_spd(nobs=20000, copula=True)
    x = synth.feasible_starts(truth, 24, check=synth.spd_start_check)
    check_parity(prob, x)


def test_cfg3_linear_latent_medium():
    prob, truth = synth.linear_latent(nobs=4000)
    x = perturbed_vectors(truth, 6, rel=0.01, seed=9)
    check_parity(prob, x)


def test_cfg3_textfile_latent_medium():
    prob, truth = synth.textfile_latent(nobs=3000)
    x = perturbed_vectors(truth, 6, rel=0.005, seed=10)
    check_parity(prob, x)


def test_cfg4_sediment_medium():
    prob, truth = synth.sediment_camenen(nobs=5000)
    x = perturbed_vectors(truth, 16, rel=0.02, seed=11)
    check_parity(prob, x)


def test_small_and_ragged_sizes():
    """n = 1, 31, 33, 257 observations (tile edges) and K = 1, 3, 65 vectors (chain-tile edges)."""
    for n in (1, 31, 33, 257):
        prob, truth = synth.baratin_2c(nobs=n, seed=20 + n)
        for K in (1, 3, 65):
            check_parity(prob, perturbed_vectors(truth, K, rel=0.01, seed=n + K))


def test_full_size_properties_cfg2():
    """BASELINE size (4096 chains x 100k gaugings): size-independent properties.
    (1) additivity: the likelihood over all observations equals the sum over two disjoint halves;
    (2) batch independence: a vector evaluated alone equals the same vector inside the batch (bitwise);
    (3) spot parity of 4 vectors against the oracle."""
    n, K = 100_000, 4096
    prob, truth = synth.baratin_spd(nobs=n)
    x = synth.feasible_starts(truth, K, check=synth.spd_start_check)
    g = BamGpu(prob)
    pr, lk, hy, feas, isnull = g.logpost_parts(x)
    assert feas.all() and not isnull.any()
    from bam_b200.capi import Problem

    def half(sel):
        return Problem("BaRatin", prob.X[sel], prob.Y[sel], Yu=prob.Yu[sel], partype=prob.partype, nval=prob.nval,
                       var_indx=prob.var_indx[sel], priors_theta=[(int(d), list(p)) for d, p in
                                                                  zip(prob.prior_theta_dist, prob.prior_theta_par)],
                       sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000])] * 2, xtra_i=prob.xtra_i)

    idx = np.arange(n)
    lkA = BamGpu(half(idx % 2 == 0)).logpost_parts(x)[1]
    lkB = BamGpu(half(idx % 2 == 1)).logpost_parts(x)[1]
    assert (np.abs(lkA + lkB - lk) <= 1e-12 * np.abs(lk)).all()
    fx_all = g.logpost(x)[0]
    for k in (0, 17, 4095):  # K = 1 runs the observation-parallel kernel, K = 4096 the chain-parallel one
        assert abs(g.logpost(x[k])[0][0] - fx_all[k]) <= 1e-12 * abs(fx_all[k])
    assert (g.logpost(x[:2048])[0] == fx_all[:2048]).all()  # same kernel: bitwise independent of the batch
    o = Oracle(prob)
    ofx = o.logpost(x[:4], nthreads=4)[0]
    assert (np.abs(fx_all[:4] - ofx) <= 1e-12 * np.abs(ofx)).all()


def test_full_size_properties_cfg3():
    """BASELINE configs[2] size (10^6 observations, every input latent: P = 2 000 008): size-independent properties of
    the cfg-3 fast path.  (1) the specialised kernel equals the generic one on the same vectors (parts included);
    (2) bitwise repeatability and batch independence; (3) moving ONE latent input changes the posterior by exactly the
    local term of its observation (the factorisation the latent sweep relies on); (4) spot parity against the oracle."""
    n, K = 1_000_000, 8
    prob, truth = synth.linear_latent(nobs=n)
    rng = np.random.default_rng(5)
    x = np.tile(truth, (K, 1))
    x[:, :8] *= 1.0 + 0.01 * rng.standard_normal((K, 8))
    x[:, 8:] += 0.05 * rng.standard_normal((K, 2 * n))
    g = BamGpu(prob)
    pr, lk, hy, feas, isnull = g.logpost_parts(x)
    assert feas.all() and not isnull.any()
    gg = BamGpu(prob)
    gg.set_option("force_generic", 1)
    pr2, lk2, hy2, _, _ = gg.logpost_parts(x)
    assert (pr == pr2).all() and np.allclose(lk, lk2, rtol=1e-12, atol=0) and np.allclose(hy, hy2, rtol=1e-12, atol=0)
    fx = g.logpost(x)[0]
    assert (g.logpost(x)[0] == fx).all() and (g.logpost(x[:3])[0] == fx[:3]).all()
    # (3) one latent input of observation i: f changes by [lkh_i + hyper_i](new) - [lkh_i + hyper_i](old)
    i, d = 123_457, 0.07
    y = x[:2].copy()
    y[:, 8 + i] += d  # first input column of observation i
    fy = g.logpost(y)[0]
    X, Y, Yu, Xu = (np.asarray(a) for a in (prob.X, prob.Y, prob.Yu, prob.Xu))

    def local(v):
        th = v[:4].reshape(2, 2, order="F")
        xt = np.array([v[8 + i], v[8 + n + i]])
        yh = xt @ th
        sig = v[[4, 6]] + v[[5, 7]] * np.abs(yh)
        var = Yu[i] ** 2 + sig**2
        lkh = np.sum(-0.5 * np.log(2 * np.pi * var) - 0.5 * (Y[i] - yh) ** 2 / var)
        hyp = np.sum(-0.5 * np.log(2 * np.pi * Xu[i] ** 2) - 0.5 * ((X[i] - xt) / Xu[i]) ** 2)
        return lkh + hyp

    for k in range(2):
        assert abs((fy[k] - fx[k]) - (local(y[k]) - local(x[k]))) <= 1e-9 * abs(fx[k]) * 1e-3 + 1e-6
    ofx = Oracle(prob).logpost(x[:2], nthreads=2)[0]
    assert (np.abs(fx[:2] - ofx) <= 1e-12 * np.abs(ofx)).all()


@pytest.mark.parametrize("name,copula", [("baratin", False), ("baratin_spd", True), ("baratin_spd", False)])
def test_chain_parallel_fast_path(name, copula):
    """K >= 128 BaRatin problems without latent inputs / biases run the chain-parallel kernel with the
    table-driven log/exp: same flags, 1e-12 parity against the oracle and against the generic kernel."""
    prob, d = problem_from_fixture(name, with_copula=copula)
    x = perturbed_vectors(d["start"], 300, rel=0.03, seed=4)
    x[7, 3 if name == "baratin" else 7] = -5.0   # infeasible continuity for one vector
    x[9, len(d["start"]) - 2] = -1.0             # gamma1 < 0: null prior
    check_parity(prob, x)
    g = BamGpu(prob)
    fast = g.logpost(x)
    g.set_option("force_generic", 1)
    gen = g.logpost(x)
    assert (fast[1] == gen[1]).all() and (fast[2] == gen[2]).all()
    ok = fast[1] == 1
    assert (np.abs(fast[0][ok] - gen[0][ok]) <= 1e-12 * np.abs(gen[0][ok])).all()


def test_fast_path_edge_observations():
    """H exactly on an activation stage (upper range), H == b1 (zero discharge through log(0) = -inf),
    H below every control (Q = 0), missing Y, 1..4 controls, Constant / Proportional sigma."""
    from bam_b200.capi import Problem

    rng = np.random.default_rng(3)
    for nc, funk, gam in ((1, "Constant", [0.8]), (2, "Proportional", [0.3]), (3, "Linear", [0.5, 0.1]),
                          (4, "Linear", [0.5, 0.1])):
        lowest = 0.2 if funk == "Proportional" else -3.0  # Proportional: Q = 0 gives sigma = 0 -> infeasible for all
        k = np.array([-0.5, 0.7, 1.9, 3.1])[:nc]
        a = np.array([20.0, 35.0, 50.0, 65.0])[:nc]
        c = np.array([1.5, 1.67, 1.67, 2.0])[:nc]
        M = np.tril(np.ones((nc, nc), dtype=np.int32))  # every control stays active once activated
        theta = np.column_stack([k, a, c]).ravel()
        H = np.concatenate([rng.uniform(max(-1.0, lowest), 4.5, 500), k[1:], [np.nextafter(k[1], 9) if nc > 1 else 1.0, lowest]])
        if funk != "Proportional":
            H = np.concatenate([H, [-0.5, np.nextafter(-0.5, 1)]])
        Y = np.abs(rng.standard_normal(H.size)) * 50 + 1.0
        Y[5] = -9999.0
        Yu = 0.05 * Y
        prob = Problem("BaRatin", H, Y, Yu=np.abs(Yu), partype=[0] * (3 * nc),
                       priors_theta=[("Gaussian", [t, abs(t) * 0.3 + 0.1]) for t in theta], sigma_func=[funk],
                       priors_gamma=[("Uniform", [0, 100])] * len(gam), xtra_i=M.flatten(order="F"))
        x = perturbed_vectors(np.concatenate([theta, gam]), 200, rel=0.01, seed=nc)
        x[0] = np.concatenate([theta, gam])  # exact k: H == k rows hit the boundary bit-exactly
        check_parity(prob, x)


@pytest.mark.parametrize("formula,css", [("Camenen", "Soulsby"), ("MPM", "SoulsbyWhitehouse"), ("Nielsen", "Soulsby"), ("Smart", "Soulsby")])
def test_sediment_fast_logpost_against_oracle(formula, css):
    """the chain-parallel Sediment kernel (per-row invariants hoisted, one table-driven exp per chain x row) at a batch
    that takes it (K >= 128): parity with the oracle, which keeps Sediment_Apply's own formulation
    (SedimentTransport_model.f90:100-192), and with the generic kernel; rows with tau* <= css (Q = 0), a missing output,
    vectors with a null prior / sigma <= 0, and a data set with an infeasible row included"""
    from bam_b200.capi import SED_CO, SED_CSS, SED_FORMULA, SED_PPO, Problem

    prob0, _ = synth.sediment_camenen(nobs=300)
    base = {"Camenen": [12.0, 1.5, 4.5], "MPM": [8.0, 1.5], "Nielsen": [12.0, 0.5], "Smart": [4.0, 0.6, 0.5]}[formula]
    nbase = len(base)
    X, Y = np.asarray(prob0.X).copy(), np.asarray(prob0.Y).copy()
    X[:12, 0] *= 1e-4   # very shallow rows: tau* <= css -> Q = 0
    Y[7, 0] = -9999.0   # a missing output

    def make(Xm):
        return Problem("Sediment", Xm, Y, partype=[0] * nbase, priors_theta=[("Gaussian", [b, 0.5 * b]) for b in base],
                       sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1]), ("Uniform", [0, 10])],
                       xtra_i=[SED_FORMULA[formula], SED_CSS[css], SED_CO["FIX"], SED_PPO["IN"], SED_PPO["IN"], SED_PPO["FIX"],
                               SED_PPO["FIX"]], xtra_d=[5.06e-4, 2.65, 1e-6, 9.81])

    prob = make(X)
    truth = np.array(base + [1e-4, 0.1])
    x = synth.feasible_starts(truth, 200, rel=0.02, seed=31)
    x[9, nbase] = -0.5      # gamma1 outside its Uniform prior: null density
    x[11, nbase] = 0.0      # gamma1 = 0 with Q = 0 rows: sigma = 0 -> infeasible (:3211)
    x[11, nbase + 1] = 0.05
    # 300 terms of either sign and magnitude 1..100 sum to totals of a few tens: measured against max(|total|, n)
    check_parity(prob, x, floor=float(prob.nobs))
    g, g2 = BamGpu(prob), BamGpu(prob)
    g2.set_option("force_generic", 1)
    fx, fg = g.logpost(x)[0], g2.logpost(x)[0]
    ok = fx != -999999999.0
    assert (ok == (fg != -999999999.0)).all() and ok.sum() >= 190
    assert np.allclose(fx[ok], fg[ok], rtol=1e-12, atol=1e-12 * prob.nobs)
    X2 = X.copy()
    X2[40, 1] = -1e-3       # a non-positive input: the row, hence every vector, is infeasible
    gb, ob = BamGpu(make(X2)), Oracle(make(X2))
    assert (gb.logpost(x)[1] == ob.logpost(x)[1]).all() and not gb.logpost(x)[1][[0, 1, 2]].any()


def test_taylor_subtile_paths_of_the_fast_kernel():
    """generation 3 of the BaRatin fast path: dense, stage-sorted gaugings make most (chain, sub-tile) pairs take the
    Taylor expansion; this data set is built to visit every branch around it -- degree classes 5 .. 16 (sub-tiles close
    to an offset b need the high degrees), the group path (arguments H - b too small for any degree, exponents outside
    (0, 5]), stage boundaries inside a sub-tile and inside a group of five, chains that fail the remnant gate, a partial
    last chain block, infeasible chains in the middle of a warp, three sigma functions -- against the oracle (which
    evaluates one libm pow per gauging and control, BaRatin_model.f90:229-250) and against the generic kernel."""
    from bam_b200.capi import Problem

    rng = np.random.default_rng(12)
    n = 30_000
    k = np.array([-0.4, 0.9, 1.6])
    a = np.array([18.0, 30.0, 42.0])
    c = np.array([1.5, 1.67, 2.2])
    M = np.array([[1, 0, 0], [0, 1, 0], [0, 1, 1]], dtype=np.int32)
    # dense stages, with a cluster right above the third activation stage (tiny H - b3) and a sparse upper tail
    H = np.concatenate([rng.uniform(-0.39, 2.4, n - 3000), 1.6 + np.abs(rng.standard_normal(2000)) * 2e-3, rng.uniform(2.4, 30.0, 1000)])
    theta = np.column_stack([k, a, c]).ravel()
    for funk, gam in (("Linear", [0.6, 0.08]), ("Constant", [1.5]), ("Proportional", [0.2])):
        Y = np.abs(rng.standard_normal(n)) * 40 + 2.0
        prob = Problem("BaRatin", H, Y, Yu=0.05 * Y, partype=[0] * 9,
                       priors_theta=[("Gaussian", [t, abs(t) * 0.5 + 0.1]) for t in theta], sigma_func=[funk],
                       priors_gamma=[("Uniform", [0, 1e7])] * len(gam), xtra_i=M.flatten(order="F"))
        x = perturbed_vectors(np.concatenate([theta, gam]), 300, rel=0.02, seed=5)
        x[3, 3] = -1.0            # k2 < k1: infeasible vector inside a warp
        x[40, 2] = 6.5            # c1 outside (0, 5]: no expansion for that chain, group path with c = 6.5
        x[41, 5] = -0.8           # a negative exponent: (H - b)^c grows without bound next to the offset
        x[77, 9] = 3e6            # gamma1 beyond the gate of the group path (2^20): every gauging through the checked path
        x[78, 9] = 1e-19          # ... and below it
        x[130, 0] = -0.3999999    # k1 a hair above the lowest stages: the first group straddles the activation stage
        err = check_parity(prob, x, tol=1e-12, floor=float(n) * 1e-3)
        g, g2 = BamGpu(prob), BamGpu(prob)
        g2.set_option("force_generic", 1)
        fx, fg = g.logpost(x), g2.logpost(x)
        assert (fx[1] == fg[1]).all() and (fx[2] == fg[2]).all()
        ok = fx[1] == 1
        assert ok.sum() >= 200
        assert (np.abs(fx[0][ok] - fg[0][ok]) <= 1e-12 * np.maximum(np.abs(fg[0][ok]), n * 1e-3)).all()
        assert (g.logpost(x)[0] == fx[0]).all()  # bitwise repeatable (dynamic item order, fixed-order sums)


def _baratin_latent_bias(n, seed=5):
    """BaRatin with X random error on half the gaugings (latent stages), X bias, Y bias, one missing Y."""
    from bam_b200.capi import Problem

    rng = np.random.default_rng(seed)
    prob0, truth = synth.baratin_2c(nobs=n, seed=11)
    Xu = np.where(rng.random(n) < 0.5, 0.02, 0.0).reshape(-1, 1)
    Xbindx = rng.integers(0, 3, n).reshape(-1, 1)
    Xb = np.where(Xbindx > 0, 0.01, 0.0)
    Ybindx = rng.integers(0, 4, n).reshape(-1, 1)
    Yb = np.where(Ybindx > 0, 0.05 * np.abs(prob0.Y), 0.0)
    Y = prob0.Y.copy()
    Y[7, 0] = -9999.0
    prob = Problem("BaRatin", prob0.X, Y, Yu=prob0.Yu, Xu=Xu, Xb=Xb, Xbindx=Xbindx, Yb=Yb, Ybindx=Ybindx,
                   partype=[0] * 6, priors_theta=[("Gaussian", [t, abs(t) * 0.2 + 0.1]) for t in truth[:6]],
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000]), ("LogNormal", [-2.0, 1.0])],
                   xtra_i=[1, 0, 0, 1])
    start = np.concatenate([truth, prob0.X[Xu[:, 0] > 0, 0], np.zeros(5)])
    return prob, start


@pytest.mark.parametrize("case", ["textfile", "baratin_bias", "linear_partial"])
def test_latent_tile_kernel(case):
    """logpost_tile (latent inputs, n >= 16384: register-resident per-chain sums over a tile of observations) against the
    oracle and against logpost_main; K = 10 leaves a ragged chain tile; an infeasible and a null vector."""
    n, K = 20000, 10
    rng = np.random.default_rng(3)
    if case == "textfile":
        prob, truth = synth.textfile_latent(nobs=n)
        x = perturbed_vectors(truth, K, rel=0.005, seed=10)
        x[4, 2] = -1.0  # gamma1 < 0: outside its Uniform(0, 100) prior
    elif case == "baratin_bias":
        prob, start = _baratin_latent_bias(n)
        x = perturbed_vectors(start, K, rel=0.01, seed=6)
        x[:, -5:] = 0.3 * rng.standard_normal((K, 5))
        x[4, 3] = x[4, 0] - 1.0  # k2 < k1: infeasible rating curve
    else:
        from bam_b200.capi import Problem

        X = rng.uniform(1.0, 5.0, (n, 2))
        Xu = np.column_stack([np.where(rng.random(n) < 0.3, 0.05, 0.0), np.full(n, 0.1)])  # column 0 partly latent
        th = np.array([1.5, -0.7, 0.4, 2.0])
        Y = np.column_stack([X @ th[:2], X @ th[2:]]) + 0.1 * rng.standard_normal((n, 2))
        Y[11, 1] = -9999.0
        prob = Problem("Linear", X, Y, Yu=np.full((n, 2), 0.05), Xu=Xu, partype=[0] * 4,
                       priors_theta=[("Gaussian", [t, 1.0]) for t in th], sigma_func=["Constant", "Proportional"],
                       priors_gamma=[("Uniform", [0, 10]), ("Uniform", [0, 10])])
        lat = np.concatenate([X[Xu[:, 0] > 0, 0], X[:, 1]])
        x = perturbed_vectors(np.concatenate([th, [0.1, 0.05], lat]), K, rel=0.01, seed=7)
        x[4, 4] = -0.5
    check_parity(prob, x, floor=float(n) * 1e-3)
    g, g2 = BamGpu(prob), BamGpu(prob)
    g2.set_option("force_generic", 1)
    a, b = g.logpost_parts(x), g2.logpost_parts(x)
    assert (a[3] == b[3]).all() and (a[4] == b[4]).all()
    ok = (a[3] == 1) & (a[4] == 0)
    assert ok.sum() >= K - 2
    for u, v in ((a[1], b[1]), (a[2], b[2])):  # likelihood, hyper
        assert (np.abs(u[ok] - v[ok]) <= 1e-12 * np.maximum(np.abs(v[ok]), n * 1e-3)).all()
    assert (g.logpost(x[3:9])[0] == g.logpost(x)[0][3:9]).all()  # a chain's value does not depend on its tile mates
