"""CPU tests of the oracle: pinned against everything the reference offers for this path
(SURVEY.md 8c): known-answer columns in the sediment data files, control matrices / VAR index
columns, and the independent surveyor restatement of SURVEY.md App. F."""
import numpy as np
import pytest

from oracle import oracle_api
from oracle.oracle_api import Oracle
from tests.helpers import load_fixture, problem_from_fixture


# SURVEY.md App. F (independent scratch restatement, sequential python sums)
APP_F = {
    "baratin": dict(fx=-208.221439619687, lkh=-192.240250729638, prior=-15.981188890049, b2=0.47563645511404),
    "sediment_camenen_x4": dict(fx=2843.24519325832, lkh=2894.82100383739, prior=-51.5758105790742),
    "regression_x2y2": dict(fx=-306.503863198421, lkh=-252.165941407072, prior=-54.3379217913487),
}


@pytest.mark.parametrize("name", sorted(APP_F))
def test_logpost_matches_independent_restatement(name):
    p, d = problem_from_fixture(name)
    r = Oracle(p).logpost_one(d["start"])
    ref = APP_F[name]
    assert r["feas"] == 1 and r["isnull"] == 0
    for key in ("fx", "lkh", "prior"):
        assert abs(r[key] - ref[key]) <= 2e-12 * abs(ref[key]), (key, r[key], ref[key])
    if "b2" in ref:
        assert abs(r["dpar"][1] - ref["b2"]) < 1e-13
        assert r["dpar"][0] == d["start"][0]  # b1 = k1 exactly (BaRatin_model.f90:375)


def test_long_double_arbiter_agrees():
    for name in APP_F:
        p, d = problem_from_fixture(name)
        o = Oracle(p)
        a, b = o.logpost_one(d["start"]), o.logpost_one(d["start"], long_double=True)
        assert abs(a["fx"] - b["fx"]) <= 1e-12 * abs(a["fx"])


def test_sediment_known_answer_columns():
    """tests/BaM_Sediment_Camenen_X4/XY.BAD carries tau, qsCL, sediment_D, tau_cr_soulsby (3 digits).
    tau and D pin the inputs/units; qsCL was generated with the ADDITIVE Soulsby formula while the
    reference CODE multiplies the two terms (SedimentTransport_model.f90:302 vs :39-40).  Parity
    target is the code; the discrepancy is asserted explicitly so that it stays documented."""
    p, d = problem_from_fixture("sediment_camenen_x4")
    n, ncol = d["nobs"], d["data_ncol"]
    W = np.asarray(d["data_table"]).reshape(ncol, n).T
    h, S0, dd, s, v, g, qs, tau_col, qsCL, SD_col, css_col = W.T
    tau = h * S0 / ((s - 1.0) * dd)
    SD = (g * (s - 1.0) / v**2) ** (1.0 / 3.0) * dd
    assert np.allclose(tau, tau_col, rtol=1.5e-2)  # inputs and answers are printed with 3 digits
    assert np.allclose(SD, SD_col, rtol=6e-3)
    css_add = 0.3 / (1 + 1.2 * SD) + 0.055 * (1 - np.exp(-0.02 * SD))
    css_mul = (0.3 / (1 + 1.2 * SD)) * 0.055 * (1 - np.exp(-0.02 * SD))
    assert np.allclose(css_add, css_col, rtol=1.5e-2)
    o = Oracle(p)
    theta = np.array([12.0, 1.5, 4.5])
    Y, feas, _ = o.apply_model(np.column_stack([h, S0, dd, s]), theta)
    assert feas.all()
    code = np.sqrt(g * (s - 1) * dd**3) * 12.0 * tau**1.5 * np.exp(-4.5 * css_mul / tau)
    assert np.allclose(Y[:, 0], code, rtol=1e-13)
    # undo the css difference -> the data file's qsCL (3 significant digits)
    as_file = Y[:, 0] * np.exp(-4.5 * (css_add - css_mul) / tau)
    assert np.allclose(as_file, qsCL, rtol=4e-2)


def test_spd_fixture_structure():
    """control matrix / VAR index column / prior correlation of tests/BaM_BaRatin_SPD"""
    p, d = problem_from_fixture("baratin_spd")
    o = Oracle(p)
    assert d["xtra_i"] == [1, 0, 0, 0, 1, 1, 0, 0, 1]  # column-major [[1,0,0],[0,1,0],[0,1,1]]
    assert (o.sizes.nFit, o.sizes.nGamma, o.sizes.nInfer, o.sizes.nCombo, o.sizes.nDpar) == (17, 2, 19, 5, 15)
    vi = np.asarray(d["var_indx"]).reshape(9, d["nobs"]).T
    assert set(np.unique(vi[:, 0])) == {1, 2, 3, 4, 5} and (vi[:, 0] == vi[:, 3]).all()
    C = np.asarray(d["priorC"]).reshape(19, 19)
    assert np.allclose(C, C.T) and np.allclose(np.diag(C), 1.0)
    r = o.logpost_one(d["start"])
    assert r["feas"] == 1
    # b3 = k3 because range 3 keeps control 2 and adds control 3 (BaRatin_model.f90:386-393)
    assert all(r["dpar"][3 * k + 2] == 1.2 for k in range(5))
    # copula term changes the prior, not the likelihood
    p0, _ = problem_from_fixture("baratin_spd", with_copula=False)
    r0 = Oracle(p0).logpost_one(d["start"])
    assert r0["lkh"] == r["lkh"] and r0["prior"] != r["prior"]


def test_early_exit_flags():
    p, d = problem_from_fixture("baratin")
    o = Oracle(p)
    x = np.array(d["start"])
    bad = x.copy(); bad[6] = -1.0          # gamma1 outside Uniform(0,1000) -> isnull, feas stays true
    r = o.logpost_one(bad)
    assert (r["feas"], r["isnull"]) == (1, 1) and r["fx"] == -999999999.0
    bad = x.copy(); bad[3] = -1.0          # k2 < k1 -> continuity infeasible
    r = o.logpost_one(bad)
    assert (r["feas"], r["isnull"]) == (0, 0)
    bad = x.copy(); bad[6] = 0.0; bad[7] = 0.0   # sigma = 0 -> infeasible (src/BaM_tools.f90:3210)
    assert o.logpost_one(bad)["feas"] == 0


def test_baratin_range_boundaries():
    """H == k_i belongs to the upper range (strict <, BaRatin_model.f90:427); below k1 -> Q = 0."""
    p, d = problem_from_fixture("baratin")
    o = Oracle(p)
    th = np.array(d["start"][:6])
    H = np.array([-1.0, -0.5, np.nextafter(-0.5, 1), 1.5, np.nextafter(1.5, -1), 3.0])
    Y, feas, dpar = o.apply_model(H, th)
    b2 = dpar[1]
    assert Y[0, 0] == 0.0 and Y[1, 0] == 0.0  # (H-b1)=0 -> 53*0**1.5 = 0
    assert Y[3, 0] == 144.0 * (1.5 - b2) ** 1.67          # range 2: control 2 only
    assert Y[4, 0] == 53.0 * (H[4] + 0.5) ** 1.5          # still range 1
    assert abs(Y[3, 0] - Y[4, 0]) < 1e-9 * Y[3, 0]        # continuity at k2
    assert feas.all()


def test_textfile_compiler_semantics():
    """a+b-c -> a+(b-c); a*b/c -> a*(b/c); a^b^c -> (a^b)^c; leading minus negates the product
    (TextFile_model.f90:702-726)."""
    from bam_b200.capi import Problem

    def ev(formula, vals):
        text = f"3\nx,y,z\n0\n\n1\n{formula}\n"
        p = Problem("TextFile", np.array([vals]), np.array([[0.0]]), partype=[], priors_theta=[],
                    sigma_func=["Constant"], priors_gamma=[("FlatPrior", [])], xtra_text=text)
        Y, feas, _ = Oracle(p).apply_model(np.array([vals]), np.array([]))
        return Y[0, 0], feas[0]

    x, y, z = 0.1, 0.2, 0.3
    assert ev("x+y-z", (x, y, z))[0] == x + (y - z)
    assert ev("x-y+z", (x, y, z))[0] == (x - y) + z
    assert ev("x*y/z", (x, y, z))[0] == x * (y / z)
    assert ev("x**y**z", (x, y, z))[0] == (x**y) ** z
    assert ev("-x*y", (x, y, z))[0] == -(x * y)
    assert ev("-x^2", (3.0, y, z))[0] == -9.0
    assert ev("2.5e-1*x+exp(-y)", (x, y, z))[0] == 2.5e-1 * x + np.exp(-y)
    assert ev("Abs(-x)+SQRT(y)", (x, y, z))[0] == abs(-x) + np.sqrt(y)
    assert ev("ispos(x)*is0(y-y)+issneg(z)", (x, y, z))[0] == 1.0
    assert ev("x/(y-y)", (x, y, z))[1] == 0      # division by zero -> row infeasible
    assert ev("log(-x)", (x, y, z))[1] == 0
    assert ev("sqrt(-x)", (x, y, z))[1] == 0
    assert ev("asin(x+2)", (x, y, z))[1] == 0


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    assert oracle_api.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle_api.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle_api.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normal_stream_moments_and_quantile():
    z = np.array([oracle_api.normal(7, i, 3, 1) for i in range(20000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    u = np.array([oracle_api.uniform(7, i, 0, 0) for i in range(20000)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    from scipy.stats import norm
    for p in (1e-12, 1e-6, 0.025, 0.3, 0.5, 0.7, 0.975, 1 - 1e-9):
        assert abs(oracle_api.normal_quantile(p) - norm.ppf(p)) <= 1e-12 * max(1.0, abs(norm.ppf(p)))


def test_envelope_rule():
    """< 75 % bad -> dropped; >= 75 % bad -> row of flags (src/BaM_tools.f90:1393-1404); Hazen quantiles."""
    f = -666.666
    col = np.arange(1.0, 101.0)
    e = oracle_api.envelope(col)
    assert e[0] == 50.5 and abs(e[5] - 50.5) < 1e-12 and abs(e[6] - col.std(ddof=1)) < 1e-12
    assert abs(e[1] - (0.025 * 100 + 0.5 - 3 + 3)) < 1e-12  # h = 3.0 -> x_3 = 3
    c2 = col.copy(); c2[:70] = f
    e2 = oracle_api.envelope(c2)
    assert e2[0] == np.median(col[70:]) and (e2 != f).all()
    c3 = col.copy(); c3[:75] = f
    assert (oracle_api.envelope(c3) == f).all()
    c4 = col.copy(); c4[3] = np.nan
    assert np.isfinite(oracle_api.envelope(c4)).all()


# ---------------------------------------------------------------------------------------------------
# Vegetation: known-answer columns of the reference's own tests/BaM_Vegetation/SyntheticData.txt
# (tests/golden/make_vegetation_fixture.py explains how the generator's parameters were recovered)
# ---------------------------------------------------------------------------------------------------
def _vegetation_fixture_problem():
    from bam_b200.capi import VEG_GROWTH, Problem

    d = load_fixture("vegetation_synthetic")
    n, nyear = d["nobs"], d["nYear"]
    X = np.column_stack([d["time"], d["stage"], d["Temp"], d["Temp"]])
    Y = np.column_stack([d["Q"], d["g"], d["cosg"], d["sing"], d["g0"]])
    nt = len(d["theta"])
    nw = d["newton"]
    prob = Problem("Vegetation", X, Y, partype=[0] * nt, priors_theta=[("FlatPrior", [])] * nt,
                   sigma_func=["Constant"] * 5, priors_gamma=[("FlatPrior", [])] * 5,
                   xtra_i=[VEG_GROWTH[d["growth"]], nyear, int(nw[4])], xtra_d=nw[:4])
    return prob, d, X, Y, n


def test_vegetation_known_answer_columns():
    prob, d, X, Y, n = _vegetation_fixture_problem()
    o = Oracle(prob)
    Yh, feas, dpar = o.apply_model(X, np.asarray(d["theta"]))
    assert feas.all()
    # the file prints 9-10 significant digits; the time column is rounded too, which moves the single sample that
    # sits on t_f (g0 = 0 in the file) by 2e-8
    tfrac = X[:, 0] - np.floor(X[:, 0])
    # the time column is rounded to 1e-9; on the steep declining limb (t_e < t <= t_f = t_e + 35 days) the factor
    # (t_f - t) is small, so that rounding alone moves g0 by up to ~2e-7: looser bound there, file precision elsewhere
    edge = (tfrac > 164.5 / 365.0 + 1e-8) & (tfrac < 199.5 / 365.0 + 1e-8)
    assert 60 <= edge.sum() <= 72
    for j, name in enumerate(["Q", "g", "cosg", "sing", "g0"]):
        e = np.abs(Yh[:, j] - Y[:, j]) / np.maximum(np.abs(Y[:, j]), 1.0)
        season = (tfrac > 106.5 / 365.0) & ~edge & (tfrac < 0.5)  # rising limb: dg0/dt ~ 10/year x 5e-10
        assert e[~edge & ~season].max() <= 2e-9, (name, e[~edge & ~season].max())  # no vegetation: Q = Q0
        assert e[season].max() <= 3e-8, (name, e[season].max())
        assert e[edge].max() <= 3e-7, (name, e[edge].max())
    assert (Yh[:, 1] > 0).sum() > 100  # the growth season is really exercised
    # the degree-day sum of Apply_growth against the file's Tcumul column at the two crossings it selects
    tc, t = np.asarray(d["Tcumul"]), np.asarray(d["time"])
    yr0 = np.floor(t) == 0
    tb = t[yr0][np.argmax(tc[yr0] >= 10.0)]
    te = t[yr0][np.argmax(tc[yr0] >= 200.0)]
    g0 = np.asarray(d["g0"])
    assert g0[yr0][t[yr0] <= tb].max() == 0.0 and abs(g0[yr0][t[yr0] == te][0] - 1.0) < 1e-9
    assert abs(dpar[0] - 10.0 * np.sqrt(0.001) * 35.0) < 1e-12


def test_vegetation_newton_root_is_the_root():
    """bending branch (chi < 0): the restated znewton returns the root of Vegetation_fNewton (:559) to rounding"""
    from workloads import synth

    prob, truth = synth.vegetation(nobs=300, growth="Growth_Yin1_temporal")
    o = Oracle(prob)
    th = truth[:prob.nFit]
    Yh, feas, _ = o.apply_model(prob.X, th)
    assert feas.all()
    B, S0, n1, b1, c1, chi, U = th[:7]
    y = prob.X[:, 1] - b1
    g = Yh[:, 1]
    sel = (y > 0) & (g > 0)
    Q0 = B * np.sqrt(S0) / n1 * y[sel] ** c1
    gam = g[sel] * y[sel] ** (1.0 / 3.0) / (8 * 9.81 * n1 * n1)
    gam2 = B**chi * y[sel] ** chi * U**chi
    x = Yh[sel, 0]
    f = x * x + gam * x * x * np.minimum(1.0, x**chi / gam2) - Q0 * Q0
    assert sel.sum() > 50 and np.max(np.abs(f) / (Q0 * Q0)) < 1e-14
    assert (np.minimum(1.0, x**chi / gam2) < 1.0).any() and (x < Q0).all()


def test_baratinbac_is_consistent_with_baratin():
    """Two parameterisations of the same rating curve, restated from two different reference modules
    (BaRatin_model.f90 k-a-c, BaRatinBAC_model.f90 b-a-c): feeding BaRatin's derived offsets b back into BaRatinBAC
    must return the original activation stages k (root of the continuity equation for the REPLACED control, b itself
    for the ADDED one) and the same discharges."""
    from bam_b200.capi import Problem

    for name, ktypes in (("baratin", [0, 3]), ("baratin_spd", None)):
        prob, fx = problem_from_fixture(name, with_copula=False)
        o = Oracle(prob)
        x = np.asarray(fx["start"], dtype=np.float64)
        _, feas, _, dpar = o.logpost(x.reshape(1, -1), want_dpar=True)
        assert feas[0] == 1
        nc = prob.ntheta // 3
        if name == "baratin_spd":  # first VAR combination: k1_1, k2_1
            th = np.array([x[0], x[5], x[6], x[7], x[12], x[13], x[14], x[15], x[16]])
            b = dpar[0][:3]
        else:
            th, b = x[:6], dpar[0][:2]
        k, a, c = th[0::3], th[1::3], th[2::3]
        bac_theta = np.column_stack([b, a, c]).flatten()
        H = np.linspace(float(k[0]) - 0.5, 4.0, 200)
        M = np.asarray(prob.xtra_i)
        pb = Problem("BaRatinBAC", H, np.zeros_like(H), partype=[0] * (3 * nc), priors_theta=[("FlatPrior", [])] * (3 * nc),
                     sigma_func=["Constant"], priors_gamma=[("FlatPrior", [])], xtra_i=M, xtra_d=[50.0])
        pk = Problem("BaRatin", H, np.zeros_like(H), partype=[0] * (3 * nc), priors_theta=[("FlatPrior", [])] * (3 * nc),
                     sigma_func=["Constant"], priors_gamma=[("FlatPrior", [])], xtra_i=M)
        Yb, fb, dpb = Oracle(pb).apply_model(H, bac_theta)
        Yk, fk, dpk = Oracle(pk).apply_model(H, th)
        assert fb.all() and fk.all()
        assert np.allclose(dpb[:nc], k, rtol=1e-11, atol=1e-12), (dpb[:nc], k)
        assert np.allclose(dpk[:nc], b, rtol=0, atol=0)
        # discharges agree except within rounding of a stage boundary
        away = np.min(np.abs(H[:, None] - k[None, :]), axis=1) > 1e-9
        assert np.allclose(Yb[away, 0], Yk[away, 0], rtol=1e-12, atol=1e-12)
        if ktypes is not None:
            assert list(dpb[nc:2 * nc]) == ktypes


def _mixture_numpy(X, th, nS, nEvt, nElt, missing):
    """Independent numpy restatement of Mixture_Apply (Mixture_model.f90:110-133), pack() pairing included."""
    evt = np.rint(X[:, 0]).astype(int)
    Y, last = np.zeros(nEvt * nElt), np.zeros(nEvt)
    for j in range(nEvt):
        w = th[j * nS:(j + 1) * nS] if missing else np.append(th[j * (nS - 1):(j + 1) * (nS - 1)], 0.0)
        if not missing:
            w[-1] = 1.0 - np.sum(w[:-1])
        last[j] = 1.0 - np.sum(w) if missing else w[-1]
        rows = np.flatnonzero(evt == j + 1)
        Y[j * nElt:(j + 1) * nElt] = X[rows, 2:] @ w + (last[j] * th[nEvt * nS:] if missing else 0.0)
    return Y, last


def test_mixture_on_the_reference_workspace():
    """tests/BaM_Mixture (5 sources, 27 events, 13 elements, missing source: 148 weights): the workspace ships no
    known-answer column, so the restatement is pinned on its DATA through an independent numpy statement of
    Mixture_Apply, at the configured starting point and at random feasible weights."""
    prob, fx = problem_from_fixture("mixture")
    assert (prob.nobs, prob.nX, prob.ntheta) == (351, 7, 148) and list(fx["xtra_i"]) == [5, 27, 13, 1]
    o = Oracle(prob)
    assert o.sizes.nDpar == 27 and o.sizes.nInfer == 150
    X = np.asarray(fx["X"]).reshape(7, 351).T
    rng = np.random.default_rng(5)
    th0 = np.asarray(fx["start"][:148])
    th1 = np.concatenate([(rng.dirichlet(np.ones(6), 27)[:, :5]).ravel(), rng.uniform(0.5, 2.0, 13)])
    for th in (th0, th1):
        Y, feas, dpar = o.apply_model(X, th)
        Yn, last = _mixture_numpy(X, th, 5, 27, 13, True)
        assert feas.all()
        assert np.max(np.abs(Y[:, 0] - Yn) / np.abs(Yn)) <= 4e-16 * 8
        assert np.max(np.abs(dpar[:27] - last)) <= 2e-16
    r = o.logpost_one(fx["start"])
    assert r["feas"] == 1 and r["isnull"] == 0 and np.isfinite(r["fx"])
    # infeasible weights: a weight above one, a negative weight, weights of one event summing above one (:112,120)
    for pos, val in ((3, 1.2), (7, -0.1), (10, 0.9)):
        th = th0.copy()
        th[pos] = val
        if pos == 10:
            th[11] = 0.9
        assert not o.apply_model(X, th)[1].any()
        x = np.asarray(fx["start"], dtype=np.float64).copy()
        x[:148] = th
        rr = o.logpost_one(x)
        assert rr["feas"] == 0 or rr["isnull"] == 1  # Uniform(0,1) prior: null before the model is run


@pytest.mark.parametrize("missing", [True, False])
def test_mixture_row_pairing(missing):
    """rows interleaved across events: output row (j-1)*nElt+e is the e-th INPUT row of event j (pack order)"""
    from workloads import synth

    prob, truth = synth.mixture(nS=3, nEvt=5, nElt=7, missing=missing, shuffle=True)
    X = np.asarray(prob.X)
    th = truth[:prob.ntheta]
    Y, feas, dpar = Oracle(prob).apply_model(X, th)
    Yn, last = _mixture_numpy(X, th, 3, 5, 7, missing)
    assert feas.all() and np.max(np.abs(Y[:, 0] - Yn)) <= 1e-15 and np.max(np.abs(dpar[:5] - last)) <= 2e-16
    if not missing:  # derived last weight outside [0,1]
        bad = th.copy()
        bad[0], bad[1] = 0.7, 0.6
        assert not Oracle(prob).apply_model(X, bad)[1].any()


def test_sfd_state_variable():
    """the transition stage kappa (SFD's state) solves fNewton_Val_General = 0 (StageFallDischarge_model.f90:253-283)
    and is undefRN on the rows that never solve for it (:126,148)"""
    from workloads import synth

    prob, truth = synth.sfd(nobs=120)
    o = Oracle(prob)
    assert o.sizes.nState == 1
    Xt, Ys, St, feas = o.calibration_run_states(truth)
    th = truth[:8]
    k = St[:, 0]
    solved = (k != -999999999.0) & (k != -666.666)
    assert solved.sum() > 60
    h2 = np.asarray(prob.X)[:, 1]
    lhs = th[0] * (k[solved] - th[1]) ** th[2] * np.sqrt((k[solved] - h2[solved] - th[6]) / th[3])
    rhs = th[5] * (k[solved] - th[4]) ** th[7]
    assert np.max(np.abs(lhs - rhs) / rhs) < 1e-10
    h1 = np.asarray(prob.X)[:, 0]
    assert ((h1 - th[1] <= 0) | (h2 - th[4] <= 0) | (h1 - h2 - th[6] <= 0))[~solved].all()


def test_sfdtidal_qmec_recursion():
    """time-recursive model (row f4): the restatement of SFDTidal_Qmec_Apply against an independent numpy run of the same
    scheme (first-order start, second-order Adams-Bashforth; bam_b200/synth.py), its three states, and the exits"""
    from workloads import synth

    prob, truth = synth.sfdtidal_qmec(nobs=400, missing_every=1)
    o = Oracle(prob)
    assert o.sizes.nState == 3 and o.P == 7
    Xt, Ys, St, feas = o.calibration_run_states(truth)
    assert feas and St.shape == (400, 3) and (St[0] == 0.0).all() and Ys[0, 0] == 300.0
    # the synthetic observations are the numpy run + noise: recover it through the noise-free part
    dt = 60.0
    assert np.allclose(np.diff(Ys[:, 0]), St[1:].sum(axis=1), rtol=1e-9, atol=1e-9)  # Q(m) - Q(m-1) = dt (pg + bf + ad)
    th = np.array([2.0, 1800.0, -6.0, 0.02, 4e-6, 0.0, 0.05, 4.0 / 3.0, 9.81, 300.0, 4000.0, 60.0])
    h1, h2 = np.asarray(prob.X)[:, 0], np.asarray(prob.X)[:, 1]
    width = th[1] / (th[0] - th[2])
    ym = (h1 + th[5] + h2 + th[6]) / 2
    Ae = width * (ym - th[2])
    pg1 = -th[8] * Ae[0] * ((h2[0] + th[6]) - (h1[0] + th[5]) + th[3]) / th[10]
    assert np.isclose(St[1, 0], dt * pg1, rtol=1e-13)
    # infeasible: non-positive area / friction coefficient, a bed above the water (negative wetted area somewhere)
    for pos, val in ((0, -5.0), (3, -1e-6), (1, 3.0)):  # fitted vector positions: A0, phi, be
        x = truth.copy()
        x[pos] = val
        assert o.logpost_one(x)["feas"] == 0


def test_sfdtidal_qmec2_on_the_reference_workspace():
    """tests/BaM_SFDTidal_Qmec2 (first 8000 of 14401 one-minute steps, 128 gauged): the workspace ships no simulated
    column, but its ADCP discharges give a physical check of the restatement -- run at the configured initial guesses, the
    recursion started 7745 steps earlier arrives at the gauged tidal discharge within a factor 1.6"""
    prob, d = problem_from_fixture("sfdtidal_qmec2")
    o = Oracle(prob)
    assert (prob.nobs, prob.ntheta, o.P, o.sizes.nState) == (8000, 11, 6, 3)
    x0 = np.asarray(d["start"])
    Xt, Ys, St, feas = o.calibration_run_states(x0)
    Y = np.asarray(d["Y"])
    ok = Y != -9999
    assert feas and ok.sum() == 128
    ratio = Ys[ok, 0] / Y[ok]  # same sign and magnitude over the gauged window (prior guesses, not a fit)
    assert ratio.min() > 0.5 and ratio.max() < 1.6
    assert np.allclose(np.diff(Ys[:, 0]), St[1:].sum(axis=1), rtol=1e-9, atol=1e-6)
    r = o.logpost_one(x0)
    assert r["feas"] == 1 and np.isfinite(r["fx"])


def test_algae_biomass_recursion():
    """time-recursive model (row f4): the restatement of AlgaeBiomass_Apply (AlgaeBiomass_model.f90:55-197) against an
    independent numpy run of the same daily budget (workloads/synth.py), its five state variables, the clamps and the
    exits.  The shipped workspace (tests/BaM_AlgaeBiomass: 17 parameters, one output) predates the code (18 parameters:
    katt + kcw instead of one attenuation; two outputs), so nothing in the reference pins this model beyond its source."""
    from workloads import synth

    prob, truth = synth.algae_biomass(nobs=500, missing_every=1)
    o = Oracle(prob)
    assert o.sizes.nState == 5 and o.P == 15
    Xt, Ys, St, feas = o.calibration_run_states(truth)
    assert feas and St.shape == (500, 5) and Ys.shape == (500, 2)
    th = np.array([1.0, 0.3, 0.001, 1.0, 0.12, 6.0, 16.5, 30.0, 150.0, 0.02, 0.5, 0.02, 9.81, 1000.0, 1e-4, 0.8, 1.5, 0.05])
    X = np.asarray(prob.X)
    assert np.allclose(Ys[:, 1], Ys[:, 0] / th[3], rtol=1e-15)
    assert (St[:, 3] == 1.0).all()                                      # Fn
    assert np.allclose(St[1:, 0], 1.0 - Ys[:-1, 0] / th[3], rtol=1e-14)  # Fb(t) = 1 - b(t-1) / bmax
    assert St[0, 0] == 1.0 - th[1] / th[3]
    tau = X[:, 2] * th[12] * th[14] * th[13]
    assert ((St[:, 4] > 0) == (tau > th[15])).all() and (St[:, 4] > 0).any()   # scour only above the critical shear stress
    cold = (X[:, 0] <= th[5]) | (X[:, 0] >= th[7])
    assert (St[cold, 1] == 0.0).all() and cold.any() and (St[~cold, 1] > 0).all()
    # the budget itself, step by step
    b_prev = np.concatenate([[th[1]], Ys[:-1, 0]])
    want = b_prev + th[4] * St[:, 0] * St[:, 1] * St[:, 2] * b_prev * th[0] - (th[11] + X[:, 4] + St[:, 4]) * (b_prev - th[2]) * th[0]
    assert np.allclose(Ys[:, 0], np.clip(want, th[2], th[3]), rtol=1e-13)
    # infeasible parameter sets (:140-146): b0 > bmax, Topt outside (Tmin, Tmax), tau0 <= 0
    for pos, val in ((0, 2.0), (4, 40.0), (10, -0.1)):  # fitted positions: b0, Topt, tau0
        x = truth.copy()
        x[pos] = val
        assert o.logpost_one(x)["feas"] == 0
    # input errors are fatal, as in the reference (err = 1, :79-99)
    import copy
    bad = copy.deepcopy(prob)
    bad.X[7, 4] = 1.5  # removal > 1
    with pytest.raises(Exception):
        Oracle(bad).logpost_one(truth)


def test_dynamic_vegetation():
    """row f4: DynamicVegetation_Apply (DynamicVegetation_model.f90:74-191) = AlgaeBiomass on depth = max(stage - b1, 0) plus,
    per time step, the root of Vegetation_fNewton on [0, Q0].  Checked against the pieces: the biomass rows equal an
    AlgaeBiomass run on the clamped depth, the discharge satisfies the Newton equation, the special cases (:150-153)."""
    from bam_b200.capi import Problem
    from workloads import synth

    prob, truth = synth.dynamic_vegetation(nobs=400, missing_every=1)
    o = Oracle(prob)
    assert o.sizes.nState == 5 and o.P == 19 + 4
    Xt, Ys, St, feas = o.calibration_run_states(truth)
    assert feas and Ys.shape == (400, 3) and St.shape == (400, 5)
    X = np.asarray(prob.X)
    B, n1, b1, c1, chi, U = truth[13:19]
    g, S0 = 9.81, 1e-4
    depth = np.maximum(X[:, 2] - b1, 0.0)
    # biomass rows = AlgaeBiomass on the clamped depth
    pa, ta = synth.algae_biomass(nobs=400, missing_every=1)
    Xa = X.copy(); Xa[:, 2] = depth
    pa2 = Problem("AlgaeBiomass", Xa, np.zeros((400, 2)), Yu=np.full((400, 2), 0.01), partype=[int(v) for v in pa.partype],
                  priors_theta=[("Gaussian", [0, 1])] * 13, fix_value=pa.fix_value, sigma_func=["Constant", "Constant"],
                  priors_gamma=[("Uniform", [0, 10])] * 2)
    _, Ya, Sa, fa = Oracle(pa2).calibration_run_states(np.concatenate([truth[:13], [0.02, 0.02]]))
    assert fa and np.array_equal(Ya, Ys[:, 1:]) and np.array_equal(Sa, St)
    # discharge: 0 on a dry bed, Q0 without biomass (none here: bmin > 0), else the root of Vegetation_fNewton
    dry = depth <= 0.0
    assert dry.sum() >= 5 and (Ys[dry, 0] == 0.0).all()
    y = depth[~dry]
    Q0 = (B * np.sqrt(S0) / n1) * y**c1
    gam = Ys[~dry, 1] * y ** (1.0 / 3.0) / (8.0 * g * n1**2)
    gam2 = (B * y * U) ** chi
    Q = Ys[~dry, 0]
    assert ((Q > 0) & (Q < Q0)).all()
    m = np.minimum(Q**chi / gam2, 1.0)
    assert np.max(np.abs(Q**2 * (1.0 + gam * m) - Q0**2) / Q0**2) < 1e-9
    # a Newton budget of 2 evaluations cannot converge: every wet row is infeasible -> infeasible vector
    tight = Problem("DynamicVegetation", X, np.asarray(prob.Y), Yu=np.asarray(prob.Yu), partype=[int(v) for v in prob.partype],
                    priors_theta=[("Gaussian", [0, 1])] * 19, fix_value=prob.fix_value, sigma_func=["Linear", "Constant", "Constant"],
                    priors_gamma=[("Uniform", [0, 10])] * 4, xtra_d=[1.0, 1.0, 1e-10, 1e-12], xtra_i=[2])
    assert Oracle(tight).logpost_one(truth)["feas"] == 0
    x = truth.copy(); x[0] = 2.0  # b0 > bmax (AlgaeBiomass :140-146)
    assert o.logpost_one(x)["feas"] == 0


def test_sfdtidal_qmec0_and_qmecms():
    """row f4: SFDTidal_Qmec0 (SFDTidal_Qmec0_model.f90:61-178) is SFDTidal_Qmec2's recursion with stages MINUS datums and
    Ae = Be (he + hm): the two restatements must agree bit for bit under that change of variables; SFDTidal_QmecMS
    (SFDTidal_QmecMS_model.f90:62-154) against its closed form in numpy; the exits of both."""
    from bam_b200.capi import Problem
    from workloads import synth

    prob, truth = synth.sfdtidal_variant("SFDTidal_Qmec0", nobs=500, missing_every=1)
    o = Oracle(prob)
    assert o.sizes.nState == 0 and o.P == 5 + 2
    _, Ys, feas = o.calibration_run(truth)
    assert feas and np.isfinite(Ys).all() and np.ptp(Ys[:, 0]) > 100.0
    Be, he, dzeta, ne, Q0 = truth[:5]
    fix = np.asarray(prob.fix_value)
    th2 = [Be, -he, dzeta, ne, -fix[4], -fix[5], fix[6], fix[7], Q0, fix[9], fix[10]]  # Qmec2: (Be, bevs, delta, ne, d1, d2, c, g, Q0, dx, dt)
    p2 = Problem("SFDTidal_Qmec2", np.asarray(prob.X), np.asarray(prob.Y), Yu=np.asarray(prob.Yu), partype=[-1] * 11, priors_theta=[],
                 fix_value=th2, sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000])] * 2)
    _, Y2, f2 = Oracle(p2).calibration_run(np.array([10.0, 0.02]))
    assert f2 and np.array_equal(Y2, Ys)
    for pos, val in ((1, -1.0), (2, 0.01), (3, 0.0)):  # he <= 0, dzeta >= 0, ne <= 0 (:96-98)
        x = truth.copy(); x[pos] = val
        assert o.logpost_one(x)["feas"] == 0
    x = truth.copy(); x[1] = 0.5  # he so small that the wetted area turns negative at low tide (:113-115)
    assert o.logpost_one(x)["feas"] == 0

    prob, truth = synth.sfdtidal_variant("SFDTidal_QmecMS", nobs=500, missing_every=1)
    o = Oracle(prob)
    assert o.sizes.nState == 0 and o.P == 4 + 2
    _, Ys, feas = o.calibration_run(truth)
    A0, be, delta, phi = truth[:4]
    fix = np.asarray(prob.fix_value)
    y0, d1, d2, c, g, dx = fix[0], fix[5], fix[6], fix[7], fix[8], fix[9]
    X = np.asarray(prob.X)
    width = A0 / (y0 - be)
    R0 = A0 / (width + 2 * (y0 - be))
    ne = np.sqrt(phi / (8 * g) * A0 * R0**c)
    y1, y2 = X[:, 0] + d1, X[:, 1] + d2
    Ae = width * ((y1 + y2) / 2 - be)
    Rh = Ae / (width + 2 * ((y1 + y2) / 2 - be))
    sl = (y2 - y1) + delta
    want = np.where(sl > 0, -1.0, 1.0) * (1 / ne) * Rh ** (0.5 * c) * Ae * np.sqrt(np.abs(sl / dx))
    assert feas and np.allclose(Ys[:, 0], want, rtol=1e-13) and (want > 0).any() and (want < 0).any()
    x = truth.copy(); x[0] = -5.0  # A0 <= 0
    assert o.logpost_one(x)["feas"] == 0


def test_sfdtidal_lagged_family():
    """row f4: SFDTidal / SFDTidal2 / SFDTidal4 / SFDTidalJones against vectorised numpy restatements of their formulas
    (the lag clamp, SFDTidal4's window rule at the ends of the series, the dry-bed zero) and their parameter exits."""
    from workloads import synth

    def lagged(a, Dt):
        n = a.size
        return a[np.clip(np.arange(n) - Dt, 0, n - 1)]

    for model in ("SFDTidal", "SFDTidal2", "SFDTidal4", "SFDTidalJones"):
        prob, truth = synth.sfdtidal_lag(model, nobs=500, missing_every=1)
        o = Oracle(prob)
        assert o.sizes.nState == 0
        _, Ys, feas = o.calibration_run(truth)
        X, fix = np.asarray(prob.X), np.asarray(prob.fix_value)
        h1, h2 = X[:, 0], X[:, 1]
        Dt = int(fix[0])
        sg = lambda x: np.where(x < 0, -1.0, 1.0)  # noqa: E731
        if model == "SFDTidal":
            K, B1, B2, h0, M, dz = truth[:6]; L = fix[6]
            fall = lagged(h1, Dt) - lagged(h2, Dt) - dz
            want = sg(fall) * K * (B2 + B1) / 2 * np.sqrt(np.abs(fall) / L) * ((lagged(h1, Dt) + lagged(h2, Dt)) / 2 - h0) ** M
            bad = (1, -1.0)
        elif model == "SFDTidal2":
            K, B, h0, M, delta, alpha = truth[:6]; L = fix[5]
            fall = h1 - alpha * lagged(h2, Dt) - delta
            want = sg(fall) * K * B * np.sqrt(np.abs(fall) / L) * (h1 - h0) ** M
            bad = (5, 1.0)   # alpha >= 1
        elif model == "SFDTidal4":
            K, B, h0, M, delta = truth[:5]; L, P = fix[5], int(fix[7])
            Sf = (lagged(h1, Dt) - lagged(h2, Dt) - delta) / L
            Sm = np.array([Sf[t - P:t + P + 1].mean() if (t - P >= 0 and t + P <= 499) else Sf[t] for t in range(500)])
            want = sg(Sm) * K * B * np.sqrt(np.abs(Sm)) * (h1 - h0) ** M
            bad = (0, 0.0)   # K <= 0
        else:
            K, B, h0, M, delta, alpha = truth[:6]; L = fix[5]
            Sf = (lagged(h1, Dt) - lagged(h2, Dt) - delta) / L
            want = sg(Sf) * K * B * np.sqrt(np.abs(Sf) * (1 - alpha * X[:, 2])) * (h1 - h0) ** M
            bad = (5, -1.0)  # alpha <= 0
        assert feas and np.allclose(Ys[:, 0], want, rtol=1e-12) and (want > 0).any() and (want < 0).any()
        x = truth.copy(); x[bad[0]] = bad[1]
        assert o.logpost_one(x)["feas"] == 0
        x = truth.copy(); x[2 if model != "SFDTidal" else 3] = 10.0  # bed above every stage: Q = 0 everywhere, still feasible
        _, Y0, f0 = o.calibration_run(x)
        assert f0 and (Y0 == 0.0).all()


def test_lagged_sfdtidal_on_the_reference_workspaces():
    """tests/BaM_SFDTidal2 (Saigon, 1372 steps, 70 gauged), BaM_SFDTidal4 (Seine, 2881 steps, 78 gauged) and
    BaM_SFDTidal_Jones (Saigon + dh/dt): the shipped workspaces whose parameter lists match the code.  At the configured
    initial guesses the restatement is feasible, equals a vectorised numpy evaluation of the formulas, and follows the
    ADCP discharges (correlation; prior guesses, not a fit)."""
    def lagged(a, Dt):
        n = a.size
        return a[np.clip(np.arange(n) - Dt, 0, n - 1)]

    sg = lambda x: np.where(x < 0, -1.0, 1.0)  # noqa: E731
    for name, nobs, gauged, mincorr in (("sfdtidal2", 1372, 70, 0.2), ("sfdtidal4", 2881, 78, 0.8), ("sfdtidal_jones", 1372, 70, 0.8)):
        prob, d = problem_from_fixture(name)
        o = Oracle(prob)
        x0 = np.asarray(d["start"])
        assert (prob.nobs, prob.ntheta, o.sizes.nState) == (nobs, 8, 0)
        _, Ys, feas = o.calibration_run(x0)
        Y = np.asarray(d["Y"]); ok = Y != -9999
        assert feas and ok.sum() == gauged and np.isfinite(Ys).all()
        r = o.logpost_one(x0)
        assert r["feas"] == 1 and np.isfinite(r["fx"])
        assert np.corrcoef(Ys[ok, 0], Y[ok])[0, 1] > mincorr
        # full theta (FIX values merged) and the formulas
        th, k = np.asarray(prob.fix_value, dtype=float).copy(), 0
        for i, pt in enumerate(prob.partype):
            if pt == 0:
                th[i] = x0[k]; k += 1
        X = np.asarray(prob.X)
        h1, h2, Dt = X[:, 0], X[:, 1], int(th[0])
        K, B, h0, M, L, delta = th[1:7]
        wet = h1 - h0 > 0
        with np.errstate(invalid="ignore"):
            if name == "sfdtidal2":
                fall = h1 - th[7] * lagged(h2, Dt) - delta
                want = sg(fall) * K * B * np.sqrt(np.abs(fall) / L) * (h1 - h0) ** M
            elif name == "sfdtidal4":
                P = int(th[7])
                Sf = (lagged(h1, Dt) - lagged(h2, Dt) - delta) / L
                Sm = np.array([Sf[t - P:t + P + 1].sum() / (2 * P + 1) if (t - P >= 0 and t + P <= nobs - 1) else Sf[t] for t in range(nobs)])
                want = sg(Sm) * K * B * np.sqrt(np.abs(Sm)) * (h1 - h0) ** M
            else:
                Sf = (lagged(h1, Dt) - lagged(h2, Dt) - delta) / L
                want = sg(Sf) * K * B * np.sqrt(np.abs(Sf) * (1 - th[7] * X[:, 2])) * (h1 - h0) ** M
        want = np.where(wet, want, 0.0)
        assert np.allclose(Ys[:, 0], want, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name,P", [("hydraulic_control_section", 3), ("orthorectification", 9), ("regression_x2y1", 4), ("sediment", 5),
                                    ("textfilemodel", 8), ("baratin_qbias", 9), ("baratin_qbiashbias", 10), ("baratinbac", 11),
                                    ("baratinbac_spd", 19), ("suspendedload", 5), ("sediment_camenen_x2", 6), ("regression_catchments", 5),
                                    ("regression_x6", 7), ("segmentation", 4)])
def test_remaining_shipped_workspaces_run(name, P):
    """the other shipped workspaces of models on the device (tests/BaM_HydraulicControl_section, BaM_Orthorectification,
    BaM_Regression_X2Y1, BaM_Sediment, BaM_TextFileModel, BaM_BaRatin_Qbias, BaM_BaRatin_QbiasHbias with its X and Y
    biases, BaM_BaRatinBAC, BaM_BaRatinBAC_SPD with VAR offsets, BaM_SuspendedLoad, BaM_Sediment_Camenen_X2,
    BaM_Regression_Catchments, BaM_Emeline (6 inputs), BaM_Segmentation) through the fixture reader and the
    oracle: feasible with a finite log-posterior at the configured starting point"""
    prob, d = problem_from_fixture(name)
    o = Oracle(prob)
    assert o.P == P
    r = o.logpost_one(np.asarray(d["start"]))
    assert r["feas"] == 1 and r["isnull"] == 0 and np.isfinite(r["fx"])
