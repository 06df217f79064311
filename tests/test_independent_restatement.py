"""A SECOND, independent restatement of the reference's log-posterior and envelope rule, in mpmath at 50 digits.

Written from the Fortran itself -- GetPriorLogPdf / GetLogLikelihood / GetHyperLogPdf / GetPostLogPdf
(src/BaM_tools.f90:3006-3382), UnfoldParVector (:3386-3463), ApplyModel_BaM (:4115-4177), Sigmafunk_Apply (:3521-3571),
Gcop_getV (:4246-4270), BaM_ProcessSpag (:1393-1426), ApplyContinuity / BaRatin_computeQ / GetHRange
(src/Models/BaRatin_model.f90:194-436), LM_Apply (src/Models/Linear_model.f90:96-99) and TxtMdl_Apply for the shipped
formula (tests/BaM_TextFile_InputUncertainty: a*T+b) -- and NOT from oracle/bam_oracle.c: different language, different
arithmetic (50 decimal digits, so the restatement has no rounding of its own), different data structures.  The C
oracle (what the GPU is compared with) must agree with it to double precision on the reference's own workspaces.

What this does and does not pin: it pins the ORACLE against a second reading of the vendored Fortran.  The
conventions of the un-vendored BMSL (Gaussian / Uniform / LogNormal log-densities and cdfs, the Hazen plotting
position of GetEmpiricalQuantile, n-1 in the standard deviation) are the same DECLARED ones; each is
one function below (`BMSL_*`), and test_declared_conventions_matter shows the comparison would fail if one were flipped,
i.e. that the agreement is not vacuous.
"""
import mpmath as mp
import numpy as np
import pytest

from oracle.oracle_api import Oracle
from tests.helpers import load_fixture, perturbed_vectors, problem_from_fixture

mp.mp.dps = 50
F = mp.mpf
HALF_LOG_2PI = mp.log(2 * mp.pi) / 2


# ---- declared BMSL conventions (un-vendored) -------------------------------------------------
def BMSL_logpdf(dist, par, x, uniform_open=False):
    """returns (logp, feas, isnull)"""
    if dist == "Gaussian":
        mu, sd = F(par[0]), F(par[1])
        if sd <= 0:
            return None, False, False
        return -HALF_LOG_2PI - mp.log(sd) - ((x - mu) / sd) ** 2 / 2, True, False
    if dist == "Uniform":
        a, b = F(par[0]), F(par[1])
        if b <= a:
            return None, False, False
        outside = (x <= a or x >= b) if uniform_open else (x < a or x > b)
        if outside:
            return None, True, True
        return -mp.log(b - a), True, False
    if dist == "LogNormal":
        mu, sd = F(par[0]), F(par[1])
        if sd <= 0:
            return None, False, False
        if x <= 0:
            return None, True, True
        return -HALF_LOG_2PI - mp.log(sd) - mp.log(x) - ((mp.log(x) - mu) / sd) ** 2 / 2, True, False
    if dist == "FlatPrior":
        return F(0), True, False
    raise NotImplementedError(dist)


def BMSL_cdf(dist, par, x):
    if dist == "Gaussian":
        return mp.ncdf((x - F(par[0])) / F(par[1]))
    if dist == "Uniform":
        a, b = F(par[0]), F(par[1])
        return min(max((x - a) / (b - a), F(0)), F(1))
    if dist == "LogNormal":
        return mp.ncdf((mp.log(x) - F(par[0])) / F(par[1])) if x > 0 else F(0)
    raise NotImplementedError(dist)


def BMSL_qnorm(u):
    return mp.sqrt(2) * mp.erfinv(2 * u - 1)


def BMSL_quantile_sorted(xs, p, rule="hazen"):
    """GetEmpiricalQuantile on a sorted sample: plotting position (i - 0.5) / n (Hazen), linear interpolation"""
    n = len(xs)
    pp = [(F(i) - (F(1) / 2 if rule == "hazen" else F(0))) / (n if rule == "hazen" else n + 1) for i in range(1, n + 1)]
    if p <= pp[0]:
        return xs[0]
    if p >= pp[-1]:
        return xs[-1]
    for i in range(1, n):
        if p <= pp[i]:
            return xs[i - 1] + (xs[i] - xs[i - 1]) * (p - pp[i - 1]) / (pp[i] - pp[i - 1])
    return xs[-1]


# ---- the vendored Fortran, restated ----------------------------------------------------------------------------
def sigmafunk(name, g, y):  # Sigmafunk_Apply :3521-3571
    ay = abs(y)
    if name == "Constant":
        return g[0]
    if name == "Linear":
        return g[0] + g[1] * ay
    if name == "Proportional":
        return g[0] * ay
    if name == "Power":
        return g[0] + g[1] * ay ** g[2]
    if name == "Exponential":
        return g[0] + (g[2] - g[0]) * (1 - mp.exp(-(ay / g[1])))
    if name == "Gaussian":
        return g[0] + (g[2] - g[0]) * (1 - mp.exp(-((ay / g[1]) ** 2)))
    raise NotImplementedError(name)


def baratin(theta, M, H):
    """BaRatin_Apply: theta = (k, a, c) per control; M[r][i] control matrix; returns (Q list, feas)"""
    nc = len(theta) // 3
    k, a, c = [theta[3 * j] for j in range(nc)], [theta[3 * j + 1] for j in range(nc)], [theta[3 * j + 2] for j in range(nc)]
    # ApplyContinuity :368-394
    if a[0] <= 0 or c[0] == 0:
        return None, False
    b = [k[0]] + [None] * (nc - 1)
    for j in range(1, nc):
        if k[j] <= k[j - 1]:
            return None, False
        if any(k[j] < b[i] for i in range(j)) or a[j] <= 0 or c[j] == 0:
            return None, False
        som = F(0)
        for i in range(j):
            som += (M[j - 1][i] - M[j][i]) * a[i] * (k[j] - b[i]) ** c[i]
        if som < 0:
            return None, False
        b[j] = k[j] - (som / a[j]) ** (1 / c[j])
    Q = []
    for h in H:  # BaRatin_computeQ :229-250, GetHRange :420-434
        rng = nc
        for i in range(nc):
            if h < k[i]:
                rng = i
                break
        q = F(0)
        for i in range(rng):
            if M[rng - 1][i] == 1:
                if h - b[i] < 0:
                    q = F(0)
                    break
                q += a[i] * (h - b[i]) ** c[i]
            elif h - b[i] < 0:
                return None, False
        Q.append(q)
    return Q, True


def log_posterior(d, x, uniform_open=False):
    """GetPostLogPdf on fixture `d` at vector x; returns (prior, lkh, hyper, feas, isnull) or flags"""
    n, nX, nY, nth = d["nobs"], d["nX"], d["nY"], len(d["partype"])
    col = lambda key, j: [F(v) for v in d[key][j * n:(j + 1) * n]]  # noqa: E731  column-major
    icol = lambda key, j: d[key][j * n:(j + 1) * n]  # noqa: E731
    x = [F(float(v)) for v in x]
    # UnfoldParVector :3386-3463 / Packer :4202-4242
    nFit = sum(nv for pt, nv in zip(d["partype"], d["nval"]) if pt != -1)
    nrem = {"Constant": 1, "Linear": 2, "Proportional": 1, "Power": 3, "Exponential": 3, "Gaussian": 3}
    pos = nFit
    gamma = []
    for j in range(nY):
        m = nrem[d["sigma_func"][j]]
        gamma.append(x[pos:pos + m])
        pos += m
    Xtrue = [col("X", j) for j in range(nX)]
    for j in range(nX):
        Xu = col("Xu", j)
        for i in range(n):
            if Xu[i] > 0:
                Xtrue[j][i] = x[pos]
                pos += 1
    assert all(max(icol("Xbindx", j)) == 0 for j in range(nX)) and all(max(icol("Ybindx", j)) == 0 for j in range(nY))
    assert pos == len(x)
    # GetPriorLogPdf :3062-3112: remnant parameters, then theta, then the copula
    prior = F(0)
    pg = d["priors_gamma"]
    kk = 0
    for j in range(nY):
        for m in range(len(gamma[j])):
            lp, feas, isnull = BMSL_logpdf(pg[kk][0], pg[kk][1], gamma[j][m], uniform_open)
            kk += 1
            if not feas or isnull:
                return None, None, None, feas, isnull
            prior += lp
    for q in range(nFit):
        lp, feas, isnull = BMSL_logpdf(d["priors_theta"][q][0], d["priors_theta"][q][1], x[q], uniform_open)
        if not feas or isnull:
            return None, None, None, feas, isnull
        prior += lp
    if d.get("priorC") is not None:
        k = int(round(len(d["priorC"]) ** 0.5))
        Cm = mp.matrix(k, k)
        for r in range(k):
            for c in range(k):
                Cm[r, c] = F(d["priorC"][r + c * k])
        Mm = Cm ** -1 - mp.eye(k)  # LoadBamObjects :398-414
        v = []
        for q in range(nFit):  # theta first, then gamma (:3084-3101)
            v.append(BMSL_qnorm(BMSL_cdf(d["priors_theta"][q][0], d["priors_theta"][q][1], x[q])))
        kk = 0
        for j in range(nY):
            for m in range(len(gamma[j])):
                v.append(BMSL_qnorm(BMSL_cdf(pg[kk][0], pg[kk][1], gamma[j][m])))
                kk += 1
        vv = mp.matrix(v)
        prior -= (vv.T * Mm * vv)[0] / 2  # :3111 (log-determinant deliberately omitted)
    # ApplyModel_BaM :4115-4177: full theta per observation (FIX from init, STD from the vector, VAR by index)
    off, q = [], 0
    for pt, nv in zip(d["partype"], d["nval"]):
        off.append(q)
        if pt != -1:
            q += nv
    vi = d["var_indx"]

    def full_theta(i):
        th = []
        for t in range(nth):
            if d["partype"][t] == -1:
                th.append(F(d["fix_value"][t]))
            elif d["partype"][t] == 0:
                th.append(x[off[t]])
            else:
                th.append(x[off[t] + vi[i + t * n] - 1])
        return th

    Yhat = [[None] * n for _ in range(nY)]
    if d["model_id"] == "BaRatin":
        nc = nth // 3
        M = [[d["xtra_i"][r + c * nc] for c in range(nc)] for r in range(nc)]
        for i in range(n):
            Q, feas = baratin(full_theta(i), M, [Xtrue[0][i]])
            if not feas:
                return None, None, None, False, False
            Yhat[0][i] = Q[0]
    elif d["model_id"] == "Linear":  # theta reshaped column-major to (nX, nY)
        for i in range(n):
            th = full_theta(i)
            for y in range(nY):
                Yhat[y][i] = sum(Xtrue[j][i] * th[y * nX + j] for j in range(nX))
    elif d["model_id"] == "TextFile":  # the shipped formula a*T+b
        assert "a*T+b" in d["xtra_text"]
        for i in range(n):
            th = full_theta(i)
            Yhat[0][i] = th[0] * Xtrue[0][i] + th[1]
    else:
        raise NotImplementedError(d["model_id"])
    # GetLogLikelihood :3204-3227
    lkh = F(0)
    Y, Yu = [col("Y", j) for j in range(nY)], [col("Yu", j) for j in range(nY)]
    for i in range(n):
        for j in range(nY):
            if Y[j][i] == F(d["mv"]):
                continue
            sig = sigmafunk(d["sigma_func"][j], gamma[j], Yhat[j][i])
            if sig <= 0:
                return None, None, None, False, False
            s = mp.sqrt(Yu[j][i] ** 2 + sig ** 2)
            lkh += -HALF_LOG_2PI - mp.log(s) - ((Y[j][i] - Yhat[j][i]) / s) ** 2 / 2
    # GetHyperLogPdf :3285-3303
    hyper = F(0)
    for j in range(nX):
        Xo, Xu = col("X", j), col("Xu", j)
        if not any(u > 0 for u in Xu):
            continue
        for i in range(n):
            if Xu[i] == 0:
                continue
            hyper += -HALF_LOG_2PI - mp.log(Xu[i]) - ((Xo[i] - Xtrue[j][i]) / Xu[i]) ** 2 / 2
    return prior, lkh, hyper, True, False


CASES = ["baratin", "baratin_spd", "regression_x2y2", "textfile_inputuncertainty"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_agrees_with_the_independent_restatement(name):
    prob, d = problem_from_fixture(name)
    x = perturbed_vectors(d["start"], 4, rel=0.01, seed=11)
    o = Oracle(prob)
    fx, feas, isnull = o.logpost(x)
    pr, lk, hy, _, _ = o.logpost_parts(x)
    for r in range(x.shape[0]):
        p2, l2, h2, f2, n2 = log_posterior(d, x[r])
        assert bool(feas[r]) == f2 and bool(isnull[r]) == n2
        if not f2 or n2:
            continue
        tot = p2 + l2 + h2
        # double-precision sums of ~100 terms against an exact evaluation: 1e-13 relative (the parity bar is 1e-12)
        assert abs(F(float(fx[r])) - tot) <= F("1e-13") * abs(tot), (name, r, float(fx[r]), float(tot))
        for got, want in ((pr[r], p2), (lk[r], l2), (hy[r], h2)):
            assert abs(F(float(got)) - want) <= F("1e-13") * max(abs(want), abs(tot) * F("1e-3")), (name, r, float(got), float(want))


def test_infeasible_and_null_flags_follow_the_fortran_order():
    prob, d = problem_from_fixture("baratin")
    o = Oracle(prob)
    x = np.tile(np.asarray(d["start"], dtype=np.float64), (3, 1))
    x[0, 3] = x[0, 0] - 0.1  # k2 <= k1: ApplyContinuity infeasible
    x[1, 6] = -1.0           # gamma1 outside its Uniform(0, 1000) prior: null density before anything else
    x[2, 1] = -5.0           # a1 <= 0
    fx, feas, isnull = o.logpost(x)
    for r in range(3):
        _, _, _, f2, n2 = log_posterior(d, x[r])
        assert bool(feas[r]) == f2 and bool(isnull[r]) == n2


def test_declared_conventions_matter():
    """flipping one declared BMSL convention breaks the agreement: the comparison above is not vacuous"""
    prob, d = problem_from_fixture("baratin")
    o = Oracle(prob)
    x = np.asarray(d["start"], dtype=np.float64).copy()
    x[7] = 0.0  # gamma2 on the closed end of Uniform(0, 1000): inside under the declared (closed) convention
    fx, feas, isnull = o.logpost(x)
    p2, l2, h2, f2, n2 = log_posterior(d, x)
    assert feas[0] and not isnull[0] and f2 and not n2
    assert abs(F(float(fx[0])) - (p2 + l2 + h2)) <= F("1e-13") * abs(p2 + l2 + h2)
    _, _, _, f3, n3 = log_posterior(d, x, uniform_open=True)  # the other convention: zero density on the boundary
    assert n3 and not isnull[0]


def test_envelope_rule_against_the_independent_restatement():
    """BaM_ProcessSpag :1393-1426: < 75 % bad values are dropped, else 7 x unfeasFlag; Hazen quantiles, mean, sd (n-1)"""
    prob, d = problem_from_fixture("baratin")
    o = Oracle(prob)
    rng = np.random.default_rng(8)
    mode = np.asarray(d["start"], dtype=np.float64)
    mc = mode * (1.0 + 0.03 * rng.standard_normal((61, mode.size)))
    mc[:9, 3] = -2.0  # 9 of 61 replications infeasible (k2 < k1): dropped
    H = np.linspace(-1.0, 5.0, 7)
    env, spag = o.predict(mc, mode, [H], 1, [0], want_spag=True)
    FLAG = F(d["unfeas_flag"])
    M = [[d["xtra_i"][r + c * 2] for c in range(2)] for r in range(2)]
    for t, h in enumerate(H):
        vals = []
        for r in range(mc.shape[0]):
            Q, feas = baratin([F(float(v)) for v in mc[r, :6]], M, [F(float(h))])
            vals.append(Q[0] if feas else FLAG)
        good = sorted(v for v in vals if v != FLAG)
        assert len(vals) - len(good) == 9
        want = [BMSL_quantile_sorted(good, F(p)) for p in ("0.5", "0.025", "0.975", "0.16", "0.84")]
        mean = sum(good) / len(good)
        want += [mean, mp.sqrt(sum((v - mean) ** 2 for v in good) / (len(good) - 1))]
        for s in range(7):
            got = F(float(env[0, s, t]))
            assert abs(got - want[s]) <= F("1e-12") * max(abs(want[s]), F("1e-300")), (t, s, float(got), float(want[s]))
        # the other plotting position (Weibull i / (n + 1)) gives different quantiles: the rule is not vacuous either
        if good[0] != good[-1]:
            assert BMSL_quantile_sorted(good, F("0.025"), rule="weibull") != want[1]
    mc2 = mc.copy()
    mc2[:50, 3] = -2.0  # 82 % infeasible: the row is flagged
    env2, _ = o.predict(mc2, mode, [H], 1, [0])
    assert (env2 == d["unfeas_flag"]).all()
