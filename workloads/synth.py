"""Synthetic workloads of the BASELINE.json configs: std::mt19937_64, seed 20261019.

These only build *inputs* (numpy); they contain no compute path.  Used by bench.py and by the
tests (so that the parity tests and the bench run the same shapes).
"""
from __future__ import annotations

import numpy as np

from workloads.mt64 import default_rng  # std::mt19937_64, the generator SURVEY.md 8(d) names
from bam_b200.capi import PARTYPE, SED_CO, SED_CSS, SED_FORMULA, SED_PPO, VEG_GROWTH, Problem

SEED = 20261019

# priors of tests/BaM_BaRatin_SPD (Config_Model.txt, Config_k1_VAR.txt, Config_k2_VAR.txt, Config_RemnantSigma.txt)
_SPD_K1_SD = [0.5, 0.62, 0.71, 0.8, 0.87]
_SPD_K2 = [1.05, 1.08, 1.11, 1.15, 1.19]
_SPD_CONTROL = [[1, 0, 0], [0, 1, 0], [0, 1, 1]]


def _baratin_q(H, k, a, c, M):
    """numpy BaRatin used only to synthesise observations."""
    nc = len(k)
    b = np.zeros(nc)
    b[0] = k[0]
    for j in range(1, nc):
        som = sum((M[j - 1][i] - M[j][i]) * a[i] * (k[j] - b[i]) ** c[i] for i in range(j))
        b[j] = k[j] - (som / a[j]) ** (1.0 / c[j])
    Q = np.zeros_like(H)
    rng_idx = np.sum(H[:, None] >= np.asarray(k)[None, :], axis=1)
    for r in range(1, nc + 1):
        sel = rng_idx == r
        for i in range(r):
            if M[r - 1][i]:
                Q[sel] += a[i] * np.maximum(H[sel] - b[i], 0.0) ** c[i]
    return Q


def baratin_spd(nobs: int = 100_000, copula: bool = False, seed: int = SEED):
    """cfg 2: BaRatin, 3 controls, 5 periods, VAR k1 and k2 (multi-period rating curves).

    Returns (Problem, truth vector of length 19)."""
    rng = default_rng(seed)
    nper = 5
    period = 1 + (nper * np.arange(nobs)) // max(nobs, 1)
    H = rng.uniform(-0.5, 3.0, nobs)
    a = [14.17, 26.52, 31.82]
    c = [1.5, 1.67, 1.67]
    k1, k3 = -0.6, 1.2
    gam = (0.1, 0.05)
    Q = np.zeros(nobs)
    for p in range(nper):
        sel = period == p + 1
        Q[sel] = _baratin_q(H[sel], [k1, _SPD_K2[p], k3], a, c, _SPD_CONTROL)
    uQ = 0.05 * Q + 0.01
    Qobs = Q + rng.standard_normal(nobs) * np.sqrt(uQ**2 + (gam[0] + gam[1] * Q) ** 2)
    ntheta = 9
    partype = [PARTYPE["VAR"], 0, 0, PARTYPE["VAR"], 0, 0, 0, 0, 0]
    nval = [nper, 1, 1, nper, 1, 1, 1, 1, 1]
    var_indx = np.ones((nobs, ntheta), dtype=np.int32)
    var_indx[:, 0] = period
    var_indx[:, 3] = period
    priors = ([("Gaussian", [-0.6, s]) for s in _SPD_K1_SD] + [("Gaussian", [14.17, 5.01]), ("Gaussian", [1.5, 0.025])]
              + [("Gaussian", [k, 0.3]) for k in _SPD_K2] + [("Gaussian", [26.52, 8.4]), ("Gaussian", [1.67, 0.025]),
                                                             ("Gaussian", [1.2, 0.2]), ("Gaussian", [31.82, 10.93]),
                                                             ("Gaussian", [1.67, 0.025])])
    priorM = None
    if copula:
        k = 19
        C = np.eye(k)
        for i in range(nper):
            for j in range(nper):
                if i != j:
                    C[i, j] = 0.8 ** abs(i - j)
                    C[7 + i, 7 + j] = 0.6 ** abs(i - j)
        priorM = np.linalg.inv(C) - np.eye(k)
    prob = Problem("BaRatin", H, Qobs, Yu=uQ, partype=partype, nval=nval, var_indx=var_indx, priors_theta=priors,
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000]), ("Uniform", [0, 1000])], priorM=priorM,
                   xtra_i=np.asarray(_SPD_CONTROL, dtype=np.int32).flatten(order="F"))
    truth = np.array([k1] * nper + [a[0], c[0]] + _SPD_K2 + [a[1], c[1], k3, a[2], c[2], gam[0], gam[1]])
    return prob, truth


def feasible_starts(truth, K: int, rel: float = 0.01, seed: int = SEED + 1, check=None):
    """K starting vectors = truth * (1 + rel * N(0,1)), redrawn until `check(x)` says feasible."""
    rng = default_rng(seed)
    truth = np.asarray(truth, dtype=np.float64)
    x = truth * (1.0 + rel * rng.standard_normal((K, truth.size)))
    if check is not None:
        for _ in range(50):
            bad = ~check(x)
            if not bad.any():
                break
            x[bad] = truth * (1.0 + rel * rng.standard_normal((int(bad.sum()), truth.size)))
    return x


def spd_start_check(x):
    """k2_p < k3 and k1 < k2 for every period (the only way a 1 % perturbation turns infeasible)."""
    x = np.asarray(x)
    return (x[:, 7:12].max(axis=1) < x[:, 14]) & (x[:, 0:5].max(axis=1) < x[:, 7:12].min(axis=1)) & (x[:, 17] > 0)


def linear_latent(nobs: int = 1_000_000, seed: int = SEED):
    """cfg 3: Linear nX=nY=2 with input uncertainty on every cell -> 2*nobs latent inputs."""
    rng = default_rng(seed)
    Xtrue = rng.standard_normal((nobs, 2))
    Xu = np.full((nobs, 2), 0.1)
    Xobs = Xtrue + rng.standard_normal((nobs, 2)) * Xu
    theta = np.array([1.0, -2.0, 0.5, 3.0])  # column-major (nX, nY)
    Yhat = Xtrue @ theta.reshape(2, 2, order="F")
    Yu = np.full((nobs, 2), 0.05)
    gam = (0.1, 0.02)
    Y = Yhat + rng.standard_normal((nobs, 2)) * np.sqrt(Yu**2 + (gam[0] + gam[1] * np.abs(Yhat)) ** 2)
    prob = Problem("Linear", Xobs, Y, Yu=Yu, Xu=Xu, partype=[0, 0, 0, 0],
                   priors_theta=[("Gaussian", [0, 10])] * 4, sigma_func=["Linear", "Linear"],
                   priors_gamma=[("Uniform", [0, 100])] * 4)
    truth = np.concatenate([theta, [gam[0], gam[1], gam[0], gam[1]], Xtrue.flatten(order="F")])
    return prob, truth


def textfile_latent(nobs: int = 100_000, seed: int = SEED):
    """cfg 3 twin: TextFile 'a*T+b' with input uncertainty (tests/BaM_TextFile_InputUncertainty shape)."""
    rng = default_rng(seed)
    T = rng.uniform(5.0, 50.0, nobs)
    Tu = np.full(nobs, 1.0)
    Tobs = T + rng.standard_normal(nobs) * Tu
    a, b = 2.0, 1.0
    Yhat = a * T + b
    Yu = np.full(nobs, 0.5)
    gam = (0.5, 0.05)
    Y = Yhat + rng.standard_normal(nobs) * np.sqrt(Yu**2 + (gam[0] + gam[1] * np.abs(Yhat)) ** 2)
    text = "1\nT\n2\na,b\n1\na*T+b   ! concentration\n"
    prob = Problem("TextFile", Tobs, Y, Yu=Yu, Xu=Tu, partype=[0, 0], priors_theta=[("Gaussian", [0, 10])] * 2,
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 100])] * 2, xtra_text=text)
    truth = np.concatenate([[a, b, gam[0], gam[1]], T])
    return prob, truth


def sediment_camenen(nobs: int = 812, seed: int = SEED):
    """cfg 4: Sediment Camenen + Soulsby, d and s as inputs (tests/BaM_Sediment_Camenen_X4 options)."""
    rng = default_rng(seed)
    h = rng.uniform(0.02, 0.5, nobs)
    S0 = rng.uniform(1e-3, 8e-3, nobs)
    d = rng.choice([3.05e-4, 5.06e-4, 1.0e-3], nobs)
    s = np.full(nobs, 2.65)
    X = np.column_stack([h, S0, d, s])
    theta = np.array([12.0, 1.5, 4.5])
    v, g = 1e-6, 9.81
    SD = (g * (s - 1) / v**2) ** (1 / 3) * d
    css = (0.3 / (1 + 1.2 * SD)) * 0.055 * (1 - np.exp(-0.02 * SD))
    tau = h * S0 / ((s - 1) * d)
    q = np.sqrt(g * (s - 1) * d**3) * theta[0] * tau ** theta[1] * np.exp(-theta[2] * css / tau)
    q[tau <= css] = 0.0
    gam = (1e-4, 0.1)
    Y = q + rng.standard_normal(nobs) * (gam[0] + gam[1] * q)
    prob = Problem("Sediment", X, Y, partype=[0, 0, 0],
                   priors_theta=[("Gaussian", [12, 6]), ("Gaussian", [1.5, 0.5]), ("Gaussian", [4.5, 2])],
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1]), ("Uniform", [0, 10])],
                   xtra_i=[SED_FORMULA["Camenen"], SED_CSS["Soulsby"], SED_CO["FIX"], SED_PPO["IN"], SED_PPO["IN"],
                           SED_PPO["FIX"], SED_PPO["FIX"]], xtra_d=[5.06e-4, 2.65, 1e-6, 9.81])
    truth = np.concatenate([theta, gam])
    return prob, truth


def sediment_posterior(Nrep: int, seed: int = SEED + 4):
    """cfg 4 posterior sample: theta ~ N((12,1.5,4.5), diag(1,0.05,0.3)^2) truncated > 0, gamma ~ |N|."""
    rng = default_rng(seed)
    th = np.abs(np.array([12.0, 1.5, 4.5]) + rng.standard_normal((Nrep, 3)) * np.array([1.0, 0.05, 0.3]))
    gm = np.abs(np.array([1e-4, 0.1]) + rng.standard_normal((Nrep, 2)) * np.array([2e-5, 0.01]))
    return np.column_stack([th, gm])


def baratin_2c(nobs: int = 38, seed: int = SEED):
    """cfg 1/5 model: 2-control BaRatin with the priors of tests/BaM_BaRatin; synthetic gaugings."""
    rng = default_rng(seed)
    H = rng.uniform(-0.3, 4.0, nobs)
    k, a, c = [-0.5, 1.5], [53.0, 144.0], [1.5, 1.67]
    Q = _baratin_q(H, k, a, c, [[1, 0], [0, 1]])
    uQ = 0.05 * Q + 0.1
    gam = (1.0, 0.1)
    Qobs = Q + rng.standard_normal(nobs) * np.sqrt(uQ**2 + (gam[0] + gam[1] * Q) ** 2)
    priors = [("Gaussian", [-0.5, 0.25]), ("Gaussian", [53, 10]), ("Gaussian", [1.5, 0.025]), ("Gaussian", [1.5, 0.5]),
              ("Gaussian", [144, 45]), ("Gaussian", [1.67, 0.025])]
    prob = Problem("BaRatin", H, Qobs, Yu=uQ, partype=[0] * 6, priors_theta=priors, sigma_func=["Linear"],
                   priors_gamma=[("Uniform", [0, 1000]), ("Uniform", [0, 1000])], xtra_i=[1, 0, 0, 1])
    truth = np.array([k[0], a[0], c[0], k[1], a[1], c[1], gam[0], gam[1]])
    return prob, truth


def baratin_spaghetti(Nrep: int = 1_000_000, nobs: int = 100_000, nspag: int = 1, seed: int = SEED + 5):
    """cfg 5: Nrep posterior rows ~ N(maxpost, (2 % of value)^2), feasibility filtered; stage series
    H_t = 0.5 + 1.5 sin^2(pi t / 5000) + 0.05 N(0,1); optionally nspag noisy copies."""
    rng = default_rng(seed)
    maxpost = np.array([-0.5, 53.0, 1.5, 1.5, 144.0, 1.67, 1.0, 0.1])
    mc = maxpost * (1.0 + 0.02 * rng.standard_normal((Nrep, maxpost.size)))
    mc[:, 6:] = np.abs(mc[:, 6:])
    t = np.arange(nobs)
    H = 0.5 + 1.5 * np.sin(np.pi * t / 5000.0) ** 2 + 0.05 * rng.standard_normal(nobs)
    if nspag > 1:
        Hs = H[:, None] + 0.01 * rng.standard_normal((nobs, nspag))
    else:
        Hs = H.reshape(-1, 1)
    return mc, maxpost, Hs


def vegetation(nobs: int = 600, growth: str = "Growth_Yin1_temporal", nyear: int = 3, var_gmax: bool = False,
               nY: int = 5, seed: int = SEED + 6):
    """Vegetation-affected rating curve (tests/BaM_Vegetation_static shape; theta with U_chi as the CODE expects,
    Vegetation_model.f90:40-41,156-181).  Inputs: time (years), stage [, raw and smoothed temperature].
    Outputs: Q (observed with noise), g / cos / sin / g0 (g observed, the rest missing).  Returns (Problem, truth)."""
    rng = default_rng(seed)
    temporal = growth == "Growth_Yin1_temporal"
    time = np.sort(rng.uniform(0.0, nyear, nobs))
    for y in range(nyear):  # make sure every year starts and ends with a point
        time[np.searchsorted(time, y)] = y + 0.002
        time[np.searchsorted(time, y + 1) - 1] = y + 0.998
    tt = time - np.floor(time)
    h = 1.6 + 0.8 * np.sin(2 * np.pi * (tt - 0.1)) ** 2 + 0.2 * rng.random(nobs)
    Ts = 10.0 + 10.0 * np.sin(2 * np.pi * (tt - 0.25))
    Tr = Ts + rng.standard_normal(nobs)
    base = [25.0, 0.001, 0.02, 1.0, 1.67, -0.5, 0.5]
    if temporal:
        gpar, ng = [0.2, 0.45, 0.7], 1
        X = np.column_stack([time, h])
    else:
        gpar = {"Growth_Yin1": [5.0, 200.0, 1500.0, 15.0], "Growth_Yin2": [5.0, 200.0, 1500.0, 2.0],
                "Growth_Yin3": [5.0, 200.0, 1500.0, 0.2]}[growth]
        ng = nyear
        X = np.column_stack([time, h, Tr, Ts])
    gmax = [1.0 + 0.2 * i for i in range(ng)]
    theta = np.array(base + gpar + gmax)
    ntheta = theta.size
    partype = [0] * ntheta
    nval = [1] * ntheta
    var_indx = None
    pri = [("Gaussian", [25, 5]), ("Gaussian", [0.001, 0.01]), ("Gaussian", [0.02, 0.01]), ("Gaussian", [1.0, 0.5]),
           ("Gaussian", [1.67, 0.025]), ("Uniform", [-1, 0]), ("LogNormal", [-0.7, 0.5])]
    if temporal:
        pri += [("Uniform", [0, 1])] * 3
    else:
        pri += [("Gaussian", [5, 5]), ("Gaussian", [200, 100]), ("Gaussian", [1500, 500]),
                ("Gaussian", [gpar[3], abs(gpar[3])])]
    truth_theta = list(theta)
    if var_gmax and temporal:  # gmax varies from year to year (Config_VAR_gmax.txt)
        partype[-1] = PARTYPE["VAR"]
        nval[-1] = nyear
        var_indx = np.ones((nobs, ntheta), dtype=np.int32)
        var_indx[:, -1] = 1 + np.floor(time).astype(np.int32)
        gm = [1.0 + 0.3 * i for i in range(nyear)]
        pri += [("Uniform", [0, 100])] * nyear
        truth_theta = list(theta[:-1]) + gm
    else:
        pri += [("Uniform", [0, 100])] * ng
    # observations: generated with a plain numpy forward model good to ~1e-6 (only used to make data)
    gm_t = np.asarray(truth_theta[-(nyear if (var_gmax and temporal) else ng):])
    yy = np.maximum(h - 1.0, 0.0)
    Q0 = 25.0 * np.sqrt(0.001) / 0.02 * yy**1.67
    if temporal:
        t0, ti, tm = gpar
        alf, tf = (tm - t0) / (tm - ti), 2 * tm - ti
        g0 = np.where((tt >= t0) & (tt <= tf), np.abs((tt - t0) / (tm - t0)) ** alf * (tf - tt) / (tf - tm), 0.0)
        g = g0 * (gm_t[np.floor(time).astype(int)] if (var_gmax and temporal) else gm_t[0])
    else:
        g0 = np.where((tt > 0.3) & (tt < 0.9), np.sin(np.pi * (tt - 0.3) / 0.6) ** 2, 0.0)
        g = g0 * gm_t[np.floor(time).astype(int)]
    Q = Q0 / np.sqrt(1.0 + g * yy ** (1 / 3) / (8 * 9.81 * 0.02**2) * 0.5)
    uQ = 0.03 * Q + 0.05
    Qobs = Q + rng.standard_normal(nobs) * np.sqrt(uQ**2 + (0.5 + 0.05 * Q) ** 2)
    Y = np.full((nobs, nY), -9999.0)
    Y[:, 0] = Qobs
    Y[:, 1] = g + 0.05 * rng.standard_normal(nobs)
    Yu = np.zeros((nobs, nY))
    Yu[:, 0] = uQ
    sig = ["Linear"] + ["Constant"] * (nY - 1)
    pg = [("Uniform", [0, 1000]), ("Uniform", [0, 1000])] + [("Uniform", [0, 1000])] * (nY - 1)
    prob = Problem("Vegetation", X, Y, Yu=Yu, partype=partype, nval=nval, var_indx=var_indx, priors_theta=pri,
                   sigma_func=sig, priors_gamma=pg,
                   xtra_i=[VEG_GROWTH[growth], nyear, 100], xtra_d=[5.0, 1000.0, 1e-5, 1e-10])
    truth = np.array(truth_theta + [0.5, 0.05] + [0.1] * (nY - 1))
    return prob, truth


def second_wave(model: str, nobs: int = 400, seed: int = SEED + 7, **kw):
    """Synthetic problems for the second-wave pointwise models.
    Returns (Problem, truth vector)."""
    from bam_b200.capi import ORTHO_DISTO, SWOT_ID

    rng = default_rng(seed)
    flat = lambda k: [("FlatPrior", [])] * k  # noqa: E731
    if model == "SuspendedLoad":  # Qs = t1 (h-t2)^t3 / (h-t4)^t5
        th = np.array([2.0, 0.2, 2.5, -0.5, 1.2])
        h = rng.uniform(0.1, 4.0, nobs)  # a few stages below t2: infeasible rows -> infeasible vector unless avoided
        h = np.maximum(h, 0.25)
        q = th[0] * (h - th[1]) ** th[2] / (h - th[3]) ** th[4]
        X, xtra = h.reshape(-1, 1), {}
        pri = [("LogNormal", [0.7, 1.0]), ("Gaussian", [0.2, 0.1]), ("Gaussian", [2.5, 0.5]), ("Gaussian", [-0.5, 0.3]),
               ("Gaussian", [1.2, 0.3])]
    elif model == "SWOT":
        variant = kw.get("variant", "SWOT_hSB")
        th = np.array([30.0, 5.0 if variant == "SWOT_hSB" else 80.0, 1e-5, 1.0, 1.67])
        h = rng.uniform(0.5, 6.0, nobs)
        S = rng.uniform(5e-5, 3e-4, nobs)
        B = rng.uniform(60.0, 120.0, nobs)
        Bt, B0 = (B, th[1]) if variant == "SWOT_hSB" else (th[1], 0.0)
        q = np.where(h > th[3], th[0] * (Bt - B0) * np.sqrt(S - th[2]) * np.abs(h - th[3]) ** th[4], 0.0)
        X = np.column_stack([h, S, B]) if variant == "SWOT_hSB" else np.column_stack([h, S])
        xtra = dict(xtra_i=[SWOT_ID[variant]])
        pri = [("Gaussian", [30, 10]), ("Gaussian", [th[1], 0.3 * th[1]]), ("Gaussian", [1e-5, 1e-5]), ("Gaussian", [1, 0.5]),
               ("Gaussian", [1.67, 0.1])]
    elif model == "SGD":
        th = np.array([30.0, 50.0, 1e-3, 0.5, 1.67])
        h = rng.uniform(0.8, 5.0, nobs)
        dhdt = rng.normal(0.0, 2e-5, nobs)
        q0 = th[0] * th[1] * np.sqrt(th[2]) * (h - th[3]) ** th[4]
        c = th[4] * th[0] * np.sqrt(th[2]) * (h - th[3]) ** (th[4] - 1)
        q = q0 * np.sqrt(np.maximum(1 + dhdt / (c * th[2]), 1e-3))
        X, xtra = np.column_stack([h, dhdt]), {}
        pri = [("Gaussian", [30, 10]), ("Gaussian", [50, 10]), ("LogNormal", [-6.9, 0.5]), ("Gaussian", [0.5, 0.2]),
               ("Gaussian", [1.67, 0.1])]
    elif model == "Recession_h":
        th = np.array([0.5, 2.0, 0.1, 0.2])
        t = np.sort(rng.uniform(0.0, 20.0, nobs))
        q = th[0] * np.exp(-th[1] * t) + (1 - th[0] - th[2]) * np.exp(-th[3] * t) + th[2]
        X, xtra = t.reshape(-1, 1), {}
        pri = [("Uniform", [0, 1]), ("LogNormal", [0.7, 1.0]), ("Uniform", [0, 1]), ("LogNormal", [-1.6, 1.0])]
    elif model == "Orthorectification":
        npd = int(kw.get("disto_npar", 2))
        disto = "Radial" if npd > 0 else "None"
        th = np.array([10.0, -1.0, 12.6, 3.14, -1.57, -1.0, 0.01, 960.0, 540.0, 5e-6, 1.0] + [1e-8, -1e-15][:npd])
        xyz = np.column_stack([rng.uniform(-20, 5, nobs), rng.uniform(5, 40, nobs), rng.uniform(0, 2, nobs)])
        X = xyz
        xtra = dict(xtra_i=[ORTHO_DISTO[disto], npd])
        pri = ([("Gaussian", [v, 0.1 * abs(v) + 1e-9]) for v in th[:11]] + [("Gaussian", [0, 1e-7]), ("Gaussian", [0, 1e-14])][:npd])
        q = None
    elif model == "HydraulicControl_section":
        # bathymetry table of a parabolic-ish section: h increasing, A and w increasing, A/w increasing
        hh = np.linspace(0.0, 5.0, kw.get("ntab", 60))
        ww = 4.0 + 6.0 * hh
        AA = 4.0 * hh + 3.0 * hh**2
        th = np.array([1.0, 9.81])
        H = rng.uniform(-0.2, 6.0, nobs)  # a few stages below the lowest point (Q = 0); a stage above the table is an infeasible row
        H = np.minimum(H, 4.9)
        lhs = hh + 0.5 * np.where(hh > 0, AA / np.maximum(ww, 1e-300), 0.0)
        hc = np.interp(np.clip(H, lhs[0], lhs[-1]), lhs, hh)
        q = np.where(H > 0, th[0] * np.sqrt(2 * th[1] * np.maximum(H - hc, 0)) * np.interp(hc, hh, AA), 0.0)
        X = H.reshape(-1, 1)
        xtra = dict(xtra_d=np.column_stack([hh, AA, ww]).flatten(order="F"))
        pri = [("LogNormal", [0, 0.2]), ("Gaussian", [9.81, 0.01])]
    elif model == "Segmentation":
        nS, option = int(kw.get("nS", 3)), int(kw.get("option", 1))
        t = np.sort(rng.uniform(0.0, 100.0, nobs))
        mu = [1.0, 3.0, 2.0, 4.0][:nS]
        tau = [30.0, 65.0, 85.0][: nS - 1]
        shifts = tau if option == 1 else list(np.diff([0.0] + tau))
        th = np.array(mu + shifts)
        seg = np.searchsorted(np.asarray(tau), t, side="right")
        q = np.asarray(mu)[seg]
        X = t.reshape(-1, 1)
        xtra = dict(xtra_i=[nS, int(kw.get("nmin", 3)), option], xtra_d=[float(kw.get("tmin", 0.5))])  # option 1 sets duration(1) = 1, so tmin must stay below 1 (:86,95)
        pri = [("Gaussian", [2, 5])] * nS + [("Uniform", [0, 100])] * (nS - 1)
    else:
        raise ValueError(model)
    nt = th.size
    if model == "Orthorectification":
        # observations: generated through a throw-away numpy forward model (rotation + projection, no distortion)
        az, ro, ti = th[3], th[4], th[5]
        R = np.array([[np.cos(az) * np.cos(ro) + np.sin(az) * np.cos(ti) * np.sin(ro), -np.sin(az) * np.cos(ro) + np.cos(az) * np.cos(ti) * np.sin(ro), np.sin(ti) * np.sin(ro)],
                      [-np.cos(az) * np.sin(ro) + np.sin(az) * np.cos(ti) * np.cos(ro), np.sin(az) * np.sin(ro) + np.cos(az) * np.cos(ti) * np.cos(ro), np.sin(ti) * np.cos(ro)],
                      [np.sin(az) * np.sin(ti), np.cos(az) * np.sin(ti), -np.cos(ti)]])
        d = (X - th[:3]) @ R.T
        c = th[7] - (th[6] / th[9]) * d[:, 1] / d[:, 2]
        l = th[8] - (th[6] / (th[10] * th[9])) * d[:, 0] / d[:, 2]
        Y = np.column_stack([c, l]) + rng.standard_normal((nobs, 2))
        prob = Problem(model, X, Y, partype=[0] * nt, priors_theta=pri, sigma_func=["Constant", "Constant"],
                       priors_gamma=[("Uniform", [0, 100])] * 2, **xtra)
        return prob, np.concatenate([th, [1.0, 1.0]])
    uq = 0.03 * np.abs(q) + 1e-3 * np.abs(q).mean()
    Y = q + rng.standard_normal(nobs) * np.sqrt(uq**2 + (0.01 * np.abs(q).mean() + 0.05 * np.abs(q)) ** 2)
    prob = Problem(model, X, Y, Yu=uq, partype=[0] * nt, priors_theta=pri, sigma_func=["Linear"],
                   priors_gamma=[("Uniform", [0, 1e6])] * 2, **xtra)
    return prob, np.concatenate([th, [0.01 * np.abs(q).mean(), 0.05]])


def mixture(nS: int = 3, nEvt: int = 6, nElt: int = 9, missing: bool = True, shuffle: bool = True, seed: int = SEED + 11):
    """Mixture (Mixture_model.f90): nEvt events x nElt elements, nS sources; rows optionally interleaved across events so
    that the reference's pack() pairing (e-th input row of event j -> output row (j-1)*nElt+e) is exercised.
    Returns (Problem, truth)."""
    rng = default_rng(seed)
    n = nEvt * nElt
    evt = np.repeat(np.arange(1, nEvt + 1), nElt)
    elt = np.tile(np.arange(1, nElt + 1), nEvt)
    if shuffle:
        o = rng.permutation(n)
        evt, elt = evt[o], elt[o]
    src = rng.uniform(0.3, 2.0, (n, nS))
    if missing:
        w = rng.dirichlet(np.ones(nS + 1), nEvt)[:, :nS] * 0.9  # sum < 1: the rest goes to the missing source
        miss = rng.uniform(0.5, 1.5, nElt)
        th = np.concatenate([w.ravel(), miss])
        pri = [("Uniform", [0, 1])] * (nS * nEvt) + [("FlatPrior+", [])] * nElt
    else:
        wf = rng.dirichlet(np.ones(nS), nEvt)
        w = wf[:, : nS - 1]
        th = w.ravel()
        pri = [("Uniform", [0, 1])] * ((nS - 1) * nEvt)
    # truth in the reference's own output order: row (j-1)*nElt+e = e-th input row of event j
    y = np.zeros(n)
    for j in range(nEvt):
        rows = np.flatnonzero(evt == j + 1)
        wj = w[j] if missing else np.append(w[j], 1.0 - w[j].sum())
        y[j * nElt:(j + 1) * nElt] = src[rows] @ wj
        if missing:
            y[j * nElt:(j + 1) * nElt] += (1.0 - w[j].sum()) * miss
    X = np.column_stack([evt.astype(float), elt.astype(float), src])
    Y = y + 0.02 * rng.standard_normal(n)
    prob = Problem("Mixture", X, Y, partype=[0] * th.size, priors_theta=pri, sigma_func=["Linear"],
                   priors_gamma=[("Uniform", [0, 1000])] * 2, xtra_i=[nS, nEvt, nElt, int(missing)])
    return prob, np.concatenate([th, [0.02, 0.01]])


def sfdtidal_qmec(nobs: int = 600, seed: int = SEED + 12, missing_every: int = 3):
    """SFDTidal_Qmec (time-recursive momentum balance between two tide gauges, SFDTidal_Qmec_model.f90:61-197):
    theta = (y0, A0, be, delta, phi, d1, d2, c, g, Q0, dx, dt).  Two synthetic tidal stage series at a 60 s step; the
    discharge "observations" are the model's own output plus noise, with gaps (most of a real record is ungauged).
    Returns (Problem, truth)."""
    rng = default_rng(seed)
    th = np.array([2.0, 1800.0, -6.0, 0.02, 4e-6, 0.0, 0.05, 4.0 / 3.0, 9.81, 300.0, 4000.0, 60.0])
    t = np.arange(nobs) * th[11]
    h1 = 1.5 * np.sin(2 * np.pi * t / 44700.0) + 0.3 * np.sin(2 * np.pi * t / 7000.0)
    h2 = 1.5 * np.sin(2 * np.pi * (t - 600.0) / 44700.0) + 0.3 * np.sin(2 * np.pi * (t - 500.0) / 7000.0) - 0.03
    y0, A0, be, delta, phi, d1, d2, c, g, Q0, dx, dt = th
    width = A0 / (y0 - be)
    R0 = A0 / (width + 2 * (y0 - be))
    ne2 = phi / (8 * g) * A0 * R0**c
    y1, y2 = h1 + d1, h2 + d2
    dy, ym = y2 - y1, (y1 + y2) / 2
    Ae = width * (ym - be)
    Rhe = Ae / (width + 2 * (ym - be))
    Q = np.zeros(nobs)
    Q[0] = Q0
    pgf = lambda m: -g * Ae[m] * (dy[m] + delta) / dx  # noqa: E731
    bff = lambda m: -g * (ne2 / Rhe[m] ** c) * Q[m] * abs(Q[m]) / Ae[m]  # noqa: E731
    for m in range(1, nobs):
        if m == 1:
            f = pgf(0) + bff(0) + 2 * (Q[0] / Ae[0]) * width * (ym[1] - ym[0]) / dt
        else:
            d1t = (ym[m] - ym[m - 2]) / (2 * dt)
            d2t = (ym[m - 1] - ym[m - 2]) / dt if m == 2 else (ym[m - 1] - ym[m - 3]) / (2 * dt)
            f = (1.5 * pgf(m - 1) - 0.5 * pgf(m - 2) + 1.5 * bff(m - 1) - 0.5 * bff(m - 2) +
                 1.5 * 2 * (Q[m - 1] / Ae[m - 1]) * width * d1t - 0.5 * 2 * (Q[m - 2] / Ae[m - 2]) * width * d2t)
        Q[m] = Q[m - 1] + dt * f
    uq = 0.03 * np.abs(Q) + 5.0
    Y = Q + rng.standard_normal(nobs) * np.sqrt(uq**2 + (10.0 + 0.02 * np.abs(Q)) ** 2)
    if missing_every > 1:
        Y[np.arange(nobs) % missing_every != 0] = -9999.0
    pri = [("Gaussian", [v, 0.05 * abs(v) + 0.01]) for v in th]
    partype = [-1, 0, 0, 0, 0, -1, -1, -1, -1, 0, -1, -1]  # y0, datums, exponent, gravity, dx, dt fixed by the user (:96-107)
    prob = Problem("SFDTidal_Qmec", np.column_stack([h1, h2]), Y, Yu=uq, partype=partype,
                   priors_theta=[pri[i] for i in range(12) if partype[i] == 0], fix_value=[v if partype[i] == -1 else 0.0 for i, v in enumerate(th)],
                   sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000])] * 2)
    truth = np.concatenate([[th[i] for i in range(12) if partype[i] == 0], [10.0, 0.02]])
    return prob, truth


def baratin_bac(nobs: int = 300, seed: int = SEED + 8, hmax: float = 6.0):
    """BaRatinBAC: the 3-control matrix of tests/BaM_BaRatinBAC (control 2 REPLACES control 1, control 3 is ADDED),
    theta = (b, a, c) per control.  Returns (Problem, truth)."""
    rng = default_rng(seed)
    M = [[1, 0, 0], [0, 1, 0], [0, 1, 1]]
    b, a, c = [-0.6, -0.1, 1.2], [14.17, 26.52, 31.82], [1.5, 1.67, 1.67]
    # activation stages: k1 = b1; k2 solves a1 (k-b1)^c1 = a2 (k-b2)^c2 (replacement); k3 = b3 (addition)
    from scipy.optimize import brentq
    k2 = brentq(lambda x: c[1] * np.log(x - b[1]) - c[0] * np.log(x - b[0]) - np.log(a[0] / a[1]), -0.1 + 1e-9, 1.2)
    H = rng.uniform(-1.0, 4.0, nobs)
    Q = np.zeros(nobs)
    r1 = (H >= b[0]) & (H < k2)
    r2 = (H >= k2) & (H < b[2])
    r3 = H >= b[2]
    Q[r1] = a[0] * (H[r1] - b[0]) ** c[0]
    Q[r2] = a[1] * (H[r2] - b[1]) ** c[1]
    Q[r3] = a[1] * (H[r3] - b[1]) ** c[1] + a[2] * (H[r3] - b[2]) ** c[2]
    uQ = 0.05 * Q + 0.05
    gam = (0.5, 0.05)
    Qobs = Q + rng.standard_normal(nobs) * np.sqrt(uQ**2 + (gam[0] + gam[1] * Q) ** 2)
    th = np.array([b[0], a[0], c[0], b[1], a[1], c[1], b[2], a[2], c[2]])
    pri = [("Gaussian", [v, 0.2 * abs(v) + 0.05]) for v in th]
    prob = Problem("BaRatinBAC", H, Qobs, Yu=uQ, partype=[0] * 9, priors_theta=pri, sigma_func=["Linear"],
                   priors_gamma=[("Uniform", [0, 1000])] * 2, xtra_i=np.asarray(M, dtype=np.int32).flatten(order="F"),
                   xtra_d=[hmax])
    return prob, np.concatenate([th, gam])


def sfd(nobs: int = 300, seed: int = SEED + 9):
    """SFD (twin-gauge stage-fall-discharge, SFD_Val_General) with the parameter values of tests/BaM_SFD."""
    rng = default_rng(seed)
    th = np.array([6000.0, -5.0, 1.6667, 3900.0, -5.0, 171.0, -0.01, 1.6667])
    h2 = rng.uniform(1.0, 3.0, nobs)
    h1 = h2 + rng.uniform(0.02, 1.5, nobs)
    # throw-away forward model for the observations: the smaller of the two regimes is a fair proxy
    qv = th[0] * (h1 - th[1]) ** th[2] * np.sqrt((h1 - h2 - th[6]) / th[3])
    qc = th[5] * (h1 - th[4]) ** th[7]
    q = np.minimum(qv, qc)
    uq = 0.05 * q
    Y = q + rng.standard_normal(nobs) * np.sqrt(uq**2 + (10.0 + 0.05 * q) ** 2)
    pri = [("Gaussian", [6000, 901.5]), ("Gaussian", [-5, 1]), ("Gaussian", [1.6667, 0.025]), ("Gaussian", [3900, 50]),
           ("Gaussian", [-5, 1]), ("Gaussian", [171, 46]), ("Gaussian", [0, 0.025]), ("Gaussian", [1.6667, 0.025])]
    prob = Problem("SFD", np.column_stack([h1, h2]), Y, Yu=uq, partype=[0] * 8, priors_theta=pri, sigma_func=["Linear"],
                   priors_gamma=[("Uniform", [0, 10000])] * 2, xtra_i=[1, 100], xtra_d=[20.0, 5.0, 1000.0, 1e-5, 1e-10])
    return prob, np.concatenate([th, [10.0, 0.05]])


def algae_biomass(nobs: int = 730, seed: int = SEED + 14, missing_every: int = 4):
    """AlgaeBiomass (time-recursive daily biomass budget, AlgaeBiomass_model.f90:55-197): 5 input series (temperature,
    irradiance, depth, turbidity, removal), 18 parameters, outputs biomass and biomass / bmax.  Synthetic two-year record
    with a few floods (hydraulic scour) and cuts (removal); the "observations" are the model's own output plus noise,
    with gaps.  Returns (Problem, truth)."""
    rng = default_rng(seed)
    day = np.arange(nobs, dtype=np.float64)
    temp = 14.0 - 10.0 * np.cos(2 * np.pi * (day - 20.0) / 365.0) + 1.5 * rng.standard_normal(nobs)
    irr = np.maximum(140.0 - 110.0 * np.cos(2 * np.pi * (day - 10.0) / 365.0) + 25.0 * rng.standard_normal(nobs), 1.0)
    depth = 0.6 + 0.15 * np.sin(2 * np.pi * day / 90.0) ** 2
    flood = (rng.uniform(0.0, 1.0, nobs) < 0.02)
    depth = depth + np.convolve(flood.astype(float), np.array([1.2, 0.8, 0.4, 0.2]), mode="full")[:nobs]
    turb = np.maximum(5.0 + 40.0 * (depth - 0.6), 0.0)
    rem = np.where(rng.uniform(0.0, 1.0, nobs) < 0.01, 0.4, 0.0)
    th = np.array([1.0, 0.3, 0.001, 1.0, 0.12, 6.0, 16.5, 30.0, 150.0, 0.02, 0.5, 0.02, 9.81, 1000.0, 1e-4, 0.8, 1.5, 0.05])
    dt, b0, bmin, bmax, mu0, Tmin, Topt, Tmax, Iopt, katt, kcw, Rsen, g, rho, J, tau0, bHyd, cHyd = th
    Texp = (Topt - Tmin) / (Tmax - Topt)
    b = np.zeros(nobs)
    bt = b0
    for t in range(nobs):
        Fb = 1.0 - bt / bmax
        Ft = 0.0 if (temp[t] <= Tmin or temp[t] >= Tmax) else ((Tmax - temp[t]) / (Tmax - Topt)) * ((temp[t] - Tmin) / (Topt - Tmin)) ** Texp
        Ir = irr[t] * np.exp(-(katt * turb[t] + kcw) * depth[t]) / Iopt
        Fi = Ir * np.exp(1.0 - Ir)
        tau = depth[t] * g * J * rho
        Rh = 0.0 if tau <= tau0 else cHyd * ((tau - tau0) / tau0) ** bHyd
        bt = min(max(bt + mu0 * Fb * Ft * Fi * bt * dt - (Rsen + rem[t] + Rh) * (bt - bmin) * dt, bmin), bmax)
        b[t] = bt
    Y = np.column_stack([b, b / bmax]) + 0.02 * rng.standard_normal((nobs, 2))
    if missing_every > 1:
        Y[np.arange(nobs) % missing_every != 0, :] = -9999.0
    Y[:, 0] = np.where(np.arange(nobs) % 2 == 0, -9999.0, Y[:, 0])  # the second output is the better observed one
    partype = [-1, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, -1, -1, 0, 0, 0]  # dt, bmin, g, rho, J fixed (Config_Model_FIX.txt)
    pri = [("Gaussian", [v, 0.05 * abs(v) + 1e-3]) for v in th]
    prob = Problem("AlgaeBiomass", np.column_stack([temp, irr, depth, turb, rem]), Y, Yu=np.full((nobs, 2), 0.01), partype=partype,
                   priors_theta=[pri[i] for i in range(18) if partype[i] == 0],
                   fix_value=[v if partype[i] == -1 else 0.0 for i, v in enumerate(th)], sigma_func=["Constant", "Constant"],
                   priors_gamma=[("Uniform", [0, 10]), ("Uniform", [0, 10])])
    truth = np.concatenate([[th[i] for i in range(18) if partype[i] == 0], [0.02, 0.02]])
    return prob, truth


def dynamic_vegetation(nobs: int = 730, seed: int = SEED + 15, missing_every: int = 3):
    """DynamicVegetation (DynamicVegetation_model.f90:74-191): the AlgaeBiomass budget on depth = max(stage - b1, 0) drives
    the vegetation-affected rating curve.  5 input series (temperature, irradiance, STAGE, turbidity, removal), 18 + 6
    parameters, outputs Q, biomass, biomass / bmax.  The "observations" come from the oracle's own run of the model (the
    generator needs the Newton solve) plus noise; a dry spell puts some stages below b1.  Returns (Problem, truth)."""
    from oracle.oracle_api import Oracle

    rng = default_rng(seed)
    day = np.arange(nobs, dtype=np.float64)
    temp = 14.0 - 10.0 * np.cos(2 * np.pi * (day - 20.0) / 365.0) + 1.5 * rng.standard_normal(nobs)
    irr = np.maximum(140.0 - 110.0 * np.cos(2 * np.pi * (day - 10.0) / 365.0) + 25.0 * rng.standard_normal(nobs), 1.0)
    stage = 0.9 + 0.15 * np.sin(2 * np.pi * day / 90.0) ** 2
    flood = (rng.uniform(0.0, 1.0, nobs) < 0.02)
    stage = stage + np.convolve(flood.astype(float), np.array([1.2, 0.8, 0.4, 0.2]), mode="full")[:nobs]
    stage[200:206] = 0.25  # dry spell: stage below b1 -> depth 0, Q = 0
    turb = np.maximum(5.0 + 40.0 * (stage - 0.9), 0.0)
    rem = np.where(rng.uniform(0.0, 1.0, nobs) < 0.01, 0.4, 0.0)
    alg = [1.0, 0.3, 0.001, 1.0, 0.12, 6.0, 16.5, 30.0, 150.0, 0.02, 0.5, 0.02, 9.81, 1000.0, 1e-4, 0.8, 1.5, 0.05]
    th = np.array(alg + [12.0, 0.03, 0.3, 1.67, -0.8, 0.5])  # B, n1, b1, c1, chi, U_chi
    partype = [-1, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, -1, -1, 0, 0, 0] + [0] * 6
    pri = [("Gaussian", [v, 0.05 * abs(v) + 1e-3]) for v in th]
    X = np.column_stack([temp, irr, stage, turb, rem])

    def make(Y):
        return Problem("DynamicVegetation", X, Y, Yu=np.column_stack([np.full(nobs, 0.002), np.full((nobs, 2), 0.01)]),
                       partype=partype, priors_theta=[pri[i] for i in range(24) if partype[i] == 0],
                       fix_value=[v if partype[i] == -1 else 0.0 for i, v in enumerate(th)],
                       sigma_func=["Linear", "Constant", "Constant"],
                       priors_gamma=[("Uniform", [0, 10])] * 4, xtra_d=[1.0, 1.0, 1e-10, 1e-12], xtra_i=[100])

    truth = np.concatenate([[th[i] for i in range(24) if partype[i] == 0], [0.001, 0.03, 0.02, 0.02]])
    _, Ys, feas = Oracle(make(np.zeros((nobs, 3)))).calibration_run(truth)
    assert feas
    Y = Ys * (1.0 + 0.03 * rng.standard_normal((nobs, 3))) + 0.002 * rng.standard_normal((nobs, 3))
    if missing_every > 1:
        Y[np.arange(nobs) % missing_every != 0, 1:] = -9999.0
    return make(Y), truth


def sfdtidal_variant(model: str, nobs: int = 600, seed: int = SEED + 16, missing_every: int = 3):
    """SFDTidal_Qmec0 (original Qmec parameterisation: Be, he, dzeta, ne, d1, d2, c, g, Q0, dx, dt; stages minus datums) and
    SFDTidal_QmecMS (steady Manning-Strickler version: y0, A0, be, delta, phi, d1, d2, c, g, dx) on the two synthetic tidal
    stage series of sfdtidal_qmec.  The discharge "observations" are the oracle's own run plus noise.  Returns (Problem, truth)."""
    from oracle.oracle_api import Oracle

    rng = default_rng(seed)
    t = np.arange(nobs) * 60.0
    h1 = 1.5 * np.sin(2 * np.pi * t / 44700.0) + 0.3 * np.sin(2 * np.pi * t / 7000.0)
    h2 = 1.5 * np.sin(2 * np.pi * (t - 600.0) / 44700.0) + 0.3 * np.sin(2 * np.pi * (t - 500.0) / 7000.0) - 0.03
    if model == "SFDTidal_Qmec0":
        th = np.array([225.0, 8.0, -0.02, 0.03, 0.0, -0.05, 4.0 / 3.0, 9.81, 300.0, 4000.0, 60.0])
        partype = [0, 0, 0, 0, -1, -1, -1, -1, 0, -1, -1]
    elif model == "SFDTidal_QmecMS":
        th = np.array([2.0, 1800.0, -6.0, 0.02, 4e-6, 0.0, 0.05, 4.0 / 3.0, 9.81, 4000.0])
        partype = [-1, 0, 0, 0, 0, -1, -1, -1, -1, -1]
    else:
        raise ValueError(model)
    nt = th.size
    pri = [("Gaussian", [v, 0.05 * abs(v) + 0.01]) for v in th]

    def make(Y, uq):
        return Problem(model, np.column_stack([h1, h2]), Y, Yu=uq, partype=partype,
                       priors_theta=[pri[i] for i in range(nt) if partype[i] == 0],
                       fix_value=[v if partype[i] == -1 else 0.0 for i, v in enumerate(th)],
                       sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000])] * 2)

    truth = np.concatenate([[th[i] for i in range(nt) if partype[i] == 0], [10.0, 0.02]])
    _, Ys, feas = Oracle(make(np.zeros(nobs), np.ones(nobs))).calibration_run(truth)
    assert feas
    Q = Ys[:, 0]
    uq = 0.03 * np.abs(Q) + 5.0
    Y = Q + rng.standard_normal(nobs) * np.sqrt(uq**2 + (10.0 + 0.02 * np.abs(Q)) ** 2)
    if missing_every > 1:
        Y[np.arange(nobs) % missing_every != 0] = -9999.0
    return make(Y, uq), truth


def sfdtidal_lag(model: str, nobs: int = 600, seed: int = SEED + 17, missing_every: int = 3):
    """SFDTidal / SFDTidal2 / SFDTidal4 / SFDTidalJones: stage-fall-discharge curves for tidal rivers that read the stages
    at a lagged time step (8 parameters each; the lag and SFDTidal4's averaging half-width are integers carried in real
    parameters, fixed by the user).  Two synthetic tidal stage series (+ dh/dt for Jones); the discharge "observations"
    are the oracle's own run plus noise.  Returns (Problem, truth)."""
    from oracle.oracle_api import Oracle

    rng = default_rng(seed)
    t = np.arange(nobs) * 60.0
    h1 = 2.0 + 1.5 * np.sin(2 * np.pi * t / 44700.0) + 0.3 * np.sin(2 * np.pi * t / 7000.0)
    h2 = 2.0 + 1.5 * np.sin(2 * np.pi * (t - 600.0) / 44700.0) + 0.3 * np.sin(2 * np.pi * (t - 500.0) / 7000.0) - 0.03
    X = [h1, h2]
    if model == "SFDTidal":        # Dt, K, B1, B2, h0, M, L, deltaz
        th = np.array([7.0, 30.0, 200.0, 240.0, -5.0, 1.67, 4000.0, 0.02])
        partype = [-1, 0, 0, 0, 0, 0, -1, 0]
    elif model == "SFDTidal2":     # Dt, K, B, h0, M, L, delta, alpha
        th = np.array([7.0, 30.0, 220.0, -5.0, 1.67, 4000.0, 0.05, 0.97])
        partype = [-1, 0, 0, 0, 0, -1, 0, 0]
    elif model == "SFDTidal4":     # Dt, K, B, h0, M, L, delta, P
        th = np.array([7.0, 30.0, 220.0, -5.0, 1.67, 4000.0, 0.02, 5.0])
        partype = [-1, 0, 0, 0, 0, -1, 0, -1]
    elif model == "SFDTidalJones":  # Dt, K, B, h0, M, L, delta, alpha; third input dh/dt
        th = np.array([7.0, 30.0, 220.0, -5.0, 1.67, 4000.0, 0.02, 800.0])
        partype = [-1, 0, 0, 0, 0, -1, 0, 0]
        X.append(np.gradient(h1, 60.0))
    else:
        raise ValueError(model)
    pri = [("Gaussian", [v, 0.05 * abs(v) + 0.01]) for v in th]

    def make(Y, uq):
        return Problem(model, np.column_stack(X), Y, Yu=uq, partype=partype,
                       priors_theta=[pri[i] for i in range(8) if partype[i] == 0],
                       fix_value=[v if partype[i] == -1 else 0.0 for i, v in enumerate(th)],
                       sigma_func=["Linear"], priors_gamma=[("Uniform", [0, 1000])] * 2)

    truth = np.concatenate([[th[i] for i in range(8) if partype[i] == 0], [10.0, 0.02]])
    _, Ys, feas = Oracle(make(np.zeros(nobs), np.ones(nobs))).calibration_run(truth)
    assert feas
    Q = Ys[:, 0]
    uq = 0.03 * np.abs(Q) + 5.0
    Y = Q + rng.standard_normal(nobs) * np.sqrt(uq**2 + (10.0 + 0.02 * np.abs(Q)) ** 2)
    if missing_every > 1:
        Y[np.arange(nobs) % missing_every != 0] = -9999.0
    return make(Y, uq), truth
