#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference's own example workspaces.

Run HERE (the container that has /root/reference); the JSON files are committed and travel
to the GPU box, /root/reference does not.  Each fixture is the *problem description* the
reference would hold in its INFER/MODEL globals after LoadBamObjects
(src/BaM_tools.f90:133-420) for one workspace of /root/reference/tests, plus the configured
starting point (theta0, gamma0, Xobs for latent inputs, zero biases: Packer, :4202-4242) and
the known-answer columns some data files carry.

This reader is deliberately independent of the C++ host's config reader
(bam_b200/csrc/host) so that the two cross-check each other in tests/test_host_driver.py.

Matrices are stored flat in COLUMN-MAJOR order.
"""
import json
import os
import re
import sys

import numpy as np

REF = os.environ.get("BAM_REFERENCE", "/root/reference")
TESTS = os.path.join(REF, "tests")
OUT = os.path.dirname(os.path.abspath(__file__))

SED_FORMULA = {"MPM": 1, "Camenen": 2, "Nielsen": 3, "Smart": 4}
SED_CSS = {"Constant": 1, "Soulsby": 2, "SoulsbyWhitehouse": 3}
SED_PPO = {"FIX": 1, "IN": 2, "PAR": 3}
SED_CO = {"FIX": 1, "PAR": 2}
DIST_NPAR = {"Gaussian": 2, "Uniform": 2, "Triangle": 3, "LogNormal": 2, "Exponential": 2, "FlatPrior": 0,
             "FlatPrior+": 0, "FlatPrior-": 0}
SIG_NPAR = {"Constant": 1, "Linear": 2, "Proportional": 1, "Power": 3, "Exponential": 3, "Gaussian": 3}


def items(line):
    """Fortran list-directed record -> list of strings (quotes stripped, '!' comment dropped)."""
    out, i, n = [], 0, len(line)
    while i < n:
        ch = line[i]
        if ch in " \t,\r\n":
            i += 1
        elif ch == "!":
            break
        elif ch in "\"'":
            j = line.index(ch, i + 1)
            out.append(line[i + 1:j])
            i = j + 1
        else:
            j = i
            while j < n and line[j] not in " \t,\r\n!":
                j += 1
            out.append(line[i:j])
            i = j
    return out


def fnum(s):
    return float(s.replace("d", "e").replace("D", "e"))


def path(ws, name):
    name = name.replace("\\", "/")
    for cand in (os.path.join(TESTS, name), os.path.join(ws, name), os.path.join(ws, os.path.basename(name))):
        if os.path.exists(cand):
            return cand
    raise FileNotFoundError(name)


def read_lines(fn):
    with open(fn, errors="replace") as f:
        return f.read().splitlines()


def read_par_block(lines, k):
    """Config_Read_Par (src/BaM_tools.f90:3751-3821): name, init, dist|FIX|VAR, pars|file."""
    name = items(lines[k])[0]
    init = fnum(items(lines[k + 1])[0])
    dist = items(lines[k + 2])[0]
    rest = items(lines[k + 3]) if k + 3 < len(lines) else []
    return name, init, dist, rest


def load_workspace(cfg_name, max_rows=None):
    cfg = read_lines(os.path.join(TESTS, cfg_name))
    ws = os.path.join(TESTS, items(cfg[0])[0].replace("\\", "/"))
    f_model, f_xtra, f_data = (items(cfg[i])[0] for i in (2, 3, 4))
    f_remnant = items(cfg[5])
    f_mcmc = items(cfg[6])[0]

    # ---- Config_Data (:2204-2270) + BaM_ReadData
    d = read_lines(path(ws, f_data))
    datafile = path(ws, items(d[0])[0])
    nhead, nobs, ncol = (int(items(d[i])[0]) for i in (1, 2, 3))
    cols = [[int(v) for v in items(d[4 + i])] for i in range(8)]
    raw = read_lines(datafile)
    header = raw[:nhead]
    W = np.array([[fnum(v) for v in re.split(r"[\s,;]+", ln.strip())[:ncol]] for ln in raw[nhead:nhead + nobs]])
    assert W.shape == (nobs, ncol), (W.shape, nobs, ncol)
    if max_rows is not None and nobs > max_rows:  # a prefix of a long time series keeps the fixture small
        W, nobs = W[:max_rows], max_rows

    def extract(c):  # Wextract (:4181-4198)
        return np.column_stack([np.zeros(nobs) if j == 0 else W[:, j - 1] for j in c])

    X, Xu, Xb, Xbi, Y, Yu, Yb, Ybi = (extract(c) for c in cols)
    nX, nY = X.shape[1], Y.shape[1]

    # ---- Config_Model (:2096-2160)
    m = read_lines(path(ws, f_model))
    model_id = items(m[0])[0]
    assert int(items(m[1])[0]) == nX and int(items(m[2])[0]) == nY
    ntheta = int(items(m[3])[0])
    partype, nval, fix_value, names = [], [], [], []
    priors_theta, theta0, var_cols = [], [], []
    for i in range(ntheta):
        name, init, dist, rest = read_par_block(m, 4 + 4 * i)
        names.append(name)
        if dist == "FIX":
            partype.append(-1); nval.append(1); fix_value.append(init); var_cols.append(0)
        elif dist == "VAR":  # Config_Finalize (:2833-2944)
            v = read_lines(path(ws, rest[0]))
            nv, col = int(items(v[0])[0]), int(items(v[1])[0])
            partype.append(1); nval.append(nv); fix_value.append(0.0); var_cols.append(col)
            for j in range(nv):
                _, init_j, dist_j, rest_j = read_par_block(v, 2 + 4 * j)
                priors_theta.append([dist_j, [fnum(x) for x in rest_j[:DIST_NPAR[dist_j]]]])
                theta0.append(init_j)
        else:
            partype.append(0); nval.append(1); fix_value.append(0.0); var_cols.append(0)
            priors_theta.append([dist, [fnum(x) for x in rest[:DIST_NPAR[dist]]]])
            theta0.append(init)
    var_indx = None
    if any(t == 1 for t in partype):
        var_indx = np.column_stack([W[:, c - 1].astype(int) if c > 0 else np.ones(nobs, dtype=int) for c in var_cols])

    # ---- Config_RemnantSigma (:3880-3946), one per output
    sigma_func, priors_gamma, gamma0 = [], [], []
    for f in f_remnant:
        r = read_lines(path(ws, f))
        funk = items(r[0])[0]
        npar = int(items(r[1])[0])
        assert npar == SIG_NPAR[funk]
        sigma_func.append(funk)
        for j in range(npar):
            _, init_j, dist_j, rest_j = read_par_block(r, 2 + 4 * j)
            priors_gamma.append([dist_j, [fnum(x) for x in rest_j[:DIST_NPAR[dist_j]]]])
            gamma0.append(init_j)

    # ---- xtra (ModelLibrary_tools.f90:436-544)
    xtra_i, xtra_d, xtra_text = [], [], None
    if model_id == "BaRatin":
        rows = [[int(round(fnum(v))) for v in items(ln)] for ln in read_lines(path(ws, f_xtra)) if items(ln)]
        nc = len(rows)
        M = np.array([r[:nc] for r in rows])
        xtra_i = [int(v) for v in M.flatten(order="F")]
    elif model_id == "BaRatinBAC":  # BaRatinBAC_XtraRead (BaRatinBAC_model.f90:150-180): control matrix rows, then hmax
        rows = [[fnum(v) for v in items(ln)] for ln in read_lines(path(ws, f_xtra)) if items(ln)]
        nc = len(rows) - 1
        M = np.array([[int(round(v)) for v in r[:nc]] for r in rows[:nc]])
        xtra_i = [int(v) for v in M.flatten(order="F")]
        xtra_d = [rows[nc][0]]
    elif model_id == "Sediment":
        s = read_lines(path(ws, f_xtra))
        ppo = items(s[3])[:4]
        xtra_i = [SED_FORMULA[items(s[0])[0]], SED_CSS[items(s[1])[0]], SED_CO[items(s[2])[0]]] + [SED_PPO[p] for p in ppo]
        xtra_d = [fnum(v) for v in items(s[4])[:4]]
    elif model_id == "TextFile":
        with open(path(ws, f_xtra)) as f:
            xtra_text = f.read()
    elif model_id == "SFD":  # SFD_XtraRead (StageFallDischarge_model.f90:191-247): ID, 5 Newton options, maxiter
        s = read_lines(path(ws, f_xtra))
        assert items(s[0])[0] == "SFD_Val_General"
        xtra_d = [fnum(items(s[i])[0]) for i in range(1, 6)]
        xtra_i = [1, int(items(s[6])[0])]
    elif model_id == "Vegetation":  # Vegetation_XtraRead (Vegetation_model.f90:244-289): growth formula, years, Newton options
        s = read_lines(path(ws, f_xtra))
        growth = {"Growth_Yin1": 1, "Growth_Yin2": 2, "Growth_Yin3": 3, "Growth_Yin1_temporal": 4}[items(s[0])[0]]
        xtra_i = [growth, int(items(s[1])[0]), int(items(s[6])[0])]
        xtra_d = [fnum(items(s[i])[0]) for i in range(2, 6)]
    elif model_id == "Orthorectification":  # Ortho_XtraRead (Orthorectification_model.f90:192-230): distortion ID, npar
        s = read_lines(path(ws, f_xtra))
        xtra_i = [{"None": 1, "Radial": 2}[items(s[0])[0]], int(items(s[1])[0])]
    elif model_id == "HydraulicControl_section":  # h-A-w table, one row per line, stored column-major (:114-150)
        rows = [[fnum(v) for v in items(ln)[:3]] for ln in read_lines(path(ws, f_xtra)) if items(ln)]
        xtra_d = [r[c] for c in range(3) for r in rows]
    elif model_id == "Segmentation":  # Segmentation_XtraRead (Segmentation_model.f90:142-180): nS, tmin, nmin, option
        s = read_lines(path(ws, f_xtra))
        xtra_i = [int(items(s[0])[0]), int(items(s[2])[0]), int(items(s[3])[0])]
        xtra_d = [fnum(items(s[1])[0])]
    elif model_id == "Mixture":  # Mixture_XtraRead (Mixture_model.f90:137-190): nS, nEvt, nElt, missingSource
        s = read_lines(path(ws, f_xtra))
        xtra_i = [int(items(s[i])[0]) for i in range(3)] + [1 if items(s[3])[0].strip(".").lower().startswith("t") else 0]

    # ---- optional prior correlation (src/BaM_main.f90:22,302; src/BaM_tools.f90:386-414)
    priorC = None
    pc = os.path.join(ws, "PriorCorrelation.txt")
    if os.path.exists(pc):
        priorC = np.array([[fnum(v) for v in ln.split()] for ln in read_lines(pc) if ln.strip()])

    # ---- Config_MCMC (:2391-2489)
    mc = read_lines(path(ws, f_mcmc))
    mcmc = dict(nAdapt=int(items(mc[1])[0]), nCycles=int(items(mc[2])[0]), MinMoveRate=fnum(items(mc[3])[0]),
                MaxMoveRate=fnum(items(mc[4])[0]), DownMult=fnum(items(mc[5])[0]), UpMult=fnum(items(mc[6])[0]),
                InitStdMode=int(items(mc[7])[0]), MultFactor=fnum(items(mc[9])[0]))

    # ---- starting vector: Packer(theta0, gamma0, X, bias_init=0) (:605, :4202-4242)
    start = list(theta0) + list(gamma0)
    for j in range(nX):
        start += [float(X[i, j]) for i in range(nobs) if Xu[i, j] > 0]
    for j in range(nX):
        start += [0.0] * int(Xbi[:, j].max() if nobs else 0)
    for j in range(nY):
        start += [0.0] * int(Ybi[:, j].max() if nobs else 0)

    col = lambda A: [float(v) for v in A.flatten(order="F")]  # noqa: E731
    coli = lambda A: [int(v) for v in A.flatten(order="F")]  # noqa: E731
    fx = dict(
        source=f"tests/{cfg_name}", model_id=model_id, nobs=nobs, nX=nX, nY=nY, parnames=names,
        X=col(X), Xu=col(Xu), Xb=col(Xb), Xbindx=coli(Xbi), Y=col(Y), Yu=col(Yu), Yb=col(Yb), Ybindx=coli(Ybi),
        partype=partype, nval=nval, fix_value=fix_value, var_indx=None if var_indx is None else coli(var_indx),
        priors_theta=priors_theta, sigma_func=sigma_func, priors_gamma=priors_gamma,
        priorC=None if priorC is None else col(priorC),
        xtra_i=xtra_i, xtra_d=xtra_d, xtra_text=xtra_text, mv=-9999.0, unfeas_flag=-666.666,
        start=start, mcmc=mcmc, data_header=header, data_table=col(W), data_ncol=ncol,
    )
    return fx


WORKSPACES = {
    "baratin": "Config_BaM_BaRatin.txt",
    "baratin_spd": "Config_BaM_BaRatin_SPD.txt",
    "regression_x2y2": "Config_BaM_Regression_X2Y2.txt",
    "sediment_camenen_x4": None,  # built below: no top-level Config_BaM for it
    "sediment_mpm_x2": "Config_BaM_Sediment.txt",
    "textfile_inputuncertainty": "Config_BaM_TextFile_InputUncertainty.txt",
    "mixture": "Config_BaM_Mixture.txt",
    "sfd": "Config_BaM_SFD.txt",
    "sfdtidal_qmec2": "Config_BaM_SFDTidal_Qmec2.txt",  # first 8000 of its 14401 time steps: the gauged window starts at step 7745 (MAX_ROWS)
    "sfdtidal2": "Config_BaM_SFDTidal2.txt",            # Saigon, 1372 steps: the lagged stage-fall-discharge curves whose
    "sfdtidal4": "Config_BaM_SFDTidal4.txt",            # shipped workspaces match the code (8 parameters); Seine, 2881 steps
    "sfdtidal_jones": "Config_BaM_SFDTidal_Jones.txt",  # third input dh/dt
    # the remaining shipped workspaces of models on the device (the synthetic generators cover the models; these pin the readers
    # and the xtra handling on the reference's own files)
    "hydraulic_control_section": "Config_BaM_HydraulicControl_section.txt",
    "orthorectification": "Config_BaM_Orthorectification.txt",
    "regression_x2y1": "Config_BaM_Regression_X2Y1.txt",
    "sediment": "Config_BaM_Sediment.txt",
    "textfilemodel": "Config_BaM_TextFileModel.txt",
    # workspaces without a top-level Config_BaM_*.txt: (directory, xtra file, remnant files) -> an equivalent config in /tmp
    "baratin_qbias": ("BaM_BaRatin_Qbias", "Config_ControlMatrix.txt", ("Config_RemnantSigma.txt",)),            # Y biases
    "baratin_qbiashbias": ("BaM_BaRatin_QbiasHbias", "Config_ControlMatrix.txt", ("Config_RemnantSigma.txt",)),  # X errors, X and Y biases
    "baratinbac": ("BaM_BaRatinBAC", "Config_ControlMatrix.txt", ("Config_RemnantSigma.txt",)),
    "baratinbac_spd": ("BaM_BaRatinBAC_SPD", "Config_ControlMatrix.txt", ("Config_RemnantSigma.txt",)),          # VAR offsets
    "suspendedload": ("BaM_SuspendedLoad", "", ("Config_RemnantSigma.txt",)),
    "sediment_camenen_x2": ("BaM_Sediment_Camenen_X2", "Config_Options.txt", ("Config_RemnantSigma.txt",)),
    "regression_catchments": ("BaM_Regression_Catchments", "", ("Config_RemnantSigma.txt",)),
    "regression_x6": ("BaM_Emeline", "", ("Config_RemnantSigma.txt",)),
    # (BaM_Vegetation_static lists a parameter set of an older version of Growth_Yin1_temporal: the reference refuses it)
    "segmentation": "BaM_Segmentation/Config_BaM.txt",
    # (Config_BaM_Vegetation.txt -- the Ladhof workspace -- lists 18 parameters where Growth_Yin2 with 8 years takes 19: the
    #  reference itself refuses it; Vegetation is pinned by tests/BaM_Vegetation/SyntheticData.txt, make_vegetation_fixture.py)
}
MAX_ROWS = {"sfdtidal_qmec2": 8000}


def synth_cfg(ws_name, remnant=("Config_RemnantSigma.txt",), xtra="Config_Options.txt"):
    """Some workspaces have no top-level Config_BaM_*.txt: write an equivalent one to /tmp."""
    lines = [f'"{ws_name}/"', '"Config_RunOptions.txt"', '"Config_Model.txt"', f'"{xtra}"', '"Config_Data.txt"',
             ",".join(f'"{r}"' for r in remnant), '"Config_MCMC.txt"', '"Config_Cooking.txt"', '"Config_Summary.txt"',
             '"Config_Residuals.txt"', '"Config_Pred_Master.txt"']
    return lines


def main():
    global TESTS
    made = []
    for key, cfg in WORKSPACES.items():
        if cfg is None or isinstance(cfg, tuple):
            tmp = f"/tmp/_cfg_{key}.txt"
            with open(tmp, "w") as f:
                lines = synth_cfg("BaM_Sediment_Camenen_X4") if cfg is None else synth_cfg(cfg[0], remnant=cfg[2], xtra=cfg[1])
                f.write("\n".join(lines) + "\n")
            # load_workspace joins TESTS with cfg_name: give it a relative path out of TESTS
            cfg = os.path.relpath(tmp, TESTS)
        fx = load_workspace(cfg, MAX_ROWS.get(key))
        if key.startswith("sediment"):
            fx["source"] = "tests/BaM_Sediment_Camenen_X4" if key.endswith("x4") else fx["source"]
        with open(os.path.join(OUT, key + ".json"), "w") as f:
            json.dump(fx, f)
        made.append((key, fx["model_id"], fx["nobs"], len(fx["start"])))
    # BaRatin prediction inputs of the reference workspace (Config_Pred_*.txt): stage grid + noisy limnigraph
    ws = os.path.join(TESTS, "BaM_BaRatin")
    hgrid = np.loadtxt(os.path.join(ws, "Hgrid.txt"))
    limni = np.loadtxt(os.path.join(ws, "Limni.txt"))
    noisy = np.loadtxt(os.path.join(ws, "Limni_noisy.txt"))
    with open(os.path.join(OUT, "baratin_pred_inputs.json"), "w") as f:
        json.dump(dict(Hgrid=hgrid.tolist(), Limni=limni.tolist(), Limni_noisy_colmajor=noisy.flatten(order="F").tolist(),
                       Limni_noisy_shape=list(noisy.shape)), f)
    for m in made:
        print("fixture", *m)


if __name__ == "__main__":
    sys.exit(main())
