"""The C++ host driver `bam_gpu` (bam_b200/csrc/host): config reader cross-checked against the
independent Python reader that produced the golden fixtures, and (GPU) a full run that writes the
reference's result files."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import load_fixture
from tests.workspace_writer import write_workspace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bam_b200", "bam_gpu")
FIXTURES = ["baratin", "baratin_spd", "regression_x2y2", "sediment_camenen_x4", "sediment_mpm_x2",
            "textfile_inputuncertainty", "mixture", "sfd"]
DIST = {"Gaussian": 1, "Uniform": 2, "Triangle": 3, "LogNormal": 4, "Exponential": 5, "FlatPrior": 6, "FlatPrior+": 7,
        "FlatPrior-": 8}
SIG = {"Constant": 1, "Linear": 2, "Proportional": 3, "Power": 4, "Exponential": 5, "Gaussian": 6}


@pytest.fixture(scope="module", autouse=True)
def built():
    from bam_b200 import build

    build.build_all()
    assert os.path.exists(EXE)


def dump(cfg):
    out = cfg + ".json"
    r = subprocess.run([EXE, "-cf", cfg, "--dump-problem", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.load(open(out))


def compare(d, fx):
    n = fx["nobs"]
    assert (d["model_id"], d["nobs"], d["nX"], d["nY"]) == (fx["model_id"], n, fx["nX"], fx["nY"])
    for key in ("X", "Xu", "Xb", "Y", "Yu", "Yb"):
        assert np.array_equal(np.asarray(d[key]), np.asarray(fx[key])), key
    for key in ("Xbindx", "Ybindx", "partype", "nval"):
        assert list(d[key]) == list(fx[key]), key
    if fx.get("var_indx") is not None:
        assert list(d["var_indx"]) == list(fx["var_indx"])
    assert list(d["prior_theta_dist"]) == [DIST[p[0]] for p in fx["priors_theta"]]
    assert list(d["prior_gamma_dist"]) == [DIST[p[0]] for p in fx["priors_gamma"]]
    par = np.asarray(d["prior_theta_par"]).reshape(-1, 4)
    for row, (_, p) in zip(par, fx["priors_theta"]):
        assert list(row[: len(p)]) == list(p)
    assert list(d["sigma_func"]) == [SIG[s] for s in fx["sigma_func"]]
    assert list(d["xtra_i"]) == list(fx["xtra_i"]) and np.allclose(d["xtra_d"], fx["xtra_d"], rtol=0, atol=0)
    assert np.array_equal(np.asarray(d["start"]), np.asarray(fx["start"]))
    if fx.get("priorC") is not None:
        k = int(round(len(fx["priorC"]) ** 0.5))
        C = np.asarray(fx["priorC"]).reshape(k, k).T
        M = np.asarray(d["priorM"]).reshape(k, k).T
        assert np.allclose(M, np.linalg.inv(C) - np.eye(k), rtol=1e-9, atol=1e-9)
    # jump std: stdFactor*|x0| (0.1 if zero), latent X -> Xu, biases -> 0.1 (src/BaM_tools.f90:605-606,2459-2471)
    f = fx["mcmc"]["MultFactor"]
    nfg = len(fx["priors_theta"]) + len(fx["priors_gamma"])
    exp = [f * abs(v) if f * abs(v) != 0 else 0.1 for v in fx["start"][:nfg]]
    assert np.allclose(d["start_std"][:nfg], exp, rtol=1e-15)


@pytest.mark.parametrize("name", FIXTURES)
def test_config_reader_roundtrip(name, tmp_path):
    fx = load_fixture(name)
    cfg, _ = write_workspace(fx, str(tmp_path), name="WS_" + name)
    compare(dump(cfg), fx)


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference workspaces only exist in the build container")
@pytest.mark.parametrize("cfg,name", [("Config_BaM_BaRatin.txt", "baratin"), ("Config_BaM_BaRatin_SPD.txt", "baratin_spd"),
                                      ("Config_BaM_Regression_X2Y2.txt", "regression_x2y2"),
                                      ("Config_BaM_TextFile_InputUncertainty.txt", "textfile_inputuncertainty"),
                                      ("Config_BaM_Mixture.txt", "mixture"), ("Config_BaM_SFD.txt", "sfd"),
                                      ("Config_BaM_HydraulicControl_section.txt", "hydraulic_control_section"),
                                      ("Config_BaM_Orthorectification.txt", "orthorectification"),
                                      ("Config_BaM_Regression_X2Y1.txt", "regression_x2y1"), ("Config_BaM_Sediment.txt", "sediment"),
                                      ("Config_BaM_TextFileModel.txt", "textfilemodel")])
def test_config_reader_on_shipped_workspaces(cfg, name):
    """the reference's own example workspaces, Windows path separators included"""
    compare_d = dump_ref(cfg)
    compare(compare_d, load_fixture(name))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference workspaces only exist in the build container")
def test_config_reader_on_shipped_second_wave_workspace():
    d = dump_ref("Config_BaM_Orthorectification.txt")
    assert (d["model_id"], d["nobs"], d["nX"], d["nY"]) == ("Orthorectification", 19, 3, 2)
    assert list(d["xtra_i"]) == [1, 0] and len(d["partype"]) == 11  # "None" distortion, 0 distortion parameters
    d = dump_ref("Config_BaM_HydraulicControl_section.txt")
    assert (d["model_id"], d["nobs"], d["nX"], d["nY"]) == ("HydraulicControl_section", 6, 1, 1)
    tab = np.asarray(d["xtra_d"]).reshape(3, -1)  # h | A | w
    assert tab.shape[1] >= 3 and (np.diff(tab[0]) >= 0).all() and (tab[1, 1:] > 0).all()
    d = dump_ref("Config_BaM_SFD.txt")
    assert (d["model_id"], d["nobs"], d["nX"], d["nY"]) == ("SFD", 68, 2, 1) and list(d["xtra_i"]) == [1, 100]
    assert d["xtra_d"] == [20, 5, 1000, 1e-05, 1e-10]
    # BaRatinBAC ships without a top-level Config_BaM file: point one at the workspace
    cfg = "/tmp/_Config_BaM_BaRatinBAC.txt"
    open(cfg, "w").write("\n".join(['"/root/reference/tests/BaM_BaRatinBAC/"', '"Config_RunOptions.txt"', '"Config_Model.txt"',
                                    '"Config_ControlMatrix.txt"', '"Config_Data.txt"', '"Config_RemnantSigma.txt"', '"Config_MCMC.txt"',
                                    '"Config_Cooking.txt"', '"Config_Summary.txt"', '"Config_Residuals.txt"',
                                    '"Config_Pred_Master.txt"']) + "\n")
    out = cfg + ".json"
    r = subprocess.run([EXE, "-cf", cfg, "--dump-problem", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.load(open(out))
    assert (d["model_id"], d["nobs"]) == ("BaRatinBAC", 104) and list(d["xtra_i"]) == [1, 0, 0, 0, 1, 1, 0, 0, 1] and d["xtra_d"] == [5]
    d = dump_ref(os.path.join("BaM_Segmentation", "Config_BaM.txt"))  # a Config_BaM.txt stored inside its workspace
    assert (d["model_id"], d["nobs"]) == ("Segmentation", 110) and list(d["xtra_i"]) == [2, 2, 1]


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference workspaces only exist in the build container")
def test_config_reader_on_shipped_time_recursive_workspace():
    """tests/BaM_SFDTidal_Qmec2 (14 401 steps; the fixture keeps the first 8000): same inputs, FIX pattern and start"""
    d = dump_ref("Config_BaM_SFDTidal_Qmec2.txt")
    fx = load_fixture("sfdtidal_qmec2")
    assert (d["model_id"], d["nobs"], d["nX"], d["nY"]) == ("SFDTidal_Qmec2", 14401, 2, 1)
    assert list(d["partype"]) == list(fx["partype"]) and np.array_equal(np.asarray(d["start"]), np.asarray(fx["start"]))
    X = np.asarray(d["X"]).reshape(2, 14401)
    Xf = np.asarray(fx["X"]).reshape(2, 8000)
    assert np.array_equal(X[:, :8000], Xf)
    Y = np.asarray(d["Y"])
    assert np.array_equal(Y[:8000], np.asarray(fx["Y"])) and (Y != -9999.0).sum() == 264


def dump_ref(cfg):
    out = "/tmp/_dump_" + cfg.replace("/", "_") + ".json"
    r = subprocess.run([EXE, "-cf", os.path.join("/root/reference/tests", cfg), "--dump-problem", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.load(open(out))


def test_cli_version_and_errors(tmp_path):
    r = subprocess.run([EXE, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and "version" in r.stdout
    r = subprocess.run([EXE, "-cf", str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "FATAL ERROR" in r.stderr


def read_table(path):
    lines = open(path).read().splitlines()
    head = [lines[0][i:i + 15].strip() for i in range(0, len(lines[0]), 15)]
    rows = np.array([[float(ln[i:i + 15]) for i in range(0, len(ln), 15)] for ln in lines[1:] if ln.strip()])
    return head, rows


@pytest.mark.gpu
def test_full_run_writes_reference_files(tmp_path):
    fx = load_fixture("baratin")
    H = np.linspace(-1, 6, 41)
    cfg, ws = write_workspace(fx, str(tmp_path), name="RUN", nAdapt=20, nCycles=10, do_pred=True,
                              pred=[dict(X=H, doParametric=True, doRemnant=True), dict(X=H, doParametric=False, doRemnant=False),
                                    dict(X=H, doParametric=True, doRemnant=False, nsim=50)])
    r = subprocess.run([EXE, "-cf", cfg, "-sd", "7", "--chains", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    head, rows = read_table(os.path.join(ws, "Results_MCMC.txt"))
    assert head == ["k1", "a1", "c1", "k2", "a2", "c2", "Y1_gamma1", "Y1_gamma2", "LogPost", "b1", "b2"]
    assert rows.shape == (200, 11) and np.isfinite(rows).all()
    assert np.allclose(rows[:, 9], rows[:, 0], rtol=1e-6)  # b1 = k1, 7 significant digits (e15.6)
    _, cooked = read_table(os.path.join(ws, "Results_MCMC_Cooked.txt"))
    assert cooked.shape == (4 * 50, 11)  # burn 50 %, slim 2, 4 chains pooled
    lines = open(os.path.join(ws, "Results_Summary.txt")).read().splitlines()
    assert len(lines) == 17 and lines[1].startswith("N") and lines[16].startswith("MaxPost")
    assert open(os.path.join(ws, "Config_MCMC.txt.monitor")).read().strip() == "200/200"
    _, rhat = read_table(os.path.join(ws, "Results_Rhat.txt"))
    assert rhat.shape == (1, 8)
    head, env = read_table(os.path.join(ws, "Y_1.env"))
    assert head == ["Median", "q2.5", "q97.5", "q16", "q84", "Mean", "Stdev"] and env.shape == (41, 7)
    assert (env[:, 1] <= env[:, 0]).all() and (env[:, 0] <= env[:, 2]).all()
    spag = np.loadtxt(os.path.join(ws, "Y_1.spag"))
    assert spag.shape == (41, 200)  # transposed: one row per input
    mp = np.loadtxt(os.path.join(ws, "Y_2.spag"))
    assert mp.shape == (41,) and not os.path.exists(os.path.join(ws, "Y_2.env"))  # maxpost: Nrep = 1, no envelope
    assert np.loadtxt(os.path.join(ws, "Y_3.spag")).shape == (41, 50)  # prior prediction with nsim = 50


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["baratin", "textfile_inputuncertainty", "baratin_spd"])
def test_residuals_and_dic_files(name, tmp_path):
    """BaM_Residual (:754-878) and BaM_computeDIC (:1526-1682) on top of the device kernels: the residual table
    against the oracle's calibration run at the written maxpost, the DIC identities against the extended MCMC."""
    from bam_b200.capi import BamGpu
    from oracle.oracle_api import Oracle
    from tests.helpers import problem_from_fixture

    fx = load_fixture(name)
    cfg, ws = write_workspace(fx, str(tmp_path), name="RD", nAdapt=10, nCycles=6, do_dic=True, do_residual=True)
    r = subprocess.run([EXE, "-cf", cfg, "-sd", "3", "--chains", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    prob, _ = problem_from_fixture(name)
    o = Oracle(prob)
    P = o.P
    head, cooked = read_table(os.path.join(ws, "Results_MCMC_Cooked.txt"))
    # ---- DIC
    hx, ext = read_table(os.path.join(ws, "Results_MCMC_Xtended.txt"))
    assert hx[P:] == ["LogPost", "LogPrior", "LogHyper", "LogLkh", "Deviance"] and ext.shape == (cooked.shape[0], P + 5)
    lp, pr, hy, lk, dev = (ext[:, P + i] for i in range(5))
    assert np.allclose(lk, lp - pr - hy, rtol=1e-5, atol=2e-3) and np.allclose(dev, -2 * lk, rtol=1e-5)  # e15.6: 6 digits
    opr, olk, ohy, ofe, onu = o.logpost_parts(cooked[:6, :P])  # the file holds 7 digits of every parameter
    assert ofe.all()
    dic = dict((ln[:15].strip(), float(ln[15:])) for ln in open(os.path.join(ws, "Results_DIC.txt")).read().splitlines())
    assert list(dic) == ["DIC1", "DIC2", "DIC3", "D(maxpost)", "E[D]", "VAR[D]", "MEDIAN[D]", "SDEV[D]"]
    k = int(np.argmax(lp))
    assert np.isclose(dic["E[D]"], dev.mean(), rtol=1e-5) and np.isclose(dic["VAR[D]"], dev.var(ddof=1), rtol=1e-4)
    assert np.isclose(dic["D(maxpost)"], dev[k], rtol=1e-5)
    assert np.isclose(dic["DIC1"], 2 * dic["E[D]"] - dic["D(maxpost)"], rtol=1e-5)
    assert np.isclose(dic["DIC3"], 0.5 * (dic["DIC1"] + dic["DIC2"]), rtol=1e-5)
    # ---- residuals: device calibration run vs oracle at full precision, then the written table
    g = BamGpu(prob)
    x = cooked[k, :P]
    Xt, Ys, feas = g.calibration_run(x)
    oXt, oYs, ofeas = o.calibration_run(x)
    assert feas and ofeas and np.array_equal(Xt, oXt)
    assert np.allclose(Ys, oYs, rtol=1e-12, atol=1e-300)
    hr, res = read_table(os.path.join(ws, "Results_Residuals.txt"))
    nX, nY, n = prob.nX, prob.nY, prob.nobs
    assert res.shape == (n, 2 * nX + 5 * nY) and hr[0] == "X1_obs" and hr[-1] == f"Y{nY}_stdres"
    Ysim_file = res[:, 2 * nX + 2 * nY:2 * nX + 3 * nY]
    # the driver's maxpost carries full precision, the cooked file 7 digits: compare loosely with the oracle run
    ok = np.asarray(prob.Y) != -9999.0
    assert np.allclose(Ysim_file, oYs, rtol=2e-3, atol=1e-6)
    resid = res[:, 2 * nX + 3 * nY:2 * nX + 4 * nY]
    yunb = res[:, 2 * nX + nY:2 * nX + 2 * nY]
    assert np.allclose(resid[ok], (yunb - Ysim_file)[ok], rtol=1e-4, atol=1e-5 * np.abs(yunb[ok]).max())
    assert (resid[~ok] == -9999.0).all()
    std = res[:, 2 * nX + 4 * nY:]
    assert np.isfinite(std[ok]).all() and np.abs(std[ok]).max() < 50


@pytest.mark.gpu
def test_onerun_dontrun_saveprior_modes(tmp_path):
    """the remaining command-line modes of BaM_main (src/BaM_main.f90:100-160): --onerun (model run at the initial
    values, :215-247), --dontrun (INFO_BaM.txt only, :308), --saveprior (prior sample of a prior prediction, :984-1004)"""
    from oracle.oracle_api import Oracle
    from tests.helpers import problem_from_fixture

    fx = load_fixture("baratin")
    H = np.linspace(-1, 6, 41)
    cfg, ws = write_workspace(fx, str(tmp_path), name="OR", nAdapt=5, nCycles=2, do_pred=True,
                              pred=[dict(X=H, doParametric=True, doRemnant=False, nsim=30)])
    # --onerun
    np.savetxt(os.path.join(ws, "Hin.txt"), np.column_stack([H, 2 * H]), header="H junk", comments="")
    open(os.path.join(ws, "Config_Inputs.txt"), "w").write('"Hin.txt"\n1\n41\n2\n')
    xtra = open(cfg).read().splitlines()[3].split('"')[1]  # 4th record of Config_BaM: the model's Xtra file
    one = os.path.join(str(tmp_path), "Config_OneRun.txt")
    open(one, "w").write(f'"{ws}/"\n"Config_Model.txt"\n"{xtra}"\n"Config_Inputs.txt"\n')
    r = subprocess.run([EXE, "-cf", one, "-or", "Y_onerun.txt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    head, Y = read_table(os.path.join(ws, "Y_onerun.txt"))
    prob, _ = problem_from_fixture("baratin")
    Yo, feas, _ = Oracle(prob).apply_model(H, np.asarray(fx["start"])[:6])
    assert head == ["Y1"] and Y.shape == (41, 1) and np.allclose(Y[:, 0], Yo[:, 0], rtol=1e-5, atol=1e-9)  # e15.6: 6 digits
    # --dontrun: only INFO_BaM.txt
    r = subprocess.run([EXE, "-cf", cfg, "-dr"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(os.path.join(ws, "INFO_BaM.txt")) and not os.path.exists(os.path.join(ws, "Results_MCMC.txt"))
    # --saveprior with a prior-prediction experiment (nsim = 30)
    r = subprocess.run([EXE, "-cf", cfg, "-sd", "5", "-sp", "PriorSample.txt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hp, pri = read_table(os.path.join(ws, "PriorSample.txt"))
    assert hp[:3] == ["k1", "a1", "c1"] and pri.shape == (30, 8)
    assert abs(pri[:, 1].mean() - 53.0) < 10.0  # a1 ~ N(53, 10) in Config_Model of tests/BaM_BaRatin


@pytest.mark.gpu
def test_state_variable_files(tmp_path):
    """SFD (tests/BaM_SFD): the state variable kappa goes to its own spaghetti / envelope files (BaM_Prediction
    :1071-1075,1121-1132, Config_Read_Pred :2808-2826) and to the TransitionStage column of the residual file."""
    from oracle.oracle_api import Oracle
    from tests.helpers import problem_from_fixture

    fx = load_fixture("sfd")
    h1 = np.linspace(1.0, 7.0, 30)
    h2 = np.full(30, 1.93)
    cfg, ws = write_workspace(fx, str(tmp_path), name="ST", nAdapt=10, nCycles=4, do_pred=True, do_residual=True,
                              pred=[dict(Xs=[h1, h2], doParametric=True, doRemnant=True, state=True),
                                    dict(Xs=[h1, h2], doParametric=False, doRemnant=False, state=True),
                                    dict(Xs=[h1, h2], doParametric=True, doRemnant=False)])
    r = subprocess.run([EXE, "-cf", cfg, "-sd", "11", "--chains", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _, cooked = read_table(os.path.join(ws, "Results_MCMC_Cooked.txt"))
    nrep = cooked.shape[0]
    S = np.loadtxt(os.path.join(ws, "S_1.spag"))
    Y = np.loadtxt(os.path.join(ws, "Y_1.spag"))
    assert S.shape == (30, nrep) and Y.shape == (30, nrep)
    head, env = read_table(os.path.join(ws, "S_1.env"))
    assert head[0] == "Median" and env.shape == (30, 7)
    assert os.path.exists(os.path.join(ws, "Y_3.spag")) and not os.path.exists(os.path.join(ws, "S_3.spag"))
    # maxpost run: the state column equals the oracle's kappa at the written maxpost (7 digits in the file)
    prob, _ = problem_from_fixture("sfd")
    o = Oracle(prob)
    k = int(np.argmax(cooked[:, o.P]))
    mp = cooked[k, :o.P]
    _, ospag = o.predict(None, mp, [h1, h2], 0, [0], want_env=False, want_spag=True, states=True)
    S2 = np.loadtxt(os.path.join(ws, "S_2.spag"))
    assert S2.shape == (30,) and np.allclose(S2, ospag[1, :, 0], rtol=1e-4)
    head, res = read_table(os.path.join(ws, "Results_Residuals.txt"))
    assert head[-1] == "TransitionStage" and res.shape == (68, 2 * 2 + 5 * 1 + 1)
    _, _, oSt, _ = o.calibration_run_states(mp)
    assert np.allclose(res[:, -1], oSt[:, 0], rtol=1e-4)


@pytest.mark.gpu
def test_time_recursive_workspace_end_to_end(tmp_path):
    """tests/BaM_SFDTidal_Qmec2 (8000-step prefix) through the driver: MCMC on the 6 free parameters (5 are FIX), cooked
    sample, residual file with the three state columns of the model (pressure, friction, advection)"""
    from oracle.oracle_api import Oracle
    from tests.helpers import problem_from_fixture

    fx = load_fixture("sfdtidal_qmec2")
    cfg, ws = write_workspace(fx, str(tmp_path), name="TR", nAdapt=5, nCycles=2, do_residual=True)
    compare(dump(cfg), fx)  # the written workspace reads back as the fixture (FIX parameters included)
    r = subprocess.run([EXE, "-cf", cfg, "-sd", "5", "--chains", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    head, rows = read_table(os.path.join(ws, "Results_MCMC.txt"))
    assert head[:6] == ["Be", "bevs", "delta", "ne", "Q0", "Y1_gamma1"][:6] or head[4] == "Q0"
    assert rows.shape[0] == 10 and np.isfinite(rows).all()
    hr, res = read_table(os.path.join(ws, "Results_Residuals.txt"))
    assert hr[-3:] == ["pressure", "friction", "advection"] and res.shape == (8000, 2 * 2 + 5 * 1 + 3)
    _, cooked = read_table(os.path.join(ws, "Results_MCMC_Cooked.txt"))
    prob, _ = problem_from_fixture("sfdtidal_qmec2")
    o = Oracle(prob)
    k = int(np.argmax(cooked[:, o.P]))
    _, oYs, oSt, ofeas = o.calibration_run_states(cooked[k, :o.P])
    assert ofeas
    # the cooked file holds 7 digits of every parameter and the recursion integrates over 8000 steps: loose comparison
    assert np.allclose(res[:, 6], oYs[:, 0], rtol=5e-3, atol=5.0)
    assert np.allclose(res[1:, -3:].sum(axis=1), np.diff(res[:, 6]), rtol=1e-4, atol=0.1)  # Q(m) - Q(m-1) = sum of the states (7 digits in the file)


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference workspaces only exist in the build container")
@pytest.mark.parametrize("cfg,name,model,n,nX", [("Config_BaM_SFDTidal2.txt", "sfdtidal2", "SFDTidal2", 1372, 2),
                                                 ("Config_BaM_SFDTidal4.txt", "sfdtidal4", "SFDTidal4", 2881, 2),
                                                 ("Config_BaM_SFDTidal_Jones.txt", "sfdtidal_jones", "SFDTidalJones", 1372, 3)])
def test_config_reader_on_the_lagged_sfdtidal_workspaces(cfg, name, model, n, nX):
    """the C++ config reader and the independent Python fixture reader agree on the shipped workspaces of the lagged
    stage-fall-discharge models"""
    d = dump_ref(cfg)
    fx = load_fixture(name)
    assert (d["model_id"], d["nobs"], d["nX"], d["nY"]) == (model, n, nX, 1)
    assert list(d["partype"]) == list(fx["partype"]) and np.array_equal(np.asarray(d["start"]), np.asarray(fx["start"]))
    assert np.array_equal(np.asarray(d["X"]), np.asarray(fx["X"])) and np.array_equal(np.asarray(d["Y"]), np.asarray(fx["Y"]))
