"""Write a BaM workspace (Config_BaM.txt + Config_*.txt + data files, reference grammar: SURVEY.md App. C)
from a golden fixture, so that the C++ host driver can be exercised on the GPU box, where the reference's
own example workspaces (/root/reference/tests) do not exist."""
import os

import numpy as np


def _par_block(name, init, dist, pars):
    return [f'"{name}"   ! Parameter Name', f"{init!r}   ! Initial guess", f"'{dist}'   ! Prior distribution",
            ",".join(repr(float(p)) for p in pars) + "   ! Prior parameters"]


def write_workspace(fx: dict, root: str, name: str = "WS", nAdapt=None, nCycles=None, do_pred=False, pred=None,
                    windows_paths=True, do_dic=False, do_residual=False):
    ws = os.path.join(root, name)
    os.makedirs(ws, exist_ok=True)
    n, nX, nY = fx["nobs"], fx["nX"], fx["nY"]
    ncol = fx["data_ncol"]
    W = np.asarray(fx["data_table"]).reshape(ncol, n).T
    with open(os.path.join(ws, "data.txt"), "w") as f:
        f.write("\t".join(f"c{j + 1}" for j in range(ncol)) + "\n")
        for r in W:
            f.write("\t".join(repr(float(v)) for v in r) + "\n")

    def find_cols(key, ncols, is_int=False):
        if fx.get(key) is None:
            return [0] * ncols
        M = np.asarray(fx[key], dtype=float).reshape(ncols, n).T
        out = []
        for j in range(ncols):
            if not M[:, j].any():
                out.append(0)
                continue
            hit = [c + 1 for c in range(ncol) if np.array_equal(W[:, c], M[:, j])]
            assert hit, (key, j)
            out.append(hit[0])
        return out

    sep = "\\" if windows_paths else "/"
    lines = [f"'{name}{sep}data.txt'   !!! data file", "1  ! header lines", f"{n} ! nobs", f"{ncol} ! ncol"]
    for key, nc in (("X", nX), ("Xu", nX), ("Xb", nX), ("Xbindx", nX), ("Y", nY), ("Yu", nY), ("Yb", nY), ("Ybindx", nY)):
        lines.append(",".join(str(c) for c in find_cols(key, nc)) + f"   !!! columns for {key}")
    open(os.path.join(ws, "Config_Data.txt"), "w").write("\n".join(lines) + "\n")

    # model
    names = fx.get("parnames") or [f"p{i}" for i in range(len(fx["partype"]))]
    lines = [f'"{fx["model_id"]}"  ! Model ID', f"{nX}", f"{nY}", f"{len(fx['partype'])}"]
    kfit = 0
    start = fx["start"]
    vi = None if fx.get("var_indx") is None else np.asarray(fx["var_indx"]).reshape(len(fx["partype"]), n).T
    for i, t in enumerate(fx["partype"]):
        if t == -1:
            lines += [f'"{names[i]}"', repr(float(fx["fix_value"][i])), "'FIX'", "0  ! ignored"]
        elif t == 1:
            nv = fx["nval"][i]
            col = [c + 1 for c in range(ncol) if np.array_equal(W[:, c].astype(int), vi[:, i])][0]
            vf = f"Config_{names[i]}_VAR.txt"
            lines += [f'"{names[i]}"', repr(float(start[kfit])), "'VAR'", f"'{vf}'"]
            vl = [f"{nv}  ! nVal", f"{col}  ! col"]
            for j in range(nv):
                d, p = fx["priors_theta"][kfit + j]
                vl += _par_block(f"{names[i]}_{j + 1}", float(start[kfit + j]), d, p)
            open(os.path.join(ws, vf), "w").write("\n".join(vl) + "\n")
            kfit += nv
        else:
            d, p = fx["priors_theta"][kfit]
            lines += _par_block(names[i], float(start[kfit]), d, p)
            kfit += 1
    open(os.path.join(ws, "Config_Model.txt"), "w").write("\n".join(lines) + "\n")

    # remnant sigma, one file per output
    NP = {"Constant": 1, "Linear": 2, "Proportional": 1, "Power": 3, "Exponential": 3, "Gaussian": 3}
    kg, rfiles = 0, []
    for y, funk in enumerate(fx["sigma_func"]):
        rl = [f"'{funk}'  ! function", f"{NP[funk]}  ! npar"]
        for j in range(NP[funk]):
            d, p = fx["priors_gamma"][kg]
            rl += _par_block(f"gamma{j + 1}", float(start[kfit + kg]), d, p)
            kg += 1
        rf = f"Config_RemnantSigma_Y{y + 1}.txt"
        open(os.path.join(ws, rf), "w").write("\n".join(rl) + "\n")
        rfiles.append(rf)

    # xtra
    xf = "Config_Xtra.txt"
    if fx["model_id"] == "BaRatin":
        nc = int(round(len(fx["xtra_i"]) ** 0.5))
        M = np.asarray(fx["xtra_i"]).reshape(nc, nc).T
        open(os.path.join(ws, xf), "w").write("\n".join(" ".join(str(int(v)) for v in r) for r in M) + "\n")
    elif fx["model_id"] == "Sediment":
        F = {1: "MPM", 2: "Camenen", 3: "Nielsen", 4: "Smart"}
        C = {1: "Constant", 2: "Soulsby", 3: "SoulsbyWhitehouse"}
        O = {1: "FIX", 2: "IN", 3: "PAR"}
        xi = fx["xtra_i"]
        open(os.path.join(ws, xf), "w").write(
            f'"{F[xi[0]]}"\n"{C[xi[1]]}"\n"{ {1: "FIX", 2: "PAR"}[xi[2]] }"\n' + ",".join(f'"{O[v]}"' for v in xi[3:7]) +
            "\n" + ",".join(repr(float(v)) for v in fx["xtra_d"]) + "\n")
    elif fx["model_id"] == "TextFile":
        open(os.path.join(ws, xf), "w").write(fx["xtra_text"])
    elif fx["model_id"] == "SFD":
        xd = fx["xtra_d"]
        open(os.path.join(ws, xf), "w").write('"SFD_Val_General"\n' + "\n".join(repr(float(v)) for v in xd) + f"\n{fx['xtra_i'][1]}\n")
    elif fx["model_id"] == "Mixture":
        xi = fx["xtra_i"]
        open(os.path.join(ws, xf), "w").write(f"{xi[0]} ! nS\n{xi[1]} ! nEvt\n{xi[2]} ! nElt\n{'.true.' if xi[3] else '.false.'} ! missing source\n")
    else:
        open(os.path.join(ws, xf), "w").write("\n")

    if fx.get("priorC") is not None:
        k = int(round(len(fx["priorC"]) ** 0.5))
        Cm = np.asarray(fx["priorC"]).reshape(k, k).T
        open(os.path.join(ws, "PriorCorrelation.txt"), "w").write("\n".join(" ".join(repr(float(v)) for v in r) for r in Cm) + "\n")

    m = fx["mcmc"]
    open(os.path.join(ws, "Config_MCMC.txt"), "w").write("\n".join([
        '"Results_MCMC.txt"', str(nAdapt or m["nAdapt"]), str(nCycles or m["nCycles"]), repr(m["MinMoveRate"]),
        repr(m["MaxMoveRate"]), repr(m["DownMult"]), repr(m["UpMult"]), "0", "**** cosmetics ****",
        repr(m["MultFactor"]), "0.1", "0.1"]) + "\n")
    open(os.path.join(ws, "Config_Cooking.txt"), "w").write('"Results_MCMC_Cooked.txt"\n0.5\n2\n')
    open(os.path.join(ws, "Config_Summary.txt"), "w").write(
        '"Results_Summary.txt"\n' + ('"Results_DIC.txt"\n"Results_MCMC_Xtended.txt"\n' if do_dic else ""))
    open(os.path.join(ws, "Config_Residuals.txt"), "w").write('"Results_Residuals.txt"\n')
    open(os.path.join(ws, "Config_RunOptions.txt"), "w").write(
        f".true.\n.true.\n{'.true.' if do_residual else '.false.'}\n{'.true.' if do_pred else '.false.'}\n")
    plines = ["0"]
    if do_pred and pred:
        plines = [str(len(pred))]
        for i, pc in enumerate(pred):
            pf = f"Config_Pred_{i + 1}.txt"
            plines.append(f"'{pf}'")
            tf = lambda b: ".true." if b else ".false."  # noqa: E731
            if "Xs" in pc:  # one series per input variable
                cols = [np.asarray(c, dtype=float).reshape(-1, 1) for c in pc["Xs"]]
                xfiles = []
                for j, c in enumerate(cols):
                    xfiles.append(f"Xpred_{i + 1}_{j + 1}.txt")
                    np.savetxt(os.path.join(ws, xfiles[-1]), c, fmt="%.17g")
                nobs, nsp = cols[0].shape[0], ",".join("1" for _ in cols)
                xlist = ",".join(f"'{name}{sep}{f}'" for f in xfiles)
            else:
                X = np.atleast_2d(np.asarray(pc["X"], dtype=float))
                if X.shape[0] == 1 and X.shape[1] > 1 and pc.get("column", True):
                    X = X.T
                xfile = f"Xpred_{i + 1}.txt"
                np.savetxt(os.path.join(ws, xfile), X, fmt="%.17g")
                nobs, nsp, xlist = X.shape[0], str(X.shape[1]), f"'{name}{sep}{xfile}'"
            lines = [xlist, str(nobs), nsp, tf(pc["doParametric"]), tf(pc["doRemnant"]),
                     str(pc.get("nsim", -1)), f"'Y_{i + 1}.spag'", tf(pc.get("transpose", True)), ".true.", f"'Y_{i + 1}.env'",
                     ".false."]
            if pc.get("state") is not None:  # state section (Config_Read_Pred :2808-2826)
                lines += [tf(pc["state"]), f"'S_{i + 1}.spag'", tf(pc.get("transpose", True)), ".true.", f"'S_{i + 1}.env'"]
            else:
                lines += [".false."]
            open(os.path.join(ws, pf), "w").write("\n".join(lines) + "\n")
    open(os.path.join(ws, "Config_Pred_Master.txt"), "w").write("\n".join(plines) + "\n")

    cfg = os.path.join(root, f"Config_BaM_{name}.txt")
    open(cfg, "w").write("\n".join([
        f'"{name}{sep}"   ! workspace', '"Config_RunOptions.txt"', '"Config_Model.txt"', f'"{xf}"', '"Config_Data.txt"',
        ",".join(f'"{r}"' for r in rfiles), '"Config_MCMC.txt"', '"Config_Cooking.txt"', '"Config_Summary.txt"',
        '"Config_Residuals.txt"', '"Config_Pred_Master.txt"']) + "\n")
    return cfg, ws
