import sys; sys.path.insert(0, '.')
import numpy as np
from bam_b200.capi import BamGpu, Problem
from oracle.oracle_api import Oracle
from tests.helpers import perturbed_vectors
rng = np.random.default_rng(3)
for nc, funk, gam in ((1, "Constant", [0.8]), (2, "Proportional", [0.3]), (3, "Linear", [0.5, 0.1]), (4, "Linear", [0.5, 0.1]), (2, "Linear", [0.5, 0.1])):
    lowest = 0.2 if funk == "Proportional" else -3.0
    k = np.array([-0.5, 0.7, 1.9, 3.1])[:nc]
    a = np.array([20.0, 35.0, 50.0, 65.0])[:nc]
    c = np.array([1.5, 1.67, 1.67, 2.0])[:nc]
    M = np.tril(np.ones((nc, nc), dtype=np.int32))
    theta = np.column_stack([k, a, c]).ravel()
    H = np.concatenate([rng.uniform(max(-1.0, lowest), 4.5, 500), k[1:], [np.nextafter(k[1], 9) if nc > 1 else 1.0, lowest]])
    if funk != "Proportional":
        H = np.concatenate([H, [-0.5, np.nextafter(-0.5, 1)]])
    Y = np.abs(rng.standard_normal(H.size)) * 50 + 1.0
    Y[5] = -9999.0
    Yu = 0.05 * Y
    prob = Problem("BaRatin", H, Y, Yu=np.abs(Yu), partype=[0] * (3 * nc), priors_theta=[("Gaussian", [t, abs(t) * 0.3 + 0.1]) for t in theta],
                   sigma_func=[funk], priors_gamma=[("Uniform", [0, 100])] * len(gam), xtra_i=M.flatten(order="F"))
    x = perturbed_vectors(np.concatenate([theta, gam]), 200, rel=0.01, seed=nc)
    g, o = BamGpu(prob), Oracle(prob)
    ofx, ofeas, _ = o.logpost(x, nthreads=8)
    fx, feas, _ = g.logpost(x)
    g.set_option("force_generic", 1)
    gx, gfeas, _ = g.logpost(x)
    ok = ofeas == 1
    print(nc, funk, "feas", ok.sum(), "fast err", np.abs(fx[ok]-ofx[ok]).max()/np.abs(ofx[ok]).max() if ok.any() else None,
          "generic err", np.abs(gx[ok]-ofx[ok]).max()/np.abs(ofx[ok]).max() if ok.any() else None, "flags", (feas==ofeas).all(), (gfeas==ofeas).all())
    if ok.any():
        e = np.abs(fx-ofx)/np.abs(ofx)
        print("   worst idx", np.argsort(-e)[:6], e[np.argsort(-e)[:6]])
    # subset search: which observation breaks it
    if ok.any() and (np.abs(fx[ok]-ofx[ok]).max()/np.abs(ofx[ok]).max() > 1e-10):
        for sel in (slice(0, 500), slice(500, None)):
            P2 = Problem("BaRatin", H[sel], Y[sel], Yu=np.abs(Yu)[sel], partype=[0] * (3 * nc), priors_theta=[("Gaussian", [t, abs(t) * 0.3 + 0.1]) for t in theta],
                         sigma_func=[funk], priors_gamma=[("Uniform", [0, 100])] * len(gam), xtra_i=M.flatten(order="F"))
            g2, o2 = BamGpu(P2), Oracle(P2)
            a1 = g2.logpost(x)[0]; b1 = o2.logpost(x, nthreads=8)[0]
            print("   subset", sel, np.abs(a1-b1).max()/np.abs(b1).max(), "H tail", H[sel][-6:] if sel.start else "")
