#!/usr/bin/env python
"""tests/golden/vegetation_synthetic.json from the reference's tests/BaM_Vegetation/SyntheticData.txt.

Run HERE (needs /root/reference).  The data file carries generator-side known-answer columns
(g0, g, Q0, gamma, Q, cosg, sing) printed with 9-10 significant digits.  The generator's parameters are not
shipped; they are recovered from parameter-free relations of the columns (all verified to the file's precision):
    Q0 = K (stage-1)^(5/3),  K = 11.0679718 = B sqrt(S0)/n1 with B=10, S0=0.001, n1=1/35
    gamma / (g (stage-1)^(1/3)) = 1/(8 * 9.81 * n1^2)
    Q = Q0 / sqrt(1+gamma)                      (no bending: chi = 0)
    cosg = cos(pi g0), sing = +-sin(pi g0)
    degree-days above Tmin=14 (column Tcumul), t_b = first Tcumul >= 10, t_e = first Tcumul >= 200, t_f = t_e + 35/365,
    g0 = ((t-t_b)/(t_e-t_b))^alpha (t_f-t)/(t_f-t_e), alpha = (t_e-t_b)/(t_f-t_e)   == Growth_Yin3 of the code
    g = gmax g0 with gmax = 1.
Only the first two years (730 rows) are kept to bound the fixture size.
"""
import json
import os

import numpy as np

REF = os.environ.get("BAM_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "tests", "BaM_Vegetation", "SyntheticData.txt")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vegetation_synthetic.json")


def main():
    hdr = open(SRC).readline().split()
    d = np.loadtxt(SRC, skiprows=1)
    d = d[np.floor(d[:, hdr.index("time")]) < 2]
    col = lambda name: d[:, hdr.index(name)].tolist()  # noqa: E731
    fx = {
        "source": "tests/BaM_Vegetation/SyntheticData.txt (rows with floor(time) < 2)",
        "nobs": int(len(d)),
        "nYear": 2,
        "growth": "Growth_Yin3",
        "theta": [10.0, 0.001, 1.0 / 35.0, 1.0, 5.0 / 3.0, 0.0, 1.0, 14.0, 10.0, 200.0, 35.0 / 365.0, 1.0, 1.0],
        "theta_names": ["B", "S0", "n1", "b1", "c1", "chi", "U_chi", "Tmin", "Tb", "Te", "tf-te", "gmax_1", "gmax_2"],
        "newton": [5.0, 1000.0, 1e-5, 1e-10, 100],
    }
    for name in ("time", "stage", "Temp", "Tcumul", "g0", "g", "Q0", "gamma", "Q", "cosg", "sing"):
        fx[name] = col(name)
    with open(OUT, "w") as f:
        json.dump(fx, f)
    print("fixture vegetation_synthetic", fx["nobs"], os.path.getsize(OUT))


if __name__ == "__main__":
    main()
