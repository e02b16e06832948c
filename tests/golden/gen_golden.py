#!/usr/bin/env python
"""Generate tests/golden/known_answers.json: 50-digit mpmath values of every closed-form expression
on the hot path, evaluated in exact real arithmetic on double inputs.

The reference ships no golden vectors, so these known-answer vectors are what pins
the oracle's transcription and rounding behaviour.  Re-run:  python tests/golden/gen_golden.py
"""
import json
import os
import struct

import mpmath as mp

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))

SMALL, VSMALL = mp.mpf(1e-15), mp.mpf(1e-300)


def f(x):
    return mp.mpf(float(x))


def hexd(x):
    """nearest double of an mpf, as hex (exact round trip)"""
    return float(mp.nstr(x, 40)).hex() if mp.isfinite(x) else str(x)


def omax(a, b):
    return a if a > b else b


def omin(a, b):
    return a if a < b else b


# ---- parameter sets (same numbers as solidstateamfoam_b200/cases.py) --------------------------------
SW = dict(sigmaR=2.222e7, Q=1.45e5, R=8.314, A=2.409e8, n=3.55, rho=2700.0, nuMin=1e-6, nuMax=1e4, eDotMin=1e-6)
FKS = dict(a1=3.24e8, a2=1.14e8, b1=0.01, b2=1.0, b3=1.0, e0=1.0, c1=6.0, c2=1.0, Tm=855.0, Ta=293.0, rho=2700.0,
           nuMin=1e-6, nuMax=1e4, eDotMin=1e-6)
FKS2 = dict(FKS, b2=0.7, c2=2.5, c1=3.0, e0=0.01, b1=0.05)
NH = dict(mu0=1e8, m=0.2, rho=2700.0, nuMin=1e-6, nuMax=1e4, strainRateEffMin=1e-6)
MIX = dict(rho1=2700.0, rho2=1.2, cp1=896.0, cp2=1005.0)
KCUB = dict(k0=100.0, k1=0.2, k2=-1e-4, k3=1e-8, Tmin=293.0, Tmax=855.0)
KCUB_C = dict(k0=100.0, k1=0.2, k2=-1e-4, k3=1e-8, Tmin=20.0, Tmax=582.0)

T_GRID = [250.0, 293.0, 300.0, 412.5, 600.0, 777.7, 855.0, 900.0, 1200.0]
E_GRID = [0.0, 1e-12, 1e-8, 1e-6, 3.3e-4, 0.1, 1.0, 37.5, 1e4]
A_GRID = [-1e-3, 0.0, 1e-9, 0.25, 0.5, 0.999, 1.0, 1.001]


def sw(T, e):
    p = {k: f(v) for k, v in SW.items()}
    T, e = f(T), f(e)
    ex = mp.exp(p["Q"] / (p["R"] * T))
    sf = p["sigmaR"] * mp.asinh(mp.power(e * ex / p["A"], 1 / p["n"]))
    sy = p["sigmaR"] * mp.asinh(mp.power(f(0.1) * ex / p["A"], 1 / p["n"]))
    nu = omax(p["nuMin"], omin(p["nuMax"], sf / (3 * p["rho"] * omax(e, p["eDotMin"]))))
    return nu, sf, sy


def fks(P, T, e):
    p = {k: f(v) for k, v in P.items()}
    T, e = f(T), f(e)
    Tl = omin(omax(T, p["Ta"]), p["Tm"])
    Th = (Tl - p["Ta"]) / (p["Tm"] - p["Ta"])
    H = p["a1"] + p["a2"] * f(1.571)
    pw = p["b1"] * mp.power(Th + SMALL, p["b2"])
    Lam = 1 + pw * (p["b3"] * mp.log(e / p["e0"] + VSMALL))
    Theta = 1 - 1 / mp.power(1 + mp.exp(-p["c1"] * Th), 1 / p["c2"])
    sf = H * Lam * Theta
    sy = p["a1"] * Theta * (1 + pw * (p["b3"] * mp.log(f(0.1) / p["e0"] + VSMALL)))
    nu = omax(p["nuMin"], omin(p["nuMax"], sf / (3 * p["rho"] * omax(e, p["eDotMin"]))))
    return nu, sf, sy


def nh(sr):
    p = {k: f(v) for k, v in NH.items()}
    sr = f(sr)
    return omax(p["nuMin"], omin(p["nuMax"], (p["mu0"] / p["rho"]) * mp.power(mp.sqrt(3) * omax(sr, p["strainRateEffMin"]), p["m"] - 1)))


def mixture(alpha, nu1, nu2, T, kc, shift):
    m = {k: f(v) for k, v in MIX.items()}
    a, nu1, nu2, T = f(alpha), f(nu1), f(nu2), f(T)
    la = omin(omax(a, mp.mpf(0)), mp.mpf(1))
    mu = la * la * m["rho1"] * nu1 + (1 - la) * (1 - la) * m["rho2"] * nu2
    rho = la * m["rho1"] + (1 - la) * m["rho2"]
    nu = mu / rho
    cp = la * m["cp1"] + (1 - la) * m["cp2"]
    k = {kk: f(v) for kk, v in kc.items()}
    Tl = omin(omax(T + f(shift), k["Tmin"]), k["Tmax"])
    k1 = k["k0"] + k["k1"] * Tl + k["k2"] * Tl ** 2 + k["k3"] * Tl ** 3
    kmix = la * k1 + (1 - la) * f(0.026)
    return mu, nu, rho, cp, kmix


def qfric(model, alpha, sy, p, r):
    alpha, sy, p, r = f(alpha), f(sy), f(p), f(r)
    om = f(41.9)
    betaTQ = f(0.9)
    if model == "constant":
        d, muf = f(0.5), f(0.3)
    else:
        d = 1 - mp.exp(-om * r / (f(0.3) * f(41.9) * f(19e-3)))
        muf = f(0.4) * mp.exp(-f(0.5) * d * om * r)
    return alpha * ((1 - d) * betaTQ * sy / mp.sqrt(3) + d * muf * p) * (om * r)


def main():
    out = {"note": "mpmath 60-digit known answers; values stored as hex of the nearest double", "sw": [], "fks": [],
           "fks2": [], "nh": [], "mixture": [], "mixture_celsius": [], "qvisc": [], "qfric": []}
    for T in T_GRID:
        for e in E_GRID:
            out["sw"].append([T, e] + [hexd(v) for v in sw(T, e)])
            out["fks"].append([T, e] + [hexd(v) for v in fks(FKS, T, e)])
            out["fks2"].append([T, e] + [hexd(v) for v in fks(FKS2, T, e)])
    for sr in E_GRID:
        out["nh"].append([sr, hexd(nh(sr))])
    nus = [(1e-6, 1.5e-5), (30.5, 1.5e-5), (1e4, 1.5e-5)]
    for a in A_GRID:
        for (n1, n2) in nus:
            for T in (250.0, 300.0, 700.0, 900.0):
                out["mixture"].append([a, n1, n2, T] + [hexd(v) for v in mixture(a, n1, n2, T, KCUB, 0.0)])
                out["mixture_celsius"].append([a, n1, n2, T] + [hexd(v) for v in mixture(a, n1, n2, T, KCUB_C, -273.0)])
    for mu in (1.8e-5, 123.4, 2.7e7):
        for e in E_GRID:
            out["qvisc"].append([mu, e, hexd(3 * f(0.9) * f(mu) * f(e) ** 2)])
    for model in ("constant", "exponential"):
        for a in (0.0, 0.3, 1.0):
            for r in (0.0, 1e-4, 5e-3, 2.5e-2):
                out["qfric"].append([model, a, 3.1e7, 1.013e5, r, hexd(qfric(model, a, 3.1e7, 1.013e5, r))])
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", sum(len(v) for v in out.values() if isinstance(v, list)), "vectors")


if __name__ == "__main__":
    main()
