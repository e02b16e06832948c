"""Pins the oracle (the reference ships no golden vectors): every closed-form expression against
60-digit mpmath known answers (tests/golden/known_answers.json, made by tests/golden/gen_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "known_answers.json")))

# double evaluation of exp(Q/(R*T)) with an argument near 58 carries ~1e-14 of argument rounding;
# everything else is a few ulp
TOL = 5e-14


def fromhex(s):
    return float.fromhex(s)


def line_mesh(n):
    """n cells in a row: only used as a container for n independent pointwise evaluations"""
    return M.hex_block(n, 1, 1, length=(float(n), 1.0, 1.0))


def pointwise_case(n):
    m = line_mesh(n)
    return m, O.OracleCase(m)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / np.abs(b)
    return np.where(a == b, 0.0, r)


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED])
@pytest.mark.parametrize("key,params", [("sw", cases.sheppard_wright_aa6061()), ("fks", cases.fks_placeholder()),
                                        ("fks2", P.fks(a1=3.24e8, a2=1.14e8, b1=0.05, b2=0.7, b3=1.0, e0=0.01, c1=3.0,
                                                       c2=2.5, Tm=855.0, Ta=293.0, rho=2700.0, nuMin=1e-6, nuMax=1e4,
                                                       eDotMin=1e-6))])
def test_viscosity_models(key, params, form):
    rows = GOLD[key]
    n = len(rows)
    m, oc = pointwise_case(n)
    T = np.array([r[0] for r in rows])
    e = np.array([r[1] for r in rows])
    nb = m.nBoundaryFaces
    oc.set(P.F_T, T, np.full(nb, 300.0))
    oc.set(P.F_EPSDOT, e, np.full(nb, 1.0))
    oc.viscosity_correct(1, params, form)
    for col, f in ((2, P.F_NU1), (3, P.F_SIGMAF), (4, P.F_SIGMAY)):
        ref = np.array([fromhex(r[col]) for r in rows])
        got = oc.internal(f)
        r = rel(got, ref)
        assert r.max() <= TOL, (key, P.FIELD_NAMES[f], float(r.max()), rows[int(np.argmax(r))][:2])


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED])
def test_norton_hoff_formula(form):
    """nu(sr) through a uniform-shear velocity field whose strainRate() is known exactly."""
    rows = GOLD["nh"]
    p = cases.norton_hoff_placeholder()
    for sr, ref_hex in rows:
        # U = (g*y, 0, 0): grad U has one entry g, symm = g/2 twice, mag = g/sqrt(2), strainRate = g
        m = M.hex_block(3, 3, 3)
        oc = O.OracleCase(m)
        X = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
        U = np.zeros_like(X)
        U[:, 0] = sr * X[:, 1]
        oc.set(P.F_U, U[: m.nCells], U[m.nCells:])
        oc.viscosity_correct(1, p, form)
        got = oc.internal(P.F_NU1)[13]  # centre cell
        srv = oc.internal(P.F_STRAINRATE)[13]
        assert abs(srv - sr) <= 4e-15 * max(sr, 1e-300) + 1e-30
        ref = fromhex(ref_hex)
        # sr itself carries ~1e-15 of gradient rounding, amplified by |m-1| = 0.8
        assert abs(got - ref) <= 2e-14 * abs(ref), (sr, got, ref)


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED])
@pytest.mark.parametrize("key,units", [("mixture", "Kelvin"), ("mixture_celsius", "Celsius")])
def test_mixture(key, units, form):
    rows = GOLD[key]
    n = len(rows)
    m, oc = pointwise_case(n)
    nb = m.nBoundaryFaces
    a = np.array([r[0] for r in rows])
    oc.set(P.F_ALPHA1, a, np.zeros(nb))
    oc.set(P.F_NU1, np.array([r[1] for r in rows]), np.ones(nb))
    oc.set(P.F_NU2, np.array([r[2] for r in rows]), np.ones(nb))
    oc.set(P.F_T, np.array([r[3] for r in rows]), np.full(nb, 300.0))
    k1 = P.k_cubic(100.0, 0.2, -1e-4, 1e-8, 293.0 if units == "Kelvin" else 20.0,
                   855.0 if units == "Kelvin" else 582.0, units)
    mix = P.mixture(2700.0, 1.2, 896.0, 1005.0, k1, P.k_constant(0.026))
    for what in ("mu", "calc_nu", "rho", "cp", "k"):
        oc.mixture(what, mix, form)
    for col, f in ((4, P.F_MUEFF), (5, P.F_NU), (6, P.F_RHO), (7, P.F_CP), (8, P.F_K)):
        ref = np.array([fromhex(r[col]) for r in rows])
        r = rel(oc.internal(f), ref)
        assert r.max() <= 1e-15 * 8, (key, P.FIELD_NAMES[f], float(r.max()))


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED])
def test_qvisc(form):
    rows = GOLD["qvisc"]
    m, oc = pointwise_case(len(rows))
    nb = m.nBoundaryFaces
    oc.set(P.F_MUEFF, np.array([r[0] for r in rows]), np.ones(nb))
    oc.set(P.F_EPSDOT, np.array([r[1] for r in rows]), np.ones(nb))
    oc.friction_correct(P.no_friction_patch(0.9), form=form)
    ref = np.array([fromhex(r[2]) for r in rows])
    assert rel(oc.internal(P.F_QVISC), ref).max() <= 8e-16


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED])
def test_qfric(form):
    """qfric at the cell under a one-face friction patch, radius r set through the cell-centre array."""
    for model, alpha, sy, pr, r, ref_hex in GOLD["qfric"]:
        m = M.hex_block(1, 1, 1)
        top = m.patch_id("zmax")
        # centre of the patch is Cf of its single face; put the cell centre at distance r below it
        m.C[0] = m.Cf[m.patchStart[top]] - np.array([0.0, 0.0, r])
        oc = O.OracleCase(m)
        nb = m.nBoundaryFaces
        oc.set(P.F_ALPHA1, [alpha], np.zeros(nb))
        oc.set(P.F_SIGMAY, [sy], np.zeros(nb))
        oc.set(P.F_P, [pr], np.zeros(nb))
        oc.set(P.F_MUEFF, [1.0], np.zeros(nb))
        oc.set(P.F_EPSDOT, [1.0], np.zeros(nb))
        if model == "constant":
            sp = P.constant_slip_stick(top, 0.9, 0.5, (0.0, 0.0, 41.9), 0.3)
        else:
            sp = P.exponential_slip_stick(top, 0.9, (0.0, 0.0, 41.9), 41.9, 0.3, 19e-3, 0.4, 0.5)
        oc.friction_correct(sp, form=form)
        got = oc.internal(P.F_QFRIC)[0]
        ref = fromhex(ref_hex)
        rr = np.sqrt(np.sum((m.C[0] - m.Cf[m.patchStart[top]]) ** 2))
        # r is recomputed from coordinates: allow its rounding (|r| ~ 1e-2 from coordinates ~ 1)
        tol = 1e-12 if r > 0 else 0.0
        assert abs(got - ref) <= tol * abs(ref) + (0 if r > 0 else 0), (model, alpha, r, got, ref, rr)


def test_edge_semantics():
    """Foam::max/min argument order with NaN, T/T at 0, Celsius clamp at exactly 0 (SURVEY D-4)."""
    m, oc = pointwise_case(4)
    nb = m.nBoundaryFaces
    oc.set(P.F_ALPHA1, [0.5, 0.5, np.nan, 2.0], np.zeros(nb))
    oc.set(P.F_T, [273.0, 0.0, 300.0, 300.0], np.full(nb, 300.0))
    k1 = P.k_cubic(100.0, 0.2, -1e-4, 1e-8, 0.0, 582.0, "Celsius")
    mix = P.mixture(2700.0, 1.2, 896.0, 1005.0, k1, P.k_constant(0.026))
    for form in (O.FAITHFUL, O.FUSED):
        oc.mixture("k", mix, form)
        k = oc.internal(P.F_K)
        assert np.isnan(k[0])          # clamp(273-273.0, 0, 582) = 0 -> k0*(0/0)
        assert np.isnan(k[1])          # phase 2: k*(T/T) with T = 0
        oc.mixture("rho", mix, form)
        rho = oc.internal(P.F_RHO)
        # min(max(NaN,0),1): max(NaN,0) = (NaN>0)?NaN:0 = 0 -> la = 0 -> rho2
        assert rho[2] == 1.2 and rho[3] == 2700.0
    bad = P.mixture(1.0, 1.0, 1.0, 1.0, P.k_cubic(1, 1, 1, 1, 0, 1, units="Rankine"), P.k_constant(1.0))
    with pytest.raises(O.OracleError) as e:
        oc.mixture("k", bad)
    assert e.value.status == P.ERR_UNITS
