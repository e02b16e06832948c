"""The oracle's two independent forms (field-at-a-time vs per-cell) must agree bit for bit, on every
mesh family and model; analytic gradient cases pin the Gauss-linear restatement."""
import math

import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P

from helpers import EPS_FACTOR, oracle_case, oracle_update

SHEAR = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)

CASES = {
    "cfg1_hex": lambda: cases.cfg1(dims=(14, 12, 10)),
    "cfg1_sheared_graded": lambda: cases.cfg1(dims=(12, 10, 8), affine=SHEAR, grading=(2.0, 0.5, 3.0)),
    "cfg2_fks_kelvin": lambda: cases.cfg2(dims=(12, 12, 8)),
    "cfg2_fks_celsius": lambda: cases.cfg2(dims=(10, 8, 8), units="Celsius"),
    "cfg3_poly": lambda: cases.cfg3(base=(10, 10, 10)),
    "cfg4_nortonhoff": lambda: cases.cfg4(dims=(10, 10, 10)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_faithful_equals_fused_bitwise(name):
    case = CASES[name]()
    a = oracle_update(case, O.FAITHFUL)
    b = oracle_update(case, O.FUSED)
    for f in case.outputs + [P.F_GRADU]:
        x, y = a.field(f), b.field(f)
        assert np.array_equal(x, y, equal_nan=True), P.FIELD_NAMES[f]
        assert np.isfinite(x).all(), P.FIELD_NAMES[f]
    # the configured inputs must exercise the intended ranges
    eps = a.internal(P.F_EPSDOT)
    assert eps.max() > 1.0 and eps.min() >= 0.0
    if case.params.stick.fricPatch >= 0:
        assert a.internal(P.F_QFRIC).max() > 0.0, "friction patch must touch material"


def _linear_case(affine, grading):
    m = M.hex_block(6, 5, 4, length=(1.2, 1.0, 0.8), affine=affine, grading=grading)
    G = np.array([[1.5, -2.0, 0.25], [0.5, 3.0, -1.0], [-0.75, 0.125, 2.0]])  # G[i][j] = dU_j/dx_i
    X = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    U = X @ G + np.array([0.1, -0.2, 0.3])
    return m, G, U


@pytest.mark.parametrize("affine,grading", [(M.IDENTITY, (1, 1, 1)), (M.IDENTITY, (3.0, 0.4, 2.0)), (SHEAR, (1, 1, 1))])
@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED])
def test_gradient_of_linear_field_is_exact(affine, grading, form):
    """Gauss linear reproduces a linear field's gradient whenever the face centre lies on the line
    between the two cell centres with the linear weights (orthogonal / affine-mapped hex)."""
    m, G, U = _linear_case(affine, grading)
    oc = O.OracleCase(m)  # all patches fixedValue with the exact boundary values
    oc.set(P.F_U, U[: m.nCells], U[m.nCells:])
    oc.grad_u(form)
    g = oc.field(P.F_GRADU).reshape(-1, 3, 3)
    assert np.abs(g[: m.nCells] - G[None]).max() <= 2e-12
    if affine == M.IDENTITY:
        # boundary value g_P + n (x) (snGrad - n.g_P) is exact only where d = Cf - C_P is parallel to n:
        # OpenFOAM's boundary snGrad is the uncorrected deltaCoeffs*(U_b - U_P)
        assert np.abs(g[m.nCells:] - G[None]).max() <= 2e-12


def test_rigid_rotation_has_zero_strain():
    m = M.hex_block(8, 8, 6, length=(1.0, 1.0, 0.75))
    X = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    om = np.array([0.3, -0.2, 41.9])
    U = np.cross(np.broadcast_to(om, X.shape), X - np.array([0.5, 0.5, 0.0]))
    oc = O.OracleCase(m)
    oc.set(P.F_U, U[: m.nCells], U[m.nCells:])
    oc.strain_rate(EPS_FACTOR, P.F_EPSDOT)
    eps = oc.field(P.F_EPSDOT)
    assert eps.max() <= 1e-12 * np.linalg.norm(om)
    g = oc.field(P.F_GRADU).reshape(-1, 3, 3)
    # grad U of omega x r is the skew tensor: G[i][j] = dU_j/dx_i = -eps_ijk omega_k
    W = np.array([[0, om[2], -om[1]], [-om[2], 0, om[0]], [om[1], -om[0], 0]])
    assert np.abs(g - W[None]).max() <= 1e-11


def test_zero_gradient_patch_and_boundary_gradient():
    """boundary value of grad U: g_b = g_P + n (x) (snGrad - n.g_P); its normal component equals snGrad."""
    case = cases.cfg1(dims=(8, 8, 6), affine=SHEAR)
    oc = oracle_case(case)
    oc.grad_u(O.FAITHFUL)
    m = case.mesh
    g = oc.field(P.F_GRADU).reshape(-1, 3, 3)
    U = oc.field(P.F_U)
    nif = m.nInternalFaces
    for p in range(m.nPatches):
        sl = m.patch_slice_b(p)
        fc = m.face_cells(p)
        n = m.Sf[nif:][sl] / m.magSf[nif:][sl][:, None]
        gb = g[m.nCells:][sl]
        ng = np.einsum("fi,fij->fj", n, gb)
        if case.patch_u_bc[p] == P.BC_ZERO_GRADIENT:
            sn = np.zeros_like(ng)
        else:
            sn = m.deltaCoeffs[nif:][sl][:, None] * (U[m.nCells:][sl] - U[fc])
        scale = np.abs(g).max()
        assert np.abs(ng - sn).max() <= 1e-12 * scale


def test_stale_strain_rate_call_sequence():
    """thermo.correct() consumes the *stored* epsilonDotEq: nu1 changes only after
    TEqn.H:3 refreshed the field."""
    case = cases.cfg2(dims=(8, 8, 6))
    oc = oracle_case(case)
    n = case.mesh.nCells + case.mesh.nBoundaryFaces
    oc.set(P.F_EPSDOT, np.full(case.mesh.nCells, 1.0), np.full(case.mesh.nBoundaryFaces, 1.0))
    oc.viscosity_correct(1, case.params.visc1)
    nu_stale = oc.internal(P.F_NU1)
    oc.strain_rate(EPS_FACTOR, P.F_EPSDOT)
    oc.viscosity_correct(1, case.params.visc1)
    assert not np.array_equal(nu_stale, oc.internal(P.F_NU1))
    assert n == oc.nTot


def test_norton_hoff_has_no_sigmay():
    case = cases.cfg4(dims=(6, 6, 6))
    oc = oracle_update(case)
    stick = P.constant_slip_stick(case.mesh.patch_id("zmax"), 0.9, 0.5, cases.OMEGA, 0.3)
    with pytest.raises(O.OracleError) as e:
        oc.friction_correct(stick)
    assert e.value.status == P.ERR_MISSING_FIELD


def stress_strain_inputs(case, oc):
    n, nb = case.mesh.nCells, case.mesh.nBoundaryFaces
    rng = np.random.default_rng(5)
    prgh = 1e5 * (1.0 + 0.05 * rng.standard_normal(n + nb))
    oc.set(P.F_PRGH, prgh[:n], prgh[n:])
    v = case.params.visc1
    return P.StressStrainParams(Q=v.Q, R=v.R, sigmaR=v.sigmaR, A=v.A, n=v.n, deltaT=1e-3), prgh


def test_stress_strain_forms_agree_and_match_numpy():
    """fluid/calculateStrain.H: field-at-a-time == per-cell bitwise; values against a numpy evaluation"""
    case = CASES["cfg1_sheared_graded"]()
    res = []
    for form in (O.FAITHFUL, O.FUSED):
        oc = oracle_update(case, form)
        prm, prgh = stress_strain_inputs(case, oc)
        oc.stress_strain(prm, form)
        oc.stress_strain(prm, form)  # second step accumulates epsilonEq
        res.append(oc)
    for f in (P.F_Z, P.F_SIGMAF_PP, P.F_DEPSEQ, P.F_EPSEQ, P.F_SIGMAVM, P.F_GRADU):
        assert np.array_equal(res[0].field(f), res[1].field(f)), P.FIELD_NAMES[f]
    oc = res[0]
    assert np.array_equal(oc.field(P.F_EPSEQ), 2.0 * oc.field(P.F_DEPSEQ))
    G = oc.field(P.F_GRADU).reshape(-1, 3, 3)
    dt = 1e-3
    sym = 0.5 * (G + G.transpose(0, 2, 1))
    dE = sym * dt
    I3 = np.eye(3)[None]
    dD = dE - np.trace(dE, axis1=1, axis2=2)[:, None, None] / 3.0 * I3
    J2s = 0.5 * (np.trace(dD, axis1=1, axis2=2) ** 2 - np.einsum("nij,nij->n", dD, dD))
    dEq = np.sqrt(4.0 / 3.0 * np.abs(J2s))
    assert np.allclose(oc.field(P.F_DEPSEQ), dEq, rtol=1e-10, atol=1e-300)
    mu, pr = oc.field(P.F_MUEFF), prgh
    dev = sym - np.trace(sym, axis1=1, axis2=2)[:, None, None] / 3.0 * I3
    st = pr[:, None, None] * I3 + 2.0 * mu[:, None, None] * dev
    sd = st - np.trace(st, axis1=1, axis2=2)[:, None, None] / 3.0 * I3
    svm = np.sqrt(2.3 / 3.0) * np.sqrt(np.einsum("nij,nij->n", sd, sd))
    assert np.allclose(oc.field(P.F_SIGMAVM), svm, rtol=1e-6, atol=1e-3)  # p*I cancels: conditioning
    # the solver's own sigmaf equals the SheppardWright model's field (same formula, same inputs)
    assert np.array_equal(oc.field(P.F_SIGMAF_PP), oc.field(P.F_SIGMAF))
