"""The oracle against the reference's OWN model sources.

oracle/_ref/libssam_ref.so is the reference's viscoPlasticModels, frictionModels and
incompressibleTwoPhaseThermalMixture translation units, compiled unmodified from /root/reference against the
stand-in OpenFOAM headers of oracle/ofstub/ (oracle/Makefile `ref`).  Both sides are CPU libm code evaluating
the same expression trees on the same inputs, so every field must agree BIT FOR BIT -- any transcription error
in oracle/oracle.cpp (operator order, max/min argument order, a field where a scalar belongs) shows up here.
Skipped where neither the reference tree nor a prebuilt library exists; tests/test_reference_golden.py then
still checks the oracle against the outputs frozen from this library.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from oracle import pyref as R
from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import params as P

from ref_sequence import run_oracle, run_reference

pytestmark = pytest.mark.skipif(not R.available(), reason="no reference tree and no prebuilt oracle/_ref")

SHEAR = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)
CASES = {
    "cfg1_sheppardwright_constantslip": lambda: cases.cfg1(dims=(9, 7, 6), affine=SHEAR, grading=(2.0, 0.5, 3.0)),
    "cfg2_fks_cubic_kelvin": lambda: cases.cfg2(dims=(8, 7, 5)),
    "cfg2_fks_cubic_celsius": lambda: cases.cfg2(dims=(6, 6, 4), units="Celsius"),
    "cfg3_poly_exponentialslip": lambda: cases.cfg3(base=(6, 6, 6)),
    "cfg4_nortonhoff": lambda: cases.cfg4(dims=(6, 5, 5)),
}
# other corners of the parameter space (FKS with b2, c2 != 1, another cubic k(T), other alloys / exponents)
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from gen_reference_golden import GOLDEN_CASES  # noqa: E402

for _k in ("cfg2_fks_general_exponents", "cfg1_sheppardwright_other_alloy", "cfg4_nortonhoff_other_exponent"):
    CASES[_k] = GOLDEN_CASES[_k]


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED], ids=["faithful", "fused"])
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_equals_reference_sources_bit_for_bit(name, form, tmp_path):
    case = CASES[name]()
    ref, info = run_reference(case, str(tmp_path))
    orc = run_oracle(case, form)
    assert set(ref) == set(orc)
    assert len(ref) >= 9
    for k in ref:
        same = (ref[k] == orc[k]) | (np.isnan(ref[k]) & np.isnan(orc[k]))
        assert same.all(), f"{name}/{k}: {np.count_nonzero(~same)} of {same.size} values differ from the reference sources"
    # the reference's own run-time selection and Info output
    assert "Selecting incompressible transport model" in info
    if "qfric" in ref:
        assert "Selecting stick model" in info and "Center for patch zmax found to be at" in info
        assert np.count_nonzero(ref["qfric"]) > 0


@pytest.mark.parametrize("name", ["cfg1_sheppardwright_constantslip", "cfg3_poly_exponentialslip"])
def test_rotational_slip_bc_equals_reference_sources(name, tmp_path):
    """boundaryConditions/*RotationalSlip*::updateCoeffs() against the oracle's restatement: bit-exact U_b"""
    case = CASES[name]()
    d = str(tmp_path)
    cases.write_case_dir(d, case)
    rc = R.ReferenceCase(case.mesh, d)
    ub_ref = rc.bc_rotational_slip("zmax")
    oc = O.OracleCase(case.mesh, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    oc.bc_rotational_slip(case.rotslip)
    sl = case.mesh.patch_slice_b(case.mesh.patch_id("zmax"))
    ub = oc.boundary(P.F_U)[sl]
    assert ub_ref.shape == ub.shape and np.abs(ub_ref).max() > 0
    assert np.array_equal(ub_ref, ub)


def _flatten_coeffs(path, name):
    """calculateStrain.H looks Q, R, sigmaR, A, n up in the phase dictionary itself (:6-10): move the entries of
    the optional <model>Coeffs sub-dictionary one level up, which the models accept too (optionalSubDict)."""
    txt = open(path).read()
    i = txt.index(name)
    j = txt.index("{", i)
    depth, k = 1, j + 1
    while depth:
        depth += {"{": 1, "}": -1}.get(txt[k], 0)
        k += 1
    open(path, "w").write(txt[:i] + txt[j + 1:k - 1] + txt[k:])


def test_calculate_strain_fragment_equals_oracle(tmp_path):
    """fluid/calculateStrain.H included verbatim (transport solve stubbed out) against oracle stress_strain: bit-exact"""
    case = CASES["cfg1_sheppardwright_constantslip"]()
    d = str(tmp_path)
    cases.write_case_dir(d, case)
    _flatten_coeffs(os.path.join(d, "constant", "transportProperties"), "SheppardWrightCoeffs")
    m = case.mesh
    n, nb = m.nCells, m.nBoundaryFaces
    rng = np.random.default_rng(9)
    prgh = 1e5 * (1.0 + 0.05 * rng.standard_normal(n + nb))
    mu = 10.0 ** rng.uniform(2, 7, n + nb)
    oc = O.OracleCase(m, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    oc.bc_rotational_slip(case.rotslip)
    oc.set(P.F_T, case.T_i, case.T_b)
    oc.strain_rate(float(np.sqrt(2.0 / 3.0)), P.F_EPSDOT)
    oc.set(P.F_MUEFF, mu[:n], mu[n:])
    oc.set(P.F_PRGH, prgh[:n], prgh[n:])
    v = case.params.visc1
    prm = P.StressStrainParams(Q=v.Q, R=v.R, sigmaR=v.sigmaR, A=v.A, n=v.n, deltaT=2.5e-4)
    rc = R.ReferenceCase(m, d)
    cat = np.concatenate
    rc.set("temp", cat([case.T_i, case.T_b]))
    rc.set("epsilonDotEq", oc.field(P.F_EPSDOT).copy())
    rc.set("muEff", mu)
    rc.set("p", cat([case.p_i, case.p_b]))
    rc.set("__magSymmGradU", np.ones(n + nb))
    rc.set("alpha.metal", cat([case.alpha1_i, case.alpha1_b]), as_file=True)
    rc.set("qfric", np.zeros(n + nb), as_file=True)
    rc.construct(with_friction=False)
    eq = np.zeros(n + nb)
    for step in range(2):  # the second pass accumulates epsilonEq
        oc.stress_strain(prm)
        ref = rc.calculate_strain(prm.deltaT, oc.field(P.F_GRADU), prgh, eq)
        eq = ref["epsilonEq"]
        for key, f in (("Z", P.F_Z), ("sigmaf", P.F_SIGMAF_PP), ("epsilonEq", P.F_EPSEQ), ("sigmavm", P.F_SIGMAVM)):
            assert np.array_equal(ref[key], oc.field(f)), (step, key)
        if step == 0:
            assert np.array_equal(ref["epsilonEq"], oc.field(P.F_DEPSEQ))  # 0 + dEpsilonEq
    assert "Max. von Mises stress" in rc.info()


def test_reference_extreme_inputs(tmp_path):
    """limiters, alpha outside [0,1], temperatures outside [Ta,Tm] / [Tmin,Tmax], tiny and huge strain rates"""
    case = cases.cfg2(dims=(6, 5, 4))
    rng = np.random.default_rng(5)
    n, nb = case.mesh.nCells, case.mesh.nBoundaryFaces
    case.T_i = rng.choice([1.0, 150.0, 293.0, 500.0, 855.0, 2000.0], n)
    case.T_b = rng.choice([1.0, 293.0, 700.0, 3000.0], nb)
    case.alpha1_i = rng.choice([-0.3, 0.0, 1e-12, 0.5, 1.0, 1.7], n)
    case.alpha1_b = rng.choice([-0.3, 0.0, 0.5, 1.0, 1.7], nb)
    case.U_i = case.U_i * rng.choice([0.0, 1e-12, 1.0, 1e6], n)[:, None]
    ref, _ = run_reference(case, str(tmp_path))
    orc = run_oracle(case)
    for k in ref:
        same = (ref[k] == orc[k]) | (np.isnan(ref[k]) & np.isnan(orc[k]))
        assert same.all(), k


def _break(tmp_path, case, fname, old, new):
    d = str(tmp_path)
    cases.write_case_dir(d, case)
    p = os.path.join(d, "constant", fname)
    txt = open(p).read()
    assert old in txt
    open(p, "w").write(txt.replace(old, new))
    m = case.mesh
    n = m.nCells + m.nBoundaryFaces
    rc = R.ReferenceCase(m, d)
    rc.set("temp", np.full(n, 500.0))
    rc.set("epsilonDotEq", np.ones(n))
    rc.set("muEff", np.ones(n))
    rc.set("p", np.ones(n))
    rc.set("__magSymmGradU", np.ones(n))
    rc.set("alpha.metal", np.ones(n), as_file=True)
    rc.set("qfric", np.zeros(n), as_file=True)
    return rc


def test_reference_error_messages_match_the_host_mirror(tmp_path):
    """the messages solidstateamfoam_b200/host reproduces are the reference's own (stickModelNew.C, ...Mixture.C)"""
    case = cases.cfg2(dims=(3, 3, 3))
    rc = _break(tmp_path / "a", case, "frictionProperties", "stickModel constantSlipStick;", "stickModel stickySlip;")
    with pytest.raises(R.ReferenceError_) as e:
        rc.construct()
    assert "Unknown stickModel type stickySlip" in str(e.value)
    assert "constantSlipStick exponentialSlipStick" in str(e.value)
    rc = _break(tmp_path / "b", case, "transportProperties", "units Kelvin;", "units Rankine;")
    rc.construct()
    with pytest.raises(R.ReferenceError_) as e:
        rc.eval("k")
    assert "Unknown conductivity function units Rankine for phase 1" in str(e.value)
    rc = _break(tmp_path / "c", case, "transportProperties", "        a1 ", "        a1missing ")
    with pytest.raises(R.ReferenceError_) as e:
        rc.construct()
    assert "keyword a1 is undefined in dictionary" in str(e.value)
