"""The C++ host-side mirror of the reference interfaces (solidstateamfoam_b200/host): dictionaries in,
RTS-selected models, the solver's call sequence replayed through the C-ABI, results vs the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import _build, cases
from solidstateamfoam_b200 import params as P

from helpers import EPS_FACTOR, BIT_EXACT, check_field, oracle_case

DRIVER = os.path.join(_build.LIB, "ssam_case_driver")


def run_driver(tmp_path, case, steps=2, expect_fail=False):
    _build.build_host()
    d = str(tmp_path / "case")
    cases.write_case_dir(d, case)
    out = str(tmp_path / "out.bin")
    r = subprocess.run([DRIVER, d, str(steps), out], capture_output=True, text=True, timeout=300)
    if expect_fail:
        return r
    assert r.returncode == 0, r.stdout + r.stderr
    n_tot = case.mesh.nCells + case.mesh.nBoundaryFaces
    res = {}
    with open(out, "rb") as fh:
        while True:
            h = np.fromfile(fh, dtype=np.int32, count=1)
            if h.size == 0:
                break
            res[int(h[0])] = np.fromfile(fh, dtype=np.float64, count=n_tot)
    return r, res


def oracle_sequence(case, steps=2, U=None):
    """the same call sequence on the oracle (case_driver.cpp main loop)"""
    oc = oracle_case(case)
    n, nb = case.mesh.nCells, case.mesh.nBoundaryFaces
    oc.set(P.F_EPSDOT, np.full(n, 1.0), np.full(nb, 1.0))
    oc.set(P.F_MUEFF, np.full(n, 1.0), np.full(nb, 1.0))
    p = case.params
    nh = p.visc1.model == P.VISC_NORTON_HOFF

    dn = O.interface_delta_n([oc])

    def thermo_correct():
        oc.viscosity_correct(1, p.visc1)
        oc.viscosity_correct(2, p.visc2)
        oc.mixture("calc_nu", p.mix)
        oc.interface_correct(dn)  # interfaceProperties::correct(), immiscible...H:80

    thermo_correct()  # constructors
    for _ in range(steps):
        thermo_correct()
        oc.solver_rho_rhocp(p.mix)
        thermo_correct()
        oc.strain_rate(EPS_FACTOR, P.F_EPSDOT)
        oc.mixture("mu", p.mix)
        if not nh:
            oc.friction_correct(p.stick)
        oc.mixture("cp", p.mix)
        oc.mixture("k", p.mix)
    return oc


def test_driver_fails_loudly_without_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_driver(tmp_path, cases.cfg1(dims=(4, 4, 4)), expect_fail=True)
    assert r.returncode == 1 and "FOAM FATAL ERROR" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg4"])
def test_driver_call_sequence_matches_oracle(tmp_path, name):
    case = {"cfg1": lambda: cases.cfg1(dims=(12, 10, 8)), "cfg2": lambda: cases.cfg2(dims=(10, 10, 8), units="Celsius"),
            "cfg4": lambda: cases.cfg4(dims=(8, 8, 8))}[name]()
    r, res = run_driver(tmp_path, case)
    assert "Selecting stick model" in r.stdout and "End" in r.stdout
    if name != "cfg4":
        assert "Center for patch zmax found to be at" in r.stdout
    else:
        assert "friction.correct() skipped" in r.stdout
    oc = oracle_sequence(case)
    assert res, "driver wrote no fields"
    lo, hi = O.field_min_max([oc], P.F_T)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("Min/max T:")][-1]
    assert tuple(float(v) for v in line.split(":")[1].split()) == (lo, hi)
    for f, got in res.items():
        check_field(P.FIELD_NAMES[f], got, oc.field(f), f in BIT_EXACT)


@pytest.mark.gpu
def test_driver_reference_error_behaviour(tmp_path):
    """unknown stickModel / viscosityModel / units abort with the reference's messages"""
    case = cases.cfg1(dims=(4, 4, 4))
    d = str(tmp_path / "case")
    r = run_driver(tmp_path, case, expect_fail=True)
    assert r.returncode == 0
    fp = os.path.join(d, "constant", "frictionProperties")
    txt = open(fp).read()
    open(fp, "w").write(txt.replace("stickModel constantSlipStick;", "stickModel stickySlip;"))
    r = subprocess.run([DRIVER, d, "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown stickModel type stickySlip" in r.stderr
    assert "constantSlipStick exponentialSlipStick" in r.stderr
    open(fp, "w").write(txt)
    tp = os.path.join(d, "constant", "transportProperties")
    ttxt = open(tp).read()
    open(tp, "w").write(ttxt.replace("transportModel SheppardWright;", "transportModel Carreau;"))
    r = subprocess.run([DRIVER, d, "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown viscosityModel type Carreau" in r.stderr
    open(tp, "w").write(ttxt.replace("kModel constant;\n    k 167.0;",
                                     "kModel cubic; k0 1; k1 1; k2 1; k3 1; Tmin 0; Tmax 1; units Rankine;"))
    r = subprocess.run([DRIVER, d, "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown conductivity function units Rankine for phase 1" in r.stderr
    open(tp, "w").write(ttxt.replace("        A ", "        Amissing "))
    r = subprocess.run([DRIVER, d, "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "keyword A is undefined in dictionary" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("binary", [False, True])
def test_driver_on_openfoam_case_directory(tmp_path, binary):
    """a real OpenFOAM case layout (polyMesh + 0/ fields + dictionaries) in, OpenFOAM fields out"""
    from solidstateamfoam_b200 import mesh as M

    shear = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)
    case = cases.cfg1(dims=(10, 8, 6), affine=shear)
    h = M.hex_block_handle(10, 8, 6, length=case.mesh.meta["length"], affine=shear)
    d = str(tmp_path / "foamCase")
    cases.write_foam_case(d, case, h, binary=binary)
    _build.build_host()
    r = subprocess.run([DRIVER, d, "2", str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # the oracle runs on the mesh as the driver saw it: geometry recomputed from points and faces
    rd = M.read_polymesh(os.path.join(d, "constant", "polyMesh"))
    import copy
    ing = copy.copy(case)
    ing.mesh = rd.mesh()
    oc = oracle_sequence(ing)
    for name, f in (("epsilonDotEq", P.F_EPSDOT), ("nu", P.F_NU), ("muEff", P.F_MUEFF), ("k", P.F_K),
                    ("qvisc", P.F_QVISC), ("qfric", P.F_QFRIC), ("sigmay", P.F_SIGMAY)):
        gi, gb, _ = rd.read_field(os.path.join(d, "2", name), 1)
        check_field(name, np.concatenate([gi, gb]), oc.field(f), f in BIT_EXACT)
