"""Shared by tests/test_reference_sources.py and tests/golden/gen_reference_golden.py: runs one cases.Case
through the reference's own model sources (oracle/pyref.py) in the solver's call order and returns every field.

Inputs the reference's library code does not produce itself come from the oracle: mag(symm(fvc::grad(U)))
(OpenFOAM's gaussGrad, restated in oracle.cpp) is computed once by the oracle and handed to both sides."""
import os
import tempfile

import numpy as np

from oracle import pyoracle as O
from oracle import pyref as R
from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import params as P

EPS_FACTOR = float(np.sqrt(2.0 / 3.0))
VOL_OUTPUTS = {"nu1": P.F_NU1, "nu2": P.F_NU2, "sigmaf": P.F_SIGMAF, "sigmay": P.F_SIGMAY, "nu": P.F_NU, "mu": P.F_MUEFF,
               "rho": P.F_RHO, "cp": P.F_CP, "k": P.F_K, "qvisc": P.F_QVISC, "qfric": P.F_QFRIC}


def run_reference(case, tmpdir=None, gpu_shim=False):
    """-> (dict name -> array, Info text).  The call order of createFluidFields.H + one time step.
    gpu_shim: the same reference mixture / friction sources, but with the GPU-backed viscosity models of
    integration/openfoam selected by the reference's own viscosityModel::New (oracle/Makefile target `shim`)."""
    d = tmpdir or tempfile.mkdtemp(prefix="ssam_ref_")
    cases.write_case_dir(d, case, eps0=1.0, mu0=1.0)
    m = case.mesh
    n = m.nCells + m.nBoundaryFaces
    nh = case.params.visc1.model == P.VISC_NORTON_HOFF
    # grad U pieces from the oracle (identical on both sides)
    oc = O.OracleCase(m, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    if case.rotslip is not None:
        oc.bc_rotational_slip(case.rotslip)
    oc.strain_rate(1.0, P.F_STRAINRATE)
    mag_symm = oc.field(P.F_STRAINRATE).copy()
    U_all = oc.field(P.F_U).copy()
    oc.close()

    rc = R.ReferenceCase(m, d, gpu=gpu_shim)
    if gpu_shim:
        rc.set_U(U_all, case.patch_u_bc)  # NortonHoff takes its strain rate from U on the device
    cat = np.concatenate
    rc.set("temp", cat([case.T_i, case.T_b]))
    rc.set("epsilonDotEq", np.full(n, 1.0))            # createFluidFields.H initial value (here: 1)
    rc.set("muEff", np.full(n, 1.0))
    rc.set("p", cat([case.p_i, case.p_b]))
    rc.set("__magSymmGradU", mag_symm)
    rc.set("alpha.metal", cat([case.alpha1_i, case.alpha1_b]), as_file=True)
    rc.set("qfric", np.zeros(n), as_file=True)
    has_fric = not nh and case.params.stick.fricPatch >= 0
    rc.construct(with_friction=has_fric)
    out = {}
    rc.thermo_correct()                                 # alphaEqn.H:219 / solveFluid.H:52
    rc.set("epsilonDotEq", EPS_FACTOR * mag_symm)       # fluid/TEqn.H:3
    rc.thermo_correct()
    mu = rc.eval("mu")                                  # TEqn.H:6  muEff = thermo.mu()
    rc.set("muEff", mu)
    if has_fric:
        rc.friction_correct()                           # TEqn.H:9
    for name in VOL_OUTPUTS:
        if name in ("qvisc", "qfric") and not has_fric:
            continue
        if name in ("sigmaf", "sigmay") and nh:
            continue
        out[name] = rc.eval(name)
    out["muf"], out["nuf"] = rc.eval("muf"), rc.eval("nuf")
    info = rc.info()
    rc.close()
    return out, info


def run_oracle(case, form=O.FAITHFUL):
    """the same sequence on oracle/oracle.cpp"""
    m = case.mesh
    n, nb = m.nCells, m.nBoundaryFaces
    p = case.params
    nh = p.visc1.model == P.VISC_NORTON_HOFF
    oc = O.OracleCase(m, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    oc.set(P.F_T, case.T_i, case.T_b)
    oc.set(P.F_ALPHA1, case.alpha1_i, case.alpha1_b)
    oc.set(P.F_P, case.p_i, case.p_b)
    if case.rotslip is not None:
        oc.bc_rotational_slip(case.rotslip)
    oc.set(P.F_EPSDOT, np.full(n, 1.0), np.full(nb, 1.0))

    def thermo_correct():
        oc.viscosity_correct(1, p.visc1, form)
        oc.viscosity_correct(2, p.visc2, form)
        oc.mixture("calc_nu", p.mix, form)

    thermo_correct()
    thermo_correct()
    oc.strain_rate(EPS_FACTOR, P.F_EPSDOT, form)
    thermo_correct()
    oc.mixture("mu", p.mix, form)
    has_fric = not nh and p.stick.fricPatch >= 0
    if has_fric:
        oc.friction_correct(p.stick, form=form)
    for w in ("rho", "cp", "k"):
        oc.mixture(w, p.mix, form)
    out = {}
    for name, f in VOL_OUTPUTS.items():
        if name in ("qvisc", "qfric") and not has_fric:
            continue
        if name in ("sigmaf", "sigmay") and nh:
            continue
        out[name] = oc.field(f).copy()
    out["muf"], out["nuf"] = oc.mixture_face_visc(p.mix, False), oc.mixture_face_visc(p.mix, True)
    oc.close()
    return out
