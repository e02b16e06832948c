"""The drop-in boundary, exercised with the reference's own code in the loop (GPU).

oracle/_ref/libssam_ref_gpu.so (oracle/Makefile target `shim`) is the reference's mixture, friction-model, stick-model
and boundary-condition sources, UNMODIFIED, compiled against the stand-in OpenFOAM headers, together with
integration/openfoam/gpuViscosityModels.C and gpuStickModels.C -- GPU-backed SheppardWright / FKS / NortonHoff and
constantSlipStick / exponentialSlipStick classes with the reference's names, TypeNames, constructor signatures and
dictionary keys, built INSTEAD of viscoPlasticModels/*.C and stickModels/{constant,exponential}SlipStick/*.C.  The
reference's own `viscosityModel::New` / `stickModel::New` select them from unchanged transportProperties /
frictionProperties, the reference's own `incompressibleTwoPhaseThermalMixture::calcNu()` and
`incompressibleFrictionModel::correct()` call them, and the GPU stick models look up the `sigmay` the GPU viscosity
model registered.  Every field is compared with the outputs frozen from the all-CPU reference sources
(tests/golden/reference_sources.npz): <= 1e-11 relative.
"""
import os
import sys

import numpy as np
import pytest

from oracle import pyref as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from gen_reference_golden import GOLDEN_CASES, inputs_digest  # noqa: E402

from helpers import check_field  # noqa: E402
from ref_sequence import run_reference  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sources.npz"))


def _shim_available():
    try:
        return R.build_shim()
    except RuntimeError:
        return False


@pytest.mark.skipif(not _shim_available(), reason="no reference tree and no prebuilt oracle/_ref/libssam_ref_gpu.so")
def test_shim_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback anywhere: without a CUDA device the GPU-backed model's constructor raises the library's error
    as a FatalError through the reference's own viscosityModel::New / mixture constructor."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from solidstateamfoam_b200 import cases

    case = GOLDEN_CASES["cfg2_fks_cubic_kelvin"]()
    d = str(tmp_path)
    cases.write_case_dir(d, case)
    n = case.mesh.nCells + case.mesh.nBoundaryFaces
    rc = R.ReferenceCase(case.mesh, d, gpu=True)
    rc.set("temp", np.full(n, 500.0))
    rc.set("epsilonDotEq", np.ones(n))
    rc.set("alpha.metal", np.ones(n), as_file=True)
    with pytest.raises(R.ReferenceError_) as e:
        rc.construct(with_friction=False)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


@pytest.mark.gpu
@pytest.mark.skipif(not _shim_available(), reason="no reference tree and no prebuilt oracle/_ref/libssam_ref_gpu.so")
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_reference_code_with_gpu_backed_models(name, tmp_path):
    case = GOLDEN_CASES[name]()
    assert inputs_digest(case) == bytes(GOLDEN[name + "/inputs_sha256"]).decode()
    got, info = run_reference(case, str(tmp_path), gpu_shim=True)
    assert "Selecting incompressible transport model" in info  # the reference's run-time selection picked them
    if "qfric" in got:
        assert "Selecting stick model" in info and "Center for patch zmax found to be at" in info
    keys = [k.split("/", 1)[1] for k in GOLDEN.files
            if k.startswith(name + "/") and not k.endswith("inputs_sha256") and not k.endswith("bc_Ub")]
    assert sorted(keys) == sorted(got)
    for k in keys:
        # nu2 (OpenFOAM's Newtonian, CPU on both sides), rho, cp are pure CPU reference code: identical
        check_field(f"{name}/{k}", got[k], GOLDEN[name + "/" + k], k in ("nu2", "rho", "cp"))


@pytest.mark.gpu
@pytest.mark.skipif(not _shim_available(), reason="no reference tree and no prebuilt oracle/_ref/libssam_ref_gpu.so")
@pytest.mark.parametrize("name", [k for k in GOLDEN_CASES if k + "/bc_Ub" in GOLDEN.files])
def test_gpu_backed_rotational_slip_patch_fields(name, tmp_path):
    """integration/openfoam's *RotationalSlipFvPatchVectorField (built instead of boundaryConditions/*.C): constructed
    from the patch's dictionary in 0/U, updateCoeffs() on the device, values mirrored into the patch field"""
    from solidstateamfoam_b200 import cases
    from solidstateamfoam_b200 import params as P

    case = GOLDEN_CASES[name]()
    d = str(tmp_path)
    cases.write_case_dir(d, case)
    rc = R.ReferenceCase(case.mesh, d, gpu=True)
    ub = rc.bc_rotational_slip("zmax").reshape(-1)
    check_field(f"{name}/U_b", ub, GOLDEN[name + "/bc_Ub"], case.rotslip.model == P.ROTSLIP_CONSTANT)
    assert np.abs(ub).max() > 0


# ---- the mixture's own translation unit on the GPU, fields device-resident (oracle/Makefile target `shim2`) ------------
def _shim2_available():
    try:
        return R.build_shim2()
    except RuntimeError:
        return False


@pytest.mark.gpu
@pytest.mark.skipif(not _shim2_available(), reason="no reference tree and no prebuilt oracle/_ref/libssam_ref_gpu2.so")
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_reference_code_with_gpu_backed_mixture(name, tmp_path):
    """integration/openfoam/gpuIncompressibleTwoPhaseThermalMixture.C built INSTEAD of the reference's
    incompressibleTwoPhaseThermalMixture.C against the reference's unmodified header: calcNu / mu / k / cp / rho / muf /
    nuf on the device, called by the same harness in the solver's order, against the all-CPU reference outputs."""
    case = GOLDEN_CASES[name]()
    got, info = run_reference(case, str(tmp_path), gpu_shim=2)
    assert "Selecting incompressible transport model" in info
    keys = [k.split("/", 1)[1] for k in GOLDEN.files
            if k.startswith(name + "/") and not k.endswith("inputs_sha256") and not k.endswith("bc_Ub")]
    assert sorted(keys) == sorted(got)
    for k in keys:
        # rho and cp are IEEE-only on both sides; nu2 is OpenFOAM's Newtonian (CPU) on both sides
        check_field(f"{name}/{k}", got[k], GOLDEN[name + "/" + k], k in ("nu2", "rho", "cp"))


def _teqn_case(tmp_path, name="cfg2_fks_cubic_kelvin", sub="r0"):
    from oracle import pyoracle as O
    from solidstateamfoam_b200 import cases
    from solidstateamfoam_b200 import params as P

    case = GOLDEN_CASES[name]()
    d = str(tmp_path / sub)
    os.makedirs(d, exist_ok=True)
    cases.write_case_dir(d, case, eps0=1.0, mu0=1.0)
    m = case.mesh
    n = m.nCells + m.nBoundaryFaces
    oc = O.OracleCase(m, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    if case.rotslip is not None:
        oc.bc_rotational_slip(case.rotslip)
    U_all = oc.field(P.F_U).copy()
    oc.close()
    rc = R.ReferenceCase(m, d, gpu=2)
    rc.set_U(U_all, case.patch_u_bc)
    cat = np.concatenate
    rc.set("temp", cat([case.T_i, case.T_b]))
    rc.set("epsilonDotEq", np.full(n, 1.0))
    rc.set("muEff", np.full(n, 1.0))
    rc.set("p", cat([case.p_i, case.p_b]))
    rc.set("alpha.metal", cat([case.alpha1_i, case.alpha1_b]), as_file=True)
    rc.set("qfric", np.zeros(n), as_file=True)
    rc.construct(with_friction=True)
    return case, rc, n


@pytest.mark.gpu
@pytest.mark.skipif(not _shim2_available(), reason="no reference tree and no prebuilt oracle/_ref/libssam_ref_gpu2.so")
def test_teqn_block_is_device_resident(tmp_path):
    """fluid/TEqn.H:3-15 through ssamTEqnUpdate.H, replayed for three time steps in the solver's order
    (thermo.correct() -> TEqn block), with the PCIe traffic of every step counted by the shim:
      * the first step uploads each input once (U, temp, alpha1, p, the Newtonian nu2 and the initial epsilonDotEq);
      * a later step uploads ONLY what the solver changed (here temp, as after the T solve) -- nothing else crosses;
      * every step downloads each consumed output exactly once (nu1, sigmaf, sigmay, nu | epsilonDotEq, muEff, qvisc,
        qfric, cpf, kf);
    and the fields equal the all-CPU reference sequence."""
    name = "cfg2_fks_cubic_kelvin"
    case, rc, n = _teqn_case(tmp_path, name)
    F = 8 * n  # bytes of one scalar field
    rc.gpu_traffic(reset=True)
    per_step = []
    T = np.concatenate([case.T_i, case.T_b])
    for step in range(3):
        if step:
            rc.set("temp", T + 0.5 * step)        # the solver solved TEqn: temp changed, U / alpha1 / p did not
        rc.thermo_correct()                        # solveFluid.H:52
        rc.teqn_update(0)                          # TEqn.H:3-15
        per_step.append(rc.gpu_traffic(reset=True))
    h2d0, d2h0 = per_step[0]
    assert h2d0 <= (3 + 1 + 1 + 1 + 1 + 1) * F, per_step      # U, temp, alpha1, p, nu2, the initial epsilonDotEq
    for h2d, d2h in per_step[1:]:
        assert h2d == F, per_step                              # temp only
    for h2d, d2h in per_step:
        assert d2h == 10 * F, per_step                         # ten consumed outputs, once each
    # values: the same three steps on the all-CPU reference build
    got = {k: rc.eval(k) for k in ("epsilonDotEq", "muEff", "cpf", "kf", "qvisc", "qfric", "nu", "nu1", "sigmay")}
    rc.close()

    # all-CPU: thermo.correct() with the stale epsilonDotEq, then the five TEqn lines, per step
    from oracle import pyoracle as O
    from solidstateamfoam_b200 import cases
    from solidstateamfoam_b200 import params as P

    d = str(tmp_path / "cpu")
    os.makedirs(d, exist_ok=True)
    cases.write_case_dir(d, case, eps0=1.0, mu0=1.0)
    oc = O.OracleCase(case.mesh, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    oc.bc_rotational_slip(case.rotslip)
    oc.strain_rate(1.0, P.F_STRAINRATE)
    mag_symm = oc.field(P.F_STRAINRATE).copy()
    oc.close()
    cpu = R.ReferenceCase(case.mesh, d, gpu=False)
    cat = np.concatenate
    cpu.set("temp", T)
    cpu.set("epsilonDotEq", np.full(n, 1.0))
    cpu.set("muEff", np.full(n, 1.0))
    cpu.set("p", cat([case.p_i, case.p_b]))
    cpu.set("__magSymmGradU", mag_symm)
    cpu.set("alpha.metal", cat([case.alpha1_i, case.alpha1_b]), as_file=True)
    cpu.set("qfric", np.zeros(n), as_file=True)
    cpu.construct(with_friction=True)
    for step in range(3):
        if step:
            cpu.set("temp", T + 0.5 * step)
        cpu.thermo_correct()
        cpu.set("epsilonDotEq", float(np.sqrt(2.0 / 3.0)) * mag_symm)
        cpu.set("muEff", cpu.eval("mu"))
        cpu.friction_correct()
    ref = {"epsilonDotEq": cpu.eval("epsilonDotEq"), "muEff": cpu.eval("muEff"), "cpf": cpu.eval("cp"), "kf": cpu.eval("k"),
           "qvisc": cpu.eval("qvisc"), "qfric": cpu.eval("qfric"), "nu": cpu.eval("nu"), "nu1": cpu.eval("nu1"),
           "sigmay": cpu.eval("sigmay")}
    cpu.close()
    for k in got:
        check_field(f"TEqn block {k}", got[k], ref[k], k in ("epsilonDotEq", "cpf"))


@pytest.mark.gpu
@pytest.mark.skipif(not _shim2_available(), reason="no reference tree and no prebuilt oracle/_ref/libssam_ref_gpu2.so")
def test_fused_teqn_block_and_two_regions(tmp_path):
    """(1) ssamFusedUpdate: ONE ssam_update_fused behind the reference's classes -- its TEqn outputs equal the staged
    block's once the viscosity models have seen the same epsilonDotEq; per step it uploads what changed and downloads
    five fields.  (2) two regions = two meshes = two GPU contexts alive side by side (chtMultiRegionInterFoam.C:114-129),
    interleaved, without disturbing each other."""
    a_case, a, na = _teqn_case(tmp_path, "cfg2_fks_cubic_kelvin", "fluidA")
    b_case, b, nb = _teqn_case(tmp_path, "cfg1_sheppardwright_constantslip", "fluidB")
    assert a.gpu_contexts() == 2
    # region A: staged reference of "the next thermo.correct() folded in": TEqn block, then thermo.correct(), then block
    a.thermo_correct()
    a.teqn_update(0)
    a.thermo_correct()       # now nu1 / sigmay / nu reflect the new epsilonDotEq
    a.teqn_update(0)
    staged = {k: a.eval(k) for k in ("epsilonDotEq", "muEff", "cpf", "kf", "qvisc", "nu1", "sigmay", "nu", "qfric")}
    # region B runs in between, on its own context
    b.thermo_correct()
    b.teqn_update(0)
    b_first = {k: b.eval(k) for k in ("epsilonDotEq", "muEff", "kf")}
    # region A again, fused this time
    a.gpu_traffic(reset=True)
    a.teqn_update(2)
    h2d, d2h = a.gpu_traffic(reset=True)
    assert h2d == 0, h2d                              # nothing changed on the host since the last call
    assert d2h == (5 + 4) * 8 * na, d2h               # five TEqn outputs + nu1, sigmaf, sigmay, qfric (allOutputs)
    fused = {k: a.eval(k) for k in ("epsilonDotEq", "muEff", "cpf", "kf", "qvisc", "nu1", "sigmay", "qfric")}
    for k in fused:
        check_field(f"fused vs staged {k}", fused[k], staged[k], k in ("epsilonDotEq", "cpf"))
    a.teqn_update(1)
    h2d, d2h = a.gpu_traffic(reset=True)
    assert (h2d, d2h) == (0, 5 * 8 * na)
    # region B was not disturbed
    b.teqn_update(0)
    for k, v in b_first.items():
        assert np.array_equal(b.eval(k), v, equal_nan=True), k
    a.close()
    assert b.gpu_contexts() == 1
    b.close()
