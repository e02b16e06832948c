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
