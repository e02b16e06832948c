"""The oracle against outputs frozen from the reference's own model sources (tests/golden/reference_sources.npz,
written by tests/golden/gen_reference_golden.py where /root/reference is mounted).  Needs neither the reference
tree nor oracle/_ref, so it also runs on the GPU box; bit-exact, both oracle forms."""
import os
import sys

import numpy as np
import pytest

from oracle import pyoracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from gen_reference_golden import GOLDEN_CASES, inputs_digest  # noqa: E402

from ref_sequence import run_oracle  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sources.npz"))


@pytest.mark.parametrize("form", [O.FAITHFUL, O.FUSED], ids=["faithful", "fused"])
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_oracle_reproduces_frozen_reference_outputs(name, form):
    case = GOLDEN_CASES[name]()
    assert inputs_digest(case) == bytes(GOLDEN[name + "/inputs_sha256"]).decode(), \
        "the synthetic case changed: regenerate tests/golden/reference_sources.npz"
    orc = run_oracle(case, form)
    if case.rotslip is not None:
        from solidstateamfoam_b200 import params as P

        oc = O.OracleCase(case.mesh, case.patch_u_bc)
        oc.set(P.F_U, case.U_i, case.U_b)
        oc.bc_rotational_slip(case.rotslip)
        orc["bc_Ub"] = oc.boundary(P.F_U)[case.mesh.patch_slice_b(case.mesh.patch_id("zmax"))].reshape(-1)
    keys = [k.split("/", 1)[1] for k in GOLDEN.files if k.startswith(name + "/") and not k.endswith("inputs_sha256")]
    assert sorted(keys) == sorted(orc)
    for k in keys:
        ref = GOLDEN[name + "/" + k]
        same = (ref == orc[k]) | (np.isnan(ref) & np.isnan(orc[k]))
        assert same.all(), f"{name}/{k}: {np.count_nonzero(~same)} values differ from the frozen reference output"
